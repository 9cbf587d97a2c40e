// Video write-out, device side: un-normalise + clamp + *255 + uint8 + 'b c f h w -> f h w c' in ONE pass.
//
// Replaces the tensor arithmetic of utils/video_op.py:167-192 (save_i2vgen_video_safe; identical in
// save_t2vhigen_video_safe :263-288 and save_video_local :215-240), which the reference runs on the CPU after a
// 173 MB fp32 D2H copy ([1,3,16,704,1280]):
//     gen_video.mul_(std).add_(mean); gen_video.clamp_(0, 1); gen_video * 255.0; rearrange; .numpy().astype('uint8')
// Here the frames leave the device as RGB bytes (43 MB for the same video, 4x less PCIe traffic, no host arithmetic).
// Byte work, HBM-bound: 4 B read + 1 B written per element; a thread owns 4 consecutive pixels of a frame:
// three coalesced 16-byte plane loads, three 4-byte stores into 12 contiguous output bytes.
// Bit-exactness: separate fp32 multiply and add (the reference's mul_ / add_ are two roundings, no FMA), clamp,
// fp32 * 255, truncation toward zero (numpy float32 -> uint8 cast).
// The kernel also counts, per frame, the bytes in [117, 137] -- the "last frame is grey" anomaly test of :199-201.
#include "common.h"
#include "ptx.cuh"

namespace vg {

__device__ __forceinline__ unsigned to_u8(float x, float sd, float mn) {
  float v = __fadd_rn(__fmul_rn(x, sd), mn);
  v = fminf(fmaxf(v, 0.0f), 1.0f);
  return __float2uint_rz(__fmul_rn(v, 255.0f));
}

__global__ void __launch_bounds__(256) video_to_rgb8_kernel(const float* __restrict__ vid, uint8_t* __restrict__ out,
                                                            unsigned long long* __restrict__ band, int C, long F, long HW,
                                                            float m0, float m1, float m2, float s0, float s1, float s2) {
  // grid.y = frame; x over groups of 4 pixels
  const long f = blockIdx.y;
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long p0 = g * 4;
  unsigned cnt = 0;
  if (p0 < HW) {
    const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
    unsigned px[4][3];
    const bool full = (p0 + 4 <= HW) && ((HW & 3) == 0);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* src = vid + ((long)c * F + f) * HW + p0;     // plane c (channels >= C replicate the last one)
      if (c >= C) src = vid + ((long)(C - 1) * F + f) * HW + p0;
      float v[4];
      if (full) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(src));
        v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = (p0 + i < HW) ? src[i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) px[i][c] = to_u8(v[i], sd[c], mean[c]);
    }
    uint8_t* dst = out + (f * HW + p0) * 3;
    if (full) {
      // 12 bytes: R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
      uint32_t w0 = px[0][0] | (px[0][1] << 8) | (px[0][2] << 16) | (px[1][0] << 24);
      uint32_t w1 = px[1][1] | (px[1][2] << 8) | (px[2][0] << 16) | (px[2][1] << 24);
      uint32_t w2 = px[2][2] | (px[3][0] << 8) | (px[3][1] << 16) | (px[3][2] << 24);
      uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
      d32[0] = w0, d32[1] = w1, d32[2] = w2;
    } else {
      for (int i = 0; i < 4 && p0 + i < HW; ++i)
        for (int c = 0; c < 3; ++c) dst[i * 3 + c] = (uint8_t)px[i][c];
    }
    for (int i = 0; i < 4; ++i)
      if (p0 + i < HW)
        for (int c = 0; c < 3; ++c) cnt += (px[i][c] >= 117u && px[i][c] <= 137u) ? 1u : 0u;
  }
  if (band) {
    // warp-shuffle reduction, one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(band + f, (unsigned long long)cnt);
  }
}

}  // namespace vg

using namespace vg;

extern "C" int vgen_video_to_rgb8(const float* video, int64_t c, int64_t f, int64_t h, int64_t w, const float* mean3,
                                  const float* std3, uint8_t* out, unsigned long long* band_count, void* stream) {
  VG_REQUIRE(video && out && mean3 && std3, "vgen_video_to_rgb8: null pointer");
  VG_REQUIRE(c == 3 && f >= 0 && h > 0 && w > 0, "vgen_video_to_rgb8: video must be [3, f, h, w]");
  VG_REQUIRE((reinterpret_cast<uintptr_t>(video) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0,
             "vgen_video_to_rgb8: video must be 16-byte and out 4-byte aligned");
  VG_REQUIRE(f <= 65535, "vgen_video_to_rgb8: too many frames");
  if (f == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (band_count) VG_CUDA(cudaMemsetAsync(band_count, 0, sizeof(unsigned long long) * f, st));
  const long hw = h * w;
  dim3 grid((unsigned)cdiv(cdiv(hw, 4), 256), (unsigned)f);
  launch_kernel(video_to_rgb8_kernel, dim3(grid), dim3(256), 0, st, video, out, band_count, (int)c, f, hw, mean3[0], mean3[1], mean3[2], std3[0],
                                             std3[1], std3[2]);
  VG_LAUNCH_CHECK("video_to_rgb8_kernel");
  return 0;
}
