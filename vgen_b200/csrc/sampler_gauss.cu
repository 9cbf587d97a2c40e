// Per-step tensor arithmetic of GaussianDiffusion (tools/modules/diffusions/diffusion_gauss.py), the
// sampler pair of the SR600 pipeline.  All HBM-bound elementwise /
// reduction work on the latent (a few MB); the UNet forwards between these calls dominate the step.
//   vgen_cfg_combine : out = u + g*(y-u) (fp16, rounded like torch does per op) + per-sample sum / sum of
//                      squares of y and out for the std-ratio guidance rescale (:212-218)
//   vgen_gauss_x0    : rescale by guide_rescale*std(y)/std(out) + (1-guide_rescale), then the x0 prediction
//                      for prediction_type v | eps | x0 (:220-230)
//   vgen_lincomb_f32 : out = a0*x0 + a1*x1 + a2*x2 + a3*x3 in fp32 -- model-input scaling (:113), the
//                      DPM-Solver++(2M) SDE update (:124-139) and the DDIM inversion step (:408-410)
#include "common.h"
#include "ptx.cuh"

namespace vg {

__global__ void cfg_combine_kernel(const __half* __restrict__ y, const __half* __restrict__ u, __half* __restrict__ out,
                                   long n_per, float g, double* __restrict__ stats) {
  const long b = blockIdx.y;
  const __half* yb = y + b * n_per;
  const __half* ub = u + b * n_per;
  __half* ob = out + b * n_per;
  double sy = 0.0, syy = 0.0, so = 0.0, soo = 0.0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_per; i += (long)gridDim.x * blockDim.x) {
    const __half yv = yb[i], uv = ub[i];
    // u_out + guide_scale * (y_out - u_out): every intermediate is an fp16 tensor upstream
    const __half d = __float2half_rn(__half2float(yv) - __half2float(uv));
    const __half gd = __float2half_rn(g * __half2float(d));
    const __half o = __float2half_rn(__half2float(uv) + __half2float(gd));
    ob[i] = o;
    const float yf = __half2float(yv), of = __half2float(o);
    sy += yf, syy += (double)yf * yf, so += of, soo += (double)of * of;
  }
  __shared__ double red[4][32];
  double v[4] = {sy, syy, so, soo};
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], off);
    if (lane == 0) red[q][wid] = v[q];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double t = lane < (int)(blockDim.x >> 5) ? red[q][lane] : 0.0;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) t += __shfl_down_sync(0xffffffffu, t, off);
      if (lane == 0) atomicAdd(&stats[b * 4 + q], t);
    }
  }
}

// pred: 0 = x0, 1 = eps, 2 = v
__global__ void gauss_x0_kernel(const float* __restrict__ xt, const __half* __restrict__ out, const double* __restrict__ stats,
                                float guide_rescale, float alpha, float sigma, int pred, float* __restrict__ x0, long n_per,
                                long total) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  float o = __half2float(out[idx]);
  if (stats) {
    const long b = idx / n_per;
    const double n = (double)n_per;
    const double vy = (stats[b * 4 + 1] - stats[b * 4 + 0] * stats[b * 4 + 0] / n) / (n - 1.0);
    const double vo = (stats[b * 4 + 3] - stats[b * 4 + 2] * stats[b * 4 + 2] / n) / (n - 1.0);
    // .std() of an fp16 tensor is an fp16 value; the ratio, its scaling and the sum are fp16 ops too
    const __half sy = __float2half_rn((float)sqrt(vy > 0.0 ? vy : 0.0));
    const __half so = __float2half_rn((float)sqrt(vo > 0.0 ? vo : 0.0));
    const __half ratio = __float2half_rn(__half2float(sy) / __half2float(so));
    const __half gr = __float2half_rn(guide_rescale * __half2float(ratio));
    const __half fac = __float2half_rn(__half2float(gr) + (1.0f - guide_rescale));
    o = __half2float(__float2half_rn(o * __half2float(fac)));
  }
  const float x = xt[idx];
  float r;
  if (pred == 2) r = alpha * x - sigma * o;
  else if (pred == 1) r = (x - sigma * o) / alpha;
  else r = o;
  x0[idx] = r;
}

__global__ void lincomb_f32_kernel(float* __restrict__ out, long n, const float* __restrict__ x0, float a0,
                                   const float* __restrict__ x1, float a1, const float* __restrict__ x2, float a2,
                                   const float* __restrict__ x3, float a3) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  float r = a0 * x0[idx];
  if (x1) r += a1 * x1[idx];
  if (x2) r += a2 * x2[idx];
  if (x3) r += a3 * x3[idx];
  out[idx] = r;
}

}  // namespace vg

using namespace vg;

extern "C" {

int vgen_cfg_combine(const void* y, const void* u, void* out, int64_t batch, int64_t n_per, float guide_scale, double* stats,
                     void* stream) {
  VG_REQUIRE(y && u && out && stats && batch > 0 && batch <= 65535 && n_per > 1, "vgen_cfg_combine: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VG_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 4 * batch, st));
  long blocks = (n_per + 256 * 8 - 1) / (256 * 8);
  const long cap = (long)sm_count() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  dim3 grid((unsigned)blocks, (unsigned)batch);
  launch_kernel(cfg_combine_kernel, dim3(grid), dim3(256), 0, st, reinterpret_cast<const __half*>(y), reinterpret_cast<const __half*>(u),
                                           reinterpret_cast<__half*>(out), n_per, guide_scale, stats);
  VG_LAUNCH_CHECK("cfg_combine_kernel");
  return 0;
}

int vgen_gauss_x0(const float* xt, const void* out, const double* stats, float guide_rescale, float alpha, float sigma,
                  int pred_type, float* x0, int64_t batch, int64_t n_per, void* stream) {
  VG_REQUIRE(xt && out && x0 && batch > 0 && n_per > 1 && pred_type >= 0 && pred_type <= 2, "vgen_gauss_x0: bad arguments");
  const long total = batch * n_per;
  launch_kernel(gauss_x0_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      xt, reinterpret_cast<const __half*>(out), stats, guide_rescale, alpha, sigma, pred_type, x0, n_per, total);
  VG_LAUNCH_CHECK("gauss_x0_kernel");
  return 0;
}

int vgen_lincomb_f32(float* out, int64_t n, const float* x0, float a0, const float* x1, float a1, const float* x2, float a2,
                     const float* x3, float a3, void* stream) {
  VG_REQUIRE(out && x0 && n >= 0, "vgen_lincomb_f32: bad arguments");
  if (n == 0) return 0;
  launch_kernel(lincomb_f32_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), out, n, x0, a0, x1, a1,
                                                                                                   x2, a2, x3, a3);
  VG_LAUNCH_CHECK("lincomb_f32_kernel");
  return 0;
}

}  // extern "C"
