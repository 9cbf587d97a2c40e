// Flash attention (head_dim 64, no mask) on tcgen05: S = Q K^T and O = P V on the 5th-gen tensor
// cores with accumulators in TMEM, online softmax in registers (one thread per query row, so no
// shuffles), operands staged by TMA.
//
// Replaces xformers.ops.memory_efficient_attention as called by MemoryEfficientCrossAttention
// (/root/reference/tools/modules/unet/util.py:231-269): exact softmax(q k^T / sqrt(64)) v per
// (batch, head); fp16 inputs, fp32 scores / softmax / accumulation, P rounded to fp16 for the PV
// product (what the CUTLASS/FA2 kernels behind xformers do).
//
// One CTA = 256 query rows (two 128-row tiles) of one (batch, head), 10 warps:
//   warps 0-3 / 4-7  softmax + output for q-tile 0 / 1 (thread r <-> TMEM lane r)
//   warp 8           TMA producer: Q once, then a 2-stage ring of (K,V) blocks of 128 keys
//   warp 9           tcgen05.mma issuer (one lane)
// The two q-tiles ping-pong: while the softmax warps of tile 0 work on S0(j+1), the tensor core runs
// PV1(j) and QK1(j+1).  Per key block and tile: S (128x128 fp32) lives in TMEM, P (fp16) goes through
// shared memory in the 128B-swizzled K-major UMMA layout, P V is computed into a scratch TMEM buffer
// and folded into a register accumulator with the usual exp(m_old - m_new) rescale.
#include "common.h"
#include "ptx.cuh"

namespace vg {

static constexpr int kAttnThreads = 320;
static constexpr int kD = 64;
static constexpr int kTileQ = 128;   // rows per q-tile (UMMA M)
static constexpr int kTileK = 128;   // keys per block (UMMA N of QK^T)
static constexpr int kQBytes = kTileQ * kD * 2;   // 16 KB
static constexpr int kKBytes = kTileK * kD * 2;   // 16 KB
static constexpr int kPBytes = kTileQ * kTileK * 2;  // 32 KB per q-tile
static constexpr int kKvStages = 2;

struct alignas(64) AttnParams {
  CUtensorMap map_q;  // {64, H, Lq, B}
  CUtensorMap map_k;  // {64, H, Lk, Bkv}
  CUtensorMap map_v;
  __half* out;
  long ldo;           // elements between consecutive tokens of out
  long out_batch_stride;
  int lq, lk, heads;
  int kv_batch_div;   // kv batch index = batch / kv_batch_div (context shared by the frames of a video)
  float scale_log2;   // softmax scale * log2(e)
};

__global__ void __launch_bounds__(kAttnThreads, 1) attn_sm100_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // 2 x 16 KB
  uint8_t* sKV = sQ + 2 * kQBytes;                     // stages x (K 16 KB + V 16 KB)
  uint8_t* sP = sKV + kKvStages * 2 * kKBytes;         // 2 x 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kPBytes);
  uint64_t* q_full = bars;            // [1]
  uint64_t* kv_full = bars + 1;       // [2]
  uint64_t* kv_empty = bars + 3;      // [2]
  uint64_t* s_full = bars + 5;        // [2] per q-tile
  uint64_t* p_full = bars + 7;        // [2]
  uint64_t* o_full = bars + 9;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qblk = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
  const int q0 = qblk * 2 * kTileQ;
  const int nkv = (p.lk + kTileK - 1) / kTileK;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&p.map_q);
    tma_prefetch_desc(&p.map_k);
    tma_prefetch_desc(&p.map_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384)

  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * kQBytes);
      tma_load_4d(sQ, &p.map_q, q_full, 0, head, q0, batch);
      tma_load_4d(sQ + kQBytes, &p.map_q, q_full, 0, head, q0 + kTileQ, batch);
      const int kvb = batch / p.kv_batch_div;
      for (int j = 0; j < nkv; ++j) {
        const int st = j % kKvStages;
        const uint32_t ph = (j / kKvStages) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1, 10);
        mbar_expect_tx(&kv_full[st], 2 * kKBytes);
        uint8_t* dst = sKV + st * 2 * kKBytes;
        tma_load_4d(dst, &p.map_k, &kv_full[st], 0, head, j * kTileK, kvb);
        tma_load_4d(dst + kKBytes, &p.map_v, &kv_full[st], 0, head, j * kTileK, kvb);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      const uint32_t idesc_qk = umma_idesc_f16(kTileQ, kTileK, 0, 0);
      const uint32_t idesc_pv = umma_idesc_f16(kTileQ, kD, 0, 1);  // B (= V) is MN-major
      auto issue_qk = [&](int i, int j) {
        const int st = j % kKvStages;
        const uint64_t a_desc = umma_desc_sw128(smem_u32(sQ + i * kQBytes), 16, 1024);
        const uint64_t b_desc = umma_desc_sw128(smem_u32(sKV + st * 2 * kKBytes), 16, 1024);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) umma_f16_ss(tmem + i * 128, a_desc + 2 * k, b_desc + 2 * k, idesc_qk, k != 0);
        umma_commit(&s_full[i]);
      };
      mbar_wait(q_full, 0, 11);
      mbar_wait(&kv_full[0], 0, 12);
      tc_fence_after();
      issue_qk(0, 0);
      issue_qk(1, 0);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % kKvStages;
        if (j + 1 < nkv) {
          mbar_wait(&kv_full[(j + 1) % kKvStages], ((j + 1) / kKvStages) & 1, 13);
        }
        for (int i = 0; i < 2; ++i) {
          mbar_wait(&p_full[i], j & 1, 14);
          tc_fence_after();
          const uint32_t v_addr = smem_u32(sKV + st * 2 * kKBytes + kKBytes);
          const uint32_t p_addr = smem_u32(sP + i * kPBytes);
#pragma unroll
          for (int ks = 0; ks < kTileK / 16; ++ks) {
            // A: P chunk (ks/4) of 64 keys, 16-key step (ks%4) inside the 128B swizzle row
            const uint64_t a_desc = umma_desc_sw128(p_addr + (ks >> 2) * (kTileQ * 128) + (ks & 3) * 32, 16, 1024);
            // B: V rows [16*ks, 16*ks+16) (K dimension), 64 contiguous d (MN-major), 8-row groups 1024 B apart
            const uint64_t b_desc = umma_desc_sw128(v_addr + ks * 16 * 128, 1024, 1024);
            umma_f16_ss(tmem + 256 + i * 64, a_desc, b_desc, idesc_pv, ks != 0);
          }
          umma_commit(&o_full[i]);
          if (i == 1) umma_commit(&kv_empty[st]);  // K(j), V(j) fully consumed by both tiles
          if (j + 1 < nkv) issue_qk(i, j + 1);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax / output warps
    const int i = warp >> 2;          // q-tile
    const int q = warp & 3;           // TMEM lane quadrant
    const int r = q * 32 + lane;      // row within the tile
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t t_s = tmem + i * 128 + lane_base;
    const uint32_t t_o = tmem + 256 + i * 64 + lane_base;
    uint8_t* prow = sP + i * kPBytes + r * 128;
    const int sw = r & 7;

    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;
    float o_acc[kD];
#pragma unroll
    for (int d = 0; d < kD; ++d) o_acc[d] = 0.f;

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&s_full[i], j & 1, 20);
      tc_fence_after();
      const int valid = p.lk - j * kTileK;  // keys valid in this block (>= 128 unless last)
      // pass 1: row max
      float m_blk = -INFINITY;
#pragma unroll
      for (int c = 0; c < kTileK; c += 32) {
        uint32_t v[32];
        tmem_ld32(t_s + c, v);
        tmem_ld_wait();
        if (valid >= c + 32) {
#pragma unroll
          for (int t = 0; t < 32; ++t) m_blk = fmaxf(m_blk, __uint_as_float(v[t]));
        } else {
#pragma unroll
          for (int t = 0; t < 32; ++t)
            if (c + t < valid) m_blk = fmaxf(m_blk, __uint_as_float(v[t]));
        }
      }
      const float m_new = fmaxf(m_run, m_blk);
      const float alpha = fast_exp2((m_run - m_new) * p.scale_log2);  // 0 on the first block (m_run = -inf)
      const float neg_ms = -m_new * p.scale_log2;

      // fold the previous block's P V into the accumulator (also guarantees P(j-1) has been consumed)
      if (j > 0) {
        mbar_wait(&o_full[i], (j - 1) & 1, 21);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < kD; c += 32) {
          uint32_t v[32];
          tmem_ld32(t_o + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t) o_acc[c + t] = o_acc[c + t] * alpha_prev + __uint_as_float(v[t]);
        }
      }
      // pass 2: p = exp2(s*scale*log2e - m*scale*log2e), row sum, P -> smem (swizzled K-major)
      float l_blk = 0.f;
#pragma unroll
      for (int c = 0; c < kTileK; c += 32) {
        uint32_t v[32];
        tmem_ld32(t_s + c, v);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          float e = fast_exp2(fmaf(__uint_as_float(v[t]), p.scale_log2, neg_ms));
          if (c + t >= valid) e = 0.f;
          pv[t] = e;
          l_blk += e;
        }
        uint8_t* chunk = prow + (c >> 6) * (kTileQ * 128);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_half2(pv[g * 8 + 0], pv[g * 8 + 1]);
          u.y = pack_half2(pv[g * 8 + 2], pv[g * 8 + 3]);
          u.z = pack_half2(pv[g * 8 + 4], pv[g * 8 + 5]);
          u.w = pack_half2(pv[g * 8 + 6], pv[g * 8 + 7]);
          const int piece = ((c & 63) >> 3) + g;  // 16-byte piece index inside the 128 B row
          *reinterpret_cast<uint4*>(chunk + ((piece ^ sw) << 4)) = u;
        }
      }
      l_run = l_run * alpha + l_blk;
      m_run = m_new;
      alpha_prev = alpha;
      // S(j) fully read and P(j) written: publish to the MMA warp
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[i]);
    }
    // last block's P V
    mbar_wait(&o_full[i], (nkv - 1) & 1, 22);
    tc_fence_after();
    const int row = q0 + i * kTileQ + r;
    const float inv_l = 1.0f / l_run;
    __half* orow = p.out + (long)batch * p.out_batch_stride + (long)row * p.ldo + head * kD;
#pragma unroll
    for (int c = 0; c < kD; c += 32) {
      uint32_t v[32];
      tmem_ld32(t_o + c, v);
      tmem_ld_wait();
      if (row < p.lq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float f[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) f[t] = (o_acc[c + g * 8 + t] * alpha_prev + __uint_as_float(v[g * 8 + t])) * inv_l;
          uint4 u;
          u.x = pack_half2(f[0], f[1]);
          u.y = pack_half2(f[2], f[3]);
          u.z = pack_half2(f[4], f[5]);
          u.w = pack_half2(f[6], f[7]);
          *reinterpret_cast<uint4*>(orow + c + g * 8) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

}  // namespace vg

using namespace vg;

extern "C" int vgen_attention_d64(const void* q, const void* k, const void* v, void* out, int64_t batch, int64_t heads,
                                  int64_t lq, int64_t lk, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                                  int64_t kv_batch_div, float scale, void* stream) {
  VG_REQUIRE(q && k && v && out, "vgen_attention_d64: null pointer");
  VG_REQUIRE(batch >= 0 && heads > 0 && lq > 0 && lk > 0 && kv_batch_div >= 1 && batch % kv_batch_div == 0,
             "vgen_attention_d64: bad shape");
  VG_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "vgen_attention_d64: strides must be multiples of 8");
  VG_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(v) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "vgen_attention_d64: pointers must be 16-byte aligned");
  VG_REQUIRE(heads <= 65535 && batch <= 65535, "vgen_attention_d64: grid too large");
  if (batch == 0) return 0;
  AttnParams p;
  // dims ordered so that the byte strides ascend: {d, head, token, batch}
  const uint32_t box[4] = {64, 1, 128, 1};
  {
    const uint64_t dims[4] = {64, (uint64_t)heads, (uint64_t)lq, (uint64_t)batch};
    const uint64_t str[3] = {128, (uint64_t)ldq * 2, (uint64_t)lq * ldq * 2};
    int rc = make_tmap_f16(&p.map_q, q, 4, dims, str, box);
    if (rc) return rc;
  }
  const uint64_t bkv = batch / kv_batch_div;
  {
    const uint64_t dims[4] = {64, (uint64_t)heads, (uint64_t)lk, bkv};
    const uint64_t str[3] = {128, (uint64_t)ldk * 2, (uint64_t)lk * ldk * 2};
    int rc = make_tmap_f16(&p.map_k, k, 4, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {64, (uint64_t)heads, (uint64_t)lk, bkv};
    const uint64_t str[3] = {128, (uint64_t)ldv * 2, (uint64_t)lk * ldv * 2};
    int rc = make_tmap_f16(&p.map_v, v, 4, dims, str, box);
    if (rc) return rc;
  }
  p.out = reinterpret_cast<__half*>(out);
  p.ldo = ldo;
  p.out_batch_stride = lq * ldo;
  p.lq = (int)lq;
  p.lk = (int)lk;
  p.heads = (int)heads;
  p.kv_batch_div = (int)kv_batch_div;
  p.scale_log2 = scale * 1.4426950408889634f;
  const size_t smem = 2 * kQBytes + kKvStages * 2 * kKBytes + 2 * kPBytes + 12 * 8 + 16 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    VG_CUDA(cudaFuncSetAttribute(attn_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  dim3 grid((unsigned)cdiv(lq, 2 * kTileQ), (unsigned)heads, (unsigned)batch);
  attn_sm100_kernel<<<grid, kAttnThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  VG_LAUNCH_CHECK("attn_sm100_kernel");
  return 0;
}
