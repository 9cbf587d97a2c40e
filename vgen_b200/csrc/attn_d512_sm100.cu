// Single-head flash attention with head_dim 512 on tcgen05: the mid-block attention of the SD-2.1 VAE
// (AttnBlock, /root/reference/tools/modules/autoencoder.py:365-389: softmax(q k^T / sqrt(512)) v over h*w = 14 080 tokens
// at 1280x704).  The reference materialises the [hw, hw] score matrix in fp32 (793 MB per image); round 1 of this
// repo materialised it in fp16.  Here the scores never leave the SM: S lives in TMEM, P in shared memory, O in TMEM.
//
// One CTA = 128 query rows x ONE HALF (256 columns) of the output, all keys in blocks of 64:
//   S[128 x 64]  = sum over the 8 channel chunks c of  Q_c[128 x 64] . K_c[64 x 64]^T      (32 tcgen05.mma, K = 16 each)
//   P            = exp2(S * scale*log2e - m)  (fp32 -> fp16, thread-per-row online softmax with lazy rescaling)
//   O[128 x 256] += P[128 x 64] . V[64 x 256]                                               (4 sub-blocks x 4 mma)
// The two output halves of a query tile are independent CTAs (blockIdx.y): each recomputes S and the exponentials --
// 1.5x the MMA work of an unsplit kernel, which O[128 x 512] fp32 = all 512 TMEM columns (no room for S) rules out.
// The op is ~4 % of a decode, so simplicity wins: single q-tile, no ping-pong, tensor-bound (1 536 MMA cycles vs 512
// MUFU cycles per block).
//
//   warps 0-3  softmax / output (thread r <-> TMEM lane r)
//   warp 4     TMA producer for Q (8 chunks, once) and K: 8 chunks per key block through a 6-slot ring
//   warp 5     tcgen05.mma issue; S is double-buffered in TMEM so QK(j+1) runs under the softmax of block j
//   warp 6     TMA producer for V: the two 128-column halves of a block through 2 slots.  K and V have their OWN producers:
//              V(j)'s slot frees only when PV(j-1) has run, i.e. after the softmax of block j-1, and an in-order producer
//              that blocks there holds back K(j+1) -- measured 6 000 cycles per block against 1 536 cycles of MMA work
// TMEM columns: S0 [0,64)  S1 [64,128)  O [128,384).
#include "common.h"
#include "ptx.cuh"

namespace vg {

static constexpr int kT5Threads = 224;
static constexpr int kT5D = 512;
static constexpr int kT5Chunks = kT5D / 64;        // 8 channel chunks of 64
static constexpr int kT5Half = 256;                // output columns per CTA
static constexpr int kT5TileQ = 128;
static constexpr int kT5TileK = 64;
static constexpr int kT5QChunkBytes = kT5TileQ * 64 * 2;   // 16 KB
static constexpr int kT5KChunkBytes = kT5TileK * 64 * 2;   // 8 KB
static constexpr int kT5VSlotBytes = kT5TileK * 128 * 2;   // 16 KB: 64 keys x 128 channels = two 64-channel sub-blocks
static constexpr int kT5PBytes = kT5TileQ * kT5TileK * 2;  // 16 KB
static constexpr int kT5KSlots = 6;
static constexpr float kT5Rescale = 8.0f;

struct alignas(64) Attn512Params {
  CUtensorMap map_q;   // {512, Lq, B}, box {64, 128, 1}
  CUtensorMap map_k;   // {512, Lk, B}, box {64, 64, 1}
  CUtensorMap map_v;   // {512, Lk, B}, box {64, 64, 1}
  __half* out;
  long ldo, out_batch_stride;
  int lq, lk;
  float scale_log2;
};

__global__ void __launch_bounds__(kT5Threads, 1) attn_d512_sm100_kernel(const __grid_constant__ Attn512Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                     // 8 x 16 KB
  uint8_t* sK = sQ + kT5Chunks * kT5QChunkBytes;          // 6 x 8 KB
  uint8_t* sV = sK + kT5KSlots * kT5KChunkBytes;          // 2 x 16 KB
  uint8_t* sP = sV + 2 * kT5VSlotBytes;                   // 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kT5PBytes);
  uint64_t* q_full = bars;                       // [1]
  uint64_t* k_full = bars + 1;                   // [kT5KSlots]
  uint64_t* k_empty = k_full + kT5KSlots;        // [kT5KSlots]
  uint64_t* v_full = k_empty + kT5KSlots;        // [2]
  uint64_t* v_empty = v_full + 2;                // [2]
  uint64_t* s_full = v_empty + 2;                // [2] per S buffer
  uint64_t* s_free = s_full + 2;                 // [2] per S buffer: scores are in registers
  uint64_t* p_full = s_free + 2;                 // [1] P(j) in shared memory, O rescaled
  uint64_t* o_full = p_full + 1;                 // [1] PV(j) finished: P buffer free, O readable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kT5TileQ, half = blockIdx.y, batch = blockIdx.z;
  const int nkv = (p.lk + kT5TileK - 1) / kT5TileK;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&p.map_q);
    tma_prefetch_desc(&p.map_k);
    tma_prefetch_desc(&p.map_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kT5KSlots; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
    }
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer: Q once, then the K chunk ring
    if (elect_one()) {
      mbar_expect_tx(q_full, kT5Chunks * kT5QChunkBytes);
      for (int c = 0; c < kT5Chunks; ++c) tma_load_3d(sQ + c * kT5QChunkBytes, &p.map_q, q_full, c * 64, q0, batch);
    }
    __syncwarp();
    int slot = 0, use = 0;                   // ring position; `use` = how many times this slot has been filled before
    for (int j = 0; j < nkv; ++j) {
      for (int c = 0; c < kT5Chunks; ++c) {
        if (use > 0) mbar_wait(&k_empty[slot], (use - 1) & 1, 30);
        if (elect_one()) {
          mbar_expect_tx(&k_full[slot], kT5KChunkBytes);
          tma_load_3d(sK + slot * kT5KChunkBytes, &p.map_k, &k_full[slot], c * 64, j * kT5TileK, batch);
        }
        __syncwarp();
        if (++slot == kT5KSlots) { slot = 0; ++use; }
      }
    }
  } else if (warp == 6) {
    // ------------------------------------------------------------------ TMA producer: V halves (slot h)
    for (int j = 0; j < nkv; ++j) {
      for (int h = 0; h < 2; ++h) {   // channels [half*256 + h*128, +128) as two 64-channel sub-blocks
        if (j >= 1) mbar_wait(&v_empty[h], (j - 1) & 1, 31);
        if (elect_one()) {
          mbar_expect_tx(&v_full[h], kT5VSlotBytes);
          const int ch0 = half * kT5Half + h * 128;
          tma_load_3d(sV + h * kT5VSlotBytes, &p.map_v, &v_full[h], ch0, j * kT5TileK, batch);
          tma_load_3d(sV + h * kT5VSlotBytes + kT5KChunkBytes, &p.map_v, &v_full[h], ch0 + 64, j * kT5TileK, batch);
        }
        __syncwarp();
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issue (whole warp walks the loop)
    const uint32_t idesc_qk = umma_idesc_f16(kT5TileQ, kT5TileK, 0, 0);   // 128 x 64, both K-major
    const uint32_t idesc_pv = umma_idesc_f16(kT5TileQ, 64, 0, 1);         // 128 x 64, B (= V) MN-major
    const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP);
    const uint32_t t_o = tmem + 128;
    int kslot = 0, kuse = 0;
    auto issue_qk = [&](int j) {
      const uint32_t t_s = tmem + (j & 1) * 64;
      for (int c = 0; c < kT5Chunks; ++c) {
        const int slot = kslot;
        mbar_wait(&k_full[slot], kuse & 1, 32);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a_desc = umma_desc_sw128(q_addr + c * kT5QChunkBytes, 16, 1024);
          const uint64_t b_desc = umma_desc_sw128(k_addr + slot * kT5KChunkBytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(t_s, a_desc + 2 * k, b_desc + 2 * k, idesc_qk, (c | k) != 0);
          umma_commit(&k_empty[slot]);
          if (c == kT5Chunks - 1) umma_commit(&s_full[j & 1]);
        }
        __syncwarp();
        if (++kslot == kT5KSlots) { kslot = 0; ++kuse; }
      }
    };
    mbar_wait(q_full, 0, 33);
    issue_qk(0);
    for (int j = 0; j < nkv; ++j) {
      if (j + 1 < nkv) {
        if (j + 1 >= 2) mbar_wait(&s_free[(j + 1) & 1], ((j - 1) >> 1) & 1, 34);   // S(j-1) is in registers
        issue_qk(j + 1);
      }
      mbar_wait(p_full, j & 1, 35);
      for (int h = 0; h < 2; ++h) {
        mbar_wait(&v_full[h], j & 1, 36);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int sb = 0; sb < 2; ++sb) {        // 64-channel sub-block: its own single-atom MN-major B tile
            const uint32_t vb = v_addr + h * kT5VSlotBytes + sb * kT5KChunkBytes;
#pragma unroll
            for (int ks = 0; ks < kT5TileK / 16; ++ks) {
              const uint64_t a_desc = umma_desc_sw128(p_addr + ks * 32, 16, 1024);
              const uint64_t b_desc = umma_desc_sw128(vb + ks * 16 * 128, 1024, 1024);
              umma_f16_ss(t_o + (h * 2 + sb) * 64, a_desc, b_desc, idesc_pv, (j | ks) != 0);
            }
          }
          umma_commit(&v_empty[h]);
          if (h == 1) umma_commit(o_full);
        }
        __syncwarp();
      }
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------------ softmax / output warps
    const int r = warp * 32 + lane;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t t_o = tmem + 128 + lane_base;
    const uint32_t prow = smem_u32(sP + r * 128);
    const int sw = r & 7;
    const float sl2 = p.scale_log2;
    float m_ref = -INFINITY, l_run = 0.f;

    for (int j = 0; j < nkv; ++j) {
      const uint32_t t_s = tmem + (j & 1) * 64 + lane_base;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1, 40);
      tc_fence_after();
      const int valid = p.lk - j * kT5TileK;
      uint32_t s0[32], s1[32];
      tmem_ld32(t_s, s0);
      tmem_ld32(t_s + 32, s1);
      tmem_ld_wait();
      if (valid < kT5TileK) {
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          if (t >= valid) s0[t] = 0xff800000u;
          if (32 + t >= valid) s1[t] = 0xff800000u;
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        mx0 = fmaxf(mx0, __uint_as_float(s0[t]));
        mx1 = fmaxf(mx1, __uint_as_float(s1[t]));
      }
      const float m_blk = fmaxf(mx0, mx1);
      const bool grow = (m_blk - m_ref) * sl2 > kT5Rescale;      // always true on the first block
      // the scores are in registers: QK(j+2) may overwrite this S buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[j & 1]);
      // PV(j-1) done: the P buffer is free and O is up to date
      if (j >= 1) {
        mbar_wait(o_full, (j - 1) & 1, 41);
        tc_fence_after();
      }
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? m_blk : m_ref;
        const float alpha = (j == 0) ? 0.f : fast_exp2((m_ref - m_new) * sl2);
        if (j > 0) {
          uint32_t o[32];
#pragma unroll 1
          for (int c = 0; c < kT5Half; c += 32) {
            tmem_ld32(t_o + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 32; ++t) o[t] = __float_as_uint(__uint_as_float(o[t]) * alpha);
            tmem_st32(t_o + c, o);
          }
          tmem_st_wait();
        }
        l_run *= alpha;
        m_ref = m_new;
      }
      const float neg_ms = -m_ref * sl2;
      float l0 = 0.f, l1 = 0.f;
#define VG_T5_EMIT(ARR, C0, LSUM)                                                                     \
  _Pragma("unroll") for (int g = 0; g < 4; ++g) {                                                     \
    float e[8];                                                                                       \
    _Pragma("unroll") for (int t = 0; t < 8; ++t) {                                                   \
      e[t] = fast_exp2(fmaf(__uint_as_float(ARR[g * 8 + t]), sl2, neg_ms));                           \
      LSUM += e[t];                                                                                   \
    }                                                                                                 \
    const int piece = ((C0) >> 3) + g;                                                                \
    st_shared_v4(prow + ((piece ^ sw) << 4), pack_half2(e[0], e[1]), pack_half2(e[2], e[3]),          \
                 pack_half2(e[4], e[5]), pack_half2(e[6], e[7]));                                     \
  }
      VG_T5_EMIT(s0, 0, l0)
      VG_T5_EMIT(s1, 32, l1)
#undef VG_T5_EMIT
      l_run += l0 + l1;
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- output: O / l for this CTA's 256 columns
    mbar_wait(o_full, (nkv - 1) & 1, 42);
    tc_fence_after();
    const int row = q0 + r;
    const float inv_l = 1.0f / l_run;
    __half* orow = p.out + (long)batch * p.out_batch_stride + (long)row * p.ldo + half * kT5Half;
#pragma unroll 1
    for (int c = 0; c < kT5Half; c += 32) {
      uint32_t o[32];
      tmem_ld32(t_o + c, o);
      tmem_ld_wait();
      if (row < p.lq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_half2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
          u.y = pack_half2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
          u.z = pack_half2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
          u.w = pack_half2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c + g * 8) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

}  // namespace vg

using namespace vg;

extern "C" int vgen_attention_d512(const void* q, const void* k, const void* v, void* out, int64_t batch, int64_t lq,
                                   int64_t lk, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float scale, void* stream) {
  VG_REQUIRE(q && k && v && out, "vgen_attention_d512: null pointer");
  VG_REQUIRE(batch >= 0 && lq > 0 && lk > 0, "vgen_attention_d512: bad shape");
  VG_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && ldq >= 512 && ldk >= 512 && ldv >= 512 && ldo >= 512,
             "vgen_attention_d512: row strides must be multiples of 8 and >= 512");
  VG_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
               reinterpret_cast<uintptr_t>(out)) & 15) == 0,
             "vgen_attention_d512: pointers must be 16-byte aligned");
  VG_REQUIRE(batch <= 65535, "vgen_attention_d512: grid too large");
  if (batch == 0) return 0;
  Attn512Params p;
  {
    const uint64_t dims[3] = {512, (uint64_t)lq, (uint64_t)batch};
    const uint64_t str[2] = {(uint64_t)ldq * 2, (uint64_t)lq * ldq * 2};
    const uint32_t box[3] = {64, 128, 1};
    int rc = make_tmap_f16(&p.map_q, q, 3, dims, str, box);
    if (rc) return rc;
  }
  const uint32_t boxkv[3] = {64, 64, 1};
  {
    const uint64_t dims[3] = {512, (uint64_t)lk, (uint64_t)batch};
    const uint64_t str[2] = {(uint64_t)ldk * 2, (uint64_t)lk * ldk * 2};
    int rc = make_tmap_f16(&p.map_k, k, 3, dims, str, boxkv);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {512, (uint64_t)lk, (uint64_t)batch};
    const uint64_t str[2] = {(uint64_t)ldv * 2, (uint64_t)lk * ldv * 2};
    int rc = make_tmap_f16(&p.map_v, v, 3, dims, str, boxkv);
    if (rc) return rc;
  }
  p.out = reinterpret_cast<__half*>(out);
  p.ldo = ldo;
  p.out_batch_stride = lq * ldo;
  p.lq = (int)lq;
  p.lk = (int)lk;
  p.scale_log2 = scale * 1.4426950408889634f;
  const size_t smem = kT5Chunks * kT5QChunkBytes + kT5KSlots * kT5KChunkBytes + 2 * kT5VSlotBytes + kT5PBytes + (2 * kT5KSlots + 12) * 8 + 16 + 1024;
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    VG_CUDA(cudaFuncSetAttribute(attn_d512_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_once.mark();
  }
  dim3 grid((unsigned)cdiv(lq, kT5TileQ), 2, (unsigned)batch);
  launch_kernel(attn_d512_sm100_kernel, dim3(grid), dim3(kT5Threads), smem, reinterpret_cast<cudaStream_t>(stream), p);
  VG_LAUNCH_CHECK("attn_d512_sm100_kernel");
  return 0;
}
