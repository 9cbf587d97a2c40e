// Flash attention (head_dim 64, no mask) on tcgen05: S = Q K^T and O = P V on the 5th-gen tensor
// cores with BOTH accumulators in TMEM, online softmax in registers (one thread per query row, so no
// shuffles), operands staged by TMA.
//
// Replaces xformers.ops.memory_efficient_attention as called by MemoryEfficientCrossAttention
// (/root/reference/tools/modules/unet/util.py:231-269): exact softmax(q k^T / sqrt(64)) v per
// (batch, head); fp16 inputs, fp32 scores / softmax / accumulation, P rounded to fp16 for the PV
// product (what the CUTLASS/FA2 kernels behind xformers do).
//
// Two kernel families live in this file; vgen_attention_d64 picks by key length (attn_default_tiles, table at the launcher):
//   SS family (first body): P goes through shared memory (SS-form tcgen05.mma).  Serves lk > 4096.
//   TS family (second body, "TS" comment below): P stays in tensor memory (TS-form tcgen05.mma), small CTAs, 2-3 per SM.
//
// SS family.  The body is a template on <kNT q-tiles per CTA, kTileK keys per block>; two instantiations are built:
//   <2, 128>  256 query rows / CTA, 11 warps, 128-key blocks (168 registers / thread): the default
//   <3, 64>   384 query rows / CTA, 16 warps, 64-key blocks (118 registers / thread): three softmax warps per SM sub-partition.
//             Built to test whether a third warp fills the MUFU pipe while the others sit in their MUFU-free phases; it does
//             not (9.79 vs 9.70 ms on the 14080^2 launch: the doubled barrier traffic per key eats the gain) -- knob only.
// One CTA = kNT 128-row tiles of one (batch, head), 5 kNT + 1 warps:
//   warps 4i .. 4i+3   softmax + output for q-tile i (thread r <-> TMEM lane r)
//   warp 4 kNT         TMA producer: Q once, then separate rings of K and V blocks of kTileK keys (K(j) is
//                      released as soon as every tile's QK(j) ran, long before V(j))
//   warps 4 kNT+1+i    tcgen05.mma issue for q-tile i (whole warp walks the loop, one elected lane issues); one
//                      issuer per tile halves the event -> issue reaction time (a single warp polling both tiles'
//                      four barriers needed ~600 cycles per round, behind the softmax warps' MUFU traffic)
// The two q-tiles ping-pong on the tensor pipe: while the softmax warps of tile 0 work on S0(j+1), the
// tensor core runs PV1(j) and QK1(j+1).  P is double-buffered per tile, so the exponentials of block j never
// wait for PV(j-1) (that wait was 23% of the softmax warps' time with a single P buffer: the PV round trip
// -- poll, issue, 256 tensor cycles, commit, wake-up -- is ~900 cycles); only the rare O rescale does.
//
// Per key block and tile the softmax thread reads its 128 scores from TMEM ONCE (four tcgen05.ld in flight,
// one wait), takes the row max, and exponentiates against a reference max m_ref that is only advanced when
// the true max has grown by more than 2^8 ("lazy rescaling", as in FlashAttention-4): P stays <= 256, which
// fp16 holds, so O can accumulate in TMEM across key blocks (tcgen05.mma accumulate) without a per-block
// read-modify-write; on the rare advance the thread rescales its O row in TMEM (tcgen05.ld / st).  The final
// O / l is exact: every term of row r carries the same 2^-m_ref factor.
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace vg {

static constexpr int kD = 64;
static constexpr int kTileQ = 128;   // rows per q-tile (UMMA M)
static constexpr int kQBytes = kTileQ * kD * 2;   // 16 KB
static constexpr float kRescaleThreshold = 8.0f;     // log2 units: P <= 2^8

// Compile-time geometry of one instantiation (shared by the kernel and its launcher).
template <int kNT_, int kTileK_>
struct AttnCfg {
  static constexpr int kNT = kNT_;                      // q-tiles per CTA
  static constexpr int kTileK = kTileK_;                // keys per block (UMMA N of QK^T)
  static constexpr int kThreads = 32 * (5 * kNT + 1);   // 4 softmax warps + 1 issuer per tile, 1 producer
  static constexpr int kKBytes = kTileK * kD * 2;       // one K (or V) block
  static constexpr int kPBytes = kTileQ * kTileK * 2;   // one P buffer of one q-tile
  static constexpr int kStageLog2 = (kTileK == 128) ? 1 : 2;
  static constexpr int kKvStages = 1 << kStageLog2;     // K ring and V ring: 2 x 16 KB or 4 x 8 KB each
  static constexpr int kNumBars = 1 + 4 * kKvStages + 7 * kNT;
  static constexpr int kDataBytes = kNT * kQBytes + 2 * kKvStages * kKBytes + 2 * kNT * kPBytes;
  static_assert(kTileK == 64 || kTileK == 128, "keys per block");
  static_assert(kNT * kTileK + kNT * 80 <= 512, "TMEM columns");
  static constexpr int kTcSlack = 816;
  static_assert(kDataBytes + kNumBars * 8 + 16 + 2048 + kTcSlack <= 227 * 1024, "shared memory");
  static constexpr size_t smem_bytes(bool tcsum) {
    // tcsum: + 2 KB ones atom; the 1024-byte alignment slack shrinks to kTcSlack so the <2,128> total stays within 227 KB
    // (the kernel traps if the dynamic shared memory base is ever less aligned than that)
    return (size_t)kDataBytes + kNumBars * 8 + 16 + (tcsum ? 2048 + kTcSlack : 1024);
  }
};

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
  int tcsum;          // timing twin only: which variant to instrument
  int stagger;        // tile i's first QK waits until tile i-1 has pulled its first scores (phase offset between the tiles)
  long long* timing;  // kTiming only: per-phase cycle counters of one CTA's first two tiles (tools/bench_attn.py)
};

// kTcSum: the softmax row sums are taken by the TENSOR CORE: the PV product runs with N = 80 instead of 64, the 16 extra
// columns of "V" being a constant shared-memory atom whose first column is 1.0 (the second MN-atom of the B descriptor, reached
// through its leading-byte-offset), so O[:, 64] accumulates sum_k P[r, k] in fp32 -- exactly the normaliser of the fp16 P the
// product used.  That removes one dependent FADD per exponential from the softmax warps (the FMA pipe shares issue slots
// with the MUFU-bound exponentials) at the cost of 25 % more PV tensor time, which has slack (tensor pipe 33 % busy).
template <class Cfg, bool kTiming, bool kTcSum>
__device__ __forceinline__ void attn_sm100_body(const AttnParams& p) {
  constexpr int kNT = Cfg::kNT, kTileK = Cfg::kTileK, kKBytes = Cfg::kKBytes, kPBytes = Cfg::kPBytes;
  constexpr int kKvStages = Cfg::kKvStages, kStageLog2 = Cfg::kStageLog2;
  constexpr int kSG = kTileK / 32;             // 32-column score groups per block
  constexpr int kOCols = kTcSum ? 80 : 64;     // TMEM columns of one q-tile's O accumulator
  constexpr int kOBase = kNT * kTileK;         // TMEM columns: S_i [i kTileK, (i+1) kTileK), O_i [kOBase + i kOCols, ...)
  constexpr int kProducerWarp = 4 * kNT;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // kNT x 16 KB
  uint8_t* sK = sQ + kNT * kQBytes;                    // ring of K blocks
  uint8_t* sV = sK + kKvStages * kKBytes;              // ring of V blocks
  uint8_t* sP = sV + kKvStages * kKBytes;              // kNT tiles x 2 buffers
  uint8_t* sOnes = sP + 2 * kNT * kPBytes;             // kTcSum: 2 KB constant MN-major atom (16 keys x 64 columns, column 0 = 1)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + (kTcSum ? 2048 : 0));
  if (kTcSum && (reinterpret_cast<uintptr_t>(smem) - reinterpret_cast<uintptr_t>(smem_raw)) > Cfg::kTcSlack) __trap();  // launcher's slack
  uint64_t* q_full = bars;                        // [1]
  uint64_t* k_full = q_full + 1;                  // [stages]
  uint64_t* k_empty = k_full + kKvStages;         // [stages] every tile's QK(j) finished
  uint64_t* v_full = k_empty + kKvStages;         // [stages]
  uint64_t* v_empty = v_full + kKvStages;         // [stages] every tile's PV(j) finished
  uint64_t* s_full = v_empty + kKvStages;         // [kNT] per q-tile: S(j) written by QK
  uint64_t* p_full = s_full + kNT;                // [kNT][2] per q-tile and P buffer: P(j) in smem, S(j) consumed, O rescaled
  uint64_t* o_full = p_full + 2 * kNT;            // [kNT][2] per q-tile and P buffer: PV(j) finished (buffer j&1 free, O readable)
  uint64_t* s_free = o_full + 2 * kNT;            // [kNT] per q-tile: S(j) is in registers, QK(j+1) may overwrite it
  uint64_t* stag = s_free + kNT;                  // [kNT] per q-tile: its FIRST scores are in registers (one phase, p.stagger)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stag + kNT);
  static_assert(1 + 4 * Cfg::kKvStages + 7 * Cfg::kNT == Cfg::kNumBars, "barrier count");
  // (one barrier per P buffer: the softmax warps may publish P(j+1) before the MMA warp has looked at P(j), and a
  // single barrier two phases ahead of its observer aliases)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qblk = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
  const int q0 = qblk * kNT * kTileQ;
  const int nkv = (p.lk + kTileK - 1) / kTileK;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&p.map_q);
    tma_prefetch_desc(&p.map_k);
    tma_prefetch_desc(&p.map_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKvStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], kNT);   // one commit per tile's issuer
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], kNT);
    }
    for (int i = 0; i < kNT; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[2 * i], 4);
      mbar_init(&p_full[2 * i + 1], 4);
      mbar_init(&o_full[2 * i], 1);
      mbar_init(&o_full[2 * i + 1], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&stag[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == kProducerWarp + 1) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  if (kTcSum && warp < 4) {
    // rows = keys (two 8-row groups 1024 B apart, like the V tile), 128 B per row, 128B-swizzled: the 16-byte piece holding
    // column 0 of row r sits at piece index (0 ^ (r & 7))
    const int t = threadIdx.x;                       // 128 threads x 16 B = 2 KB
    const int row = t >> 3, piece = t & 7;
    const uint32_t val = (piece == (row & 7)) ? 0x00003C00u : 0u;   // fp16 1.0 in the low half
    st_shared_v4(smem_u32(sOnes) + t * 16, val, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kProducerWarp) {
    // ------------------------------------------------------------------ TMA producer
    const int kvb = batch / p.kv_batch_div;
    if (elect_one()) {
      mbar_expect_tx(q_full, kNT * kQBytes);
#pragma unroll
      for (int i = 0; i < kNT; ++i) tma_load_4d(sQ + i * kQBytes, &p.map_q, q_full, 0, head, q0 + i * kTileQ, batch);
    }
    __syncwarp();
    // K(j) then V(j), blocking on the slot's release (use u of slot s is block u * stages + s).  K(j)'s slot frees after
    // every tile's QK(j - stages), V(j)'s after every PV(j - stages) -- each at least a block before the data is needed again.
    for (int j = 0; j < nkv; ++j) {
      const int st = j & (kKvStages - 1);
      const uint32_t reuse_parity = ((j >> kStageLog2) - 1) & 1;
      if (j >= kKvStages) mbar_wait(&k_empty[st], reuse_parity, 10);
      if (elect_one()) {
        mbar_expect_tx(&k_full[st], kKBytes);
        tma_load_4d(sK + st * kKBytes, &p.map_k, &k_full[st], 0, head, j * kTileK, kvb);
      }
      __syncwarp();
      if (j >= kKvStages) mbar_wait(&v_empty[st], reuse_parity, 16);
      if (elect_one()) {
        mbar_expect_tx(&v_full[st], kKBytes);
        tma_load_4d(sV + st * kKBytes, &p.map_v, &v_full[st], 0, head, j * kTileK, kvb);
      }
      __syncwarp();
    }
  } else if (warp > kProducerWarp) {
    // ------------------------------------------------------------------ MMA issue for q-tile i (warp-uniform loop)
    const int i = warp - kProducerWarp - 1;
    const uint32_t idesc_qk = umma_idesc_f16(kTileQ, kTileK, 0, 0);
    const uint32_t idesc_pv = umma_idesc_f16(kTileQ, kOCols, 0, 1);  // B (= V [+ the ones atom]) is MN-major
    const uint32_t ones_addr = smem_u32(sOnes);
    const uint32_t q_addr = smem_u32(sQ) + i * kQBytes, k_addr = smem_u32(sK), v_addr0 = smem_u32(sV);
    const uint32_t p_addr0 = smem_u32(sP) + 2 * i * kPBytes;
    const uint32_t t_s = tmem + i * kTileK, t_o = tmem + kOBase + i * kOCols;
    // Event-driven issue: QK(j+1) goes out as soon as the softmax warps of this tile have pulled S(j) into registers
    // (s_free), i.e. it overlaps their exponentials; PV(j) goes out when P(j) is in shared memory.
    // Per tile the events arrive in a fixed order -- s_free(j) (scores in registers) always precedes p_full(j) (P
    // written) -- so the issuer simply blocks on them in turn (hardware-assisted mbarrier wait: no polling loop
    // competing with the softmax warps for issue slots, and a short wake-up):  QK(0); { QK(j+1); PV(j) } ...
    auto issue_qk = [&](int j) {
      const int st = j & (kKvStages - 1);
      mbar_wait(&k_full[st], (j >> kStageLog2) & 1, 12);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t a_desc = umma_desc_sw128(q_addr, 16, 1024);
        const uint64_t b_desc = umma_desc_sw128(k_addr + st * kKBytes, 16, 1024);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) umma_f16_ss(t_s, a_desc + 2 * k, b_desc + 2 * k, idesc_qk, k != 0);
        umma_commit(&s_full[i]);
        umma_commit(&k_empty[st]);  // this tile is done with K(j); the slot frees when every tile is
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0, 11);
    if (i > 0 && p.stagger) mbar_wait(&stag[i - 1], 0, 17);   // start one "load + max" phase behind the previous tile
    issue_qk(0);
    for (int j = 0; j < nkv; ++j) {
      if (j + 1 < nkv) {
        mbar_wait(&s_free[i], j & 1, 13);
        issue_qk(j + 1);
      }
      const int st = j & (kKvStages - 1);
      mbar_wait(&p_full[2 * i + (j & 1)], (j >> 1) & 1, 14);
      mbar_wait(&v_full[st], (j >> kStageLog2) & 1, 15);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t v_addr = v_addr0 + st * kKBytes;
        const uint32_t p_addr = p_addr0 + (j & 1) * kPBytes;
#pragma unroll
        for (int ks = 0; ks < kTileK / 16; ++ks) {
          // A: P chunk (ks/4) of 64 keys, 16-key step (ks%4) inside the 128B swizzle row
          const uint64_t a_desc = umma_desc_sw128(p_addr + (ks >> 2) * (kTileQ * 128) + (ks & 3) * 32, 16, 1024);
          // B: V rows [16*ks, 16*ks+16) (K dimension), 64 contiguous d (MN-major), 8-row groups 1024 B apart
          // (kTcSum: the second 64-column atom of B is the ones tile: leading byte offset = its distance from this K step)
          const uint32_t vk = v_addr + ks * 16 * 128;
          const uint64_t b_desc = umma_desc_sw128(vk, kTcSum ? ones_addr - vk : 1024u, 1024);
          umma_f16_ss(t_o, a_desc, b_desc, idesc_pv, (j | ks) != 0);  // O accumulates over blocks
        }
        umma_commit(&o_full[2 * i + (j & 1)]);
        umma_commit(&v_empty[st]);
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- softmax / output warps
    const int i = warp >> 2;          // q-tile
    const int q = warp & 3;           // TMEM lane quadrant
    const int r = q * 32 + lane;      // row within the tile
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t t_s = tmem + i * kTileK + lane_base;
    const uint32_t t_o = tmem + kOBase + i * kOCols + lane_base;
    const uint32_t prow0 = smem_u32(sP + 2 * i * kPBytes + r * 128);
    const int sw = r & 7;
    const float sl2 = p.scale_log2;

    float m_ref = -INFINITY;  // reference max (raw score units) all of this row's exponentials are relative to
    float l_run = 0.f;
    long long tm_wait_s = 0, tm_ldmax = 0, tm_wait_p = 0, tm_exp = 0, tm_c0 = 0, tm_c1 = 0, tm_start = 0;
    if (kTiming) tm_start = clock64();

    for (int j = 0; j < nkv; ++j) {
      if (kTiming) tm_c0 = clock64();
      mbar_wait(&s_full[i], j & 1, 20);
      tc_fence_after();
      if (kTiming) { tm_c1 = clock64(); tm_wait_s += tm_c1 - tm_c0; }
      const int valid = p.lk - j * kTileK;  // keys valid in this block (>= kTileK unless last)
      uint32_t s[kSG][32];
      auto load_scores = [&]() {
#pragma unroll
        for (int c = 0; c < kSG; ++c) tmem_ld32(t_s + 32 * c, s[c]);
        tmem_ld_wait();
        if (valid < kTileK) {  // last, partial block: masked keys must not contribute
#pragma unroll
          for (int c = 0; c < kSG; ++c) {
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              if (32 * c + t >= valid) s[c][t] = 0xff800000u;
            }
          }
        }
      };
      load_scores();
      float mx[kSG];
#pragma unroll
      for (int c = 0; c < kSG; ++c) mx[c] = -INFINITY;
#pragma unroll
      for (int t = 0; t < 32; ++t) {
#pragma unroll
        for (int c = 0; c < kSG; ++c) mx[c] = fmaxf(mx[c], __uint_as_float(s[c][t]));
      }
      float m_blk = fmaxf(mx[0], mx[1]);
      if (kSG == 4) m_blk = fmaxf(m_blk, fmaxf(mx[kSG - 2], mx[kSG - 1]));

      // advance the reference max only when the true max has outgrown it by more than the threshold
      const bool grow = (m_blk - m_ref) * sl2 > kRescaleThreshold;   // always true on the first block
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? m_blk : m_ref;
        const float alpha = (j == 0) ? 0.f : fast_exp2((m_ref - m_new) * sl2);  // 1 for rows that do not advance
        if (j > 0) {
          // O must be complete up to PV(j-1) before it is touched (the (j-1)/2-th use of that P buffer's barrier)
          mbar_wait(&o_full[2 * i + ((j - 1) & 1)], ((j - 1) >> 1) & 1, 21);
          tc_fence_after();
          // rescale this row of O in TMEM; the score registers are dead here and reloaded afterwards
          uint32_t o[32];
#pragma unroll 1
          for (int c = 0; c < kD; c += 32) {
            tmem_ld32(t_o + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 32; ++t) o[t] = __float_as_uint(__uint_as_float(o[t]) * alpha);
            tmem_st32(t_o + c, o);
          }
          if (kTcSum) {       // the row-sum column (and its 15 zero companions) rescales with O
            uint32_t o16[16];
            tmem_ld16(t_o + kD, o16);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 16; ++t) o16[t] = __float_as_uint(__uint_as_float(o16[t]) * alpha);
            tmem_st16(t_o + kD, o16);
          }
          tmem_st_wait();
          load_scores();
        }
        l_run *= alpha;
        m_ref = m_new;
      }
      // the scores are in registers for good: let the tensor core start QK(j+1) into this S buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s_free[i]);
        if (j == 0) mbar_arrive(&stag[i]);
      }
      if (kTiming) { tm_c0 = clock64(); tm_ldmax += tm_c0 - tm_c1; }
      // P buffer j&1 was last read by PV(j-2).  (The wait is implied by the s_full(j) this thread has observed -- a tcgen05.commit
      // covers all earlier MMAs of the issuing thread and PV(j-2) precedes QK(j) -- but dropping it measured slower: 9.95 vs
      // 9.70 ms <2,128>, 10.40 vs 9.79 ms <3,64> on the 14080^2 launch, profiles/r02q_attn_ss.log.)
      if (j >= 2) mbar_wait(&o_full[2 * i + (j & 1)], ((j >> 1) - 1) & 1, 23);
      if (kTiming) { tm_c1 = clock64(); tm_wait_p += tm_c1 - tm_c0; }
      const uint32_t prow = prow0 + (j & 1) * kPBytes;
      const float neg_ms = -m_ref * sl2;
      float lsum[kSG];
      // p = exp2(s*scale*log2e - m_ref*scale*log2e), row sum, P -> smem (128B-swizzled K-major UMMA layout: 64-key chunks of
      // [128 rows x 128 B])
#pragma unroll
      for (int c = 0; c < kSG; ++c) {
        lsum[c] = 0.f;
        const uint32_t chunk = prow + ((32 * c) >> 6) * (kTileQ * 128);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float e[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            e[t] = fast_exp2(fmaf(__uint_as_float(s[c][g * 8 + t]), sl2, neg_ms));
            if (!kTcSum) lsum[c] += e[t];
          }
          const int piece = (((32 * c) & 63) >> 3) + g;   // 16-byte piece inside the 128 B row
          st_shared_v4(chunk + ((piece ^ sw) << 4), pack_half2(e[0], e[1]), pack_half2(e[2], e[3]), pack_half2(e[4], e[5]),
                       pack_half2(e[6], e[7]));
        }
      }
      if (kSG == 4) l_run += (lsum[0] + lsum[1]) + (lsum[kSG - 2] + lsum[kSG - 1]);
      else l_run += lsum[0] + lsum[1];
      // S(j) consumed, P(j) written, O rescaled: publish to the MMA warp
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[2 * i + (j & 1)]);
      if (kTiming) tm_exp += clock64() - tm_c1;
    }
    if (kTiming && p.timing && i < 2 && q == 0 && lane == 0 && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0) {
      long long* t = p.timing + i * 8;
      t[0] = tm_wait_s, t[1] = tm_ldmax, t[2] = tm_wait_p, t[3] = tm_exp, t[4] = clock64() - tm_start, t[5] = nkv;
    }
    // ---- output: O / l
    mbar_wait(&o_full[2 * i + ((nkv - 1) & 1)], ((nkv - 1) >> 1) & 1, 22);
    tc_fence_after();
    const int row = q0 + i * kTileQ + r;
    if (kTcSum) {
      uint32_t o16[16];
      tmem_ld16(t_o + kD, o16);
      tmem_ld_wait();
      l_run = __uint_as_float(o16[0]);
    }
    const float inv_l = 1.0f / l_run;
    __half* orow = p.out + (long)batch * p.out_batch_stride + (long)row * p.ldo + head * kD;
    uint32_t oa[32], ob[32];
    tmem_ld32(t_o + 0, oa);
    tmem_ld32(t_o + 32, ob);
    tmem_ld_wait();
    if (row < p.lq) {
#define VG_ATTN_STORE(ARR, C0)                                                                                  \
  _Pragma("unroll") for (int g = 0; g < 4; ++g) {                                                               \
    uint4 u;                                                                                                    \
    u.x = pack_half2(__uint_as_float(ARR[g * 8 + 0]) * inv_l, __uint_as_float(ARR[g * 8 + 1]) * inv_l);         \
    u.y = pack_half2(__uint_as_float(ARR[g * 8 + 2]) * inv_l, __uint_as_float(ARR[g * 8 + 3]) * inv_l);         \
    u.z = pack_half2(__uint_as_float(ARR[g * 8 + 4]) * inv_l, __uint_as_float(ARR[g * 8 + 5]) * inv_l);         \
    u.w = pack_half2(__uint_as_float(ARR[g * 8 + 6]) * inv_l, __uint_as_float(ARR[g * 8 + 7]) * inv_l);         \
    *reinterpret_cast<uint4*>(orow + (C0) + g * 8) = u;                                                         \
  }
      VG_ATTN_STORE(oa, 0)
      VG_ATTN_STORE(ob, 32)
#undef VG_ATTN_STORE
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kProducerWarp + 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// "TS" family: P never touches shared memory.  The exponentials are packed to fp16 in registers and stored with
// tcgen05.st into the TMEM columns the block's scores came from; the PV product reads its A operand from tensor memory
// (TS-form tcgen05.mma).  Against the SS form above that removes, per 64-key block and tile, eight STS.128, the
// MEMBAR + generic->async proxy fence that publishes them (the costliest fixed item of the softmax warps' block, see the
// SASS), and two of the four mbarrier round trips: the scores are double-buffered in TMEM (2 x 64 columns per tile) and P(j)
// aliases S(j), so the tensor pipe's issue order -- PV(j), then QK(j+2) into the same buffer -- is the only hand-back the
// S buffers need.  Per tile: S/P buffers [128 i, 128 i + 128), O (+ the row-sum column) [128 kNT + 80 i, +80).
//   kNT = 2: one CTA per SM, 11 warps, K/V blocks shared by both tiles (416 TMEM columns)
//   kNT = 1: TWO CTAs per SM (256 TMEM columns and 82 KB of shared memory each), 6 warps: the two tiles of an SM are
//            independent CTAs whose prologues / epilogues overlap (the 145-key cross attention is all prologue): the default
//            for lk <= 4096
//   kNT = 1, ONE score buffer, row sums in the softmax warps ("TS3"): 128 TMEM columns, 49 KB, 96 registers: THREE CTAs per SM;
//            QK(j+1) then has to wait for PV(j), which costs the tile ~990 cycles per block -- ahead only on the 145-key launches
// Measured and rejected on top of this body (source: experiments/attn_ts_ffma2_fmnmx3_poly_earlyld.cu.txt; numbers for the 14080^2
// launch, TS1 / TS2, profiles/r02q_attn_*.log vs r02r_attn_ts_microopt.log): issuing the tcgen05.ld of S(j+1) in the middle of block
// j's exponentials into a second register set (10.48 -> 10.70 / 9.83 -> 10.38 ms: a warp's TMEM operations complete in order, so
// the early load only delays the P store queued behind it); FFMA2 for the scale-and-subtract, FMNMX3 for the row maximum, one
// register set and the next block's load queued before the store wait (10.48 -> 10.99 / 9.83 -> 13.06 ms); every 4th / 8th
// exponential as a Cody-Waite + degree-3 polynomial on the FMA pipe (11.2 / 11.5 ms and 11.7 / 12.0 ms on that base).
template <int kNT_, int kSBufs_ = 2, bool kTcSum_ = true, int kCtasPerSm_ = (kNT_ == 1 ? 2 : 1)>
struct AttnTsCfg {
  static constexpr int kNT = kNT_;
  static constexpr int kSBufs = kSBufs_;                // score buffers per tile (2: QK runs two blocks ahead; 1: one)
  static constexpr bool kTcSum = kTcSum_;               // row sums on the tensor core (16 more O columns) or in the softmax warps
  static constexpr int kTileK = 64;
  static constexpr int kThreads = 32 * (5 * kNT + 1);
  static constexpr int kKBytes = kTileK * kD * 2;       // 8 KB
  static constexpr int kStageLog2 = kSBufs == 2 ? 2 : 1, kKvStages = 1 << kStageLog2;
  static constexpr int kNumBars = 1 + 4 * kKvStages + 3 * kSBufs * kNT;
  static constexpr int kDataBytes = kNT * kQBytes + 2 * kKvStages * kKBytes + (kTcSum ? 2048 : 0);   // Q, K ring, V ring, ones atom
  static constexpr int kTmemNeed = kNT * (64 * kSBufs + (kTcSum ? 80 : 64));
  static constexpr uint32_t kTmemCols = kTmemNeed <= 128 ? 128 : kTmemNeed <= 256 ? 256 : 512;
  static constexpr int kCtasPerSm = kCtasPerSm_;
  static_assert(kTmemNeed <= 512 && kCtasPerSm * (int)kTmemCols <= 512, "TMEM columns");
  static constexpr size_t smem_bytes() { return (size_t)kDataBytes + kNumBars * 8 + 16 + 1024; }
  static_assert(kCtasPerSm * (kDataBytes + kNumBars * 8 + 16 + 1024 + 1024) <= 227 * 1024, "shared memory of the co-resident CTAs");
};

template <class Cfg, bool kTiming>
__device__ __forceinline__ void attn_sm100_ts_body(const AttnParams& p) {
  constexpr int kNT = Cfg::kNT, kTileK = Cfg::kTileK, kKBytes = Cfg::kKBytes;
  constexpr int kKvStages = Cfg::kKvStages, kStageLog2 = Cfg::kStageLog2;
  constexpr int kSBufs = Cfg::kSBufs;
  constexpr bool kTcSum = Cfg::kTcSum;
  constexpr int kOCols = kTcSum ? 80 : 64;
  constexpr int kTileCols = 64 * kSBufs;          // S / P columns of one q-tile
  constexpr int kOBase = kTileCols * kNT;
  // block j lives in score buffer sbuf(j); it is the (j / kSBufs)-th use of that buffer's barriers
  auto sbuf = [](int j) { return kSBufs == 2 ? (j & 1) : 0; };
  auto sphase = [](int j) { return (uint32_t)((kSBufs == 2 ? (j >> 1) : j) & 1); };
  constexpr int kProducerWarp = 4 * kNT;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // kNT x 16 KB
  uint8_t* sK = sQ + kNT * kQBytes;                    // ring of K blocks
  uint8_t* sV = sK + kKvStages * kKBytes;              // ring of V blocks
  uint8_t* sOnes = sV + kKvStages * kKBytes;           // kTcSum: 2 KB constant MN-major atom (16 keys x 64 columns, column 0 = 1)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + (kTcSum ? 2048 : 0));
  uint64_t* q_full = bars;                        // [1]
  uint64_t* k_full = q_full + 1;                  // [stages]
  uint64_t* k_empty = k_full + kKvStages;         // [stages] every tile's QK(j) finished
  uint64_t* v_full = k_empty + kKvStages;         // [stages]
  uint64_t* v_empty = v_full + kKvStages;         // [stages] every tile's PV(j) finished
  uint64_t* s_full = v_empty + kKvStages;         // [kNT][kSBufs] per q-tile and S buffer: S(j) written by QK(j)
  uint64_t* p_full = s_full + kSBufs * kNT;       // [kNT][kSBufs] per q-tile and S buffer: P(j) stored over S(j), O rescaled
  uint64_t* o_full = p_full + kSBufs * kNT;       // [kNT][kSBufs] per q-tile and S buffer: PV(j) finished (O readable)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + kSBufs * kNT);
  static_assert(1 + 4 * Cfg::kKvStages + 3 * Cfg::kSBufs * Cfg::kNT == Cfg::kNumBars, "barrier count");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qblk = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
  const int q0 = qblk * kNT * kTileQ;
  const int nkv = (p.lk + kTileK - 1) / kTileK;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&p.map_q);
    tma_prefetch_desc(&p.map_k);
    tma_prefetch_desc(&p.map_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKvStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], kNT);   // one commit per tile's issuer
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], kNT);
    }
    for (int i = 0; i < kSBufs * kNT; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == kProducerWarp + 1) {
    tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tmem_relinquish();
  }
  if (kTcSum && warp < 4) {
    // rows = keys (two 8-row groups 1024 B apart, like the V tile), 128 B per row, 128B-swizzled: the 16-byte piece holding
    // column 0 of row r sits at piece index (0 ^ (r & 7))
    const int t = threadIdx.x;                       // 128 threads x 16 B = 2 KB
    const int row = t >> 3, piece = t & 7;
    const uint32_t val = (piece == (row & 7)) ? 0x00003C00u : 0u;   // fp16 1.0 in the low half
    st_shared_v4(smem_u32(sOnes) + t * 16, val, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kProducerWarp) {
    // ------------------------------------------------------------------ TMA producer
    const int kvb = batch / p.kv_batch_div;
    if (elect_one()) {
      mbar_expect_tx(q_full, kNT * kQBytes);
#pragma unroll
      for (int i = 0; i < kNT; ++i) tma_load_4d(sQ + i * kQBytes, &p.map_q, q_full, 0, head, q0 + i * kTileQ, batch);
    }
    __syncwarp();
    for (int j = 0; j < nkv; ++j) {
      const int st = j & (kKvStages - 1);
      const uint32_t reuse_parity = ((j >> kStageLog2) - 1) & 1;
      if (j >= kKvStages) mbar_wait(&k_empty[st], reuse_parity, 30);
      if (elect_one()) {
        mbar_expect_tx(&k_full[st], kKBytes);
        tma_load_4d(sK + st * kKBytes, &p.map_k, &k_full[st], 0, head, j * kTileK, kvb);
      }
      __syncwarp();
      if (j >= kKvStages) mbar_wait(&v_empty[st], reuse_parity, 31);
      if (elect_one()) {
        mbar_expect_tx(&v_full[st], kKBytes);
        tma_load_4d(sV + st * kKBytes, &p.map_v, &v_full[st], 0, head, j * kTileK, kvb);
      }
      __syncwarp();
    }
  } else if (warp > kProducerWarp) {
    // ------------------------------------------------------------------ MMA issue for q-tile i (warp-uniform loop)
    // QK(0) [QK(1)] { PV(j) QK(j + kSBufs) } ...: the tensor pipe executes one thread's MMAs in issue order, so QK(j + kSBufs)
    // overwrites the S buffer only after PV(j) has read P(j) out of it.
    const int i = warp - kProducerWarp - 1;
    const uint32_t idesc_qk = umma_idesc_f16(kTileQ, kTileK, 0, 0);
    const uint32_t idesc_pv = umma_idesc_f16(kTileQ, kOCols, 0, 1);  // A = P (TMEM, K-major), B = V + the ones atom (MN-major)
    const uint32_t ones_addr = smem_u32(sOnes);
    const uint32_t q_addr = smem_u32(sQ) + i * kQBytes, k_addr = smem_u32(sK), v_addr0 = smem_u32(sV);
    const uint32_t t_s = tmem + kTileCols * i, t_o = tmem + kOBase + i * kOCols;
    auto issue_qk = [&](int j) {
      const int st = j & (kKvStages - 1);
      mbar_wait(&k_full[st], (j >> kStageLog2) & 1, 32);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t a_desc = umma_desc_sw128(q_addr, 16, 1024);
        const uint64_t b_desc = umma_desc_sw128(k_addr + st * kKBytes, 16, 1024);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) umma_f16_ss(t_s + 64 * sbuf(j), a_desc + 2 * k, b_desc + 2 * k, idesc_qk, k != 0);
        umma_commit(&s_full[kSBufs * i + sbuf(j)]);
        umma_commit(&k_empty[st]);  // this tile is done with K(j); the slot frees when every tile is
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0, 33);
    issue_qk(0);
    if (kSBufs == 2 && nkv > 1) issue_qk(1);
    for (int j = 0; j < nkv; ++j) {
      const int st = j & (kKvStages - 1);
      mbar_wait(&p_full[kSBufs * i + sbuf(j)], sphase(j), 34);
      mbar_wait(&v_full[st], (j >> kStageLog2) & 1, 35);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t v_addr = v_addr0 + st * kKBytes;
        const uint32_t a_tmem = t_s + 64 * sbuf(j);      // P(j): 64 keys = 32 columns of packed fp16 pairs
#pragma unroll
        for (int ks = 0; ks < kTileK / 16; ++ks) {
          // B: V rows [16*ks, 16*ks+16) (K dimension), 64 contiguous d (MN-major), 8-row groups 1024 B apart; the second
          // 64-column atom of B is the ones tile (leading byte offset = its distance from this K step)
          const uint32_t vk = v_addr + ks * 16 * 128;
          const uint64_t b_desc = umma_desc_sw128(vk, kTcSum ? ones_addr - vk : 1024u, 1024);
          umma_f16_ts(t_o, a_tmem + 8 * ks, b_desc, idesc_pv, (j | ks) != 0);  // O accumulates over blocks
        }
        umma_commit(&o_full[kSBufs * i + sbuf(j)]);
        umma_commit(&v_empty[st]);
      }
      __syncwarp();
      if (j + kSBufs < nkv) issue_qk(j + kSBufs);
    }
  } else {
    // ---------------------------------------------------------------- softmax / output warps
    const int i = warp >> 2;          // q-tile
    const int q = warp & 3;           // TMEM lane quadrant
    const int r = q * 32 + lane;      // row within the tile
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t t_s = tmem + kTileCols * i + lane_base;
    const uint32_t t_o = tmem + kOBase + i * kOCols + lane_base;
    const float sl2 = p.scale_log2;

    float m_ref = -INFINITY;  // reference max (raw score units) all of this row's exponentials are relative to
    float l_run = 0.f;        // !kTcSum: the row sum
    long long tm_wait_s = 0, tm_ldmax = 0, tm_exp = 0, tm_c0 = 0, tm_c1 = 0, tm_start = 0;
    if (kTiming) tm_start = clock64();

    auto mask_scores = [&](uint32_t (&s)[2][32], int valid) {
      if (valid < kTileK) {  // last, partial block: masked keys must not contribute
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            if (32 * c + t >= valid) s[c][t] = 0xff800000u;
          }
        }
      }
    };
    auto wait_and_load = [&](uint32_t (&s)[2][32], int j) {   // tcgen05.ld of S(j) issued, NOT waited for
      mbar_wait(&s_full[kSBufs * i + sbuf(j)], sphase(j), 40);
      tc_fence_after();
      tmem_ld32(t_s + 64 * sbuf(j), s[0]);
      tmem_ld32(t_s + 64 * sbuf(j) + 32, s[1]);
    };
    // exponentials of 32 scores -> 16 packed fp16 pairs -> TMEM columns [16 c, 16 c + 16) of the block's S buffer
    auto exp_store = [&](const uint32_t (&sc)[32], int c, float neg_ms, uint32_t t_p) {
      uint32_t pk[16];
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const float e0 = fast_exp2(fmaf(__uint_as_float(sc[2 * t]), sl2, neg_ms));
        const float e1 = fast_exp2(fmaf(__uint_as_float(sc[2 * t + 1]), sl2, neg_ms));
        if (!kTcSum) {
          l0 += e0;
          l1 += e1;
        }
        pk[t] = pack_half2(e0, e1);
      }
      tmem_st16(t_p + 16 * c, pk);
      if (!kTcSum) l_run += l0 + l1;
    };
    // One key block whose scores are in `cur` (loaded and waited for); the next block's scores are loaded into `nxt` at the end.
    auto step = [&](uint32_t (&cur)[2][32], uint32_t (&nxt)[2][32], int j) {
      if (kTiming) tm_c1 = clock64();
      mask_scores(cur, p.lk - j * kTileK);
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int t = 0; t < 32; t += 2) {
        mx0 = fmaxf(mx0, __uint_as_float(cur[0][t]));
        mx1 = fmaxf(mx1, __uint_as_float(cur[0][t + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(cur[1][t]));
        mx3 = fmaxf(mx3, __uint_as_float(cur[1][t + 1]));
      }
      const float m_blk = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      // advance the reference max only when the true max has outgrown it by more than the threshold
      const bool grow = (m_blk - m_ref) * sl2 > kRescaleThreshold;   // always true on the first block
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? m_blk : m_ref;
        if (j > 0) {
          const float alpha = fast_exp2((m_ref - m_new) * sl2);  // 1 for rows that do not advance
          // O must be complete up to PV(j-1) before it is touched
          mbar_wait(&o_full[kSBufs * i + sbuf(j - 1)], sphase(j - 1), 41);
          tc_fence_after();
          uint32_t o[16];
#pragma unroll 1
          for (int c = 0; c < kOCols; c += 16) {   // 64 output columns (+ kTcSum: the row-sum column and its 15 zero companions)
            tmem_ld16(t_o + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 16; ++t) o[t] = __float_as_uint(__uint_as_float(o[t]) * alpha);
            tmem_st16(t_o + c, o);
          }
          tmem_st_wait();
          l_run *= alpha;
        }
        m_ref = m_new;
      }
      if (kTiming) { tm_c0 = clock64(); tm_ldmax += tm_c0 - tm_c1; }
      const float neg_ms = -m_ref * sl2;
      const uint32_t t_p = t_s + 64 * sbuf(j);
      exp_store(cur[0], 0, neg_ms, t_p);
      exp_store(cur[1], 1, neg_ms, t_p);
      // P(j) stored (and O rescaled): publish to the MMA warp
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[kSBufs * i + sbuf(j)]);
      if (kTiming) { tm_c1 = clock64(); tm_exp += tm_c1 - tm_c0; }
      if (j + 1 < nkv) {
        wait_and_load(nxt, j + 1);
        tmem_ld_wait();
        reg_pin32(nxt[0]);
        reg_pin32(nxt[1]);
      }
      if (kTiming) { tm_c0 = clock64(); tm_wait_s += tm_c0 - tm_c1; }
    };

    uint32_t sa[2][32];
    wait_and_load(sa, 0);
    tmem_ld_wait();
    reg_pin32(sa[0]);
    reg_pin32(sa[1]);
    if constexpr (Cfg::kCtasPerSm >= 3) {
      // <= 96 registers per thread: one score register set (a block's scores are dead once its exponentials are packed)
      for (int j = 0; j < nkv; ++j) step(sa, sa, j);
    } else {
      uint32_t sb[2][32];   // alternating score register sets (measured faster than one set: 10.48 vs 10.99 ms, TS1 at 14080^2)
      for (int j = 0; j < nkv; j += 2) {
        step(sa, sb, j);
        if (j + 1 < nkv) step(sb, sa, j + 1);
      }
    }
    if (kTiming && p.timing && i < 2 && q == 0 && lane == 0 && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0) {
      long long* t = p.timing + (kNT == 1 ? 0 : i) * 8;
      t[0] = tm_wait_s, t[1] = tm_ldmax, t[2] = 0, t[3] = tm_exp, t[4] = clock64() - tm_start, t[5] = nkv;
    }
    // ---- output: O / l  (l = the row-sum column, accumulated by the tensor core)
    mbar_wait(&o_full[kSBufs * i + sbuf(nkv - 1)], sphase(nkv - 1), 42);
    tc_fence_after();
    const int row = q0 + i * kTileQ + r;
    uint32_t o16[16];
    if (kTcSum) tmem_ld16(t_o + kD, o16);
    uint32_t oa[32], ob[32];
    tmem_ld32(t_o + 0, oa);
    tmem_ld32(t_o + 32, ob);
    tmem_ld_wait();
    const float inv_l = 1.0f / (kTcSum ? __uint_as_float(o16[0]) : l_run);
    __half* orow = p.out + (long)batch * p.out_batch_stride + (long)row * p.ldo + head * kD;
    if (row < p.lq) {
#define VG_ATTN_STORE(ARR, C0)                                                                                  \
  _Pragma("unroll") for (int g = 0; g < 4; ++g) {                                                               \
    uint4 u;                                                                                                    \
    u.x = pack_half2(__uint_as_float(ARR[g * 8 + 0]) * inv_l, __uint_as_float(ARR[g * 8 + 1]) * inv_l);         \
    u.y = pack_half2(__uint_as_float(ARR[g * 8 + 2]) * inv_l, __uint_as_float(ARR[g * 8 + 3]) * inv_l);         \
    u.z = pack_half2(__uint_as_float(ARR[g * 8 + 4]) * inv_l, __uint_as_float(ARR[g * 8 + 5]) * inv_l);         \
    u.w = pack_half2(__uint_as_float(ARR[g * 8 + 6]) * inv_l, __uint_as_float(ARR[g * 8 + 7]) * inv_l);         \
    *reinterpret_cast<uint4*>(orow + (C0) + g * 8) = u;                                                         \
  }
      VG_ATTN_STORE(oa, 0)
      VG_ATTN_STORE(ob, 32)
#undef VG_ATTN_STORE
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kProducerWarp + 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem);
  }
}

using AttnTs1 = AttnTsCfg<1>;
using AttnTs2 = AttnTsCfg<2>;
// one score buffer, row sums in the softmax warps: 128 TMEM columns and 49 KB of shared memory per CTA, THREE CTAs per SM
// (three softmax warps per sub-partition without doubling the per-key barrier traffic)
using AttnTs3 = AttnTsCfg<1, 1, false, 3>;
#define VG_ATTN_TS_KERNEL(NAME, CFG, TIMING)                                                                 \
  __global__ void __launch_bounds__(CFG::kThreads, CFG::kCtasPerSm) NAME(const __grid_constant__ AttnParams p) { \
    attn_sm100_ts_body<CFG, TIMING>(p);                                                                      \
  }
VG_ATTN_TS_KERNEL(attn_sm100_ts1_kernel, AttnTs1, false)
VG_ATTN_TS_KERNEL(attn_sm100_ts2_kernel, AttnTs2, false)
VG_ATTN_TS_KERNEL(attn_sm100_ts3_kernel, AttnTs3, false)
VG_ATTN_TS_KERNEL(attn_sm100_ts1_timing_kernel, AttnTs1, true)
VG_ATTN_TS_KERNEL(attn_sm100_ts2_timing_kernel, AttnTs2, true)
VG_ATTN_TS_KERNEL(attn_sm100_ts3_timing_kernel, AttnTs3, true)
#undef VG_ATTN_TS_KERNEL

using AttnCfg2 = AttnCfg<2, 128>;
using AttnCfg3 = AttnCfg<3, 64>;

__global__ void __launch_bounds__(AttnCfg2::kThreads, 1) attn_sm100_kernel(const __grid_constant__ AttnParams p) {
  attn_sm100_body<AttnCfg2, false, false>(p);
}
__global__ void __launch_bounds__(AttnCfg2::kThreads, 1) attn_sm100_tcsum_kernel(const __grid_constant__ AttnParams p) {
  attn_sm100_body<AttnCfg2, false, true>(p);
}
// three q-tiles, 64-key blocks (always with the tensor-core row sums)
__global__ void __launch_bounds__(AttnCfg3::kThreads, 1) attn_sm100_x3_kernel(const __grid_constant__ AttnParams p) {
  attn_sm100_body<AttnCfg3, false, true>(p);
}
// instrumented twins (phase cycle counters); only tools/bench_attn.py launches them (vgen_attention_d64_debug)
__global__ void __launch_bounds__(AttnCfg2::kThreads, 1) attn_sm100_timing_kernel(const __grid_constant__ AttnParams p) {
  if (p.tcsum) attn_sm100_body<AttnCfg2, true, true>(p);
  else attn_sm100_body<AttnCfg2, true, false>(p);
}
__global__ void __launch_bounds__(AttnCfg3::kThreads, 1) attn_sm100_x3_timing_kernel(const __grid_constant__ AttnParams p) {
  attn_sm100_body<AttnCfg3, true, true>(p);
}

}  // namespace vg

using namespace vg;

// Which instantiation serves a shape when VGEN_ATTN_TILES is not set (internal codes: 2 = SS <2 tiles, 128 keys>, 3 = SS <3, 64>,
// 10 / 20 / 30 = TS1 / TS2 / TS3).  Measured on B200 (profiles/r02q_attn_*.log; TS3: r02t / r02u_attn_ts3.log: 11.32, 1.525, 0.246,
// 0.0442, 0.334, 0.176 -- ahead only on the two cross-attention rows), ms:
//   lq x lk        SS<2,128>  SS<3,64>   TS1      TS2
//   14080 x 14080   9.70       9.79     10.48     9.83
//    3520 x 3520    1.446      1.396     1.401    1.406
//     880 x 880     0.308      0.277     0.2335   0.293
//     220 x 220     0.0545     0.0562    0.0428   0.0487
//   14080 x 145     0.553      0.385     0.3725   0.450
//    3520 x 145     0.284      0.204     0.190    0.230
// Short key sequences are all prologue / epilogue: two small CTAs per SM overlap them.  Long ones are bound by the MUFU pipe
// at 73-78 % utilisation in every variant; the 128-key blocks of the SS kernel amortise the per-block barriers best.
static int attn_default_tiles(int64_t lq, int64_t lk) {
  (void)lq;
  return lk <= 4096 ? 10 : 2;
}

static int attention_d64_impl(const void* q, const void* k, const void* v, void* out, int64_t batch, int64_t heads,
                              int64_t lq, int64_t lk, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                              int64_t kv_batch_div, float scale, void* stream, long long* timing) {
  VG_REQUIRE(q && k && v && out, "vgen_attention_d64: null pointer");
  VG_REQUIRE(batch >= 0 && heads > 0 && lq > 0 && lk > 0 && kv_batch_div >= 1 && batch % kv_batch_div == 0,
             "vgen_attention_d64: bad shape");
  VG_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "vgen_attention_d64: strides must be multiples of 8");
  VG_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(v) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "vgen_attention_d64: pointers must be 16-byte aligned");
  VG_REQUIRE(heads <= 65535 && batch <= 65535, "vgen_attention_d64: grid too large");
  if (batch == 0) return 0;
  // Variant: VGEN_ATTN_TILES=2 -> SS <2 tiles, 128-key blocks>, 3 -> SS <3 tiles, 64-key blocks>, t1 / t2 -> the TS family (1 or
  // 2 tiles per CTA); unset: by shape.
  // VGEN_ATTN_TCSUM=0 keeps the row sums in the softmax warps (SS <2,128> only); VGEN_ATTN_STAGGER=0/1 forces the start-up
  // phase offset between the tiles of SS <3,64> off / on.  A/B knobs.
  int tiles_mode, tcsum_mode, stagger_mode;   // (getenv per call: ~100 ns against a >= 50 us kernel; lets one process A/B)
  {
    const char* e = getenv("VGEN_ATTN_TILES");
    tiles_mode = (e && e[0] == '3') ? 3 : (e && e[0] == '2') ? 2 : 0;
    if (e && e[0] == 't' && (e[1] == '1' || e[1] == '2' || e[1] == '3')) tiles_mode = 10 * (e[1] - '0');
    e = getenv("VGEN_ATTN_TCSUM");
    tcsum_mode = (e && e[0] == '0') ? 0 : 1;
    e = getenv("VGEN_ATTN_STAGGER");
    stagger_mode = e ? (e[0] != '0') : -1;
  }
  const int tiles = tiles_mode ? tiles_mode : attn_default_tiles(lq, lk);
  const int tile_k = tiles == 2 ? AttnCfg2::kTileK : 64;
  AttnParams p;
  // dims ordered so that the byte strides ascend: {d, head, token, batch}
  const uint32_t box[4] = {64, 1, 128, 1};
  const uint32_t box_kv[4] = {64, 1, (uint32_t)tile_k, 1};
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
    int rc = make_tmap_f16(&p.map_k, k, 4, dims, str, box_kv);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {64, (uint64_t)heads, (uint64_t)lk, bkv};
    const uint64_t str[3] = {128, (uint64_t)ldv * 2, (uint64_t)lk * ldv * 2};
    int rc = make_tmap_f16(&p.map_v, v, 4, dims, str, box_kv);
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
  p.timing = timing;
  p.tcsum = tiles != 2 ? 1 : tcsum_mode;
  p.stagger = stagger_mode >= 0 ? stagger_mode : (tiles == 3 ? 1 : 0);
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    VG_CUDA(cudaFuncSetAttribute(attn_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VG_CUDA(cudaFuncSetAttribute(attn_sm100_tcsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VG_CUDA(cudaFuncSetAttribute(attn_sm100_timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VG_CUDA(cudaFuncSetAttribute(attn_sm100_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VG_CUDA(cudaFuncSetAttribute(attn_sm100_x3_timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
#define VG_ATTN_TS_ATTR(K, CFG) VG_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CFG::smem_bytes()))
    VG_ATTN_TS_ATTR(attn_sm100_ts1_kernel, AttnTs1);
    VG_ATTN_TS_ATTR(attn_sm100_ts2_kernel, AttnTs2);
    VG_ATTN_TS_ATTR(attn_sm100_ts1_timing_kernel, AttnTs1);
    VG_ATTN_TS_ATTR(attn_sm100_ts2_timing_kernel, AttnTs2);
    VG_ATTN_TS_ATTR(attn_sm100_ts3_kernel, AttnTs3);
    VG_ATTN_TS_ATTR(attn_sm100_ts3_timing_kernel, AttnTs3);
#undef VG_ATTN_TS_ATTR
    attr_once.mark();
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (tiles >= 10) {
    if (tiles / 10 == 1) {
      dim3 grid((unsigned)cdiv(lq, kTileQ), (unsigned)heads, (unsigned)batch);
      launch_kernel(timing ? attn_sm100_ts1_timing_kernel : attn_sm100_ts1_kernel, dim3(grid), dim3(AttnTs1::kThreads),
                    AttnTs1::smem_bytes(), st, p);
    } else if (tiles / 10 == 3) {
      dim3 grid((unsigned)cdiv(lq, kTileQ), (unsigned)heads, (unsigned)batch);
      launch_kernel(timing ? attn_sm100_ts3_timing_kernel : attn_sm100_ts3_kernel, dim3(grid), dim3(AttnTs3::kThreads),
                    AttnTs3::smem_bytes(), st, p);
    } else {
      dim3 grid((unsigned)cdiv(lq, 2 * kTileQ), (unsigned)heads, (unsigned)batch);
      launch_kernel(timing ? attn_sm100_ts2_timing_kernel : attn_sm100_ts2_kernel, dim3(grid), dim3(AttnTs2::kThreads),
                    AttnTs2::smem_bytes(), st, p);
    }
  } else if (tiles == 3) {
    const size_t smem = AttnCfg3::smem_bytes(true);
    dim3 grid((unsigned)cdiv(lq, AttnCfg3::kNT * kTileQ), (unsigned)heads, (unsigned)batch);
    if (timing) launch_kernel(attn_sm100_x3_timing_kernel, dim3(grid), dim3(AttnCfg3::kThreads), smem, st, p);
    else launch_kernel(attn_sm100_x3_kernel, dim3(grid), dim3(AttnCfg3::kThreads), smem, st, p);
  } else {
    const size_t smem = AttnCfg2::smem_bytes(tcsum_mode != 0);
    dim3 grid((unsigned)cdiv(lq, AttnCfg2::kNT * kTileQ), (unsigned)heads, (unsigned)batch);
    if (timing) launch_kernel(attn_sm100_timing_kernel, dim3(grid), dim3(AttnCfg2::kThreads), smem, st, p);
    else if (tcsum_mode) launch_kernel(attn_sm100_tcsum_kernel, dim3(grid), dim3(AttnCfg2::kThreads), smem, st, p);
    else launch_kernel(attn_sm100_kernel, dim3(grid), dim3(AttnCfg2::kThreads), smem, st, p);
  }
  VG_LAUNCH_CHECK("attn_sm100_kernel");
  return 0;
}

extern "C" int vgen_attention_d64(const void* q, const void* k, const void* v, void* out, int64_t batch, int64_t heads,
                                  int64_t lq, int64_t lk, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                                  int64_t kv_batch_div, float scale, void* stream) {
  return attention_d64_impl(q, k, v, out, batch, heads, lq, lk, ldq, ldk, ldv, ldo, kv_batch_div, scale, stream, nullptr);
}

// Diagnosis entry (tools/bench_attn.py): the instrumented twin with per-phase cycle counters of one mid-grid CTA
// (timing16: 16 int64, per q-tile {wait S, load+max, wait P buffer, exponentials+store, total, blocks, -, -}).
extern "C" int vgen_attention_d64_debug(const void* q, const void* k, const void* v, void* out, int64_t batch, int64_t heads,
                                        int64_t lq, int64_t lk, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                                        int64_t kv_batch_div, float scale, long long* timing16, void* stream) {
  VG_REQUIRE(timing16, "vgen_attention_d64_debug: timing16 must not be null");
  return attention_d64_impl(q, k, v, out, batch, heads, lq, lk, ldq, ldk, ldv, ldo, kv_batch_div, scale, stream, timing16);
}
