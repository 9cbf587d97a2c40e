// Epilogue shared by the 1-CTA and 2-CTA tap-GEMM kernels: one thread owns one accumulator row (TMEM
// lane), reads it 32 columns at a time with tcgen05.ld and applies alpha, bias, per-frame bias (ResBlock
// "h + emb_out"), residual, or GEGLU, rounding to fp16 exactly where the reference materialises an fp16
// tensor.  Two warps share a TMEM lane quadrant and interleave the 32-column chunks; the four warps that
// work on the same chunk form a "column group" (128 threads = the 128 rows of the tile).
//
// Stores.  A thread-per-row epilogue writes 64 contiguous bytes per thread, i.e. every warp store touches
// 32 different 128-byte lines: the LSU retires ~16 B/clk/SM, which caps a short-K layer (k=320, n=320)
// at ~2.9 TB/s of the 6.5 TB/s HBM can deliver.  So the chunk is written to a 64B-swizzled staging tile in
// shared memory (conflict-free 16-byte stores) and leaves the SM as ONE TMA tile store per (chunk, group):
// rows x 64 bytes, clipped by the TMA unit at the tensor edges.  Two staging tiles per group ring so the
// store of chunk i overlaps the math of chunk i+1.  (An earlier variant that re-read the staging tile with
// ld.shared and issued coalesced st.global from the same warps was 1.4-1.6x slower than direct stores; the
// TMA store has no second pass through the register file.)  The direct path remains for outputs the TMA
// cannot address (row stride or base not 16-byte aligned).
#pragma once
#include "ptx.cuh"
#include "tapgemm.h"

namespace vg {

// exact-erf GELU (F.gelu default, util.py:714) with erf from Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, far below
// the fp16 rounding the result receives).  Written so that no cancellation and no sign fix-up is needed:
//   c = 1 - erf(|x|/sqrt2) = poly(t) * t * exp(-x^2/2),  t = 1 / (1 + p |x|/sqrt2)
//   gelu(x) = max(x, 0) - (|x|/2) * c
// and exp(-z^2) = ex2(-u^2) with u = z * sqrt(log2 e): 2 MUFU (rcp, ex2, both .ftz: no denormal fix-up code) + 13
// FMA-pipe instructions (libm erff: ~30; the straightforward 0.5x(1+erf) form with __expf: ~20).
__device__ __forceinline__ float gelu_erf(float x) {
#ifdef VG_GELU_LIBM
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
#endif
  constexpr float kU = 0.70710678118654752f * 1.2011224087864498f;   // 1/sqrt2 * sqrt(log2 e)
  constexpr float kP = 0.3275911f / 1.2011224087864498f;             // p / sqrt(log2 e)
  const float ax = fabsf(x);
  const float u = ax * kU;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(u, kP, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float c = poly * t * fast_exp2(-u * u);
  return fmaxf(x, 0.f) - (0.5f * ax) * c;
}

static constexpr int kEpiStageBytes = 128 * 64;              // one staging tile: 128 rows x 32 fp16
static constexpr int kEpiStagingTotal = 4 * kEpiStageBytes;  // 2 column groups x 2 ring slots

struct EpiRow {
  int nb_i;        // n-block index of the tile
  int i3;          // outermost coordinate (frame) of the tile
  int i1_0, i2_0;  // origin of the tile along d1 / d2 (TMA store coordinates)
  long row;        // output row of this thread (TMEM lane)
  bool row_ok;     // that row is inside the tensor
  uint32_t t_row;  // TMEM address of the thread's lane, column 0 of the accumulator buffer
};

struct EpiStore {
  bool tma;                // TMA tile stores (else direct per-thread stores)
  const CUtensorMap* map;  // output [out_n, d1, d2, d3], box {32, box1, box2, 1}, 64B swizzle
  uint32_t stage;          // shared address of this column group's two staging tiles
  uint32_t row_off;        // r * 64: this thread's row inside a staging tile
  uint32_t row_xor;        // (r >> 1) & 3: 64B-swizzle term of that row
  int bar_id;              // named barrier of the column group (128 threads)
  bool leader;             // the one thread of the group that issues / tracks the bulk stores
  uint32_t slot;           // ring position, carried across chunks and tiles
  // residual tiles arrive by TMA as well (same box / swizzle as the output), one chunk ahead of their use
  bool res_tma;
  const CUtensorMap* map_res;
  uint32_t rstage;         // shared address of this group's two residual tiles
  uint64_t* rbar;          // [2] full barriers of those tiles
  uint32_t box_bytes;      // bytes one residual tile load delivers (box1 * box2 * 64)
  uint32_t r_issue, r_cons;  // loads issued / consumed so far (slot = count & 1, phase = (count >> 1) & 1)
};

__device__ __forceinline__ bool tapgemm_vec_ok(const TapGemmEpilogue& e, int out_n) {
  return ((reinterpret_cast<uintptr_t>(e.bias) & 15) == 0) && ((e.ldo & 7) == 0) && ((out_n & 7) == 0) &&
         ((reinterpret_cast<uintptr_t>(e.out) & 15) == 0) && ((reinterpret_cast<uintptr_t>(e.group_bias) & 15) == 0) &&
         ((e.ld_group_bias & 7) == 0) &&
         (e.residual == nullptr || (((e.ldr & 7) == 0) && ((reinterpret_cast<uintptr_t>(e.residual) & 15) == 0)));
}

// host side: can the output be written with TMA tile stores?
inline bool tapgemm_out_tma_ok(const TapGemmEpilogue& e) {
  return (reinterpret_cast<uintptr_t>(e.out) & 15) == 0 && (e.ldo & 7) == 0;
}

// ---- staging-tile protocol of one column group -------------------------------------------------------
// acquire: the slot written two chunks ago must have been read by its TMA store
__device__ __forceinline__ void epi_stage_acquire(EpiStore& st) {
  if (st.leader) bulk_wait_group_read<1>();
  named_bar_sync(st.bar_id, 128);
}
// 16-byte piece j8 (0..3) of this thread's row
__device__ __forceinline__ void epi_stage_put(const EpiStore& st, int j8, const uint4& o) {
  const uint32_t addr = st.stage + st.slot * kEpiStageBytes + st.row_off + ((((uint32_t)j8) ^ st.row_xor) << 4);
  st_shared_v4(addr, o.x, o.y, o.z, o.w);
}
// publish: make the generic-proxy writes visible to the TMA unit, then one thread stores the tile
__device__ __forceinline__ void epi_stage_publish(EpiStore& st, int col, const EpiRow& t) {
  fence_proxy_async_smem();
  named_bar_sync(st.bar_id, 128);
  if (st.leader) {
    tma_store_4d(st.map, st.stage + st.slot * kEpiStageBytes, col, t.i1_0, t.i2_0, t.i3);
    bulk_commit_group();
  }
  st.slot ^= 1;
}
// residual tile for the chunk at column `col` of tile t: issued by the group leader one chunk ahead (or, for a tile's
// first chunk, before the accumulator-ready wait); every thread counts the issue so the slot / phase stay in step
__device__ __forceinline__ void epi_res_issue(EpiStore& st, int col, const EpiRow& t) {
  if (st.leader) {
    uint64_t* bar = st.rbar + (st.r_issue & 1);
    mbar_expect_tx(bar, st.box_bytes);
    tma_load_4d(reinterpret_cast<void*>(0), st.map_res, bar, col, t.i1_0, t.i2_0, t.i3, st.rstage + (st.r_issue & 1) * kEpiStageBytes);
  }
  ++st.r_issue;
}
// wait for the oldest outstanding residual tile and fetch piece j8 of this thread's row
__device__ __forceinline__ void epi_res_wait(const EpiStore& st) {
  mbar_wait(st.rbar + (st.r_cons & 1), (st.r_cons >> 1) & 1, 35);
}
__device__ __forceinline__ uint4 epi_res_get(const EpiStore& st, int j8) {
  const uint32_t addr = st.rstage + (st.r_cons & 1) * kEpiStageBytes + st.row_off + ((((uint32_t)j8) ^ st.row_xor) << 4);
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// before the kernel exits every bulk store must have completed
__device__ __forceinline__ void epi_stage_drain(const EpiStore& st) {
  if (st.tma && st.leader) bulk_wait_group<0>();
}

// chunk0 / chunk_step: this warp handles the 32-column chunks chunk0, chunk0 + chunk_step, ... (two warps share
// a TMEM lane quadrant and split the columns between them).
// kGeglu is a template parameter of the kernels: the two epilogues have very different register needs, and
// compiled into one kernel each paid for the other's allocation and schedule.
// kLn: LayerNorm folded into the GEMM (TapGemmEpilogue.row_stats / col_sum): the accumulator row is scaled by the row's rstd and
// shifted by (-mean * rstd) * col_sum[n]; such launches carry no residual / per-frame bias, whose registers the column sums use.
template <bool kGeglu, bool kLn>
__device__ __forceinline__ void tapgemm_epilogue_tile(const TapGemmShape& s, const TapGemmEpilogue& e, const EpiRow& t,
                                                      EpiStore& st, bool vec_ok, int out_n, int chunk0, int chunk_step) {
  const int BN = s.bn;
  if constexpr (!kGeglu) {
    const int n0 = t.nb_i * BN;
    __half* orow = e.out + t.row * e.ldo;
    const __half* rrow = e.residual ? e.residual + t.row * e.ldr : nullptr;
    const __half* grow = e.group_bias ? e.group_bias + (long)fd_div(e.f_group_bias_div, t.i3) * e.ld_group_bias : nullptr;
    float ln_scale = e.alpha, ln_shift = 0.f;   // kLn: rstd (* alpha) and -mean * rstd of this thread's row
    if (kLn && t.row_ok) {
      const float2 rs = __ldg(e.row_stats + t.row);
      ln_scale = rs.x * e.alpha;
      ln_shift = rs.y * e.alpha;
    }
    for (int c0 = chunk0 * 32; c0 < BN; c0 += chunk_step * 32) {
      const int nbase = n0 + c0;
      if (nbase >= s.n) break;  // uniform over the column group
      uint32_t v[32];
      tmem_ld32(t.t_row + c0, v);
      const bool full = nbase + 32 <= s.n;
      const bool fast = vec_ok && t.row_ok && full;
      // issue the (HBM / L2 latency) residual and per-frame-bias loads before blocking on the TMEM load
      uint4 r4[4], g4[4];
      float4 b4[8], c4[8];
      const bool staged = st.tma && vec_ok && full;  // uniform over the column group
      const bool res_tile = st.res_tma && staged;    // this chunk's residual was requested by TMA
      if (res_tile) {
        // request the NEXT chunk of this tile (its slot was last read two chunks ago, before a publish barrier)
        const int nnext = nbase + chunk_step * 32;
        if (c0 + chunk_step * 32 < BN && nnext + 32 <= s.n) epi_res_issue(st, nnext, t);
      }
      if (fast) {
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          if (!kLn && rrow && !res_tile) r4[j8] = __ldg(reinterpret_cast<const uint4*>(rrow + nbase + j8 * 8));
          if (!kLn && grow) g4[j8] = __ldg(reinterpret_cast<const uint4*>(grow + nbase + j8 * 8));
          if (kLn) {  // col_sum is 16-byte aligned (checked by the launcher), nbase a multiple of 32
            c4[2 * j8] = __ldg(reinterpret_cast<const float4*>(e.col_sum + nbase + j8 * 8));
            c4[2 * j8 + 1] = __ldg(reinterpret_cast<const float4*>(e.col_sum + nbase + j8 * 8 + 4));
          }
          if (e.bias) {  // vec_ok implies n % 8 == 0; bias tensors are 16-byte aligned
            b4[2 * j8] = __ldg(reinterpret_cast<const float4*>(e.bias + nbase + j8 * 8));
            b4[2 * j8 + 1] = __ldg(reinterpret_cast<const float4*>(e.bias + nbase + j8 * 8 + 4));
          }
        }
      }
      if (staged) epi_stage_acquire(st);
      if (res_tile) {
        epi_res_wait(st);
        if (fast) {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) r4[j8] = epi_res_get(st, j8);
        }
        ++st.r_cons;
      }
      tmem_ld_wait();
      if (fast) {
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[j8 * 8 + j]) * ln_scale;
          if (e.bias) {
            const float4 b0 = b4[2 * j8], b1 = b4[2 * j8 + 1];
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
            f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
          }
          if (kLn) {
            const float4 c0v = c4[2 * j8], c1v = c4[2 * j8 + 1];
            f[0] = fmaf(ln_shift, c0v.x, f[0]); f[1] = fmaf(ln_shift, c0v.y, f[1]);
            f[2] = fmaf(ln_shift, c0v.z, f[2]); f[3] = fmaf(ln_shift, c0v.w, f[3]);
            f[4] = fmaf(ln_shift, c1v.x, f[4]); f[5] = fmaf(ln_shift, c1v.y, f[5]);
            f[6] = fmaf(ln_shift, c1v.z, f[6]); f[7] = fmaf(ln_shift, c1v.w, f[7]);
          }
          if (!kLn && grow) {
            // reference: h (fp16 conv output) + emb_out (fp16) -> fp16  (util.py:909-919)
            const __half* gh = reinterpret_cast<const __half*>(&g4[j8]);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __half2float(__float2half_rn(f[j])) + __half2float(gh[j]);
          }
          if (!kLn && rrow) {
            const __half* rh = reinterpret_cast<const __half*>(&r4[j8]);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __half2float(__float2half_rn(f[j])) + __half2float(rh[j]);
          }
          uint4 o;
          o.x = pack_half2(f[0], f[1]);
          o.y = pack_half2(f[2], f[3]);
          o.z = pack_half2(f[4], f[5]);
          o.w = pack_half2(f[6], f[7]);
          if (staged) epi_stage_put(st, j8, o);
          else *reinterpret_cast<uint4*>(orow + nbase + j8 * 8) = o;
        }
      } else if (t.row_ok) {
        for (int j = 0; j < 32; ++j) {
          const int n = nbase + j;
          if (n >= s.n) break;
          float a = __uint_as_float(v[j]) * ln_scale;
          if (e.bias) a += e.bias[n];
          if (kLn) a = fmaf(ln_shift, e.col_sum[n], a);
          if (grow) a = __half2float(__float2half_rn(a)) + __half2float(grow[n]);
          if (rrow) a = __half2float(__float2half_rn(a)) + __half2float(rrow[n]);
          orow[n] = __float2half_rn(a);
        }
      }
      if (staged) epi_stage_publish(st, nbase, t);  // rows outside the tensor are clipped by the TMA unit
    }
  } else {
    // GEGLU: columns [0,BN/2) of this tile are "value" j, columns [BN/2,BN) the matching "gate" j
    // (host interleaves the weight rows per BN block).  out = value * gelu(gate)   (util.py:707-714)
    const int hb = BN >> 1;
    const int o0 = t.nb_i * hb;
    __half* orow = e.out + t.row * e.ldo;
    float ln_scale = e.alpha, ln_shift = 0.f;
    if (kLn && t.row_ok) {
      const float2 rs = __ldg(e.row_stats + t.row);
      ln_scale = rs.x * e.alpha;
      ln_shift = rs.y * e.alpha;
    }
    for (int c0 = chunk0 * 32; c0 < hb; c0 += chunk_step * 32) {
      const int obase = o0 + c0;
      if (obase >= out_n) break;  // uniform over the column group
      uint32_t v[32], g[32];
      tmem_ld32(t.t_row + c0, v);
      tmem_ld32(t.t_row + hb + c0, g);
      const bool full = obase + 32 <= out_n;
      const bool staged = st.tma && vec_ok && full;
      if (staged) epi_stage_acquire(st);
      tmem_ld_wait();
      if (t.row_ok) {
        const int wbase = t.nb_i * BN + c0;  // packed weight-row index of value j (gate is + hb)
        // reference: the projection is an fp16 tensor, gelu(gate) an fp16 tensor, their product an fp16 tensor.
        // Pairs are kept packed: value and gelu(gate) are rounded by the f32x2 -> f16x2 pack, the product is one
        // HMUL2 (an fp16 x fp16 product is exact in fp32, so rounding it once is what the reference computes).
        uint32_t o[16];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), bg = bv;
          if (e.bias) {  // packed bias is 16-byte aligned and wbase is a multiple of 32
            bv = __ldg(reinterpret_cast<const float4*>(e.bias + wbase + j4 * 4));
            bg = __ldg(reinterpret_cast<const float4*>(e.bias + wbase + hb + j4 * 4));
          }
          if (kLn) {  // shift = bias + (-mean rstd) * col_sum, per packed weight row
            const float4 cv = __ldg(reinterpret_cast<const float4*>(e.col_sum + wbase + j4 * 4));
            const float4 cg = __ldg(reinterpret_cast<const float4*>(e.col_sum + wbase + hb + j4 * 4));
            bv.x = fmaf(ln_shift, cv.x, bv.x); bv.y = fmaf(ln_shift, cv.y, bv.y);
            bv.z = fmaf(ln_shift, cv.z, bv.z); bv.w = fmaf(ln_shift, cv.w, bv.w);
            bg.x = fmaf(ln_shift, cg.x, bg.x); bg.y = fmaf(ln_shift, cg.y, bg.y);
            bg.z = fmaf(ln_shift, cg.z, bg.z); bg.w = fmaf(ln_shift, cg.w, bg.w);
          }
          const float bva[4] = {bv.x, bv.y, bv.z, bv.w}, bga[4] = {bg.x, bg.y, bg.z, bg.w};
#pragma unroll
          for (int jj = 0; jj < 4; jj += 2) {
            const int j = j4 * 4 + jj;
            const __half2 a2 = __floats2half2_rn(fmaf(__uint_as_float(v[j]), ln_scale, bva[jj]),
                                                 fmaf(__uint_as_float(v[j + 1]), ln_scale, bva[jj + 1]));
            const float2 b2 = __half22float2(__floats2half2_rn(fmaf(__uint_as_float(g[j]), ln_scale, bga[jj]),
                                                               fmaf(__uint_as_float(g[j + 1]), ln_scale, bga[jj + 1])));
            const __half2 o2 = __hmul2(a2, __floats2half2_rn(gelu_erf(b2.x), gelu_erf(b2.y)));
            o[j >> 1] = *reinterpret_cast<const uint32_t*>(&o2);
          }
        }
        if (vec_ok && full) {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            const uint4 u4 = make_uint4(o[j8 * 4 + 0], o[j8 * 4 + 1], o[j8 * 4 + 2], o[j8 * 4 + 3]);
            if (staged) epi_stage_put(st, j8, u4);
            else *reinterpret_cast<uint4*>(orow + obase + j8 * 8) = u4;
          }
        } else {
          const __half* oh = reinterpret_cast<const __half*>(o);
          for (int j = 0; j < 32 && obase + j < out_n; ++j) orow[obase + j] = oh[j];
        }
      }
      if (staged) epi_stage_publish(st, obase, t);
    }
  }
}

}  // namespace vg
