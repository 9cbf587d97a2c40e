// Epilogue shared by the 1-CTA and 2-CTA tap-GEMM kernels: one thread owns one accumulator row (TMEM
// lane), reads it 32 columns at a time with tcgen05.ld and applies alpha, bias, per-frame bias (ResBlock
// "h + emb_out"), residual, or GEGLU, rounding to fp16 exactly where the reference materialises an fp16
// tensor.  Two warps share a TMEM lane quadrant and interleave the 32-column chunks.
// (A variant that transposed through shared memory for fully coalesced stores was measured 1.4-1.6x SLOWER on
// the short-K layers - the epilogue is latency-bound, not store-bound - and was removed; see DESIGN.md.)
#pragma once
#include "ptx.cuh"
#include "tapgemm.h"

namespace vg {

// exact-erf GELU (F.gelu default, util.py:714) with erf from Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, far below
// the fp16 rounding the result receives): 2 MUFU + ~14 FMA-pipe instructions instead of libm erff's ~30.
__device__ __forceinline__ float gelu_erf(float x) {
#ifdef VG_GELU_LIBM
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
#endif
  const float z = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.0f - poly * t * __expf(-z * z);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

struct EpiRow {
  int nb_i;        // n-block index of the tile
  int i3;          // outermost coordinate (frame) of the tile
  long row;        // output row of this thread (TMEM lane)
  bool row_ok;     // that row is inside the tensor
  uint32_t t_row;  // TMEM address of the thread's lane, column 0 of the accumulator buffer
};

__device__ __forceinline__ bool tapgemm_vec_ok(const TapGemmEpilogue& e, int out_n) {
  return ((reinterpret_cast<uintptr_t>(e.bias) & 15) == 0) && ((e.ldo & 7) == 0) && ((out_n & 7) == 0) &&
         ((reinterpret_cast<uintptr_t>(e.out) & 15) == 0) && ((reinterpret_cast<uintptr_t>(e.group_bias) & 15) == 0) &&
         ((e.ld_group_bias & 7) == 0) &&
         (e.residual == nullptr || (((e.ldr & 7) == 0) && ((reinterpret_cast<uintptr_t>(e.residual) & 15) == 0)));
}

// chunk0 / chunk_step: this warp handles the 32-column chunks chunk0, chunk0 + chunk_step, ... (two warps share
// a TMEM lane quadrant and split the columns between them).
__device__ __forceinline__ void tapgemm_epilogue_tile(const TapGemmShape& s, const TapGemmEpilogue& e, const EpiRow& t,
                                                      bool vec_ok, int out_n, int chunk0, int chunk_step) {
  const int BN = s.bn;
  if (!e.geglu) {
    const int n0 = t.nb_i * BN;
    __half* orow = e.out + t.row * e.ldo;
    const __half* rrow = e.residual ? e.residual + t.row * e.ldr : nullptr;
    const __half* grow = e.group_bias ? e.group_bias + (long)(t.i3 / e.group_bias_div) * e.ld_group_bias : nullptr;
    for (int c0 = chunk0 * 32; c0 < BN; c0 += chunk_step * 32) {
      uint32_t v[32];
      tmem_ld32(t.t_row + c0, v);
      const int nbase = n0 + c0;
      const bool fast = vec_ok && t.row_ok && nbase + 32 <= s.n;
      // issue the (HBM / L2 latency) residual and per-frame-bias loads before blocking on the TMEM load
      uint4 r4[4], g4[4];
      if (fast) {
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          if (rrow) r4[j8] = __ldg(reinterpret_cast<const uint4*>(rrow + nbase + j8 * 8));
          if (grow) g4[j8] = __ldg(reinterpret_cast<const uint4*>(grow + nbase + j8 * 8));
        }
      }
      tmem_ld_wait();
      if (!t.row_ok) continue;
      if (nbase >= s.n) continue;
      if (fast) {
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[j8 * 8 + j]) * e.alpha;
          if (e.bias) {  // vec_ok implies n % 8 == 0; bias tensors are 16-byte aligned
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(e.bias + nbase + j8 * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(e.bias + nbase + j8 * 8 + 4));
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
            f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
          }
          if (grow) {
            // reference: h (fp16 conv output) + emb_out (fp16) -> fp16  (util.py:909-919)
            const __half* gh = reinterpret_cast<const __half*>(&g4[j8]);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __half2float(__float2half_rn(f[j])) + __half2float(gh[j]);
          }
          if (rrow) {
            const __half* rh = reinterpret_cast<const __half*>(&r4[j8]);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __half2float(__float2half_rn(f[j])) + __half2float(rh[j]);
          }
          uint4 o;
          o.x = pack_half2(f[0], f[1]);
          o.y = pack_half2(f[2], f[3]);
          o.z = pack_half2(f[4], f[5]);
          o.w = pack_half2(f[6], f[7]);
          *reinterpret_cast<uint4*>(orow + nbase + j8 * 8) = o;
        }
      } else {
        for (int j = 0; j < 32; ++j) {
          const int n = nbase + j;
          if (n >= s.n) break;
          float a = __uint_as_float(v[j]) * e.alpha;
          if (e.bias) a += e.bias[n];
          if (grow) a = __half2float(__float2half_rn(a)) + __half2float(grow[n]);
          if (rrow) a = __half2float(__float2half_rn(a)) + __half2float(rrow[n]);
          orow[n] = __float2half_rn(a);
        }
      }
    }
  } else {
    // GEGLU: columns [0,BN/2) of this tile are "value" j, columns [BN/2,BN) the matching "gate" j
    // (host interleaves the weight rows per BN block).  out = value * gelu(gate)   (util.py:707-714)
    const int hb = BN >> 1;
    const int o0 = t.nb_i * hb;
    __half* orow = e.out + t.row * e.ldo;
    for (int c0 = chunk0 * 32; c0 < hb; c0 += chunk_step * 32) {
      uint32_t v[32], g[32];
      tmem_ld32(t.t_row + c0, v);
      tmem_ld32(t.t_row + hb + c0, g);
      tmem_ld_wait();
      if (!t.row_ok) continue;
      const int obase = o0 + c0;
      if (obase >= out_n) continue;
      const int wbase = t.nb_i * BN + c0;  // packed weight-row index of value j (gate is + hb)
      float f[32];
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), bg = bv;
        if (e.bias) {  // packed bias is 16-byte aligned and wbase is a multiple of 32
          bv = __ldg(reinterpret_cast<const float4*>(e.bias + wbase + j4 * 4));
          bg = __ldg(reinterpret_cast<const float4*>(e.bias + wbase + hb + j4 * 4));
        }
        const float bva[4] = {bv.x, bv.y, bv.z, bv.w}, bga[4] = {bg.x, bg.y, bg.z, bg.w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = j4 * 4 + jj;
          const float a = fmaf(__uint_as_float(v[j]), e.alpha, bva[jj]);
          const float b = fmaf(__uint_as_float(g[j]), e.alpha, bga[jj]);
          // reference rounds the projection to fp16, gelu to fp16, product to fp16
          const float a16 = __half2float(__float2half_rn(a));
          const float b16 = __half2float(__float2half_rn(b));
          const float ge = __half2float(__float2half_rn(gelu_erf(b16)));
          f[j] = a16 * ge;
        }
      }
      if (vec_ok && obase + 32 <= out_n) {
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          uint4 o;
          o.x = pack_half2(f[j8 * 8 + 0], f[j8 * 8 + 1]);
          o.y = pack_half2(f[j8 * 8 + 2], f[j8 * 8 + 3]);
          o.z = pack_half2(f[j8 * 8 + 4], f[j8 * 8 + 5]);
          o.w = pack_half2(f[j8 * 8 + 6], f[j8 * 8 + 7]);
          *reinterpret_cast<uint4*>(orow + obase + j8 * 8) = o;
        }
      } else {
        for (int j = 0; j < 32 && obase + j < out_n; ++j) orow[obase + j] = __float2half_rn(f[j]);
      }
    }
  }
}

}  // namespace vg
