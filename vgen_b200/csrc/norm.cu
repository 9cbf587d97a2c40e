// GroupNorm(32) (+SiLU) and LayerNorm on channels-last fp16 activations: HBM-bound kernels with
// 16-byte vectorised, fully coalesced accesses, several independent loads in flight per thread and
// fp32 statistics.
//
// Reference ops replaced:
//   nn.GroupNorm(32, C) + nn.SiLU in ResBlock in_layers/out_layers   tools/modules/unet/util.py:845-876
//   GroupNorm over 5-D [b,c,f,h,w] (statistics span ALL frames)       util.py:1248 (TemporalTransformer),
//                                                                      util.py:1662-1680 (TemporalConvBlock_v2)
//   GroupNorm eps=1e-6 in SpatialTransformer / VAE                     util.py:329, autoencoder.py:15-16
//   nn.LayerNorm in BasicTransformerBlock                              util.py:694-696
// Under the reference's autocast these run in fp32 and the result is rounded to fp16 once by the
// consuming conv/linear; the kernels below do the same (fp32 maths, one fp16 rounding at the store).
//
// GroupNorm is three launches:
//   gn_stats    grid (splits, N): each thread owns fixed channel vectors (no atomics) and accumulates
//               per-channel fp32 sum / sum-of-squares over its pixels, 4 loads in flight; folded to
//               per-group (mean, M2) partials.  The split count scales with the tensor so that even the
//               N = 1 "all frames jointly" norms fill the 148 SMs.
//   gn_finalize grid (N): one warp per group merges the partials with Chan's formula (robust to
//               mean >> std) -> (mean, rstd).
//   gn_apply    same thread->channel ownership, scale/shift in registers: y = silu?(x*scale + shift).
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace vg {

static constexpr int kGroups = 32;
static constexpr int kMaxSplits = 1024;

struct GnGeom {
  int C, C8, cpg;      // channels, C/8, channels per group
  int vpt;             // channel vectors per thread
  int cols;            // thread columns (= ceil(C8 / vpt))
  int rows;            // pixel rows processed per block iteration
  long P;              // positions per sample
  int splits;
  long chunk;          // positions per split
};

__device__ __forceinline__ void acc8(const uint4& u, float (&s)[8], float (&q)[8]) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h2[i]);
    s[2 * i] += f.x;
    q[2 * i] = fmaf(f.x, f.x, q[2 * i]);
    s[2 * i + 1] += f.y;
    q[2 * i + 1] = fmaf(f.y, f.y, q[2 * i + 1]);
  }
}

template <int VPT>
__global__ void __launch_bounds__(256) gn_stats_kernel(const __half* __restrict__ x, float2* __restrict__ partial, GnGeom g) {
  extern __shared__ float sm[];  // [2][rows][C]
  const int n = blockIdx.y, sp = blockIdx.x;
  const int tid = threadIdx.x;
  const int r = tid / g.cols, cv = tid - r * g.cols;
  const long p0 = (long)sp * g.chunk;
  long p1 = p0 + g.chunk;
  if (p1 > g.P) p1 = g.P;
  float s[VPT][8], q[VPT][8];
#pragma unroll
  for (int j = 0; j < VPT; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) s[j][i] = q[j][i] = 0.f;

  if (r < g.rows) {
    const __half* base = x + ((long)n * g.P) * g.C;
    const long step = g.rows;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int v = cv + j * g.cols;
      if (v >= g.C8) continue;
      const __half* col = base + v * 8;
      long p = p0 + r;
      for (; p + 3 * step < p1; p += 4 * step) {  // 4 independent 16-byte loads in flight
        const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(col + p * g.C));
        const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(col + (p + step) * g.C));
        const uint4 u2 = __ldg(reinterpret_cast<const uint4*>(col + (p + 2 * step) * g.C));
        const uint4 u3 = __ldg(reinterpret_cast<const uint4*>(col + (p + 3 * step) * g.C));
        acc8(u0, s[j], q[j]);
        acc8(u1, s[j], q[j]);
        acc8(u2, s[j], q[j]);
        acc8(u3, s[j], q[j]);
      }
      for (; p < p1; p += step) {
        const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(col + p * g.C));
        acc8(u0, s[j], q[j]);
      }
    }
    float* ss = sm + (long)r * g.C;
    float* qq = sm + (long)(g.rows + r) * g.C;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int v = cv + j * g.cols;
      if (v < g.C8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          ss[v * 8 + i] = s[j][i];
          qq[v * 8 + i] = q[j][i];
        }
      }
    }
  }
  __syncthreads();
  if (tid < kGroups) {
    float S = 0.f, Q = 0.f;
    for (int rr = 0; rr < g.rows; ++rr) {
      const float* ss = sm + (long)rr * g.C + tid * g.cpg;
      const float* qq = sm + (long)(g.rows + rr) * g.C + tid * g.cpg;
      for (int c = 0; c < g.cpg; ++c) {
        S += ss[c];
        Q += qq[c];
      }
    }
    const float cnt = (float)((p1 > p0 ? p1 - p0 : 0) * g.cpg);
    const float mean = cnt > 0 ? S / cnt : 0.f;
    float m2 = Q - S * mean;  // sum (x-mean)^2
    if (m2 < 0.f) m2 = 0.f;
    partial[((long)n * g.splits + sp) * kGroups + tid] = make_float2(mean, m2);
  }
}

// one warp per group: lanes stride over the splits, then a shuffle tree of Chan merges
__global__ void __launch_bounds__(1024) gn_finalize_kernel(const float2* __restrict__ partial, float2* __restrict__ stats,
                                                           GnGeom g, float eps) {
  const int n = blockIdx.x;
  const int grp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float cnt = 0.f, mean = 0.f, m2 = 0.f;
  const float2* pbase = partial + (long)n * g.splits * kGroups + grp;
  for (int sp0 = lane; sp0 < g.splits; sp0 += 32 * 8) {
    float2 pm[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {  // 8 independent loads in flight, then the (serial) Chan merges
      const int sp = sp0 + u * 32;
      pm[u] = sp < g.splits ? __ldg(pbase + (long)sp * kGroups) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int sp = sp0 + u * 32;
      if (sp >= g.splits) continue;
      const long q0 = (long)sp * g.chunk;
      long q1 = q0 + g.chunk;
      if (q1 > g.P) q1 = g.P;
      if (q1 <= q0) continue;
      const float cb = (float)((q1 - q0) * g.cpg);
      const float tot = cnt + cb;
      const float delta = pm[u].x - mean;
      mean += delta * (cb / tot);
      m2 += pm[u].y + delta * delta * (cnt * cb / tot);
      cnt = tot;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float cb = __shfl_xor_sync(0xffffffffu, cnt, o);
    const float mb = __shfl_xor_sync(0xffffffffu, mean, o);
    const float qb = __shfl_xor_sync(0xffffffffu, m2, o);
    const float tot = cnt + cb;
    if (tot > 0.f) {
      const float delta = mb - mean;
      mean += delta * (cb / tot);
      m2 += qb + delta * delta * (cnt * cb / tot);
      cnt = tot;
    }
  }
  if (lane == 0) stats[(long)n * kGroups + grp] = make_float2(mean, rsqrtf(m2 / cnt + eps));
}

// x * sigmoid(x); the quotient uses the approximate reciprocal (<= 1 ulp, the result is rounded to fp16): the
// IEEE division sequence (~10 instructions) made gn_apply issue-bound (ncu r01k: 65% issue, 67% XU at 3.5 TB/s)
__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }
__device__ __forceinline__ uint4 norm8(const uint4& u, const float (&sc)[8], const float (&sh)[8], int silu) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&u);
  float f[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = __half22float2(h2[k]);
    f[2 * k] = fmaf(t.x, sc[2 * k], sh[2 * k]);
    f[2 * k + 1] = fmaf(t.y, sc[2 * k + 1], sh[2 * k + 1]);
  }
  if (silu) {
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = silu_f(f[k]);
  }
  uint4 o;
  o.x = pack_half2(f[0], f[1]);
  o.y = pack_half2(f[2], f[3]);
  o.z = pack_half2(f[4], f[5]);
  o.w = pack_half2(f[6], f[7]);
  return o;
}

template <int VPT>
__global__ void __launch_bounds__(256) gn_apply_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                       const float2* __restrict__ stats, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, GnGeom g, int silu, long apply_chunk) {
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int r = tid / g.cols, cv = tid - r * g.cols;
  if (r >= g.rows) return;
  const long p0 = (long)blockIdx.x * apply_chunk;
  long p1 = p0 + apply_chunk;
  if (p1 > g.P) p1 = g.P;
  const long step = g.rows;
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const int v = cv + j * g.cols;
    if (v >= g.C8) continue;
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = v * 8 + k;
      const float2 st = __ldg(stats + (long)n * kGroups + c / g.cpg);
      sc[k] = st.y * __ldg(gamma + c);
      sh[k] = __ldg(beta + c) - st.x * sc[k];
    }
    const __half* xc = x + ((long)n * g.P) * g.C + v * 8;
    __half* yc = y + ((long)n * g.P) * g.C + v * 8;
    long p = p0 + r;
    for (; p + 3 * step < p1; p += 4 * step) {
      const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(xc + p * g.C));
      const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(xc + (p + step) * g.C));
      const uint4 u2 = __ldg(reinterpret_cast<const uint4*>(xc + (p + 2 * step) * g.C));
      const uint4 u3 = __ldg(reinterpret_cast<const uint4*>(xc + (p + 3 * step) * g.C));
      *reinterpret_cast<uint4*>(yc + p * g.C) = norm8(u0, sc, sh, silu);
      *reinterpret_cast<uint4*>(yc + (p + step) * g.C) = norm8(u1, sc, sh, silu);
      *reinterpret_cast<uint4*>(yc + (p + 2 * step) * g.C) = norm8(u2, sc, sh, silu);
      *reinterpret_cast<uint4*>(yc + (p + 3 * step) * g.C) = norm8(u3, sc, sh, silu);
    }
    for (; p < p1; p += step) {
      const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(xc + p * g.C));
      *reinterpret_cast<uint4*>(yc + p * g.C) = norm8(u0, sc, sh, silu);
    }
  }
}

static int gn_geometry(long N, long P, int C, GnGeom* g) {
  VG_REQUIRE(C % 32 == 0 && C % 8 == 0, "group_norm: C must be a multiple of 32");
  g->C = C;
  g->C8 = C / 8;
  g->cpg = C / kGroups;
  g->vpt = (g->C8 + 255) / 256;
  VG_REQUIRE(g->vpt <= 2, "group_norm: C > 4096 not supported");
  g->cols = (g->C8 + g->vpt - 1) / g->vpt;
  g->rows = 256 / g->cols;
  if (g->rows < 1) g->rows = 1;
  g->P = P;
  long want = (8L * sm_count()) / (N > 0 ? N : 1);  // ~8 blocks per SM over the whole launch
  if (want < 1) want = 1;
  if (want > kMaxSplits) want = kMaxSplits;
  long min_chunk = 8L * g->rows;  // at least two unrolled iterations per thread
  long chunk = (P + want - 1) / want;
  if (chunk < min_chunk) chunk = min_chunk;
  g->chunk = chunk;
  g->splits = (int)((P + chunk - 1) / chunk);
  return 0;
}

// ------------------------------------------------------------------------------- LayerNorm
// One warp per row, the row lives in registers (C <= 2560), two-pass mean / variance like ATen.  Warps
// walk rows with a grid stride and prefetch the next row's raw vectors before reducing the current one,
// so every warp always has loads in flight.
// kStats: nothing is written but the row's statistics {rstd, -mean * rstd} (vgen_row_stats: the LayerNorm itself is folded
// into the GEMM that consumes the row, see vgen_epilogue.row_stats).
template <int MAXV, bool kStats>
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        long rows, int C, long ldx, long ldy, float eps,
                                                        float2* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int C8 = C >> 3;
  const long warps_total = (long)gridDim.x * (blockDim.x >> 5);
  long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  uint4 nxt[MAXV];
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int v = lane + j * 32;
    if (v < C8) nxt[j] = __ldg(reinterpret_cast<const uint4*>(x + row * ldx + v * 8));
  }
  for (; row < rows; row += warps_total) {
    float f[MAXV][8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int v = lane + j * 32;
      if (v < C8) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&nxt[j]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 t = __half22float2(h2[k]);
          f[j][2 * k] = t.x;
          f[j][2 * k + 1] = t.y;
          s += t.x + t.y;
        }
      }
    }
    const long nrow = row + warps_total;
    if (nrow < rows) {
#pragma unroll
      for (int j = 0; j < MAXV; ++j) {
        const int v = lane + j * 32;
        if (v < C8) nxt[j] = __ldg(reinterpret_cast<const uint4*>(x + nrow * ldx + v * 8));
      }
    }
    s = warp_sum(s);
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int v = lane + j * 32;
      if (v < C8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float d = f[j][k] - mean;
          q = fmaf(d, d, q);
        }
      }
    }
    q = warp_sum(q);
    const float rstd = rsqrtf(q / (float)C + eps);
    if (kStats) {
      if (lane == 0) stats[row] = make_float2(rstd, -mean * rstd);
      continue;
    }
    __half* yr = y + row * ldy;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int v = lane + j * 32;
      if (v < C8) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaf((f[j][k] - mean) * rstd, gg[k], bb[k]);
        uint4 u;
        u.x = pack_half2(o[0], o[1]);
        u.y = pack_half2(o[2], o[3]);
        u.z = pack_half2(o[4], o[5]);
        u.w = pack_half2(o[6], o[7]);
        *reinterpret_cast<uint4*>(yr + v * 8) = u;
      }
    }
  }
}

// read-only 16-byte load the compiler may not hoist out of a loop (ptxas kept 80 loop-invariant gamma / beta values in
// registers and spilled; the lines stay L1-resident, re-loading them per row group is cheaper than the lost occupancy)
__device__ __forceinline__ float4 ldg_f4_pinned(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// Rows narrower than a warp's worth of 16-byte vectors (C = 320 / 640: 40 / 80 vectors): LPR lanes share a row and a
// warp normalises 32/LPR rows at once, so every load instruction is fully populated and each lane keeps VPL (= 5)
// 16-byte loads in flight (the warp-per-row kernel had 1.25 per lane at C = 320 and ran at 3.9 TB/s).
template <int LPR, int VPL, bool kStats>
__global__ void __launch_bounds__(256) layernorm_grouped_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                long rows, int C, long ldx, long ldy, float eps,
                                                                float2* __restrict__ stats) {
  constexpr int RW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, li = lane % LPR;
  const long warps_total = (long)gridDim.x * (blockDim.x >> 5);
  long grp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long ngroups = (rows + RW - 1) / RW;
  if (grp >= ngroups) return;
  uint4 nxt[VPL];
  {
    const long row = grp * RW + sub;
    if (row < rows) {
#pragma unroll
      for (int j = 0; j < VPL; ++j) nxt[j] = __ldg(reinterpret_cast<const uint4*>(x + row * ldx + (li + j * LPR) * 8));
    }
  }
  for (; grp < ngroups; grp += warps_total) {
    const long row = grp * RW + sub;
    const bool live = row < rows;
    uint4 cur[VPL];  // the row stays packed (fp16) in registers; each pass converts on the fly
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      cur[j] = live ? nxt[j] : make_uint4(0u, 0u, 0u, 0u);
      const __half2* h2 = reinterpret_cast<const __half2*>(&cur[j]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 t = __half22float2(h2[k]);
        s += t.x + t.y;
      }
    }
    const long nrow = (grp + warps_total) * RW + sub;
    if (grp + warps_total < ngroups && nrow < rows) {
#pragma unroll
      for (int j = 0; j < VPL; ++j) nxt[j] = __ldg(reinterpret_cast<const uint4*>(x + nrow * ldx + (li + j * LPR) * 8));
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const __half2* h2 = reinterpret_cast<const __half2*>(&cur[j]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 t = __half22float2(h2[k]);
        const float d0 = t.x - mean, d1 = t.y - mean;
        q = fmaf(d0, d0, q);
        q = fmaf(d1, d1, q);
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)C + eps);
    if (kStats) {
      if (live && li == 0) stats[row] = make_float2(rstd, -mean * rstd);
      continue;
    }
    if (live) {
      __half* yr = y + row * ldy;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int v = li + j * LPR;
        const float4 g0 = ldg_f4_pinned(gamma + v * 8);
        const float4 g1 = ldg_f4_pinned(gamma + v * 8 + 4);
        const float4 b0 = ldg_f4_pinned(beta + v * 8);
        const float4 b1 = ldg_f4_pinned(beta + v * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const __half2* h2 = reinterpret_cast<const __half2*>(&cur[j]);
        float o[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 t = __half22float2(h2[k]);
          o[2 * k] = fmaf((t.x - mean) * rstd, gg[2 * k], bb[2 * k]);
          o[2 * k + 1] = fmaf((t.y - mean) * rstd, gg[2 * k + 1], bb[2 * k + 1]);
        }
        uint4 u;
        u.x = pack_half2(o[0], o[1]);
        u.y = pack_half2(o[2], o[3]);
        u.z = pack_half2(o[4], o[5]);
        u.w = pack_half2(o[6], o[7]);
        *reinterpret_cast<uint4*>(yr + v * 8) = u;
      }
    }
  }
}

// generic small-C LayerNorm (any C, scalar): one thread per row
template <bool kStats>
__global__ void layernorm_small_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                       const float* __restrict__ gamma, const float* __restrict__ beta, long rows, int C,
                                       long ldx, long ldy, float eps, float2* __restrict__ stats) {
  const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const __half* xr = x + row * ldx;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += __half2float(xr[c]);
  const float mean = s / C;
  float q = 0.f;
  for (int c = 0; c < C; ++c) {
    const float d = __half2float(xr[c]) - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(q / C + eps);
  if (kStats) {
    stats[row] = make_float2(rstd, -mean * rstd);
  } else {
    for (int c = 0; c < C; ++c) y[row * ldy + c] = __float2half_rn((__half2float(xr[c]) - mean) * rstd * gamma[c] + beta[c]);
  }
}

// vgen_layer_norm (kStats = false) and vgen_row_stats (kStats = true) share the kernel choice.
template <bool kStats>
static int layer_norm_launch(const void* x, void* y, int64_t rows, int64_t c, int64_t ldx, int64_t ldy, const float* gamma,
                             const float* beta, float eps, float2* stats, void* stream) {
  if (rows == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __half* xp = reinterpret_cast<const __half*>(x);
  __half* yp = reinterpret_cast<__half*>(y);
  const bool vec = (c % 8 == 0) && (ldx % 8 == 0) && (ldy % 8 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(y) & 15) == 0) && c <= 2560 &&
                   ((reinterpret_cast<uintptr_t>(gamma) & 15) == 0) && ((reinterpret_cast<uintptr_t>(beta) & 15) == 0);
  if (vec) {
    const int wpb = 8;
    long blocks = (rows + wpb - 1) / wpb;
    const long cap = 16L * sm_count();  // grid-stride: ~16 resident blocks' worth per SM
    if (blocks > cap) blocks = cap;
    const int c8 = (int)c / 8;
    if (c8 == 40 || c8 == 80) {
      const int rw = c8 == 40 ? 4 : 2;  // rows per warp
      long gblocks = ((rows + rw - 1) / rw + wpb - 1) / wpb;
      if (gblocks > cap) gblocks = cap;
      if (c8 == 40)
        launch_kernel(layernorm_grouped_kernel<8, 5, kStats>, dim3((unsigned)gblocks), dim3(32 * wpb), 0, st, xp, yp, gamma, beta, rows, (int)c, ldx, ldy, eps, stats);
      else
        launch_kernel(layernorm_grouped_kernel<16, 5, kStats>, dim3((unsigned)gblocks), dim3(32 * wpb), 0, st, xp, yp, gamma, beta, rows, (int)c, ldx, ldy, eps, stats);
      VG_LAUNCH_CHECK("layernorm_grouped_kernel");
      return 0;
    }
    if (c8 <= 32)
      launch_kernel(layernorm_kernel<1, kStats>, dim3((unsigned)blocks), dim3(32 * wpb), 0, st, xp, yp, gamma, beta, rows, (int)c, ldx, ldy, eps, stats);
    else if (c8 <= 64)
      launch_kernel(layernorm_kernel<2, kStats>, dim3((unsigned)blocks), dim3(32 * wpb), 0, st, xp, yp, gamma, beta, rows, (int)c, ldx, ldy, eps, stats);
    else if (c8 <= 160)
      launch_kernel(layernorm_kernel<5, kStats>, dim3((unsigned)blocks), dim3(32 * wpb), 0, st, xp, yp, gamma, beta, rows, (int)c, ldx, ldy, eps, stats);
    else
      launch_kernel(layernorm_kernel<10, kStats>, dim3((unsigned)blocks), dim3(32 * wpb), 0, st, xp, yp, gamma, beta, rows, (int)c, ldx, ldy, eps, stats);
    VG_LAUNCH_CHECK("layernorm_kernel");
  } else {
    const unsigned blocks = (unsigned)((rows + 127) / 128);
    launch_kernel(layernorm_small_kernel<kStats>, dim3(blocks), dim3(128), 0, st, xp, yp, gamma, beta, rows, (int)c, ldx, ldy, eps, stats);
    VG_LAUNCH_CHECK("layernorm_small_kernel");
  }
  return 0;
}

}  // namespace vg

using namespace vg;

extern "C" {

// workspace: [n][kMaxSplits][32] partials + [n][32] final (mean, rstd)
int64_t vgen_group_norm_workspace_bytes(int64_t n) {
  const int64_t nn = n < 1 ? 1 : n;
  return nn * (int64_t)(kMaxSplits + 1) * kGroups * (int64_t)sizeof(float2);
}

int vgen_group_norm(const void* x, void* y, int64_t n, int64_t p, int64_t c, const float* gamma, const float* beta,
                    float eps, int silu, void* workspace, void* stream) {
  VG_REQUIRE(x && y && gamma && beta && workspace, "vgen_group_norm: null pointer");
  VG_REQUIRE(n >= 0 && p > 0 && c > 0, "vgen_group_norm: bad shape");
  if (n == 0) return 0;
  VG_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
             "vgen_group_norm: x/y must be 16-byte aligned");
  GnGeom g;
  int rc = gn_geometry(n, p, (int)c, &g);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float2* part = reinterpret_cast<float2*>(workspace);
  float2* stats = part + (long)n * kMaxSplits * kGroups;
  const size_t smem_stats = (size_t)2 * g.rows * g.C * sizeof(float);
  VG_REQUIRE(smem_stats <= 48 * 1024, "vgen_group_norm: stats smem too large");
  dim3 grid_s(g.splits, (unsigned)n);
  const int threads = g.rows * g.cols < 32 ? 32 : g.rows * g.cols;
  if (g.vpt == 1)
    launch_kernel(gn_stats_kernel<1>, dim3(grid_s), dim3(threads), smem_stats, st, reinterpret_cast<const __half*>(x), part, g);
  else
    launch_kernel(gn_stats_kernel<2>, dim3(grid_s), dim3(threads), smem_stats, st, reinterpret_cast<const __half*>(x), part, g);
  VG_LAUNCH_CHECK("gn_stats_kernel");
  launch_kernel(gn_finalize_kernel, dim3((unsigned)n), dim3(1024), 0, st, part, stats, g, eps);
  VG_LAUNCH_CHECK("gn_finalize_kernel");
  long blocks = (8L * sm_count()) / n;
  if (blocks < 1) blocks = 1;
  long apply_chunk = (p + blocks - 1) / blocks;
  if (apply_chunk < 8L * g.rows) apply_chunk = 8L * g.rows;
  blocks = (p + apply_chunk - 1) / apply_chunk;
  dim3 grid_a((unsigned)blocks, (unsigned)n);
  if (g.vpt == 1)
    launch_kernel(gn_apply_kernel<1>, dim3(grid_a), dim3(threads), 0, st, reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y), stats, gamma,
                                                   beta, g, silu, apply_chunk);
  else
    launch_kernel(gn_apply_kernel<2>, dim3(grid_a), dim3(threads), 0, st, reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y), stats, gamma,
                                                   beta, g, silu, apply_chunk);
  VG_LAUNCH_CHECK("gn_apply_kernel");
  return 0;
}

int vgen_layer_norm(const void* x, void* y, int64_t rows, int64_t c, int64_t ldx, int64_t ldy, const float* gamma,
                    const float* beta, float eps, void* stream) {
  VG_REQUIRE(x && y && gamma && beta, "vgen_layer_norm: null pointer");
  VG_REQUIRE(rows >= 0 && c > 0 && ldx >= c && ldy >= c, "vgen_layer_norm: bad shape");
  return layer_norm_launch<false>(x, y, rows, c, ldx, ldy, gamma, beta, eps, nullptr, stream);
}

int vgen_row_stats(const void* x, int64_t rows, int64_t c, int64_t ldx, float eps, float* stats, void* stream) {
  VG_REQUIRE(x && stats, "vgen_row_stats: null pointer");
  VG_REQUIRE(rows >= 0 && c > 0 && ldx >= c, "vgen_row_stats: bad shape");
  VG_REQUIRE((reinterpret_cast<uintptr_t>(stats) & 7) == 0, "vgen_row_stats: stats must be 8-byte aligned");
  return layer_norm_launch<true>(x, nullptr, rows, c, ldx, c, nullptr, nullptr, eps, reinterpret_cast<float2*>(stats), stream);
}

}  // extern "C"
