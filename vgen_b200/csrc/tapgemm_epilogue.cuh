// Epilogue shared by the 1-CTA and 2-CTA tap-GEMM kernels.
//
// Accumulators live in TMEM with one row per lane, so the natural read-back gives every thread one
// OUTPUT ROW: storing from that layout makes each warp-level store touch 32 different rows (32 x 16 B
// scattered pieces), which costs ~8x the LSU cycles of a coalesced store and made the short-K layers
// (K = 320 linears at 225 280 rows) epilogue-bound (profiles/r01c_shapes.json).  So each epilogue warp
// transposes through a small padded shared-memory tile:
//   1. tcgen05.ld 64 accumulator columns of the thread's row; alpha, bias, per-frame bias (ResBlock
//      "h + emb_out"), GEGLU; round to fp16 (where the reference materialises an fp16 tensor); write
//      the 128-byte row segment to the warp's staging tile (pitch 144 B: conflict-free both ways);
//   2. re-read the tile so that 8 consecutive lanes cover one row segment (4 rows per instruction), add
//      the residual (coalesced 16-byte loads) and store coalesced 16-byte pieces.
// Two warps share a TMEM lane quadrant and split the 64-column super-chunks between them.
#pragma once
#include "ptx.cuh"
#include "tapgemm.h"

namespace vg {

static constexpr int kEpiPitch = 144;                       // bytes per staged row (128 + 16 pad)
static constexpr int kEpiStageBytes = 32 * kEpiPitch;       // per epilogue warp
static constexpr int kEpiWarps = 8;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

struct EpiRow {
  int nb_i;        // n-block index of the tile
  int i3;          // outermost coordinate (frame) of the tile
  long row;        // output row of this thread (TMEM lane)
  bool row_ok;     // that row is inside the tensor
  uint32_t t_row;  // TMEM address of the thread's lane, column 0 of the accumulator buffer
  // geometry of the tile, for the transposed read-out (lane -> 8 different rows of the warp's 32)
  int t1_i, t2_i;  // tile indices along d1 / d2
  int q;           // TMEM lane quadrant (rows q*32 .. q*32+31 of the tile)
};

__device__ __forceinline__ bool tapgemm_vec_ok(const TapGemmEpilogue& e, int out_n) {
  return ((reinterpret_cast<uintptr_t>(e.bias) & 15) == 0) && ((e.ldo & 7) == 0) && ((out_n & 7) == 0) &&
         ((reinterpret_cast<uintptr_t>(e.out) & 15) == 0) && ((reinterpret_cast<uintptr_t>(e.group_bias) & 15) == 0) &&
         ((e.ld_group_bias & 7) == 0) &&
         (e.residual == nullptr || (((e.ldr & 7) == 0) && ((reinterpret_cast<uintptr_t>(e.residual) & 15) == 0)));
}

// scalar path (tiny / unaligned N): direct per-row stores
__device__ __forceinline__ void tapgemm_epilogue_scalar(const TapGemmShape& s, const TapGemmEpilogue& e, const EpiRow& t,
                                                        int out_n, int chunk0, int chunk_step) {
  const int BN = s.bn;
  if (!e.geglu) {
    const int n0 = t.nb_i * BN;
    __half* orow = e.out + t.row * e.ldo;
    const __half* rrow = e.residual ? e.residual + t.row * e.ldr : nullptr;
    const __half* grow = e.group_bias ? e.group_bias + (long)(t.i3 / e.group_bias_div) * e.ld_group_bias : nullptr;
    for (int c0 = chunk0 * 32; c0 < BN; c0 += chunk_step * 32) {
      uint32_t v[32];
      tmem_ld32(t.t_row + c0, v);
      tmem_ld_wait();
      if (!t.row_ok) continue;
      for (int j = 0; j < 32; ++j) {
        const int n = n0 + c0 + j;
        if (n >= s.n) break;
        float a = __uint_as_float(v[j]) * e.alpha;
        if (e.bias) a += e.bias[n];
        if (grow) a = __half2float(__float2half_rn(a)) + __half2float(grow[n]);
        if (rrow) a = __half2float(__float2half_rn(a)) + __half2float(rrow[n]);
        orow[n] = __float2half_rn(a);
      }
    }
  } else {
    const int hb = BN >> 1;
    const int o0 = t.nb_i * hb;
    __half* orow = e.out + t.row * e.ldo;
    for (int c0 = chunk0 * 32; c0 < hb; c0 += chunk_step * 32) {
      uint32_t v[32], g[32];
      tmem_ld32(t.t_row + c0, v);
      tmem_ld32(t.t_row + hb + c0, g);
      tmem_ld_wait();
      if (!t.row_ok) continue;
      const int wbase = t.nb_i * BN + c0;
      for (int j = 0; j < 32 && o0 + c0 + j < out_n; ++j) {
        float a = __uint_as_float(v[j]) * e.alpha, b = __uint_as_float(g[j]) * e.alpha;
        if (e.bias) {
          a += e.bias[wbase + j];
          b += e.bias[wbase + hb + j];
        }
        const float a16 = __half2float(__float2half_rn(a)), b16 = __half2float(__float2half_rn(b));
        orow[o0 + c0 + j] = __float2half_rn(a16 * __half2float(__float2half_rn(gelu_erf(b16))));
      }
    }
  }
}

// 8 accumulator values v[OFF..OFF+8) -> alpha, bias, per-frame bias -> packed fp16
// (template offset: the accumulator array must stay in registers, so it is never indexed through a pointer)
template <int OFF>
__device__ __forceinline__ uint4 epi_pack8(const uint32_t (&v)[32], float alpha, const float* bias, const __half* grow) {
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[OFF + j]) * alpha;
  if (bias) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + 4));
    f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
    f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
  }
  if (grow) {
    // reference: h (fp16 conv output) + emb_out (fp16) -> fp16  (util.py:909-919)
    const uint4 g4 = __ldg(reinterpret_cast<const uint4*>(grow));
    const __half* gh = reinterpret_cast<const __half*>(&g4);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = __half2float(__float2half_rn(f[j])) + __half2float(gh[j]);
  }
  uint4 o;
  o.x = pack_half2(f[0], f[1]);
  o.y = pack_half2(f[2], f[3]);
  o.z = pack_half2(f[4], f[5]);
  o.w = pack_half2(f[6], f[7]);
  return o;
}

// chunk0 / chunk_step: this warp handles the 32-column chunks chunk0, chunk0 + chunk_step, ... (two warps share
// a TMEM lane quadrant and split the columns between them).
__device__ __forceinline__ void tapgemm_epilogue_direct(const TapGemmShape& s, const TapGemmEpilogue& e, const EpiRow& t,
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
      tmem_ld_wait();
      if (!t.row_ok) continue;
      const int nbase = n0 + c0;
      if (nbase >= s.n) continue;
      if (vec_ok && nbase + 32 <= s.n) {
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
            const uint4 g4 = __ldg(reinterpret_cast<const uint4*>(grow + nbase + j8 * 8));
            const __half* gh = reinterpret_cast<const __half*>(&g4);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __half2float(__float2half_rn(f[j])) + __half2float(gh[j]);
          }
          if (rrow) {
            const uint4 r4 = __ldg(reinterpret_cast<const uint4*>(rrow + nbase + j8 * 8));
            const __half* rh = reinterpret_cast<const __half*>(&r4);
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
      for (int j = 0; j < 32; ++j) {
        float a = __uint_as_float(v[j]) * e.alpha;
        float b = __uint_as_float(g[j]) * e.alpha;
        if (e.bias) {
          a += __ldg(e.bias + wbase + j);
          b += __ldg(e.bias + wbase + hb + j);
        }
        // reference rounds the projection to fp16, gelu to fp16, product to fp16
        const float a16 = __half2float(__float2half_rn(a));
        const float b16 = __half2float(__float2half_rn(b));
        const float ge = __half2float(__float2half_rn(gelu_erf(b16)));
        f[j] = a16 * ge;
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


// stage: warp-private staging tile in shared memory (kEpiStageBytes).
// chunk0 / chunk_step: this warp handles the 64-column super-chunks chunk0, chunk0 + chunk_step, ...
__device__ __forceinline__ void tapgemm_epilogue_tile(const TapGemmShape& s, const TapGemmEpilogue& e, const EpiRow& t,
                                                      bool vec_ok, int out_n, int chunk0, int chunk_step, uint8_t* stage) {
  if (!vec_ok || !e.staged) {
    // direct path: every thread stores its own row (32-column chunks interleaved between the two warps)
    tapgemm_epilogue_direct(s, e, t, vec_ok, out_n, chunk0, chunk_step);
    return;
  }
  const int lane = threadIdx.x & 31;
  const int BN = s.bn;
  const int width = e.geglu ? (BN >> 1) : BN;                 // output columns produced by this tile
  const int o_base = e.geglu ? t.nb_i * (BN >> 1) : t.nb_i * BN;  // first output column of the tile
  // transposed read-out geometry: lane handles rows rr = i*4 + lane/8 (i = 0..7), 16-byte piece lane%8
  const int piece = lane & 7;
  long rowoff[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = t.q * 32 + i * 4 + (lane >> 3);
    const int i1 = t.t1_i * s.box1 + (r % s.box1);
    const int i2 = t.t2_i * s.box2 + (r / s.box1);
    const bool ok = (r < s.box1 * s.box2) && (i1 < s.d1) && (i2 < s.d2) && (t.i3 < s.d3);
    rowoff[i] = ok ? (((long)t.i3 * s.d2 + i2) * s.d1 + i1) : -1;
  }
  const __half* grow = (!e.geglu && e.group_bias) ? e.group_bias + (long)(t.i3 / e.group_bias_div) * e.ld_group_bias : nullptr;
  uint8_t* my_row = stage + lane * kEpiPitch;

  for (int c0 = chunk0 * 64; c0 < width; c0 += chunk_step * 64) {
    const int cols = (width - c0) < 64 ? (width - c0) : 64;    // 32 or 64 (BN is a multiple of 32)
    // ---- 1. TMEM -> registers -> fp16 row segment in the staging tile
    if (!e.geglu) {
      uint32_t v0[32], v1[32];
      tmem_ld32(t.t_row + c0, v0);
      if (cols > 32) tmem_ld32(t.t_row + c0 + 32, v1);
      tmem_ld_wait();
      const int nbase = o_base + c0;
#define VG_EPI_PIECE(ARR, OFF, BYTE0)                                                                         \
  {                                                                                                          \
    const int n = nbase + (BYTE0) / 2 + (OFF);                                                               \
    if (n < s.n)                                                                                             \
      *reinterpret_cast<uint4*>(my_row + (BYTE0) + (OFF) * 2) =                                              \
          epi_pack8<OFF>(ARR, e.alpha, e.bias ? e.bias + n : nullptr, grow ? grow + n : nullptr);            \
  }
      VG_EPI_PIECE(v0, 0, 0) VG_EPI_PIECE(v0, 8, 0) VG_EPI_PIECE(v0, 16, 0) VG_EPI_PIECE(v0, 24, 0)
      if (cols > 32) {
        VG_EPI_PIECE(v1, 0, 64) VG_EPI_PIECE(v1, 8, 64) VG_EPI_PIECE(v1, 16, 64) VG_EPI_PIECE(v1, 24, 64)
      }
#undef VG_EPI_PIECE
    } else {
      // GEGLU: accumulator columns [0,BN/2) are "value" j, [BN/2,BN) the matching "gate" j (host interleaves
      // the weight rows per BN block).  out = value * gelu(gate)   (util.py:707-714)
      const int hb = BN >> 1;
      const int wbase = t.nb_i * BN + c0;  // packed weight-row index of value j (gate is + hb)
#pragma unroll 1
      for (int h32 = 0; h32 < cols; h32 += 32) {
        uint32_t v[32], g[32];
        tmem_ld32(t.t_row + c0 + h32, v);
        tmem_ld32(t.t_row + hb + c0 + h32, g);
        tmem_ld_wait();
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int jj = j8 * 8 + j;
            float a = __uint_as_float(v[jj]) * e.alpha;
            float b = __uint_as_float(g[jj]) * e.alpha;
            if (e.bias) {
              a += __ldg(e.bias + wbase + h32 + jj);
              b += __ldg(e.bias + wbase + hb + h32 + jj);
            }
            // reference rounds the projection to fp16, gelu to fp16, product to fp16
            const float a16 = __half2float(__float2half_rn(a));
            const float b16 = __half2float(__float2half_rn(b));
            f[j] = a16 * __half2float(__float2half_rn(gelu_erf(b16)));
          }
          uint4 o;
          o.x = pack_half2(f[0], f[1]);
          o.y = pack_half2(f[2], f[3]);
          o.z = pack_half2(f[4], f[5]);
          o.w = pack_half2(f[6], f[7]);
          *reinterpret_cast<uint4*>(my_row + (h32 + j8 * 8) * 2) = o;
        }
      }
    }
    __syncwarp();
    // ---- 2. transposed read-out: 8 lanes per row segment -> coalesced residual loads and stores
    const int col = o_base + c0 + piece * 8;
    if (piece * 8 < cols && col < out_n) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (rowoff[i] < 0) continue;
        uint4 u = *reinterpret_cast<const uint4*>(stage + (i * 4 + (lane >> 3)) * kEpiPitch + piece * 16);
        if (e.residual) {
          const uint4 r4 = __ldg(reinterpret_cast<const uint4*>(e.residual + rowoff[i] * e.ldr + col));
          const __half2* uh = reinterpret_cast<const __half2*>(&u);
          const __half2* rh = reinterpret_cast<const __half2*>(&r4);
          uint32_t o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 a = __half22float2(uh[k]), b = __half22float2(rh[k]);
            o[k] = pack_half2(a.x + b.x, a.y + b.y);
          }
          u = make_uint4(o[0], o[1], o[2], o[3]);
        }
        *reinterpret_cast<uint4*>(e.out + rowoff[i] * e.ldo + col) = u;
      }
    }
    __syncwarp();
  }
}

}  // namespace vg
