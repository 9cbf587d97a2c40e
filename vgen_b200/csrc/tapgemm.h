// Internal description of one tap-GEMM problem (see tapgemm_sm100.cu for the maths).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace vg {

static constexpr int kMaxTaps = 9;

struct TapGemmShape {
  int c;            // channels of A (K per tap)
  int kc;           // c / 64
  int d1, d2, d3;   // logical extents of A / of the output rows (row = (i3*d2 + i2)*d1 + i1)
  int box1, box2;   // output tile = box1 x box2 positions (<= 128) of one i3
  int t1, t2;       // tiles along d1 / d2
  int num_taps;
  int tap1[kMaxTaps], tap2[kMaxTaps], tap3[kMaxTaps];  // coordinate offsets per tap
  int n;            // output channels (rows of W)
  int bn;           // N tile
  int nb;           // ceil(n / bn)
  int total_tiles;  // d3 * t2 * t1 * nb
};

struct TapGemmEpilogue {
  __half* out;
  long ldo;                  // elements between consecutive output rows
  float alpha;               // accumulator scale
  const float* bias;         // [n] or null   (GEGLU: indexed like the packed W rows)
  const __half* group_bias;  // [d3][n] or null: added per outermost coordinate (frame) after rounding
  long ld_group_bias;
  int group_bias_div;        // bias row = i3 / group_bias_div
  const __half* residual;    // [rows][n] or null: added after rounding
  long ldr;
  int geglu;                 // out has n/2 columns: value * gelu(gate)
};

struct TapGemmArgs {
  const __half* a;
  long a_stride1, a_stride2, a_stride3;  // element strides of d1, d2, d3 (channel stride is 1)
  const __half* w;                       // [n][num_taps * c], K-major
  TapGemmShape shape;
  TapGemmEpilogue epi;
};

int tapgemm_sm100_launch(const TapGemmArgs& a, cudaStream_t stream);       // tcgen05 + TMA, one CTA per tile
int tapgemm_sm100_2cta_launch(const TapGemmArgs& a, cudaStream_t stream);  // tcgen05 cta_group::2, CTA pair per 256-row tile
int tapgemm_simt_launch(const TapGemmArgs& a, cudaStream_t stream);   // plain SIMT cross-check kernel

}  // namespace vg
