// C-ABI entry points of the tap-GEMM family: they only describe the problem (extents, taps, tile
// shape) and hand it to the tcgen05 kernel (or, for debugging, the SIMT cross-check kernel).
#include <stdlib.h>

#include "common.h"
#include "tapgemm.h"

namespace vg {

static int g_tapgemm_impl = 0;  // 0 = auto (CTA pairs when there are >= 2 M-tiles), 1 = simt, 2 = 1-CTA tcgen05, 3 = 2-CTA tcgen05

// N tile.  With many M tiles the cost is the padded N extent plus a per-tile constant (~24 columns' worth of
// epilogue / pipeline fill).  With FEW M tiles (the 11x20 / 22x40 levels of the UNet, every layer of the 448x256
// configs) the persistent grid runs ceil(tiles / CTA pairs) waves and the widest tile leaves most of the last wave
// idle -- e.g. 32 row pairs x 1280 columns: BN 256 -> 160 tiles = 3 waves of 74 pairs (72 % busy), BN 160 -> 256
// tiles = 4 waves at 0.66 of the per-wave time.  So the cost is waves x (BN + overhead), waves counted on this device.
static int pick_bn(long n, int geglu, long m_pair_tiles, int clusters) {
  if (geglu) {
    for (int bn = 256; bn >= 64; bn -= 64)
      if (n % bn == 0) return bn;
    return 0;
  }
  if (n <= 96) return (int)((n + 31) / 32) * 32;
  if (clusters < 1) clusters = 1;
  int best = 128;
  double best_cost = 1e30;
  for (int bn = 256; bn >= 96; bn -= 32) {
    const long nb = (n + bn - 1) / bn;
    const long tiles = nb * (m_pair_tiles > 0 ? m_pair_tiles : 1);
    const long waves = (tiles + clusters - 1) / clusters;
    // many waves: the tail wave is amortised, fall back to total padded work
    const double work = tiles > 8L * clusters ? (double)tiles / clusters : (double)waves;
    const double cost = work * (bn + 24.0);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

// choose the (box1, box2) output tile (<=128 positions) that wastes the fewest MMA rows
static void pick_box(long d1, long d2, int* b1, int* b2) {
  double best = -1;
  int bb1 = 1, bb2 = 1;
  const int max1 = (int)(d1 < 128 ? d1 : 128);
  for (int x = max1; x >= 1; --x) {
    int y = 128 / x;
    if (y > d2) y = (int)d2;
    if (y > 256) y = 256;
    const double tiles = (double)cdiv(d1, x) * cdiv(d2, y);
    const double eff = (double)d1 * d2 / (tiles * 128.0);
    if (eff > best + 1e-9) {
      best = eff;
      bb1 = x;
      bb2 = y;
    }
  }
  *b1 = bb1;
  *b2 = bb2;
}

static int fill_epilogue(TapGemmArgs* t, void* out, long ldo, const vgen_epilogue* epi) {
  TapGemmEpilogue& e = t->epi;
  e.out = reinterpret_cast<__half*>(out);
  e.ldo = ldo;
  e.alpha = epi ? epi->alpha : 1.0f;
  e.bias = epi ? epi->bias : nullptr;
  e.group_bias = epi ? reinterpret_cast<const __half*>(epi->group_bias) : nullptr;
  e.ld_group_bias = epi ? epi->group_bias_ld : 0;
  e.group_bias_div = (epi && epi->group_bias_div > 1) ? (int)epi->group_bias_div : 1;
  e.residual = epi ? reinterpret_cast<const __half*>(epi->residual) : nullptr;
  e.ldr = epi ? epi->residual_ld : 0;
  e.geglu = epi ? epi->geglu : 0;
  e.row_stats = epi ? reinterpret_cast<const float2*>(epi->row_stats) : nullptr;
  e.col_sum = epi ? epi->col_sum : nullptr;
  VG_REQUIRE((e.row_stats == nullptr) == (e.col_sum == nullptr), "tapgemm: row_stats and col_sum go together");
  VG_REQUIRE((reinterpret_cast<uintptr_t>(e.row_stats) & 7) == 0 && (reinterpret_cast<uintptr_t>(e.col_sum) & 15) == 0,
             "tapgemm: row_stats must be 8-byte and col_sum 16-byte aligned");
  return 0;
}

static int finish_and_launch(TapGemmArgs* t, const vgen_epilogue* epi, void* stream) {
  static bool env_read = false;
  if (!env_read) {  // VGEN_TAPGEMM_IMPL=0..3 overrides the default once (A/B measurements)
    env_read = true;
    const char* ev = getenv("VGEN_TAPGEMM_IMPL");
    if (ev && ev[0] >= '0' && ev[0] <= '3') g_tapgemm_impl = ev[0] - '0';
  }
  TapGemmShape& s = t->shape;
  s.kc = s.c / 64;
  const long m_tiles_all = (long)s.d3 * cdiv(s.d2, s.box2) * cdiv(s.d1, s.box1);
  int bn = (epi && epi->bn > 0) ? epi->bn : pick_bn(s.n, t->epi.geglu, (m_tiles_all + 1) / 2, sm_count() / 2);
  VG_REQUIRE(bn > 0, "tapgemm: no valid N tile (GEGLU needs n % 64 == 0)");
  s.bn = bn;
  s.nb = cdiv(s.n, bn);
  s.t1 = cdiv(s.d1, s.box1);
  s.t2 = cdiv(s.d2, s.box2);
  const long tiles = (long)s.d3 * s.t2 * s.t1 * s.nb;
  VG_REQUIRE(tiles < (1L << 31), "tapgemm: too many tiles");
  s.total_tiles = (int)tiles;
  if (tiles == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (g_tapgemm_impl == 1 || (s.c % 64) != 0) return tapgemm_simt_launch(*t, st);
  const long m_tiles = (long)s.d3 * s.t2 * s.t1;
  const bool pair = g_tapgemm_impl == 3 || (g_tapgemm_impl == 0 && m_tiles >= 2);
  return pair ? tapgemm_sm100_2cta_launch(*t, st) : tapgemm_sm100_launch(*t, st);
}

}  // namespace vg

using namespace vg;

extern "C" {

int vgen_set_tapgemm_impl(int impl) {
  if (impl < 0 || impl > 3) return fail("vgen_set_tapgemm_impl: impl must be 0 (auto), 1 (simt), 2 (1-CTA) or 3 (2-CTA)");
  g_tapgemm_impl = impl;
  return 0;
}

int vgen_linear(const void* a, int64_t m, int64_t k, int64_t lda, const void* w, int64_t n, void* out, int64_t ldo,
                const vgen_epilogue* epi, void* stream) {
  VG_REQUIRE(a && w && out, "vgen_linear: null pointer");
  VG_REQUIRE(m >= 0 && k > 0 && n > 0 && lda >= k, "vgen_linear: bad shape");
  if (m == 0) return 0;
  TapGemmArgs t{};
  t.a = reinterpret_cast<const __half*>(a);
  t.w = reinterpret_cast<const __half*>(w);
  t.a_stride1 = lda;
  t.a_stride2 = lda * m;
  t.a_stride3 = lda * m;
  TapGemmShape& s = t.shape;
  s.c = (int)k;
  s.d1 = (int)m;
  s.d2 = 1;
  s.d3 = 1;
  s.box1 = 128;
  s.box2 = 1;
  s.num_taps = 1;
  s.tap1[0] = s.tap2[0] = s.tap3[0] = 0;
  s.n = (int)n;
  {
    int rc = fill_epilogue(&t, out, ldo, epi);
    if (rc) return rc;
  }
  VG_REQUIRE(!t.epi.group_bias, "vgen_linear: group_bias is only defined for conv entries");
  return finish_and_launch(&t, epi, stream);
}

int vgen_conv2d_3x3(const void* x, int64_t nimg, int64_t h, int64_t w_, int64_t c, const void* w, int64_t n, void* out,
                    int64_t ldo, const vgen_epilogue* epi, void* stream) {
  VG_REQUIRE(x && w && out, "vgen_conv2d_3x3: null pointer");
  VG_REQUIRE(nimg >= 0 && h > 0 && w_ > 0 && c > 0 && n > 0, "vgen_conv2d_3x3: bad shape");
  if (nimg == 0) return 0;
  TapGemmArgs t{};
  t.a = reinterpret_cast<const __half*>(x);
  t.w = reinterpret_cast<const __half*>(w);
  t.a_stride1 = c;
  t.a_stride2 = c * w_;
  t.a_stride3 = c * w_ * h;
  TapGemmShape& s = t.shape;
  s.c = (int)c;
  s.d1 = (int)w_;
  s.d2 = (int)h;
  s.d3 = (int)nimg;
  pick_box(w_, h, &s.box1, &s.box2);
  s.num_taps = 9;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      s.tap1[ky * 3 + kx] = kx - 1;
      s.tap2[ky * 3 + kx] = ky - 1;
      s.tap3[ky * 3 + kx] = 0;
    }
  s.n = (int)n;
  {
    int rc = fill_epilogue(&t, out, ldo, epi);
    if (rc) return rc;
  }
  VG_REQUIRE(!t.epi.row_stats, "vgen_conv2d_3x3: row_stats (folded LayerNorm) is only defined for vgen_linear");
  return finish_and_launch(&t, epi, stream);
}

int vgen_tconv3_batch(const void* x, int64_t batch, int64_t f, int64_t hw, int64_t c, const void* w, int64_t n, void* out,
                      int64_t ldo, const vgen_epilogue* epi, void* stream) {
  VG_REQUIRE(x && w && out, "vgen_tconv3: null pointer");
  VG_REQUIRE(batch >= 0 && f >= 0 && hw > 0 && c > 0 && n > 0, "vgen_tconv3: bad shape");
  if (f == 0 || batch == 0) return 0;
  TapGemmArgs t{};
  t.a = reinterpret_cast<const __half*>(x);
  t.w = reinterpret_cast<const __half*>(w);
  t.a_stride1 = c;
  t.a_stride2 = c * hw;
  t.a_stride3 = c * hw * f;
  TapGemmShape& s = t.shape;
  s.c = (int)c;
  s.d1 = (int)hw;
  s.d2 = (int)f;
  s.d3 = (int)batch;  // videos: the frame taps are zero-padded by the TMA unit at each video's first / last frame
  s.box1 = (int)(hw < 128 ? hw : 128);
  s.box2 = 1;
  if (hw < 128) {  // small planes: put several frames in one tile
    s.box2 = (int)(128 / hw);
    if (s.box2 > f) s.box2 = (int)f;
  }
  s.num_taps = 3;
  for (int kt = 0; kt < 3; ++kt) {
    s.tap1[kt] = 0;
    s.tap2[kt] = kt - 1;
    s.tap3[kt] = 0;
  }
  s.n = (int)n;
  {
    int rc = fill_epilogue(&t, out, ldo, epi);
    if (rc) return rc;
  }
  VG_REQUIRE(!t.epi.group_bias, "vgen_tconv3: group_bias not supported");
  VG_REQUIRE(!t.epi.row_stats, "vgen_tconv3: row_stats (folded LayerNorm) is only defined for vgen_linear");
  return finish_and_launch(&t, epi, stream);
}

int vgen_tconv3(const void* x, int64_t f, int64_t hw, int64_t c, const void* w, int64_t n, void* out, int64_t ldo,
                const vgen_epilogue* epi, void* stream) {
  return vgen_tconv3_batch(x, 1, f, hw, c, w, n, out, ldo, epi, stream);
}

}  // extern "C"
