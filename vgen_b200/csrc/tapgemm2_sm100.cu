// tapgemm, CTA-pair variant (tcgen05 cta_group::2): same maths and epilogue as tapgemm_sm100.cu, but two
// CTAs of a cluster (the two SMs of a TPC) compute a 256 x BN tile together.  Each CTA stages its own 128
// rows of A and only HALF of the W tile; the leader's single tcgen05.mma.cta_group::2 reads both CTAs'
// shared memory.  Per FLOP this moves (128*64 + BN/2*64) instead of (128*64 + BN*64) operand elements
// L2 -> SM per CTA, which is what bounds the 1-CTA kernel (profiles/r01a_ncu_summary.md).
//
//   warp 0 (both CTAs)   TMA producer: own A box + own half of W, completion signalled on the LEADER's
//                        full barrier (cp.async.bulk.tensor ... cta_group::2)
//   warp 1 (leader only) MMA issuer: M = 256, commits multicast to both CTAs' empty / accumulator-full barriers
//   warps 2-5 (both)     epilogue on the CTA's own 128 accumulator rows (its own TMEM)
#include "common.h"
#include "ptx.cuh"
#include "tapgemm.h"
#include "tapgemm_epilogue.cuh"

namespace vg {

static constexpr int kBM2 = 128;       // rows per CTA (UMMA M = 256 per pair)
static constexpr int kBK2 = 64;
static constexpr int kThreads2 = 320;  // TMA, MMA, 8 epilogue warps
static constexpr uint32_t kTmemCols2 = 512;
static constexpr int kABytes2 = kBM2 * kBK2 * 2;

struct alignas(64) TapGemm2Params {
  CUtensorMap map_a;
  CUtensorMap map_b;     // box {64, BN/2}
  CUtensorMap map_out;   // output, box {32, box1, box2, 1}, 64B swizzle (TMA tile stores of the epilogue)
  int out_tma;           // 1: map_out is valid
  CUtensorMap map_res;   // residual, same box / swizzle as map_out (valid when res_tma)
  int res_tma;
  int staging_bytes;     // kEpiStagingTotal, doubled when residual tiles are staged too
  TapGemmShape s;
  TapGemmEpilogue e;
  int stages;
  int b_slot_bytes;
  int m_tiles;           // d3 * t2 * t1
  int total_pair_tiles;  // ceil(m_tiles / 2) * nb
};

// kEpi: bit 0 = GEGLU epilogue, bit 1 = LayerNorm folded into the epilogue (TapGemmEpilogue.row_stats); separate
// instantiations because the epilogues have very different register needs.
template <int kEpi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
    tapgemm_sm100_2cta_kernel(const __grid_constant__ TapGemm2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr bool kGeglu = (kEpi & 1) != 0, kLn = (kEpi & 2) != 0;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int stages = p.stages;
  const int stage_bytes = kABytes2 + p.b_slot_bytes;

  uint8_t* tiles = smem;
  uint8_t* staging = smem + (size_t)stages * stage_bytes;   // epilogue staging tiles (1024-aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + p.staging_bytes);
  uint64_t* full_bar = bars;                      // [stages]  (the leader's copy is the one in use)
  uint64_t* empty_bar = bars + stages;            // [stages]  local, arrived by the multicast commit
  uint64_t* tfull_bar = bars + 2 * stages;        // [2]       local, arrived by the multicast commit
  uint64_t* tempty_bar = bars + 2 * stages + 2;   // [2]       the leader's copy collects 8 warp arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 4);
  uint64_t* res_bar = bars + 2 * stages + 6;      // [2 column groups][2 slots] residual tiles landed

  const TapGemmShape& s = p.s;
  const int k_iters = s.num_taps * s.kc;
  const int BN = s.bn;
  const int half_n = BN >> 1;
  const uint32_t stage_tx = (uint32_t)(s.box1 * s.box2 * kBK2 * 2 + half_n * kBK2 * 2);  // bytes landing in ONE CTA

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_a);
    tma_prefetch_desc(&p.map_b);
    if (p.out_tma) tma_prefetch_desc(&p.map_out);
    if (p.res_tma) tma_prefetch_desc(&p.map_res);
    for (int i = 0; i < 4; ++i) mbar_init(&res_bar[i], 1);
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 16);  // 8 epilogue warps x 2 CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm<kTmemCols2>(tmem_slot);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    // the whole warp walks the loop (warp-uniform operands); one elected lane issues the TMA
    int st = 0;
    uint32_t ph = 0;  // ring position / phase, carried across tiles
    const uint32_t full_leader0 = mapa_shared(smem_u32(&full_bar[0]), 0);
    for (int pt = cluster_id; pt < p.total_pair_tiles; pt += num_clusters) {
      int mq, nb_i, rest, t1_i, i3, t2_i;
      fd_divmod(s.f_nb, pt, mq, nb_i);
      const int m = 2 * mq + (int)rank;   // this CTA's M-tile (may be one past the end: loads zeros)
      fd_divmod(s.f_t1, m, rest, t1_i);
      fd_divmod(s.f_t2, rest, i3, t2_i);
      const int i1_0 = t1_i * s.box1, i2_0 = t2_i * s.box2;
      const int n0 = nb_i * BN + (int)rank * half_n;
      for (int tap = 0; tap < s.num_taps; ++tap) {
        const int c1 = i1_0 + s.tap1[tap], c2 = i2_0 + s.tap2[tap], c3 = i3 + s.tap3[tap];
        const int wk0 = tap * s.c;
        for (int kc_i = 0; kc_i < s.kc; ++kc_i) {
          mbar_wait(&empty_bar[st], ph ^ 1, 31);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(&full_bar[st], 2 * stage_tx);
            const uint32_t full_leader = full_leader0 + (uint32_t)st * 8u;
            uint8_t* sa = tiles + (size_t)st * stage_bytes;
            tma_load_4d_2sm(sa, &p.map_a, full_leader, kc_i * kBK2, c1, c2, c3);
            tma_load_2d_2sm(sa + kABytes2, &p.map_b, full_leader, wk0 + kc_i * kBK2, n0);
          }
          __syncwarp();
          if (++st == stages) {
            st = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    // whole warp in the loop, one elected lane issues tcgen05.mma / commit
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_f16(2 * kBM2, BN, 0, 0);
      const uint32_t tiles_addr = smem_u32(tiles);
      int st = 0;
      uint32_t ph = 0;
      uint32_t lt = 0;
      for (int pt = cluster_id; pt < p.total_pair_tiles; pt += num_clusters, ++lt) {
        const uint32_t as = lt & 1, aph = (lt >> 1) & 1;
        mbar_wait(&tempty_bar[as], aph ^ 1, 32);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full_bar[st], ph, 33);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = tiles_addr + (uint32_t)st * (uint32_t)stage_bytes;
            const uint64_t a_desc = umma_desc_sw128(a_addr, 16, 1024);
            const uint64_t b_desc = umma_desc_sw128(a_addr + kABytes2, 16, 1024);
#pragma unroll
            for (int k = 0; k < kBK2 / 16; ++k)
              umma_f16_ss_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);
            umma_commit_2sm(&empty_bar[st], 3);                         // frees the stage in BOTH CTAs
            if (it == k_iters - 1) umma_commit_2sm(&tfull_bar[as], 3);  // accumulator ready in BOTH CTAs
          }
          __syncwarp();
          if (++st == stages) {
            st = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5, both CTAs)
    const TapGemmEpilogue& e = p.e;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int rows_in_tile = s.box1 * s.box2;
    const int out_n = kGeglu ? (s.n >> 1) : s.n;
    const bool vec_ok = tapgemm_vec_ok(e, out_n);
    int r1, r2;                       // position of this thread's row inside the box (tile-invariant)
    fd_divmod(s.f_box1, r, r2, r1);
    const int cg = (warp - 2) >> 2;   // column group: the 4 warps (all lane quadrants) working on the same chunk
    EpiStore est;
    est.tma = p.out_tma != 0;
    est.map = &p.map_out;
    est.stage = smem_u32(staging) + (uint32_t)cg * 2u * kEpiStageBytes;
    est.row_off = (uint32_t)r * 64u;
    est.row_xor = ((uint32_t)r >> 1) & 3u;
    est.bar_id = 1 + cg;
    est.leader = (r == 0);
    est.slot = 0;
    est.res_tma = !kGeglu && p.res_tma != 0;
    est.map_res = &p.map_res;
    est.rstage = smem_u32(staging) + kEpiStagingTotal + (uint32_t)cg * 2u * kEpiStageBytes;
    est.rbar = res_bar + 2 * cg;
    est.box_bytes = (uint32_t)(s.box1 * s.box2 * 64);
    est.r_issue = est.r_cons = 0;
    uint32_t lt = 0;
    for (int pt = cluster_id; pt < p.total_pair_tiles; pt += num_clusters, ++lt) {
      const uint32_t as = lt & 1, aph = (lt >> 1) & 1;
      EpiRow t;
      int mq, rest, t1_i, t2_i;
      fd_divmod(s.f_nb, pt, mq, t.nb_i);
      const int m = 2 * mq + (int)rank;
      fd_divmod(s.f_t1, m, rest, t1_i);
      fd_divmod(s.f_t2, rest, t.i3, t2_i);
      t.i1_0 = t1_i * s.box1;
      t.i2_0 = t2_i * s.box2;
      const int i1 = t.i1_0 + r1;
      const int i2 = t.i2_0 + r2;
      t.row_ok = (r < rows_in_tile) && (i1 < s.d1) && (i2 < s.d2) && (t.i3 < s.d3);
      t.row = ((long)t.i3 * s.d2 + i2) * s.d1 + i1;
      if (!kGeglu && est.res_tma) {
        // the residual tile of this group's first chunk: requested now, while the tile's MMAs are still running
        const int col = t.nb_i * BN + cg * 32;
        if (vec_ok && cg * 32 < BN && col + 32 <= s.n) epi_res_issue(est, col, t);
      } else if (!kGeglu && e.residual && t.row_ok) {
        // the residual row segment comes from HBM: start pulling it into L2 while the tile's MMAs are still running
        const __half* rp = e.residual + t.row * e.ldr + t.nb_i * BN;
        for (int c0 = cg * 32; c0 < BN && t.nb_i * BN + c0 < s.n; c0 += 64)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + c0));
      }
      mbar_wait(&tfull_bar[as], aph, 34);
      tc_fence_after();
      t.t_row = tmem_base + as * 256 + ((uint32_t)(q * 32) << 16);
      tapgemm_epilogue_tile<kGeglu, kLn>(s, e, t, est, vec_ok, out_n, cg, 2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed_cluster(mapa_shared(smem_u32(&tempty_bar[as]), 0));
    }
    epi_stage_drain(est);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<kTmemCols2>(tmem_base);
  }
}

int tapgemm_sm100_2cta_launch(const TapGemmArgs& a, cudaStream_t stream) {
  const TapGemmShape& s = a.shape;
  VG_REQUIRE(s.c % 64 == 0, "tapgemm_2cta: C must be a multiple of 64");
  VG_REQUIRE(s.bn >= 32 && s.bn <= 256 && s.bn % 32 == 0, "tapgemm_2cta: BN must be in [32,256], multiple of 32");
  VG_REQUIRE(s.box1 >= 1 && s.box2 >= 1 && s.box1 * s.box2 <= kBM2 && s.box1 <= 256 && s.box2 <= 256, "tapgemm_2cta: bad box");
  VG_REQUIRE(s.num_taps >= 1 && s.num_taps <= kMaxTaps, "tapgemm_2cta: bad tap count");
  VG_REQUIRE((reinterpret_cast<uintptr_t>(a.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.w) & 15) == 0,
             "tapgemm_2cta: A and W must be 16-byte aligned");
  if (a.epi.geglu) VG_REQUIRE(s.bn % 64 == 0 && s.n % s.bn == 0, "tapgemm_2cta: GEGLU needs BN%64==0 and N%BN==0");

  TapGemm2Params p;
  p.s = s;
  tapgemm_prepare_shape(p.s);
  p.e = a.epi;
  p.e.f_group_bias_div = make_fastdiv(p.e.group_bias_div > 1 ? p.e.group_bias_div : 1);
  {
    const uint64_t dims[4] = {(uint64_t)s.c, (uint64_t)s.d1, (uint64_t)s.d2, (uint64_t)s.d3};
    const uint64_t strides[3] = {(uint64_t)a.a_stride1 * 2, (uint64_t)a.a_stride2 * 2, (uint64_t)a.a_stride3 * 2};
    const uint32_t box[4] = {64, (uint32_t)s.box1, (uint32_t)s.box2, 1};
    int rc = make_tmap_f16(&p.map_a, a.a, 4, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)s.num_taps * s.c, (uint64_t)s.n};
    const uint64_t strides[1] = {(uint64_t)s.num_taps * s.c * 2};
    const uint32_t box[2] = {64, (uint32_t)(s.bn / 2)};
    int rc = make_tmap_f16(&p.map_b, a.w, 2, dims, strides, box);
    if (rc) return rc;
  }
  p.b_slot_bytes = (((s.bn / 2) * kBK2 * 2 + 1023) / 1024) * 1024;
  const int stage_bytes = kABytes2 + p.b_slot_bytes;
  const int out_n_h = a.epi.geglu ? s.n / 2 : s.n;
  const bool out_tma_h = tapgemm_out_tma_ok(a.epi) && out_n_h >= 32 && !getenv("VGEN_TAPGEMM_DIRECT_STORE");
  const bool res_tma_h = out_tma_h && !a.epi.geglu && a.epi.residual && (a.epi.ldr & 7) == 0 &&
                         (reinterpret_cast<uintptr_t>(a.epi.residual) & 15) == 0 && !getenv("VGEN_TAPGEMM_DIRECT_RESIDUAL");
  p.staging_bytes = kEpiStagingTotal * (res_tma_h ? 2 : 1);
  int stages = (224 * 1024 - p.staging_bytes) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  p.stages = stages;
  p.m_tiles = s.d3 * s.t2 * s.t1;
  p.total_pair_tiles = ((p.m_tiles + 1) / 2) * s.nb;
  const size_t smem = (size_t)stages * stage_bytes + p.staging_bytes + (2 * stages + 10) * 8 + 1024;
  {
    const int out_n = a.epi.geglu ? s.n / 2 : s.n;
    p.out_tma = out_tma_h ? 1 : 0;
    p.res_tma = res_tma_h ? 1 : 0;
    if (p.res_tma) {
      const uint64_t ld = (uint64_t)a.epi.ldr * 2;
      const uint64_t dims[4] = {(uint64_t)out_n, (uint64_t)s.d1, (uint64_t)s.d2, (uint64_t)s.d3};
      const uint64_t strides[3] = {ld, ld * s.d1, ld * s.d1 * s.d2};
      const uint32_t box[4] = {32, (uint32_t)s.box1, (uint32_t)s.box2, 1};
      int rc = make_tmap_f16(&p.map_res, a.epi.residual, 4, dims, strides, box, 64);
      if (rc) return rc;
    }
    if (p.out_tma) {
      const uint64_t ld = (uint64_t)a.epi.ldo * 2;
      const uint64_t dims[4] = {(uint64_t)out_n, (uint64_t)s.d1, (uint64_t)s.d2, (uint64_t)s.d3};
      const uint64_t strides[3] = {ld, ld * s.d1, ld * s.d1 * s.d2};
      const uint32_t box[4] = {32, (uint32_t)s.box1, (uint32_t)s.box2, 1};
      int rc = make_tmap_f16(&p.map_out, a.epi.out, 4, dims, strides, box, 64);
      if (rc) return rc;
    }
  }

  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    VG_CUDA(cudaFuncSetAttribute(tapgemm_sm100_2cta_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VG_CUDA(cudaFuncSetAttribute(tapgemm_sm100_2cta_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VG_CUDA(cudaFuncSetAttribute(tapgemm_sm100_2cta_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VG_CUDA(cudaFuncSetAttribute(tapgemm_sm100_2cta_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_once.mark();
  }
  int clusters = sm_count() / 2;
  if (clusters > p.total_pair_tiles) clusters = p.total_pair_tiles;
  if (clusters < 1) return 0;
  const int kind = (a.epi.geglu ? 1 : 0) | (a.epi.row_stats ? 2 : 0);
  if (a.epi.row_stats) VG_REQUIRE(!a.epi.residual && !a.epi.group_bias && s.num_taps == 1 && s.d2 == 1 && s.d3 == 1,
                                  "tapgemm: a folded LayerNorm goes with a plain linear (no residual / per-frame bias)");
  if (kind == 3) launch_kernel(tapgemm_sm100_2cta_kernel<3>, dim3(2 * clusters), dim3(kThreads2), smem, stream, p);
  else if (kind == 2) launch_kernel(tapgemm_sm100_2cta_kernel<2>, dim3(2 * clusters), dim3(kThreads2), smem, stream, p);
  else if (kind == 1) launch_kernel(tapgemm_sm100_2cta_kernel<1>, dim3(2 * clusters), dim3(kThreads2), smem, stream, p);
  else launch_kernel(tapgemm_sm100_2cta_kernel<0>, dim3(2 * clusters), dim3(kThreads2), smem, stream, p);
  VG_LAUNCH_CHECK("tapgemm_sm100_2cta_kernel");
  return 0;
}

}  // namespace vg
