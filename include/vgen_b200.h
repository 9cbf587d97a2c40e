/*
 * vgen_b200 -- C ABI of libvgen_b200.so: hand-written sm_100a kernels for the VGen sampling hot path
 * (DDIM loop -> spatio-temporal UNet forward -> AutoencoderKL decode).
 *
 * The reference (ali-vilab/VGen) is pure PyTorch and has no FFI for this path: every op below replaces
 * a torch/cuDNN/cuBLAS/xformers library call made from a reference nn.Module.  Each entry cites the
 * reference call site it stands in for (paths relative to the reference root).  The Python classes in
 * vgen_b200/ (registered under the reference's MODEL / DIFFUSION / AUTO_ENCODER registry names) are the
 * only callers; they bind this header with ctypes.
 *
 * Conventions
 *   - plain C: raw device pointers, sizes and a cudaStream_t passed as void*; no torch types.
 *   - activations are fp16, channels-last ("tokens x channels", NHWC / [f][h][w][c]); weights fp16,
 *     K-major [out][taps*in]; biases / norm affine fp32; all accumulation fp32.
 *   - every function returns 0 on success; otherwise vgen_last_error() describes the failure.
 *   - no hidden synchronisation, no allocation: the caller owns every buffer (incl. workspaces).
 *   - not thread-safe per stream; one process per GPU.
 */
#ifndef VGEN_B200_H_
#define VGEN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VGEN_B200_ABI_VERSION 1

/* ---- library ---------------------------------------------------------------------------------- */
int vgen_abi_version(void);
const char* vgen_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
int64_t vgen_launch_count(void);
/* 0 = tcgen05/TMA kernels (default), 1 = SIMT cross-check kernels for the tap-GEMM family (debug only). */
int vgen_set_tapgemm_impl(int impl);

/* ---- tap-GEMM family (tcgen05.mma + TMA; tapgemm_sm100.cu) ------------------------------------ */
typedef struct vgen_epilogue {
  float alpha;              /* accumulator scale (1.0f = none)                                        */
  const float* bias;        /* [n] fp32 or NULL                                                       */
  const void* group_bias;   /* fp16 [groups][n] or NULL: per-frame bias, added after fp16 rounding     */
  int64_t group_bias_ld;    /*   (ResBlock "h + emb_out": tools/modules/unet/util.py:909-919)          */
  const void* residual;     /* fp16 [rows][n] or NULL: added after fp16 rounding (skip / x_in adds)    */
  int64_t residual_ld;
  int geglu;                /* 1: out[rows][n/2] = value * gelu(gate); W rows interleaved per bn block */
  int bn;                   /* N tile (32..256, multiple of 32; GEGLU: multiple of 64); 0 = auto       */
} vgen_epilogue;

/* out[m][n] = epi(a[m][:k] . w[n][:k])       -- nn.Linear / 1x1 conv / Conv1d(k=1)
 * replaces: util.py:224-229 (to_q/k/v/out), :711 (GEGLU.proj), :736 (FF out), :338,:353 (proj_in/out),
 *           :1213,:1229 (Conv1d proj), autoencoder.py:344-363 (q/k/v/proj_out 1x1), :305 (nin_shortcut) */
int vgen_linear(const void* a, int64_t m, int64_t k, int64_t lda, const void* w, int64_t n, void* out,
                int64_t ldo, const vgen_epilogue* epi, void* stream);

/* 3x3 / stride 1 / pad 1 conv on [nimg][h][w][c] fp16; w is [n][9*c] with k = (ky*3+kx)*c + ci.
 * replaces: nn.Conv2d in ResBlock util.py:845-876, Upsample.conv :761, head unet_t2v.py:204-207,
 *           VAE ResnetBlock / Upsample / conv_in / conv_out autoencoder.py:282-300,448-452,603,647 */
int vgen_conv2d_3x3(const void* x, int64_t nimg, int64_t h, int64_t w_, int64_t c, const void* w, int64_t n,
                    void* out, int64_t ldo, const vgen_epilogue* epi, void* stream);

/* temporal (3,1,1) conv, pad (1,0,0), on [f][hw][c] fp16 (one video); w is [n][3*c], k = kt*c + ci.
 * replaces: nn.Conv3d in TemporalConvBlock_v2 util.py:1662-1680 */
int vgen_tconv3(const void* x, int64_t f, int64_t hw, int64_t c, const void* w, int64_t n, void* out,
                int64_t ldo, const vgen_epilogue* epi, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VGEN_B200_H_ */
