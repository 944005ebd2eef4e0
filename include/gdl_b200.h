/* gdl_b200.h — C ABI of libgdlb200.so: the B200-native (sm_100a) kernels behind
 * NRCan/geo-deep-learning's segmentation hot path.
 *
 * The reference has no FFI today (it is pure Python on top of ATen/cuDNN/cuBLAS); every
 * entry point below cites the reference call site whose arithmetic it replaces
 * (paths relative to the reference tree, geo_deep_learning/...).  INTEGRATION.md shows the
 * ctypes stub a maintainer would add on the reference side.
 *
 * Conventions
 *   - plain pointers + sizes only; no torch types.  All pointers are DEVICE pointers unless
 *     a parameter says "host".  The caller owns every buffer; the library allocates nothing
 *     and keeps no pointer after a call returns.
 *   - activations are NHWC ("channels last"): element (n,h,w,c) at base[((n*H+h)*W+w)*ld+c],
 *     `ld` = pixel stride in elements (>= C, multiple of 8) so channel slices are addressable.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously.
 *   - return value: 0 = ok, otherwise a GDL_ERR_* code; gdl_last_error() returns the text
 *     (thread-local).  Nothing throws across the boundary.  The Python wrapper maps
 *     GDL_ERR_INVALID -> ValueError, GDL_ERR_UNSUPPORTED -> NotImplementedError,
 *     GDL_ERR_CUDA -> RuntimeError, mirroring the exception types the reference raises
 *     (mix_transformer.py:82-84, multilevel_neck.py:141-146).
 *   - re-entrant: no global mutable state apart from lazily initialised function attributes.
 */
#ifndef GDL_B200_H_
#define GDL_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define GDL_B200_VERSION 100

#define GDL_OK 0
#define GDL_ERR_INVALID 1     /* bad shape / argument                      */
#define GDL_ERR_UNSUPPORTED 2 /* valid request outside the implemented set */
#define GDL_ERR_CUDA 3        /* CUDA runtime / driver failure             */

#define GDL_BF16 0
#define GDL_F32 1
#define GDL_F16 2

#define GDL_MAX_SRC 6

const char* gdl_last_error(void);
int gdl_version(void);
int gdl_device_info(int* sm_count, int* cc_major, int* cc_minor, unsigned long long* total_mem);

/* One member of a "virtual concat": a conv reads its input channels from up to GDL_MAX_SRC
 * NHWC tensors of identical N,H,W (replaces torch.cat([...], dim=1) in smp UnetPlusPlus'
 * DecoderBlock, segformer_mlp.py:127, upernet.py:145 — the concat is never materialised). */
typedef struct {
  const void* ptr; /* 16-bit NHWC activations                        */
  int channels;    /* channels taken from this source (multiple of 16) */
  int ld;          /* pixel stride in elements                       */
} gdl_src_t;

/* ---- tensor-core implicit-GEMM convolution, stride 1 ---------------------------------------
 * out[n,ho,wo,k] = act( bias[k] + sum_{r,s,c} in[n,ho+r-pad_h,wo+s-pad_w,c] * w[k,r,s,c] )
 * Replaces nn.Conv2d inside models/utils.py:10-52 (ConvModule), smp Conv2dReLU / ResNet convs
 * (segmentation_unetplus.py:126-131), the 1x1 convs / nn.Linear of segformer_mlp.py:8-19,63-75
 * and mix_transformer.py Attention/Mlp projections (R=S=1 == GEMM).  tcgen05.mma (bf16/f16 in,
 * fp32 accumulate in TMEM), operands staged by TMA with out-of-bounds zero fill as the padding.
 * weight: packed [Cout][R][S][Ctot] 16-bit (see gdl_pack_conv_weight), Ctot = sum(src.channels).
 * Strided convolutions go through gdl_im2col + this entry point with R=S=1.               */
typedef struct {
  int N, H, W;
  int num_src;
  gdl_src_t src[GDL_MAX_SRC];
  int Cout;
  int R, S, pad_h, pad_w;
  const void* weight;
  int dtype;       /* GDL_BF16 or GDL_F16: operand type        */
  void* out;       /* NHWC [N][Ho][Wo][ldo], Ho = H+2*pad_h-R+1 */
  int out_dtype;   /* GDL_BF16 / GDL_F16 / GDL_F32             */
  int ldo;
  const float* bias; /* optional [Cout] fp32                   */
  int relu;
} gdl_conv_fwd_t;
int gdl_conv2d_nhwc_fwd(const gdl_conv_fwd_t* d, void* stream);

/* ---- weight gradient ------------------------------------------------------------------------
 * dw[k,r,s,c] += sum_{n,ho,wo} dy[n,ho,wo,k] * in[n,ho+r-pad_h,wo+s-pad_w,c]
 * (autograd of the convs above: a20 in SURVEY.md §8).  dw is fp32 [Cout][R][S][Ctot] and is
 * ACCUMULATED into (split-K partial sums are added with red.global.add.f32): zero it first. */
typedef struct {
  int N, H, W;
  int num_src;
  gdl_src_t src[GDL_MAX_SRC];
  int Cout;
  int R, S, pad_h, pad_w;
  const void* dy; /* NHWC [N][Ho][Wo][ld_dy] 16-bit */
  int ld_dy;
  int dtype;
  float* dw;
} gdl_conv_wgrad_t;
int gdl_conv2d_nhwc_wgrad(const gdl_conv_wgrad_t* d, void* stream);

/* Input gradient of a stride-1 conv is gdl_conv2d_nhwc_fwd applied to dy with the
 * tap-flipped, channel-transposed weights produced by gdl_pack_conv_weight(transpose=1). */

/* ---- weight layout transforms ---------------------------------------------------------------
 * src: fp32 OIHW [Cout][Cin][R][S] (the nn.Parameter / state_dict layout of the reference).
 * mode 0: dst[k][(r,s,c)]          = src[k][c][r][s]   forward operand  (rows Cout)
 * mode 1: dst[c][(R-1-r,S-1-s,k)]  = src[k][c][r][s]   dgrad operand    (rows Cin, taps flipped)
 * mode 2: dst[(r,s,c)][k]          = src[k][c][r][s]   dgrad operand of an im2col'd conv
 * dst is 16-bit (dtype); dst_ld = row stride in elements (0 = dense), pad columns are zeroed. */
int gdl_pack_conv_weight(const float* src, void* dst, int Cout, int Cin, int R, int S, int mode,
                         int dst_ld, int dtype, void* stream);
/* grad fp32 [Cout][src_ld] (columns (r,s,c), from wgrad) -> fp32 OIHW; dst = (accumulate? dst:0)+src */
int gdl_unpack_conv_wgrad(const float* src, float* dst, int Cout, int Cin, int R, int S, int src_ld,
                          int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GDL_B200_H_ */
