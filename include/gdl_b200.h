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
/* tuning switches: "conv_halo" / "wgrad_halo" (0/1: 3x3 convs reuse one halo row tile for the 3 horizontal
 * taps), "wgrad_l2_mb" (L2 budget of the wgrad pixel split), "conv_epilogue" (0 direct row stores / 1 smem-transposed),
 * "conv_rows" / "wgrad_rows" (0/1: 3x3 convs with Cout <= 64 use the weight-stationary row-rolling forward kernel / the
 * paired-tap row-streaming weight-gradient kernel), "pdl" (0/1: every kernel is launched with the programmatic-stream-
 * serialization attribute and begins with griddepcontrol.wait — the launch latency between dependent kernels of a stream
 * or a captured graph is hidden, results are unchanged; off until set). */
int gdl_set_option(const char* name, long long value);

/* ---- deterministic reductions: caller-owned scratch workspace ------------------------------------------------------
 * SURVEY.md 8(b) `gdl_query_workspace_bytes`; reference behaviour being matched: torch's BatchNorm / cuDNN wgrad are
 * run-to-run reproducible under torch.use_deterministic_algorithms, which the parity tests of the fused step rely on
 * (tasks_with_models/segmentation_unetplus.py:223-248 replayed from a CUDA graph must equal the eager step bit for bit).
 * With a workspace registered for the current device (and option "deterministic" != 0, the default) every cross-block
 * sum — BatchNorm statistics, BN/LN/bias/LayerScale parameter gradients, loss statistics, gradient norms, the pixel
 * split of the weight-gradient kernels — is formed in a fixed order (per-block slots + ordered tree, or an ordered
 * turnstile per output tile) instead of with fp32 atomics.  Without one the atomic paths are used.
 * The ONE exception to "the library keeps no pointer": the registered buffer must stay valid until it is replaced or
 * cleared (ptr = NULL), and every launch that uses it must be stream-ordered with the others on that device.
 * gdl_set_workspace zeroes the first 256 KiB (tickets) on `stream`. */
long long gdl_query_workspace_bytes(void);
int gdl_set_workspace(void* ptr, long long bytes, void* stream);

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
  int relu;          /* activation: 0 none, 1 ReLU, 2 exact-erf GELU (timm Mlp / nn.GELU) */
  /* --- optional extensions (zero = off) ---------------------------------------------------------
   * residual: added in the epilogue before the activation (x + proj(...) / x + fc2(...) of
   *   mix_transformer.py:218-221); res_dtype GDL_F32 (fp32 residual stream) or 16-bit, stride ldr.
   * w_ld / w_rows: row stride (elements) and row count of the weight matrix when it is a slice of a
   *   wider tensor (default: dense [Cout][R*S*Ctot]).
   * w_rows_per_img: batched B operand — image n uses weight rows [n*w_rows_per_img, ...): one GEMM per
   *   (image, head) for q.k^T and dP = dO.v^T (mix_transformer.py:151-155).
   * w_mn_major: the weight matrix is stored [K rows][Cout cols] (Cout = 16/32/64 contiguous), i.e. the
   *   B operand is used transposed without a copy: P.V and dS.K of the attention. */
  const void* residual;
  int res_dtype;
  int ldr;
  const float* oscale; /* optional fp32 [Cout]: out = act((acc + bias) * oscale + residual) — timm LayerScale gamma */
  int w_ld;
  int w_rows;
  int w_rows_per_img;
  int w_mn_major;
  /* grouped (multi-head) launch of a pointwise GEMM: `groups` independent products in ONE launch, group g reading the
   * source channels [g*g_src_stride, +src[0].channels), the weight matrix shifted by g*g_w_stride elements along its
   * contiguous dimension, and writing the output channels [g*g_out_stride, +Cout).  All attention heads of
   * mix_transformer.py:139-160 / timm Attention in one launch.  0 or 1 = not grouped; needs R = S = 1, one source, no
   * bias / oscale / residual. */
  int groups;
  int g_src_stride;
  int g_w_stride;
  int g_out_stride;
  /* optional: BatchNorm batch statistics of the (16-bit rounded) output, produced by the conv itself — the training-mode
   * half of ConvModule's Conv2d -> BatchNorm2d (models/utils.py:10-52, smp Conv2dReLU):
   *   bn_sums[c] = sum(out[..., c] - bn_pivot[c]),  bn_sums[Cout + c] = sum((out[..., c] - bn_pivot[c])^2)
   * exactly what gdl_bn_stats computes from the stored tensor (bn_pivot may be NULL = 0).  Where the epilogue stages its
   * tile in shared memory for a TMA store the sums are taken from that tile (no second pass over the tensor in HBM);
   * otherwise the library runs the statistics kernel after the convolution.  NULL = not wanted. */
  float* bn_sums;
  const float* bn_pivot;
} gdl_conv_fwd_t;
int gdl_conv2d_nhwc_fwd(const gdl_conv_fwd_t* d, void* stream);
/* Plans the same launch without running it: *fused = 1 when d->bn_sums would come out of the conv's own epilogue, 0 when
 * the statistics kernel would run after the conv (a caller that times the convolution alone then runs gdl_bn_stats itself). */
int gdl_conv2d_bn_fusable(const gdl_conv_fwd_t* d, int* fused);

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
  /* optional (zero = dense / not batched): dw row stride in elements; batched = one independent
   * dw per image at dw + n*dw_img_stride (attention: dV = P^T dO, dK = dS^T q per image and head). */
  int dw_ld;
  long long dw_img_stride;
  int batched;
  /* optional, batched 1x1 only (0 or 1 = not grouped): `groups` independent products per image in one launch — group g reads the
   * source channels shifted by g*g_src_stride, the dy channels shifted by g*g_dy_stride and adds into dw + g*g_dw_stride
   * (all attention heads of dV = P^T dO / dK = dS^T q at once). */
  int groups;
  int g_src_stride;
  int g_dy_stride;
  int g_dw_stride;
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

/* Every packed operand of a model refreshed in ONE launch (the fused trainer calls it after its Adam step instead of
 * ~150-200 gdl_pack_conv_weight launches at the start of the next step; same element mapping as gdl_pack_conv_weight).
 * table_dev: DEVICE array of n_entries descriptors; chunk0_dev: DEVICE int[n_entries + 1], chunk0[e] = first 4096-element
 * chunk of entry e (prefix sums of ceil(rows * dst_ld / 4096)), chunk0[n_entries] = total_chunks. */
typedef struct {
  const float* src; /* fp32 OIHW master                                  */
  void* dst;        /* 16-bit packed operand, rows x dst_ld              */
  int Cout, Cin, R, S;
  int mode;         /* 0 / 1 / 2 as gdl_pack_conv_weight                 */
  int dst_ld;       /* row stride of dst (>= columns, zero padded)       */
} gdl_repack_t;
int gdl_repack_weights(const gdl_repack_t* table_dev, const int* chunk0_dev, int n_entries, int total_chunks, int dtype,
                       void* stream);
/* grad fp32 [Cout][src_ld] (columns (r,s,c), from wgrad) -> fp32 OIHW; dst = (accumulate? dst:0)+src */
int gdl_unpack_conv_wgrad(const float* src, float* dst, int Cout, int Cin, int R, int S, int src_ld,
                          int accumulate, void* stream);

/* ---- pixel-packed form of a narrow conv (engine-internal re-layout; no reference counterpart: the reference calls
 * cuDNN through torch.nn.Conv2d, segmentation_models_pytorch decoder blocks with 16/32 channels) ----
 * A Rx3 / pad-1 conv over Ci channels on (N,H,W,Ci) is the same memory as an Rx3 conv over f*Ci channels on
 * (N,H,W/f,f*Ci) with block-Toeplitz weights.  src: 16-bit packed weights [Co][R][3][Ci] (mode 0 or mode 1 output of
 * gdl_pack_conv_weight); dst: [f*Co][R][3][f*Ci], dst[(j,o)][ky][sx][(j',c)] = src[o][ky][f*(sx-1)+j'-j+1][c] or 0. */
int gdl_widen_conv_weight(const void* src, void* dst, int Co, int Ci, int R, int f, int dtype, void* stream);
/* fp32 gradient of widened weights [f*src_co][src_ld] (rows (j,o), o < src_co; columns (ky,sx,(j',c))) -> fp32 OIHW
 * [Co][Ci][R][3] of the conv; src_co >= Co when the output channels were zero-padded (0 = Co) */
int gdl_fold_widened_wgrad(const float* src, int src_ld, int src_co, float* dst, int Co, int Ci, int R, int f,
                           int accumulate, void* stream);

/* =============================================================================================
 * HBM-bound kernels (csrc/elementwise.cu).  16-bit NHWC activations, C % 8 == 0 unless noted;
 * `dtype` = GDL_BF16 / GDL_F16.
 * ============================================================================================= */

/* Patch normalise + layout change: y = ((x / image_max) - mean[c]) / std[c] written as 16-bit NHWC
 * with pixel stride ld (channels >= C are zero).  Replaces utils/tensors.py:10-35
 * (`normalization`, `standardization`) and their call sites datasets/wds_dataset.py:230-236,
 * datasets/csv_dataset.py:149-153 (true fp32 divisions, same order).  image_max <= 0 skips the
 * first step, mean == NULL the second (then it is a pure cast/transposition of batch["image"]).
 * in_kind: 0 = uint8 NHWC, 1 = f32 NHWC, 2 = f32 NCHW, 3 = uint8 NCHW. */
int gdl_normalize_to_nhwc(const void* x, int in_kind, void* y, int out_dtype, long long N, long long H,
                          long long W, int C, int ld, const float* mean, const float* stdv, float image_max,
                          void* stream);

/* GPU-side batch augmentation fused with the patch normalisation.  Replaces the kornia pipeline every task builds in
 * `_apply_aug` and runs on the HOST in `on_before_batch_transfer` (segmentation_segformer.py:95-125,206-216;
 * segmentation_unetplus.py:92-122; segmentation_dofa.py:91-121): RandomHorizontalFlip / RandomVerticalFlip /
 * RandomRotation90(times 1..3) / RandomResizedCrop (image: bilinear, align_corners=False; mask: nearest), one of them
 * per batch (`random_apply=1`), each applied per sample with its own p.  The random draws stay on the host
 * (gdl_b200/augment.py); the kernel applies them:  params = int32 [N][6] = {op, k, y0, x0, ch, cw} per sample,
 *   op 0 identity | 1 horizontal flip | 2 vertical flip | 3 torch.rot90 by k quarter turns (H == W) |
 *   4 crop rows y0..y0+ch-1, columns x0..x0+cw-1, resized back to (H, W) with ATen's interpolate index arithmetic.
 * x: in_kind as gdl_normalize_to_nhwc (0 uint8 NHWC, 1 f32 NHWC, 2 f32 NCHW, 3 uint8 NCHW); y: out_dtype GDL_BF16/GDL_F16 ->
 * 16-bit NHWC with pixel stride ld (channels >= C zero, same arithmetic as gdl_normalize_to_nhwc: op 0 is bit-identical
 * to it); GDL_F32 -> f32 NCHW (ld unused).  mask / mask_out: (N,H,W) int64 (mask_kind 0) or uint8 (1), both or neither.
 * Not in place. */
int gdl_augment_normalize(const void* x, int in_kind, const void* mask, int mask_kind, const int* params, void* y,
                          int out_dtype, void* mask_out, long long N, long long H, long long W, int C, int ld,
                          const float* mean, const float* stdv, float image_max, void* stream);

/* nn.Dropout2d in training mode (models/decoders/segformer_mlp.py:73,129 before linear_pred; models/heads/fcn_head.py:69-83
 * before cls_seg): y[n][p][c] = x[n][p][c] * mask[n][c], mask (fp32 [N][C]) = 0 or 1 / (1 - p) drawn by the caller.  The same
 * call is the backward (dx = dy * mask).  16-bit NHWC with pixel strides ldx / ldy; may run in place. */
int gdl_dropout2d_apply(const void* x, long long ldx, const float* mask, void* y, long long ldy, int dtype, long long N,
                        long long HW, int C, void* stream);

/* im2col / col2im for strided convs and the 3/4/6-band stem (torchvision ResNet conv1 7x7/2, the
 * 3x3/2 and 1x1/2 convs of layer2-4): col[(n,ho,wo)][(r,s,c)] zero padded to Kpad columns; the conv
 * itself is then gdl_conv2d_nhwc_fwd with R=S=1.  col2im is the gather-form adjoint (C % 8 == 0). */
int gdl_im2col_nhwc(const void* x, void* col, int dtype, int N, int H, int W, int C, int ld, int R, int S,
                    int stride, int pad, int Kpad, void* stream);
int gdl_col2im_nhwc(const void* dcol, void* dx, int dtype, int N, int H, int W, int C, int ld, int R, int S,
                    int stride, int pad, int Kpad, void* stream);

/* ---- SyncBatchNorm statistics over NVLink peer memory ------------------------------------------------------------
 * Replaces the per-layer statistics collectives of torch SyncBatchNorm (every shipped YAML: `sync_batchnorm: true`,
 * configs/segformer_config_RGB.yaml:6-14): sums[0..n) := sum over ranks of sums[0..n), added in RANK order (bit-identical on
 * every rank).  peer_bufs (HOST array of `world` device pointers): rank r's exchange buffer of gdl_p2p_exchange_bytes()
 * bytes, zero-initialised, mapped for peer access in this process (torch.distributed._symmetric_memory); counter: one
 * zero-initialised device word per process (the call counter).  One small kernel: store own sums, signal the peers, wait for
 * theirs, read them over NVLink.  Every rank must make the same sequence of calls (as with any collective). */
int gdl_p2p_allreduce_sums(float* sums, int n, const void* const* peer_bufs, int rank, int world, int slot_floats,
                           unsigned* counter, void* stream);
long long gdl_p2p_exchange_bytes(int world, int slot_floats);

/* BatchNorm2d, training mode (nn.BatchNorm2d inside models/utils.py:10-52 ConvModule,
 * segformer_mlp.py:63-72, smp Conv2dReLU, torchvision ResNet): split into
 *   stats    sums[0:C] = sum(x - pivot), sums[C:2C] = sum((x - pivot)^2)   (fp32, zeroed inside)
 *   finalize mean/biased var -> scale = gamma*invstd, shift = beta - mean*scale, saved mean/invstd,
 *            running stats update (momentum, unbiased var); `count` = elements per channel the sums
 *            cover (after an optional cross-rank all-reduce of `sums` = SyncBatchNorm)
 *   apply    y = act(x*scale + shift [+ res | + res*rscale + rshift]); optional second output y_up =
 *            nearest x2 upsample of y (smp DecoderBlock's F.interpolate(scale_factor=2, "nearest"))
 * pivot (fp32 [C], may be NULL, may alias running_mean) only conditions the variance formula. */
int gdl_bn_stats(const void* x, int dtype, long long M, int C, int ld, float* sums, const float* pivot,
                 void* stream);
int gdl_bn_finalize(const float* pivot, const float* sums, long long count, int C, const float* gamma,
                    const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                    float* scale, float* shift, float* save_mean, float* save_invstd, void* stream);
int gdl_bn_eval_coeffs(int C, const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float eps, float* scale, float* shift, void* stream);
int gdl_bn_apply(const void* x, int ldx, const float* scale, const float* shift, const void* res, int ldr,
                 const float* rscale, const float* rshift, int relu, void* y, int ldy, void* y_up, int ldu,
                 int dtype, int N, int H, int W, int C, void* stream);

/* Backward of [concat-consumers] -> ReLU -> BatchNorm in two passes (autograd of the above, a20):
 *   grad_gather  g = (sum_k src_k) * (y > 0); src mode 1 = gradient of the x2-upsampled copy
 *                (2H x 2W), contributes its 2x2 sum.  With sums != NULL also reduces
 *                sums[0:C] = sum g, sums[C:2C] = sum g*(x-mean)*invstd in the same pass.
 *   bn_bwd_apply dx = gamma*invstd*(g - sums[0:C]/count - xhat*sums[C:2C]/count); optional
 *                dgamma = sums[C:2C], dbeta = sums[0:C]. */
int gdl_grad_gather(int num_src, const void* const* src_ptr, const int* src_ld, const int* src_mode,
                    const void* y, int ldy, const void* x, int ldx, const float* mean, const float* invstd,
                    void* g, int ldg, float* sums, int dtype, int N, int H, int W, int C, void* stream);
int gdl_bn_bwd_apply(const void* g, int ldg, const void* x, int ldx, const float* mean, const float* invstd,
                     const float* gamma, const float* sums, void* dx, int ldd, float* dgamma, float* dbeta,
                     int accumulate_param_grads, int dtype, long long M, long long count, int C, void* stream);
int gdl_bn_param_grads(const float* sums, int C, float* dgamma, float* dbeta, int accumulate, void* stream);

/* MaxPool2d(kernel 3, stride 2, pad 1) of the ResNet stem; idx (1 byte/element, may be NULL) holds
 * the winning tap (first maximum in scan order, as ATen) for the backward. */
int gdl_maxpool3x3s2_fwd(const void* x, int ldx, void* y, int ldy, unsigned char* idx, int dtype, int N, int H,
                         int W, int C, void* stream);
int gdl_maxpool3x3s2_bwd(const void* dy, int ldy, const unsigned char* idx, void* dx, int ldx, int dtype, int N,
                         int H, int W, int C, void* stream);

/* =============================================================================================
 * losses, eval post-processing, optimizer (csrc/loss_optim.cu).  logits: fp32 [M][ld], K classes.
 * ============================================================================================= */

/* loss = w_ce * CE + w_dice * Dice over M pixels (training_step's `self.loss(y_hat, y)`,
 * segmentation_segformer.py:226-229; torch.nn.CrossEntropyLoss / smp SoftCrossEntropyLoss /
 * smp DiceLoss semantics, see csrc/loss_optim.cu).  K == 1 is the binary (sigmoid) mode.
 * target_kind: 0 = int64, 1 = uint8.  stats: 4+3K floats, coeff: 2+2K floats, coeff[0] = loss.
 * bwd writes d(loss)/d(logits) * grad_scale[0] as [M][ldd] of out_dtype (columns >= K untouched). */
int gdl_seg_loss_fwd(const float* logits, int ld, const void* target, int target_kind, long long M, int K,
                     long long ignore_index, int has_ignore, float w_ce, float w_dice, float label_smoothing,
                     int ce_mean_over_all, float dice_smooth, float dice_eps, float* stats, float* coeff,
                     void* stream);
int gdl_seg_loss_bwd(const float* logits, int ld, const void* target, int target_kind, long long M, int K,
                     long long ignore_index, int has_ignore, float w_ce, float w_dice, float label_smoothing,
                     int ce_mean_over_all, float dice_smooth, float dice_eps, const float* coeff,
                     const float* grad_scale, void* dlogits, int ldd, int out_dtype, void* stream);

/* softmax(dim=1).argmax(dim=1) (K > 1) or sigmoid > threshold (K == 1):
 * segmentation_segformer.py:268-271, segmentation_unetplus.py test/validation steps. */
int gdl_argmax_classes(const float* logits, int ld, long long M, int K, float threshold, long long* out,
                       void* stream);

/* ---- fused head: bilinear upsample (align_corners=False) + loss / argmax, SURVEY.md 8(b) gdl_upsample_ce_* ----------
 * Replaces  F.interpolate(logits, size=image.shape[2:], mode="bilinear", align_corners=False)
 * (models/segmentation/segformer.py:47-57, models/segmentation/dofa.py:90-105) FOLLOWED BY the loss of training_step
 * (tasks_with_models/segmentation_segformer.py:218-243; CrossEntropy / smp Dice / SoftCE as gdl_seg_loss_*) or by
 * softmax(dim=1).argmax(dim=1) | sigmoid > threshold (segmentation_segformer.py:268-271,288-291).  The (N,K,H,W) fp32
 * logits are never materialised: each full-resolution pixel is interpolated from its 4 low-resolution taps in registers.
 * logits_lr: fp32 NHWC [N][h][w][ld] (ld >= K); target [N][H][W] (target_kind 0 = int64, 1 = uint8); stats / coeff as
 * gdl_seg_loss_fwd (coeff[0] = loss).  _bwd writes d(loss)/d(logits_lr) * grad_scale[0] as [N][h][w][ldd] of
 * out_dtype (channels >= K are left untouched: pre-zero a padded operand); gather form, no atomics. */
int gdl_upsample_ce_fwd(const float* logits_lr, int ld, int N, int h, int w, int H, int W, const void* target,
                        int target_kind, int K, long long ignore_index, int has_ignore, float w_ce, float w_dice,
                        float label_smoothing, int ce_mean_over_all, float dice_smooth, float dice_eps, float* stats,
                        float* coeff, void* stream);
int gdl_upsample_ce_bwd(const float* logits_lr, int ld, int N, int h, int w, int H, int W, const void* target,
                        int target_kind, int K, long long ignore_index, int has_ignore, float w_ce, float w_dice,
                        float label_smoothing, int ce_mean_over_all, float dice_smooth, float dice_eps,
                        const float* coeff, const float* grad_scale, void* dlogits_lr, int ldd, int out_dtype,
                        void* stream);
int gdl_upsample_argmax(const float* logits_lr, int ld, int N, int h, int w, int H, int W, int K, float threshold,
                        long long* out, void* stream);

/* The same post-processing fused with the confusion counts of the evaluation metric: torchmetrics
 * `MeanIoU(num_classes, per_class=True, input_format="index")` as used by the three tasks' test_step
 * (segmentation_segformer.py:78-92,283-296) reduces per-SAMPLE intersection / union counts, so the counts are kept per
 * sample: conf[n][t][p] (int64, Kc x Kc with Kc = max(K, 2)) += #pixels of sample n with target t predicted p; targets
 * equal to ignore_index (when has_ignore) or outside [0, Kc) are skipped.  logits: fp32 [N*HW][ld]; target_kind 0 = int64,
 * 1 = uint8 (as gdl_seg_loss_fwd); classes (int64 [N*HW]) and (target, conf) are each optional.  conf must be zeroed by the caller. */
int gdl_argmax_confusion(const float* logits, int ld, long long N, long long HW, int K, float threshold,
                         const void* target, int target_kind, long long ignore_index, int has_ignore,
                         long long* classes, long long* conf, void* stream);

/* torch.optim.Adam step on flat fp32 buffers; g is multiplied by grad_scale[0] first (may be NULL).
 * gdl_grad_clip_coef: scale[0] = min(1, max_norm / (||g||_2 + 1e-6))  (clip_grad_norm_). */
int gdl_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, const float* grad_scale, void* stream);
/* same update with the step counter on the device: state[0] = step, state[1..2] = bias corrections, advanced
 * by the call itself (nothing step-dependent in kernel parameters: the step can live in a CUDA graph).
 * lr_scale (device, may be NULL): the step uses lr * lr_scale[0] — how a learning-rate scheduler (the reference
 * configures ReduceLROnPlateau / OneCycleLR in its YAML files) reaches a step that is replayed from a captured graph. */
int gdl_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                      float eps, float weight_decay, float* state, const float* grad_scale, const float* lr_scale,
                      void* stream);
int gdl_grad_clip_coef(const float* g, long long n, float max_norm, float* sumsq_scratch, float* scale,
                       void* stream);

/* =============================================================================================
 * MixTransformer / SegFormer HBM-bound kernels (csrc/transformer.cu).  Tokens (B,N,C) and maps
 * (B,h,w,C) are the same NHWC memory, so the reference's reshape/permute/contiguous copies
 * (mix_transformer.py:499,541-546; segformer_mlp.py:78-121) do not exist here.
 * ============================================================================================= */

/* nn.LayerNorm over the last dim (eps 1e-6 for block/stage norms, 1e-5 for OverlapPatchEmbed.norm and
 * Attention.norm: mix_transformer.py:100,251,611).  x: fp32 (residual stream) or 16-bit, y: 16-bit or
 * fp32; mean/rstd [M] saved for the backward.  bwd: dx (+ add) as fp32 and/or a 16-bit copy; pgrads
 * [2][C] = (dgamma, dbeta) accumulated (zero first). */
int gdl_layernorm_fwd(const void* x, int x_dtype, long long ldx, const float* gamma, const float* beta, float eps,
                      void* y, int y_dtype, long long ldy, float* mean, float* rstd, long long M, int C,
                      void* stream);
int gdl_layernorm_bwd(const void* g, int g_dtype, long long ldg, const void* x, int x_dtype, long long ldx,
                      const float* mean, const float* rstd, const float* gamma, const float* add, long long lda,
                      float* dx32, long long ld32, void* dx16, int dx16_dtype, long long ld16, float* pgrads,
                      long long M, int C, void* stream);

/* attention scores: p = softmax(scale * s) over rows of length L (mix_transformer.py:151-152); columns
 * [L, Lpad) of p are written as zeros so p can be the K-padded operand of the P.V GEMM.
 * bwd: ds = scale * p * (dp - sum_j dp_j p_j). */
int gdl_softmax_fwd(const void* s, long long lds, float scale, void* p, long long ldp, int dtype, long long M,
                    int L, int Lpad, void* stream);
int gdl_softmax_bwd(const void* p, long long ldp, const void* dp, long long lddp, float scale, void* ds,
                    long long ldds, int dtype, long long M, int L, int Lpad, void* stream);

/* Fused spatial-reduction attention forward (Attention.forward, mix_transformer.py:131-159): for every image b and head g
 *   o[b, :, g*64:(g+1)*64] = softmax(scale * q_g . k_g^T) . v_g
 * in ONE kernel (scores stay in TMEM / shared memory).  q: (B, N, c) tokens, row stride ldq; kv: (B*nk, 2c) rows = reduced
 * tokens, K in columns [0, c), V in [c, 2c), row stride ldkv; o: (B, N, c), row stride ldo.  p_out (optional, training):
 * (B, N, heads*nk) normalised probabilities of head g in columns [g*nk, (g+1)*nk) — what the backward's dV = P^T.dO and
 * softmax_bwd read.  Head dim 64 (c == 64*heads).  With p_out: nk a multiple of 64 up to 256 (every MiT stage of a 256x256 or
 * 512x512 tile), any N; without (inference): any N and nk.  Anything else returns GDL_ERR_UNSUPPORTED and the caller uses
 * gdl_conv2d_nhwc_fwd + gdl_softmax_fwd + gdl_conv2d_nhwc_fwd. */
int gdl_sra_attention_fwd(const void* q, long long ldq, const void* kv, long long ldkv, void* o, long long ldo,
                          void* p_out, long long ldp, int B, int N, int heads, int nk, int c, float scale, int dtype,
                          void* stream);

/* The query-tile-local half of the attention backward in one kernel (same pipeline, roles swapped): dP = dO.V^T in TMEM ->
 * dS = scale * P * (dP - rowsum(P * dP)) written over the TMA-loaded P tile in shared memory -> dQ = dS.K.  Outputs dq (B, N, c) and
 * ds (B, N, heads*nk) 16-bit (ds feeds the dK = dS^T.q weight-gradient launch; dV = P^T.dO reads the saved P): replaces
 * gdl_conv2d_nhwc_fwd (dP) + gdl_softmax_bwd + gdl_conv2d_nhwc_fwd (dQ); dP never reaches HBM.  Same shape limits as the forward
 * with p_out. */
int gdl_sra_attention_bwd(const void* d_o, long long lddo, const void* kv, long long ldkv, const void* p_saved, long long ldp,
                          void* dq, long long lddq, void* ds, long long ldds, int B, int N, int heads, int nk, int c, float scale,
                          int dtype, void* stream);

/* The same kernel for plain multi-head self-attention with many keys (timm's Attention inside the DOFA ViT blocks,
 * dofa_v2.py:445-487 -> timm vision_transformer.Block): keys are streamed in blocks of 128 with the online-softmax recurrence, so
 * neither scores nor probabilities exist in HBM.  Forward only (no probabilities are saved).  q / k / v / o point at head 0's 64
 * columns of token 0 of image 0 (e.g. qkv, qkv + c, qkv + 2c of a fused (B*N, 3c) projection with ld = 3c); head g sits 64*g
 * columns to the right; images are N rows apart.  Head dim 64; any N (ragged query tiles and key blocks are zero-filled / clipped
 * by the TMA unit and masked in the softmax). */
int gdl_mha_flash_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* o,
                      long long ldo, int B, int N, int heads, float scale, int dtype, void* stream);

/* Mix-FFN middle: y = GELU(depthwise3x3(x) + b) (Mlp.dwconv + act, mix_transformer.py:56-63,533-546), exact
 * erf GELU; `pre` keeps the 16-bit pre-activation for the backward.  w: fp32 [C][3][3].
 * bwd: dx and pgrads [C][10] (9 taps + bias, accumulated; zero first); dpre_scratch: [M][C] 16-bit. */
int gdl_dwconv3x3_gelu_fwd(const void* x, int ldx, const float* w, const float* bias, void* pre, void* y, int dtype,
                           int N, int H, int W, int C, void* stream);
int gdl_dwconv3x3_gelu_bwd(const void* dy, const void* pre, const void* x, int ldx, const float* w,
                           void* dpre_scratch, void* dx, int lddx, float* pgrads, int dtype, int N, int H, int W,
                           int C, void* stream);

/* F.interpolate(mode="bilinear", align_corners=False) on NHWC (segformer.py:51-57, segformer_mlp.py:88-119,
 * dofa.py:90-105) for 16-bit features or fp32 logits, and its adjoint in gather form. */
int gdl_bilinear_fwd(const void* x, long long ldx, void* y, long long ldy, int dtype, int N, int Hi, int Wi, int Ho,
                     int Wo, int C, void* stream);
int gdl_bilinear_bwd(const void* dy, long long ldy, void* dx, long long ldx, int dtype, int N, int Hi, int Wi,
                     int Ho, int Wo, int C, void* stream);

/* out = base + sum_i resize_i(src_i): `num_src` (<= 3) low-resolution maps, each bilinearly resized (align_corners=False)
 * to (Ho, Wo), added to the optional full-resolution `base` in fp32 in a fixed order and rounded once (16-bit NHWC, C % 8 ==
 * 0).  The SegFormer decoder's fuse layer (segformer_mlp.py:77-128: four projections, three resizes to the stride-4 grid,
 * concat, 1x1 conv without bias) with the 1x1 conv moved in front of the resizes — both are linear, so
 * conv1x1(concat_i resize(y_i)) == sum_i resize(conv1x1_i(y_i)) — which never writes the three resized 768-channel maps
 * and runs the 3072 -> 768 product at each map's own resolution.  Adjoint: gdl_bilinear_bwd per source. */
typedef struct {
  const void* ptr; /* (N, H, W, C) 16-bit */
  int H, W;
  long long ld;    /* elements between pixels */
} gdl_lowres_t;
int gdl_bilinear_sum_fwd(const void* base, long long ld_base, int num_src, const gdl_lowres_t* src, void* out,
                         long long ld_out, int dtype, int N, int Ho, int Wo, int C, void* stream);

/* nn.AdaptiveAvgPool2d(S) of the UperNet pyramid pooling module (models/utils.py:55-93) and its adjoint;
 * y/dy are dense [N][S][S][C]. */
int gdl_adaptive_avgpool_fwd(const void* x, long long ldx, void* y, int dtype, int N, int H, int W, int C, int S,
                             void* stream);
int gdl_adaptive_avgpool_bwd(const void* dy, void* dx, int dtype, int N, int H, int W, int C, int S, void* stream);
/* y = a + b on NHWC rows (UperNet top-down path, upernet.py:128-135) */
int gdl_add_nhwc(const void* a, long long lda, const void* b, long long ldb, void* y, long long ldy, int dtype,
                 long long M, int C, void* stream);

/* DOFA ViT token glue (dofa_v2.py:445-468): tokens[b] = [cls ; patch[b] + pos_embed[1:]] (fp32 stream), and the
 * feature tap tokens[:, 1:, :] -> dense (B, P, C) map in `feat_dtype`. */
int gdl_vit_assemble_tokens(const void* patch, int patch_dtype, const float* pos, const float* cls, float* tokens,
                            int B, int P, int C, void* stream);
int gdl_vit_extract_feature(const float* tokens, void* feat, int feat_dtype, int B, int P, int C, void* stream);

/* Gradient of a feature tap into the fp32 stream gradient g (B, P+1, C): init != 0 starts it (g[b][0] = 0,
 * g[b][1+p] = dfeat[b][p]), init == 0 accumulates (g[b][1+p] += dfeat[b][p]).  Backward of gdl_vit_extract_feature. */
int gdl_vit_feature_grad(const void* dfeat, int dtype, float* g, int B, int P, int C, int init, void* stream);

/* ViT block pieces needed when the DOFA encoder is TRAINED (timm.models.vision_transformer.Block as used by
 * dofa_v2.py:249-263: x = x + drop_path(ls(attn(norm1(x)))); x = x + drop_path(ls(mlp(norm2(x))))).  The forward-only
 * path fuses GELU and LayerScale + residual into the GEMM epilogues; the training path keeps the pre-activation and the
 * un-scaled branch output for the backward:
 *   gelu_fwd / gelu_bwd:  y = gelu(x) exact (erf);  dpre = dy * gelu'(pre)            (16-bit, n % 8 == 0)
 *   layerscale_add:       out[r][c] = res[r][c] + s(r) * gamma[c] * u[r][c]           (fp32 stream, u 16-bit [M][C])
 *   layerscale_bwd:       du[r][c] = s(r) * gamma[c] * g[r][c] (16-bit);  dgamma[c] += sum_r s(r) * g[r][c] * u[r][c]
 * s(r) = sscale[r / rows_per_sample] is timm's DropPath (per-sample keep mask / keep probability); sscale == NULL -> 1.
 * dgamma (fp32 [C], may be NULL) is accumulated with atomics: zero it first. */
int gdl_gelu_fwd(const void* x, void* y, int dtype, long long n, void* stream);
int gdl_gelu_bwd(const void* dy, const void* pre, void* dpre, int dtype, long long n, void* stream);
int gdl_layerscale_add(const float* res, const void* u, int dtype, const float* gamma, const float* sscale,
                       long long rows_per_sample, float* out, long long M, int C, void* stream);
int gdl_layerscale_bwd(const float* g, const void* u, int dtype, const float* gamma, const float* sscale,
                       long long rows_per_sample, void* du, float* dgamma, long long M, int C, void* stream);

/* DynamicChannelEmbed of DynamicMixTransformer (models/encoders/mix_transformer.py:762-865), the parts that are not GEMMs:
 * the softmax over the C <= 16 input bands of the channel-attention logits and the attention-weighted sum of the per-band
 * feature maps (:843-848).  xw: 16-bit [P][C*E] (band c in columns c*E..c*E+E-1), scores: fp32 [P][16] (columns >= C
 * unused), out: 16-bit [P][E], attn: fp32 [P][16] (kept for the backward, may be NULL at inference).
 *   fwd: a = softmax_c(scores); out[p][e] = sum_c a[c] xw[p][c*E+e]
 *   bwd: dxw[p][c*E+e] = a[c] dout[p][e]; dscores[p][c] = a[c] (dattn[c] - sum_k a[k] dattn[k]), dattn[c] = sum_e dout[p][e] xw[p][c*E+e]
 *        (dscores 16-bit [P][16], columns >= C zero: the dY operand of the score conv's backward)
 * E % 8 == 0 and E / 8 a power of two <= 32 (MiT stage-1 widths 32 / 64).  relu_bwd: dx = dy * (y > 0), n % 8 == 0. */
int gdl_channel_pool_fwd(const void* xw, const float* scores, void* out, float* attn, int dtype, long long P, int C, int E,
                         void* stream);
int gdl_channel_pool_bwd(const void* dout, const void* xw, const float* attn, void* dxw, void* dscores, int dtype, long long P,
                         int C, int E, void* stream);
int gdl_relu_bwd(const void* dy, const void* y, void* dx, int dtype, long long n, void* stream);

/* fp32 -> dtype cast of a flat buffer (residual-stream gradient -> 16-bit GEMM operand) */
int gdl_cast_f32(const float* x, void* y, int dtype, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GDL_B200_H_ */
