// loss_optim.cu — per-pixel segmentation losses (softmax cross-entropy with optional label
// smoothing, soft Dice, or a weighted sum) on NHWC fp32 logits, their fused backward, and a
// flat fused Adam step.  All HBM-bound: one pass to reduce, one pass to emit d(logits).
//
// Loss semantics restate (a17 in SURVEY.md §8):
//   torch.nn.CrossEntropyLoss (mean over non-ignored pixels, label_smoothing),
//   smp.losses.SoftCrossEntropyLoss (mean over ALL pixels, ignored ones contribute 0),
//   smp.losses.DiceLoss(mode="multiclass"|"binary", from_logits=True, smooth, eps=1e-7, log_loss=False).
#include <math.h>
#include <string.h>

#include <type_traits>

#include "../../include/gdl_b200.h"
#include "common.cuh"
#include "det_reduce.cuh"
#include "loss_cfg.cuh"

namespace gdl {

template <int KMAX, typename TT>
__global__ void seg_loss_stats_kernel(const float* __restrict__ logits, int ld, const TT* __restrict__ target,
                                      long long M, LossCfg cfg, float* __restrict__ stats, const DetCtx det) {
  GDL_PDL_ENTRY();
  const int K = cfg.K;
  LossAcc<KMAX> acc;
  acc.init();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M;
       i += (long long)gridDim.x * blockDim.x) {
    float z[KMAX];
#pragma unroll
    for (int c = 0; c < KMAX; ++c) z[c] = c < K ? logits[i * ld + c] : -INFINITY;
    acc.add(z, load_target(target, i), cfg);
  }
  acc.commit(cfg, stats, det);
}

__global__ void seg_loss_finalize_kernel(const float* __restrict__ stats, LossCfg cfg, float* __restrict__ coeff) {
  GDL_PDL_ENTRY();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int K = cfg.K;
  float loss = 0.f;
  float denom = cfg.ce_mean_over_all ? stats[3] : stats[2];
  denom = fmaxf(denom, 1.f);
  coeff[1] = denom;
  if (cfg.w_ce != 0.f) {
    const float eps = cfg.label_smoothing;
    const float ce = (1.f - eps) * stats[0] / denom + (cfg.binary ? 0.f : eps / (float)K * stats[1] / denom);
    loss += cfg.w_ce * ce;
  }
  for (int c = 0; c < K; ++c) {
    float a = 0.f, b = 0.f;
    if (cfg.w_dice != 0.f) {
      const float inter = stats[4 + c], card = stats[4 + K + c], tsum = stats[4 + 2 * K + c];
      const float num = 2.f * inter + cfg.dice_smooth;
      const float den_raw = card + cfg.dice_smooth;
      const float den = fmaxf(den_raw, cfg.dice_eps);
      const float mask = tsum > 0.f ? 1.f : 0.f;
      loss += cfg.w_dice * mask * (1.f - num / den) / (float)K;
      // d(1 - num/den)/dp = -(2 t den - num * dden/dp) / den^2 ; dden/dp = 1 unless clamped
      const float dd = den_raw > cfg.dice_eps ? 1.f : 0.f;
      a = -mask / (float)K * (0.f * den - num * dd) / (den * den);
      b = -mask / (float)K * (2.f * den - num * dd) / (den * den);
    }
    coeff[2 + c] = a;
    coeff[2 + K + c] = b;
  }
  coeff[0] = loss;
}

template <int KMAX, typename TT, typename TO>
__global__ void seg_loss_bwd_kernel(const float* __restrict__ logits, int ld, const TT* __restrict__ target,
                                    long long M, LossCfg cfg, const float* __restrict__ coeff,
                                    const float* __restrict__ grad_scale, TO* __restrict__ dlogits, int ldd) {
  GDL_PDL_ENTRY();
  const int K = cfg.K;
  const float gs = grad_scale ? grad_scale[0] : 1.f;
  const float inv_denom = 1.f / coeff[1];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M;
       i += (long long)gridDim.x * blockDim.x) {
    const long long t = load_target(target, i);
    const bool ign = cfg.has_ignore && t == cfg.ignore_index;
    float z[KMAX], d[KMAX];
#pragma unroll
    for (int c = 0; c < KMAX; ++c) z[c] = c < K ? logits[i * ld + c] : -INFINITY;
    loss_pixel_grad<KMAX>(z, t, ign, cfg, coeff, inv_denom, d);
#pragma unroll
    for (int c = 0; c < KMAX; ++c) {
      if (c < K) {
        if constexpr (std::is_same<TO, float>::value)
          dlogits[i * ldd + c] = d[c] * gs;
        else if constexpr (std::is_same<TO, __nv_bfloat16>::value)
          dlogits[i * ldd + c] = __float2bfloat16_rn(d[c] * gs);
        else
          dlogits[i * ldd + c] = __float2half_rn(d[c] * gs);
      }
    }
  }
}

// argmax over classes (eval post-processing, segmentation_segformer.py:268-271): int64 class map.
// softmax is monotone, so argmax(logits) == argmax(softmax(logits)); first maximum wins (torch).
__global__ void argmax_kernel(const float* __restrict__ logits, int ld, long long M, int K, float threshold,
                              long long* __restrict__ out) {
  GDL_PDL_ENTRY();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M;
       i += (long long)gridDim.x * blockDim.x) {
    if (K == 1) {
      const float p = 1.f / (1.f + expf(-logits[i * ld]));
      out[i] = p > threshold ? 1 : 0;
    } else {
      float best = logits[i * ld];
      int bi = 0;
      for (int c = 1; c < K; ++c) {
        const float v = logits[i * ld + c];
        if (v > best) {
          best = v;
          bi = c;
        }
      }
      out[i] = bi;
    }
  }
}

// ------------------------------------------------------------------------------------------
// fused Adam over a flat fp32 parameter buffer (torch.optim.Adam semantics, no amsgrad)
// ------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                            float weight_decay, float bc1, float bc2_sqrt, const float* __restrict__ grad_scale) {
  GDL_PDL_ENTRY();
  const float gsc = grad_scale ? grad_scale[0] : 1.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * gsc;
    const float pi = p[i];
    if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
    const float mi = m[i] + (gi - m[i]) * (1.f - beta1);
    const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// device-side step counter variant (CUDA-graph friendly: nothing step-dependent is baked into kernel
// parameters).  state[0] = step (as float), state[1] = 1 - beta1^step, state[2] = sqrt(1 - beta2^step)
__global__ void adam_advance_kernel(float* __restrict__ state, float beta1, float beta2) {
  GDL_PDL_ENTRY();
  const float step = state[0] + 1.f;
  state[0] = step;
  state[1] = 1.f - powf(beta1, step);
  state[2] = sqrtf(1.f - powf(beta2, step));
}

__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                                float weight_decay, const float* __restrict__ state,
                                const float* __restrict__ grad_scale, const float* __restrict__ lr_scale) {
  GDL_PDL_ENTRY();
  const float gsc = grad_scale ? grad_scale[0] : 1.f;
  const float bc1 = state[1], bc2_sqrt = state[2];
  if (lr_scale != nullptr) lr *= lr_scale[0];  // learning-rate schedule as a device value: a captured graph follows it
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * gsc;
    const float pi = p[i];
    if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
    const float mi = m[i] + (gi - m[i]) * (1.f - beta1);
    const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// sum of squares of a flat fp32 buffer (for clip_grad_norm_); result accumulated into out[0]
__global__ void sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out, const DetCtx det) {
  GDL_PDL_ENTRY();
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    s = fmaf(g[i], g[i], s);
  s = warp_sum(s);
  __shared__ float shw[8];  // blockDim.x == 256: warp partials added in warp order
  if ((threadIdx.x & 31) == 0) shw[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += shw[w];
    if (det.s0 != nullptr) det_put(det, 1, 0, v);
    else atomicAdd(out, v);
  }
  if (det.s0 != nullptr) det_finish(det, 1, out);
}

// scale[0] = min(1, max_norm / (sqrt(sumsq) + 1e-6))  (torch.nn.utils.clip_grad_norm_)
__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float max_norm, float* __restrict__ scale) {
  GDL_PDL_ENTRY();
  const float nrm = sqrtf(sumsq[0]);
  const float c = max_norm / (nrm + 1e-6f);
  scale[0] = c < 1.f ? c : 1.f;
}

int loss_blocks(long long M) {
  long long b = (M + 255) / 256;
  if (b > 4 * kNumSMsB200) b = 4 * kNumSMsB200;
  return (int)(b < 1 ? 1 : b);
}

void launch_loss_finalize(const float* stats, const LossCfg& cfg, float* coeff, cudaStream_t s) {
  GDL_LAUNCH(seg_loss_finalize_kernel, 1, 32, 0, s, stats, cfg, coeff);
}

int make_cfg(LossCfg* c, int K, long long ignore_index, int has_ignore, float w_ce, float w_dice,
             float label_smoothing, int ce_mean_over_all, float dice_smooth, float dice_eps) {
  GDL_REQUIRE(K >= 1 && K <= kLossMaxK, GDL_ERR_UNSUPPORTED, "loss: number of classes %d outside [1,%d]", K, kLossMaxK);
  c->K = K;
  c->binary = K == 1;
  c->ignore_index = ignore_index;
  c->has_ignore = has_ignore;
  c->w_ce = w_ce;
  c->w_dice = w_dice;
  c->label_smoothing = label_smoothing;
  c->ce_mean_over_all = ce_mean_over_all;
  c->dice_smooth = dice_smooth;
  c->dice_eps = dice_eps;
  return 0;
}

}  // namespace gdl

using namespace gdl;

// target_kind: 0 = int64, 1 = uint8
extern "C" int gdl_seg_loss_fwd(const float* logits, int ld, const void* target, int target_kind, long long M,
                                int K, long long ignore_index, int has_ignore, float w_ce, float w_dice,
                                float label_smoothing, int ce_mean_over_all, float dice_smooth, float dice_eps,
                                float* stats /* 4+3K */, float* coeff /* 2+2K; coeff[0] = loss */, void* stream) {
  GDL_REQUIRE(logits && target && stats && coeff && M > 0 && ld >= K, GDL_ERR_INVALID, "seg_loss_fwd: bad args");
  LossCfg cfg;
  int st = make_cfg(&cfg, K, ignore_index, has_ignore, w_ce, w_dice, label_smoothing, ce_mean_over_all,
                    dice_smooth, dice_eps);
  if (st) return st;
  cudaStream_t s = (cudaStream_t)stream;
  GDL_CHECK_CUDA(cudaMemsetAsync(stats, 0, (4 + 3 * (size_t)K) * sizeof(float), s));
  int blocks = loss_blocks(M);
  DetCtx det = det_none();
  {
    const DetWs ws = det_workspace();
    if (const int gd = det_grid(ws, blocks, 4 + 3 * K)) {
      blocks = gd;
      det = det_ctx(ws, gd, 4 + 3 * K);
    }
  }
#define LAUNCH_STATS(KMAX)                                                                                  \
  do {                                                                                                      \
    if (target_kind == 0)                                                                                   \
      GDL_LAUNCH((seg_loss_stats_kernel<KMAX, long long>), blocks, 256, 0, s, logits, ld, (const long long*)target, M, cfg, stats, det); \
    else                                                                                                    \
      GDL_LAUNCH((seg_loss_stats_kernel<KMAX, uint8_t>), blocks, 256, 0, s, logits, ld, (const uint8_t*)target, M, cfg, stats, det); \
  } while (0)
  if (K <= 2) LAUNCH_STATS(2);
  else if (K <= 8) LAUNCH_STATS(8);
  else LAUNCH_STATS(32);
#undef LAUNCH_STATS
  GDL_CHECK_CUDA(cudaGetLastError());
  GDL_LAUNCH(seg_loss_finalize_kernel, 1, 32, 0, s, stats, cfg, coeff);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_seg_loss_bwd(const float* logits, int ld, const void* target, int target_kind, long long M,
                                int K, long long ignore_index, int has_ignore, float w_ce, float w_dice,
                                float label_smoothing, int ce_mean_over_all, float dice_smooth, float dice_eps,
                                const float* coeff, const float* grad_scale, void* dlogits, int ldd, int out_dtype,
                                void* stream) {
  GDL_REQUIRE(logits && target && coeff && dlogits && M > 0 && ld >= K && ldd >= K, GDL_ERR_INVALID,
              "seg_loss_bwd: bad args");
  LossCfg cfg;
  int st = make_cfg(&cfg, K, ignore_index, has_ignore, w_ce, w_dice, label_smoothing, ce_mean_over_all,
                    dice_smooth, dice_eps);
  if (st) return st;
  cudaStream_t s = (cudaStream_t)stream;
  const int blocks = loss_blocks(M);
#define LAUNCH_BWD2(KMAX, TT, TO) \
  GDL_LAUNCH((seg_loss_bwd_kernel<KMAX, TT, TO>), blocks, 256, 0, s, logits, ld, (const TT*)target, M, cfg, coeff, grad_scale, (TO*)dlogits, ldd)
#define LAUNCH_BWD1(KMAX, TT)                                     \
  do {                                                            \
    if (out_dtype == GDL_F32) LAUNCH_BWD2(KMAX, TT, float);       \
    else if (out_dtype == GDL_BF16) LAUNCH_BWD2(KMAX, TT, __nv_bfloat16); \
    else LAUNCH_BWD2(KMAX, TT, __half);                           \
  } while (0)
#define LAUNCH_BWD(KMAX)                                          \
  do {                                                            \
    if (target_kind == 0) LAUNCH_BWD1(KMAX, long long);           \
    else LAUNCH_BWD1(KMAX, uint8_t);                              \
  } while (0)
  if (K <= 2) LAUNCH_BWD(2);
  else if (K <= 8) LAUNCH_BWD(8);
  else LAUNCH_BWD(32);
#undef LAUNCH_BWD
#undef LAUNCH_BWD1
#undef LAUNCH_BWD2
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_argmax_classes(const float* logits, int ld, long long M, int K, float threshold, long long* out,
                                  void* stream) {
  GDL_REQUIRE(logits && out && M > 0 && K >= 1 && ld >= K, GDL_ERR_INVALID, "argmax: bad args");
  GDL_LAUNCH(argmax_kernel, loss_blocks(M), 256, 0, (cudaStream_t)stream, logits, ld, M, K, threshold, out);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int step, const float* grad_scale,
                             void* stream) {
  GDL_REQUIRE(p && g && m && v && n > 0 && step >= 1, GDL_ERR_INVALID, "adam: bad args");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  long long b = (n + 255) / 256;
  if (b > 8 * kNumSMsB200) b = 8 * kNumSMsB200;
  GDL_LAUNCH(adam_kernel, (int)b, 256, 0, (cudaStream_t)stream, p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                       (float)bc1, (float)sqrt(bc2), grad_scale);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                                 float beta2, float eps, float weight_decay, float* state /* 3 floats */,
                                 const float* grad_scale, const float* lr_scale, void* stream) {
  GDL_REQUIRE(p && g && m && v && state && n > 0, GDL_ERR_INVALID, "adam_dev: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  GDL_LAUNCH(adam_advance_kernel, 1, 1, 0, st, state, beta1, beta2);
  long long b = (n + 255) / 256;
  if (b > 8 * kNumSMsB200) b = 8 * kNumSMsB200;
  GDL_LAUNCH(adam_dev_kernel, (int)b, 256, 0, st, p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, state, grad_scale, lr_scale);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_grad_clip_coef(const float* g, long long n, float max_norm, float* sumsq_scratch, float* scale,
                                  void* stream) {
  GDL_REQUIRE(g && sumsq_scratch && scale && n > 0, GDL_ERR_INVALID, "grad_clip: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  GDL_CHECK_CUDA(cudaMemsetAsync(sumsq_scratch, 0, sizeof(float), s));
  long long b = (n + 255) / 256;
  if (b > 4 * kNumSMsB200) b = 4 * kNumSMsB200;
  DetCtx det = det_none();
  {
    const DetWs ws = det_workspace();
    if (const int gd = det_grid(ws, b, 1)) {
      b = gd;
      det = det_ctx(ws, gd, 1);
    }
  }
  GDL_LAUNCH(sumsq_kernel, (int)b, 256, 0, s, g, n, sumsq_scratch, det);
  GDL_LAUNCH(clip_coef_kernel, 1, 1, 0, s, sumsq_scratch, max_norm, scale);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}
