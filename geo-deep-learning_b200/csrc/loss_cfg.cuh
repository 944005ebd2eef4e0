// loss_cfg.cuh — configuration / statistics layout of the segmentation losses, shared by loss_optim.cu (logits in HBM)
// and upsample_head.cu (logits interpolated on the fly from the low-resolution map).
#pragma once
#include "common.cuh"

namespace gdl {

constexpr int kLossMaxK = 32;

// stats layout (floats): [0] nll_sum [1] smooth_sum [2] valid_count [3] total_count
//                        [4 + c] inter_c   [4 + K + c] card_c   [4 + 2K + c] tsum_c
// coeff layout (floats): [0] loss  [1] ce_denominator  [2 + c] dice_a_c (dL/dp_c for t=0)
//                        [2 + K + c] dice_b_c (dL/dp_c for t=1)
struct LossCfg {
  int K;
  int binary;         // K == 1: sigmoid instead of softmax
  long long ignore_index;
  int has_ignore;
  float w_ce, w_dice;
  float label_smoothing;
  int ce_mean_over_all;  // smp SoftCrossEntropyLoss: mean over all pixels
  float dice_smooth, dice_eps;
};

template <typename TT>
GDL_DEVINL long long load_target(const TT* t, long long i) {
  return (long long)t[i];
}

// host helpers defined in loss_optim.cu
int make_cfg(LossCfg* c, int K, long long ignore_index, int has_ignore, float w_ce, float w_dice, float label_smoothing,
             int ce_mean_over_all, float dice_smooth, float dice_eps);
int loss_blocks(long long M);
void launch_loss_finalize(const float* stats, const LossCfg& cfg, float* coeff, cudaStream_t s);

}  // namespace gdl

#include "det_reduce.cuh"

namespace gdl {

// per-thread accumulators of the loss statistics + their ordered block / grid reduction (blockDim.x == 256)
template <int KMAX>
struct LossAcc {
  float nll, smooth, valid, total;
  float inter[KMAX], card[KMAX], tsum[KMAX];
  GDL_DEVINL void init() {
    nll = smooth = valid = total = 0.f;
#pragma unroll
    for (int c = 0; c < KMAX; ++c) inter[c] = card[c] = tsum[c] = 0.f;
  }
  // one pixel: z[c] = logits (c >= K: -inf), t = target class
  GDL_DEVINL void add(const float (&z)[KMAX], long long t, const LossCfg& cfg) {
    const int K = cfg.K;
    const bool ign = cfg.has_ignore && t == cfg.ignore_index;
    total += 1.f;
    if (cfg.binary) {
      // p = sigmoid(z) computed as exp(logsigmoid(z)) like smp
      const float ls = fminf(z[0], 0.f) - log1pf(expf(-fabsf(z[0])));
      const float p = expf(ls);
      if (!ign) {
        const float tt = t != 0 ? 1.f : 0.f;
        inter[0] += p * tt;
        card[0] += p + tt;
        tsum[0] += tt;
        // BCE-with-logits as the "ce" term
        nll += -(tt * ls + (1.f - tt) * (ls - z[0]));
        valid += 1.f;
      }
    } else {
      float mx = z[0];
#pragma unroll
      for (int c = 1; c < KMAX; ++c) mx = fmaxf(mx, z[c]);
      float se = 0.f;
#pragma unroll
      for (int c = 0; c < KMAX; ++c) se += c < K ? expf(z[c] - mx) : 0.f;
      const float lse = mx + logf(se);
      if (!ign) {
        float sl = 0.f;
#pragma unroll
        for (int c = 0; c < KMAX; ++c) {
          if (c < K) {
            const float lp = z[c] - lse;
            const float p = expf(lp);
            const float tt = (t == c) ? 1.f : 0.f;
            inter[c] += p * tt;
            card[c] += p + tt;
            tsum[c] += tt;
            sl += -lp;
            if (t == c) nll += -lp;
          }
        }
        smooth += sl;
        valid += 1.f;
      }
    }
  }
  // every thread of every block, uniform control flow
  __device__ inline void commit(const LossCfg& cfg, float* __restrict__ stats, const DetCtx& det) {
    const int K = cfg.K;
    // per-warp partials, added in warp order (no shared-memory atomics: their arrival order is not reproducible)
    __shared__ float shw[8][4 + 3 * kLossMaxK];
  // block reduction: warp shuffles (fixed tree), then the 8 warps' partials in warp order
  const int wid = threadIdx.x >> 5;
  nll = warp_sum(nll);
  smooth = warp_sum(smooth);
  valid = warp_sum(valid);
  total = warp_sum(total);
  if ((threadIdx.x & 31) == 0) {
    shw[wid][0] = nll;
    shw[wid][1] = smooth;
    shw[wid][2] = valid;
    shw[wid][3] = total;
  }
#pragma unroll
  for (int c = 0; c < KMAX; ++c) {
    if (c < K) {
      const float a = warp_sum(inter[c]), b = warp_sum(card[c]), d = warp_sum(tsum[c]);
      if ((threadIdx.x & 31) == 0) {
        shw[wid][4 + c] = a;
        shw[wid][4 + K + c] = b;
        shw[wid][4 + 2 * K + c] = d;
      }
    }
  }
  __syncthreads();
  const int nvals = 4 + 3 * K;
  for (int i = threadIdx.x; i < nvals; i += blockDim.x) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += shw[w][i];
    if (det.s0 != nullptr) det_put(det, nvals, i, v);
    else atomicAdd(&stats[i], v);
  }
  if (det.s0 != nullptr) det_finish(det, nvals, stats);
  }
};

// d[c] = d(loss)/d(z_c) of one pixel (before the caller's gradient scale)
template <int KMAX>
GDL_DEVINL void loss_pixel_grad(const float (&z)[KMAX], long long t, bool ign, const LossCfg& cfg,
                                const float* __restrict__ coeff, float inv_denom, float (&d)[KMAX]) {
  const int K = cfg.K;
#pragma unroll
    for (int c = 0; c < KMAX; ++c) d[c] = 0.f;
    if (!ign) {
      if (cfg.binary) {
        const float p = 1.f / (1.f + expf(-z[0]));
        const float tt = t != 0 ? 1.f : 0.f;
        float g = 0.f;
        if (cfg.w_ce != 0.f) g += cfg.w_ce * (p - tt) * inv_denom;
        if (cfg.w_dice != 0.f) g += cfg.w_dice * (tt > 0.f ? coeff[2 + K] : coeff[2]) * p * (1.f - p);
        d[0] = g;
      } else {
        float mx = z[0];
#pragma unroll
        for (int c = 1; c < KMAX; ++c) mx = fmaxf(mx, z[c]);
        float p[KMAX];
        float se = 0.f;
#pragma unroll
        for (int c = 0; c < KMAX; ++c) {
          p[c] = c < K ? expf(z[c] - mx) : 0.f;
          se += p[c];
        }
        const float inv = 1.f / se;
        float dot = 0.f;
        float dldp[KMAX];
#pragma unroll
        for (int c = 0; c < KMAX; ++c) {
          p[c] *= inv;
          dldp[c] = 0.f;
          if (c < K && cfg.w_dice != 0.f) {
            dldp[c] = cfg.w_dice * ((t == c) ? coeff[2 + K + c] : coeff[2 + c]);
            dot += dldp[c] * p[c];
          }
        }
        const float eps = cfg.label_smoothing;
#pragma unroll
        for (int c = 0; c < KMAX; ++c) {
          if (c < K) {
            float g = 0.f;
            if (cfg.w_ce != 0.f) {
              // d/dz [ (1-eps) * nll + eps/K * sum_c(-log p_c) ] = p - (1-eps) onehot - eps/K
              const float tt = (t == c) ? 1.f : 0.f;
              g += cfg.w_ce * (p[c] - (1.f - eps) * tt - eps / (float)K) * inv_denom;
            }
            if (cfg.w_dice != 0.f) g += p[c] * (dldp[c] - dot);
            d[c] = g;
          }
        }
      }
    }
}

}  // namespace gdl
