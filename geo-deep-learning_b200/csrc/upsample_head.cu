// upsample_head.cu — the segmentation head's last step fused with what consumes it:
//   logits = F.interpolate(low_res_logits, size=(H, W), mode="bilinear", align_corners=False)   (segformer.py:47-57,
//   models/segmentation/dofa.py:90-105), then CrossEntropy / Dice (training_step) or softmax.argmax (validation / test,
//   segmentation_segformer.py:268-271).
// The (N, K, H, W) fp32 logits are never written: every full-resolution pixel is interpolated from its 4 low-resolution
// taps in registers (the low-resolution map is a few MB and stays in L2).
//   gdl_upsample_ce_fwd   statistics pass  -> the same stats / coeff layout as gdl_seg_loss_fwd (loss = coeff[0])
//   gdl_upsample_ce_bwd   d(loss)/d(low-res logits) in GATHER form: one warp per low-resolution pixel walks the
//                         full-resolution pixels whose taps touch it, recomputes their softmax and adds
//                         weight x d(loss)/d(logit); lanes are combined with a fixed shuffle tree (bit-reproducible)
//   gdl_upsample_argmax   class map of the upsampled logits (K == 1: sigmoid > threshold)
// Interpolation arithmetic is bilinear_fwd_kernel<float>'s, operation for operation, so the fused loss equals the
// unfused one bit for bit on the statistics pass.
#include <math.h>
#include <string.h>

#include <type_traits>

#include "../../include/gdl_b200.h"
#include "bilinear.cuh"
#include "loss_cfg.cuh"

namespace gdl {

struct UpGeom {
  int N, h, w, H, W;  // low-resolution map h x w, target resolution H x W
  float sh, sw;       // h / H, w / W
  int ld;             // pixel stride of the low-resolution logits (floats)
};

// z[c] of full-resolution pixel (n, Y, X)
template <int KMAX>
GDL_DEVINL void interp_logits(const float* __restrict__ lr, const UpGeom& g, int K, long long n, int Y, int X, float (&z)[KMAX]) {
  int h0, h1, w0, w1;
  float a0, a1, b0, b1;
  bil_src(Y, g.sh, g.h, h0, h1, a0, a1);
  bil_src(X, g.sw, g.w, w0, w1, b0, b1);
  const float* base = lr + n * g.h * g.w * g.ld;
  const float* p00 = base + ((long long)h0 * g.w + w0) * g.ld;
  const float* p01 = base + ((long long)h0 * g.w + w1) * g.ld;
  const float* p10 = base + ((long long)h1 * g.w + w0) * g.ld;
  const float* p11 = base + ((long long)h1 * g.w + w1) * g.ld;
#pragma unroll
  for (int c = 0; c < KMAX; ++c)
    z[c] = c < K ? bil_mix(a0, a1, b0, b1, __ldg(p00 + c), __ldg(p01 + c), __ldg(p10 + c), __ldg(p11 + c)) : -INFINITY;
}

template <int KMAX, typename TT>
__global__ void __launch_bounds__(256) upsample_ce_stats_kernel(const float* __restrict__ lr, UpGeom g,
                                                                 const TT* __restrict__ target, LossCfg cfg,
                                                                 float* __restrict__ stats, const DetCtx det) {
  GDL_PDL_ENTRY();
  const long long M = (long long)g.N * g.H * g.W;
  LossAcc<KMAX> acc;
  acc.init();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % g.W);
    const long long t = i / g.W;
    const int Y = (int)(t % g.H);
    float z[KMAX];
    interp_logits<KMAX>(lr, g, cfg.K, t / g.H, Y, X, z);
    acc.add(z, load_target(target, i), cfg);
  }
  acc.commit(cfg, stats, det);
}

// one warp per low-resolution pixel
template <int KMAX, typename TT, typename TO>
__global__ void __launch_bounds__(256) upsample_ce_bwd_kernel(const float* __restrict__ lr, UpGeom g,
                                                               const TT* __restrict__ target, LossCfg cfg,
                                                               const float* __restrict__ coeff,
                                                               const float* __restrict__ grad_scale, TO* __restrict__ dlr, int ldd) {
  GDL_PDL_ENTRY();
  const int K = cfg.K;
  const float gs = grad_scale ? grad_scale[0] : 1.f;
  const float inv_denom = 1.f / coeff[1];
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long npix = (long long)g.N * g.h * g.w;
  for (long long pidx = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pidx < npix; pidx += nwarps) {
    const int x = (int)(pidx % g.w);
    const long long t0 = pidx / g.w;
    const int y = (int)(t0 % g.h);
    const long long n = t0 / g.h;
    int Ylo, Yhi, Xlo, Xhi;
    bil_range(y, g.sh, g.H, Ylo, Yhi);
    bil_range(x, g.sw, g.W, Xlo, Xhi);
    const int nX = Xhi - Xlo + 1;
    const int cnt = (Yhi - Ylo + 1) * nX;
    float acc[KMAX];
#pragma unroll
    for (int c = 0; c < KMAX; ++c) acc[c] = 0.f;
    for (int q = lane; q < cnt; q += 32) {
      const int Y = Ylo + q / nX, X = Xlo + q % nX;
      int h0, h1, w0, w1;
      float a0, a1, b0, b1;
      bil_src(Y, g.sh, g.h, h0, h1, a0, a1);
      bil_src(X, g.sw, g.w, w0, w1, b0, b1);
      const float wy = (h0 == y ? a0 : 0.f) + (h1 == y ? a1 : 0.f);
      const float wx = (w0 == x ? b0 : 0.f) + (w1 == x ? b1 : 0.f);
      const float wgt = wy * wx;
      if (wgt == 0.f) continue;
      const long long i = (n * g.H + Y) * g.W + X;
      const long long tt = load_target(target, i);
      const bool ign = cfg.has_ignore && tt == cfg.ignore_index;
      if (ign) continue;
      float z[KMAX], d[KMAX];
      interp_logits<KMAX>(lr, g, K, n, Y, X, z);
      loss_pixel_grad<KMAX>(z, tt, false, cfg, coeff, inv_denom, d);
#pragma unroll
      for (int c = 0; c < KMAX; ++c) acc[c] = fmaf(wgt, d[c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < KMAX; ++c)
      if (c < K) acc[c] = warp_sum(acc[c]);
    if (lane == 0) {
      TO* o = dlr + pidx * ldd;
#pragma unroll
      for (int c = 0; c < KMAX; ++c) {
        if (c < K) {
          const float v = acc[c] * gs;
          if constexpr (std::is_same<TO, float>::value) o[c] = v;
          else if constexpr (std::is_same<TO, __nv_bfloat16>::value) o[c] = __float2bfloat16_rn(v);
          else o[c] = __float2half_rn(v);
        }
      }
    }
  }
}

template <int KMAX>
__global__ void __launch_bounds__(256) upsample_argmax_kernel(const float* __restrict__ lr, UpGeom g, int K, float threshold,
                                                               long long* __restrict__ out) {
  GDL_PDL_ENTRY();
  const long long M = (long long)g.N * g.H * g.W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % g.W);
    const long long t = i / g.W;
    const int Y = (int)(t % g.H);
    float z[KMAX];
    interp_logits<KMAX>(lr, g, K, t / g.H, Y, X, z);
    if (K == 1) {
      out[i] = 1.f / (1.f + expf(-z[0])) > threshold ? 1 : 0;
    } else {
      float best = z[0];
      int bi = 0;
#pragma unroll
      for (int c = 1; c < KMAX; ++c)
        if (c < K && z[c] > best) {  // first maximum wins (torch.argmax)
          best = z[c];
          bi = c;
        }
      out[i] = bi;
    }
  }
}

static int make_geom(UpGeom* g, int N, int h, int w, int H, int W, int ld, int K) {
  GDL_REQUIRE(N > 0 && h > 0 && w > 0 && H > 0 && W > 0 && ld >= K, GDL_ERR_INVALID,
              "upsample head: bad geometry N=%d %dx%d -> %dx%d ld=%d K=%d", N, h, w, H, W, ld, K);
  GDL_REQUIRE((long long)N * H * W < (1ll << 40), GDL_ERR_UNSUPPORTED, "upsample head: too many pixels");
  g->N = N;
  g->h = h;
  g->w = w;
  g->H = H;
  g->W = W;
  g->sh = (float)h / (float)H;
  g->sw = (float)w / (float)W;
  g->ld = ld;
  return 0;
}

}  // namespace gdl

using namespace gdl;

extern "C" int gdl_upsample_ce_fwd(const float* logits_lr, int ld, int N, int h, int w, int H, int W, const void* target,
                                   int target_kind, int K, long long ignore_index, int has_ignore, float w_ce, float w_dice,
                                   float label_smoothing, int ce_mean_over_all, float dice_smooth, float dice_eps,
                                   float* stats, float* coeff, void* stream) {
  GDL_REQUIRE(logits_lr && target && stats && coeff, GDL_ERR_INVALID, "upsample_ce_fwd: null argument");
  LossCfg cfg;
  int st = make_cfg(&cfg, K, ignore_index, has_ignore, w_ce, w_dice, label_smoothing, ce_mean_over_all, dice_smooth, dice_eps);
  if (st) return st;
  UpGeom g;
  st = make_geom(&g, N, h, w, H, W, ld, K);
  if (st) return st;
  cudaStream_t s = (cudaStream_t)stream;
  GDL_CHECK_CUDA(cudaMemsetAsync(stats, 0, (4 + 3 * (size_t)K) * sizeof(float), s));
  int blocks = loss_blocks((long long)N * H * W);
  DetCtx det = det_none();
  {
    const DetWs ws = det_workspace();
    if (const int gd = det_grid(ws, blocks, 4 + 3 * K)) {
      blocks = gd;
      det = det_ctx(ws, gd, 4 + 3 * K);
    }
  }
#define LAUNCH_STATS(KMAX)                                                                                              \
  do {                                                                                                                  \
    if (target_kind == 0)                                                                                               \
      GDL_LAUNCH((upsample_ce_stats_kernel<KMAX, long long>), blocks, 256, 0, s, logits_lr, g, (const long long*)target, cfg, stats, det); \
    else                                                                                                                \
      GDL_LAUNCH((upsample_ce_stats_kernel<KMAX, uint8_t>), blocks, 256, 0, s, logits_lr, g, (const uint8_t*)target, cfg, stats, det);     \
  } while (0)
  if (K <= 2) LAUNCH_STATS(2);
  else if (K <= 8) LAUNCH_STATS(8);
  else LAUNCH_STATS(32);
#undef LAUNCH_STATS
  GDL_CHECK_CUDA(cudaGetLastError());
  launch_loss_finalize(stats, cfg, coeff, s);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_upsample_ce_bwd(const float* logits_lr, int ld, int N, int h, int w, int H, int W, const void* target,
                                   int target_kind, int K, long long ignore_index, int has_ignore, float w_ce, float w_dice,
                                   float label_smoothing, int ce_mean_over_all, float dice_smooth, float dice_eps,
                                   const float* coeff, const float* grad_scale, void* dlogits_lr, int ldd, int out_dtype,
                                   void* stream) {
  GDL_REQUIRE(logits_lr && target && coeff && dlogits_lr && ldd >= K, GDL_ERR_INVALID, "upsample_ce_bwd: bad args");
  LossCfg cfg;
  int st = make_cfg(&cfg, K, ignore_index, has_ignore, w_ce, w_dice, label_smoothing, ce_mean_over_all, dice_smooth, dice_eps);
  if (st) return st;
  UpGeom g;
  st = make_geom(&g, N, h, w, H, W, ld, K);
  if (st) return st;
  cudaStream_t s = (cudaStream_t)stream;
  const long long npix = (long long)N * h * w;
  long long blocks = (npix + 7) / 8;  // 8 warps per block, one low-resolution pixel per warp and iteration
  if (blocks > 32ll * kNumSMsB200) blocks = 32ll * kNumSMsB200;
#define LAUNCH_BWD(KMAX, TT, TO)                                                                                        \
  GDL_LAUNCH((upsample_ce_bwd_kernel<KMAX, TT, TO>), (int)blocks, 256, 0, s, logits_lr, g, (const TT*)target, cfg, coeff, grad_scale, \
                                                                   (TO*)dlogits_lr, ldd)
#define DISPATCH_TO(KMAX, TT)                                            \
  do {                                                                   \
    if (out_dtype == GDL_F32) LAUNCH_BWD(KMAX, TT, float);               \
    else if (out_dtype == GDL_BF16) LAUNCH_BWD(KMAX, TT, __nv_bfloat16); \
    else LAUNCH_BWD(KMAX, TT, __half);                                   \
  } while (0)
#define DISPATCH_K(TT)                    \
  do {                                    \
    if (K <= 2) DISPATCH_TO(2, TT);       \
    else if (K <= 8) DISPATCH_TO(8, TT);  \
    else DISPATCH_TO(32, TT);             \
  } while (0)
  GDL_REQUIRE(out_dtype == GDL_F32 || out_dtype == GDL_BF16 || out_dtype == GDL_F16, GDL_ERR_INVALID, "upsample_ce_bwd: out_dtype");
  if (target_kind == 0) DISPATCH_K(long long);
  else DISPATCH_K(uint8_t);
#undef DISPATCH_K
#undef DISPATCH_TO
#undef LAUNCH_BWD
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_upsample_argmax(const float* logits_lr, int ld, int N, int h, int w, int H, int W, int K, float threshold,
                                   long long* out, void* stream) {
  GDL_REQUIRE(logits_lr && out && K >= 1 && K <= kLossMaxK, GDL_ERR_INVALID, "upsample_argmax: bad args (1 <= K <= %d)", kLossMaxK);
  UpGeom g;
  int st = make_geom(&g, N, h, w, H, W, ld, K);
  if (st) return st;
  const int blocks = loss_blocks((long long)N * H * W);
  cudaStream_t s = (cudaStream_t)stream;
  if (K <= 2) GDL_LAUNCH(upsample_argmax_kernel<2>, blocks, 256, 0, s, logits_lr, g, K, threshold, out);
  else if (K <= 8) GDL_LAUNCH(upsample_argmax_kernel<8>, blocks, 256, 0, s, logits_lr, g, K, threshold, out);
  else GDL_LAUNCH(upsample_argmax_kernel<32>, blocks, 256, 0, s, logits_lr, g, K, threshold, out);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}
