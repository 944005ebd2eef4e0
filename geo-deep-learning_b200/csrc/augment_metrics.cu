// augment_metrics.cu — the two HBM-bound kernels either side of the model:
//   * augment_kernel: the kornia augmentation of `on_before_batch_transfer` (flip / rot90 / resized crop, image
//     bilinear + mask nearest) fused with the patch normalisation: one pass from the raw tile to the 16-bit NHWC
//     operand of the stem (or to an f32 NCHW batch for the Lightning route);
//   * argmax_confusion_kernel: eval post-processing (softmax.argmax / sigmoid > t) fused with the per-sample
//     confusion counts torchmetrics' MeanIoU is computed from.
// Both are one thread per pixel with coalesced stores; reads of a resized crop hit 4 neighbouring source pixels.
#include <stdint.h>

#include "../../include/gdl_b200.h"
#include "common.cuh"

namespace gdl {

static int am_blocks(long long work_items, int threads, int max_waves) {
  long long b = (work_items + threads - 1) / threads;
  const long long cap = (long long)kNumSMsB200 * max_waves;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

template <typename T>
struct Out16;
template <>
struct Out16<__nv_bfloat16> {
  static GDL_DEVINL uint32_t pack2(float a, float b) { return pack_bf16x2(a, b); }
};
template <>
struct Out16<__half> {
  static GDL_DEVINL uint32_t pack2(float a, float b) { return pack_f16x2(a, b); }
};

constexpr int kAugParams = 6;  // {op, k, y0, x0, ch, cw} per sample
enum { kAugIdentity = 0, kAugHFlip = 1, kAugVFlip = 2, kAugRot90 = 3, kAugCrop = 4 };

// source taps of one output pixel: exact ops have one tap (w = 1), a resized crop four
struct AugTaps {
  int y0, y1, x0, x1;      // source rows / columns (already offset by the crop origin)
  float ly, lx;            // weights of row y1 / column x1
  int my, mx;              // nearest-neighbour source of the mask
  bool interp;
};

GDL_DEVINL AugTaps aug_taps(const int* __restrict__ q, int oy, int ox, int H, int W) {
  AugTaps t;
  t.interp = false;
  t.ly = t.lx = 0.f;
  int sy = oy, sx = ox;
  const int op = q[0];
  if (op == kAugHFlip) {
    sx = W - 1 - ox;
  } else if (op == kAugVFlip) {
    sy = H - 1 - oy;
  } else if (op == kAugRot90 && H == W) {
    // torch.rot90(x, k, dims=(H, W)): k = 1 -> out[i][j] = in[j][W-1-i]; k = 2 -> in[H-1-i][W-1-j]; k = 3 -> in[H-1-j][i]
    const int k = q[1] & 3;
    if (k == 1) {
      sy = ox;
      sx = W - 1 - oy;
    } else if (k == 2) {
      sy = H - 1 - oy;
      sx = W - 1 - ox;
    } else if (k == 3) {
      sy = H - 1 - ox;
      sx = oy;
    }
  } else if (op == kAugCrop) {
    // crop rows y0..y0+ch-1, columns x0..x0+cw-1 (clamped into the tile), resized to (H, W):
    // image = F.interpolate(bilinear, align_corners=False), mask = F.interpolate(nearest) — ATen's index arithmetic
    int cy0 = min(max(q[2], 0), H - 1), cx0 = min(max(q[3], 0), W - 1);
    const int ch = min(max(q[4], 1), H - cy0), cw = min(max(q[5], 1), W - cx0);
    const float scy = (float)ch / (float)H, scx = (float)cw / (float)W;
    float fy = scy * ((float)oy + 0.5f) - 0.5f, fx = scx * ((float)ox + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    int iy = min((int)floorf(fy), ch - 1), ix = min((int)floorf(fx), cw - 1);
    t.ly = fminf(fmaxf(fy - (float)iy, 0.f), 1.f);
    t.lx = fminf(fmaxf(fx - (float)ix, 0.f), 1.f);
    t.y0 = cy0 + iy;
    t.y1 = cy0 + min(iy + 1, ch - 1);
    t.x0 = cx0 + ix;
    t.x1 = cx0 + min(ix + 1, cw - 1);
    t.my = cy0 + min((int)floorf((float)oy * scy), ch - 1);
    t.mx = cx0 + min((int)floorf((float)ox * scx), cw - 1);
    t.interp = true;
    return t;
  }
  t.y0 = t.y1 = t.my = sy;
  t.x0 = t.x1 = t.mx = sx;
  return t;
}

// TIn: uint8_t / float; in_chw: NCHW input; OUT_F32: f32 NCHW output (else 16-bit NHWC with pixel stride ld)
template <typename T, typename TIn, typename TM, bool OUT_F32>
__global__ void augment_kernel(const TIn* __restrict__ x, const TM* __restrict__ mask, const int* __restrict__ params,
                               void* __restrict__ yv, TM* __restrict__ mask_out, long long N, int H, int W, int C, int ld,
                               const float* __restrict__ mean, const float* __restrict__ stdv, float image_max,
                               int in_chw) {
  GDL_PDL_ENTRY();
  const long long hw = (long long)H * W, total = N * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / hw;
    const int p = (int)(i - n * hw);
    const int oy = p / W, ox = p - oy * W;
    const AugTaps t = aug_taps(params + n * kAugParams, oy, ox, H, W);
    const long long p00 = (long long)t.y0 * W + t.x0, p01 = (long long)t.y0 * W + t.x1;
    const long long p10 = (long long)t.y1 * W + t.x0, p11 = (long long)t.y1 * W + t.x1;
    if (mask != nullptr) mask_out[i] = mask[n * hw + (long long)t.my * W + t.mx];
    const int cend = OUT_F32 ? C : ld;
    for (int cb = 0; cb < cend; cb += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = cb + j;
        float r = 0.f;
        if (c < C) {
          const long long base = in_chw ? (n * C + c) * hw : n * hw * C + c;
          const long long s = in_chw ? 1 : C;
          r = (float)x[base + p00 * s];
          if (t.interp) {
            // ATen upsample_bilinear2d: w0y * (w0x * a + w1x * b) + w1y * (w0x * c + w1x * d)
            const float a = r, b = (float)x[base + p01 * s], c2 = (float)x[base + p10 * s], d = (float)x[base + p11 * s];
            const float w0x = 1.f - t.lx, w0y = 1.f - t.ly;
            r = w0y * (w0x * a + t.lx * b) + t.ly * (w0x * c2 + t.lx * d);
          }
          // same arithmetic as normalize_kernel (utils/tensors.py:10-35: true fp32 divisions, same order)
          if (image_max > 0.f) r = __fdiv_rn(r, image_max);
          if (mean != nullptr) r = __fdiv_rn(r - mean[c], stdv[c]);
        }
        v[j] = r;
      }
      if (OUT_F32) {
        float* y = reinterpret_cast<float*>(yv);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (cb + j < C) y[(n * C + cb + j) * hw + p] = v[j];
      } else {
        T* y = reinterpret_cast<T*>(yv);
        *reinterpret_cast<uint4*>(y + i * ld + cb) =
            make_uint4(Out16<T>::pack2(v[0], v[1]), Out16<T>::pack2(v[2], v[3]), Out16<T>::pack2(v[4], v[5]),
                       Out16<T>::pack2(v[6], v[7]));
      }
    }
  }
}

template <typename T, typename TIn, bool OUT_F32>
static void launch_augment(const void* x, const void* mask, int mask_kind, const int* params, void* y, void* mask_out,
                           long long N, int H, int W, int C, int ld, const float* mean, const float* stdv,
                           float image_max, int in_chw, cudaStream_t st) {
  const int blocks = am_blocks(N * H * W, 256, 16);
  if (mask_kind == 0)
    GDL_LAUNCH((augment_kernel<T, TIn, long long, OUT_F32>), blocks, 256, 0, st, 
        (const TIn*)x, (const long long*)mask, params, y, (long long*)mask_out, N, H, W, C, ld, mean, stdv, image_max, in_chw);
  else
    GDL_LAUNCH((augment_kernel<T, TIn, uint8_t, OUT_F32>), blocks, 256, 0, st, 
        (const TIn*)x, (const uint8_t*)mask, params, y, (uint8_t*)mask_out, N, H, W, C, ld, mean, stdv, image_max, in_chw);
}

// ------------------------------------------------------------------------------------------
// argmax + per-sample confusion counts.  One block works on pixels of ONE sample at a time (blockIdx.y = sample),
// counts in a shared-memory histogram (Kc*Kc <= 1024 bins), flushed with 64-bit atomics.
// ------------------------------------------------------------------------------------------
constexpr int kMaxConfClasses = 32;

template <typename TT>
__global__ void argmax_confusion_kernel(const float* __restrict__ logits, int ld, long long hw, int K, float threshold,
                                        const TT* __restrict__ target, long long ignore_index, int has_ignore,
                                        long long* __restrict__ classes, unsigned long long* __restrict__ conf) {
  GDL_PDL_ENTRY();
  extern __shared__ unsigned int hist[];
  const int Kc = K == 1 ? 2 : K;
  const int bins = target != nullptr ? Kc * Kc : 0;  // no histogram (and no shared memory) for a classes-only call
  for (int b = threadIdx.x; b < bins; b += blockDim.x) hist[b] = 0u;
  __syncthreads();
  const long long n = blockIdx.y;
  const float* lg = logits + n * hw * ld;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < hw; p += (long long)gridDim.x * blockDim.x) {
    int bi = 0;
    if (K == 1) {
      const float pr = 1.f / (1.f + expf(-lg[p * ld]));
      bi = pr > threshold ? 1 : 0;
    } else {
      float best = lg[p * ld];
      for (int c = 1; c < K; ++c) {
        const float v = lg[p * ld + c];
        if (v > best) {
          best = v;
          bi = c;
        }
      }
    }
    if (classes != nullptr) classes[n * hw + p] = bi;
    if (target != nullptr) {
      const long long tv = (long long)target[n * hw + p];
      if (!(has_ignore && tv == ignore_index) && tv >= 0 && tv < Kc) atomicAdd(&hist[(int)tv * Kc + bi], 1u);
    }
  }
  __syncthreads();
  if (target != nullptr)
    for (int b = threadIdx.x; b < bins; b += blockDim.x)
      if (hist[b]) atomicAdd(&conf[n * bins + b], (unsigned long long)hist[b]);
}

}  // namespace gdl

using namespace gdl;

extern "C" int gdl_augment_normalize(const void* x, int in_kind, const void* mask, int mask_kind, const int* params,
                                     void* y, int out_dtype, void* mask_out, long long N, long long H, long long W,
                                     int C, int ld, const float* mean, const float* stdv, float image_max,
                                     void* stream) {
  GDL_REQUIRE(x && y && params && N > 0 && H > 0 && W > 0 && C > 0, GDL_ERR_INVALID, "augment: bad args");
  GDL_REQUIRE(H * W < (1ll << 31), GDL_ERR_INVALID, "augment: tile too large");
  GDL_REQUIRE((mean == nullptr) == (stdv == nullptr), GDL_ERR_INVALID, "augment: mean and std go together");
  GDL_REQUIRE((mask == nullptr) == (mask_out == nullptr), GDL_ERR_INVALID, "augment: mask and mask_out go together");
  GDL_REQUIRE(mask_kind == 0 || mask_kind == 1, GDL_ERR_INVALID, "augment: mask_kind must be 0 (int64) or 1 (uint8)");
  GDL_REQUIRE(in_kind >= 0 && in_kind <= 3, GDL_ERR_INVALID, "augment: unknown in_kind %d", in_kind);
  GDL_REQUIRE(x != y && (mask == nullptr || mask != mask_out), GDL_ERR_INVALID, "augment: in-place operation is not supported");
  cudaStream_t st = (cudaStream_t)stream;
  const int in_chw = in_kind >= 2;
  const bool in_u8 = in_kind == 0 || in_kind == 3;
  if (out_dtype == GDL_F32) {  // f32 NCHW out (the batch["image"] layout of the Lightning route)
    if (in_u8)
      launch_augment<__nv_bfloat16, uint8_t, true>(x, mask, mask_kind, params, y, mask_out, N, (int)H, (int)W, C, ld, mean, stdv, image_max, in_chw, st);
    else
      launch_augment<__nv_bfloat16, float, true>(x, mask, mask_kind, params, y, mask_out, N, (int)H, (int)W, C, ld, mean, stdv, image_max, in_chw, st);
  } else {
    GDL_REQUIRE(out_dtype == GDL_BF16 || out_dtype == GDL_F16, GDL_ERR_INVALID, "augment: bad out_dtype %d", out_dtype);
    GDL_REQUIRE(ld >= C && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, GDL_ERR_INVALID,
                "augment: 16-bit NHWC output needs ld >= C, ld %% 8 == 0 and a 16-byte aligned base (ld %d)", ld);
    if (out_dtype == GDL_BF16) {
      if (in_u8)
        launch_augment<__nv_bfloat16, uint8_t, false>(x, mask, mask_kind, params, y, mask_out, N, (int)H, (int)W, C, ld, mean, stdv, image_max, in_chw, st);
      else
        launch_augment<__nv_bfloat16, float, false>(x, mask, mask_kind, params, y, mask_out, N, (int)H, (int)W, C, ld, mean, stdv, image_max, in_chw, st);
    } else {
      if (in_u8)
        launch_augment<__half, uint8_t, false>(x, mask, mask_kind, params, y, mask_out, N, (int)H, (int)W, C, ld, mean, stdv, image_max, in_chw, st);
      else
        launch_augment<__half, float, false>(x, mask, mask_kind, params, y, mask_out, N, (int)H, (int)W, C, ld, mean, stdv, image_max, in_chw, st);
    }
  }
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_argmax_confusion(const float* logits, int ld, long long N, long long HW, int K, float threshold,
                                    const void* target, int target_kind, long long ignore_index, int has_ignore,
                                    long long* classes, long long* conf, void* stream) {
  GDL_REQUIRE(logits && N > 0 && HW > 0 && K >= 1 && ld >= K, GDL_ERR_INVALID, "argmax_confusion: bad args");
  GDL_REQUIRE(N <= 65535, GDL_ERR_INVALID, "argmax_confusion: at most 65535 samples per call");
  GDL_REQUIRE((target == nullptr) == (conf == nullptr), GDL_ERR_INVALID, "argmax_confusion: target and conf go together");
  GDL_REQUIRE(classes || conf, GDL_ERR_INVALID, "argmax_confusion: nothing to write");
  GDL_REQUIRE(target_kind == 0 || target_kind == 1, GDL_ERR_INVALID, "argmax_confusion: target_kind must be 0 (int64) or 1 (uint8)");
  const int Kc = K == 1 ? 2 : K;
  GDL_REQUIRE(conf == nullptr || Kc <= kMaxConfClasses, GDL_ERR_UNSUPPORTED,
              "argmax_confusion: at most %d classes (got %d)", kMaxConfClasses, Kc);
  // blocks per sample: enough to fill the machine a few times, few enough that the histogram flush stays cheap
  long long per = (HW + 256 * 8 - 1) / (256 * 8);
  const long long cap = ((long long)kNumSMsB200 * 8 + N - 1) / N;
  if (per > cap) per = cap;
  if (per < 1) per = 1;
  dim3 grid((unsigned)per, (unsigned)N);
  const size_t smem = conf != nullptr ? (size_t)Kc * Kc * sizeof(unsigned int) : 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (target_kind == 0)
    GDL_LAUNCH(argmax_confusion_kernel<long long>, grid, 256, smem, st, logits, ld, HW, K, threshold, (const long long*)target,
                                                               ignore_index, has_ignore, classes, (unsigned long long*)conf);
  else
    GDL_LAUNCH(argmax_confusion_kernel<uint8_t>, grid, 256, smem, st, logits, ld, HW, K, threshold, (const uint8_t*)target,
                                                             ignore_index, has_ignore, classes, (unsigned long long*)conf);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}
