// bilinear.cuh — index arithmetic of F.interpolate(mode="bilinear", align_corners=False), shared by the resize kernels
// (transformer.cu) and the fused upsample + loss / argmax head (upsample_head.cu).
#pragma once
#include "common.cuh"

namespace gdl {

// Source index rule of ATen: src = scale*(dst+0.5)-0.5, clamped at 0 (segformer.py:51-57, segformer_mlp.py:88-119).
GDL_DEVINL void bil_src(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 < in_size - 1 ? i0 + 1 : i0;
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

// the 4-tap mix with its rounding points pinned (explicit mul / fma): every kernel that interpolates — the resize
// kernels and the fused upsample + loss / argmax head — produces bit-identical values
GDL_DEVINL float bil_mix(float a0, float a1, float b0, float b1, float x00, float x01, float x10, float x11) {
  const float top = __fmaf_rn(b1, x01, __fmul_rn(b0, x00));
  const float bot = __fmaf_rn(b1, x11, __fmul_rn(b0, x10));
  return __fmaf_rn(a1, bot, __fmul_rn(a0, top));
}

// outputs d whose taps can touch input i (a superset; callers test the taps):
GDL_DEVINL void bil_range(int i, float scale, int out_size, int& lo, int& hi) {
  // outputs d with floor(src(d)) in {i-1, i}: src(d) in [i-1, i+1)  ->  d in [(i-0.5)/scale-0.5, (i+1.5)/scale-0.5)
  float a = ((float)i - 0.5f) / scale - 0.5f, b = ((float)i + 1.5f) / scale - 0.5f;
  lo = (int)floorf(a) - 1;
  hi = (int)ceilf(b) + 1;
  if (lo < 0) lo = 0;
  if (hi > out_size - 1) hi = out_size - 1;
}

}  // namespace gdl
