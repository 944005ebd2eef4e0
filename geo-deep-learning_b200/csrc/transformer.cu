// transformer.cu — the HBM-bound kernels of the MixTransformer / SegFormer path:
// LayerNorm (fwd/bwd, fp32 residual stream in, 16-bit normalised tokens out), row softmax for the
// spatial-reduction attention scores (fwd/bwd), depthwise 3x3 conv + bias + exact-erf GELU of the
// Mix-FFN (fwd/bwd, NHWC so no NLC<->NCHW transposes exist), bilinear resize with
// align_corners=False (fwd/bwd, 16-bit features and fp32 logits).
//
// Tokens (B, N, C) and maps (B, h, w, C) are the same memory (NHWC), which is why the reference's
// reshape/permute/contiguous copies (mix_transformer.py:499,541-546; segformer_mlp.py:78-121) vanish.
#include <math.h>
#include <string.h>

#include <type_traits>

#include "../../include/gdl_b200.h"
#include "common.cuh"
#include "det_reduce.cuh"
#include "bilinear.cuh"

namespace gdl {

template <typename T>
GDL_DEVINL float to_f(T v);
template <>
GDL_DEVINL float to_f<float>(float v) { return v; }
template <>
GDL_DEVINL float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
GDL_DEVINL float to_f<__half>(__half v) { return __half2float(v); }
template <typename T>
GDL_DEVINL T from_f(float v);
template <>
GDL_DEVINL float from_f<float>(float v) { return v; }
template <>
GDL_DEVINL __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
GDL_DEVINL __half from_f<__half>(float v) { return __float2half_rn(v); }

static int row_blocks(long long rows, int rows_per_block, int max_waves = 8) {
  long long b = (rows + rows_per_block - 1) / rows_per_block;
  long long cap = (long long)kNumSMsB200 * max_waves;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

// ------------------------------------------------------------------------------------------
// LayerNorm over the last dim.  One warp per row; C <= 32 * kLnMaxPerLane.
// x: TI (fp32 residual stream or 16-bit), y: TO (16-bit operand of the next GEMM, or fp32).
// ------------------------------------------------------------------------------------------
constexpr int kLnMaxPerLane = 32;  // C <= 1024

template <typename TI, typename TO>
__global__ void layernorm_fwd_kernel(const TI* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float eps, TO* __restrict__ y, long long ldy,
                                     float* __restrict__ mean_out, float* __restrict__ rstd_out, long long M, int C) {
  GDL_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int per = (C + 31) / 32;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += (long long)gridDim.x * wpb) {
    float v[kLnMaxPerLane];
    float s = 0.f;
#pragma unroll 4
    for (int i = 0; i < per; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < C ? to_f<TI>(x[row * ldx + c]) : 0.f;
      s += v[i];
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll 4
    for (int i = 0; i < per; ++i) {
      const int c = lane + 32 * i;
      const float d = c < C ? v[i] - mean : 0.f;
      q = fmaf(d, d, q);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll 4
    for (int i = 0; i < per; ++i) {
      const int c = lane + 32 * i;
      if (c < C) y[row * ldy + c] = from_f<TO>((v[i] - mean) * rstd * gamma[c] + beta[c]);
    }
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
  }
}

// backward: dx = rstd * (g*gamma - mean_c(g*gamma) - xhat * mean_c(g*gamma*xhat)) [+ add]
//   written as fp32 (residual-stream gradient) and/or as a 16-bit copy (operand of the previous GEMM's
//   dgrad/wgrad).  dgamma/dbeta partial sums per block -> atomics (pre-zeroed [2][C]: dgamma then dbeta).
template <typename TI, typename TG>
__global__ void layernorm_bwd_kernel(const TG* __restrict__ g, long long ldg, const TI* __restrict__ x, long long ldx,
                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                     const float* __restrict__ gamma, const float* __restrict__ add, long long lda,
                                     float* __restrict__ dx32, long long ld32, void* __restrict__ dx16, long long ld16,
                                     int dx16_is_half, float* __restrict__ pgrads, long long M, int C, const DetCtx det) {
  GDL_PDL_ENTRY();
  extern __shared__ float sh[];  // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int per = (C + 31) / 32;
  float dga[kLnMaxPerLane], dbe[kLnMaxPerLane];
#pragma unroll 4
  for (int i = 0; i < per; ++i) dga[i] = dbe[i] = 0.f;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += (long long)gridDim.x * wpb) {
    const float mu = mean[row], rs = rstd[row];
    float gg[kLnMaxPerLane], xh[kLnMaxPerLane];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
    for (int i = 0; i < per; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        const float gv = to_f<TG>(g[row * ldg + c]);
        xh[i] = (to_f<TI>(x[row * ldx + c]) - mu) * rs;
        dga[i] = fmaf(gv, xh[i], dga[i]);
        dbe[i] += gv;
        gg[i] = gv * gamma[c];
        s1 += gg[i];
        s2 = fmaf(gg[i], xh[i], s2);
      } else {
        gg[i] = xh[i] = 0.f;
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
#pragma unroll 4
    for (int i = 0; i < per; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        float d = rs * (gg[i] - s1 - xh[i] * s2);
        if (add) d += add[row * lda + c];
        if (dx32) dx32[row * ld32 + c] = d;
        if (dx16) {
          if (dx16_is_half)
            reinterpret_cast<__half*>(dx16)[row * ld16 + c] = __float2half_rn(d);
          else
            reinterpret_cast<__nv_bfloat16*>(dx16)[row * ld16 + c] = __float2bfloat16_rn(d);
        }
      }
    }
  }
  if (pgrads != nullptr) {
    // the block's warps add their partials in warp order (shared-memory atomics would land in arrival order)
    for (int w = 0; w < wpb; ++w) {
      if ((int)(threadIdx.x >> 5) == w) {
#pragma unroll 4
        for (int i = 0; i < per; ++i) {
          const int c = lane + 32 * i;
          if (c < C) {
            sh[c] += dga[i];
            sh[C + c] += dbe[i];
          }
        }
      }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
      if (det.s0 != nullptr) det_put(det, 2 * C, i, sh[i]);
      else atomicAdd(&pgrads[i], sh[i]);
    }
    if (det.s0 != nullptr) det_finish(det, 2 * C, pgrads);
  }
}

// ------------------------------------------------------------------------------------------
// softmax over rows of length L (attention scores, mix_transformer.py:151-152):
//   p = softmax(scale * s); rows are stored with stride ld >= L, pad columns of p are written 0.
// backward: ds = scale * p * (dp - sum_j dp_j p_j)
// ------------------------------------------------------------------------------------------
constexpr int kSmMaxPerLane = 48;  // L <= 1536 (DOFA: 1297 tokens)

// PER = compile-time bound of the per-lane element count: the row lives in registers (a runtime trip count put the
// array in local memory)
template <typename T, int PER>
__global__ void softmax_fwd_kernel(const T* __restrict__ s, long long lds, float scale, T* __restrict__ p,
                                   long long ldp, long long M, int L, int Lpad) {
  GDL_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += (long long)gridDim.x * wpb) {
    float v[PER];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < L ? to_f<T>(s[row * lds + c]) * scale : -INFINITY;
      mx = fmaxf(mx, v[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < L ? expf(v[i] - mx) : 0.f;
      sum += v[i];
    }
    const float inv = 1.f / warp_sum(sum);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = lane + 32 * i;
      if (c < Lpad) p[row * ldp + c] = from_f<T>(v[i] * inv);
    }
  }
}

template <typename T, int PER>
__global__ void softmax_bwd_kernel(const T* __restrict__ p, long long ldp, const T* __restrict__ dp, long long lddp,
                                   float scale, T* __restrict__ ds, long long ldds, long long M, int L, int Lpad) {
  GDL_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += (long long)gridDim.x * wpb) {
    float pv[PER], dv[PER];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = lane + 32 * i;
      pv[i] = c < L ? to_f<T>(p[row * ldp + c]) : 0.f;
      dv[i] = c < L ? to_f<T>(dp[row * lddp + c]) : 0.f;
      dot = fmaf(pv[i], dv[i], dot);
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = lane + 32 * i;
      if (c < Lpad) ds[row * ldds + c] = from_f<T>(c < L ? scale * pv[i] * (dv[i] - dot) : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------
// depthwise 3x3 (pad 1, stride 1, bias) + exact GELU, NHWC 16-bit, C % 8 == 0  (Mlp.dwconv + act,
// mix_transformer.py:56-63,533-546).  Weights fp32 [C][3][3] (the nn.Conv2d(groups=C) layout squeezed).
// fwd stores the pre-activation (16-bit, as the autocast reference does) for the backward.
// ------------------------------------------------------------------------------------------
GDL_DEVINL float gelu_f(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752f)); }
GDL_DEVINL float gelu_grad_f(float v) {
  return 0.5f * (1.f + erff(v * 0.70710678118654752f)) + v * 0.39894228040143268f * expf(-0.5f * v * v);
}

template <typename T>
GDL_DEVINL void ld8(const T* p, float (&f)[8]);
template <>
GDL_DEVINL void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = bf16_lo(w[i]);
    f[2 * i + 1] = bf16_hi(w[i]);
  }
}
template <>
GDL_DEVINL void ld8<__half>(const __half* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
template <typename T>
GDL_DEVINL void st8(T* p, const float (&f)[8]);
template <>
GDL_DEVINL void st8<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                            pack_bf16x2(f[6], f[7]));
}
template <>
GDL_DEVINL void st8<__half>(__half* p, const float (&f)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_f16x2(f[0], f[1]), pack_f16x2(f[2], f[3]), pack_f16x2(f[4], f[5]),
                                            pack_f16x2(f[6], f[7]));
}

template <typename T>
GDL_DEVINL void unpack8(const uint4& u, float (&f)[8]);
template <>
GDL_DEVINL void unpack8<__nv_bfloat16>(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = bf16_lo(w[i]);
    f[2 * i + 1] = bf16_hi(w[i]);
  }
}
template <>
GDL_DEVINL void unpack8<__half>(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// ---- 128-bit versions of the softmax kernels (16-bit scores, rows padded to a multiple of 8): lane l owns the
// 8-element chunks l, l+32, ... of a row; PV = chunks per lane (compile time, the row lives in registers).  The scalar
// version moved 2 bytes per lane and load: 2.5 TB/s on the 1344-key DOFA rows (run 14).
template <typename T, int PV>
__global__ void softmax_fwd_vec_kernel(const T* __restrict__ s, long long lds, float scale, T* __restrict__ p,
                                       long long ldp, long long M, int L, int Lpad) {
  GDL_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += (long long)gridDim.x * wpb) {
    float v[PV][8];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < PV; ++i) {
      const int c = (lane + 32 * i) * 8;
      if (c < L) {
        unpack8<T>(*reinterpret_cast<const uint4*>(s + row * lds + c), v[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[i][j] = (c + j < L) ? v[i][j] * scale : -INFINITY;
          mx = fmaxf(mx, v[i][j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = -INFINITY;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < PV; ++i) {
      const int c = (lane + 32 * i) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[i][j] = (c + j < L) ? expf(v[i][j] - mx) : 0.f;
        sum += v[i][j];
      }
    }
    const float inv = 1.f / warp_sum(sum);
#pragma unroll
    for (int i = 0; i < PV; ++i) {
      const int c = (lane + 32 * i) * 8;
      if (c < Lpad) {
        float o8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o8[j] = v[i][j] * inv;
        st8(p + row * ldp + c, o8);
      }
    }
  }
}

template <typename T, int PV>
__global__ void softmax_bwd_vec_kernel(const T* __restrict__ p, long long ldp, const T* __restrict__ dp, long long lddp,
                                       float scale, T* __restrict__ ds, long long ldds, long long M, int L, int Lpad) {
  GDL_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += (long long)gridDim.x * wpb) {
    float pv[PV][8], dv[PV][8];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < PV; ++i) {
      const int c = (lane + 32 * i) * 8;
      if (c < L) {
        unpack8<T>(*reinterpret_cast<const uint4*>(p + row * ldp + c), pv[i]);
        unpack8<T>(*reinterpret_cast<const uint4*>(dp + row * lddp + c), dv[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (c + j >= L) pv[i][j] = dv[i][j] = 0.f;
          dot = fmaf(pv[i][j], dv[i][j], dot);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) pv[i][j] = dv[i][j] = 0.f;
      }
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int i = 0; i < PV; ++i) {
      const int c = (lane + 32 * i) * 8;
      if (c < Lpad) {
        float o8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o8[j] = (c + j < L) ? scale * pv[i][j] * (dv[i][j] - dot) : 0.f;
        st8(ds + row * ldds + c, o8);
      }
    }
  }
}

constexpr int kDwThreads = 128;  // ~160 registers per thread: 3 blocks of 128 per SM instead of 1 of 256
constexpr int kDwStrip = 32;  // pixels of one image row a thread walks with a rolling 3x3 window

// Depthwise 3x3 (pad 1) over NHWC, one thread = one 8-channel vector x one strip of kDwStrip consecutive pixels of an
// image row.  The 3x3 neighbourhood is a rolling window of packed 16-byte vectors: 3 new loads per pixel instead of 9,
// pointer increments instead of per-tap index arithmetic (the per-pixel version ran at 0.7 TB/s).
//   GELU = true : y = GELU(round16(conv(x) + bias)), pre = round16(conv(x) + bias)      (forward, dwconv+GELU)
//   FLIP = true : taps mirrored (dx = conv-transpose of dpre with the same filter), no bias, no GELU   (backward dx)
template <typename T, bool GELU, bool FLIP>
__global__ void __launch_bounds__(kDwThreads) dwconv_strip_kernel(const T* __restrict__ x, int ldx, const float* __restrict__ w,
                                                           const float* __restrict__ bias, T* __restrict__ pre,
                                                           T* __restrict__ y, int ldy, int N, int H, int W, int C) {
  GDL_PDL_ENTRY();
  const int cv = C / 8;
  const int tpr = cv < (int)blockDim.x ? cv : blockDim.x;
  const int rows_per_block = blockDim.x / tpr;
  const int tc = threadIdx.x % tpr;
  const int tr = threadIdx.x / tpr;
  if (tr >= rows_per_block) return;
  const int strips_w = (W + kDwStrip - 1) / kDwStrip;
  const long long nstrips = (long long)N * H * strips_w;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (int c8 = tc; c8 < cv; c8 += tpr) {
    const int c0 = c8 * 8;
    float wr[8][9], bz[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      bz[j] = (!FLIP && bias) ? bias[c0 + j] : 0.f;
#pragma unroll
      for (int k = 0; k < 9; ++k) wr[j][k] = w[(c0 + j) * 9 + (FLIP ? 8 - k : k)];
    }
    for (long long sidx = (long long)blockIdx.x * rows_per_block + tr; sidx < nstrips;
         sidx += (long long)gridDim.x * rows_per_block) {
      const int ws = (int)(sidx % strips_w);
      const long long t = sidx / strips_w;
      const int h = (int)(t % H);
      const long long n = t / H;
      const int w_begin = ws * kDwStrip;
      const int w_end = min(W, w_begin + kDwStrip);
      const T* rowp[3];
      bool rv[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int hh = h + r - 1;
        rv[r] = hh >= 0 && hh < H;
        rowp[r] = x + ((n * H + (rv[r] ? hh : h)) * W) * ldx + c0;
      }
      uint4 colA[3], colB[3], colC[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        colA[r] = (rv[r] && w_begin > 0) ? *reinterpret_cast<const uint4*>(rowp[r] + (long long)(w_begin - 1) * ldx) : zero;
        colB[r] = rv[r] ? *reinterpret_cast<const uint4*>(rowp[r] + (long long)w_begin * ldx) : zero;
      }
      const long long pix0 = (n * H + h) * W;
      for (int wq = w_begin; wq < w_end; ++wq) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
          colC[r] = (rv[r] && wq + 1 < W) ? *reinterpret_cast<const uint4*>(rowp[r] + (long long)(wq + 1) * ldx) : zero;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = bz[j];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          float f[8];
          unpack8<T>(colA[r], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(f[j], wr[j][r * 3 + 0], acc[j]);
          unpack8<T>(colB[r], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(f[j], wr[j][r * 3 + 1], acc[j]);
          unpack8<T>(colC[r], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(f[j], wr[j][r * 3 + 2], acc[j]);
        }
        const long long pix = pix0 + wq;
        if (GELU) {
          // the autocast reference rounds the conv output to 16 bits before GELU: do the same
          const float (&a)[8] = acc;
          st8(pre + pix * C + c0, a);
          float out[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) out[j] = gelu_f(to_f<T>(from_f<T>(acc[j])));
          st8(y + pix * ldy + c0, out);
        } else {
          st8(y + pix * ldy + c0, acc);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          colA[r] = colB[r];
          colB[r] = colC[r];
        }
      }
    }
  }
}

// dpre = dy * gelu'(pre)   (in place on dy allowed)
template <typename T>
__global__ void gelu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ pre, T* __restrict__ dpre,
                                long long n8) {
  GDL_PDL_ENTRY();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    float a[8], b[8], o[8];
    ld8(dy + i * 8, a);
    ld8(pre + i * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = a[j] * gelu_grad_f(b[j]);
    st8(dpre + i * 8, o);
  }
}

// dw[c][tap] += sum_p dpre[p] * x[p + tap], db[c] += sum_p dpre[p]; pgrads fp32 [C][10], accumulated.
// Same strip walk with a rolling window on x; the 80 partial sums of a thread are reduced across the block's strips in
// shared memory before the atomics.
template <typename T>
__global__ void __launch_bounds__(kDwThreads) dwconv_bwd_dw_kernel(const T* __restrict__ dpre, const T* __restrict__ x, int ldx,
                                                            float* __restrict__ pgrads, int N, int H, int W, int C,
                                                            const DetCtx det) {
  GDL_PDL_ENTRY();
  const int cv = C / 8;
  const int tpr = cv < (int)blockDim.x ? cv : blockDim.x;
  const int rows_per_block = blockDim.x / tpr;
  const int tc = threadIdx.x % tpr;
  const int tr = threadIdx.x / tpr;
  const int strips_w = (W + kDwStrip - 1) / kDwStrip;
  const long long nstrips = (long long)N * H * strips_w;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  __shared__ float red[kDwThreads][10];
  for (int cbase = 0; cbase < cv; cbase += tpr) {  // uniform trip count (barriers inside)
    const int c8 = cbase + tc;
    const bool active = c8 < cv && tr < rows_per_block;
    const int c0 = c8 * 8;
    float wg[8][10];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int k = 0; k < 10; ++k) wg[j][k] = 0.f;
    if (active) {
      for (long long sidx = (long long)blockIdx.x * rows_per_block + tr; sidx < nstrips;
           sidx += (long long)gridDim.x * rows_per_block) {
        const int ws = (int)(sidx % strips_w);
        const long long t = sidx / strips_w;
        const int h = (int)(t % H);
        const long long n = t / H;
        const int w_begin = ws * kDwStrip;
        const int w_end = min(W, w_begin + kDwStrip);
        const T* rowp[3];
        bool rv[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int hh = h + r - 1;
          rv[r] = hh >= 0 && hh < H;
          rowp[r] = x + ((n * H + (rv[r] ? hh : h)) * W) * ldx + c0;
        }
        uint4 colA[3], colB[3], colC[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          colA[r] = (rv[r] && w_begin > 0) ? *reinterpret_cast<const uint4*>(rowp[r] + (long long)(w_begin - 1) * ldx) : zero;
          colB[r] = rv[r] ? *reinterpret_cast<const uint4*>(rowp[r] + (long long)w_begin * ldx) : zero;
        }
        const T* gp = dpre + ((n * H + h) * W) * C + c0;
        for (int wq = w_begin; wq < w_end; ++wq) {
          const uint4 gu = *reinterpret_cast<const uint4*>(gp + (long long)wq * C);
#pragma unroll
          for (int r = 0; r < 3; ++r)
            colC[r] = (rv[r] && wq + 1 < W) ? *reinterpret_cast<const uint4*>(rowp[r] + (long long)(wq + 1) * ldx) : zero;
          float g0[8];
          unpack8<T>(gu, g0);
#pragma unroll
          for (int j = 0; j < 8; ++j) wg[j][9] += g0[j];
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            float f[8];
            unpack8<T>(colA[r], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) wg[j][r * 3 + 0] = fmaf(g0[j], f[j], wg[j][r * 3 + 0]);
            unpack8<T>(colB[r], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) wg[j][r * 3 + 1] = fmaf(g0[j], f[j], wg[j][r * 3 + 1]);
            unpack8<T>(colC[r], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) wg[j][r * 3 + 2] = fmaf(g0[j], f[j], wg[j][r * 3 + 2]);
          }
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            colA[r] = colB[r];
            colB[r] = colC[r];
          }
        }
      }
    }
    // block reduction over the strips that share a channel vector, one channel (10 sums) at a time
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 10; ++k) red[threadIdx.x][k] = active ? wg[j][k] : 0.f;
      __syncthreads();
      if (tr == 0 && c8 < cv) {
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          float a = 0.f;
          for (int r2 = 0; r2 < rows_per_block; ++r2) a += red[r2 * tpr + tc][k];
          if (det.s0 != nullptr) det_put(det, 10 * C, (c0 + j) * 10 + k, a);
          else atomicAdd(&pgrads[(long long)(c0 + j) * 10 + k], a);
        }
      }
    }
  }
  if (det.s0 != nullptr) det_finish(det, 10 * C, pgrads);
}

// ------------------------------------------------------------------------------------------
// bilinear resize, align_corners=False (F.interpolate(mode="bilinear"), segformer.py:51-57,
// segformer_mlp.py:88-119).  Source index rule of ATen: src = scale*(dst+0.5)-0.5, clamped at 0.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void bilinear_fwd_kernel(const T* __restrict__ x, long long ldx, T* __restrict__ y, long long ldy, int N,
                                    int Hi, int Wi, int Ho, int Wo, int C, float sh, float sw) {
  GDL_PDL_ENTRY();
  const long long total = (long long)N * Ho * Wo * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int wo = (int)(t % Wo);
    t /= Wo;
    const int ho = (int)(t % Ho);
    const long long n = t / Ho;
    int h0, h1, w0, w1;
    float a0, a1, b0, b1;
    bil_src(ho, sh, Hi, h0, h1, a0, a1);
    bil_src(wo, sw, Wi, w0, w1, b0, b1);
    const T* base = x + n * Hi * Wi * ldx + c;
    const float v = bil_mix(a0, a1, b0, b1, to_f<T>(base[((long long)h0 * Wi + w0) * ldx]), to_f<T>(base[((long long)h0 * Wi + w1) * ldx]),
                            to_f<T>(base[((long long)h1 * Wi + w0) * ldx]), to_f<T>(base[((long long)h1 * Wi + w1) * ldx]));
    y[((n * Ho + ho) * Wo + wo) * ldy + c] = from_f<T>(v);
  }
}

// 8 channels per thread (16-bit features, C % 8 == 0)
template <typename T>
__global__ void bilinear_fwd_vec8_kernel(const T* __restrict__ x, long long ldx, T* __restrict__ y, long long ldy, int N,
                                         int Hi, int Wi, int Ho, int Wo, int C, float sh, float sw) {
  GDL_PDL_ENTRY();
  const int cv = C / 8;
  const long long total = (long long)N * Ho * Wo * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cv) * 8;
    long long t = i / cv;
    const int wo = (int)(t % Wo);
    t /= Wo;
    const int ho = (int)(t % Ho);
    const long long n = t / Ho;
    int h0, h1, w0, w1;
    float a0, a1, b0, b1;
    bil_src(ho, sh, Hi, h0, h1, a0, a1);
    bil_src(wo, sw, Wi, w0, w1, b0, b1);
    const T* base = x + n * Hi * Wi * ldx + c0;
    float f00[8], f01[8], f10[8], f11[8], o[8];
    ld8(base + ((long long)h0 * Wi + w0) * ldx, f00);
    ld8(base + ((long long)h0 * Wi + w1) * ldx, f01);
    ld8(base + ((long long)h1 * Wi + w0) * ldx, f10);
    ld8(base + ((long long)h1 * Wi + w1) * ldx, f11);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = bil_mix(a0, a1, b0, b1, f00[j], f01[j], f10[j], f11[j]);
    st8(y + ((n * Ho + ho) * Wo + wo) * ldy + c0, o);
  }
}

// out = base + sum_i bilinear(src_i -> (Ho, Wo)): the SegFormer decoder's fuse layer after the 1x1 conv has been moved in
// front of the (linear) resizes (segformer_mlp.py:77-128; gdl_bilinear_sum_fwd).  8 channels per thread, fp32 sum in a
// fixed order (base, src 0, 1, 2), ONE rounding to 16 bits; every tap uses bil_mix, so one term equals bilinear_fwd's value
// before its rounding.
struct BilinearSumSrc {
  const void* ptr;
  long long ld;
  int H, W;
  float sh, sw;
};
struct BilinearSumParams {
  BilinearSumSrc src[3];
  int num_src;
};
template <typename T>
__global__ void bilinear_sum_vec8_kernel(const T* __restrict__ base, long long ldb, const BilinearSumParams p,
                                         T* __restrict__ y, long long ldy, int N, int Ho, int Wo, int C) {
  GDL_PDL_ENTRY();
  const int cv = C / 8;
  const long long total = (long long)N * Ho * Wo * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cv) * 8;
    long long t = i / cv;
    const int wo = (int)(t % Wo);
    t /= Wo;
    const int ho = (int)(t % Ho);
    const long long n = t / Ho;
    const long long pix = (n * Ho + ho) * Wo + wo;
    float o[8];
    if (base != nullptr) {
      ld8(base + pix * ldb + c0, o);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (k < p.num_src) {
        const BilinearSumSrc& s = p.src[k];
        int h0, h1, w0, w1;
        float a0, a1, b0, b1;
        bil_src(ho, s.sh, s.H, h0, h1, a0, a1);
        bil_src(wo, s.sw, s.W, w0, w1, b0, b1);
        const T* x = reinterpret_cast<const T*>(s.ptr) + n * s.H * s.W * s.ld + c0;
        float f00[8], f01[8], f10[8], f11[8];
        ld8(x + ((long long)h0 * s.W + w0) * s.ld, f00);
        ld8(x + ((long long)h0 * s.W + w1) * s.ld, f01);
        ld8(x + ((long long)h1 * s.W + w0) * s.ld, f10);
        ld8(x + ((long long)h1 * s.W + w1) * s.ld, f11);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = __fadd_rn(o[j], bil_mix(a0, a1, b0, b1, f00[j], f01[j], f10[j], f11[j]));
      }
    }
    st8(y + pix * ldy + c0, o);
  }
}

// gather-form adjoint: dx[n,hi,wi,c] = sum over output pixels whose taps touch (hi,wi)
template <typename T>
__global__ void bilinear_bwd_kernel(const T* __restrict__ dy, long long ldy, T* __restrict__ dx, long long ldx, int N,
                                    int Hi, int Wi, int Ho, int Wo, int C, float sh, float sw) {
  GDL_PDL_ENTRY();
  const long long total = (long long)N * Hi * Wi * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int wi = (int)(t % Wi);
    t /= Wi;
    const int hi = (int)(t % Hi);
    const long long n = t / Hi;
    int hlo, hhi, wlo, whi;
    bil_range(hi, sh, Ho, hlo, hhi);
    bil_range(wi, sw, Wo, wlo, whi);
    float acc = 0.f;
    for (int ho = hlo; ho <= hhi; ++ho) {
      int h0, h1;
      float a0, a1;
      bil_src(ho, sh, Hi, h0, h1, a0, a1);
      float wh = (h0 == hi ? a0 : 0.f) + (h1 == hi ? a1 : 0.f);
      if (wh == 0.f) continue;
      for (int wo = wlo; wo <= whi; ++wo) {
        int w0, w1;
        float b0, b1;
        bil_src(wo, sw, Wi, w0, w1, b0, b1);
        const float ww = (w0 == wi ? b0 : 0.f) + (w1 == wi ? b1 : 0.f);
        if (ww == 0.f) continue;
        acc = fmaf(wh * ww, to_f<T>(dy[((n * Ho + ho) * Wo + wo) * ldy + c]), acc);
      }
    }
    dx[((n * Hi + hi) * Wi + wi) * ldx + c] = from_f<T>(acc);
  }
}

template <typename T>
__global__ void bilinear_bwd_vec8_kernel(const T* __restrict__ dy, long long ldy, T* __restrict__ dx, long long ldx,
                                         int N, int Hi, int Wi, int Ho, int Wo, int C, float sh, float sw) {
  GDL_PDL_ENTRY();
  const int cv = C / 8;
  const long long total = (long long)N * Hi * Wi * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cv) * 8;
    long long t = i / cv;
    const int wi = (int)(t % Wi);
    t /= Wi;
    const int hi = (int)(t % Hi);
    const long long n = t / Hi;
    int hlo, hhi, wlo, whi;
    bil_range(hi, sh, Ho, hlo, hhi);
    bil_range(wi, sw, Wo, wlo, whi);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int ho = hlo; ho <= hhi; ++ho) {
      int h0, h1;
      float a0, a1;
      bil_src(ho, sh, Hi, h0, h1, a0, a1);
      const float wh = (h0 == hi ? a0 : 0.f) + (h1 == hi ? a1 : 0.f);
      if (wh == 0.f) continue;
      for (int wo = wlo; wo <= whi; ++wo) {
        int w0, w1;
        float b0, b1;
        bil_src(wo, sw, Wi, w0, w1, b0, b1);
        const float ww = (w0 == wi ? b0 : 0.f) + (w1 == wi ? b1 : 0.f);
        if (ww == 0.f) continue;
        float f[8];
        ld8(dy + ((n * Ho + ho) * Wo + wo) * ldy + c0, f);
        const float k = wh * ww;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(k, f[j], acc[j]);
      }
    }
    st8(dx + ((n * Hi + hi) * Wi + wi) * ldx + c0, acc);
  }
}

// ------------------------------------------------------------------------------------------
// nn.AdaptiveAvgPool2d(s) on NHWC (PPM, models/utils.py:55-93): bin [floor(i*H/s), ceil((i+1)*H/s))
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void adaptive_avgpool_fwd_kernel(const T* __restrict__ x, long long ldx, T* __restrict__ y, int N, int H,
                                            int W, int C, int S) {
  GDL_PDL_ENTRY();
  const long long total = (long long)N * S * S * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int ow = (int)(t % S);
    t /= S;
    const int oh = (int)(t % S);
    const long long n = t / S;
    const int h0 = (oh * H) / S, h1 = ((oh + 1) * H + S - 1) / S;
    const int w0 = (ow * W) / S, w1 = ((ow + 1) * W + S - 1) / S;
    float acc = 0.f;
    for (int h = h0; h < h1; ++h)
      for (int w = w0; w < w1; ++w) acc += to_f<T>(x[((n * H + h) * W + w) * ldx + c]);
    y[i] = from_f<T>(acc / (float)((h1 - h0) * (w1 - w0)));
  }
}

template <typename T>
__global__ void adaptive_avgpool_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int C,
                                            int S) {
  GDL_PDL_ENTRY();
  const long long total = (long long)N * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const long long n = t / H;
    float acc = 0.f;
    for (int oh = 0; oh < S; ++oh) {
      const int h0 = (oh * H) / S, h1 = ((oh + 1) * H + S - 1) / S;
      if (h < h0 || h >= h1) continue;
      for (int ow = 0; ow < S; ++ow) {
        const int w0 = (ow * W) / S, w1 = ((ow + 1) * W + S - 1) / S;
        if (w < w0 || w >= w1) continue;
        acc += to_f<T>(dy[((n * S + oh) * S + ow) * C + c]) / (float)((h1 - h0) * (w1 - w0));
      }
    }
    dx[i] = from_f<T>(acc);
  }
}

// y = a + b (16-bit NHWC with strides) — UperNet top-down path  laterals[i-1] + up(laterals[i])
template <typename T>
__global__ void add_kernel(const T* __restrict__ a, long long lda, const T* __restrict__ b, long long ldb,
                           T* __restrict__ y, long long ldy, long long M, int C) {
  GDL_PDL_ENTRY();
  const long long total = M * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    y[r * ldy + c] = from_f<T>(to_f<T>(a[r * lda + c]) + to_f<T>(b[r * ldb + c]));
  }
}

// ------------------------------------------------------------------------------------------
// ViT token glue (dofa_v2.py:445-468): tokens[b][0] = cls, tokens[b][1+p] = patch[b][p] + pos[1+p];
// feature tap: feat[b][p] = cast(tokens[b][1+p])  (drop the cls token, NHWC map = token rows)
// ------------------------------------------------------------------------------------------
template <typename TP>
__global__ void vit_assemble_tokens_kernel(const TP* __restrict__ patch, const float* __restrict__ pos,
                                           const float* __restrict__ cls, float* __restrict__ tokens, int B, int P, int C) {
  GDL_PDL_ENTRY();
  const long long total = (long long)B * (P + 1) * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long t = i / C;
    const int tok = (int)(t % (P + 1));
    const long long b = t / (P + 1);
    tokens[i] = tok == 0 ? cls[c] : to_f<TP>(patch[(b * P + tok - 1) * C + c]) + pos[(long long)tok * C + c];
  }
}

template <typename TO>
__global__ void vit_extract_feature_kernel(const float* __restrict__ tokens, TO* __restrict__ feat, int B, int P, int C) {
  GDL_PDL_ENTRY();
  const long long total = (long long)B * P * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long t = i / C;
    const int p = (int)(t % P);
    const long long b = t / P;
    feat[i] = from_f<TO>(tokens[(b * (P + 1) + p + 1) * C + c]);
  }
}

// elementwise helpers for the fp32 residual stream
template <typename T>
__global__ void cast_f32_kernel(const float* __restrict__ x, T* __restrict__ y, long long n) {
  GDL_PDL_ENTRY();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = from_f<T>(x[i]);
}

// ------------------------------------------------------------------------------------------
// ViT block glue for TRAINING the DOFA encoder (timm Block: x = x + drop_path(ls(branch(norm(x))))), where the
// GEMM-epilogue fusions of the forward-only path (GELU, LayerScale + residual) cannot be used because the backward
// needs the pre-activation / the un-scaled branch output.
// ------------------------------------------------------------------------------------------
// y = gelu(x), exact (erf) as nn.GELU
template <typename T>
__global__ void gelu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long n8) {
  GDL_PDL_ENTRY();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    float a[8], o[8];
    ld8(x + i * 8, a);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = gelu_f(a[j]);
    st8(y + i * 8, o);
  }
}

// out[r][c] = res[r][c] + s(r) * gamma[c] * u[r][c]  on the fp32 residual stream; s(r) = sscale[r / rows_per_sample]
// (DropPath: per-sample keep mask / keep probability) or 1.  One thread per 8 channels.
template <typename T>
__global__ void layerscale_add_kernel(const float* __restrict__ res, const T* __restrict__ u,
                                      const float* __restrict__ gamma, const float* __restrict__ sscale,
                                      long long rows_per_sample, float* __restrict__ out, long long M, int C) {
  GDL_PDL_ENTRY();
  const int cv = C / 8;
  const long long total = M * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cv;
    const int c = (int)(i - r * cv) * 8;
    const float s = sscale != nullptr ? sscale[r / rows_per_sample] : 1.f;
    float uu[8];
    ld8(u + r * C + c, uu);
    const float4 r0 = *reinterpret_cast<const float4*>(res + r * C + c), r1 = *reinterpret_cast<const float4*>(res + r * C + c + 4);
    // gamma is a parameter: inside a flat parameter buffer it is only 4-byte aligned, so no vector loads on it
    const float* gm = gamma + c;
    float4 o0, o1;
    o0.x = r0.x + s * gm[0] * uu[0];
    o0.y = r0.y + s * gm[1] * uu[1];
    o0.z = r0.z + s * gm[2] * uu[2];
    o0.w = r0.w + s * gm[3] * uu[3];
    o1.x = r1.x + s * gm[4] * uu[4];
    o1.y = r1.y + s * gm[5] * uu[5];
    o1.z = r1.z + s * gm[6] * uu[6];
    o1.w = r1.w + s * gm[7] * uu[7];
    *reinterpret_cast<float4*>(out + r * C + c) = o0;
    *reinterpret_cast<float4*>(out + r * C + c + 4) = o1;
  }
}

// backward of the above w.r.t. u and gamma:  du[r][c] = s(r) * gamma[c] * g[r][c] (16-bit operand of the branch's last
// GEMM backward);  dgamma[c] += sum_r s(r) * g[r][c] * u[r][c]  (fp32 atomics, one per thread and channel at the end).
// blockDim.x = 256 = rpb rows x tpr channel vectors (tpr = C/8 <= 256).
template <typename T>
__global__ void layerscale_bwd_kernel(const float* __restrict__ g, const T* __restrict__ u, const float* __restrict__ gamma,
                                      const float* __restrict__ sscale, long long rows_per_sample, T* __restrict__ du,
                                      float* __restrict__ dgamma, long long M, int C, const DetCtx det) {
  GDL_PDL_ENTRY();
  const int tpr = C / 8;
  const int rpb = blockDim.x / tpr;
  const int rl = threadIdx.x / tpr, v = threadIdx.x - rl * tpr;
  const bool act = rl < rpb;
  const int c = v * 8;
  float gm[8], acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    gm[j] = act ? gamma[c + j] : 0.f;
    acc[j] = 0.f;
  }
  if (act) {
    for (long long r = (long long)blockIdx.x * rpb + rl; r < M; r += (long long)gridDim.x * rpb) {
      const float s = sscale != nullptr ? sscale[r / rows_per_sample] : 1.f;
      float uu[8], o[8];
      ld8(u + r * C + c, uu);
      const float4 g0 = *reinterpret_cast<const float4*>(g + r * C + c), g1 = *reinterpret_cast<const float4*>(g + r * C + c + 4);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = s * gm[j] * gg[j];
        acc[j] += s * gg[j] * uu[j];
      }
      st8(du + r * C + c, o);
    }
  }
  if (dgamma != nullptr) {
    // rows of the block in row order per channel (no atomics inside the block), then block partials -> dgamma
    __shared__ float red[256][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = acc[j];
    __syncthreads();
    if (rl == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = 0.f;
        for (int r2 = 0; r2 < rpb; ++r2) a += red[r2 * tpr + v][j];
        if (det.s0 != nullptr) det_put(det, C, c + j, a);
        else atomicAdd(dgamma + c + j, a);
      }
    }
    if (det.s0 != nullptr) det_finish(det, C, dgamma);
  }
}

// gradient of a feature tap (feat[b][p] = tokens[b][1+p]) into the fp32 stream gradient (B, P+1, C):
// init != 0: g[b][0] = 0, g[b][1+p] = dfeat[b][p]   (the deepest tap starts the stream gradient)
// init == 0: g[b][1+p] += dfeat[b][p]
template <typename T>
__global__ void vit_feature_grad_kernel(const T* __restrict__ dfeat, float* __restrict__ g, int B, int P, int C, int init) {
  GDL_PDL_ENTRY();
  const long long total = (long long)B * (P + 1) * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long t = i / C;
    const int tok = (int)(t % (P + 1));
    const long long b = t / (P + 1);
    if (tok == 0) {
      if (init) g[i] = 0.f;
      continue;
    }
    const float d = to_f<T>(dfeat[(b * P + tok - 1) * C + c]);
    g[i] = init ? d : g[i] + d;
  }
}

// ------------------------------------------------------------------------------------------
// DynamicChannelEmbed (mix_transformer.py:762-865) — the per-band pieces that are not GEMMs.
//   xw     [P][C*E]  16-bit: band-weighted spatial-conv outputs, band c in columns c*E .. c*E+E-1
//   scores [P][16]   fp32:   channel-attention logits of the C <= 16 bands (columns >= C unused)
// channel_pool_fwd:  a = softmax_c(scores);  out[p][e] = sum_c a[c] * xw[p][c*E + e];  attn[p][c] = a[c] (kept for bwd)
// channel_pool_bwd:  dxw[p][c*E+e] = a[c] * dout[p][e];  dattn[c] = sum_e dout[p][e] * xw[p][c*E+e];
//                    dscores[c] = a[c] * (dattn[c] - sum_k a[k] dattn[k])   (16-bit, columns >= C zero)
// One thread per (pixel, 8 channels of E); the E/8 threads of a pixel are adjacent lanes of one warp (E/8 in {1,2,4,8,16,32}).
// ------------------------------------------------------------------------------------------
constexpr int kMaxBands = 16;

template <typename T>
__global__ void channel_pool_fwd_kernel(const T* __restrict__ xw, const float* __restrict__ scores, T* __restrict__ out,
                                        float* __restrict__ attn, long long P, int C, int E) {
  GDL_PDL_ENTRY();
  const int tpp = E / 8;
  const long long total = P * tpp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / tpp;
    const int v = (int)(i - p * tpp);
    float a[kMaxBands];
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) {
      a[c] = scores[p * 16 + c];
      mx = fmaxf(mx, a[c]);
    }
    float sum = 0.f;
    for (int c = 0; c < C; ++c) {
      a[c] = expf(a[c] - mx);
      sum += a[c];
    }
    const float inv = 1.f / sum;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < C; ++c) {
      a[c] *= inv;
      float x[8];
      ld8(xw + p * (long long)C * E + (long long)c * E + v * 8, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(a[c], x[j], acc[j]);
    }
    st8(out + p * E + v * 8, acc);
    if (v == 0 && attn != nullptr) {
      for (int c = 0; c < 16; ++c) attn[p * 16 + c] = c < C ? a[c] : 0.f;
    }
  }
}

template <typename T>
__global__ void channel_pool_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ xw, const float* __restrict__ attn,
                                        T* __restrict__ dxw, T* __restrict__ dscores, long long P, int C, int E) {
  GDL_PDL_ENTRY();
  const int tpp = E / 8;
  // whole pixels per block iteration so that the lanes of a pixel stay together and every lane of a warp takes the same
  // number of trips through the loop (the shuffles below need all 32 lanes)
  const long long total = P * tpp;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long trips = (total + stride - 1) / stride;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (long long t = 0; t < trips; ++t, i += stride) {
    const bool live = i < total;
    const long long p = live ? i / tpp : 0;
    const int v = live ? (int)(i - p * tpp) : 0;
    float a[kMaxBands], da[kMaxBands];
    float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (live) ld8(dout + p * E + v * 8, d);
    for (int c = 0; c < C; ++c) {
      a[c] = live ? attn[p * 16 + c] : 0.f;
      float part = 0.f;
      if (live) {
        float x[8], o[8];
        ld8(xw + p * (long long)C * E + (long long)c * E + v * 8, x);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          part = fmaf(d[j], x[j], part);
          o[j] = a[c] * d[j];
        }
        st8(dxw + p * (long long)C * E + (long long)c * E + v * 8, o);
      }
      for (int off = tpp >> 1; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
      da[c] = part;
    }
    if (live && v == 0) {
      float dot = 0.f;
      for (int c = 0; c < C; ++c) dot = fmaf(a[c], da[c], dot);
      float o0[8], o1[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        o0[c] = c < C ? a[c] * (da[c] - dot) : 0.f;
        o1[c] = c + 8 < C ? a[c + 8] * (da[c + 8] - dot) : 0.f;
      }
      st8(dscores + p * 16, o0);
      st8(dscores + p * 16 + 8, o1);
    }
  }
}

// dx = dy * (y > 0)  (ReLU backward from the activation; 16-bit, n % 8 == 0)
template <typename T>
__global__ void relu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, long long n8) {
  GDL_PDL_ENTRY();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    float a[8], b[8], o[8];
    ld8(dy + i * 8, a);
    ld8(y + i * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = b[j] > 0.f ? a[j] : 0.f;
    st8(dx + i * 8, o);
  }
}

}  // namespace gdl

using namespace gdl;

#define GDL_DISPATCH_T(dtype, ...)                                     \
  do {                                                                 \
    if ((dtype) == GDL_BF16) {                                         \
      using T = __nv_bfloat16;                                         \
      __VA_ARGS__;                                                     \
    } else if ((dtype) == GDL_F16) {                                   \
      using T = __half;                                                \
      __VA_ARGS__;                                                     \
    } else if ((dtype) == GDL_F32) {                                   \
      using T = float;                                                 \
      __VA_ARGS__;                                                     \
    } else {                                                           \
      ::gdl::set_last_error("unknown dtype %d", dtype);                \
      return GDL_ERR_INVALID;                                          \
    }                                                                  \
  } while (0)

extern "C" int gdl_layernorm_fwd(const void* x, int x_dtype, long long ldx, const float* gamma, const float* beta,
                                 float eps, void* y, int y_dtype, long long ldy, float* mean, float* rstd,
                                 long long M, int C, void* stream) {
  GDL_REQUIRE(x && y && gamma && beta && M > 0 && C > 0 && C <= 32 * kLnMaxPerLane, GDL_ERR_INVALID,
              "layernorm: bad args (C=%d must be <= %d)", C, 32 * kLnMaxPerLane);
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = row_blocks(M, 8);
#define LN_FWD(TI, TO) \
  GDL_LAUNCH((layernorm_fwd_kernel<TI, TO>), blocks, 256, 0, st, (const TI*)x, ldx, gamma, beta, eps, (TO*)y, ldy, mean, rstd, M, C)
  if (x_dtype == GDL_F32) {
    if (y_dtype == GDL_F32) LN_FWD(float, float);
    else if (y_dtype == GDL_BF16) LN_FWD(float, __nv_bfloat16);
    else LN_FWD(float, __half);
  } else if (x_dtype == GDL_BF16) {
    if (y_dtype == GDL_F32) LN_FWD(__nv_bfloat16, float);
    else if (y_dtype == GDL_BF16) LN_FWD(__nv_bfloat16, __nv_bfloat16);
    else LN_FWD(__nv_bfloat16, __half);
  } else {
    if (y_dtype == GDL_F32) LN_FWD(__half, float);
    else if (y_dtype == GDL_BF16) LN_FWD(__half, __nv_bfloat16);
    else LN_FWD(__half, __half);
  }
#undef LN_FWD
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_layernorm_bwd(const void* g, int g_dtype, long long ldg, const void* x, int x_dtype, long long ldx,
                                 const float* mean, const float* rstd, const float* gamma, const float* add,
                                 long long lda, float* dx32, long long ld32, void* dx16, int dx16_dtype, long long ld16,
                                 float* pgrads /* [2][C] dgamma, dbeta: accumulated */, long long M, int C,
                                 void* stream) {
  GDL_REQUIRE(g && x && mean && rstd && gamma && (dx32 || dx16) && M > 0 && C > 0 && C <= 32 * kLnMaxPerLane,
              GDL_ERR_INVALID, "layernorm_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = row_blocks(M, 8, 4);
  const int smem = 2 * C * (int)sizeof(float);
  const int is_half = dx16_dtype == GDL_F16;
  DetCtx det = det_none();
  if (pgrads) {
    const DetWs ws = det_workspace();
    if (const int gd = det_grid(ws, blocks, 2 * C)) {
      blocks = gd;
      det = det_ctx(ws, gd, 2 * C);
    }
  }
#define LN_BWD(TI, TG)                                                                                          \
  GDL_LAUNCH((layernorm_bwd_kernel<TI, TG>), blocks, 256, smem, st, (const TG*)g, ldg, (const TI*)x, ldx, mean, rstd, gamma, add, \
                                                         lda, dx32, ld32, dx16, ld16, is_half, pgrads, M, C, det)
  if (x_dtype == GDL_F32) {
    if (g_dtype == GDL_F32) LN_BWD(float, float);
    else if (g_dtype == GDL_BF16) LN_BWD(float, __nv_bfloat16);
    else LN_BWD(float, __half);
  } else if (x_dtype == GDL_BF16) {
    if (g_dtype == GDL_F32) LN_BWD(__nv_bfloat16, float);
    else if (g_dtype == GDL_BF16) LN_BWD(__nv_bfloat16, __nv_bfloat16);
    else LN_BWD(__nv_bfloat16, __half);
  } else {
    if (g_dtype == GDL_F32) LN_BWD(__half, float);
    else if (g_dtype == GDL_BF16) LN_BWD(__half, __nv_bfloat16);
    else LN_BWD(__half, __half);
  }
#undef LN_BWD
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_softmax_fwd(const void* s, long long lds, float scale, void* p, long long ldp, int dtype,
                               long long M, int L, int Lpad, void* stream) {
  GDL_REQUIRE(s && p && M > 0 && L > 0 && Lpad >= L && Lpad <= 32 * kSmMaxPerLane && ldp >= Lpad && lds >= L,
              GDL_ERR_INVALID, "softmax: bad args (L=%d Lpad=%d)", L, Lpad);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype != GDL_F32 && Lpad % 8 == 0 && lds % 8 == 0 && ldp % 8 == 0 && lds >= Lpad &&
      ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(p)) & 15) == 0) {
    const int pv = (Lpad + 255) / 256;
#define GDL_SMV_FWD(PVV)                                                                                            \
  do {                                                                                                              \
    if (dtype == GDL_BF16)                                                                                          \
      GDL_LAUNCH((softmax_fwd_vec_kernel<__nv_bfloat16, PVV>), row_blocks(M, 8), 256, 0, st,                                  \
          (const __nv_bfloat16*)s, lds, scale, (__nv_bfloat16*)p, ldp, M, L, Lpad);                                 \
    else                                                                                                            \
      GDL_LAUNCH((softmax_fwd_vec_kernel<__half, PVV>), row_blocks(M, 8), 256, 0, st, (const __half*)s, lds, scale, (__half*)p, \
                                                                            ldp, M, L, Lpad);                       \
  } while (0)
    if (pv <= 1) GDL_SMV_FWD(1);
    else if (pv <= 2) GDL_SMV_FWD(2);
    else if (pv <= 4) GDL_SMV_FWD(4);
    else GDL_SMV_FWD(6);
#undef GDL_SMV_FWD
    GDL_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const int per = (Lpad + 31) / 32;
#define GDL_SM_FWD(PERV) \
  GDL_DISPATCH_T(dtype, { GDL_LAUNCH((softmax_fwd_kernel<T, PERV>), row_blocks(M, 8), 256, 0, st, (const T*)s, lds, scale, (T*)p, ldp, M, L, Lpad); })
  if (per <= 8) GDL_SM_FWD(8);
  else if (per <= 16) GDL_SM_FWD(16);
  else if (per <= 32) GDL_SM_FWD(32);
  else GDL_SM_FWD(kSmMaxPerLane);
#undef GDL_SM_FWD
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_softmax_bwd(const void* p, long long ldp, const void* dp, long long lddp, float scale, void* ds,
                               long long ldds, int dtype, long long M, int L, int Lpad, void* stream) {
  GDL_REQUIRE(p && dp && ds && M > 0 && L > 0 && Lpad >= L && Lpad <= 32 * kSmMaxPerLane, GDL_ERR_INVALID,
              "softmax_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype != GDL_F32 && Lpad % 8 == 0 && ldp % 8 == 0 && lddp % 8 == 0 && ldds % 8 == 0 && ldp >= Lpad && lddp >= Lpad &&
      ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(dp) | reinterpret_cast<uintptr_t>(ds)) & 15) == 0) {
    const int pv = (Lpad + 255) / 256;
#define GDL_SMV_BWD(PVV)                                                                                             \
  do {                                                                                                               \
    if (dtype == GDL_BF16)                                                                                           \
      GDL_LAUNCH((softmax_bwd_vec_kernel<__nv_bfloat16, PVV>), row_blocks(M, 8), 256, 0, st,                                   \
          (const __nv_bfloat16*)p, ldp, (const __nv_bfloat16*)dp, lddp, scale, (__nv_bfloat16*)ds, ldds, M, L, Lpad); \
    else                                                                                                             \
      GDL_LAUNCH((softmax_bwd_vec_kernel<__half, PVV>), row_blocks(M, 8), 256, 0, st, (const __half*)p, ldp, (const __half*)dp, \
                                                                            lddp, scale, (__half*)ds, ldds, M, L, Lpad); \
  } while (0)
    if (pv <= 1) GDL_SMV_BWD(1);
    else if (pv <= 2) GDL_SMV_BWD(2);
    else if (pv <= 4) GDL_SMV_BWD(4);
    else GDL_SMV_BWD(6);
#undef GDL_SMV_BWD
    GDL_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  GDL_DISPATCH_T(dtype, {
    const int per = (Lpad + 31) / 32;
    if (per <= 8)
      GDL_LAUNCH((softmax_bwd_kernel<T, 8>), row_blocks(M, 8), 256, 0, st, (const T*)p, ldp, (const T*)dp, lddp, scale, (T*)ds, ldds, M, L, Lpad);
    else if (per <= 16)
      GDL_LAUNCH((softmax_bwd_kernel<T, 16>), row_blocks(M, 8), 256, 0, st, (const T*)p, ldp, (const T*)dp, lddp, scale, (T*)ds, ldds, M, L, Lpad);
    else if (per <= 32)
      GDL_LAUNCH((softmax_bwd_kernel<T, 32>), row_blocks(M, 8), 256, 0, st, (const T*)p, ldp, (const T*)dp, lddp, scale, (T*)ds, ldds, M, L, Lpad);
    else
      GDL_LAUNCH((softmax_bwd_kernel<T, kSmMaxPerLane>), row_blocks(M, 8), 256, 0, st, (const T*)p, ldp, (const T*)dp, lddp, scale, (T*)ds, ldds, M, L, Lpad);
  });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int chan_row_grid(long long rows, int C, int rows_per_thread, int threads = 256) {
  const int cv = C / 8;
  const int tpr = cv < threads ? cv : threads;
  const int rpb = threads / tpr;
  long long b = (rows + (long long)rpb * rows_per_thread - 1) / ((long long)rpb * rows_per_thread);
  const long long cap = (long long)kNumSMsB200 * 8;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

extern "C" int gdl_dwconv3x3_gelu_fwd(const void* x, int ldx, const float* w, const float* bias, void* pre, void* y,
                                      int dtype, int N, int H, int W, int C, void* stream) {
  GDL_REQUIRE(x && w && pre && y && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && ldx % 8 == 0, GDL_ERR_INVALID,
              "dwconv_gelu: bad args");
  GDL_REQUIRE(dtype == GDL_BF16 || dtype == GDL_F16, GDL_ERR_INVALID, "dwconv_gelu: 16-bit dtype expected");
  cudaStream_t st = (cudaStream_t)stream;
  const long long nstrips = (long long)N * H * ((W + kDwStrip - 1) / kDwStrip);
  const int grid = chan_row_grid(nstrips, C, 1, kDwThreads);
  if (dtype == GDL_BF16)
    GDL_LAUNCH((dwconv_strip_kernel<__nv_bfloat16, true, false>), grid, kDwThreads, 0, st, (const __nv_bfloat16*)x, ldx, w, bias, (__nv_bfloat16*)pre, (__nv_bfloat16*)y, C, N, H, W, C);
  else
    GDL_LAUNCH((dwconv_strip_kernel<__half, true, false>), grid, kDwThreads, 0, st, (const __half*)x, ldx, w, bias, (__half*)pre, (__half*)y, C, N, H, W, C);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_dwconv3x3_gelu_bwd(const void* dy, const void* pre, const void* x, int ldx, const float* w,
                                      void* dpre_scratch, void* dx, int lddx, float* pgrads /* [C][10] accumulated */,
                                      int dtype, int N, int H, int W, int C, void* stream) {
  GDL_REQUIRE(dy && pre && x && w && dpre_scratch && dx && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0,
              GDL_ERR_INVALID, "dwconv_gelu_bwd: bad args");
  GDL_REQUIRE(dtype == GDL_BF16 || dtype == GDL_F16, GDL_ERR_INVALID, "dwconv_gelu_bwd: 16-bit dtype expected");
  cudaStream_t st = (cudaStream_t)stream;
  const long long M = (long long)N * H * W;
  const long long n8 = M * (C / 8);
  long long b1 = (n8 + 255) / 256;
  if (b1 > 16 * kNumSMsB200) b1 = 16 * kNumSMsB200;
  const long long nstrips = (long long)N * H * ((W + kDwStrip - 1) / kDwStrip);
  const int g_dx = chan_row_grid(nstrips, C, 1, kDwThreads);
  int g_dw = chan_row_grid(nstrips, C, 4, kDwThreads);  // a few strips per thread: one block-reduced set of sums per block
  DetCtx det = det_none();
  if (pgrads) {
    const DetWs ws = det_workspace();
    if (const int gd = det_grid(ws, g_dw, 10 * C)) {
      g_dw = gd;
      det = det_ctx(ws, gd, 10 * C);
    }
  }
  if (dtype == GDL_BF16) {
    using T = __nv_bfloat16;
    GDL_LAUNCH(gelu_bwd_kernel<T>, (int)b1, 256, 0, st, (const T*)dy, (const T*)pre, (T*)dpre_scratch, n8);
    GDL_LAUNCH((dwconv_strip_kernel<T, false, true>), g_dx, kDwThreads, 0, st, (const T*)dpre_scratch, C, w, nullptr, nullptr, (T*)dx, lddx, N, H, W, C);
    if (pgrads) GDL_LAUNCH(dwconv_bwd_dw_kernel<T>, g_dw, kDwThreads, 0, st, (const T*)dpre_scratch, (const T*)x, ldx, pgrads, N, H, W, C, det);
  } else {
    using T = __half;
    GDL_LAUNCH(gelu_bwd_kernel<T>, (int)b1, 256, 0, st, (const T*)dy, (const T*)pre, (T*)dpre_scratch, n8);
    GDL_LAUNCH((dwconv_strip_kernel<T, false, true>), g_dx, kDwThreads, 0, st, (const T*)dpre_scratch, C, w, nullptr, nullptr, (T*)dx, lddx, N, H, W, C);
    if (pgrads) GDL_LAUNCH(dwconv_bwd_dw_kernel<T>, g_dw, kDwThreads, 0, st, (const T*)dpre_scratch, (const T*)x, ldx, pgrads, N, H, W, C, det);
  }
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_bilinear_fwd(const void* x, long long ldx, void* y, long long ldy, int dtype, int N, int Hi, int Wi,
                                int Ho, int Wo, int C, void* stream) {
  GDL_REQUIRE(x && y && N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && C > 0 && ldx >= C && ldy >= C, GDL_ERR_INVALID,
              "bilinear: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)N * Ho * Wo * C;
  long long b = (total + 255) / 256;
  if (b > 16 * kNumSMsB200) b = 16 * kNumSMsB200;
  const float sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
  if (dtype != GDL_F32 && C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0) {
    long long bv = (total / 8 + 255) / 256;
    if (bv > 16 * kNumSMsB200) bv = 16 * kNumSMsB200;
    if (bv < 1) bv = 1;
    if (dtype == GDL_BF16)
      GDL_LAUNCH(bilinear_fwd_vec8_kernel<__nv_bfloat16>, (int)bv, 256, 0, st, (const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)y, ldy, N, Hi, Wi, Ho, Wo, C, sh, sw);
    else
      GDL_LAUNCH(bilinear_fwd_vec8_kernel<__half>, (int)bv, 256, 0, st, (const __half*)x, ldx, (__half*)y, ldy, N, Hi, Wi, Ho, Wo, C, sh, sw);
    GDL_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  GDL_DISPATCH_T(dtype, { GDL_LAUNCH(bilinear_fwd_kernel<T>, (int)b, 256, 0, st, (const T*)x, ldx, (T*)y, ldy, N, Hi, Wi, Ho, Wo, C, sh, sw); });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_bilinear_sum_fwd(const void* base, long long ld_base, int num_src, const gdl_lowres_t* src, void* out,
                                    long long ld_out, int dtype, int N, int Ho, int Wo, int C, void* stream) {
  GDL_REQUIRE(out && N > 0 && Ho > 0 && Wo > 0 && C > 0 && ld_out >= C && num_src >= 0 && num_src <= 3 &&
                  (num_src == 0 || src != nullptr) && (base != nullptr || num_src > 0),
              GDL_ERR_INVALID, "bilinear_sum: bad args (at most 3 resized sources)");
  GDL_REQUIRE(dtype == GDL_BF16 || dtype == GDL_F16, GDL_ERR_UNSUPPORTED, "bilinear_sum: 16-bit features only");
  GDL_REQUIRE(C % 8 == 0 && ld_out % 8 == 0 && ((uintptr_t)out & 15) == 0 &&
                  (base == nullptr || (ld_base % 8 == 0 && ld_base >= C && ((uintptr_t)base & 15) == 0)),
              GDL_ERR_UNSUPPORTED, "bilinear_sum: channels and row strides must be multiples of 8, pointers 16-byte aligned");
  BilinearSumParams p;
  p.num_src = num_src;
  for (int k = 0; k < 3; ++k) {
    p.src[k] = BilinearSumSrc{nullptr, 0, 1, 1, 1.f, 1.f};
    if (k >= num_src) continue;
    GDL_REQUIRE(src[k].ptr && src[k].H > 0 && src[k].W > 0 && src[k].ld >= C && src[k].ld % 8 == 0 &&
                    ((uintptr_t)src[k].ptr & 15) == 0,
                GDL_ERR_INVALID, "bilinear_sum: bad source %d", k);
    p.src[k] = BilinearSumSrc{src[k].ptr, src[k].ld, src[k].H, src[k].W, (float)src[k].H / (float)Ho, (float)src[k].W / (float)Wo};
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)N * Ho * Wo * (C / 8);
  long long b = (total + 255) / 256;
  if (b > 16 * kNumSMsB200) b = 16 * kNumSMsB200;
  if (dtype == GDL_BF16)
    GDL_LAUNCH(bilinear_sum_vec8_kernel<__nv_bfloat16>, (int)b, 256, 0, st, (const __nv_bfloat16*)base, ld_base, p, (__nv_bfloat16*)out, ld_out, N, Ho, Wo, C);
  else
    GDL_LAUNCH(bilinear_sum_vec8_kernel<__half>, (int)b, 256, 0, st, (const __half*)base, ld_base, p, (__half*)out, ld_out, N, Ho, Wo, C);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_bilinear_bwd(const void* dy, long long ldy, void* dx, long long ldx, int dtype, int N, int Hi, int Wi,
                                int Ho, int Wo, int C, void* stream) {
  GDL_REQUIRE(dy && dx && N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && C > 0, GDL_ERR_INVALID, "bilinear_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)N * Hi * Wi * C;
  long long b = (total + 255) / 256;
  if (b > 16 * kNumSMsB200) b = 16 * kNumSMsB200;
  const float sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
  if (dtype != GDL_F32 && C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && ((uintptr_t)dx & 15) == 0 && ((uintptr_t)dy & 15) == 0) {
    long long bv = (total / 8 + 127) / 128;
    if (bv > 16 * kNumSMsB200) bv = 16 * kNumSMsB200;
    if (bv < 1) bv = 1;
    if (dtype == GDL_BF16)
      GDL_LAUNCH(bilinear_bwd_vec8_kernel<__nv_bfloat16>, (int)bv, 128, 0, st, (const __nv_bfloat16*)dy, ldy, (__nv_bfloat16*)dx, ldx, N, Hi, Wi, Ho, Wo, C, sh, sw);
    else
      GDL_LAUNCH(bilinear_bwd_vec8_kernel<__half>, (int)bv, 128, 0, st, (const __half*)dy, ldy, (__half*)dx, ldx, N, Hi, Wi, Ho, Wo, C, sh, sw);
    GDL_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  GDL_DISPATCH_T(dtype, { GDL_LAUNCH(bilinear_bwd_kernel<T>, (int)b, 256, 0, st, (const T*)dy, ldy, (T*)dx, ldx, N, Hi, Wi, Ho, Wo, C, sh, sw); });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_adaptive_avgpool_fwd(const void* x, long long ldx, void* y, int dtype, int N, int H, int W, int C,
                                        int S, void* stream) {
  // S may exceed H / W (PPM bin 6 on a 3x3 map of a small tile): ATen's window formula then yields 1-pixel windows
  GDL_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && S > 0, GDL_ERR_INVALID, "adaptive_avgpool: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)N * S * S * C;
  long long b = (total + 255) / 256;
  if (b > 8 * kNumSMsB200) b = 8 * kNumSMsB200;
  GDL_DISPATCH_T(dtype, { GDL_LAUNCH(adaptive_avgpool_fwd_kernel<T>, (int)b, 256, 0, st, (const T*)x, ldx, (T*)y, N, H, W, C, S); });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_adaptive_avgpool_bwd(const void* dy, void* dx, int dtype, int N, int H, int W, int C, int S,
                                        void* stream) {
  GDL_REQUIRE(dy && dx && N > 0 && H > 0 && W > 0 && C > 0 && S > 0, GDL_ERR_INVALID, "adaptive_avgpool_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)N * H * W * C;
  long long b = (total + 255) / 256;
  if (b > 8 * kNumSMsB200) b = 8 * kNumSMsB200;
  GDL_DISPATCH_T(dtype, { GDL_LAUNCH(adaptive_avgpool_bwd_kernel<T>, (int)b, 256, 0, st, (const T*)dy, (T*)dx, N, H, W, C, S); });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_add_nhwc(const void* a, long long lda, const void* b, long long ldb, void* y, long long ldy, int dtype,
                            long long M, int C, void* stream) {
  GDL_REQUIRE(a && b && y && M > 0 && C > 0, GDL_ERR_INVALID, "add: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  long long blk = (M * C + 255) / 256;
  if (blk > 16 * kNumSMsB200) blk = 16 * kNumSMsB200;
  GDL_DISPATCH_T(dtype, { GDL_LAUNCH(add_kernel<T>, (int)blk, 256, 0, st, (const T*)a, lda, (const T*)b, ldb, (T*)y, ldy, M, C); });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_vit_assemble_tokens(const void* patch, int patch_dtype, const float* pos, const float* cls,
                                       float* tokens, int B, int P, int C, void* stream) {
  GDL_REQUIRE(patch && pos && cls && tokens && B > 0 && P > 0 && C > 0, GDL_ERR_INVALID, "vit_assemble_tokens: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  long long b = ((long long)B * (P + 1) * C + 255) / 256;
  if (b > 16 * kNumSMsB200) b = 16 * kNumSMsB200;
  GDL_DISPATCH_T(patch_dtype, { GDL_LAUNCH(vit_assemble_tokens_kernel<T>, (int)b, 256, 0, st, (const T*)patch, pos, cls, tokens, B, P, C); });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_vit_extract_feature(const float* tokens, void* feat, int feat_dtype, int B, int P, int C, void* stream) {
  GDL_REQUIRE(tokens && feat && B > 0 && P > 0 && C > 0, GDL_ERR_INVALID, "vit_extract_feature: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  long long b = ((long long)B * P * C + 255) / 256;
  if (b > 16 * kNumSMsB200) b = 16 * kNumSMsB200;
  GDL_DISPATCH_T(feat_dtype, { GDL_LAUNCH(vit_extract_feature_kernel<T>, (int)b, 256, 0, st, tokens, (T*)feat, B, P, C); });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_cast_f32(const float* x, void* y, int dtype, long long n, void* stream) {
  GDL_REQUIRE(x && y && n > 0, GDL_ERR_INVALID, "cast: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  long long b = (n + 255) / 256;
  if (b > 16 * kNumSMsB200) b = 16 * kNumSMsB200;
  GDL_DISPATCH_T(dtype, { GDL_LAUNCH(cast_f32_kernel<T>, (int)b, 256, 0, st, x, (T*)y, n); });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

#define GDL_DISPATCH_T16(dtype, ...)                                   \
  do {                                                                 \
    if ((dtype) == GDL_BF16) {                                         \
      using T = __nv_bfloat16;                                         \
      __VA_ARGS__;                                                     \
    } else if ((dtype) == GDL_F16) {                                   \
      using T = __half;                                                \
      __VA_ARGS__;                                                     \
    } else {                                                           \
      ::gdl::set_last_error("16-bit dtype expected, got %d", dtype);   \
      return GDL_ERR_INVALID;                                          \
    }                                                                  \
  } while (0)

extern "C" int gdl_gelu_fwd(const void* x, void* y, int dtype, long long n, void* stream) {
  GDL_REQUIRE(x && y && n > 0 && n % 8 == 0, GDL_ERR_INVALID, "gelu_fwd: bad args (n %% 8 == 0 required)");
  GDL_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, GDL_ERR_INVALID,
              "gelu_fwd: 16-byte aligned buffers expected");
  long long b = (n / 8 + 255) / 256;
  if (b > 16 * kNumSMsB200) b = 16 * kNumSMsB200;
  GDL_DISPATCH_T16(dtype, { GDL_LAUNCH(gelu_fwd_kernel<T>, (int)b, 256, 0, (cudaStream_t)stream, (const T*)x, (T*)y, n / 8); });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_gelu_bwd(const void* dy, const void* pre, void* dpre, int dtype, long long n, void* stream) {
  GDL_REQUIRE(dy && pre && dpre && n > 0 && n % 8 == 0, GDL_ERR_INVALID, "gelu_bwd: bad args (n %% 8 == 0 required)");
  GDL_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(pre) | reinterpret_cast<uintptr_t>(dpre)) & 15) == 0,
              GDL_ERR_INVALID, "gelu_bwd: 16-byte aligned buffers expected");
  long long b = (n / 8 + 255) / 256;
  if (b > 16 * kNumSMsB200) b = 16 * kNumSMsB200;
  GDL_DISPATCH_T16(dtype, { GDL_LAUNCH(gelu_bwd_kernel<T>, (int)b, 256, 0, (cudaStream_t)stream, (const T*)dy, (const T*)pre, (T*)dpre, n / 8); });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_layerscale_add(const float* res, const void* u, int dtype, const float* gamma, const float* sscale,
                                  long long rows_per_sample, float* out, long long M, int C, void* stream) {
  GDL_REQUIRE(res && u && gamma && out && M > 0 && C > 0 && C % 8 == 0, GDL_ERR_INVALID, "layerscale_add: bad args (C %% 8 == 0)");
  GDL_REQUIRE(sscale == nullptr || rows_per_sample > 0, GDL_ERR_INVALID, "layerscale_add: rows_per_sample must be positive");
  GDL_REQUIRE(((reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
              GDL_ERR_INVALID, "layerscale_add: 16-byte aligned buffers expected");
  long long b = (M * (C / 8) + 255) / 256;
  if (b > 16 * kNumSMsB200) b = 16 * kNumSMsB200;
  GDL_DISPATCH_T16(dtype, {
    GDL_LAUNCH(layerscale_add_kernel<T>, (int)b, 256, 0, (cudaStream_t)stream, res, (const T*)u, gamma, sscale,
                                                                      rows_per_sample > 0 ? rows_per_sample : 1, out, M, C);
  });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_layerscale_bwd(const float* g, const void* u, int dtype, const float* gamma, const float* sscale,
                                  long long rows_per_sample, void* du, float* dgamma, long long M, int C, void* stream) {
  GDL_REQUIRE(g && u && gamma && du && M > 0 && C > 0 && C % 8 == 0 && C <= 2048, GDL_ERR_INVALID,
              "layerscale_bwd: bad args (C %% 8 == 0, C <= 2048)");
  GDL_REQUIRE(sscale == nullptr || rows_per_sample > 0, GDL_ERR_INVALID, "layerscale_bwd: rows_per_sample must be positive");
  GDL_REQUIRE(((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(du)) & 15) == 0,
              GDL_ERR_INVALID, "layerscale_bwd: 16-byte aligned buffers expected");
  const int rpb = 256 / (C / 8);
  long long b = (M + (long long)rpb * 8 - 1) / ((long long)rpb * 8);  // >= 8 rows per thread: few atomics per channel
  if (b > 2 * kNumSMsB200) b = 2 * kNumSMsB200;
  if (b < 1) b = 1;
  DetCtx det = det_none();
  if (dgamma) {
    const DetWs ws = det_workspace();
    if (const int gd = det_grid(ws, b, C)) {
      b = gd;
      det = det_ctx(ws, gd, C);
    }
  }
  GDL_DISPATCH_T16(dtype, {
    GDL_LAUNCH(layerscale_bwd_kernel<T>, (int)b, 256, 0, (cudaStream_t)stream, g, (const T*)u, gamma, sscale,
                                                                      rows_per_sample > 0 ? rows_per_sample : 1, (T*)du, dgamma, M, C, det);
  });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_vit_feature_grad(const void* dfeat, int dtype, float* g, int B, int P, int C, int init, void* stream) {
  GDL_REQUIRE(dfeat && g && B > 0 && P > 0 && C > 0, GDL_ERR_INVALID, "vit_feature_grad: bad args");
  long long b = ((long long)B * (P + 1) * C + 255) / 256;
  if (b > 16 * kNumSMsB200) b = 16 * kNumSMsB200;
  GDL_DISPATCH_T(dtype, { GDL_LAUNCH(vit_feature_grad_kernel<T>, (int)b, 256, 0, (cudaStream_t)stream, (const T*)dfeat, g, B, P, C, init); });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static bool pow2_le32(int v) { return v >= 1 && v <= 32 && (v & (v - 1)) == 0; }

extern "C" int gdl_channel_pool_fwd(const void* xw, const float* scores, void* out, float* attn, int dtype, long long P,
                                    int C, int E, void* stream) {
  GDL_REQUIRE(xw && scores && out && P > 0 && C >= 1 && C <= kMaxBands && E >= 8 && E % 8 == 0 && pow2_le32(E / 8),
              GDL_ERR_INVALID, "channel_pool_fwd: bad args (1 <= C <= 16 bands, E/8 a power of two <= 32)");
  GDL_REQUIRE(((reinterpret_cast<uintptr_t>(xw) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, GDL_ERR_INVALID,
              "channel_pool_fwd: 16-byte aligned buffers expected");
  long long b = (P * (E / 8) + 255) / 256;
  if (b > 16 * kNumSMsB200) b = 16 * kNumSMsB200;
  GDL_DISPATCH_T16(dtype, {
    GDL_LAUNCH(channel_pool_fwd_kernel<T>, (int)b, 256, 0, (cudaStream_t)stream, (const T*)xw, scores, (T*)out, attn, P, C, E);
  });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_channel_pool_bwd(const void* dout, const void* xw, const float* attn, void* dxw, void* dscores, int dtype,
                                    long long P, int C, int E, void* stream) {
  GDL_REQUIRE(dout && xw && attn && dxw && dscores && P > 0 && C >= 1 && C <= kMaxBands && E >= 8 && E % 8 == 0 &&
              pow2_le32(E / 8), GDL_ERR_INVALID, "channel_pool_bwd: bad args (1 <= C <= 16 bands, E/8 a power of two <= 32)");
  GDL_REQUIRE(((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(xw) | reinterpret_cast<uintptr_t>(dxw) |
                reinterpret_cast<uintptr_t>(dscores)) & 15) == 0, GDL_ERR_INVALID, "channel_pool_bwd: 16-byte aligned buffers expected");
  long long b = (P * (E / 8) + 255) / 256;
  if (b > 16 * kNumSMsB200) b = 16 * kNumSMsB200;
  GDL_DISPATCH_T16(dtype, {
    GDL_LAUNCH(channel_pool_bwd_kernel<T>, (int)b, 256, 0, (cudaStream_t)stream, (const T*)dout, (const T*)xw, attn, (T*)dxw,
                                                                        (T*)dscores, P, C, E);
  });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_relu_bwd(const void* dy, const void* y, void* dx, int dtype, long long n, void* stream) {
  GDL_REQUIRE(dy && y && dx && n > 0 && n % 8 == 0, GDL_ERR_INVALID, "relu_bwd: bad args (n %% 8 == 0 required)");
  GDL_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0,
              GDL_ERR_INVALID, "relu_bwd: 16-byte aligned buffers expected");
  long long b = (n / 8 + 255) / 256;
  if (b > 16 * kNumSMsB200) b = 16 * kNumSMsB200;
  GDL_DISPATCH_T16(dtype, { GDL_LAUNCH(relu_bwd_kernel<T>, (int)b, 256, 0, (cudaStream_t)stream, (const T*)dy, (const T*)y, (T*)dx, n / 8); });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}
