// elementwise.cu — the HBM-bound kernels around the tensor-core convs: input normalisation,
// im2col/col2im for strided convs, BatchNorm (train/eval) statistics / apply / backward,
// gradient gather (+ReLU mask, +2x2 sum-pool of an upsampled consumer), max-pool.
//
// All activations are NHWC 16-bit with C % 8 == 0 so every thread moves 128-bit vectors
// (8 channels); rows are addressed as x[row * ld + c].  Grids are sized in multiples of the
// SM count and grid-stride over rows.
#include <string.h>

#include "../../include/gdl_b200.h"
#include "common.cuh"
#include "det_reduce.cuh"

namespace gdl {

// ------------------------------------------------------------------------------------------
// 8-wide 16-bit vector helpers
// ------------------------------------------------------------------------------------------
template <typename T>
struct Vec8;

template <>
struct Vec8<__nv_bfloat16> {
  static GDL_DEVINL void unpack(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = bf16_lo(w[i]);
      f[2 * i + 1] = bf16_hi(w[i]);
    }
  }
  static GDL_DEVINL uint4 pack(const float (&f)[8]) {
    return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                      pack_bf16x2(f[6], f[7]));
  }
  static GDL_DEVINL float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
  static GDL_DEVINL __nv_bfloat16 from_float(float v) { return __float2bfloat16_rn(v); }
};
template <>
struct Vec8<__half> {
  static GDL_DEVINL void unpack(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      float2 t = __half22float2(h);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  static GDL_DEVINL uint4 pack(const float (&f)[8]) {
    return make_uint4(pack_f16x2(f[0], f[1]), pack_f16x2(f[2], f[3]), pack_f16x2(f[4], f[5]),
                      pack_f16x2(f[6], f[7]));
  }
  static GDL_DEVINL float to_float(__half v) { return __half2float(v); }
  static GDL_DEVINL __half from_float(float v) { return __float2half_rn(v); }
};

template <typename T>
GDL_DEVINL void load8(const T* p, float (&f)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  Vec8<T>::unpack(u, f);
}
template <typename T>
GDL_DEVINL void store8(T* p, const float (&f)[8]) {
  *reinterpret_cast<uint4*>(p) = Vec8<T>::pack(f);
}

static int ew_blocks(long long work_items, int threads, int max_waves = 8) {
  long long b = (work_items + threads - 1) / threads;
  long long cap = (long long)kNumSMsB200 * max_waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// grid for the "one 8-channel vector per thread, stride over rows" kernels (256 threads / block)
static int row_grid(long long rows, int C) {
  const int cv = C / 8;
  const int tpr = cv < 256 ? cv : 256;
  const int rpb = 256 / tpr;
  long long b = (rows + (long long)rpb * 4 - 1) / ((long long)rpb * 4);  // >= 4 rows per thread
  const long long cap = (long long)kNumSMsB200 * 8;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

#define GDL_DISPATCH_16(dtype, ...)                                  \
  do {                                                               \
    if ((dtype) == GDL_BF16) {                                       \
      using T = __nv_bfloat16;                                       \
      __VA_ARGS__;                                                   \
    } else if ((dtype) == GDL_F16) {                                 \
      using T = __half;                                              \
      __VA_ARGS__;                                                   \
    } else {                                                         \
      ::gdl::set_last_error("16-bit dtype expected, got %d", dtype); \
      return GDL_ERR_INVALID;                                        \
    }                                                                \
  } while (0)

// ------------------------------------------------------------------------------------------
// input normalisation: y = ((x / 255) - mean_c) / std_c, uint8 (or f32) HWC -> 16-bit NHWC.
// Arithmetic order follows utils/tensors.py:10-35 (true divisions, fp32).
// ------------------------------------------------------------------------------------------
template <typename T, typename TIn>
__global__ void normalize_kernel(const TIn* __restrict__ x, T* __restrict__ y, long long npix, int C,
                                 int ld, const float* __restrict__ mean, const float* __restrict__ stdv,
                                 float image_max, int in_is_chw, long long hw) {
  GDL_PDL_ENTRY();
  const long long total = npix * ld;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / ld;
    const int c = (int)(i - pix * ld);
    float v = 0.f;
    if (c < C) {
      float raw;
      if (in_is_chw) {
        const long long n = pix / hw, p = pix - n * hw;
        raw = (float)x[(n * C + c) * hw + p];
      } else {
        raw = (float)x[pix * C + c];
      }
      v = raw;
      if (image_max > 0.f) v = __fdiv_rn(raw, image_max);
      if (mean != nullptr) v = __fdiv_rn(v - mean[c], stdv[c]);
    }
    y[i] = Vec8<T>::from_float(v);
  }
}

// ------------------------------------------------------------------------------------------
// im2col (for strided convs and the C_in = 3/4/6 stem): col[(n,ho,wo)][(r,s,c)], zero padded
// to Kpad columns.  One thread moves one 16-byte (8 element) chunk.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void im2col_vec8_kernel(const T* __restrict__ x, T* __restrict__ col, int N, int H, int W,
                                   int C, int ld, int R, int S, int stride, int pad, int Ho, int Wo,
                                   int Kpad) {
  GDL_PDL_ENTRY();
  const int cv = C / 8;
  const int kchunks = Kpad / 8;
  const long long total = (long long)N * Ho * Wo * kchunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / kchunks;
    const int kc = (int)(i - row * kchunks);
    const int tap = kc / cv, c8 = kc - tap * cv;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (tap < R * S) {
      const int r = tap / S, s = tap - r * S;
      const int wo = (int)(row % Wo);
      const long long t = row / Wo;
      const int ho = (int)(t % Ho);
      const int n = (int)(t / Ho);
      const int h = ho * stride - pad + r, w = wo * stride - pad + s;
      if (h >= 0 && h < H && w >= 0 && w < W)
        v = *reinterpret_cast<const uint4*>(x + (((long long)n * H + h) * W + w) * ld + c8 * 8);
    }
    *reinterpret_cast<uint4*>(col + row * Kpad + kc * 8) = v;
  }
}

// 4-channel input stored with pixel stride 4 or 8 (the 4-band stem): one thread moves two taps (2 x 8 B)
template <typename T>
__global__ void im2col_c4_kernel(const T* __restrict__ x, T* __restrict__ col, int N, int H, int W, int ld, int R,
                                 int S, int stride, int pad, int Ho, int Wo, int Kpad) {
  GDL_PDL_ENTRY();
  const int pairs = Kpad / 8;  // 8 elements = 2 taps per thread
  const int taps = R * S;
  const long long total = (long long)N * Ho * Wo * pairs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / pairs;
    const int pr = (int)(i - row * pairs);
    const int wo = (int)(row % Wo);
    const long long t = row / Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    uint2 v[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int tap = pr * 2 + j;
      v[j] = make_uint2(0, 0);
      if (tap < taps) {
        const int r = tap / S, s2 = tap - r * S;
        const int h = ho * stride - pad + r, w = wo * stride - pad + s2;
        if (h >= 0 && h < H && w >= 0 && w < W)
          v[j] = *reinterpret_cast<const uint2*>(x + (((long long)n * H + h) * W + w) * ld);
      }
    }
    *reinterpret_cast<uint4*>(col + row * Kpad + pr * 8) = make_uint4(v[0].x, v[0].y, v[1].x, v[1].y);
  }
}

// generic (any C): one thread per output element
template <typename T>
__global__ void im2col_scalar_kernel(const T* __restrict__ x, T* __restrict__ col, int N, int H, int W,
                                     int C, int ld, int R, int S, int stride, int pad, int Ho, int Wo,
                                     int Kpad) {
  GDL_PDL_ENTRY();
  const long long total = (long long)N * Ho * Wo * Kpad;
  const int K = R * S * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / Kpad;
    const int k = (int)(i - row * Kpad);
    T v = Vec8<T>::from_float(0.f);
    if (k < K) {
      const int tap = k / C, c = k - tap * C;
      const int r = tap / S, s = tap - r * S;
      const int wo = (int)(row % Wo);
      const long long t = row / Wo;
      const int ho = (int)(t % Ho);
      const int n = (int)(t / Ho);
      const int h = ho * stride - pad + r, w = wo * stride - pad + s;
      if (h >= 0 && h < H && w >= 0 && w < W) v = x[(((long long)n * H + h) * W + w) * ld + c];
    }
    col[i] = v;
  }
}

// col2im (gather form): dx[n,h,w,c] = sum over taps of dcol[(n,ho,wo)][(r,s,c)]
template <typename T>
__global__ void col2im_vec8_kernel(const T* __restrict__ dcol, T* __restrict__ dx, int N, int H, int W,
                                   int C, int ld, int R, int S, int stride, int pad, int Ho, int Wo,
                                   int Kpad) {
  GDL_PDL_ENTRY();
  const int cv = C / 8;
  const long long total = (long long)N * H * W * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int c8 = (int)(i - pix * cv);
    const int w = (int)(pix % W);
    const long long t = pix / W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < R; ++r) {
      const int hn = h + pad - r;
      if (hn < 0 || hn % stride) continue;
      const int ho = hn / stride;
      if (ho >= Ho) continue;
      for (int s = 0; s < S; ++s) {
        const int wn = w + pad - s;
        if (wn < 0 || wn % stride) continue;
        const int wo = wn / stride;
        if (wo >= Wo) continue;
        float f[8];
        load8(dcol + (((long long)n * Ho + ho) * Wo + wo) * Kpad + (r * S + s) * C + c8 * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    }
    store8(dx + pix * ld + c8 * 8, acc);
  }
}

// ------------------------------------------------------------------------------------------
// BatchNorm statistics.  x: [M][C] (ld).  Per channel: pivot p_c (caller supplied, e.g. the
// running mean: identical on every rank so partial sums can be all-reduced for SyncBN),
//   S1 = sum(x - p), S2 = sum((x - p)^2)   (pivoting keeps S2/M - (S1/M)^2 well conditioned)
// Block = 256 threads = (C/8 channel-vectors) x (rows in flight); each thread strides rows.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void bn_stats_kernel(const T* __restrict__ x, long long M, int C, int ld,
                                float* __restrict__ sums /* [2][C], pre-zeroed */,
                                const float* __restrict__ pivot /* [C] or null */, const DetCtx det) {
  GDL_PDL_ENTRY();
  const int cv = C / 8;
  const int tpr = cv < (int)blockDim.x ? cv : blockDim.x;  // threads spanning the channel dim
  const int rows_per_block = blockDim.x / tpr;
  const int tc = threadIdx.x % tpr;
  const int tr = threadIdx.x / tpr;
  extern __shared__ float red[];  // [rows_per_block][tpr*8][2] reduced at the end
  for (int cbase = 0; cbase < cv; cbase += tpr) {  // uniform trip count (barriers inside)
    const int c8 = cbase + tc;
    const bool active = (c8 < cv) && (tr < rows_per_block);
    float piv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (active && pivot != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) piv[j] = pivot[c8 * 8 + j];
    }
    float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (active) {
      // 4 independent 16-byte loads in flight per thread (a single one measured 3.6 TB/s in run 8); the 4 rows of a
      // thread are rows_per_block apart, so a block reads one contiguous 4*rows_per_block-row chunk per iteration
      // (4 distant fronts per thread measured slower than the single load: run 9)
      constexpr int kInFlight = 8;  // round 2: 8 (was 4) loads in flight per thread — 2.0 TB/s measured with 4 (0.31 of peak)
      const long long chunk = (long long)kInFlight * rows_per_block;
      long long base = (long long)blockIdx.x * chunk;
      for (; base + chunk <= M; base += (long long)gridDim.x * chunk) {
        uint4 u[kInFlight];
#pragma unroll
        for (int q = 0; q < kInFlight; ++q)
          u[q] = *reinterpret_cast<const uint4*>(x + (base + q * rows_per_block + tr) * ld + c8 * 8);
#pragma unroll
        for (int q = 0; q < kInFlight; ++q) {
          float f[8];
          Vec8<T>::unpack(u[q], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d = f[j] - piv[j];
            s1[j] += d;
            s2[j] = fmaf(d, d, s2[j]);
          }
        }
      }
      if (base < M) {  // ragged last chunk (at most one block sees it)
        for (long long row = base + tr; row < M; row += rows_per_block) {
          float f[8];
          load8(x + row * ld + c8 * 8, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d = f[j] - piv[j];
            s1[j] += d;
            s2[j] = fmaf(d, d, s2[j]);
          }
        }
      }
    }
    // reduce across the rows_per_block threads that share this channel vector
    __syncthreads();
    if (tr < rows_per_block) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        red[((tr * tpr + tc) * 8 + j) * 2 + 0] = s1[j];
        red[((tr * tpr + tc) * 8 + j) * 2 + 1] = s2[j];
      }
    }
    __syncthreads();
    if (tr == 0 && c8 < cv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = 0.f, b = 0.f;
        for (int r = 0; r < rows_per_block; ++r) {
          a += red[((r * tpr + tc) * 8 + j) * 2 + 0];
          b += red[((r * tpr + tc) * 8 + j) * 2 + 1];
        }
        if (det.s0 != nullptr) {  // ordered reduction: this block's slot, summed in block order by det_finish
          det_put(det, 2 * C, c8 * 8 + j, a);
          det_put(det, 2 * C, C + c8 * 8 + j, b);
        } else {
          atomicAdd(&sums[c8 * 8 + j], a);
          atomicAdd(&sums[C + c8 * 8 + j], b);
        }
      }
    }
  }
  if (det.s0 != nullptr) det_finish(det, 2 * C, sums);
}

// finalize: mean/var -> (scale, shift) for the apply kernel, saved (mean, invstd) for backward,
// running statistics update with momentum and unbiased variance (nn.BatchNorm2d semantics).
__global__ void bn_finalize_kernel(const float* pivot /* may alias running_mean */, const float* __restrict__ sums, long long M,
                                   int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float eps, float momentum, float* running_mean,
                                   float* __restrict__ running_var, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ save_mean,
                                   float* __restrict__ save_invstd) {
  GDL_PDL_ENTRY();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float piv = pivot ? pivot[c] : 0.f;
  const float invM = 1.0f / (float)M;
  const float m1 = sums[c] * invM;
  float var = sums[C + c] * invM - m1 * m1;
  var = fmaxf(var, 0.f);
  const float mean = piv + m1;
  const float invstd = rsqrtf(var + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * invstd;
  shift[c] = b - mean * g * invstd;
  save_mean[c] = mean;
  save_invstd[c] = invstd;
  if (running_mean) {
    const float unb = M > 1 ? var * ((float)M / (float)(M - 1)) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unb;
  }
}

// eval-mode: scale/shift from running statistics
__global__ void bn_eval_coeffs_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                      float* __restrict__ scale, float* __restrict__ shift) {
  GDL_PDL_ENTRY();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = rsqrtf(rv[c] + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * invstd;
  shift[c] = b - rm[c] * g * invstd;
}

// ------------------------------------------------------------------------------------------
// BN apply (+ residual, + ReLU, + optional nearest x2 upsampled second output)
//   y = act( x*scale + shift + [ res*rscale + rshift | res ] )
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void bn_apply_kernel(const T* __restrict__ x, int ldx, const float* __restrict__ scale,
                                const float* __restrict__ shift, const T* __restrict__ res, int ldr,
                                const float* __restrict__ rscale, const float* __restrict__ rshift,
                                int relu, T* __restrict__ y, int ldy, T* __restrict__ y_up, int ldu,
                                int N, int H, int W, int C) {
  GDL_PDL_ENTRY();
  // each thread owns one 8-channel vector (coefficients live in registers) and strides over pixels
  const int cv = C / 8;
  const int tpr = cv < (int)blockDim.x ? cv : blockDim.x;
  const int rows_per_block = blockDim.x / tpr;
  const int tc = threadIdx.x % tpr;
  const int tr = threadIdx.x / tpr;
  if (tr >= rows_per_block) return;
  const long long M = (long long)N * H * W;
  const long long W2 = 2ll * W;
  for (int c8 = tc; c8 < cv; c8 += tpr) {
    const int c0 = c8 * 8;
    float sc[8], sh[8], rs[8], rh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = scale[c0 + j];
      sh[j] = shift[c0 + j];
      rs[j] = rscale ? rscale[c0 + j] : 1.f;
      rh[j] = rshift ? rshift[c0 + j] : 0.f;
    }
    for (long long pix = (long long)blockIdx.x * rows_per_block + tr; pix < M;
         pix += (long long)gridDim.x * rows_per_block) {
      float f[8];
      load8(x + pix * ldx + c0, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sh[j]);
      if (res != nullptr) {
        float g[8];
        load8(res + pix * ldr + c0, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += fmaf(g[j], rs[j], rh[j]);
      }
      if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
      }
      const uint4 packed = Vec8<T>::pack(f);
      if (y != nullptr) *reinterpret_cast<uint4*>(y + pix * ldy + c0) = packed;
      if (y_up != nullptr) {
        const int w = (int)(pix % W);
        const long long t = pix / W;
        const int h = (int)(t % H);
        const long long n = t / H;
        T* u = y_up + ((n * 2 * H + 2 * h) * W2 + 2 * w) * ldu + c0;
        *reinterpret_cast<uint4*>(u) = packed;
        *reinterpret_cast<uint4*>(u + ldu) = packed;
        *reinterpret_cast<uint4*>(u + W2 * ldu) = packed;
        *reinterpret_cast<uint4*>(u + (W2 + 1) * ldu) = packed;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// gradient gather: g = (sum_k src_k) [* (y > 0)], optionally accumulating the BN-backward sums
//   sum_g[c] and sum_gxhat[c] = sum g * (x - mean) * invstd  in the same pass.
// A source with mode 1 is the gradient of a nearest-x2-upsampled copy (shape 2H x 2W): its
// contribution is the 2x2 sum.
// ------------------------------------------------------------------------------------------
struct GatherSrcs {
  const void* ptr[GDL_MAX_SRC];
  int ld[GDL_MAX_SRC];
  int mode[GDL_MAX_SRC];
  int n;
};

// NS = number of gradient sources (compile time: the loads of one iteration are all issued before the first use,
// 2..20 independent 16-byte loads in flight per thread — run 8 measured the previous sequential version at 2.1 TB/s,
// a third of what bn_apply reaches).  Narrow gathers (NS <= 2) take two pixels per iteration.
// NUP (round 2): how the sources are laid out.  0 / 1 = the first NS - NUP sources are plain (mode 0) and the LAST one is
// the gradient of the nearest-x2 copy (mode 1, a 2x2 sum: four loads) — every gather of a UNet++ / UperNet step has this
// form — so the staging registers are 1 vector per plain source instead of 4 for every source: 155 -> ~90 registers at
// NS = 2 (one resident block of 256 threads per SM -> two or three; the round-1 kernel ran at 0.55 of the copy peak with
// 2+ sources, 0.75 with one).  -1 = any mix of modes (4 vectors staged per source).
template <typename T, int NS, int NUP>
__global__ void __launch_bounds__(256) grad_gather_kernel(GatherSrcs srcs, const T* __restrict__ y, int ldy,
                                                          const T* __restrict__ x, int ldx,
                                                          const float* __restrict__ mean,
                                                          const float* __restrict__ invstd, T* __restrict__ g, int ldg,
                                                          float* __restrict__ sums /* [2][C] or null */, int N, int H,
                                                          int W, int C, const DetCtx det) {
  GDL_PDL_ENTRY();
  constexpr bool kGeneric = NUP < 0;
  constexpr int N0 = kGeneric ? NS : NS - NUP;        // plain sources (generic: every source, staged with 4 vectors)
  constexpr int NU = kGeneric ? 0 : NUP;              // trailing 2x2-pooled sources
  constexpr int V0 = kGeneric ? 4 : 1;
  constexpr int PIX = NS <= 2 ? 2 : 1;
  const int cv = C / 8;
  const int tpr = cv < (int)blockDim.x ? cv : blockDim.x;
  const int rows_per_block = blockDim.x / tpr;
  const int tc = threadIdx.x % tpr;
  const int tr = threadIdx.x / tpr;
  const long long M = (long long)N * H * W;
  extern __shared__ float red[];
  bool any_up = NU > 0;
  if (kGeneric) {
#pragma unroll
    for (int k = 0; k < NS; ++k) any_up |= srcs.mode[k] != 0;
  }
  for (int cbase = 0; cbase < cv; cbase += tpr) {  // uniform trip count (barriers inside)
    const int c8 = cbase + tc;
    const bool active = (c8 < cv) && (tr < rows_per_block);
    const int c0 = c8 * 8;
    float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float mu[8] = {0, 0, 0, 0, 0, 0, 0, 0}, is[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (sums != nullptr && active) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        mu[j] = mean[c0 + j];
        is[j] = invstd[c0 + j];
      }
    }
    if (active) {
      // the PIX pixels of a thread are rows_per_block apart: a block covers one contiguous PIX*rows_per_block chunk
      const long long stride = rows_per_block;
      for (long long pix0 = (long long)blockIdx.x * PIX * rows_per_block + tr; pix0 < M;
           pix0 += (long long)gridDim.x * PIX * rows_per_block) {
        uint4 raw[PIX][N0 > 0 ? N0 : 1][V0];
        uint4 rawu[PIX][NU > 0 ? NU : 1][4];
        uint4 yv[PIX], xv[PIX];
        // ---- issue every load of this iteration ----
#pragma unroll
        for (int pp = 0; pp < PIX; ++pp) {
          const long long pix = pix0 + pp * stride;
          if (pix >= M) continue;
          long long up_off = 0;  // pixel offset of the 2x2 block in a (N, 2H, 2W) source
          if (any_up) {
            const unsigned pu = (unsigned)pix;  // M < 2^31 (checked by the host wrapper)
            const unsigned w = pu % (unsigned)W, t = pu / (unsigned)W;
            const unsigned h = t % (unsigned)H, n = t / (unsigned)H;
            up_off = ((long long)(n * 2 * H + 2 * h)) * (2ll * W) + 2 * w;
          }
#pragma unroll
          for (int k = 0; k < N0; ++k) {
            const T* sp = reinterpret_cast<const T*>(srcs.ptr[k]);
            const long long ld = srcs.ld[k];
            if (!kGeneric || srcs.mode[k] == 0) {
              raw[pp][k][0] = *reinterpret_cast<const uint4*>(sp + pix * ld + c0);
            } else {
              const T* b = sp + up_off * ld + c0;
              raw[pp][k][0] = *reinterpret_cast<const uint4*>(b);
              raw[pp][k][V0 > 1 ? 1 : 0] = *reinterpret_cast<const uint4*>(b + ld);
              raw[pp][k][V0 > 2 ? 2 : 0] = *reinterpret_cast<const uint4*>(b + 2ll * W * ld);
              raw[pp][k][V0 > 3 ? 3 : 0] = *reinterpret_cast<const uint4*>(b + (2ll * W + 1) * ld);
            }
          }
#pragma unroll
          for (int k = 0; k < NU; ++k) {
            const T* sp = reinterpret_cast<const T*>(srcs.ptr[N0 + k]);
            const long long ld = srcs.ld[N0 + k];
            const T* b = sp + up_off * ld + c0;
            rawu[pp][k][0] = *reinterpret_cast<const uint4*>(b);
            rawu[pp][k][1] = *reinterpret_cast<const uint4*>(b + ld);
            rawu[pp][k][2] = *reinterpret_cast<const uint4*>(b + 2ll * W * ld);
            rawu[pp][k][3] = *reinterpret_cast<const uint4*>(b + (2ll * W + 1) * ld);
          }
          if (y != nullptr) yv[pp] = *reinterpret_cast<const uint4*>(y + pix * ldy + c0);
          if (sums != nullptr) xv[pp] = *reinterpret_cast<const uint4*>(x + pix * ldx + c0);
        }
        // ---- reduce, mask, store, statistics ----
#pragma unroll
        for (int pp = 0; pp < PIX; ++pp) {
          const long long pix = pix0 + pp * stride;
          if (pix >= M) continue;
          float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
          for (int k = 0; k < N0; ++k) {
            float f[8];
            Vec8<T>::unpack(raw[pp][k][0], f);
            if (!kGeneric || srcs.mode[k] == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] += f[j];
            } else {
              float f1[8], f2[8], f3[8];
              Vec8<T>::unpack(raw[pp][k][V0 > 1 ? 1 : 0], f1);
              Vec8<T>::unpack(raw[pp][k][V0 > 2 ? 2 : 0], f2);
              Vec8<T>::unpack(raw[pp][k][V0 > 3 ? 3 : 0], f3);
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] += ((f[j] + f1[j]) + f2[j]) + f3[j];
            }
          }
#pragma unroll
          for (int k = 0; k < NU; ++k) {
            float f[8], f1[8], f2[8], f3[8];
            Vec8<T>::unpack(rawu[pp][k][0], f);
            Vec8<T>::unpack(rawu[pp][k][1], f1);
            Vec8<T>::unpack(rawu[pp][k][2], f2);
            Vec8<T>::unpack(rawu[pp][k][3], f3);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += ((f[j] + f1[j]) + f2[j]) + f3[j];
          }
          if (y != nullptr) {
            float yy[8];
            Vec8<T>::unpack(yv[pp], yy);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = yy[j] > 0.f ? acc[j] : 0.f;
          }
          // the gradient that flows on is the 16-bit rounded one: use it for the sums too
          const uint4 packed = Vec8<T>::pack(acc);
          if (g != nullptr) *reinterpret_cast<uint4*>(g + pix * ldg + c0) = packed;
          if (sums != nullptr) {
            float gr[8], xx[8];
            Vec8<T>::unpack(packed, gr);
            Vec8<T>::unpack(xv[pp], xx);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              s1[j] += gr[j];
              s2[j] = fmaf(gr[j], (xx[j] - mu[j]) * is[j], s2[j]);
            }
          }
        }
      }
    }
    if (sums != nullptr) {
      __syncthreads();
      if (tr < rows_per_block) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          red[((tr * tpr + tc) * 8 + j) * 2 + 0] = s1[j];
          red[((tr * tpr + tc) * 8 + j) * 2 + 1] = s2[j];
        }
      }
      __syncthreads();
      if (tr == 0 && c8 < cv) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = 0.f, b = 0.f;
          for (int r = 0; r < rows_per_block; ++r) {
            a += red[((r * tpr + tc) * 8 + j) * 2 + 0];
            b += red[((r * tpr + tc) * 8 + j) * 2 + 1];
          }
          if (det.s0 != nullptr) {
            det_put(det, 2 * C, c0 + j, a);
            det_put(det, 2 * C, C + c0 + j, b);
          } else {
            atomicAdd(&sums[c0 + j], a);
            atomicAdd(&sums[C + c0 + j], b);
          }
        }
      }
    }
  }
  if (sums != nullptr && det.s0 != nullptr) det_finish(det, 2 * C, sums);
}

// BN backward, elementwise part:  dx = gamma*invstd * (g - sum_g/M - xhat * sum_gxhat/M)
// (in place on g allowed); also emits dgamma = sum_gxhat, dbeta = sum_g (accumulated).
template <typename T>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ g, int ldg, const T* __restrict__ x, int ldx,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gamma, const float* __restrict__ sums,
                                    T* __restrict__ dx, int ldd, long long M, long long count, int C) {
  GDL_PDL_ENTRY();
  // dx = a*g + b*x + c with per-channel a = gamma*invstd, b = -a*invstd*s2/M, c = a*(mean*invstd*s2/M - s1/M)
  const int cv = C / 8;
  const int tpr = cv < (int)blockDim.x ? cv : blockDim.x;
  const int rows_per_block = blockDim.x / tpr;
  const int tc = threadIdx.x % tpr;
  const int tr = threadIdx.x / tpr;
  if (tr >= rows_per_block) return;
  const float invM = 1.0f / (float)count;
  for (int c8 = tc; c8 < cv; c8 += tpr) {
    const int c0 = c8 * 8;
    float ca[8], cb[8], cc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      const float is = invstd[c];
      const float a = (gamma ? gamma[c] : 1.f) * is;
      const float s1 = sums[c] * invM, s2 = sums[C + c] * invM;
      ca[j] = a;
      cb[j] = -a * is * s2;
      cc[j] = a * (mean[c] * is * s2 - s1);
    }
    const long long stride = (long long)gridDim.x * rows_per_block;
    long long pix = (long long)blockIdx.x * rows_per_block + tr;
    for (; pix + stride < M; pix += 2 * stride) {  // two rows in flight per thread
      float g0[8], x0[8], g1[8], x1[8], o0[8], o1[8];
      load8(g + pix * ldg + c0, g0);
      load8(x + pix * ldx + c0, x0);
      load8(g + (pix + stride) * ldg + c0, g1);
      load8(x + (pix + stride) * ldx + c0, x1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o0[j] = fmaf(ca[j], g0[j], fmaf(cb[j], x0[j], cc[j]));
        o1[j] = fmaf(ca[j], g1[j], fmaf(cb[j], x1[j], cc[j]));
      }
      store8(dx + pix * ldd + c0, o0);
      store8(dx + (pix + stride) * ldd + c0, o1);
    }
    if (pix < M) {
      float g0[8], x0[8], o0[8];
      load8(g + pix * ldg + c0, g0);
      load8(x + pix * ldx + c0, x0);
#pragma unroll
      for (int j = 0; j < 8; ++j) o0[j] = fmaf(ca[j], g0[j], fmaf(cb[j], x0[j], cc[j]));
      store8(dx + pix * ldd + c0, o0);
    }
  }
}

__global__ void bn_param_grads_kernel(const float* __restrict__ sums, int C, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta, int accumulate) {
  GDL_PDL_ENTRY();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + sums[C + c];
  if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + sums[c];
}

// ------------------------------------------------------------------------------------------
// max-pool 3x3 / stride 2 / pad 1 (ResNet stem), NHWC.  The argmax tap (first maximum in
// (r,s) scan order, as ATen) is saved as one byte per output element for the backward.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool3x3s2_fwd_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y, int ldy,
                                        uint8_t* __restrict__ idx, int N, int H, int W, int C, int Ho, int Wo) {
  GDL_PDL_ENTRY();
  const int cv = C / 8;
  const long long total = (long long)N * Ho * Wo * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long opix = i / cv;
    const int c0 = (int)(i - opix * cv) * 8;
    const int wo = (int)(opix % Wo);
    const long long t = opix / Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float best[8];
    int bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = -INFINITY;
      bi[j] = 0;
    }
    bool first = true;
    for (int r = 0; r < 3; ++r) {
      const int h = 2 * ho - 1 + r;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int w = 2 * wo - 1 + s;
        if (w < 0 || w >= W) continue;
        float f[8];
        load8(x + (((long long)n * H + h) * W + w) * ldx + c0, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (first || f[j] > best[j]) {
            best[j] = f[j];
            bi[j] = r * 3 + s;
          }
        }
        first = false;
      }
    }
    store8(y + opix * ldy + c0, best);
    if (idx != nullptr) {
      uint2 p;
      p.x = (uint32_t)bi[0] | ((uint32_t)bi[1] << 8) | ((uint32_t)bi[2] << 16) | ((uint32_t)bi[3] << 24);
      p.y = (uint32_t)bi[4] | ((uint32_t)bi[5] << 8) | ((uint32_t)bi[6] << 16) | ((uint32_t)bi[7] << 24);
      *reinterpret_cast<uint2*>(idx + opix * C + c0) = p;
    }
  }
}

template <typename T>
__global__ void maxpool3x3s2_bwd_kernel(const T* __restrict__ dy, int ldy, const uint8_t* __restrict__ idx,
                                        T* __restrict__ dx, int ldx, int N, int H, int W, int C, int Ho,
                                        int Wo) {
  GDL_PDL_ENTRY();
  const int cv = C / 8;
  const long long total = (long long)N * H * W * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int c0 = (int)(i - pix * cv) * 8;
    const int w = (int)(pix % W);
    const long long t = pix / W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < 3; ++r) {
      const int hn = h + 1 - r;
      if (hn < 0 || (hn & 1)) continue;
      const int ho = hn >> 1;
      if (ho >= Ho) continue;
      for (int s = 0; s < 3; ++s) {
        const int wn = w + 1 - s;
        if (wn < 0 || (wn & 1)) continue;
        const int wo = wn >> 1;
        if (wo >= Wo) continue;
        const long long opix = ((long long)n * Ho + ho) * Wo + wo;
        const uint2 p = *reinterpret_cast<const uint2*>(idx + opix * C + c0);
        float f[8];
        load8(dy + opix * ldy + c0, f);
        const int tap = r * 3 + s;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t word = j < 4 ? p.x : p.y;
          const int sel = (word >> ((j & 3) * 8)) & 0xff;
          if (sel == tap) acc[j] += f[j];
        }
      }
    }
    store8(dx + pix * ldx + c0, acc);
  }
}

// ------------------------------------------------------------------------------------------
// nn.Dropout2d in training mode (segformer_mlp.py:73,129; fcn_head.py:69-83): whole channels of a sample are zeroed,
// the rest scaled by 1 / (1 - p).  The draw m[n][c] (0 or 1 / (1 - p)) is made by the caller; y = x * m[n][c] is
// also its own backward (dx = dy * m).
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void scale_nc_kernel(const T* __restrict__ x, long long ldx, const float* __restrict__ m, T* __restrict__ y,
                                long long ldy, long long N, long long HW, int C) {
  GDL_PDL_ENTRY();
  const int cv = C / 8;
  const long long total = N * HW * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / cv;
    const int c = (int)(i - row * cv) * 8;
    const long long n = row / HW;
    float f[8];
    load8(x + row * ldx + c, f);
    const float* mm = m + n * C + c;
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= mm[j];
    store8(y + row * ldy + c, f);
  }
}

}  // namespace gdl

using namespace gdl;

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" int gdl_normalize_to_nhwc(const void* x, int in_kind, void* y, int out_dtype, long long N,
                                     long long H, long long W, int C, int ld, const float* mean,
                                     const float* stdv, float image_max, void* stream) {
  GDL_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && ld >= C, GDL_ERR_INVALID, "normalize: bad args");
  GDL_REQUIRE((mean == nullptr) == (stdv == nullptr), GDL_ERR_INVALID, "normalize: mean and std go together");
  const long long npix = N * H * W;
  const int blocks = ew_blocks(npix * ld, 256, 16);
  cudaStream_t st = (cudaStream_t)stream;
  // in_kind: 0 = uint8 NHWC, 1 = f32 NHWC, 2 = f32 NCHW, 3 = uint8 NCHW
  GDL_DISPATCH_16(out_dtype, {
    if (in_kind == 0)
      GDL_LAUNCH((normalize_kernel<T, uint8_t>), blocks, 256, 0, st, (const uint8_t*)x, (T*)y, npix, C, ld, mean, stdv, image_max, 0, H * W);
    else if (in_kind == 1)
      GDL_LAUNCH((normalize_kernel<T, float>), blocks, 256, 0, st, (const float*)x, (T*)y, npix, C, ld, mean, stdv, image_max, 0, H * W);
    else if (in_kind == 2)
      GDL_LAUNCH((normalize_kernel<T, float>), blocks, 256, 0, st, (const float*)x, (T*)y, npix, C, ld, mean, stdv, image_max, 1, H * W);
    else if (in_kind == 3)
      GDL_LAUNCH((normalize_kernel<T, uint8_t>), blocks, 256, 0, st, (const uint8_t*)x, (T*)y, npix, C, ld, mean, stdv, image_max, 1, H * W);
    else {
      set_last_error("normalize: unknown in_kind %d", in_kind);
      return GDL_ERR_INVALID;
    }
  });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_im2col_nhwc(const void* x, void* col, int dtype, int N, int H, int W, int C, int ld,
                               int R, int S, int stride, int pad, int Kpad, void* stream) {
  GDL_REQUIRE(x && col && N > 0 && H > 0 && W > 0 && C > 0 && ld >= C && R > 0 && S > 0 && stride > 0 && pad >= 0,
              GDL_ERR_INVALID, "im2col: bad args");
  GDL_REQUIRE(Kpad >= R * S * C && Kpad % 8 == 0, GDL_ERR_INVALID, "im2col: Kpad %d too small or not a multiple of 8", Kpad);
  const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S) / stride + 1;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (C % 8 == 0) && (ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  GDL_DISPATCH_16(dtype, {
    if (C == 4 && ld % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) & 7) == 0)) {
      const long long total = (long long)N * Ho * Wo * (Kpad / 8);
      GDL_LAUNCH(im2col_c4_kernel<T>, ew_blocks(total, 256, 16), 256, 0, st, (const T*)x, (T*)col, N, H, W, ld, R, S, stride, pad, Ho, Wo, Kpad);
    } else if (vec) {
      const long long total = (long long)N * Ho * Wo * (Kpad / 8);
      GDL_LAUNCH(im2col_vec8_kernel<T>, ew_blocks(total, 256, 16), 256, 0, st, (const T*)x, (T*)col, N, H, W, C, ld, R, S, stride, pad, Ho, Wo, Kpad);
    } else {
      const long long total = (long long)N * Ho * Wo * Kpad;
      GDL_LAUNCH(im2col_scalar_kernel<T>, ew_blocks(total, 256, 32), 256, 0, st, (const T*)x, (T*)col, N, H, W, C, ld, R, S, stride, pad, Ho, Wo, Kpad);
    }
  });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_col2im_nhwc(const void* dcol, void* dx, int dtype, int N, int H, int W, int C, int ld,
                               int R, int S, int stride, int pad, int Kpad, void* stream) {
  GDL_REQUIRE(dcol && dx && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && ld >= C && ld % 8 == 0,
              GDL_ERR_INVALID, "col2im: bad args (C must be a multiple of 8)");
  GDL_REQUIRE(Kpad >= R * S * C && Kpad % 8 == 0, GDL_ERR_INVALID, "col2im: bad Kpad");
  const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S) / stride + 1;
  const long long total = (long long)N * H * W * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  GDL_DISPATCH_16(dtype, {
    GDL_LAUNCH(col2im_vec8_kernel<T>, ew_blocks(total, 256, 16), 256, 0, st, (const T*)dcol, (T*)dx, N, H, W, C, ld, R, S, stride, pad, Ho, Wo, Kpad);
  });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int stats_launch_geometry(int C, int* threads, int* smem_bytes) {
  const int cv = C / 8;
  const int t = 256;
  const int tpr = cv < t ? cv : t;
  const int rpb = t / tpr;
  *threads = t;
  *smem_bytes = rpb * tpr * 8 * 2 * (int)sizeof(float);
  return rpb;
}

extern "C" int gdl_bn_stats(const void* x, int dtype, long long M, int C, int ld, float* sums,
                            const float* pivot, void* stream) {
  GDL_REQUIRE(x && sums && M > 0 && C > 0 && C % 8 == 0 && ld >= C && ld % 8 == 0, GDL_ERR_INVALID,
              "bn_stats: bad args (C=%d must be a multiple of 8)", C);
  cudaStream_t st = (cudaStream_t)stream;
  GDL_CHECK_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)C * sizeof(float), st));
  int threads, smem;
  const int rpb = stats_launch_geometry(C, &threads, &smem);
  long long blocks = (M + rpb * 16 - 1) / (rpb * 16);  // >= 16 rows per thread
  if (blocks > 4 * kNumSMsB200) blocks = 4 * kNumSMsB200;
  if (blocks < 1) blocks = 1;
  const DetWs ws = det_workspace();
  DetCtx det = det_none();
  if (const int g = det_grid(ws, blocks, 2 * C)) {
    blocks = g;
    det = det_ctx(ws, g, 2 * C);
  }
  GDL_DISPATCH_16(dtype, { GDL_LAUNCH(bn_stats_kernel<T>, (int)blocks, threads, smem, st, (const T*)x, M, C, ld, sums, pivot, det); });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_bn_finalize(const float* pivot, const float* sums, long long M, int C, const float* gamma,
                               const float* beta, float eps, float momentum, float* running_mean,
                               float* running_var, float* scale, float* shift, float* save_mean,
                               float* save_invstd, void* stream) {
  GDL_REQUIRE(sums && scale && shift && save_mean && save_invstd && M > 0 && C > 0, GDL_ERR_INVALID,
              "bn_finalize: bad args");
  GDL_LAUNCH(bn_finalize_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream, pivot, sums, M, C, gamma, beta, eps,
                                                                       momentum, running_mean, running_var, scale,
                                                                       shift, save_mean, save_invstd);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_bn_eval_coeffs(int C, const float* gamma, const float* beta, const float* running_mean,
                                  const float* running_var, float eps, float* scale, float* shift, void* stream) {
  GDL_REQUIRE(C > 0 && running_mean && running_var && scale && shift, GDL_ERR_INVALID, "bn_eval_coeffs: bad args");
  GDL_LAUNCH(bn_eval_coeffs_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream, C, gamma, beta, running_mean,
                                                                          running_var, eps, scale, shift);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_bn_apply(const void* x, int ldx, const float* scale, const float* shift, const void* res,
                            int ldr, const float* rscale, const float* rshift, int relu, void* y, int ldy,
                            void* y_up, int ldu, int dtype, int N, int H, int W, int C, void* stream) {
  GDL_REQUIRE(x && scale && shift && (y || y_up) && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0,
              GDL_ERR_INVALID, "bn_apply: bad args (C=%d must be a multiple of 8)", C);
  GDL_REQUIRE(ldx % 8 == 0 && (!y || ldy % 8 == 0) && (!y_up || ldu % 8 == 0) && (!res || ldr % 8 == 0),
              GDL_ERR_INVALID, "bn_apply: strides must be multiples of 8");
  GDL_REQUIRE((rscale == nullptr) == (rshift == nullptr), GDL_ERR_INVALID, "bn_apply: rscale/rshift go together");
  const long long total = (long long)N * H * W * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  GDL_DISPATCH_16(dtype, {
    GDL_LAUNCH(bn_apply_kernel<T>, row_grid(total / (C / 8), C), 256, 0, st, (const T*)x, ldx, scale, shift, (const T*)res, ldr,
                                                                  rscale, rshift, relu, (T*)y, ldy, (T*)y_up, ldu,
                                                                  N, H, W, C);
  });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_grad_gather(int num_src, const void* const* src_ptr, const int* src_ld, const int* src_mode,
                               const void* y, int ldy, const void* x, int ldx, const float* mean,
                               const float* invstd, void* g, int ldg, float* sums, int dtype, int N, int H,
                               int W, int C, void* stream) {
  GDL_REQUIRE(num_src >= 1 && num_src <= GDL_MAX_SRC && src_ptr && src_ld && src_mode, GDL_ERR_INVALID,
              "grad_gather: between 1 and %d sources", GDL_MAX_SRC);
  GDL_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, GDL_ERR_INVALID, "grad_gather: bad shape");
  GDL_REQUIRE(g || sums, GDL_ERR_INVALID, "grad_gather: nothing to produce");
  GDL_REQUIRE(!sums || (x && mean && invstd), GDL_ERR_INVALID, "grad_gather: sums need x, mean, invstd");
  GatherSrcs gs;
  memset(&gs, 0, sizeof(gs));
  gs.n = num_src;
  for (int i = 0; i < num_src; ++i) {
    GDL_REQUIRE(src_ptr[i] && src_ld[i] % 8 == 0, GDL_ERR_INVALID, "grad_gather: bad source %d", i);
    gs.ptr[i] = src_ptr[i];
    gs.ld[i] = src_ld[i];
    gs.mode[i] = src_mode[i];
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (sums) GDL_CHECK_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)C * sizeof(float), st));
  int threads, smem;
  const int rpb = stats_launch_geometry(C, &threads, &smem);
  const long long M = (long long)N * H * W;
  long long blocks = (M + rpb * 4 - 1) / (rpb * 4);
  if (blocks > 8 * kNumSMsB200) blocks = 8 * kNumSMsB200;
  if (blocks < 1) blocks = 1;
  GDL_REQUIRE(M < (1ll << 31), GDL_ERR_UNSUPPORTED, "grad_gather: too many pixels");
  DetCtx det = det_none();
  if (sums) {
    const DetWs ws = det_workspace();
    if (const int gd = det_grid(ws, blocks, 2 * C)) {
      blocks = gd;
      det = det_ctx(ws, gd, 2 * C);
    }
  }
  // source layout: plain sources first, at most one 2x2-pooled (mode 1) source, last -> the lean kernel; anything else -> generic
  int nup = 0;
  for (int i = 0; i < num_src; ++i) nup += src_mode[i] != 0;
  const bool lean = nup == 0 || (nup == 1 && src_mode[num_src - 1] != 0);
#define GDL_GG_LAUNCH2(NSV, NUPV)                                                                                      \
  GDL_DISPATCH_16(dtype, {                                                                                             \
    GDL_LAUNCH((grad_gather_kernel<T, NSV, NUPV>), (int)blocks, threads, smem, st, gs, (const T*)y, ldy, (const T*)x, ldx, mean, \
                                                                         invstd, (T*)g, ldg, sums, N, H, W, C, det);   \
  })
#define GDL_GG_LAUNCH(NSV)                          \
  do {                                              \
    if (!lean) GDL_GG_LAUNCH2(NSV, -1);             \
    else if (nup == 0) GDL_GG_LAUNCH2(NSV, 0);      \
    else GDL_GG_LAUNCH2(NSV, 1);                    \
  } while (0)
  switch (num_src) {
    case 1: GDL_GG_LAUNCH(1); break;
    case 2: GDL_GG_LAUNCH(2); break;
    case 3: GDL_GG_LAUNCH(3); break;
    case 4: GDL_GG_LAUNCH(4); break;
    case 5: GDL_GG_LAUNCH(5); break;
    default: GDL_GG_LAUNCH(6); break;
  }
#undef GDL_GG_LAUNCH2
#undef GDL_GG_LAUNCH
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_bn_bwd_apply(const void* g, int ldg, const void* x, int ldx, const float* mean,
                                const float* invstd, const float* gamma, const float* sums, void* dx, int ldd,
                                float* dgamma, float* dbeta, int accumulate_param_grads, int dtype, long long M,
                                long long count, int C, void* stream) {
  GDL_REQUIRE(g && x && mean && invstd && sums && dx && M > 0 && C > 0 && C % 8 == 0, GDL_ERR_INVALID,
              "bn_bwd_apply: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  GDL_DISPATCH_16(dtype, {
    GDL_LAUNCH(bn_bwd_apply_kernel<T>, row_grid(M, C), 256, 0, st, (const T*)g, ldg, (const T*)x, ldx, mean,
                                                                           invstd, gamma, sums, (T*)dx, ldd, M, count > 0 ? count : M, C);
  });
  GDL_CHECK_CUDA(cudaGetLastError());
  if (dgamma || dbeta) {
    GDL_LAUNCH(bn_param_grads_kernel, (C + 127) / 128, 128, 0, st, sums, C, dgamma, dbeta, accumulate_param_grads);
    GDL_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

extern "C" int gdl_bn_param_grads(const float* sums, int C, float* dgamma, float* dbeta, int accumulate,
                                  void* stream) {
  GDL_REQUIRE(sums && C > 0, GDL_ERR_INVALID, "bn_param_grads: bad args");
  if (dgamma || dbeta) {
    GDL_LAUNCH(bn_param_grads_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream, sums, C, dgamma, dbeta, accumulate);
    GDL_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

extern "C" int gdl_maxpool3x3s2_fwd(const void* x, int ldx, void* y, int ldy, unsigned char* idx, int dtype, int N,
                                    int H, int W, int C, void* stream) {
  GDL_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0,
              GDL_ERR_INVALID, "maxpool: bad args");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)N * Ho * Wo * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  GDL_DISPATCH_16(dtype, {
    GDL_LAUNCH(maxpool3x3s2_fwd_kernel<T>, ew_blocks(total, 256, 16), 256, 0, st, (const T*)x, ldx, (T*)y, ldy, idx, N, H, W, C, Ho, Wo);
  });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_maxpool3x3s2_bwd(const void* dy, int ldy, const unsigned char* idx, void* dx, int ldx, int dtype,
                                    int N, int H, int W, int C, void* stream) {
  GDL_REQUIRE(dy && idx && dx && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, GDL_ERR_INVALID, "maxpool bwd: bad args");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)N * H * W * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  GDL_DISPATCH_16(dtype, {
    GDL_LAUNCH(maxpool3x3s2_bwd_kernel<T>, ew_blocks(total, 256, 16), 256, 0, st, (const T*)dy, ldy, idx, (T*)dx, ldx, N, H, W, C, Ho, Wo);
  });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_dropout2d_apply(const void* x, long long ldx, const float* mask, void* y, long long ldy, int dtype,
                                   long long N, long long HW, int C, void* stream) {
  GDL_REQUIRE(x && mask && y && N > 0 && HW > 0 && C > 0 && C % 8 == 0 && ldx >= C && ldy >= C && ldx % 8 == 0 && ldy % 8 == 0,
              GDL_ERR_INVALID, "dropout2d: bad args (C, ldx, ldy multiples of 8)");
  GDL_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, GDL_ERR_INVALID,
              "dropout2d: 16-byte aligned buffers expected");
  const int blocks = ew_blocks(N * HW * (C / 8), 256, 16);
  GDL_DISPATCH_16(dtype, {
    GDL_LAUNCH(scale_nc_kernel<T>, blocks, 256, 0, (cudaStream_t)stream, (const T*)x, ldx, mask, (T*)y, ldy, N, HW, C);
  });
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}
