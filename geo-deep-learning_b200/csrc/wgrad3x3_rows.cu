// wgrad3x3_rows.cu — weight gradient of a 3x3 / pad-1 / stride-1 conv with a narrow output (Cout = 16 / 32 / 64).
//
// Why: conv_wgrad_kernel (igemm_conv.cu) puts Cout on the MMA's M dimension; with Cout <= 64 at least half of every
// 128-row MMA is empty, each of its units re-reads dY and X once per filter row, and its epilogue is not overlapped
// (run 8/9: 135-430 TFLOP/s on these layers, tensor pipe 38 % active with half of that wasted).
//
// Here  dW^T[cin][tap][cout] = sum_pixels X_tap^T . dY  with
//   A = X   (MN-major, M = 128 = [64 input channels at horizontal shift s | the same 64 channels at shift s+1]),
//   B = dY  (MN-major, N = Cout), K = the 128 pixels of one image-row segment,
// so one MMA chain produces TWO horizontal taps (rows 0-63 / 64-127 of the accumulator) and a second chain the third
// tap: 3 useful half-tiles out of 4 instead of 1 out of 2.  A unit = (one 64-channel slab of one source, one block of
// Hb image rows of one 128-pixel column): it streams the Hb + 2 input rows ONCE (each against the dY rows above, at and
// below it: filter rows 2, 1, 0) and keeps the 3 dY rows it needs in a ring.
//
// Slab-stationary CTAs (round 2): CTA c owns slab c % num_slabs for its whole life and walks the pixel blocks
// c / num_slabs, + ctas_per_slab, ...; its 6 TMEM accumulators (3 filter rows x {taps 0|1, tap 2}) keep accumulating
// across all of them, so there is ONE epilogue per CTA instead of one per unit: coalesced red.global.add into the fp32
// gradient, or — ordered mode (det_reduce.cuh) — a plain store of the CTA's partial gradient, after which
// wgrad_rows_reduce_kernel adds the partials of a slab in CTA order: bit-reproducible.
//
// The two 64-channel MN atoms of the paired A operand are two TMA boxes of the same input row (130 pixels starting at
// w0-1 and at w0) placed kXAtomBytes apart, addressed through the descriptor's leading-dimension byte offset.
#include <stdlib.h>
#include <string.h>

#include "../../include/gdl_b200.h"
#include "common.cuh"
#include "tmap.cuh"
#include "det_reduce.cuh"

namespace gdl {

constexpr int kWrThreads = 192;
constexpr int kWrXAtomBytes = 17 * 1024;  // 130 px x 64 ch x 2 B = 16 640 B, padded to the 1024-B swizzle repeat
constexpr int kWrXStages = 3;             // input-row stages (2 atoms each)
constexpr int kWrDYStages = 4;            // dY rows: 3 in use + 1 in flight

struct WgradRowsKParams {
  struct SlabOffsets {
    int coff[64];
  };
  CUtensorMap tmX[GDL_MAX_SRC];
  CUtensorMap tmDY;
  int num_slabs;
  int slab_src[64], slab_c0[64], slab_coff[64];  // source index, channel offset inside it, offset in the virtual concat
  int Ctot, Cout;
  int Nimg, H, W;
  int tiles_w, Hb, hblocks;
  int nblk;           // pixel blocks (image x column x row block); a unit = (slab, pixel block)
  int ctas_per_slab;  // gridDim.x = num_slabs * ctas_per_slab
  float* partials;    // ordered mode: [gridDim.x][3][3][Cout][64] per-CTA partial gradients (null: red.global.add)
  int dy_stage_bytes, dy_row_bytes;  // 128 px x Cout x 2 B (1024-aligned), Cout x 2 B
  int ab_fmt;
  float* dw;
  long long dw_ld;
};

__global__ void __launch_bounds__(kWrThreads, 1) wgrad3x3_rows_kernel(const __grid_constant__ WgradRowsKParams p) {
  GDL_PDL_ENTRY();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* smem_dy = smem + (size_t)kWrXStages * 2 * kWrXAtomBytes;

  __shared__ __align__(8) uint64_t x_full[kWrXStages];
  __shared__ __align__(8) uint64_t x_empty[kWrXStages];
  __shared__ __align__(8) uint64_t dy_full[kWrDYStages];
  __shared__ __align__(8) uint64_t dy_empty[kWrDYStages];
  __shared__ __align__(8) uint64_t tfull_bar;
  __shared__ __align__(8) uint64_t tempty_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmX[0]);
    tma_prefetch_desc(&p.tmDY);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kWrXStages; ++i) {
        mbar_init(&x_full[i], 1);
        mbar_init(&x_empty[i], 1);
      }
      for (int i = 0; i < kWrDYStages; ++i) {
        mbar_init(&dy_full[i], 1);
        mbar_init(&dy_empty[i], 1);
      }
      mbar_init(&tfull_bar, 1);
      mbar_init(&tempty_bar, 128);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_base_smem, 512u);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  // CTA -> (slab, first pixel block); pixel block t -> (row block fastest, column, image).  CTAs with the same
  // blockIdx.x / num_slabs stream the same dY rows at the same time (shared in L2).
  const int slab = blockIdx.x % p.num_slabs;
  const int blk0 = blockIdx.x / p.num_slabs;
  auto decode = [&](int t, int& img, int& w0, int& h0, int& rows) {
    const int hb = t % p.hblocks;
    t /= p.hblocks;
    const int wt = t % p.tiles_w;
    img = t / p.tiles_w;
    w0 = wt * 128;
    h0 = hb * p.Hb;
    rows = min(p.Hb, p.H - h0);
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int xs = 0, ds = 0;
      uint32_t xph = 0, dph = 0;
      const int src = p.slab_src[slab], c0 = p.slab_c0[slab];
      for (int t = blk0; t < p.nblk; t += p.ctas_per_slab) {
        int img, w0, h0, rows;
        decode(t, img, w0, h0, rows);
        for (int i = h0 - 1; i <= h0 + rows; ++i) {
          if (i + 1 < h0 + rows) {  // dY row i+1 (filter row 0 of input row i) — first use of that row
            mbar_wait(&dy_empty[ds], dph ^ 1);
            mbar_expect_tx(&dy_full[ds], (uint32_t)(128 * p.dy_row_bytes));
            tma_load_4d(smem_dy + (size_t)ds * p.dy_stage_bytes, &p.tmDY, &dy_full[ds], 0, w0, i + 1, img);
            if (++ds == kWrDYStages) {
              ds = 0;
              dph ^= 1;
            }
          }
          mbar_wait(&x_empty[xs], xph ^ 1);
          mbar_expect_tx(&x_full[xs], (uint32_t)(2 * 130 * 64 * 2));
          uint8_t* x_dst = smem + (size_t)xs * 2 * kWrXAtomBytes;
          tma_load_4d(x_dst, &p.tmX[src], &x_full[xs], c0, w0 - 1, i, img);                  // shifts 0 and 2
          tma_load_4d(x_dst + kWrXAtomBytes, &p.tmX[src], &x_full[xs], c0, w0, i, img);      // shift 1 (second atom)
          if (++xs == kWrXStages) {
            xs = 0;
            xph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(128, p.Cout, p.ab_fmt, 1, 1);
      const uint32_t ltA = umma_layout_type(128), ltB = umma_layout_type(p.dy_row_bytes);
      const uint32_t sboA = 8u * 128u, sboB = 8u * (uint32_t)p.dy_row_bytes;      // 8 pixel rows
      const uint32_t kstepA = 16u * 128u, kstepB = 16u * (uint32_t)p.dy_row_bytes;  // 16 pixel rows
      int xs = 0, ds = 0;
      uint32_t xph = 0, dph = 0;
      uint32_t started = 0;     // bit ky: accumulators of filter row ky already hold data (kept across pixel blocks)
      for (int t = blk0; t < p.nblk; t += p.ctas_per_slab) {
        int img, w0, h0, rows;
        decode(t, img, w0, h0, rows);
        const int ds_first = ds;  // ring slot of dY row h0; row h lives in (ds_first + h - h0) % kWrDYStages
        for (int i = h0 - 1; i <= h0 + rows; ++i) {
          if (i + 1 < h0 + rows) {
            mbar_wait(&dy_full[ds], dph);
            if (++ds == kWrDYStages) {
              ds = 0;
              dph ^= 1;
            }
          }
          mbar_wait(&x_full[xs], xph);
          tc_fence_after();
          const uint32_t x_addr = smem_u32(smem + (size_t)xs * 2 * kWrXAtomBytes);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int h = i - ky + 1;  // output row whose gradient meets input row i under filter row ky
            if (h < h0 || h >= h0 + rows) continue;
            const uint32_t b_addr = smem_u32(smem_dy + (size_t)((ds_first + h - h0) % kWrDYStages) * p.dy_stage_bytes);
            const uint32_t d_pair = tmem_base + (uint32_t)(ky * 128);
            const uint32_t d_single = d_pair + 64u;
            const uint32_t fresh = ((started >> ky) & 1u) ^ 1u;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t db = umma_smem_desc(b_addr + kk * kstepB, 128 * p.dy_row_bytes, sboB, ltB);
              // taps kx = 0 | 1: atoms = the row at shift 0 (box starting at w0-1) | at shift 1 (box starting at w0)
              const uint64_t da = umma_smem_desc(x_addr + kk * kstepA, kWrXAtomBytes, sboA, ltA);
              umma_f16(d_pair, da, db, idesc, (uint32_t)!(fresh && kk == 0));
              // tap kx = 2: first box shifted by two pixel rows (256 B); the second atom's rows are never read back
              const uint64_t da2 = umma_smem_desc(x_addr + 256 + kk * kstepA, kWrXAtomBytes, sboA, ltA);
              umma_f16(d_single, da2, db, idesc, (uint32_t)!(fresh && kk == 0));
            }
            started |= 1u << ky;
          }
          umma_commit(&x_empty[xs]);
          if (++xs == kWrXStages) {
            xs = 0;
            xph ^= 1;
          }
          // dY row i-1 was last used by this input row (filter row 2)
          if (i - 1 >= h0) umma_commit(&dy_empty[(ds_first + i - 1 - h0) % kWrDYStages]);
        }
      }
      umma_commit(&tfull_bar);  // all pixel blocks of this CTA accumulated
    }
  } else {
    // ===================== epilogue: thread = one accumulator row = one input channel (two taps) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int cin_row = row & 63;      // channel inside the slab
    const int kx_pair = row >> 6;      // rows 0-63: tap kx = 0, rows 64-127: tap kx = 1
    const long long col0 = p.slab_coff[slab] + cin_row;
    mbar_wait(&tfull_bar, 0);
    tc_fence_after();
    float* part = p.partials ? p.partials + (long long)blockIdx.x * (9ll * p.Cout * 64) : nullptr;
    for (int ky = 0; ky < 3; ++ky) {
      const uint32_t t_pair = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ky * 128);
      for (int sel = 0; sel < 2; ++sel) {  // 0: paired accumulator (taps 0|1), 1: tap 2 (rows 0-63 only)
        const int kx = sel == 0 ? kx_pair : 2;
        float* dst = p.dw + (long long)((ky * 3 + kx) * p.Ctot) + col0;
        for (int j = 0; j < p.Cout; j += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(t_pair + (uint32_t)(sel * 64 + j), v);
          tmem_ld_wait();
          if (sel == 1 && row >= 64) continue;
          // lanes = consecutive input channels: every access below is one coalesced 128-byte request per warp
          if (part != nullptr) {
            float* pd = part + ((long long)(ky * 3 + kx) * p.Cout + j) * 64 + cin_row;
#pragma unroll
            for (int o = 0; o < 16; ++o) pd[o * 64] = __uint_as_float(v[o]);
          } else {
#pragma unroll
            for (int o = 0; o < 16; ++o)
              asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + (long long)(j + o) * p.dw_ld),
                           "f"(__uint_as_float(v[o]))
                           : "memory");
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512u);
  }
}

// ordered mode: dw[cout][(ky,kx)][slab channels] += partial(CTA 0 of the slab) + partial(CTA 1) + ... in CTA order
__global__ void __launch_bounds__(256) wgrad_rows_reduce_kernel(const float* __restrict__ partials, int num_slabs, int cps,
                                                                 int Cout, int Ctot, float* __restrict__ dw, long long dw_ld,
                                                                 const __grid_constant__ WgradRowsKParams::SlabOffsets so) {
  GDL_PDL_ENTRY();
  const long long per_cta = 9ll * Cout * 64;
  const long long total = (long long)num_slabs * per_cta;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int slab = (int)(i / per_cta);
    const long long e = i - (long long)slab * per_cta;
    const int cin = (int)(e & 63);
    const int cout = (int)((e >> 6) % Cout);
    const int tap = (int)((e >> 6) / Cout);
    float a = 0.f;
    for (int j = 0; j < cps; ++j) a += __ldcg(partials + ((long long)j * num_slabs + slab) * per_cta + e);
    dw[(long long)cout * dw_ld + (long long)tap * Ctot + so.coff[slab] + cin] += a;
  }
}

// Returns 1 when the descriptor is a case this kernel covers (then *status holds the launch status), 0 otherwise.
int wgrad3x3_rows_try(const gdl_conv_wgrad_t* d, cudaStream_t stream, int* status) {
  *status = 0;
  if (d->R != 3 || d->S != 3 || d->pad_h != 1 || d->pad_w != 1 || d->batched) return 0;
  if (d->Cout != 16 && d->Cout != 32 && d->Cout != 64) return 0;
  if (d->W < 64 || d->H < 4) return 0;
  WgradRowsKParams p;
  memset(&p, 0, sizeof(p));
  int coff = 0;
  for (int i = 0; i < d->num_src; ++i) {
    if (d->src[i].channels % 64) return 0;
    for (int c = 0; c < d->src[i].channels; c += 64) {
      if (p.num_slabs >= 64) return 0;
      p.slab_src[p.num_slabs] = i;
      p.slab_c0[p.num_slabs] = c;
      p.slab_coff[p.num_slabs] = coff + c;
      ++p.num_slabs;
    }
    coff += d->src[i].channels;
  }
  p.Ctot = coff;
  p.Cout = d->Cout;
  p.Nimg = d->N;
  p.H = d->H;
  p.W = d->W;
  p.tiles_w = (d->W + 127) / 128;
  // rows per unit: long units amortise the halo rows (2 per block) and the epilogue, but keep >= ~4 waves of units
  int sms = 0, dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
    sms = kNumSMsB200;
  const int cps_max = sms / p.num_slabs;
  if (cps_max < 1) return 0;  // more slabs than SMs: the generic kernel
  // rows per unit: long units amortise the 2 halo rows per block; keep >= ~4 pixel blocks per CTA for balance
  int Hb = 32;
  while (Hb > 8 && (long long)d->N * p.tiles_w * ((d->H + Hb - 1) / Hb) < 4ll * cps_max) Hb >>= 1;
  if (Hb > d->H) Hb = d->H;
  p.Hb = Hb;
  p.hblocks = (d->H + Hb - 1) / Hb;
  const long long nblk = (long long)d->N * p.tiles_w * p.hblocks;
  if (nblk >= (1ll << 31)) return 0;
  p.nblk = (int)nblk;
  p.ctas_per_slab = nblk < cps_max ? (int)nblk : cps_max;
  {
    const DetWs ws = det_workspace();
    const long long need = (long long)p.num_slabs * p.ctas_per_slab * 9 * d->Cout * 64;
    p.partials = (ws.ok() && p.ctas_per_slab > 1 && need <= ws.slot_floats) ? ws.slots : nullptr;
  }
  p.dy_row_bytes = d->Cout * 2;
  p.dy_stage_bytes = 128 * p.dy_row_bytes;  // 4 / 8 / 16 KB
  p.ab_fmt = d->dtype == GDL_BF16 ? 1 : 0;
  p.dw = d->dw;
  p.dw_ld = d->dw_ld > 0 ? d->dw_ld : 9ll * p.Ctot;

  for (int i = 0; i < d->num_src; ++i) {
    *status = make_tmap_nhwc(&p.tmX[i], d->src[i].ptr, d->dtype, d->src[i].channels, d->W, d->H, d->N, d->src[i].ld, 64,
                             130, 1, 128);
    if (*status) return 1;
  }
  *status = make_tmap_nhwc(&p.tmDY, d->dy, d->dtype, d->Cout, d->W, d->H, d->N, d->ld_dy, d->Cout, 128, 1,
                           p.dy_row_bytes);
  if (*status) return 1;

  const int smem = kWrXStages * 2 * kWrXAtomBytes + kWrDYStages * p.dy_stage_bytes + 1024;
  static PerDeviceOnce attr_once;
  *status = check_cuda(set_max_dyn_smem_once(attr_once, wgrad3x3_rows_kernel,
                                             kWrXStages * 2 * kWrXAtomBytes + kWrDYStages * 16 * 1024 + 1024),
                       "cudaFuncSetAttribute(wgrad3x3_rows_kernel)");
  if (*status) return 1;
  const int grid = p.num_slabs * p.ctas_per_slab;
  GDL_LAUNCH(wgrad3x3_rows_kernel, grid, kWrThreads, smem, stream, p);
  *status = check_cuda(cudaGetLastError(), "wgrad3x3_rows_kernel launch");
  if (*status == 0 && p.partials != nullptr) {
    WgradRowsKParams::SlabOffsets so;
    memset(&so, 0, sizeof(so));
    for (int i = 0; i < p.num_slabs; ++i) so.coff[i] = p.slab_coff[i];
    const long long total = (long long)p.num_slabs * 9 * d->Cout * 64;
    long long rb = (total + 255) / 256;
    if (rb > 4ll * sms) rb = 4ll * sms;
    GDL_LAUNCH(wgrad_rows_reduce_kernel, (int)rb, 256, 0, stream, p.partials, p.num_slabs, p.ctas_per_slab, d->Cout, p.Ctot, p.dw,
                                                        p.dw_ld, so);
    *status = check_cuda(cudaGetLastError(), "wgrad_rows_reduce_kernel launch");
  }
  return 1;
}

}  // namespace gdl
