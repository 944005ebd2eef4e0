// conv3x3_rows.cu — weight-stationary, row-rolling 3x3 / pad-1 / stride-1 implicit GEMM for narrow outputs.
//
// Why: for Cout <= 64 the per-tile kernel in igemm_conv.cu is bound by L2->SM bandwidth, not by the tensor pipe
// (profiles/: 64-channel layers at 256^2 run at ~490 of ~1230 TFLOP/s).  Per 128-pixel tile and 64-channel k-chunk it
// pulls 3 input rows (3 x 16.6 KB) AND the 9 weight taps (9 x 8 KB = 73 KB): the weights are the larger stream.
//
// Here one CTA owns a column block of G vertically adjacent 128-pixel row tiles (G x BN <= 256 TMEM columns, double
// buffered), and per k-chunk
//   * loads the 9 weight taps ONCE (three filter-row stages of a weight ring) for all G tiles,
//   * streams the G + 2 input rows ONCE (instead of 3 G): input row i feeds the filter rows r = 0,1,2 of the output
//     tiles g = i+1, i, i-1 — up to 9 MMAs chains per loaded row, each into a different TMEM accumulator,
//   * keeps the 3 horizontal taps as row-shifted UMMA descriptors inside the 130-pixel halo tile (as the halo mode of
//     igemm_conv.cu, probe: profiles/r01_probe_shifted_umma_descriptor.json).
// L2->SM bytes per tile and k-chunk: 16.6 KB x (G+2)/G + 73 KB / G  (G = 4: 43 KB instead of 123 KB).
//
// Roles: warp 0 = TMA producer (two rings: input rows, weight filter rows), warp 1 = MMA issuer (one thread),
// warps 2..5 = epilogue (tcgen05.ld -> bias / ReLU / GELU / LayerScale / residual -> NHWC row stores).
#include <stdlib.h>
#include <string.h>

#include "../../include/gdl_b200.h"
#include "common.cuh"
#include "tmap.cuh"
#include "det_reduce.cuh"

namespace gdl {

constexpr int kRowsThreads = 192;
constexpr int kRowsBK = 64;                      // channels per k-chunk (128-byte swizzled rows)
constexpr int kRowsAStage = 17 * 1024;           // 130 px x 64 ch x 2 B = 16 640 B, padded to the 1024-B swizzle repeat
constexpr int kRowsMaxAStages = 6;
constexpr int kRowsMaxBStages = 6;
constexpr int kRowsSmemBudget = 224 * 1024;     // + ~300 B static < 227 KB

struct ConvRowsKParams {
  CUtensorMap tmA[GDL_MAX_SRC];
  CUtensorMap tmB;
  CUtensorMap tmO;       // output tile store (TMA) when tma_store != 0
  int tma_store;         // 16-bit output, 16-byte aligned rows, no residual: the epilogue stages tiles in smem
  int o_stage_bytes;     // 128 px x BN x 2 B
  int o_swz_mask;        // 16-byte-chunk XOR mask of the staging tile's hardware swizzle (7 / 3 / 1 / 0)
  int num_src;
  int src_chunks[GDL_MAX_SRC];
  int src_coff[GDL_MAX_SRC];
  int Ctot;
  int Nimg, Ho, Wo;
  int tiles_w, hblocks, G;
  int Cout, BN, n_tiles;
  long long num_jobs;
  int a_stages, b_stages, b_tap_bytes, b_stage_bytes;
  int ab_fmt;
  void* out;
  int out_dtype;
  long long ldo;
  int vec_ok;
  const float* bias;
  int relu;
  const float* oscale;
  const void* residual;
  int res_dtype;
  long long ldr;
  // BatchNorm statistics of the rounded output, read back from the staged 64-channel tiles (tma_store, BN == 64)
  int bn_on, bn_smem_off;  // dynamic-smem offset of the 4 x [2][64] float accumulators
  float* bn_sums;
  const float* bn_pivot;
  DetCtx bn_det;
};

struct RowsJob {
  int n0, img, w0, h0;
};

GDL_DEVINL RowsJob rows_job(const ConvRowsKParams& p, long long job) {
  // consecutive jobs walk down an image column block by block, then across, then images, then Cout tiles
  RowsJob j;
  const int hb = (int)(job % p.hblocks);
  long long t = job / p.hblocks;
  const int wt = (int)(t % p.tiles_w);
  t /= p.tiles_w;
  j.img = (int)(t % p.Nimg);
  j.n0 = (int)(t / p.Nimg) * p.BN;
  j.w0 = wt * 128;
  j.h0 = hb * p.G;
  return j;
}

__global__ void __launch_bounds__(kRowsThreads, 1) conv3x3_rows_kernel(const __grid_constant__ ConvRowsKParams p) {
  GDL_PDL_ENTRY();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* smem_b = smem + (size_t)p.a_stages * kRowsAStage;
  uint8_t* smem_o = smem_b + (size_t)p.b_stages * p.b_stage_bytes;  // 2 output staging tiles (1024-B aligned)

  __shared__ __align__(8) uint64_t a_full[kRowsMaxAStages];
  __shared__ __align__(8) uint64_t a_empty[kRowsMaxAStages];
  __shared__ __align__(8) uint64_t b_full[kRowsMaxBStages];
  __shared__ __align__(8) uint64_t b_empty[kRowsMaxBStages];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int G = p.G;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_src; ++s) tma_prefetch_desc(&p.tmA[s]);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < p.a_stages; ++i) {
        mbar_init(&a_full[i], 1);
        mbar_init(&a_empty[i], 1);
      }
      for (int i = 0; i < p.b_stages; ++i) {
        mbar_init(&b_full[i], 1);
        mbar_init(&b_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull_bar[i], 1);
        mbar_init(&tempty_bar[i], 128);
      }
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (long long job = blockIdx.x; job < p.num_jobs; job += gridDim.x) {
        const RowsJob j = rows_job(p, job);
        for (int src = 0; src < p.num_src; ++src) {
          for (int ch = 0; ch < p.src_chunks[src]; ++ch) {
            const int kcol = p.src_coff[src] + ch * kRowsBK;
            for (int i = -1; i <= G; ++i) {
              if (i <= 1) {  // weights of filter row r = i + 1 (3 taps), first needed by input row i
                const int r = i + 1;
                mbar_wait(&b_empty[bs], bph ^ 1);
                uint8_t* b_dst = smem_b + (size_t)bs * p.b_stage_bytes;
                mbar_expect_tx(&b_full[bs], (uint32_t)(3 * p.BN * kRowsBK * 2));
                for (int s = 0; s < 3; ++s)
                  tma_load_2d(b_dst + s * p.b_tap_bytes, &p.tmB, &b_full[bs], (r * 3 + s) * p.Ctot + kcol, j.n0);
                if (++bs == p.b_stages) {
                  bs = 0;
                  bph ^= 1;
                }
              }
              mbar_wait(&a_empty[as], aph ^ 1);
              mbar_expect_tx(&a_full[as], (uint32_t)(130 * kRowsBK * 2));
              tma_load_4d(smem + (size_t)as * kRowsAStage, &p.tmA[src], &a_full[as], ch * kRowsBK, j.w0 - 1, j.h0 + i,
                          j.img);
              if (++as == p.a_stages) {
                as = 0;
                aph ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(128, p.BN, p.ab_fmt, 0, 0);
      const uint32_t lt = umma_layout_type(kRowsBK * 2);
      const uint32_t sbo = 8u * kRowsBK * 2u;
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      for (long long job = blockIdx.x; job < p.num_jobs; job += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)(buf * 256);
        bool first_chunk = true;
        for (int src = 0; src < p.num_src; ++src) {
          for (int ch = 0; ch < p.src_chunks[src]; ++ch) {
            uint32_t b_addr[3] = {0, 0, 0};
            int b_slot[3] = {0, 0, 0};
            for (int i = -1; i <= G; ++i) {
              if (i <= 1) {
                const int r = i + 1;
                mbar_wait(&b_full[bs], bph);
                b_slot[r] = bs;
                b_addr[r] = smem_u32(smem_b + (size_t)bs * p.b_stage_bytes);
                if (++bs == p.b_stages) {
                  bs = 0;
                  bph ^= 1;
                }
              }
              mbar_wait(&a_full[as], aph);
              tc_fence_after();
              const uint32_t a_addr = smem_u32(smem + (size_t)as * kRowsAStage);
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                const int g = i - r + 1;  // output row tile fed by (input row i, filter row r)
                if (g < 0 || g >= G) continue;
                const uint32_t d_tmem = d_base + (uint32_t)(g * p.BN);
#pragma unroll
                for (int s3 = 0; s3 < 3; ++s3) {
                  const uint32_t a_s = a_addr + s3 * kRowsBK * 2;  // start shifted by s3 pixels (rows of 128 B)
                  const uint32_t b_s = b_addr[r] + s3 * p.b_tap_bytes;
#pragma unroll
                  for (int kk = 0; kk < kRowsBK / 16; ++kk) {
                    const uint64_t da = umma_smem_desc(a_s + kk * 32, 16, sbo, lt);
                    const uint64_t db = umma_smem_desc(b_s + kk * 32, 16, sbo, lt);
                    umma_f16(d_tmem, da, db, idesc, (uint32_t)(!(first_chunk && r == 0 && s3 == 0 && kk == 0)));
                  }
                }
              }
              umma_commit(&a_empty[as]);
              if (++as == p.a_stages) {
                as = 0;
                aph ^= 1;
              }
              // a filter row's weights are last used by input row G - 2 + r
              if (i >= G - 2) umma_commit(&b_empty[b_slot[i - (G - 2)]]);
            }
            first_chunk = false;
          }
        }
        umma_commit(&tfull_bar[buf]);
      }
    }
  } else {
    // ===================== epilogue: every thread owns one accumulator row (pixel) =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    if (p.tma_store) {
      // Row stores from registers cost ~4 SM-cycles per 16-byte store (every lane hits another line): 16 k cycles for
      // the 4 x 16 KB tiles of a job, 3.5x its MMA time (run 9: tensor pipe 28 % active).  Instead each thread writes
      // its pixel row into a swizzled smem tile and one thread hands the tile to the TMA store engine.
      const bool issuer = (warp == 2 && lane == 0);
      const int pitch = p.BN * 2;
      int st = 0;
      int it = 0;
      float* bn_all = reinterpret_cast<float*>(smem + p.bn_smem_off);  // [4 warps][2][64]
      float* bn_acc = bn_all + q * 128;
      if (p.bn_on) {
        for (int i = lane; i < 128; i += 32) bn_acc[i] = 0.f;
        __syncwarp();
      }
      for (long long job = blockIdx.x; job < p.num_jobs; job += gridDim.x, ++it) {
        const int buf = it & 1;
        const RowsJob j = rows_job(p, job);
        mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
        tc_fence_after();
        for (int g = 0; g < G; ++g) {
          if (issuer) bulk_wait_group_read<1>();  // the store issued two tiles ago no longer reads staging[st]
          named_bar_sync(1, 128);
          uint8_t* stg = smem_o + (size_t)st * p.o_stage_bytes;
          const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + g * p.BN);
          for (int cb = 0; cb < p.BN; cb += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(t_addr + cb, v);
            tmem_ld_wait();
            const int c0 = j.n0 + cb;
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
            if (p.bias != nullptr) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c0 + i < p.Cout) f[i] += __ldg(p.bias + c0 + i);
            }
            if (p.oscale != nullptr) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c0 + i < p.Cout) f[i] *= __ldg(p.oscale + c0 + i);
            }
            if (p.relu == 1) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
            } else if (p.relu == 2) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = 0.5f * f[i] * (1.f + erff(f[i] * 0.70710678118654752f));
            }
            uint4 lo, hi;
            if (p.out_dtype == GDL_BF16) {
              lo = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                              pack_bf16x2(f[6], f[7]));
              hi = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]),
                              pack_bf16x2(f[14], f[15]));
            } else {
              lo = make_uint4(pack_f16x2(f[0], f[1]), pack_f16x2(f[2], f[3]), pack_f16x2(f[4], f[5]),
                              pack_f16x2(f[6], f[7]));
              hi = make_uint4(pack_f16x2(f[8], f[9]), pack_f16x2(f[10], f[11]), pack_f16x2(f[12], f[13]),
                              pack_f16x2(f[14], f[15]));
            }
            uint32_t off = (uint32_t)(row * pitch + cb * 2);
            uint32_t o0 = off ^ (((off >> 7) & (uint32_t)p.o_swz_mask) << 4);
            off += 16;
            uint32_t o1 = off ^ (((off >> 7) & (uint32_t)p.o_swz_mask) << 4);
            *reinterpret_cast<uint4*>(stg + o0) = lo;
            *reinterpret_cast<uint4*>(stg + o1) = hi;
          }
          if (g == G - 1) {  // every TMEM read of this job is done: hand the accumulators back to the MMA warp
            tc_fence_before();
            mbar_arrive(&tempty_bar[buf]);
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (issuer) {
            tma_store_4d(&p.tmO, stg, j.n0, j.w0, j.h0 + g, j.img);
            bulk_commit_group();
          }
          if (p.bn_on && 2 * lane < p.Cout) {
            // the staged tile (128 pixels of one image row x 64 channels) is complete: lane = one channel pair (a 32-bit
            // word per pixel, conflict free), warp q = pixels q, q+4, ...; pixels beyond the image width are skipped
            const int c = 2 * lane;
            const float pv0 = p.bn_pivot ? __ldg(p.bn_pivot + c) : 0.f;
            const float pv1 = (p.bn_pivot && c + 1 < p.Cout) ? __ldg(p.bn_pivot + c + 1) : 0.f;
            float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
            const bool full = j.w0 + 128 <= p.Wo;
            const uint32_t xo[2] = {(uint32_t)(lane * 4) ^ ((uint32_t)q << 4), (uint32_t)(lane * 4) ^ ((uint32_t)(q + 4) << 4)};
            const bool is_bf16 = p.out_dtype == GDL_BF16;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
              const int r = q + 4 * k;  // 128-byte rows, chunk XOR mask r & 7: two masks per warp
              const uint32_t u = *reinterpret_cast<const uint32_t*>(stg + r * 128 + xo[k & 1]);
              float x0, x1;
              if (is_bf16) {
                x0 = bf16_lo(u);
                x1 = bf16_hi(u);
              } else {
                const __half2 h2 = *reinterpret_cast<const __half2*>(&u);
                x0 = __low2float(h2);
                x1 = __high2float(h2);
              }
              if (full || j.w0 + r < p.Wo) {
                x0 -= pv0;
                x1 -= pv1;
                s1a += x0;
                s1b += x1;
                s2a = fmaf(x0, x0, s2a);
                s2b = fmaf(x1, x1, s2b);
              }
            }
            bn_acc[c] += s1a;
            bn_acc[c + 1] += s1b;
            bn_acc[64 + c] += s2a;
            bn_acc[64 + c + 1] += s2b;
          }
          st ^= 1;
        }
      }
      if (issuer) bulk_wait_group<0>();
      if (p.bn_on) {
        named_bar_sync(1, 128);
        const int i = (int)threadIdx.x - 64;  // 0..127: [2][64]
        const int c = i & 63;
        if (c < p.Cout) {
          const float v = ((bn_all[i] + bn_all[128 + i]) + bn_all[256 + i]) + bn_all[384 + i];
          det_put(p.bn_det, 2 * p.Cout, i < 64 ? c : p.Cout + c, v);
        }
      }
    } else {
    int it = 0;
    for (long long job = blockIdx.x; job < p.num_jobs; job += gridDim.x, ++it) {
      const int buf = it & 1;
      const RowsJob j = rows_job(p, job);
      const int w = j.w0 + row;
      mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
      tc_fence_after();
      for (int g = 0; g < G; ++g) {
        const int h = j.h0 + g;
        const bool valid = (h < p.Ho) && (w < p.Wo);
        const long long pix = ((long long)j.img * p.Ho + h) * p.Wo + w;
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + g * p.BN);
        for (int cb = 0; cb < p.BN; cb += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(t_addr + cb, v);
          tmem_ld_wait();
          const int c0 = j.n0 + cb;
          const int nvalid = min(16, p.Cout - c0);
          if (!valid || nvalid <= 0) continue;
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < nvalid) f[i] += __ldg(p.bias + c0 + i);
          }
          if (p.oscale != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < nvalid) f[i] *= __ldg(p.oscale + c0 + i);
          }
          if (p.residual != nullptr) {
            const long long roff = pix * p.ldr + c0;
            if (p.res_dtype == GDL_F32) {
              const float* rp = reinterpret_cast<const float*>(p.residual) + roff;
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (i < nvalid) f[i] += rp[i];
            } else if (p.res_dtype == GDL_BF16) {
              const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + roff;
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (i < nvalid) f[i] += __bfloat162float(rp[i]);
            } else {
              const __half* rp = reinterpret_cast<const __half*>(p.residual) + roff;
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (i < nvalid) f[i] += __half2float(rp[i]);
            }
          }
          if (p.relu == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
          } else if (p.relu == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = 0.5f * f[i] * (1.f + erff(f[i] * 0.70710678118654752f));
          }
          const long long off = pix * p.ldo + c0;
          if (p.out_dtype == GDL_F32)
            store_row16<float>(reinterpret_cast<float*>(p.out) + off, f, nvalid, p.vec_ok);
          else if (p.out_dtype == GDL_BF16)
            store_row16<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(p.out) + off, f, nvalid, p.vec_ok);
          else
            store_row16<__half>(reinterpret_cast<__half*>(p.out) + off, f, nvalid, p.vec_ok);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[buf]);
    }
    }  // tma_store
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512u);
  }
  if (p.bn_on) det_finish(p.bn_det, 2 * p.Cout, p.bn_sums);
}

static int rows_sm_count();

static int rows_sm_count() { return device_sm_count(); }

// Returns 1 when the descriptor is a case this kernel covers (then *status holds the launch status), 0 otherwise.
// Called by gdl_conv2d_nhwc_fwd after it validated the descriptor.
// dry != nullptr: launch nothing, report in *dry whether the BatchNorm statistics would come out of the epilogue
int conv3x3_rows_try(const gdl_conv_fwd_t* d, cudaStream_t stream, int* status, int* dry) {
  *status = 0;
  if (d->R != 3 || d->S != 3 || d->pad_h != 1 || d->pad_w != 1 || d->w_mn_major || d->w_rows_per_img) return 0;
  if (d->Cout > 64 || d->W < 64) return 0;
  int Ctot = 0;
  for (int i = 0; i < d->num_src; ++i) {
    if (d->src[i].channels % kRowsBK) return 0;
    Ctot += d->src[i].channels;
  }
  const int BN = (d->Cout + 15) / 16 * 16;
  int G = 256 / BN;
  if (G > 8) G = 8;
  while (G >= 2 && d->H % G) G >>= 1;
  if (G < 2) return 0;

  ConvRowsKParams p;
  memset(&p, 0, sizeof(p));
  p.num_src = d->num_src;
  p.Ctot = Ctot;
  p.Nimg = d->N;
  p.Ho = d->H;
  p.Wo = d->W;
  p.tiles_w = (d->W + 127) / 128;
  p.hblocks = d->H / G;
  p.G = G;
  p.Cout = d->Cout;
  p.BN = BN;
  p.n_tiles = 1;
  p.num_jobs = (long long)p.n_tiles * d->N * p.tiles_w * p.hblocks;
  p.b_tap_bytes = BN * kRowsBK * 2;              // 2 / 4 / 6 / 8 KB: multiples of the 1024-B swizzle repeat
  p.b_stage_bytes = 3 * p.b_tap_bytes;
  p.ab_fmt = d->dtype == GDL_BF16 ? 1 : 0;
  p.out = d->out;
  p.out_dtype = d->out_dtype;
  p.ldo = d->ldo;
  const int esz = d->out_dtype == GDL_F32 ? 4 : 2;
  p.vec_ok = ((reinterpret_cast<uintptr_t>(d->out) & 15) == 0) && ((d->ldo * esz) % 16 == 0);
  static int opt_tma_store = -1;
  if (opt_tma_store < 0) {
    const char* e = getenv("GDL_ROWS_TMA_STORE");
    opt_tma_store = e ? atoi(e) : 1;
  }
  p.tma_store = opt_tma_store && d->out_dtype != GDL_F32 && p.vec_ok && d->residual == nullptr && d->Cout % 8 == 0;
  p.o_stage_bytes = p.tma_store ? 128 * BN * 2 : 0;  // 4 / 8 / 12 / 16 KB
  p.o_swz_mask = BN == 64 ? 7 : (BN == 32 ? 3 : (BN == 16 ? 1 : 0));
  // fused BatchNorm statistics: needs the staged 64-channel tile and a workspace slot per CTA (decided before the
  // pipeline depth: the 4 x [2][64] float accumulators come out of the dynamic shared-memory budget)
  const int sms_bn = rows_sm_count();
  const int grid_bn = p.num_jobs < sms_bn ? (int)p.num_jobs : sms_bn;
  const DetWs ws = det_workspace();
  const bool bn_fused = d->bn_sums != nullptr && p.tma_store && BN == 64 && det_grid(ws, grid_bn, 2 * d->Cout) == grid_bn;
  const int fixed = 1024 + 2 * p.o_stage_bytes + (bn_fused ? 2048 : 0);
  p.a_stages = 4;
  if (dry != nullptr) {
    const int bst = (kRowsSmemBudget - fixed - p.a_stages * kRowsAStage) / (3 * BN * kRowsBK * 2);
    if (bst < 4) return 0;  // (the same test as below: the generic kernel takes this shape)
    *dry = bn_fused ? 1 : 0;
    return 1;
  }
  p.b_stages = (kRowsSmemBudget - fixed - p.a_stages * kRowsAStage) / p.b_stage_bytes;
  if (p.b_stages > kRowsMaxBStages) p.b_stages = kRowsMaxBStages;
  if (p.b_stages < 4) return 0;
  {
    const int left = kRowsSmemBudget - fixed - p.b_stages * p.b_stage_bytes;
    p.a_stages = left / kRowsAStage;
    if (p.a_stages > kRowsMaxAStages) p.a_stages = kRowsMaxAStages;
  }
  p.bias = d->bias;
  p.relu = d->relu;
  p.oscale = d->oscale;
  p.residual = d->residual;
  p.res_dtype = d->res_dtype;
  p.ldr = d->ldr;

  int coff = 0;
  for (int i = 0; i < d->num_src; ++i) {
    p.src_chunks[i] = d->src[i].channels / kRowsBK;
    p.src_coff[i] = coff;
    coff += d->src[i].channels;
    *status = make_tmap_nhwc(&p.tmA[i], d->src[i].ptr, d->dtype, d->src[i].channels, d->W, d->H, d->N, d->src[i].ld,
                             kRowsBK, 130, 1, kRowsBK * 2);
    if (*status) return 1;
  }
  const long long Ktot = 9ll * Ctot;
  const long long w_ld = d->w_ld > 0 ? d->w_ld : Ktot;
  const long long w_rows = d->w_rows > 0 ? d->w_rows : d->Cout;
  *status = make_tmap_2d(&p.tmB, d->weight, d->dtype, Ktot, w_rows, w_ld, kRowsBK, BN, kRowsBK * 2);
  if (*status) return 1;

  if (p.tma_store) {
    const int swz = BN == 64 ? 128 : (BN == 32 ? 64 : (BN == 16 ? 32 : 0));
    *status = make_tmap_nhwc(&p.tmO, d->out, d->out_dtype, d->Cout, d->W, d->H, d->N, d->ldo, BN, 128, 1, swz);
    if (*status) return 1;
  }
  int smem = p.a_stages * kRowsAStage + p.b_stages * p.b_stage_bytes + 2 * p.o_stage_bytes + 1024;
  if (bn_fused) {
    p.bn_on = 1;
    p.bn_smem_off = p.a_stages * kRowsAStage + p.b_stages * p.b_stage_bytes + 2 * p.o_stage_bytes;  // after the staging tiles
    smem += 2048;
    p.bn_sums = d->bn_sums;
    p.bn_pivot = d->bn_pivot;
    p.bn_det = det_ctx(ws, grid_bn, 2 * d->Cout);
    *status = check_cuda(cudaMemsetAsync(d->bn_sums, 0, 2 * (size_t)d->Cout * sizeof(float), stream), "cudaMemsetAsync(bn_sums)");
    if (*status) return 1;
  }
  static PerDeviceOnce attr_once;
  *status = check_cuda(set_max_dyn_smem_once(attr_once, conv3x3_rows_kernel, kRowsSmemBudget + 1024),
                       "cudaFuncSetAttribute(conv3x3_rows_kernel)");
  if (*status) return 1;
  const int sms = rows_sm_count();
  const int grid = p.num_jobs < sms ? (int)p.num_jobs : sms;
  GDL_LAUNCH(conv3x3_rows_kernel, grid, kRowsThreads, smem, stream, p);
  *status = check_cuda(cudaGetLastError(), "conv3x3_rows_kernel launch");
  if (*status == 0 && d->bn_sums != nullptr && !bn_fused)  // the statistics kernel on the stored output
    *status = gdl_bn_stats(d->out, d->out_dtype, (long long)d->N * d->H * d->W, d->Cout, d->ldo, d->bn_sums, d->bn_pivot,
                           (void*)stream);
  return 1;
}

}  // namespace gdl
