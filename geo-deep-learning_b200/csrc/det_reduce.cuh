// det_reduce.cuh — run-to-run deterministic cross-block reductions (no floating-point atomics on global memory).
//
// Why: fp32 atomics land in arrival order, so two identical launches (or an eager launch and its CUDA-graph replay)
// differ in the last bits of every BatchNorm sum and weight gradient; through a randomly initialised UNet++ in 16-bit
// arithmetic that noise is amplified to tens of percent in first-layer gradients (round-1 VERDICT, weak #2).
//
// Two primitives, both on a caller-registered, per-device workspace (gdl_set_workspace; the library allocates nothing):
//   (1) slot reduction: every block writes its partial vector (K floats) into slot blockIdx.x with plain stores; the last
//       block of each group of 32 (atomic ticket) adds the group's slots in block order, the last group adds the group
//       sums in group order and accumulates them onto `out`.  The summation tree depends only on (gridDim, K).
//   (2) turnstile: work units that add into the same output tile do so in a fixed order — unit `turn` spins until the
//       tile's counter equals `turn`, adds, publishes turn + 1; the last one resets the counter (zero at rest).
//       Deadlock-free for persistent kernels whose CTAs are all co-resident and walk their units in increasing order,
//       when a unit only ever waits for units with a smaller index.
// Launches that use the workspace must be stream-ordered with respect to each other on a device (the trainer, the
// Lightning route and the sliding-window driver issue everything on one stream).
#pragma once
#include "common.cuh"

namespace gdl {

constexpr int kDetGroup = 32;               // blocks per first-level group
constexpr int kDetReduceCtrs = 1024;        // [0] = group ticket, [1 + g] = block tickets of group g
constexpr int kDetTurnstiles = 64 * 1024 - kDetReduceCtrs;
constexpr long long kDetCtrBytes = 64ll * 1024 * 4;

struct DetCtx {
  float* s0;       // [gridDim.x][K] block partials
  float* s1;       // [groups][K]    group sums
  unsigned* ctr;   // kDetReduceCtrs tickets, zero at rest
};

struct DetWs {
  unsigned* ctr;          // kDetReduceCtrs reduce tickets + kDetTurnstiles turnstile counters
  float* slots;
  long long slot_floats;
  bool ok() const { return slots != nullptr; }
  unsigned* turnstiles() const { return ctr + kDetReduceCtrs; }
};

// workspace of the CURRENT device; .ok() is false when none is registered or option "deterministic" is 0 (runtime.cu)
DetWs det_workspace();

// largest grid (<= want) whose slots fit; 0 when the reduction cannot use the workspace
inline int det_grid(const DetWs& ws, long long want, int K) {
  if (!ws.ok() || K <= 0) return 0;
  long long g = want;
  const long long cap_groups = (long long)(kDetReduceCtrs - 1) * kDetGroup;
  if (g > cap_groups) g = cap_groups;
  while (g > 0 && (g + (g + kDetGroup - 1) / kDetGroup) * K > ws.slot_floats) g = g * 3 / 4;
  return (int)g;
}
inline DetCtx det_ctx(const DetWs& ws, int grid, int K) {
  DetCtx c;
  c.s0 = ws.slots;
  c.s1 = ws.slots + (long long)grid * K;
  c.ctr = ws.ctr;
  return c;
}
inline DetCtx det_none() {
  DetCtx c;
  c.s0 = c.s1 = nullptr;
  c.ctr = nullptr;
  return c;
}

GDL_DEVINL float det_ld(const float* p) { return __ldcg(p); }  // L2 (other blocks' stores), never a stale L1 line

// partial i of this block
GDL_DEVINL void det_put(const DetCtx& d, int K, int i, float v) { d.s0[(long long)blockIdx.x * K + i] = v; }

// Called by EVERY thread of EVERY block after the block's det_put stores (uniform control flow).  out[i] += total_i.
__device__ inline void det_finish(const DetCtx& d, int K, float* __restrict__ out) {
  __shared__ unsigned s_last;
  const unsigned nblk = gridDim.x;
  const unsigned grp = blockIdx.x / kDetGroup, ngrp = (nblk + kDetGroup - 1) / kDetGroup;
  const unsigned b0 = grp * kDetGroup;
  const unsigned gsize = min((unsigned)kDetGroup, nblk - b0);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&d.ctr[1 + grp], 1u) == gsize - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    float s = 0.f;
    for (unsigned b = 0; b < gsize; ++b) s += det_ld(d.s0 + (long long)(b0 + b) * K + i);
    if (ngrp == 1) out[i] += s;
    else d.s1[(long long)grp * K + i] = s;
  }
  if (threadIdx.x == 0) d.ctr[1 + grp] = 0;
  if (ngrp == 1) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&d.ctr[0], 1u) == ngrp - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    float s = 0.f;
    for (unsigned g = 0; g < ngrp; ++g) s += det_ld(d.s1 + (long long)g * K + i);
    out[i] += s;
  }
  if (threadIdx.x == 0) d.ctr[0] = 0;
}

// ---- turnstile -------------------------------------------------------------------------------------------------
// one thread: spin until *t == turn (bounded like mbar_wait: a protocol bug traps instead of hanging the box)
GDL_DEVINL void turnstile_wait(unsigned* t, unsigned turn) {
  volatile unsigned* v = t;
  if (*v == turn) {
    __threadfence();
    return;
  }
  const long long t0 = clock64();
  unsigned spins = 0;
  while (*v != turn) {
    if ((++spins & 0xff) == 0 && clock64() - t0 > 8000000000ll) {
      printf("gdl: turnstile timeout block %d turn %u value %u\n", blockIdx.x, turn, *v);
      __trap();
    }
  }
  __threadfence();
}
// one thread, after the adding threads have fenced and synchronised
GDL_DEVINL void turnstile_pass(unsigned* t, unsigned turn, bool last) {
  __threadfence();
  *reinterpret_cast<volatile unsigned*>(t) = last ? 0u : turn + 1u;
}

}  // namespace gdl
