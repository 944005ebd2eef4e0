// p2p_exchange.cu — SyncBatchNorm statistics exchanged over NVLink peer memory in ONE small kernel per layer.
//
// Reference behaviour: every shipped YAML sets `sync_batchnorm: true` (configs/segformer_config_RGB.yaml:6-14), i.e. torch
// SyncBatchNorm all-gathers / all-reduces the per-channel sums of EVERY BatchNorm layer in forward and in backward:
// ~150 blocking 2C-float collectives per UNet++-ResNet50 step.  Through NCCL each costs ~25 us at 8 GPUs even inside a
// captured CUDA graph (measured round 2: step 85.1 ms at N=1 -> 90.1 ms at N=8, most of it these collectives).
//
// Here every rank owns a symmetric exchange buffer that all peers have mapped (torch.distributed._symmetric_memory:
// CUDA VMM + NVLink / NVSwitch peer access): [2 slots][slot_floats] floats followed by [2 slots][world] 32-bit flags.
// One block per call:
//   1. seq = ++(*counter) (device-side call counter: nothing call-dependent in kernel arguments -> CUDA-graph safe);
//      slot = seq & 1; copy the local sums into the own slot; fence.sys
//   2. thread r < world: store seq into flag[slot][my rank] of PEER r (release, system scope)
//   3. thread r < world: spin until the own flag[slot][r] >= seq (acquire) — bounded, traps instead of hanging the box
//   4. every thread: sums[i] = slot_0[i] + slot_1[i] + ... in RANK order, read straight from the peers' buffers
// All ranks add in the same order: the result is bit-identical everywhere and run to run.
// Slot reuse is safe with two slots: a rank can complete call k+1 only after every peer has signalled k+1, which a peer does
// after finishing its reads of call k (stream order) — so nobody is more than one call ahead of the slowest reader.
#include "../../include/gdl_b200.h"
#include "common.cuh"

namespace gdl {

constexpr int kP2PMaxWorld = 16;

struct P2PPeers {
  float* buf[kP2PMaxWorld];       // rank r's exchange buffer (mapped in this process)
  unsigned* flags[kP2PMaxWorld];  // rank r's flag array [2][world]
};

GDL_DEVINL void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
GDL_DEVINL unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
GDL_DEVINL float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
GDL_DEVINL float4 ld_relaxed_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(256) p2p_allreduce_kernel(float* __restrict__ sums, int n, const __grid_constant__ P2PPeers peers,
                                                             unsigned* __restrict__ counter, int rank, int world, int slot_floats) {
  GDL_PDL_ENTRY();
  __shared__ unsigned s_seq;
  if (threadIdx.x == 0) {
    s_seq = *counter + 1u;
    *counter = s_seq;
  }
  __syncthreads();
  const unsigned seq = s_seq;
  const int slot = (int)(seq & 1u);
  float* mine = peers.buf[rank] + (long long)slot * slot_floats;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mine[i] = sums[i];
  __threadfence_system();  // the own slot is written before anybody is told so
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int r = threadIdx.x;
    st_release_sys(peers.flags[r] + slot * world + rank, seq);  // "rank's data of call seq is in place", told to peer r
    const unsigned* f = peers.flags[rank] + slot * world + r;
    const long long t0 = clock64();
    unsigned spins = 0;
    while ((int)(ld_acquire_sys(f) - seq) < 0) {
      if ((++spins & 0x3ff) == 0 && clock64() - t0 > 20000000000ll) {  // ~10 s: a peer never arrived
        printf("gdl: p2p exchange timeout rank %d waiting for rank %d (call %u)\n", rank, r, seq);
        __trap();
      }
    }
  }
  __syncthreads();
  // every peer's values of one 4-float group are requested before the first is used: ONE NVLink round trip per group
  // (a dependent `a += load` chain serialised world x n/256 remote loads: 43 us per 4096-float exchange at N = 2, run 11)
  const long long base = (long long)slot * slot_floats;
  const int n4 = n >> 2;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    float4 v[kP2PMaxWorld];
#pragma unroll
    for (int r = 0; r < kP2PMaxWorld; ++r)
      if (r < world) v[r] = ld_relaxed_sys_v4(peers.buf[r] + base + 4 * i);
    float4 a = v[0];
#pragma unroll
    for (int r = 1; r < kP2PMaxWorld; ++r)
      if (r < world) {
        a.x += v[r].x;
        a.y += v[r].y;
        a.z += v[r].z;
        a.w += v[r].w;
      }
    *reinterpret_cast<float4*>(sums + 4 * i) = a;
  }
  for (int i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) {
    float a = 0.f;
    for (int r = 0; r < world; ++r) a += ld_relaxed_sys(peers.buf[r] + base + i);
    sums[i] = a;
  }
}

}  // namespace gdl

using namespace gdl;

extern "C" int gdl_p2p_allreduce_sums(float* sums, int n, const void* const* peer_bufs_host, int rank, int world,
                                      int slot_floats, unsigned* counter, void* stream) {
  GDL_REQUIRE(sums && peer_bufs_host && counter && n > 0 && world >= 1 && world <= kP2PMaxWorld && rank >= 0 && rank < world,
              GDL_ERR_INVALID, "p2p_allreduce_sums: bad args (world <= %d)", kP2PMaxWorld);
  GDL_REQUIRE((reinterpret_cast<uintptr_t>(sums) & 15) == 0, GDL_ERR_INVALID, "p2p_allreduce_sums: sums must be 16-byte aligned");
  GDL_REQUIRE(n <= slot_floats && slot_floats % 4 == 0, GDL_ERR_INVALID, "p2p_allreduce_sums: %d values exceed the %d-float slots", n,
              slot_floats);
  P2PPeers peers;
  for (int r = 0; r < world; ++r) {
    GDL_REQUIRE(peer_bufs_host[r] != nullptr, GDL_ERR_INVALID, "p2p_allreduce_sums: peer %d has no buffer", r);
    peers.buf[r] = reinterpret_cast<float*>(const_cast<void*>(peer_bufs_host[r]));
    peers.flags[r] = reinterpret_cast<unsigned*>(peers.buf[r] + 2ll * slot_floats);
  }
  for (int r = world; r < kP2PMaxWorld; ++r) {
    peers.buf[r] = nullptr;
    peers.flags[r] = nullptr;
  }
  GDL_LAUNCH(p2p_allreduce_kernel, 1, 256, 0, (cudaStream_t)stream, sums, n, peers, counter, rank, world, slot_floats);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" long long gdl_p2p_exchange_bytes(int world, int slot_floats) {
  return 2ll * slot_floats * 4 + 2ll * world * 4;
}
