// hostemu_tc.h — TEST INFRASTRUCTURE: a functional model of the sm_100a features the tensor-core kernels use, behind the
// wrapper names of csrc/common.cuh (tests/hostemu/build.py forwards each wrapper body to hostemu::tc::<name>):
//   * tensor maps (cuTensorMapEncodeTiled stand-in) and TMA tile loads / stores: box copy global <-> shared with the
//     32/64/128-byte swizzle applied to the ABSOLUTE shared-memory address, zero fill / clipping out of bounds;
//   * mbarrier objects (arrival count + transaction bytes + phase parity), bounded-spin waits become fiber yields;
//   * TMEM (128 lanes x 512 fp32 columns per block, column allocator) and tcgen05.ld 32x32b.x16;
//   * tcgen05.mma kind::f16 from shared-memory matrix descriptors (start address, LBO, SBO, swizzle layout type,
//     K-major and MN-major canonical layouts), M = 64/128 x N x K = 16, fp32 accumulation;
//   * named barriers, elect.sync.
// Every asynchronous operation completes at issue (one legal execution order).  The model states the semantics this code
// base relies on; it is validated by reproducing, on the CPU, the results of the kernel tests that are green on a B200.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace hostemu {
namespace tc {
uint32_t smem_u32(const void* p);
bool elect_one();
void mbar_init(uint64_t* bar, uint32_t count);
void mbar_expect_tx(uint64_t* bar, uint32_t bytes);
void mbar_arrive(uint64_t* bar);
uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity);
void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1);
void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3);
void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3);
void named_bar_sync(int id, int nthreads);
void tmem_alloc(uint32_t* smem_result, uint32_t ncols);
void tmem_dealloc(uint32_t taddr, uint32_t ncols);
void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate);
void umma_commit(uint64_t* bar);
void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]);
void red_add_f32(float* dst, float v);
void check_align(const void* p, unsigned bytes, const char* what);
void bulk_commit_group();
void bulk_wait_group(int n);
bool async_tick(bool stuck);
inline void nop() {}
}  // namespace tc
}  // namespace hostemu
