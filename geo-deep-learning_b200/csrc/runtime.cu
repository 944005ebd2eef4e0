// runtime.cu — error channel, driver entry points, tensor-map encoding.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "../../include/gdl_b200.h"
#include "tmap.cuh"
#include "det_reduce.cuh"

namespace gdl {

static thread_local char g_err[1024] = {0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_last_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return GDL_ERR_CUDA;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (PFN_encodeTiled)p;
  });
  return fn;
}

static CUtensorMapSwizzle swz(int bytes) {
  switch (bytes) {
    case 128: return CU_TENSOR_MAP_SWIZZLE_128B;
    case 64: return CU_TENSOR_MAP_SWIZZLE_64B;
    case 32: return CU_TENSOR_MAP_SWIZZLE_32B;
    default: return CU_TENSOR_MAP_SWIZZLE_NONE;
  }
}

static int encode(CUtensorMap* out, const void* base, int dtype, int rank, const cuuint64_t* dims,
                  const cuuint64_t* strides_bytes, const cuuint32_t* box, int swizzle_bytes) {
  PFN_encodeTiled fn = get_encode();
  GDL_REQUIRE(fn != nullptr, GDL_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  GDL_REQUIRE(dtype == kDtBF16 || dtype == kDtF16 || dtype == kDtF32, GDL_ERR_INVALID, "tensor map: bad dtype %d", dtype);
  GDL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, GDL_ERR_INVALID,
              "tensor map: base address must be 16-byte aligned");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out,
                  dtype == kDtBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                   : (dtype == kDtF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32),
                  (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error(
        "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u] "
        "stride1 %llu swz %d",
        (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
        (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
        box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0,
        (unsigned long long)strides_bytes[0], swizzle_bytes);
    return GDL_ERR_CUDA;
  }
  return 0;
}

int make_tmap_nhwc(CUtensorMap* out, const void* base, int dtype, long long C, long long W,
                   long long H, long long N, long long ld, int boxC, int boxW, int boxH,
                   int swizzle_bytes) {
  const long long esz = dtype == kDtF32 ? 4 : 2;
  GDL_REQUIRE((ld * esz) % 16 == 0, GDL_ERR_INVALID, "NHWC pixel stride must be a multiple of 16 bytes (got %lld elements)", ld);
  GDL_REQUIRE(boxC * esz <= swizzle_bytes || swizzle_bytes == 0, GDL_ERR_INVALID,
              "tensor map: inner box (%d B) exceeds swizzle span %d", (int)(boxC * esz), swizzle_bytes);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t str[3] = {(cuuint64_t)(ld * esz), (cuuint64_t)(W * ld * esz), (cuuint64_t)(H * W * ld * esz)};
  cuuint32_t box[4] = {(cuuint32_t)boxC, (cuuint32_t)boxW, (cuuint32_t)boxH, 1};
  return encode(out, base, dtype, 4, dims, str, box, swizzle_bytes);
}

int make_tmap_2d(CUtensorMap* out, const void* base, int dtype, long long cols, long long rows,
                 long long ld, int boxCols, int boxRows, int swizzle_bytes) {
  GDL_REQUIRE((ld * 2) % 16 == 0, GDL_ERR_INVALID, "matrix row stride must be a multiple of 8 elements (got %lld)", ld);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t str[1] = {(cuuint64_t)(ld * 2)};
  cuuint32_t box[2] = {(cuuint32_t)boxCols, (cuuint32_t)boxRows};
  return encode(out, base, dtype, 2, dims, str, box, swizzle_bytes);
}

// ---- deterministic-reduction workspace (det_reduce.cuh): registered per device by the caller -----------------------
struct WsEntry {
  void* ptr;
  long long bytes;
};
static WsEntry g_ws[64];
int g_opt_pdl = 0;  // gdl_set_option("pdl", 0/1): launches carry the programmatic-stream-serialization attribute (common.cuh)
int g_opt_deterministic = 1;  // gdl_set_option("deterministic", 0/1); only effective with a registered workspace

DetWs det_workspace() {
  DetWs w;
  w.ctr = nullptr;
  w.slots = nullptr;
  w.slot_floats = 0;
  int dev = 0;
  if (!g_opt_deterministic || cudaGetDevice(&dev) != cudaSuccess) return w;
  const WsEntry e = g_ws[dev & 63];
  if (e.ptr == nullptr || e.bytes <= kDetCtrBytes) return w;
  w.ctr = reinterpret_cast<unsigned*>(e.ptr);
  w.slots = reinterpret_cast<float*>(reinterpret_cast<char*>(e.ptr) + kDetCtrBytes);
  w.slot_floats = (e.bytes - kDetCtrBytes) / 4;
  return w;
}

}  // namespace gdl

extern "C" {

long long gdl_query_workspace_bytes(void) { return gdl::kDetCtrBytes + 192ll * 1024 * 1024; }

int gdl_set_workspace(void* ptr, long long bytes, void* stream) {
  int dev = 0;
  GDL_CHECK_CUDA(cudaGetDevice(&dev));
  if (ptr == nullptr) {
    gdl::g_ws[dev & 63] = gdl::WsEntry{nullptr, 0};
    return 0;
  }
  GDL_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 255) == 0 && bytes >= gdl::kDetCtrBytes + (1ll << 20), GDL_ERR_INVALID,
              "set_workspace: need a 256-byte aligned buffer of at least %lld bytes (gdl_query_workspace_bytes() recommends %lld)",
              gdl::kDetCtrBytes + (1ll << 20), gdl_query_workspace_bytes());
  // tickets and turnstiles are zero at rest; every kernel that uses them leaves them at zero
  GDL_CHECK_CUDA(cudaMemsetAsync(ptr, 0, (size_t)gdl::kDetCtrBytes, (cudaStream_t)stream));
  gdl::g_ws[dev & 63] = gdl::WsEntry{ptr, bytes};
  return 0;
}

const char* gdl_last_error(void) { return gdl::g_err; }

int gdl_version(void) { return GDL_B200_VERSION; }

int gdl_device_info(int* sm_count, int* cc_major, int* cc_minor, unsigned long long* total_mem) {
  int dev = 0;
  GDL_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  GDL_CHECK_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (total_mem) *total_mem = (unsigned long long)p.totalGlobalMem;
  return 0;
}

}  // extern "C"
