// tmap.cuh — host-side TMA tensor-map construction (cuTensorMapEncodeTiled through
// cudaGetDriverEntryPoint, so libgdlb200.so has no link-time dependency on libcuda).
#pragma once
#include "common.cuh"

namespace gdl {

// dtype codes shared with include/gdl_b200.h
enum : int { kDtBF16 = 0, kDtF32 = 1, kDtF16 = 2 };

// NHWC activation view: element (n,h,w,c) lives at base + ((n*H + h)*W + w)*ld + c.
// `ld` (pixel stride, in elements) lets a channel slice of a wider buffer be addressed.
int make_tmap_nhwc(CUtensorMap* out, const void* base, int dtype, long long C, long long W,
                   long long H, long long N, long long ld, int boxC, int boxW, int boxH,
                   int swizzle_bytes);

// Row-major 2D matrix [rows][cols] (cols contiguous), row stride ld elements.
int make_tmap_2d(CUtensorMap* out, const void* base, int dtype, long long cols, long long rows,
                 long long ld, int boxCols, int boxRows, int swizzle_bytes);

}  // namespace gdl
