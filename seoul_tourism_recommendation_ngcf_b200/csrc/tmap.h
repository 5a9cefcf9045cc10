// Host side of TMA: builds CUtensorMap descriptors for 2-D fp32 row-major arrays through the driver entry point
// cuTensorMapEncodeTiled (looked up with cudaGetDriverEntryPoint, so the library does not link libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

// rows x cols fp32 array with row pitch `ld` elements; box = box_rows x box_cols elements (box_cols * 4 bytes must not
// exceed the swizzle span).  Rows / columns of a box that lie outside the array are filled with zeros.
// Returns 0 on success, the CUresult (or -1 if the entry point is missing) otherwise.
int ngcf_encode_tmap_2d(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                        uint32_t box_rows, CUtensorMapSwizzle swizzle);
