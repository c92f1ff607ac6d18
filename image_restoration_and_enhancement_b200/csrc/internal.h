// Host-side helpers shared by the translation units of librestoragen.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "../../include/restoragen.h"

namespace rg {

int set_error(int code, const char* msg);                       // records msg, returns code
int set_cuda_error(cudaError_t e, const char* where);           // records "<where>: <cuda string>", returns (int)e
int check_launch(const char* kernel);                            // cudaGetLastError() after a launch
void count_launch();
int sm_count();

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked).
int encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base,
                      const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
                      const cuuint32_t* elem_strides, CUtensorMapSwizzle swizzle);

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace rg
