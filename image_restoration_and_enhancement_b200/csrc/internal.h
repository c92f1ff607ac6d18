// Host-side helpers shared by the translation units of librestoragen.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <string.h>
#include "../../include/restoragen.h"

namespace rg {

int set_error(int code, const char* msg);                       // records msg, returns code
int set_cuda_error(cudaError_t e, const char* where);           // records "<where>: <cuda string>", returns (int)e
int check_launch(const char* kernel);                            // cudaGetLastError() after a launch
void count_launch();
constexpr int kMaxDevices = 64;
int current_device();                                            // clamped to [0, kMaxDevices)
int sm_count();                                                  // of the current device
// once per (device, function): cudaFuncAttributeMaxDynamicSharedMemorySize; `done` = static std::atomic<bool>[kMaxDevices]
int ensure_smem_attr(const void* func, int bytes, std::atomic<bool>* done, const char* what);

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked).
int encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base,
                      const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
                      const cuuint32_t* elem_strides, CUtensorMapSwizzle swizzle);

// Launch with the programmatic-dependent-launch attribute (see common.cuh: pdl_trigger / pdl_wait).  ONLY for kernels
// that execute pdl_wait() before touching global memory.  rg_set_pdl(0) falls back to plain stream serialization.
bool pdl_enabled(int cls = 0);          // cls 0: GEMM / attention / glue kernels; cls 1: the HBM-bound norm kernels
template <int CLS = 0, typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled(CLS) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace rg
