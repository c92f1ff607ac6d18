// Error reporting, launch accounting and the TMA tensor-map encoder for librestoragen.so.
#include <atomic>
#include <stdio.h>
#include "common.cuh"
#include "internal.h"

namespace rg {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
int set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}
int check_launch(const char* kernel) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, kernel);
    return RG_OK;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// bit 0: GEMM / attention / glue kernels, bit 1: GroupNorm kernels.  Default 1: measured on the batch-2 UNet evaluation
// (profiles/r02_floor_*.txt) plain 5.94 ms, mode 3 5.82 ms, mode 1 5.60 ms -- the two-pass GroupNorm gets SLOWER with an
// early-scheduled dependent (15.5 -> 22.3 us at 2 x 64 x 64 x 320), everything else gains 0.4-1.7 us per launch.
static std::atomic<int> g_pdl{1};
bool pdl_enabled(int cls) { return (g_pdl.load(std::memory_order_relaxed) >> cls) & 1; }

int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev < 0 ? 0 : (dev >= kMaxDevices ? kMaxDevices - 1 : dev);
}

// per device: a process may drive several GPUs (one pipeline each)
int sm_count() {
    static std::atomic<int> n[kMaxDevices];
    const int dev = current_device();
    int v = n[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (v <= 0) v = 148;
        n[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device AND per function: `done` is the caller's
// function-local flag array (zero-initialised static), indexed by device
int ensure_smem_attr(const void* func, int bytes, std::atomic<bool>* done, const char* what) {
    const int dev = current_device();
    if (done[dev].load(std::memory_order_acquire)) return RG_OK;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return set_cuda_error(e, what);
    done[dev].store(true, std::memory_order_release);
    return RG_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base,
                      const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
                      const cuuint32_t* elem_strides, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_error(RG_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    CUresult r = fn(map, dtype, rank, const_cast<void*>(base), dims, strides_bytes, box, elem_strides,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[400];
        snprintf(msg, sizeof(msg),
                 "cuTensorMapEncodeTiled failed (CUresult %d): rank %u dims [%llu,%llu,%llu,%llu] stride0 %llu box "
                 "[%u,%u,%u,%u] base %p",
                 (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                 (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                 (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0,
                 rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
        return set_error(RG_ERR_TENSORMAP, msg);
    }
    return RG_OK;
}

}  // namespace rg

extern "C" const char* rg_last_error(void) { return rg::g_err; }
extern "C" int rg_version(void) { return 100; }
extern "C" int rg_operand_dtype(void) { return rg::kOperandF16 ? RG_DT_F16 : RG_DT_BF16; }
extern "C" int64_t rg_launch_count(void) { return rg::g_launches.load(); }
extern "C" int rg_device_sm_count(void) { return rg::sm_count(); }
extern "C" int rg_set_pdl(int mode) { return rg::g_pdl.exchange(mode & 3); }
