// capi.cu -- library-level pieces of the C ABI: error reporting, device queries, TMA descriptor encoding.
#include "common.cuh"

#include <stdarg.h>
#include <string.h>

namespace plaid {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int current_device_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev % kMaxDevices;
}

int sm_count() {
    static int cached[kMaxDevices] = {0};
    const int slot = current_device_slot();
    if (cached[slot] == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached[slot] = n;
        else
            return 148;  // B200
    }
    return cached[slot];
}

// Racing threads may both set the attribute; that is harmless (same value, idempotent call).
int ensure_dynamic_smem(const void* fn, int bytes, int (&state)[kMaxDevices], bool full_carveout) {
    const int slot = current_device_slot();
    if (bytes <= state[slot]) return PLAID_OK;
    if (bytes > 48 * 1024 || full_carveout) {
        PLAID_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        if (full_carveout)
            PLAID_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    state[slot] = bytes;
    return PLAID_OK;
}

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_bf16_2d_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    PLAID_CHECK_ARG(fn != nullptr, PLAID_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    PLAID_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0, PLAID_ERR_ARG, "TMA base must be 16-byte aligned");
    PLAID_CHECK_ARG(box_rows >= 1 && box_rows <= 256, PLAID_ERR_ARG, "TMA box rows %u out of range", box_rows);
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * sizeof(__nv_bfloat16)};
    cuuint32_t box[2] = {64, box_rows};  // 64 bf16 = 128 B = one swizzle row
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PLAID_CHECK_ARG(r == CUDA_SUCCESS, PLAID_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return PLAID_OK;
}

}  // namespace plaid

extern "C" {

int plaid_abi_version(void) { return 2; }

const char* plaid_last_error(void) { return plaid::g_err; }

const char* plaid_arch(void) { return "sm_100a"; }

}  // extern "C"
