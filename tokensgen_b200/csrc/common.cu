#include "common.h"

#include <mutex>
#include <string>

namespace tg {

static thread_local std::string t_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_last_error = buf;
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(int(e), "%s: launch failed: %s", what, cudaGetErrorString(e));
    return 0;
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

static int encode(CUtensorMap* out, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(999, "cuTensorMapEncodeTiled not available from the driver");
    cuuint32_t elem_strides[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, cuuint32_t(rank), const_cast<void*>(base), dims, strides,
                    box, elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(int(r), "cuTensorMapEncodeTiled failed (CUresult %d): base=%p rank=%d dims=(%llu,%llu,%llu)", int(r),
                    base, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                    (unsigned long long)(rank > 2 ? dims[2] : 0));
    return 0;
}

int make_tmap_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box, const uint32_t* elem_strides) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(999, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t d[5];
    cuuint64_t st[4];
    cuuint32_t b[5], es[5];
    for (int i = 0; i < rank; ++i) {
        d[i] = dims[i];
        b[i] = box[i];
        es[i] = elem_strides[i];
        if (i + 1 < rank) st[i] = strides_bytes[i];
    }
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, cuuint32_t(rank), const_cast<void*>(base), d, st, b, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(int(r), "cuTensorMapEncodeTiled failed (CUresult %d): base=%p rank=%d dims=(%llu,%llu,%llu,%llu) box=(%u,%u,%u,%u)",
                    int(r), base, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                    (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1],
                    rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return 0;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_rows) {
    cuuint64_t dims[3] = {inner, rows, 1};
    cuuint64_t strides[2] = {row_stride_bytes, 0};
    cuuint32_t box[3] = {box_inner, box_rows, 1};
    return encode(out, base, 2, dims, strides, box);
}

int make_tmap_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batch,
                 uint64_t row_stride_bytes, uint64_t batch_stride_bytes, uint32_t box_inner, uint32_t box_rows) {
    cuuint64_t dims[3] = {inner, rows, batch};
    cuuint64_t strides[2] = {row_stride_bytes, batch_stride_bytes};
    cuuint32_t box[3] = {box_inner, box_rows, 1};
    return encode(out, base, 3, dims, strides, box);
}

}  // namespace tg

extern "C" {

int tg_version(void) { return TG_VERSION; }

const char* tg_last_error(void) { return tg::t_last_error.c_str(); }

}  // extern "C"
