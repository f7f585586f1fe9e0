// Host-side TMA descriptor (CUtensorMap) encoding, resolved through the runtime so the library needs no
// link-time libcuda dependency (it must load on a machine without a driver).
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "ptx.cuh"

namespace pn {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !p) {
        (void)cudaGetLastError();
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

int encode_tmap(void *out, const void *base, int esize, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                const uint32_t *box, int swizzle, const uint32_t *elem_strides) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
        return PN_ERR_CUDA;
    }
    PN_CHECK_ARG(rank >= 1 && rank <= 5 && (esize == 2 || esize == 4), "encode_tmap: bad rank %d / element size %d", rank, esize);
    cuuint64_t d[5], s[5];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = elem_strides ? elem_strides[i] : 1; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    static const CUtensorMapSwizzle sw[4] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_SWIZZLE_64B,
                                             CU_TENSOR_MAP_SWIZZLE_128B};
    CUresult r = fn(reinterpret_cast<CUtensorMap *>(out), esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                    (cuuint32_t)rank, const_cast<void *>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, sw[swizzle & 3],
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu, box %u x %u)", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 1), box[0], rank > 1 ? box[1] : 1);
        return PN_ERR_CUDA;
    }
    return PN_OK;
}

}  // namespace pn
