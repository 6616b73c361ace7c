// Host-side TMA descriptor encoding.  cuTensorMapEncodeTiled lives in libcuda; it is resolved through the
// runtime (cudaGetDriverEntryPoint) so the library has no link-time dependency on the driver stub.
#include <mutex>

#include "tc_common.cuh"

namespace dcb {
namespace tc {

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn resolve_encode() {
    static EncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(p);
    });
    return fn;
}

int encode_tile_map_16bit(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes,
                          uint32_t box_rows) {
    EncodeFn encode = resolve_encode();
    if (!encode) return fail("cuTensorMapEncodeTiled is not available from the CUDA driver");
    // cuTensorMapEncodeTiled is a DRIVER call: it needs a current context on the calling thread.  Runtime calls bind the
    // primary context lazily, driver calls do not -- and autograd runs backward on its own worker thread, where this may be
    // the first CUDA call (CUresult 201 otherwise).  Once per thread.
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) {
        DCB_CUDA_OK(cudaFree(nullptr));
        ctx_bound = true;
    }
    if (reinterpret_cast<uintptr_t>(base) % 16 != 0) return fail("TMA: matrix base must be 16-byte aligned");
    if (row_pitch_bytes % 16 != 0) return fail("TMA: row pitch (%llu bytes) must be a multiple of 16", (unsigned long long)row_pitch_bytes);
    if (box_rows == 0 || box_rows > 256) return fail("TMA: box_rows=%u out of range", box_rows);
    const cuuint64_t gdim[2] = {cols, rows};                 // innermost first
    const cuuint64_t gstride[1] = {row_pitch_bytes};         // stride of dim 1 in bytes
    const cuuint32_t box[2] = {64, box_rows};                // 64 x 2 B = 128 B = one swizzle row
    const cuuint32_t estr[2] = {1, 1};
    // bf16 and fp16 are both plain 16-bit payloads for TMA; OOB elements are filled with zeros
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

}  // namespace tc
}  // namespace dcb
