// C-ABI glue that is not tied to one loss family: error buffer, version, deterministic finalize, and the device-to-device
// copies of the sharded exchange.
#include "common.cuh"

namespace dcb {

char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

// ---------------------------------------------------------------------------------------------
// finalize: out[k] = scale[k] * sum(partials[k]) ; out[n] = sum_k percent[k] * out[k]
// (reference model/_loss.py:195-200 -- `cal_res[n] * scale` then `loss += cal_res[n] * percent[n]`)
// One CTA; partials are summed with a fixed assignment and a fixed tree, so the result is run-to-run identical.
// ---------------------------------------------------------------------------------------------
struct FinalizeParams {
    int n_terms;
    const double* partials[DCB_MAX_TERMS];
    int counts[DCB_MAX_TERMS];
    float scale[DCB_MAX_TERMS];
    float percent[DCB_MAX_TERMS];
};

// One warp per term (<= 16 terms -> 512 threads): lane-strided loads in index order, fixed shuffle tree.
__global__ void __launch_bounds__(32 * DCB_MAX_TERMS) finalize_kernel(const __grid_constant__ FinalizeParams p,
                                                                      float* __restrict__ out) {
    __shared__ double term_sum[DCB_MAX_TERMS];
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (k < p.n_terms) {
        const double* __restrict__ src = p.partials[k];
        const int n = p.counts[k];
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
        int i = lane;
        for (; i + 96 < n; i += 128) {
            v0 += src[i];
            v1 += src[i + 32];
            v2 += src[i + 64];
            v3 += src[i + 96];
        }
        for (; i < n; i += 32) v0 += src[i];
        const double v = warp_sum((v0 + v1) + (v2 + v3));
        if (lane == 0) term_sum[k] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float total = 0.f;
        for (int t = 0; t < p.n_terms; ++t) {
            // same rounding points as the reference: fp32 value, * scale, * percent, += in fp32
            const float res = (float)term_sum[t] * p.scale[t];
            out[t] = res;
            total += res * p.percent[t];
        }
        out[p.n_terms] = total;
    }
}

// ---------------------------------------------------------------------------------------------
// gather: up to 64 (dst, src, bytes) copies in ONE launch, 16 bytes per thread per step.  The sources are peer-mapped
// buffers (loads over NVLink): the latency-optimised text-row exchange for small global batches, where per-copy
// copy-engine launches would cost more than the transfers (L-CLIP stage: 0.5 MB per peer).
// ---------------------------------------------------------------------------------------------
constexpr int kGatherMax = 64;
struct GatherParams {
    void* dst[kGatherMax];
    const void* src[kGatherMax];
    long long vecs[kGatherMax];     // 16-byte units
};
__global__ void __launch_bounds__(256) peer_gather_kernel(const __grid_constant__ GatherParams p) {
    const int k = blockIdx.y;
    const uint4* __restrict__ src = static_cast<const uint4*>(p.src[k]);
    uint4* __restrict__ dst = static_cast<uint4*>(p.dst[k]);
    const long long n = p.vecs[k];
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {                 // four NVLink round trips in flight per thread
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + i + u * stride));
#pragma unroll
        for (int u = 0; u < 4; ++u) dst[i + u * stride] = v[u];
    }
    for (; i < n; i += stride) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + i));
        dst[i] = v;
    }
}

}  // namespace dcb

extern "C" {

int dcb_peer_gather(int n_copies, void* const* dst, const void* const* src, const int64_t* bytes, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(n_copies >= 1 && n_copies <= kGatherMax, "1..%d copies per launch", kGatherMax);
    GatherParams p{};
    long long max_vecs = 0;
    for (int k = 0; k < n_copies; ++k) {
        DCB_REQUIRE(dst[k] && src[k] && bytes[k] >= 0 && bytes[k] % 16 == 0 &&
                        (reinterpret_cast<uintptr_t>(dst[k]) | reinterpret_cast<uintptr_t>(src[k])) % 16 == 0,
                    "copy %d: 16-byte aligned pointers and sizes required", k);
        p.dst[k] = dst[k];
        p.src[k] = src[k];
        p.vecs[k] = bytes[k] / 16;
        max_vecs = p.vecs[k] > max_vecs ? p.vecs[k] : max_vecs;
    }
    long long bx = (max_vecs + 255) / 256;
    const long long cap = (4LL * kNumSMs + n_copies - 1) / n_copies;      // ~4 CTAs per SM in total
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    peer_gather_kernel<<<dim3((unsigned)bx, (unsigned)n_copies), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    DCB_CUDA_OK(cudaGetLastError());
    return 0;
}

int dcb_version(void) { return 100; }
int dcb_compiled_arch(void) { return 100; }
const char* dcb_last_error(void) { return dcb::error_buffer(); }

// One stream-ordered device-to-device copy (local or peer-mapped addresses: unified addressing resolves the route; peer
// pulls over NVLink run on the copy engines and leave the SMs to the tensor kernels).
int dcb_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(dst && src && bytes >= 0, "bad arguments");
    if (bytes == 0) return 0;
    DCB_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
    return 0;
}

int dcb_finalize(int n_terms, const double* const* partials, const int32_t* counts, const float* scale,
                 const float* percent, float* out, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(n_terms >= 1 && n_terms <= DCB_MAX_TERMS, "n_terms=%d out of range [1,%d]", n_terms, DCB_MAX_TERMS);
    DCB_REQUIRE(out, "out must not be NULL");
    FinalizeParams p{};
    p.n_terms = n_terms;
    for (int k = 0; k < n_terms; ++k) {
        DCB_REQUIRE(partials[k] && counts[k] >= 0, "term %d: bad partials", k);
        p.partials[k] = partials[k];
        p.counts[k] = counts[k];
        p.scale[k] = scale ? scale[k] : 1.f;
        p.percent[k] = percent ? percent[k] : 0.f;
    }
    finalize_kernel<<<1, 32 * DCB_MAX_TERMS, 0, static_cast<cudaStream_t>(stream)>>>(p, out);
    DCB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
