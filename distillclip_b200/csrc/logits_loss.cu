// Per-module API on MATERIALISED logits: HardLabel.forward(stu_logits) and SoftLabel.forward(stu_logits, tea_logits)
// keep the reference signatures (model/loss_component/hard_label.py:10-12, soft_label.py:11-16), so they receive an
// [n, n] logits tensor (or its `.T` view, model/component/clip_model.py:44).  LossCalculator's two-tower path does not
// come through here (it runs the fused kernels from the embeddings); this is the drop-in for direct module use.
//
// One CTA per logical row, strided element access so transposed views need no copy.  Arbitrary logit ranges:
// a true running max is used (unlike the fused cosine path).
#include "common.cuh"

namespace dcb {

constexpr int kLogitThreads = 256;

template <typename T>
__device__ __forceinline__ float ld_logit(const T* base, long long rs, long long cs, long long i, long long j) {
    return Elem<T>::to_f(base[i * rs + j * cs]);
}

__device__ __forceinline__ float block_max(float v) {
    __shared__ float part[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) part[warp] = v;
    __syncthreads();
    float r = lane < (blockDim.x >> 5) ? part[lane] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
    __syncthreads();
    return r;     // valid in every thread
}
__device__ __forceinline__ float block_sum_all(float v) {
    __shared__ float part[32];
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) part[warp] = v;
    __syncthreads();
    float r = lane < (blockDim.x >> 5) ? part[lane] : 0.f;
    r = warp_sum(r);
    __syncthreads();
    return r;
}

// saved[i] = {m_s, Z_s, m_t, Z_t}  (m = row max of the raw logits, Z = sum exp((x - m) * inv_temp)); hard label uses
// inv_temp = 1 and only the first two.  rowloss[i] (double) = CE_i (hard) or KL_i without the T^2 factor (soft).
template <typename T, bool kSoft>
__global__ void __launch_bounds__(kLogitThreads) logits_stats_kernel(const T* __restrict__ s, long long srs, long long scs,
                                                                     const T* __restrict__ t, long long trs, long long tcs,
                                                                     int n, float inv_temp, float4* __restrict__ saved,
                                                                     double* __restrict__ rowloss) {
    const long long i = blockIdx.x;
    float ms = -INFINITY, mt = -INFINITY;
    for (int j = threadIdx.x; j < n; j += kLogitThreads) {
        ms = fmaxf(ms, ld_logit(s, srs, scs, i, j));
        if (kSoft) mt = fmaxf(mt, ld_logit(t, trs, tcs, i, j));
    }
    ms = block_max(ms);
    if (kSoft) mt = block_max(mt);
    float zs = 0.f, zt = 0.f, w = 0.f;
    for (int j = threadIdx.x; j < n; j += kLogitThreads) {
        const float sv = ld_logit(s, srs, scs, i, j);
        zs += __expf((sv - ms) * inv_temp);
        if (kSoft) {
            const float tv = ld_logit(t, trs, tcs, i, j);
            const float et = __expf((tv - mt) * inv_temp);
            zt += et;
            w = fmaf(et, tv - sv, w);
        }
    }
    zs = block_sum_all(zs);
    if (kSoft) {
        zt = block_sum_all(zt);
        w = block_sum_all(w);
    }
    if (threadIdx.x == 0) {
        saved[i] = make_float4(ms, zs, mt, zt);
        if (kSoft) {
            // double: KL_i is a small difference of O(1) terms
            rowloss[i] = (double)w * (double)inv_temp / (double)zt + ((double)ms - (double)mt) * (double)inv_temp +
                         log((double)zs / (double)zt);
        } else {
            rowloss[i] = (double)ms + log((double)zs) - (double)ld_logit(s, srs, scs, i, i);
        }
    }
}

// hard: g_ij = up (softmax_ij - [i==j]) / n          soft: g_ij = up T (p^s_ij - p^t_ij)
template <typename T, typename G, bool kSoft>
__global__ void __launch_bounds__(kLogitThreads) logits_grads_kernel(const T* __restrict__ s, long long srs, long long scs,
                                                                     const T* __restrict__ t, long long trs, long long tcs,
                                                                     int n, float inv_temp, const float4* __restrict__ saved,
                                                                     const float* __restrict__ upstream, G* __restrict__ grad) {
    const long long i = blockIdx.x;
    const float4 sv4 = saved[i];
    const float up = upstream[0];
    const float cs_ = kSoft ? up / (inv_temp * sv4.y) : up / ((float)n * sv4.y);
    const float ct_ = kSoft ? up / (inv_temp * sv4.w) : 0.f;
    const float lab = up / (float)n;
    G* __restrict__ g = grad + i * n;
    for (int j = threadIdx.x; j < n; j += kLogitThreads) {
        float v = __expf((ld_logit(s, srs, scs, i, j) - sv4.x) * inv_temp) * cs_;
        if (kSoft) v -= __expf((ld_logit(t, trs, tcs, i, j) - sv4.z) * inv_temp) * ct_;
        else if (j == i) v -= lab;
        g[j] = Elem<G>::from_f(v);
    }
}

// CLIPCosDiff (reference model/loss_component/clip_cos_diff.py:5-23):
//   value = mean_i relu(t_ii - s_ii) + mean_{i != j} relu(s_ij - t_ij)      (get_neg_element = all off-diagonal entries)
// rowloss[i] = relu(t_ii - s_ii) / n + sum_{j != i} relu(s_ij - t_ij) / (n (n-1));  relu'(0) = 0.
template <typename T>
__global__ void __launch_bounds__(kLogitThreads) cos_diff_stats_kernel(const T* __restrict__ s, long long srs, long long scs,
                                                                       const T* __restrict__ t, long long trs, long long tcs,
                                                                       int n, double* __restrict__ rowloss) {
    const long long i = blockIdx.x;
    float neg = 0.f;
    for (int j = threadIdx.x; j < n; j += kLogitThreads)
        if (j != i) neg += fmaxf(ld_logit(s, srs, scs, i, j) - ld_logit(t, trs, tcs, i, j), 0.f);
    neg = block_sum_all(neg);
    if (threadIdx.x == 0) {
        const float pos = fmaxf(ld_logit(t, trs, tcs, i, i) - ld_logit(s, srs, scs, i, i), 0.f);
        rowloss[i] = (double)pos / (double)n + (n > 1 ? (double)neg / ((double)n * (double)(n - 1)) : 0.0);
    }
}
template <typename T, typename G>
__global__ void __launch_bounds__(kLogitThreads) cos_diff_grads_kernel(const T* __restrict__ s, long long srs, long long scs,
                                                                       const T* __restrict__ t, long long trs, long long tcs,
                                                                       int n, const float* __restrict__ upstream, G* __restrict__ grad) {
    const long long i = blockIdx.x;
    const float up = upstream[0];
    const float gpos = -up / (float)n, gneg = n > 1 ? up / ((float)n * (float)(n - 1)) : 0.f;
    G* __restrict__ g = grad + i * n;
    for (int j = threadIdx.x; j < n; j += kLogitThreads) {
        const float sv = ld_logit(s, srs, scs, i, j), tv = ld_logit(t, trs, tcs, i, j);
        const float v = (j == i) ? (tv > sv ? gpos : 0.f) : (sv > tv ? gneg : 0.f);
        g[j] = Elem<G>::from_f(v);
    }
}

}  // namespace dcb

extern "C" {

int dcb_logits_row_stats(const void* stu_logits, int64_t stu_rs, int64_t stu_cs, const void* tea_logits, int64_t tea_rs,
                         int64_t tea_cs, int64_t n, int dtype, float temperature, int mode, float* saved, double* rowloss,
                         void* stream) {
    using namespace dcb;
    DCB_REQUIRE(stu_logits && saved && rowloss && n >= 1 && n < (1ll << 31), "bad arguments");
    DCB_REQUIRE(mode == 0 || (mode == 1 && tea_logits && temperature > 0.f) || (mode == 2 && tea_logits),
                "soft label needs teacher logits and a positive temperature; cos_diff needs teacher logits");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mode == 2) {
        switch (dtype) {
            case DCB_BF16: cos_diff_stats_kernel<__nv_bfloat16><<<(unsigned)n, kLogitThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(stu_logits), stu_rs, stu_cs, static_cast<const __nv_bfloat16*>(tea_logits), tea_rs, tea_cs, (int)n, rowloss); break;
            case DCB_F16: cos_diff_stats_kernel<__half><<<(unsigned)n, kLogitThreads, 0, st>>>(static_cast<const __half*>(stu_logits), stu_rs, stu_cs, static_cast<const __half*>(tea_logits), tea_rs, tea_cs, (int)n, rowloss); break;
            case DCB_F32: cos_diff_stats_kernel<float><<<(unsigned)n, kLogitThreads, 0, st>>>(static_cast<const float*>(stu_logits), stu_rs, stu_cs, static_cast<const float*>(tea_logits), tea_rs, tea_cs, (int)n, rowloss); break;
            default: return fail("unknown dtype %d", dtype);
        }
        DCB_CUDA_OK(cudaGetLastError());
        return 0;
    }
    const float inv_temp = mode == 1 ? 1.0f / temperature : 1.0f;
    float4* sv = reinterpret_cast<float4*>(saved);
#define DCB_LAUNCH_STATS(T)                                                                                            \
    if (mode == 1)                                                                                                     \
        logits_stats_kernel<T, true><<<(unsigned)n, kLogitThreads, 0, st>>>(static_cast<const T*>(stu_logits), stu_rs, stu_cs, \
            static_cast<const T*>(tea_logits), tea_rs, tea_cs, (int)n, inv_temp, sv, rowloss);                          \
    else                                                                                                               \
        logits_stats_kernel<T, false><<<(unsigned)n, kLogitThreads, 0, st>>>(static_cast<const T*>(stu_logits), stu_rs, stu_cs, \
            nullptr, 0, 0, (int)n, inv_temp, sv, rowloss);
    switch (dtype) {
        case DCB_BF16: DCB_LAUNCH_STATS(__nv_bfloat16) break;
        case DCB_F16: DCB_LAUNCH_STATS(__half) break;
        case DCB_F32: DCB_LAUNCH_STATS(float) break;
        default: return fail("unknown dtype %d", dtype);
    }
#undef DCB_LAUNCH_STATS
    DCB_CUDA_OK(cudaGetLastError());
    return 0;
}

int dcb_logits_row_grads(const void* stu_logits, int64_t stu_rs, int64_t stu_cs, const void* tea_logits, int64_t tea_rs,
                         int64_t tea_cs, int64_t n, int dtype, float temperature, int mode, const float* saved,
                         const float* upstream, void* grad_logits, int grad_dtype, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(stu_logits && saved && upstream && grad_logits && n >= 1, "bad arguments");
    DCB_REQUIRE(mode == 0 || (mode == 1 && tea_logits && temperature > 0.f) || (mode == 2 && tea_logits),
                "soft label needs teacher logits and a positive temperature; cos_diff needs teacher logits");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float inv_temp = mode == 1 ? 1.0f / temperature : 1.0f;
    const float4* sv = reinterpret_cast<const float4*>(saved);
    return dispatch_in_grad(dtype, grad_dtype, [&](auto tt, auto gg) -> int {
        using T = decltype(tt);
        using G = decltype(gg);
        if (mode == 2)
            cos_diff_grads_kernel<T, G><<<(unsigned)n, kLogitThreads, 0, st>>>(
                static_cast<const T*>(stu_logits), stu_rs, stu_cs, static_cast<const T*>(tea_logits), tea_rs, tea_cs, (int)n,
                upstream, static_cast<G*>(grad_logits));
        else if (mode == 1)
            logits_grads_kernel<T, G, true><<<(unsigned)n, kLogitThreads, 0, st>>>(
                static_cast<const T*>(stu_logits), stu_rs, stu_cs, static_cast<const T*>(tea_logits), tea_rs, tea_cs, (int)n,
                inv_temp, sv, upstream, static_cast<G*>(grad_logits));
        else
            logits_grads_kernel<T, G, false><<<(unsigned)n, kLogitThreads, 0, st>>>(
                static_cast<const T*>(stu_logits), stu_rs, stu_cs, nullptr, 0, 0, (int)n, inv_temp, sv, upstream,
                static_cast<G*>(grad_logits));
        DCB_CUDA_OK(cudaGetLastError());
        return 0;
    });
}

}  // extern "C"
