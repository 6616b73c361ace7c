// Row-softmax losses between the pooled outputs of student and teacher ([rows, cols], contiguous):
//   mode 0  OutKLLoss  (reference model/loss_component/out_kl.py:12-16):
//           KLDiv(sum)(log_softmax(s/T, dim=1), softmax(t/T, dim=1)) * T^2
//   mode 1  OutCELoss  (reference model/loss_component/out_ce.py:9-13):
//           CrossEntropy(mean)(s, softmax(t, dim=1)) = mean_i [ lse(s_i) - sum_j p^t_ij s_ij ]
// One CTA per row; saved[i] = {max_s, Z_s, max_t, Z_t}; rowloss[i] in double (the KL is a small difference of O(1) terms).
// Backward: mode 0: g = up T (p^s - p^t);  mode 1: g = up (softmax(s) - p^t) / rows.
#include "common.cuh"

namespace dcb {

constexpr int kRowThreads = 128;

__device__ __forceinline__ float rs_block_max(float v) {
    __shared__ float part[4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
    __syncthreads();
    const float r = fmaxf(fmaxf(part[0], part[1]), fmaxf(part[2], part[3]));
    __syncthreads();
    return r;
}
__device__ __forceinline__ float rs_block_sum(float v) {
    __shared__ float part[4];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
    __syncthreads();
    const float r = (part[0] + part[1]) + (part[2] + part[3]);
    __syncthreads();
    return r;
}

template <typename T>
__global__ void __launch_bounds__(kRowThreads) row_softmax_stats_kernel(const T* __restrict__ s, const T* __restrict__ t, int cols,
                                                                        float inv_temp, int mode, float4* __restrict__ saved,
                                                                        double* __restrict__ rowloss) {
    const long long i = blockIdx.x;
    const T* __restrict__ sp = s + i * cols;
    const T* __restrict__ tp = t + i * cols;
    float ms = -INFINITY, mt = -INFINITY;
    for (int j = threadIdx.x; j < cols; j += kRowThreads) {
        ms = fmaxf(ms, Elem<T>::to_f(sp[j]));
        mt = fmaxf(mt, Elem<T>::to_f(tp[j]));
    }
    ms = rs_block_max(ms);
    mt = rs_block_max(mt);
    float zs = 0.f, zt = 0.f, w = 0.f;
    for (int j = threadIdx.x; j < cols; j += kRowThreads) {
        const float sv = Elem<T>::to_f(sp[j]), tv = Elem<T>::to_f(tp[j]);
        const float et = __expf((tv - mt) * inv_temp);
        zs += __expf((sv - ms) * inv_temp);
        zt += et;
        w = fmaf(et, mode == 0 ? tv - sv : sv - ms, w);      // KL: sum e_t (t - s);  CE: sum e_t (s - max_s)
    }
    zs = rs_block_sum(zs);
    zt = rs_block_sum(zt);
    w = rs_block_sum(w);
    if (threadIdx.x == 0) {
        saved[i] = make_float4(ms, zs, mt, zt);
        if (mode == 0)
            rowloss[i] = (double)w * (double)inv_temp / (double)zt + ((double)ms - (double)mt) * (double)inv_temp +
                         log((double)zs / (double)zt);
        else      // lse(s) - sum p_t s = log Zs - sum p_t (s - max_s)
            rowloss[i] = log((double)zs) - (double)w / (double)zt;
    }
}

template <typename T, typename G>
__global__ void __launch_bounds__(kRowThreads) row_softmax_grads_kernel(const T* __restrict__ s, const T* __restrict__ t, int cols,
                                                                        float inv_temp, float coef, const float4* __restrict__ saved,
                                                                        const float* __restrict__ upstream, G* __restrict__ grad) {
    const long long i = blockIdx.x;
    const float4 sv4 = saved[i];
    const float k = upstream[0] * coef;
    const float cs = k / sv4.y, ct = k / sv4.w;
    const T* __restrict__ sp = s + i * cols;
    const T* __restrict__ tp = t + i * cols;
    G* __restrict__ g = grad + i * cols;
    for (int j = threadIdx.x; j < cols; j += kRowThreads)
        g[j] = Elem<G>::from_f(__expf((Elem<T>::to_f(sp[j]) - sv4.x) * inv_temp) * cs -
                               __expf((Elem<T>::to_f(tp[j]) - sv4.z) * inv_temp) * ct);
}

}  // namespace dcb

extern "C" {

int dcb_row_softmax_stats(const void* stu, const void* tea, int64_t rows, int64_t cols, int dtype, float temperature, int mode,
                          float* saved, double* rowloss, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(stu && tea && saved && rowloss && rows >= 1 && cols >= 1 && cols < (1ll << 31), "bad arguments");
    DCB_REQUIRE(mode == 1 || (mode == 0 && temperature > 0.f), "mode 0 (out_kl) needs a positive temperature; mode 1 = out_ce");
    const float inv_temp = mode == 0 ? 1.0f / temperature : 1.0f;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float4* sv = reinterpret_cast<float4*>(saved);
    switch (dtype) {
        case DCB_BF16: row_softmax_stats_kernel<__nv_bfloat16><<<(unsigned)rows, kRowThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(stu), static_cast<const __nv_bfloat16*>(tea), (int)cols, inv_temp, mode, sv, rowloss); break;
        case DCB_F16: row_softmax_stats_kernel<__half><<<(unsigned)rows, kRowThreads, 0, st>>>(static_cast<const __half*>(stu), static_cast<const __half*>(tea), (int)cols, inv_temp, mode, sv, rowloss); break;
        case DCB_F32: row_softmax_stats_kernel<float><<<(unsigned)rows, kRowThreads, 0, st>>>(static_cast<const float*>(stu), static_cast<const float*>(tea), (int)cols, inv_temp, mode, sv, rowloss); break;
        default: return fail("unknown dtype %d", dtype);
    }
    DCB_CUDA_OK(cudaGetLastError());
    return 0;
}

int dcb_row_softmax_grads(const void* stu, const void* tea, int64_t rows, int64_t cols, int dtype, float temperature, int mode,
                          const float* saved, const float* upstream, void* grad, int grad_dtype, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(stu && tea && saved && upstream && grad && rows >= 1 && cols >= 1, "bad arguments");
    DCB_REQUIRE(mode == 1 || (mode == 0 && temperature > 0.f), "mode 0 (out_kl) needs a positive temperature; mode 1 = out_ce");
    const float inv_temp = mode == 0 ? 1.0f / temperature : 1.0f;
    const float coef = mode == 0 ? temperature : 1.0f / (float)rows;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float4* sv = reinterpret_cast<const float4*>(saved);
    return dispatch_in_grad(dtype, grad_dtype, [&](auto tt, auto gg) -> int {
        using T = decltype(tt);
        using G = decltype(gg);
        row_softmax_grads_kernel<T, G><<<(unsigned)rows, kRowThreads, 0, st>>>(static_cast<const T*>(stu), static_cast<const T*>(tea),
                                                                              (int)cols, inv_temp, coef, sv, upstream, static_cast<G*>(grad));
        DCB_CUDA_OK(cudaGetLastError());
        return 0;
    });
}

}  // extern "C"
