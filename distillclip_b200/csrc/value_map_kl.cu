// LastValueMapKL (reference model/loss_component/last_value_map_kl.py:10-14):
//     KLDiv(sum)( softmax(stu, dim=1).log(), softmax(tea, dim=1) )        stu, tea: [B, H, N, N] value-relation maps
// The softmax runs over the HEAD axis (dim=1): every (sample, position) owns a column of H values at stride N*N.
// One pass: each thread takes VEC consecutive positions of one sample, reads the H student and H teacher values of each,
// and writes the gradient (p^s - p^t) * grad_scale for all heads; 6 bytes per student element as the other streaming
// losses.  With m = max over both columns, es = exp(s - m), et = exp(t - m):
//     KL = W/Zt + log(Zs/Zt) = -a + log1p(a + Q/Zt),  W = sum et (t - s),  a = -W/Zt,  Q = sum [es - et + et (t - s)]
// (Q is the second-order part of Zs - Zt: the first-order terms of the two summands cancel analytically, not in fp32).
#include <type_traits>

#include "common.cuh"

namespace dcb {

constexpr int kVmThreads = 256;
constexpr int kVmMaxHeads = 16;

template <typename T, typename G, int VEC>
__global__ void __launch_bounds__(kVmThreads) value_map_kl_kernel(const T* __restrict__ s_base, const T* __restrict__ t_base,
                                                                  G* __restrict__ g_base, long long groups, long long groups_per_b,
                                                                  long long positions, int heads, float grad_scale,
                                                                  double* __restrict__ partials, const float* __restrict__ fwd_mult,
                                                                  const float* __restrict__ upstream, float expected) {
    // forward: gradients carry grad_scale * (*fwd_mult); regrad (upstream != nullptr, partials == nullptr): recompute them
    // with the true upstream gradient unless it is what the forward pass assumed
    const float fm = fwd_mult ? __ldg(fwd_mult) : 1.f;
    if (upstream) {
        const float up = __ldg(upstream);
        if (up == expected * fm) return;
        grad_scale *= up;
    } else {
        grad_scale *= fm;
    }
    double acc = 0.0;
    for (long long gi = (long long)blockIdx.x * kVmThreads + threadIdx.x; gi < groups; gi += (long long)gridDim.x * kVmThreads) {
        const long long b = ((unsigned long long)(gi | groups_per_b) >> 32) == 0 ? (long long)((unsigned)gi / (unsigned)groups_per_b) : gi / groups_per_b;
        const long long pos = (gi - b * groups_per_b) * VEC;
        const long long off = b * heads * positions + pos;
        float sv[kVmMaxHeads][VEC], tv[kVmMaxHeads][VEC];
#pragma unroll
        for (int h = 0; h < kVmMaxHeads; ++h)
            if (h < heads) {
                load_vec<T, VEC>(s_base + off + h * positions, sv[h]);
                load_vec<T, VEC>(t_base + off + h * positions, tv[h]);
            }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            float m = -INFINITY;
#pragma unroll
            for (int h = 0; h < kVmMaxHeads; ++h)
                if (h < heads) m = fmaxf(m, fmaxf(sv[h][e], tv[h][e]));
            float zs = 0.f, zt = 0.f, w = 0.f, q = 0.f;
#pragma unroll
            for (int h = 0; h < kVmMaxHeads; ++h)
                if (h < heads) {
                    const float d = tv[h][e] - sv[h][e];
                    const float es = __expf(sv[h][e] - m), et = __expf(tv[h][e] - m);
                    zs += es;
                    zt += et;
                    w = fmaf(et, d, w);
                    q += fmaf(et, d, es - et);
                    sv[h][e] = es;
                    tv[h][e] = et;
                }
            const float a = -w / zt;
            acc += (double)(-a + log1pf(a + q / zt));
            const float cs = grad_scale / zs, ct = grad_scale / zt;
#pragma unroll
            for (int h = 0; h < kVmMaxHeads; ++h)
                if (h < heads) sv[h][e] = __fsub_rn(__fmul_rn(sv[h][e], cs), __fmul_rn(tv[h][e], ct));   // no FMA contraction: exactly 0 when s == t
        }
        if (g_base) {
#pragma unroll
            for (int h = 0; h < kVmMaxHeads; ++h)
                if (h < heads) store_vec<G, VEC>(g_base + off + h * positions, sv[h]);
        }
    }
    if (partials == nullptr) return;
    acc = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

}  // namespace dcb

extern "C" int dcb_value_map_kl_fwd_bwd(const void* stu, const void* tea, void* grad_stu, int64_t batch, int64_t heads,
                                        int64_t positions, int in_dtype, int grad_dtype, float grad_scale, double* partials,
                                        int* n_partials, const float* fwd_mult, const float* upstream, float expected,
                                        void* stream) {
    using namespace dcb;
    DCB_REQUIRE(stu && tea && ((partials && n_partials) || (upstream && grad_stu)), "NULL pointer argument");
    DCB_REQUIRE(batch >= 1 && positions >= 1, "bad shape");
    DCB_REQUIRE(heads >= 1 && heads <= kVmMaxHeads, "value-map KL supports 1..%d heads (got %lld)", kVmMaxHeads, (long long)heads);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool aligned16 = (((uintptr_t)stu | (uintptr_t)tea | (uintptr_t)grad_stu) % 16) == 0;
    return dispatch_in_grad(in_dtype, grad_dtype, [&](auto tt, auto gg) -> int {
        using T = decltype(tt);
        using G = decltype(gg);
        auto launch = [&](auto vec_tag) -> int {
            constexpr int VEC = decltype(vec_tag)::value;
            const long long groups_per_b = positions / VEC, groups = batch * groups_per_b;
            long long grid = (groups + kVmThreads - 1) / kVmThreads;
            if (grid > (long long)kNumSMs * 8) grid = (long long)kNumSMs * 8;
            if (grid > DCB_MAX_PARTIALS) grid = DCB_MAX_PARTIALS;
            value_map_kl_kernel<T, G, VEC><<<(unsigned)grid, kVmThreads, 0, st>>>(
                static_cast<const T*>(stu), static_cast<const T*>(tea), static_cast<G*>(grad_stu), groups, groups_per_b, positions,
                (int)heads, grad_scale, upstream ? nullptr : partials, fwd_mult, upstream, expected);
            DCB_CUDA_OK(cudaGetLastError());
            if (n_partials) *n_partials = (int)grid;
            return 0;
        };
        // widest vector whose byte width divides every head's start offset (positions * sizeof(T)) and the base pointers
        if (aligned16 && positions % 2 == 0) return launch(std::integral_constant<int, 2>{});
        return launch(std::integral_constant<int, 1>{});
    });
}
