// Attention-map KL: head-mean of student and teacher maps, KL(sum) value and student gradient, one pass.
//
// Replaces AttentionProbsKL.forward (reference model/loss_component/attention_probs_kl.py:10-22):
//     s = sum_h stu[b,h,p] / Hs ;  t = sum_h tea[b,h,p] / Ht
//     value += xlogy(t,t) - t*log(s)          (nn.KLDivLoss(reduction='sum'), :8, log_target=False)
//     d value / d stu[b,h,p] = -t / (s * Hs)  (identical for every head h)
// divided by len(stu_attn_probs) (:21).
//
// HBM-bound (read s, read t, write ds = 6 B per student element for bf16).  A thread owns VEC consecutive
// positions of one sample, walks the heads with independent vector loads (the head stride P*sizeof(T)
// keeps VEC*sizeof(T) alignment because P % VEC == 0), then writes the same gradient vector to every
// student head.  All layers go through one launch; per-CTA double partials are reduced by dcb_finalize.
#include "stream_tiles.cuh"

namespace dcb {

struct AttnSeg {
    const void* s;
    const void* t;
    void* g;
    long long groups;        // batch * positions / VEC
    long long groups_per_b;  // positions / VEC
    long long positions;
    long long tile_begin;
    int hs, ht;
    float inv_hs, inv_ht;
    float val_coef;    // 1 / divisor
    float grad_coef;   // grad_scale / (hs * divisor)
};
struct AttnParams {
    int n_seg;
    long long total_tiles;
    AttnSeg seg[DCB_MAX_LAYERS];
};

constexpr int kAttnThreads = kStreamThreads;

// H = compile-time head count for both maps (0 = runtime head counts)
template <typename T, typename G, int VEC, int H>
__global__ void __launch_bounds__(kAttnThreads) attn_kl_kernel(const __grid_constant__ AttnParams p,
                                                                 double* __restrict__ partials) {
    const int tid = threadIdx.x;
    double dacc = 0.0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int k = 0;
#pragma unroll 1
        while (k + 1 < p.n_seg && tile >= p.seg[k + 1].tile_begin) ++k;
        const AttnSeg& sg = p.seg[k];
        AttnShape sh{sg.groups, sg.groups_per_b, sg.positions, sg.hs, sg.ht, sg.inv_hs, sg.inv_ht};
        const float acc = attn_tile<T, G, VEC, H>(static_cast<const T*>(sg.s), static_cast<const T*>(sg.t), static_cast<G*>(sg.g),
                                                  sh, (tile - sg.tile_begin) * kAttnThreads + tid, sg.grad_coef);
        dacc += (double)acc * (double)sg.val_coef;
    }
    const double total = block_sum(dacc);
    if (tid == 0) partials[blockIdx.x] = total;
}

template <typename T, typename G, int VEC>
static int launch_attn(AttnParams& p, double* partials, int* n_partials, cudaStream_t stream) {
    long long tiles = 0;
    int common_h = p.seg[0].hs;
    for (int k = 0; k < p.n_seg; ++k) {
        p.seg[k].groups_per_b = p.seg[k].positions / VEC;
        p.seg[k].groups *= p.seg[k].groups_per_b;   // groups held the batch on entry
        p.seg[k].tile_begin = tiles;
        tiles += (p.seg[k].groups + kAttnThreads - 1) / kAttnThreads;
        if (p.seg[k].hs != common_h || p.seg[k].ht != common_h) common_h = 0;
    }
    p.total_tiles = tiles;
    long long grid = tiles < (long long)kNumSMs * 16 ? tiles : (long long)kNumSMs * 16;
    if (grid < 1) grid = 1;
    const unsigned g = (unsigned)grid;
    bool done = false;
    if constexpr (VEC <= 4) {
        if (common_h == 12) {
            attn_kl_kernel<T, G, VEC, 12><<<g, kAttnThreads, 0, stream>>>(p, partials);
            done = true;
        } else if (common_h == 8) {
            attn_kl_kernel<T, G, VEC, 8><<<g, kAttnThreads, 0, stream>>>(p, partials);
            done = true;
        }
    }
    if (!done) attn_kl_kernel<T, G, VEC, 0><<<g, kAttnThreads, 0, stream>>>(p, partials);
    DCB_CUDA_OK(cudaGetLastError());
    *n_partials = (int)grid;
    return 0;
}

}  // namespace dcb

extern "C" int dcb_attn_kl_fwd_bwd(int n_layers, const void* const* stu, const void* const* tea, void* const* grad_stu,
                                   const int64_t* batch, const int32_t* stu_heads, const int32_t* tea_heads,
                                   const int64_t* positions, int in_dtype, int grad_dtype, int divisor,
                                   float grad_scale, double* partials, int* n_partials, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(n_layers >= 1 && n_layers <= DCB_MAX_LAYERS, "n_layers=%d out of range [1,%d]", n_layers, DCB_MAX_LAYERS);
    DCB_REQUIRE(divisor >= 1, "divisor must be >= 1");
    DCB_REQUIRE(partials && n_partials, "partials / n_partials must not be NULL");
    AttnParams p{};
    p.n_seg = n_layers;
    const int isz = dtype_size(in_dtype), gsz = dtype_size(grad_dtype);
    // widest vector (in elements) every row start stays aligned to: P % VEC == 0 and base pointers aligned
    int vec = 16 / isz;
    for (int k = 0; k < n_layers; ++k) {
        DCB_REQUIRE(stu[k] && tea[k], "layer %d: NULL input", k);
        DCB_REQUIRE(batch[k] >= 1 && positions[k] >= 1 && stu_heads[k] >= 1 && tea_heads[k] >= 1, "layer %d: bad shape", k);
        p.seg[k].s = stu[k];
        p.seg[k].t = tea[k];
        p.seg[k].g = grad_stu ? grad_stu[k] : nullptr;
        p.seg[k].groups = batch[k];
        p.seg[k].positions = positions[k];
        p.seg[k].hs = stu_heads[k];
        p.seg[k].ht = tea_heads[k];
        p.seg[k].inv_hs = 1.0f / (float)stu_heads[k];
        p.seg[k].inv_ht = 1.0f / (float)tea_heads[k];
        p.seg[k].val_coef = (float)(1.0 / (double)divisor);
        p.seg[k].grad_coef = (float)((double)grad_scale / ((double)stu_heads[k] * (double)divisor));
        while (vec > 1 && (positions[k] % vec != 0 || ((uintptr_t)stu[k] | (uintptr_t)tea[k]) % (vec * isz) != 0 ||
                           (p.seg[k].g && (uintptr_t)p.seg[k].g % (vec * gsz < 16 ? vec * gsz : 16) != 0)))
            vec >>= 1;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return dispatch_in_grad(in_dtype, grad_dtype, [&](auto tt, auto gg) -> int {
        using T = decltype(tt);
        using G = decltype(gg);
        constexpr int kMax = Elem<T>::kPer16B;
        if (vec >= kMax) return launch_attn<T, G, kMax>(p, partials, n_partials, st);
        if (vec == 4) {
            if constexpr (kMax > 4) return launch_attn<T, G, 4>(p, partials, n_partials, st);
        }
        if (vec == 2) return launch_attn<T, G, 2>(p, partials, n_partials, st);
        return launch_attn<T, G, 1>(p, partials, n_partials, st);
    });
}
