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
    long long total_s, total_t;   // elements (aligned mode)
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
template <typename T, typename G, int VEC, int H, int GPT = 1>
__global__ void __launch_bounds__(kAttnThreads) attn_kl_kernel(const __grid_constant__ AttnParams p,
                                                                 double* __restrict__ partials) {
    const int tid = threadIdx.x;
    double dacc = 0.0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int k = 0;
#pragma unroll 1
        while (k + 1 < p.n_seg && tile >= p.seg[k + 1].tile_begin) ++k;
        const AttnSeg& sg = p.seg[k];
        float acc;
        if constexpr (VEC == 0) {                          // aligned 16-byte mode (stream_tiles.cuh: attn_tile_aligned)
            __shared__ uint4 xchg[kAttnThreads];
            if constexpr (sizeof(T) == 2 && sizeof(G) == 2) {
                AttnShape8 sh8{sg.groups, sg.groups_per_b, sg.positions, sg.total_s, sg.total_t, sg.hs, sg.ht, sg.inv_hs, sg.inv_ht};
                acc = attn_tile_aligned<T, G, H, false>(static_cast<const T*>(sg.s), static_cast<const T*>(sg.t), static_cast<G*>(sg.g),
                                                        sh8, (tile - sg.tile_begin) * kAttnThreads + tid, sg.grad_coef, xchg, tid);
            } else {
                acc = 0.f;
            }
        } else {
            AttnShape sh{sg.groups, sg.groups_per_b, sg.positions, sg.hs, sg.ht, sg.inv_hs, sg.inv_ht};
            if constexpr (GPT > 1 && H > 0)
                acc = attn_tile_multi<T, G, VEC, H, false, GPT>(static_cast<const T*>(sg.s), static_cast<const T*>(sg.t), static_cast<G*>(sg.g),
                                                                sh, (tile - sg.tile_begin) * (kAttnThreads * GPT) + tid, kAttnThreads, sg.grad_coef);
            else
                acc = attn_tile<T, G, VEC, H>(static_cast<const T*>(sg.s), static_cast<const T*>(sg.t), static_cast<G*>(sg.g),
                                              sh, (tile - sg.tile_begin) * kAttnThreads + tid, sg.grad_coef);
        }
        dacc += (double)acc * (double)sg.val_coef;
    }
    const double total = block_sum(dacc);
    if (tid == 0) partials[blockIdx.x] = total;
}

// groups per thread: enough independent loads in flight for narrow vectors (DCB_ATTN_GPT overrides, profiling only)
template <int VEC> static int attn_gpt() {
    if (const char* e = getenv("DCB_ATTN_GPT")) return atoi(e);
    return VEC <= 4 ? 2 : 1;      // measured: image stage (VEC 4) 0.76 -> 0.86, text stage (VEC 1) 0.62 -> 0.65; 4 is worse for both
}

template <typename T, typename G, int VEC>
static int launch_attn(AttnParams& p, double* partials, int* n_partials, cudaStream_t stream) {
    long long tiles = 0;
    int common_h = p.seg[0].hs;
    for (int k = 0; k < p.n_seg; ++k)
        if (p.seg[k].hs != common_h || p.seg[k].ht != common_h) common_h = 0;
    int gpt = (common_h == 12 || common_h == 8) && VEC <= 4 && VEC > 0 ? attn_gpt<VEC>() : 1;
    if (gpt != 2 && gpt != 4) gpt = 1;
    for (int k = 0; k < p.n_seg; ++k) {
        p.seg[k].total_s = p.seg[k].groups * p.seg[k].hs * p.seg[k].positions;      // groups holds the batch on entry
        p.seg[k].total_t = p.seg[k].groups * p.seg[k].ht * p.seg[k].positions;
        p.seg[k].groups_per_b = VEC == 0 ? (p.seg[k].positions + 7) / 8 : p.seg[k].positions / (VEC == 0 ? 1 : VEC);
        p.seg[k].groups *= p.seg[k].groups_per_b;   // groups held the batch on entry
        p.seg[k].tile_begin = tiles;
        tiles += (p.seg[k].groups + kAttnThreads * gpt - 1) / (kAttnThreads * gpt);
    }
    p.total_tiles = tiles;
    long long grid = tiles < (long long)kNumSMs * 16 ? tiles : (long long)kNumSMs * 16;
    if (grid < 1) grid = 1;
    const unsigned g = (unsigned)grid;
    bool done = false;
    if constexpr (VEC == 0) {
        if (common_h == 12) attn_kl_kernel<T, G, 0, 12, 1><<<g, kAttnThreads, 0, stream>>>(p, partials);
        else if (common_h == 8) attn_kl_kernel<T, G, 0, 8, 1><<<g, kAttnThreads, 0, stream>>>(p, partials);
        else attn_kl_kernel<T, G, 0, 0, 1><<<g, kAttnThreads, 0, stream>>>(p, partials);
        done = true;
    } else if constexpr (VEC <= 4) {
#define DCB_ATTN_CASE(HH)                                                                                          \
    if (common_h == HH) {                                                                                          \
        if (gpt == 4) attn_kl_kernel<T, G, VEC, HH, 4><<<g, kAttnThreads, 0, stream>>>(p, partials);               \
        else if (gpt == 2) attn_kl_kernel<T, G, VEC, HH, 2><<<g, kAttnThreads, 0, stream>>>(p, partials);          \
        else attn_kl_kernel<T, G, VEC, HH, 1><<<g, kAttnThreads, 0, stream>>>(p, partials);                        \
        done = true;                                                                                               \
    }
        DCB_ATTN_CASE(12)
        DCB_ATTN_CASE(8)
#undef DCB_ATTN_CASE
    }
    if constexpr (VEC > 0) {
        if (!done) attn_kl_kernel<T, G, VEC, 0><<<g, kAttnThreads, 0, stream>>>(p, partials);
    }
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
    bool aligned16 = true;
    for (int k = 0; k < n_layers; ++k)
        aligned16 = aligned16 && p.seg[k].g && (((uintptr_t)stu[k] | (uintptr_t)tea[k] | (uintptr_t)p.seg[k].g) % 16 == 0);
    return dispatch_in_grad(in_dtype, grad_dtype, [&](auto tt, auto gg) -> int {
        using T = decltype(tt);
        using G = decltype(gg);
        constexpr int kMax = Elem<T>::kPer16B;
        if (vec >= kMax) return launch_attn<T, G, kMax>(p, partials, n_partials, st);
        if constexpr (sizeof(T) == 2 && sizeof(G) == 2) {
            // head rows off the 16-byte grid (odd map sizes): aligned vectors + in-register realignment
            // (default OFF: measured slower than per-thread loads with 2 groups per thread -- image stage 0.60 vs 0.86 of HBM,
            // text stage 0.63 vs 0.65; kept as a tested option, DCB_ATTN_ALIGNED=1)
            if (aligned16 && getenv("DCB_ATTN_ALIGNED") && !getenv("DCB_ATTN_NO_ALIGNED")) return launch_attn<T, G, 0>(p, partials, n_partials, st);
        }
        if (vec == 4) {
            if constexpr (kMax > 4) return launch_attn<T, G, 4>(p, partials, n_partials, st);
        }
        if (vec == 2) return launch_attn<T, G, 2>(p, partials, n_partials, st);
        return launch_attn<T, G, 1>(p, partials, n_partials, st);
    });
}
