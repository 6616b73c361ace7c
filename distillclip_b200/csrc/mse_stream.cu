// Hidden-state / embedding MSE: forward value and student gradient in one streaming pass.
//
// Replaces HiddenMSE.forward (reference model/loss_component/hidden_mse.py:9-17), EmbedMSELoss.forward
// (embed_mse.py:9-10) and the autograd backward of nn.MSELoss below them.
//
// HBM-bound: algorithmic traffic is read s + read t + write ds = 3 * sizeof(T) bytes per element
// (SURVEY.md section 8d).  All layers go through ONE launch; each CTA walks 16-byte-vectorised,
// fully coalesced tiles (8 x 128-bit loads in flight per thread), accumulates (s-t)^2 in fp32 per
// tile / double per CTA, and writes its partial sum; dcb_finalize reduces the partials in a fixed
// order, so the value is deterministic (no float atomics).
#include "stream_tiles.cuh"

namespace dcb {

struct MseSeg {
    const void* s;
    const void* t;
    void* g;
    long long n;
    long long tile_begin;
    float val_coef;    // 1 / (n * divisor)
    float grad_coef;   // grad_scale * 2 / (n * divisor)
};
struct MseParams {
    int n_seg;
    long long total_tiles;
    MseSeg seg[DCB_MAX_LAYERS];
};

constexpr int kMseThreads = kStreamThreads;

template <typename T, typename G, int VEC, int UNROLL>
__global__ void __launch_bounds__(kMseThreads) mse_stream_kernel(const __grid_constant__ MseParams p,
                                                                   double* __restrict__ partials) {
    static_assert(UNROLL == kMseUnroll, "tile body is written for kMseUnroll");
    constexpr int kTile = kMseThreads * UNROLL * VEC;
    const int tid = threadIdx.x;
    double dacc = 0.0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int k = 0;
#pragma unroll 1
        while (k + 1 < p.n_seg && tile >= p.seg[k + 1].tile_begin) ++k;
        const long long base = (tile - p.seg[k].tile_begin) * kTile;
        G* g = p.seg[k].g ? static_cast<G*>(p.seg[k].g) + base : nullptr;
        const float acc = mse_tile<T, G, VEC>(static_cast<const T*>(p.seg[k].s) + base, static_cast<const T*>(p.seg[k].t) + base,
                                              g, p.seg[k].n - base, p.seg[k].grad_coef, tid);
        dacc += (double)acc * (double)p.seg[k].val_coef;
    }
    const double total = block_sum(dacc);
    if (tid == 0) partials[blockIdx.x] = total;
}

template <typename T, typename G, int VEC>
static int launch_mse(MseParams& p, double* partials, int* n_partials, cudaStream_t stream) {
    constexpr int kUnroll = 4;
    constexpr long long kTile = (long long)kMseThreads * kUnroll * VEC;
    long long tiles = 0;
    for (int k = 0; k < p.n_seg; ++k) {
        p.seg[k].tile_begin = tiles;
        tiles += (p.seg[k].n + kTile - 1) / kTile;
    }
    p.total_tiles = tiles;
    long long grid = tiles < (long long)kNumSMs * 8 ? tiles : (long long)kNumSMs * 8;
    if (grid < 1) grid = 1;
    mse_stream_kernel<T, G, VEC, kUnroll><<<(unsigned)grid, kMseThreads, 0, stream>>>(p, partials);
    DCB_CUDA_OK(cudaGetLastError());
    *n_partials = (int)grid;
    return 0;
}

}  // namespace dcb

extern "C" int dcb_mse_fwd_bwd(int n_layers, const void* const* stu, const void* const* tea, void* const* grad_stu,
                               const int64_t* numel, int in_dtype, int grad_dtype, int divisor, float grad_scale,
                               double* partials, int* n_partials, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(n_layers >= 1 && n_layers <= DCB_MAX_LAYERS, "n_layers=%d out of range [1,%d]", n_layers, DCB_MAX_LAYERS);
    DCB_REQUIRE(divisor >= 1, "divisor must be >= 1");
    DCB_REQUIRE(partials && n_partials, "partials / n_partials must not be NULL");
    MseParams p{};
    p.n_seg = n_layers;
    bool aligned = true;
    const int gsz = dtype_size(grad_dtype);
    for (int k = 0; k < n_layers; ++k) {
        DCB_REQUIRE(stu[k] && tea[k], "layer %d: NULL input", k);
        DCB_REQUIRE(numel[k] >= 1, "layer %d: numel must be >= 1", k);
        p.seg[k].s = stu[k];
        p.seg[k].t = tea[k];
        p.seg[k].g = grad_stu ? grad_stu[k] : nullptr;
        p.seg[k].n = numel[k];
        const double denom = (double)numel[k] * (double)divisor;
        p.seg[k].val_coef = (float)(1.0 / denom);
        p.seg[k].grad_coef = (float)(2.0 * (double)grad_scale / denom);
        aligned = aligned && (((uintptr_t)stu[k] | (uintptr_t)tea[k]) % 16 == 0);
        if (p.seg[k].g) aligned = aligned && ((uintptr_t)p.seg[k].g % 16 == 0);
    }
    (void)gsz;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return dispatch_in_grad(in_dtype, grad_dtype, [&](auto tt, auto gg) -> int {
        using T = decltype(tt);
        using G = decltype(gg);
        if (aligned) return launch_mse<T, G, Elem<T>::kPer16B>(p, partials, n_partials, st);
        return launch_mse<T, G, 1>(p, partials, n_partials, st);
    });
}
