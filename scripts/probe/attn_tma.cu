// Attention-map losses (KL or MSE of the head means) with TMA-staged tiles: the HBM3e-friendly version of attn_kl.cu.
//
// Replaces AttentionProbsKL.forward (reference model/loss_component/attention_probs_kl.py:10-22) and
// AttentionProbsMSE / AttentionScoreMSE (attention_probs_mse.py:10-22, attention_score_mse.py:10-22) + autograd.
//
// Why: a head row of an [B, H, N, N] bf16 map starts every N*N*2 bytes -- 5000 B for N = 50, 11858 B for N = 77 -- so
// per-thread vector loads are limited to 8 B resp. 2 B.  A whole SAMPLE ([H*N*N] elements) is 16-byte
// aligned, so the maps are described to TMA as 2-D tensors [B, H*P] and every (head, position chunk) is fetched as
// boxes {256 elements, 1 sample} starting at the element offset (h*P + p0) rounded DOWN to 8 elements (TMA needs a
// 16-byte aligned box start: an unaligned one raises an illegal-instruction fault, measured); the residue m_h < 8 is an
// offset into the shared-memory row, so the SM only ever sees shared memory.  Per tile (one sample, pc - 8 positions):
// the producer warp issues the loads of all Hs + Ht head rows into a 3-stage ring; 8 compute warps reduce over heads from
// shared memory (one position pair per thread) and write the head-independent gradient to every student head row with
// 4-byte (2-byte on odd element offsets) global stores, 128 contiguous bytes per warp instruction.
//
// Algorithmic traffic: read s, read t, write ds = 6 B per student element (bf16); the 2/3 that are reads go through
// the copy engine in 512-byte requests.
#include "tc_common.cuh"

namespace dcb {

namespace atma {
constexpr int kBox = 256;                 // elements per TMA box (inner extent <= 256)
constexpr int kComputeWarps = 16;
constexpr int kComputeThreads = 32 * kComputeWarps;
constexpr int kThreads = 32 + kComputeThreads;   // warp 0 producer + compute warps
constexpr int kStages = 3;
constexpr int kMaxLayers = 8;
}  // namespace atma

struct AttnTmaLayer {
    CUtensorMap map_s, map_t;             // [B, Hs*P], [B, Ht*P]
    void* grad;                           // [B, Hs, P] or nullptr
    int batch, hs, ht, positions;
    int chunks;                           // position chunks per sample
    long long tile_begin;
    float inv_hs, inv_ht, val_coef, grad_coef;
    int term;
};
struct AttnTmaParams {
    int n_layers, mode;                   // mode 0 = KL, 1 = MSE
    int pc;                               // positions per tile (multiple of kBox)
    int max_rows;                         // max over layers of hs + ht
    int max_hs;
    long long total_tiles;
    double* partials;                     // [n_terms][partial_stride], this kernel fills [term][0 .. gridDim.x)
    int partial_stride;
    AttnTmaLayer layer[atma::kMaxLayers];
};

template <typename T>
__global__ void __launch_bounds__(atma::kThreads, 1) attn_tma_kernel(const __grid_constant__ AttnTmaParams p) {
    using namespace atma;
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int in_stage_bytes = p.max_rows * p.pc * 2;                 // [rows][pc] 16-bit
    const uint32_t in_ring = smem_base;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_gen + kStages * in_stage_bytes);
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kStages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int span = p.pc - 8;                                        // positions per tile

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, kComputeWarps);          // one arrive per compute warp
        }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == 0) {
        // ---------------------------------------------------------------- producer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            int k = 0;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                while (k + 1 < p.n_layers && tile >= p.layer[k + 1].tile_begin) ++k;
                const AttnTmaLayer& L = p.layer[k];
                const long long lt = tile - L.tile_begin;
                const int b = (int)(lt / L.chunks), c = (int)(lt % L.chunks);
                const int p0 = c * span;
                const int len = min(span, L.positions - p0);
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                const uint32_t dst = in_ring + stage * in_stage_bytes;
                const uint32_t full = bar_full + 8 * stage;
                int boxes = 0;
                for (int h = 0; h < L.hs + L.ht; ++h) {
                    const int hh = h < L.hs ? h : h - L.hs;
                    boxes += (((hh * L.positions + p0) & 7) + len + kBox - 1) / kBox;
                }
                mbar_arrive_expect_tx(full, boxes * kBox * 2);
                for (int h = 0; h < L.hs + L.ht; ++h) {
                    const bool stu = h < L.hs;
                    const int e0 = (stu ? h : h - L.hs) * L.positions + p0;
                    const int a0 = e0 & ~7, need = (e0 & 7) + len;
                    for (int x = 0; x < need; x += kBox)
                        tma_load_2d(dst + (h * p.pc + x) * 2, stu ? &L.map_s : &L.map_t, full, a0 + x, b);
                }
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ---------------------------------------------------------------- compute: position pairs
        const int ct = threadIdx.x - 32;
        __shared__ int row_off[2][64];          // per tile parity: element offset of each head row inside the stage (h*pc + m_h)
        __shared__ int row_par[2][64];          // store parity of student row h: 1 if (h*P + p0) is odd
        int it = 0;
        int stage = 0;
        uint32_t phase = 0;
        int k = 0;
        double cur = 0.0;
        int cur_term = -1;
        unsigned int written = 0;               // thread ct == 0: terms whose partial this CTA has written
        __shared__ double warp_part[kComputeWarps];
        auto flush = [&]() {
            double v = warp_sum(cur);
            if (lane == 0) warp_part[warp - 1] = v;
            asm volatile("bar.sync 1, %0;" ::"n"(kComputeThreads) : "memory");
            if (ct == 0) {
                double tot = 0.0;
                for (int w = 0; w < kComputeWarps; ++w) tot += warp_part[w];
                p.partials[(size_t)cur_term * p.partial_stride + blockIdx.x] = tot;
                written |= 1u << cur_term;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kComputeThreads) : "memory");
            cur = 0.0;
        };
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            while (k + 1 < p.n_layers && tile >= p.layer[k + 1].tile_begin) ++k;
            const AttnTmaLayer& L = p.layer[k];
            if (L.term != cur_term) {
                if (cur_term >= 0) flush();
                cur_term = L.term;
            }
            const long long lt = tile - L.tile_begin;
            const int b = (int)(lt / L.chunks), c = (int)(lt % L.chunks);
            const int p0 = c * span;
            const int len = min(span, L.positions - p0);
            const int hs = L.hs, ht = L.ht;
            int* roff = row_off[it & 1];
            int* rpar = row_par[it & 1];
            if (ct < hs + ht) {
                const int e0 = (ct < hs ? ct : ct - hs) * L.positions + p0;
                roff[ct] = ct * p.pc + (e0 & 7);
                rpar[ct] = e0 & 1;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kComputeThreads) : "memory");
            mbar_wait(bar_full + 8 * stage, phase);
            const T* in = reinterpret_cast<const T*>(smem_gen + stage * in_stage_bytes);
            T* gbase = L.grad ? static_cast<T*>(L.grad) + (size_t)b * hs * L.positions + p0 : nullptr;
            float acc = 0.f;
            for (int x = 2 * ct; x < len; x += 2 * kComputeThreads) {
                const bool two = x + 1 < len;
                float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;
#pragma unroll 4
                for (int h = 0; h < hs; ++h) {
                    const T* row = in + roff[h] + x;
                    float a, bb;
                    if (rpar[h] == 0) {
                        unpack2<T>(*reinterpret_cast<const uint32_t*>(row), a, bb);
                    } else {
                        a = Elem<T>::to_f(row[0]);
                        bb = Elem<T>::to_f(row[1]);
                    }
                    s0 += a;
                    s1 += bb;
                }
#pragma unroll 4
                for (int h = hs; h < hs + ht; ++h) {
                    const T* row = in + roff[h] + x;
                    float a, bb;
                    if (rpar[h] == 0) {
                        unpack2<T>(*reinterpret_cast<const uint32_t*>(row), a, bb);
                    } else {
                        a = Elem<T>::to_f(row[0]);
                        bb = Elem<T>::to_f(row[1]);
                    }
                    t0 += a;
                    t1 += bb;
                }
                float g[2];
                const float sm[2] = {s0 * L.inv_hs, s1 * L.inv_hs}, tm[2] = {t0 * L.inv_ht, t1 * L.inv_ht};
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float v;
                    if (p.mode == 1) {
                        const float d = sm[e] - tm[e];
                        v = d * d;
                        g[e] = d * L.grad_coef;
                    } else {
                        const float tl = (tm[e] == 0.f) ? 0.f : tm[e] * logf(tm[e]);      // xlogy(t, t)
                        v = tl - tm[e] * logf(sm[e]);                                       // 0 * -inf -> NaN like the reference
                        g[e] = -L.grad_coef * (tm[e] / sm[e]);
                    }
                    if (e == 0 || two) acc += v;
                }
                if (gbase) {
                    const uint32_t packed = pack2<T>(g[0], g[1]);
#pragma unroll 4
                    for (int h = 0; h < hs; ++h) {
                        T* dstp = gbase + (size_t)h * L.positions + x;
                        if (two && rpar[h] == 0) {
                            *reinterpret_cast<uint32_t*>(dstp) = packed;
                        } else {
                            dstp[0] = Elem<T>::from_f(g[0]);
                            if (two) dstp[1] = Elem<T>::from_f(g[1]);
                        }
                    }
                }
            }
            cur += (double)acc * (double)L.val_coef;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * stage);      // input stage consumed
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (cur_term >= 0) flush();
        if (ct == 0) {
            // every CTA of the (fixed-size) grid owns one partial per term of this launch: zero the ones it never met
            for (int l = 0; l < p.n_layers; ++l)
                if (!(written & (1u << p.layer[l].term))) {
                    p.partials[(size_t)p.layer[l].term * p.partial_stride + blockIdx.x] = 0.0;
                    written |= 1u << p.layer[l].term;
                }
        }
    }
}

}  // namespace dcb

// 2-D map over [batch, heads * positions] of 16-bit elements, box {256, 1}, no swizzle
static int encode_flat_map(CUtensorMap* map, const void* base, uint64_t batch, uint64_t inner) {
    using namespace dcb;
    static void* fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return fail("cuTensorMapEncodeTiled is not available from the CUDA driver");
    }
    using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    const cuuint64_t gdim[2] = {inner, batch};
    const cuuint64_t gstride[1] = {inner * 2};
    const cuuint32_t box[2] = {dcb::atma::kBox, 1};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeFn>(fn)(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
                                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (flat map) failed with CUresult %d", (int)r);
    return 0;
}

// 1 if the TMA-staged kernel can take this layer: 16-bit elements, 16-byte aligned samples, enough positions to fill a box
extern "C" int dcb_attn_tma_supported(int dtype, int64_t stu_heads, int64_t tea_heads, int64_t positions) {
    return (dtype == DCB_BF16 || dtype == DCB_F16) && positions >= 128 && (stu_heads * positions) % 8 == 0 &&
           (tea_heads * positions) % 8 == 0 && stu_heads + tea_heads <= 64 && stu_heads * positions < (1ll << 31) &&
           tea_heads * positions < (1ll << 31);
}
extern "C" int dcb_attn_tma_grid(void) { return dcb::kNumSMs; }

// Same contract as dcb_attn_kl_fwd_bwd (mode 0) / the attention-MSE kind of dcb_tower_fwd_bwd (mode 1), restricted to
// layers accepted by dcb_attn_tma_supported.  partials: this call fills partials[term[l]][0 .. dcb_attn_tma_grid()) with
// row stride partial_stride (doubles); terms must be grouped (non-decreasing).  grad dtype == input dtype.
extern "C" int dcb_attn_tma_fwd_bwd(int n_layers, int mode, const int32_t* term, const void* const* stu, const void* const* tea,
                                    void* const* grad_stu, const int64_t* batch, const int32_t* stu_heads,
                                    const int32_t* tea_heads, const int64_t* positions, const int32_t* divisor,
                                    const float* grad_scale, int dtype, double* partials, int partial_stride, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(n_layers >= 1 && n_layers <= atma::kMaxLayers, "n_layers=%d out of range [1,%d]", n_layers, atma::kMaxLayers);
    DCB_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (KL) or 1 (MSE)");
    DCB_REQUIRE(partials && partial_stride >= dcb_attn_tma_grid(), "partials / partial_stride too small");
    AttnTmaParams p{};
    p.n_layers = n_layers;
    p.mode = mode;
    p.partials = partials;
    p.partial_stride = partial_stride;
    int max_rows = 0;
    int64_t min_pos = 1ll << 40;
    for (int k = 0; k < n_layers; ++k) {
        DCB_REQUIRE(stu[k] && tea[k] && batch[k] >= 1 && divisor[k] >= 1, "layer %d: bad arguments", k);
        DCB_REQUIRE(dcb_attn_tma_supported(dtype, stu_heads[k], tea_heads[k], positions[k]), "layer %d not supported by the TMA path", k);
        DCB_REQUIRE(((uintptr_t)stu[k] | (uintptr_t)tea[k]) % 16 == 0 && (uintptr_t)(grad_stu ? grad_stu[k] : nullptr) % 4 == 0,
                    "layer %d: inputs must be 16-byte aligned, gradients 4-byte aligned", k);
        DCB_REQUIRE(k == 0 || term[k] >= term[k - 1], "terms must be grouped");
        if (stu_heads[k] + tea_heads[k] > max_rows) max_rows = stu_heads[k] + tea_heads[k];
        if (positions[k] < min_pos) min_pos = positions[k];
    }
    // shared-memory row length per head (multiple of 256; 8 elements of it are alignment slack): as large as shared
    // memory allows and the maps need (fewer, longer bursts per head row)
    int pc = 1024;
    while (pc > atma::kBox && (pc - atma::kBox >= min_pos + 8 || atma::kStages * max_rows * pc * 2 > 200 * 1024)) pc -= atma::kBox;
    p.pc = pc;
    p.max_rows = max_rows;
    long long tiles = 0;
    for (int k = 0; k < n_layers; ++k) {
        AttnTmaLayer& L = p.layer[k];
        if (encode_flat_map(&L.map_s, stu[k], batch[k], (uint64_t)stu_heads[k] * positions[k])) return 1;
        if (encode_flat_map(&L.map_t, tea[k], batch[k], (uint64_t)tea_heads[k] * positions[k])) return 1;
        L.grad = grad_stu ? grad_stu[k] : nullptr;
        L.batch = (int)batch[k];
        L.hs = stu_heads[k];
        L.ht = tea_heads[k];
        L.positions = (int)positions[k];
        L.chunks = (int)((positions[k] + (pc - 8) - 1) / (pc - 8));
        L.tile_begin = tiles;
        tiles += (long long)batch[k] * L.chunks;
        L.inv_hs = 1.0f / (float)stu_heads[k];
        L.inv_ht = 1.0f / (float)tea_heads[k];
        L.term = term[k];
        if (mode == 0) {
            L.val_coef = (float)(1.0 / (double)divisor[k]);
            L.grad_coef = (float)((double)grad_scale[k] / ((double)stu_heads[k] * (double)divisor[k]));
        } else {
            const double denom = (double)batch[k] * (double)positions[k] * (double)divisor[k];
            L.val_coef = (float)(1.0 / denom);
            L.grad_coef = (float)(2.0 * (double)grad_scale[k] / (denom * (double)stu_heads[k]));
        }
    }
    p.total_tiles = tiles;
    const int smem = 128 + atma::kStages * max_rows * pc * 2 + 8 * 2 * atma::kStages + 64;
    const long long grid = dcb_attn_tma_grid();      // fixed: the reduction reads exactly this many partials per term
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static int max_set_bf = 0, max_set_h = 0;
    if (dtype == DCB_BF16) {
        if (smem > max_set_bf) {
            DCB_CUDA_OK(cudaFuncSetAttribute(attn_tma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            max_set_bf = smem;
        }
        attn_tma_kernel<__nv_bfloat16><<<(unsigned)grid, atma::kThreads, smem, st>>>(p);
    } else {
        if (smem > max_set_h) {
            DCB_CUDA_OK(cudaFuncSetAttribute(attn_tma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            max_set_h = smem;
        }
        attn_tma_kernel<__half><<<(unsigned)grid, atma::kThreads, smem, st>>>(p);
    }
    DCB_CUDA_OK(cudaGetLastError());
    return 0;
}
