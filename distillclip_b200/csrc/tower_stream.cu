// All streaming losses of one tower in ONE launch: hidden-state / embedding MSE, attention-map KL, attention-map /
// attention-score MSE, output L1 and output cosine, over every layer; values and student gradients in a single pass
// over HBM, plus the weighting of model/_loss.py:195-200.
//
// Replaces LossCalculator.cal_one_tower_loss's python loops over layers and loss names (reference model/_loss.py:155-202,
// loss_component/attention_probs_kl.py:10-22, attention_probs_mse.py:10-22, attention_score_mse.py:10-22,
// hidden_mse.py:9-17, embed_mse.py:9-10, out_l1.py:9-10, out_cos.py:10-11) and their autograd backward.
//
// A persistent grid (4 CTAs per SM) walks a unified tile list (MSE tiles, then attention tiles).  Every CTA keeps one
// double accumulator per loss term and writes it to partials[term][cta]; the last CTA to finish (atomic ticket) reduces
// the partials in a fixed order -- so values are run-to-run identical -- applies `scale` and `percent`, and resets the
// ticket for the next launch.  No separate finalize launch, no float atomics.
#include "stream_tiles.cuh"

namespace dcb {

constexpr int kTowerMaxSeg = 40;
constexpr int kTowerMaxTerms = 8;

struct TowerSeg {
    const void* s;
    const void* t;
    void* g;
    long long n;           // MSE: elements; ATTN: groups (batch * positions / VEC)
    long long tile_begin;
    long long positions;   // ATTN
    long long groups_per_b;
    long long total_s, total_t;   // ATTN, aligned mode: elements of the student / teacher tensor
    int kind;              // 0 = MSE, 1 = attention KL, 2 = L1, 3 = cosine rows, 4 = attention (head-mean) MSE
    int term;
    int hs, ht;
    int aligned;           // MSE: 16-byte aligned pointers -> vector path
    float inv_hs, inv_ht;
    float val_coef;
    float grad_coef;
    // regrad mode (backward): recompute this segment's gradients with the TRUE upstream gradient unless it equals what the
    // forward pass assumed: skip iff *up == expected * fwd_mult, else coefficient = grad_coef * (*up * up_mult)
    const float* up;
    float up_mult, expected;
};
struct TowerParams {
    int n_seg, n_terms;
    long long total_tiles;
    float scale[kTowerMaxTerms], percent[kTowerMaxTerms];
    double* partials;      // [n_terms][partial_stride]
    int partial_stride;
    unsigned int ext_mask; // terms whose partials were written by an earlier kernel (attn_tma.cu): ext_count entries each
    int ext_count;
    unsigned int* ticket;
    float* out;            // [n_terms + 1]
    const float* fwd_mult; // optional device scalar multiplied into every gradient written by the forward pass (the AMP
                           // GradScaler's scale tensor: gradients are rounded once, at the scaled magnitude)
    int stage_max_rows;    // staged attention tiles: max hs + ht over the attention segments
    int stage_w, stage_row_bytes;   // staged attention tiles (AVEC == -1): positions per tile, staged bytes per head row
    int regrad;            // 1 = gradients only (see TowerSeg::up): no values, no partials, no ticket
    TowerSeg seg[kTowerMaxSeg];
};

// GPT = attention groups per thread (stream_tiles.cuh: attn_tile_multi): 2 keeps twice the loads in flight (image stage
// attention tiles 0.76 -> 0.86 of HBM, scripts/attn_gpt_probe.py) at ~90 registers, hence 2 resident CTAs instead of 4.
// AVEC = 0: attention tiles on aligned 16-byte accesses with in-register realignment (attn_tile_aligned), for 16-bit maps
// whose head rows are off the 16-byte grid.  AVEC = -1: STAGED attention tiles (cp.async into shared memory, next tile in
// flight while the current one is computed; stream_tiles.cuh).
extern __shared__ __align__(16) unsigned char tower_dyn_smem[];

// resident CTAs per SM the register budget is compiled for: per-thread loads in flight = 2 heads' worth x AVEC x GPT
constexpr int tower_min_blocks(int avec, int gpt) {
    return avec == 0 ? 2 : (gpt == 1 ? 4 : (avec * gpt <= 2 ? 4 : (avec * gpt <= 4 ? 3 : 2)));
}

template <typename T, typename G, int AVEC, int AH, int GPT>
__global__ void __launch_bounds__(kStreamThreads, tower_min_blocks(AVEC, GPT)) tower_stream_kernel(const __grid_constant__ TowerParams p) {
    constexpr int MVEC = Elem<T>::kPer16B;
    constexpr long long kMseTile = (long long)kStreamThreads * kMseUnroll * MVEC;
    constexpr long long kMseTileScalar = (long long)kStreamThreads * kMseUnroll;
    const int tid = threadIdx.x;
    __shared__ float seg_gc[kTowerMaxSeg];
    __shared__ unsigned char seg_skip[kTowerMaxSeg];
    __shared__ int any_work;
    if (tid == 0) any_work = 0;
    __syncthreads();
    if (tid < p.n_seg) {
        const float fm = p.fwd_mult ? __ldg(p.fwd_mult) : 1.f;
        float gc = p.seg[tid].grad_coef * fm;
        bool skip = false;
        if (p.regrad) {
            const float up = __ldg(p.seg[tid].up);
            skip = up == p.seg[tid].expected * fm || p.seg[tid].g == nullptr;
            gc = p.seg[tid].grad_coef * (up * p.seg[tid].up_mult);
        }
        seg_gc[tid] = gc;
        seg_skip[tid] = skip ? 1 : 0;
        if (!skip) atomicOr(&any_work, 1);
    }
    __syncthreads();
    if (p.regrad && !any_work) return;             // the common case in backward: nothing to do, no HBM traffic
    // Tiles are ordered by segment and segments by term, so a CTA meets the terms in non-decreasing order: one running
    // accumulator, flushed (block reduce -> partials[term][cta]) whenever the term changes.
    double cur = 0.0;
    int cur_term = -1;
    unsigned int written = 0;          // thread 0: bit q set once partials[q][cta] has been written
    auto flush = [&]() {
        const double total = block_sum(cur);
        if (tid == 0) {
            p.partials[(size_t)cur_term * p.partial_stride + blockIdx.x] = total;
            written |= 1u << cur_term;
        }
        __syncthreads();
        cur = 0.0;
    };
    int k = 0;
    [[maybe_unused]] int stage_buf = 0;                 // staged mode: buffer holding the current attention tile
    [[maybe_unused]] bool stage_ready = false;          //   ... whose copies are already in flight
    [[maybe_unused]] __shared__ short stage_row_off[2][64];
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        while (k + 1 < p.n_seg && tile >= p.seg[k + 1].tile_begin) ++k;
        const TowerSeg& sg = p.seg[k];
        if (p.regrad && seg_skip[k]) continue;
        const float gcoef = seg_gc[k];
        if (!p.regrad && sg.term != cur_term) {        // block-uniform
            if (cur_term >= 0) flush();
            cur_term = sg.term;
        }
        const long long lt = tile - sg.tile_begin;
        float acc;
        if (sg.kind == 0 || sg.kind == 2) {
            G* g = static_cast<G*>(sg.g);
            const long long base = lt * (sg.aligned ? kMseTile : kMseTileScalar);
            const T* sp = static_cast<const T*>(sg.s) + base;
            const T* tp = static_cast<const T*>(sg.t) + base;
            G* gp = g ? g + base : nullptr;
            if (sg.kind == 0)
                acc = sg.aligned ? mse_tile<T, G, MVEC, false>(sp, tp, gp, sg.n - base, gcoef, tid)
                                 : mse_tile<T, G, 1, false>(sp, tp, gp, sg.n - base, gcoef, tid);
            else
                acc = sg.aligned ? mse_tile<T, G, MVEC, true>(sp, tp, gp, sg.n - base, gcoef, tid)
                                 : mse_tile<T, G, 1, true>(sp, tp, gp, sg.n - base, gcoef, tid);
        } else if (sg.kind == 3) {
            acc = cos_row_tile<T, G>(static_cast<const T*>(sg.s), static_cast<const T*>(sg.t), static_cast<G*>(sg.g), sg.n,
                                     (int)sg.positions, lt * (kStreamThreads / 32) + (tid >> 5), gcoef, tid & 31);
        } else {
            AttnShape sh{sg.n, sg.groups_per_b, sg.positions, sg.hs, sg.ht, sg.inv_hs, sg.inv_ht};
            if constexpr (AVEC == -1) {
                const size_t buf_bytes = (size_t)(p.stage_max_rows) * p.stage_row_bytes;
                auto staged = [&](const TowerSeg& q) {
                    return AttnStaged{q.positions, q.groups_per_b, q.total_s, q.total_t, p.stage_w, p.stage_row_bytes, q.hs, q.ht, q.inv_hs, q.inv_ht};
                };
                const AttnStaged cur_a = staged(sg);
                if (!stage_ready)
                    attn_stage_issue<T>(static_cast<const T*>(sg.s), static_cast<const T*>(sg.t), cur_a, lt,
                                        tower_dyn_smem + stage_buf * buf_bytes, stage_row_off[stage_buf], tid);
                // the CTA's next tile: start its copies now if it is an attention tile too
                const long long nt = tile + gridDim.x;
                int kn = k;
                while (kn + 1 < p.n_seg && nt >= p.seg[kn + 1].tile_begin) ++kn;
                const TowerSeg& sn = p.seg[kn];
                const bool next_staged = nt < p.total_tiles && (sn.kind == 1 || sn.kind == 4) && !(p.regrad && seg_skip[kn]);
                if (next_staged) {
                    attn_stage_issue<T>(static_cast<const T*>(sn.s), static_cast<const T*>(sn.t), staged(sn), nt - sn.tile_begin,
                                        tower_dyn_smem + (stage_buf ^ 1) * buf_bytes, stage_row_off[stage_buf ^ 1], tid);
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                } else {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                }
                __syncthreads();
                if (sg.kind == 1)
                    acc = attn_stage_compute<T, G, AH, false>(tower_dyn_smem + stage_buf * buf_bytes, stage_row_off[stage_buf],
                                                              static_cast<G*>(sg.g), cur_a, lt, gcoef, tid);
                else
                    acc = attn_stage_compute<T, G, AH, true>(tower_dyn_smem + stage_buf * buf_bytes, stage_row_off[stage_buf],
                                                             static_cast<G*>(sg.g), cur_a, lt, gcoef, tid);
                __syncthreads();                         // everyone is done with this buffer before it is refilled
                stage_ready = next_staged;
                stage_buf ^= 1;
            } else if constexpr (AVEC == 0) {
                __shared__ uint4 xchg[kStreamThreads];
                acc = 0.f;
                if constexpr (sizeof(T) == 2 && sizeof(G) == 2) {
                    AttnShape8 sh8{sg.n, sg.groups_per_b, sg.positions, sg.total_s, sg.total_t, sg.hs, sg.ht, sg.inv_hs, sg.inv_ht};
                    if (sg.kind == 1)
                        acc = attn_tile_aligned<T, G, AH, false>(static_cast<const T*>(sg.s), static_cast<const T*>(sg.t),
                                                                 static_cast<G*>(sg.g), sh8, lt * kStreamThreads + tid, gcoef, xchg, tid);
                    else
                        acc = attn_tile_aligned<T, G, AH, true>(static_cast<const T*>(sg.s), static_cast<const T*>(sg.t),
                                                                static_cast<G*>(sg.g), sh8, lt * kStreamThreads + tid, gcoef, xchg, tid);
                }
            } else if constexpr (GPT > 1 && AH > 0) {
                const long long g0 = lt * (kStreamThreads * GPT) + tid;
                if (sg.kind == 1)
                    acc = attn_tile_multi<T, G, AVEC, AH, false, GPT>(static_cast<const T*>(sg.s), static_cast<const T*>(sg.t),
                                                                      static_cast<G*>(sg.g), sh, g0, kStreamThreads, gcoef);
                else
                    acc = attn_tile_multi<T, G, AVEC, AH, true, GPT>(static_cast<const T*>(sg.s), static_cast<const T*>(sg.t),
                                                                     static_cast<G*>(sg.g), sh, g0, kStreamThreads, gcoef);
            } else {
                if (sg.kind == 1)
                    acc = attn_tile<T, G, AVEC, AH, false>(static_cast<const T*>(sg.s), static_cast<const T*>(sg.t),
                                                           static_cast<G*>(sg.g), sh, lt * kStreamThreads + tid, gcoef);
                else
                    acc = attn_tile<T, G, AVEC, AH, true>(static_cast<const T*>(sg.s), static_cast<const T*>(sg.t),
                                                          static_cast<G*>(sg.g), sh, lt * kStreamThreads + tid, gcoef);
            }
        }
        cur += (double)acc * (double)sg.val_coef;
    }
    if (p.regrad) return;
    if (cur_term >= 0) flush();
    __shared__ bool is_last;
    if (tid == 0)
        for (int q = 0; q < p.n_terms; ++q)
            if (!((written | p.ext_mask) & (1u << q))) p.partials[(size_t)q * p.partial_stride + blockIdx.x] = 0.0;
    if (tid == 0) {
        __threadfence();
        const unsigned int ticket = atomicAdd(p.ticket, 1u);
        is_last = ticket == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    __shared__ double term_sum[kTowerMaxTerms];
    for (int q = 0; q < p.n_terms; ++q) {
        const volatile double* src = p.partials + (size_t)q * p.partial_stride;
        const unsigned int cnt = (p.ext_mask & (1u << q)) ? (unsigned int)p.ext_count : gridDim.x;
        double v = 0.0;
        for (unsigned int i = tid; i < cnt; i += kStreamThreads) v += src[i];     // fixed assignment, fixed tree
        v = block_sum(v);
        if (tid == 0) term_sum[q] = v;
        __syncthreads();
    }
    if (tid == 0) {
        float total = 0.f;
        for (int q = 0; q < p.n_terms; ++q) {
            // same rounding points as the reference: fp32 value, * scale, * percent, += in fp32 (_loss.py:199-200)
            const float res = (float)term_sum[q] * p.scale[q];
            p.out[q] = res;
            total += res * p.percent[q];
        }
        p.out[p.n_terms] = total;
        *p.ticket = 0u;          // ready for the next launch (stream order / graph replay)
    }
}

// attention groups per thread for this launch (0 = no attention segment with a compile-time head count)
static int tower_gpt(int avec, int common_h) {
    if (!(common_h == 12 || common_h == 8) || avec > 4) return 1;
    // measured (scripts/attn_gpt_probe.py, round 2): with 2-byte loads (N = 77) 2 groups per thread at 4 resident CTAs lift the text
    // tower 0.72 -> 0.77 of HBM (4 groups: 0.73); with 8-byte loads (N = 50) the tower is 0.89-0.92 either way -> 1 group
    if (const char* e = getenv("DCB_ATTN_GPT")) return (atoi(e) == 2 || (atoi(e) == 4 && avec == 1)) ? atoi(e) : 1;
    if (avec == 1) return 2;
    return 1;
}

template <typename T, typename G, int AH>
static int launch_tower_staged(const TowerParams& p, unsigned grid, size_t smem, cudaStream_t st) {
    static size_t max_set = 0;
    if (smem > 48 * 1024 && smem > max_set) {
        DCB_CUDA_OK(cudaFuncSetAttribute(tower_stream_kernel<T, G, -1, AH, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        max_set = smem;
    }
    tower_stream_kernel<T, G, -1, AH, 1><<<grid, kStreamThreads, smem, st>>>(p);
    DCB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T, typename G, int AVEC>
static int launch_tower_h(const TowerParams& p, int common_h, unsigned grid, cudaStream_t st, size_t smem = 0) {
    bool done = false;
    if constexpr (AVEC == -1) {
        if (common_h == 12) return launch_tower_staged<T, G, 12>(p, grid, smem, st);
        if (common_h == 8) return launch_tower_staged<T, G, 8>(p, grid, smem, st);
        return launch_tower_staged<T, G, 0>(p, grid, smem, st);
    } else if constexpr (AVEC == 0) {
        if (common_h == 12) tower_stream_kernel<T, G, 0, 12, 1><<<grid, kStreamThreads, 0, st>>>(p);
        else if (common_h == 8) tower_stream_kernel<T, G, 0, 8, 1><<<grid, kStreamThreads, 0, st>>>(p);
        else tower_stream_kernel<T, G, 0, 0, 1><<<grid, kStreamThreads, 0, st>>>(p);
        DCB_CUDA_OK(cudaGetLastError());
        return 0;
    } else if constexpr (AVEC <= 4) {          // head loops fully unrolled for the two head counts of the BASELINE configs
        const int gpt = tower_gpt(AVEC, common_h);
#define DCB_TOWER_CASE(HH)                                                                                      \
        if (common_h == HH) {                                                                                      \
            bool g4 = false;                                                                                       \
            if constexpr (AVEC == 1) {                                                                             \
                if (gpt == 4) {                                                                                    \
                    tower_stream_kernel<T, G, AVEC, HH, 4><<<grid, kStreamThreads, 0, st>>>(p);                    \
                    g4 = true;                                                                                     \
                }                                                                                                  \
            }                                                                                                      \
            if (!g4) {                                                                                             \
                if (gpt == 2) tower_stream_kernel<T, G, AVEC, HH, 2><<<grid, kStreamThreads, 0, st>>>(p);          \
                else tower_stream_kernel<T, G, AVEC, HH, 1><<<grid, kStreamThreads, 0, st>>>(p);                   \
            }                                                                                                      \
            done = true;                                                                                           \
        }
        DCB_TOWER_CASE(12)
        DCB_TOWER_CASE(8)
#undef DCB_TOWER_CASE
    }
    if constexpr (AVEC > 0) {
        if (!done) tower_stream_kernel<T, G, AVEC, 0, 1><<<grid, kStreamThreads, 0, st>>>(p);
    }
    DCB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace dcb

extern "C" int dcb_tower_grid(void) { return dcb::kNumSMs * 4; }

extern "C" int dcb_tower_fwd_bwd(int n_seg, const int32_t* kind, const int32_t* term, const void* const* stu,
                                 const void* const* tea, void* const* grad_stu, const int64_t* numel,
                                 const int64_t* batch, const int32_t* stu_heads, const int32_t* tea_heads,
                                 const int64_t* positions, const int32_t* divisor, const float* grad_scale, int n_terms,
                                 const float* scale, const float* percent, int in_dtype, int grad_dtype,
                                 double* partials, int partial_stride, uint32_t ext_mask, int ext_count,
                                 uint32_t* ticket, float* out, const float* fwd_mult, const float* const* upstream,
                                 const float* up_mult, const float* expected, void* stream) {
    using namespace dcb;
    const bool regrad = upstream != nullptr;
    DCB_REQUIRE(n_seg >= 0 && n_seg <= kTowerMaxSeg, "n_seg=%d out of range [0,%d]", n_seg, kTowerMaxSeg);
    DCB_REQUIRE(n_terms >= 1 && n_terms <= kTowerMaxTerms, "n_terms=%d out of range [1,%d]", n_terms, kTowerMaxTerms);
    if (regrad) {
        DCB_REQUIRE(up_mult && expected, "regrad mode needs up_mult and expected per segment");
    } else {
        DCB_REQUIRE(partial_stride >= dcb_tower_grid() && ext_count <= partial_stride, "partial_stride too small");
        DCB_REQUIRE(partials && ticket && out, "partials / ticket / out must not be NULL");
    }
    TowerParams p{};
    p.n_seg = n_seg;
    p.n_terms = n_terms;
    p.partials = partials;
    p.partial_stride = partial_stride;
    p.ext_mask = ext_mask;
    p.ext_count = ext_count;
    p.ticket = ticket;
    p.out = out;
    p.fwd_mult = fwd_mult;
    p.regrad = regrad ? 1 : 0;
    for (int q = 0; q < n_terms; ++q) {
        p.scale[q] = scale ? scale[q] : 1.f;
        p.percent[q] = percent ? percent[q] : 0.f;
    }
    const int isz = dtype_size(in_dtype), gsz = dtype_size(grad_dtype);
    int avec = 16 / isz, common_h = -1;
    for (int k = 0; k < n_seg; ++k) {
        DCB_REQUIRE(stu[k] && tea[k], "segment %d: NULL input", k);
        DCB_REQUIRE(term[k] >= 0 && term[k] < n_terms && divisor[k] >= 1, "segment %d: bad term / divisor", k);
        TowerSeg& sg = p.seg[k];
        sg.s = stu[k];
        sg.t = tea[k];
        sg.g = grad_stu ? grad_stu[k] : nullptr;
        sg.kind = kind[k];
        sg.term = term[k];
        if (regrad) {
            DCB_REQUIRE(upstream[k] || !sg.g, "segment %d: regrad needs the upstream gradient pointer", k);
            sg.up = upstream[k];
            sg.up_mult = up_mult[k];
            sg.expected = expected[k];
        }
        if (kind[k] == 0 || kind[k] == 2) {
            DCB_REQUIRE(numel[k] >= 1, "segment %d: numel must be >= 1", k);
            sg.n = numel[k];
            const double denom = (double)numel[k] * (double)divisor[k];
            sg.val_coef = (float)(1.0 / denom);
            sg.grad_coef = (float)((kind[k] == 0 ? 2.0 : 1.0) * (double)grad_scale[k] / denom);
            sg.aligned = (((uintptr_t)stu[k] | (uintptr_t)tea[k] | (uintptr_t)sg.g) % 16 == 0) ? 1 : 0;
        } else if (kind[k] == 3) {
            // cosine rows: batch[k] rows of positions[k] elements; value = mean over rows of (1 - cos)
            DCB_REQUIRE(batch[k] >= 1 && positions[k] >= 1 && positions[k] < (1ll << 31), "segment %d: bad shape", k);
            sg.n = batch[k];
            sg.positions = positions[k];
            const double denom = (double)batch[k] * (double)divisor[k];
            sg.val_coef = (float)(1.0 / denom);
            sg.grad_coef = (float)((double)grad_scale[k] / denom);
        } else if (kind[k] == 1 || kind[k] == 4) {
            DCB_REQUIRE(batch[k] >= 1 && positions[k] >= 1 && stu_heads[k] >= 1 && tea_heads[k] >= 1, "segment %d: bad shape", k);
            sg.n = batch[k];
            sg.positions = positions[k];
            sg.hs = stu_heads[k];
            sg.ht = tea_heads[k];
            sg.inv_hs = 1.0f / (float)stu_heads[k];
            sg.inv_ht = 1.0f / (float)tea_heads[k];
            if (kind[k] == 1) {
                sg.val_coef = (float)(1.0 / (double)divisor[k]);
                sg.grad_coef = (float)((double)grad_scale[k] / ((double)stu_heads[k] * (double)divisor[k]));
            } else {      // MSE(mean) of the head means: mean over batch * positions elements
                const double denom = (double)batch[k] * (double)positions[k] * (double)divisor[k];
                sg.val_coef = (float)(1.0 / denom);
                sg.grad_coef = (float)(2.0 * (double)grad_scale[k] / (denom * (double)stu_heads[k]));
            }
            while (avec > 1 && (positions[k] % avec != 0 || ((uintptr_t)stu[k] | (uintptr_t)tea[k]) % (avec * isz) != 0 ||
                                (sg.g && (uintptr_t)sg.g % (avec * gsz < 16 ? avec * gsz : 16) != 0)))
                avec >>= 1;
            if (common_h == -1) common_h = stu_heads[k];
            if (stu_heads[k] != common_h || tea_heads[k] != common_h) common_h = 0;
        } else {
            return fail("segment %d: unknown kind %d", k, kind[k]);
        }
    }
    // head rows off the 16-byte grid (avec < 8 for 16-bit maps): aligned vectors + realignment when every pointer allows it
    // (default OFF -- a measured negative result: the realignment network costs more issue slots than the narrow loads cost
    // memory efficiency; DCB_ATTN_ALIGNED=1 selects it, tests keep it bit-identical to the default path)
    bool aligned_mode = common_h != -1 && isz == 2 && gsz == 2 && avec < 8 && getenv("DCB_ATTN_ALIGNED") && !getenv("DCB_ATTN_NO_ALIGNED");
    for (int k = 0; k < n_seg && aligned_mode; ++k)
        if (p.seg[k].kind == 1 || p.seg[k].kind == 4)
            aligned_mode = p.seg[k].g && (((uintptr_t)p.seg[k].s | (uintptr_t)p.seg[k].t | (uintptr_t)p.seg[k].g) % 16 == 0);
    // STAGED mode: head rows off the 16-byte grid (per-thread vectors narrower than 16 bytes) -> cp.async staging through shared
    // memory with the next tile in flight (stream_tiles.cuh).  Needs 16-byte aligned map bases and <= 64 head rows per tile.
    // (default OFF -- the second measured negative result on these tiles: bit-identical but slower, text stage 0.58 vs 0.72 of HBM,
    // image stage 0.58 vs 0.90: without the load stalls the tile is issue-bound, and the staging adds copies, syncs and index
    // arithmetic on top; DCB_ATTN_STAGED=1 selects it, tests keep it bit-identical to the default path)
    bool staged_mode = common_h != -1 && !aligned_mode && avec * isz < 16 && getenv("DCB_ATTN_STAGED") && !getenv("DCB_ATTN_NO_STAGED");
    int max_rows = 0;
    for (int k = 0; k < n_seg && staged_mode; ++k)
        if (p.seg[k].kind == 1 || p.seg[k].kind == 4) {
            staged_mode = (((uintptr_t)p.seg[k].s | (uintptr_t)p.seg[k].t) % 16 == 0) && p.seg[k].hs + p.seg[k].ht <= 64;
            max_rows = p.seg[k].hs + p.seg[k].ht > max_rows ? p.seg[k].hs + p.seg[k].ht : max_rows;
        }
    size_t stage_smem = 0;
    if (staged_mode) {
        int w = 512;
        if (const char* e = getenv("DCB_ATTN_STAGE_W")) w = atoi(e) >= 256 ? atoi(e) / 256 * 256 : 256;     // profiling only
        while (w > 256 && 2 * (size_t)max_rows * (w * isz + 32) > 56 * 1024) w -= 256;
        p.stage_w = w;
        p.stage_row_bytes = w * isz + 32;
        p.stage_max_rows = max_rows;
        stage_smem = 2 * (size_t)max_rows * p.stage_row_bytes;
        staged_mode = stage_smem <= 100 * 1024;
    }
    const int gpt = (aligned_mode || staged_mode) ? 1 : tower_gpt(avec, common_h);
    long long tiles = 0;
    const long long mse_tile_vec = (long long)kStreamThreads * kMseUnroll * (16 / isz);
    const long long mse_tile_scalar = (long long)kStreamThreads * kMseUnroll;
    for (int k = 0; k < n_seg; ++k) {
        TowerSeg& sg = p.seg[k];
        sg.tile_begin = tiles;
        if (sg.kind == 0 || sg.kind == 2) {
            const long long tl = sg.aligned ? mse_tile_vec : mse_tile_scalar;
            tiles += (sg.n + tl - 1) / tl;
        } else if (sg.kind == 3) {
            tiles += (sg.n + kStreamThreads / 32 - 1) / (kStreamThreads / 32);       // one warp per row
        } else {
            sg.total_s = sg.n * sg.hs * sg.positions;                 // sg.n holds the batch here
            sg.total_t = sg.n * sg.ht * sg.positions;
            if (staged_mode) {                                        // one tile = stage_w positions of one sample
                sg.groups_per_b = (sg.positions + p.stage_w - 1) / p.stage_w;
                sg.n *= sg.groups_per_b;
                tiles += sg.n;
                continue;
            }
            sg.groups_per_b = aligned_mode ? (sg.positions + 7) / 8 : sg.positions / avec;
            sg.n *= sg.groups_per_b;
            tiles += (sg.n + kStreamThreads * gpt - 1) / (kStreamThreads * gpt);
        }
    }
    p.total_tiles = tiles;
    const long long max_grid = (long long)kNumSMs * (staged_mode ? 4 : tower_min_blocks(aligned_mode ? 0 : avec, gpt));     // one persistent CTA per resident slot
    long long grid = tiles < max_grid ? tiles : max_grid;
    if (grid < 1) grid = 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return dispatch_in_grad(in_dtype, grad_dtype, [&](auto tt, auto gg) -> int {
        using T = decltype(tt);
        using G = decltype(gg);
        constexpr int kMax = Elem<T>::kPer16B;
        if (staged_mode) return launch_tower_h<T, G, -1>(p, common_h, (unsigned)grid, st, stage_smem);
        if constexpr (sizeof(T) == 2 && sizeof(G) == 2) {
            if (aligned_mode) return launch_tower_h<T, G, 0>(p, common_h, (unsigned)grid, st);
        }
        if (avec >= kMax) return launch_tower_h<T, G, kMax>(p, common_h, (unsigned)grid, st);
        if (avec == 4) {
            if constexpr (kMax > 4) return launch_tower_h<T, G, 4>(p, common_h, (unsigned)grid, st);
        }
        if (avec == 2) return launch_tower_h<T, G, 2>(p, common_h, (unsigned)grid, st);
        return launch_tower_h<T, G, 1>(p, common_h, (unsigned)grid, st);
    });
}
