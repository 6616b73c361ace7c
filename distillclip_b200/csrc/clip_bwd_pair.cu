// Backward of the fused contrastive / logit-KL losses for one direction, CTA-pair version (cta_group::2).
//
// Same math as clip_bwd.cu (see there for the reference lines and the definition of G_ij), different mapping:
// a cluster of two CTAs on one TPC owns a block of 128 a-side rows, 64 per CTA, and the WHOLE embedding dimension.
// With tcgen05.mma.cta_group::2 and M = 128 each CTA keeps 64 accumulator rows spread over all 128 TMEM lanes
// (lanes 0-63 hold the first half of the N columns, lanes 64-127 the second half), so an accumulator of N columns
// costs N/2 TMEM columns: the [64 x D] gradient accumulator of each CTA (D/2 columns) fits next to the 128-wide S/T
// tiles -- double buffered for D <= 512 (2 x 128 + 256), single buffered for D <= 768 (128 + 384) -- and the logits
// are recomputed exactly once per direction instead of once per 256-column chunk of D.  Each CTA stages its own half of every operand: 64 rows of a, NT/2 rows of b, and 128 of
// the 256 rows of each b_hatT slice; the leader CTA (cluster rank 0) issues all MMAs; completion is multicast to the
// barriers of both CTAs; both CTAs run the epilogue on their own 64 rows.
//
// Warp roles per CTA (384 threads): warp 0 = operand-ring producer, warp 1 = TMEM alloc (+ MMA issuer in the leader),
// warps 2-3 idle, warps 4-11 = epilogue: warp w reads TMEM lanes 32 (w % 4) .. and the
// (w - 4) / 4-th 32-column chunk of its rows, so every scheduler has two epilogue warps to hide MUFU / FMA latency
// (one warp per scheduler issued only every ~4 cycles, ncu r01d).
#include "tc_common.cuh"
#include "clip_shared.cuh"

namespace dcb {

namespace bwdp {
constexpr int kRowsPerCta = 64, kBK = 64, kUmmaK = 16;
constexpr int kThreads = 384;                              // 4 control warps + 8 epilogue warps (2 per scheduler)
constexpr int kTmemCols = 512;
constexpr int kATile = kRowsPerCta * kBK * 2;             // 8 KiB  [64 x 64] 16-bit
constexpr int kSliceRows = 128;                           // rows of one 256-row b_hatT slice held by one CTA
constexpr int kMaxSlices = 3;                             // D <= 768
// NT = logit-tile width (columns per tile), ST = number of S/T accumulator stages in TMEM (2 = double buffered).
// TMEM columns: ST * NT for S/T + D/2 for the gradient accumulator <= 512:  D <= 512 -> (128, 2), D <= 768 -> (128, 1).
template <int NT, int ST> struct Cfg {
    static_assert(NT == 128, "ring entries are sized for 128-wide tiles");
    static constexpr int kBHalf = NT / 2;                                 // b rows staged per CTA
    static constexpr int kBTile = kBHalf * kBK * 2;                       // bytes
    // ONE ring for every operand byte, entries consumed in MMA issue order.  Entry kinds (32 KiB each):
    //   S/T k-chunk:   a_stu [64 x 64], a_tea, b_stu [64 x 64], b_tea
    //   b_hatT slice:  this CTA's [128 d-rows x 128 j] of one 256-row slice (two K sub-tiles of 16 KiB)
    // so the bytes in flight from L2 stay at kStages x 32 KiB in both phases of a tile (the S/T recompute was limited
    // by a 96 KiB ring: 48 B/clk/SM, r01e).
    static constexpr int kStageBytes = 2 * kATile + 2 * kBTile;
    static constexpr int kBtSliceBytes = kSliceRows * NT * 2;
    static_assert(kStageBytes == 32768 && kBtSliceBytes == 32768, "uniform ring entries");
    static constexpr int kStages = 6;
    static constexpr int kGBytes = kRowsPerCta * NT * 2;                  // fp16 G tile of this CTA
    static constexpr int kStCols = NT;                                    // TMEM columns of one S/T stage (S NT/2 + T NT/2)
    static constexpr int kAccCol = ST * kStCols;
    static constexpr int smem_bytes() { return 1024 + kStages * kStageBytes + kGBytes + 2 * 5 * NT * 4 + 256; }
};
}  // namespace bwdp

struct ClipBwdPairParams {
    const float* a_inv_stu;
    const float* b_inv_stu;
    const float* a_inv_tea;
    const float* b_inv_tea;
    const float* coef_row;
    const float* coef_col;
    const float* gmax_row;    // legacy mode (bounds == nullptr): coefficients already carry the upstream gradients,
    const float* gmax_col;    //   gmax_row[0] + gmax_col[0] bounds |G|
    const float* bounds;      // pipeline mode: coef_row / coef_col are UNIT coefficients, multiplied here by the upstream
    ClipUpstream up;          //   gradients read from the device; bounds[6] fixes the fp16 scale (clip_shared.cuh)
    int extra;                // 1: add d(CLIPCosDiff)/dS and d(LogitsMSE)/dS to the tiles (kExtra instantiation)
    int diag0;                //    global column index of local row 0 (the diagonal carries no off-diagonal cos_diff term)
    float inv_batch, inv_pairs;   // 1/B, 1/(B (B - 1))
    int bt_block_cols;        // b_hatT is stored as row blocks [n_blocks][dim][bt_block_cols] (one block per source rank);
    int bt_block_rows;        //   bt_block_rows = rows of one block (= dim).  One block: bt_block_cols >= cols
    float* acc;               // [n_split][rows][dim] fp32
    __half* g_out;            // optional [rows][g_ld] fp16: the scaled gradient tiles G 2^k for clip_gt_gemm_kernel (else nullptr)
    long long g_ld;
    float* dump_s;            // tests only: [rows, cols] student logits as seen by the epilogue
    long long* trace;         // profiling only: clock64 timestamps of cluster 0 (see kTraceSlots), else nullptr
    int rows, cols, dim;
    int slices;               // ceil(dim / 256)
    int n_split, col_tiles;
    float inv_temp;
};

// trace layout: [tile][16] : 0-4 MMA thread (st wait done, st issued, gfull wait done, grad issued, -)
//                            8-14 epilogue warp 4 lane 0 (tile start, scales staged, stfull, tmem loaded, computed, gempty, done)
constexpr int kTraceSlots = 24, kTraceTiles = 64;
#define DCB_TRACE(tile, slot)                                                                        \
    do {                                                                                             \
        if (p.trace && blockIdx.x < 2 && (tile) < kTraceTiles)                                       \
            p.trace[((size_t)(blockIdx.x & 1) * kTraceTiles + (tile)) * kTraceSlots + (slot)] = clock64(); \
    } while (0)

__device__ __forceinline__ float pair_tile_scale(float gmax) {      // must match clip_grad_tile_scale in clip_misc.cu
    if (!(gmax > 0.f) || !isfinite(gmax)) return 1.f;
    int e;
    frexpf(gmax, &e);
    return ldexpf(1.f, 14 - e);
}
__device__ __forceinline__ float ex2p(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool kTeacher, int NT, int ST, bool kExtra = false>
__global__ void __launch_bounds__(bwdp::kThreads, 1)
clip_bwd_pair_kernel(const __grid_constant__ CUtensorMap map_a_stu, const __grid_constant__ CUtensorMap map_b_stu,
                     const __grid_constant__ CUtensorMap map_a_tea, const __grid_constant__ CUtensorMap map_b_tea,
                     const __grid_constant__ CUtensorMap map_bt, const __grid_constant__ ClipBwdPairParams p,
                     const uint32_t idesc_st, const uint32_t idesc_grad) {
    using namespace bwdp;
    using namespace tc;
    using C = Cfg<NT, ST>;
    constexpr int kSub = NT / kBK;                       // K sub-tiles of the gradient GEMM (K = NT)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t ring = smem_base;
    const uint32_t g_smem = ring + C::kStages * C::kStageBytes;
    uint8_t* g_gen = smem_gen + C::kStages * C::kStageBytes;
    float* scale_buf = reinterpret_cast<float*>(smem_gen + C::kStages * C::kStageBytes + C::kGBytes);   // [2][5][NT]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(scale_buf) + 2 * 5 * NT * 4);
    const uint32_t bar_full = smem_u32(bars);                    // [kStages]  leader: TMA bytes of both CTAs
    const uint32_t bar_empty = bar_full + 8 * C::kStages;        // [kStages]  each CTA: slot free (multicast commit)
    const uint32_t bar_stfull = bar_empty + 8 * C::kStages;      // [2] each CTA: S/T accumulators ready (multicast commit)
    const uint32_t bar_stempty = bar_stfull + 16;                // [2] leader: 8 epilogue warps of both CTAs drained them
    const uint32_t bar_gfull = bar_stempty + 16;                 // leader: 8 epilogue warps wrote their G halves
    const uint32_t bar_gempty = bar_gfull + 8;                   // each CTA: gradient MMAs done with the G tile
    const uint32_t bar_accfull = bar_gempty + 8;                 // each CTA: gradient accumulator final
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int rb = cluster_id / p.n_split, sp = cluster_id % p.n_split;
    const int tile_begin = (int)(((long long)sp * p.col_tiles) / p.n_split);
    const int tile_end = (int)(((long long)(sp + 1) * p.col_tiles) / p.n_split);
    const int n_tiles = tile_end - tile_begin;
    const int n_kc = (p.dim + kBK - 1) / kBK;
    const int row0 = rb * 128 + (int)rank * kRowsPerCta;          // first a-side row of this CTA

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_stfull + 8 * s, 1);
            mbar_init(bar_stempty + 8 * s, 16);
        }
        mbar_init(bar_gfull, 16);
        mbar_init(bar_gempty, 1);
        mbar_init(bar_accfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(smem_u32(tmem_slot), kTmemCols);
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                                           // peer barriers initialised, both TMEM allocations done
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    // leader-side barrier addresses as seen from this CTA
    const uint32_t l_full = map_to_cta(bar_full, 0), l_stempty = map_to_cta(bar_stempty, 0);
    const uint32_t l_gfull = map_to_cta(bar_gfull, 0);

    if (warp == 0 || warp == 2) {
        // ---------------------------------------------------------------- operand ring (own halves), MMA issue order.
        // Two issuing threads (warp 0: student boxes + first K sub-tile of a slice, warp 2: teacher boxes + second sub-tile):
        // one thread sustains ~50-65 B/clk of 8 KiB boxes, two ~74 (scripts/probe/tma_probe.cu).  Both walk the same
        // stage/phase sequence; warp 0 alone arms the transaction count.
        const bool second = warp == 2;
        if (elect_one()) {
            tma_prefetch_desc(&map_a_stu);
            tma_prefetch_desc(&map_b_stu);
            tma_prefetch_desc(&map_bt);
            int stage = 0;
            uint32_t phase = 0;
            constexpr uint32_t kStBytesPerCta = (kTeacher ? 2 : 1) * (kATile + C::kBTile);
            for (int tt = 0; tt <= n_tiles; ++tt) {
                if (tt < n_tiles) {                                    // S/T k-chunks of tile tt
                    const int col0 = (tile_begin + tt) * NT + (int)rank * C::kBHalf;
                    for (int kc = 0; kc < n_kc; ++kc) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        const uint32_t dst = ring + stage * C::kStageBytes;
                        if (leader && !second) mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * kStBytesPerCta);
                        const uint32_t full = l_full + 8 * stage;
                        if (!second) {
                            tma_load_2d_pair(dst, &map_a_stu, full, kc * kBK, row0);
                            tma_load_2d_pair(dst + 2 * kATile, &map_b_stu, full, kc * kBK, col0);
                        } else if (kTeacher) {
                            tma_load_2d_pair(dst + kATile, &map_a_tea, full, kc * kBK, row0);
                            tma_load_2d_pair(dst + 2 * kATile + C::kBTile, &map_b_tea, full, kc * kBK, col0);
                        }
                        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                    }
                }
                if (tt > 0) {                                          // b_hatT slices for the gradient GEMM of tile tt-1
                    const int jg = (tile_begin + tt - 1) * NT;
                    const int blk = jg / p.bt_block_cols;                  // tiles never straddle a block (host checks)
                    const int j0 = jg - blk * p.bt_block_cols;
                    const int y0 = blk * p.bt_block_rows;
                    for (int sl = 0; sl < p.slices; ++sl) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        const uint32_t dst = ring + stage * C::kStageBytes;
                        if (leader && !second) mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * C::kBtSliceBytes);
                        const uint32_t full = l_full + 8 * stage;
                        static_assert(kSub == 2, "one K sub-tile of a slice per issuing thread");
                        const int ks = second ? 1 : 0;
                        tma_load_2d_pair(dst + ks * (kSliceRows * kBK * 2), &map_bt, full, j0 + ks * kBK,
                                         y0 + sl * 256 + (int)rank * kSliceRows);
                        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer (leader CTA): the whole warp walks
        // the pipeline (waits are warp-uniform), one elected lane issues
        if (leader) {
            int stage = 0;
            uint32_t phase = 0;
            auto issue_st = [&](int t) {
                const int as = t % ST;
                mbar_wait(bar_stempty + 8 * as, ((t / ST) & 1) ^ 1);
                if (lane == 0) DCB_TRACE(t, 0);
                tc_fence_after_sync();
                const uint32_t acc_s = tmem_base + as * C::kStCols, acc_t = acc_s + NT / 2;
                long long t_wait = 0, t_mark = p.trace ? clock64() : 0;
                for (int kc = 0; kc < n_kc; ++kc) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    if (p.trace) { const long long now = clock64(); t_wait += now - t_mark; }
                    tc_fence_after_sync();
                    const uint32_t src = ring + stage * C::kStageBytes;
                    if (elect_one()) {
                        const uint64_t da_s = umma_desc_k_sw128(src), da_t = umma_desc_k_sw128(src + kATile);
                        const uint64_t db_s = umma_desc_k_sw128(src + 2 * kATile);
                        const uint64_t db_t = umma_desc_k_sw128(src + 2 * kATile + C::kBTile);
#pragma unroll
                        for (int k = 0; k < kBK / kUmmaK; ++k) {
                            const uint32_t accum = (kc > 0 || k > 0) ? 1u : 0u;
                            umma_f16_pair(acc_s, da_s + 2 * k, db_s + 2 * k, idesc_st, accum);
                            if (kTeacher) umma_f16_pair(acc_t, da_t + 2 * k, db_t + 2 * k, idesc_st, accum);
                        }
                        umma_commit_pair(bar_empty + 8 * stage, 3);
                        if (kc == n_kc - 1) umma_commit_pair(bar_stfull + 8 * as, 3);
                    }
                    __syncwarp();
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                    if (p.trace) t_mark = clock64();
                }
                if (lane == 0) {
                    DCB_TRACE(t, 1);
                    if (p.trace && blockIdx.x < 2 && t < kTraceTiles) p.trace[((size_t)(blockIdx.x & 1) * kTraceTiles + t) * kTraceSlots + 4] = t_wait;
                }
            };
            // order on the (in-order) MMA pipe: S/T(t), grad(t-1), S/T(t+1), ...  -- the gradient GEMM of tile t-1 runs
            // while the epilogue warps work on tile t; with ST == 2 the S/T GEMM of tile t+1 overlaps them as well
            for (int tt = 0; tt <= n_tiles; ++tt) {
                if (tt < n_tiles) issue_st(tt);
                if (tt == 0) continue;
                const int t = tt - 1;
                mbar_wait(bar_gfull, t & 1);
                if (lane == 0) DCB_TRACE(t, 2);
                tc_fence_after_sync();
                for (int sl = 0; sl < p.slices; ++sl) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after_sync();
                    const uint32_t src = ring + stage * C::kStageBytes;
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < kSub; ++ks) {
                            const uint64_t dg = umma_desc_k_sw128(g_smem + ks * (kRowsPerCta * kBK * 2));
                            const uint64_t dbt = umma_desc_k_sw128(src + ks * (kSliceRows * kBK * 2));
#pragma unroll
                            for (int k = 0; k < kBK / kUmmaK; ++k)
                                umma_f16_pair(tmem_base + C::kAccCol + sl * 128, dg + 2 * k, dbt + 2 * k, idesc_grad,
                                              (t > 0 || ks > 0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit_pair(bar_empty + 8 * stage, 3);
                        if (sl == p.slices - 1) {
                            umma_commit_pair(bar_gempty, 3);
                            if (t == n_tiles - 1) umma_commit_pair(bar_accfull, 3);
                        }
                    }
                    __syncwarp();
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                if (lane == 0) DCB_TRACE(t, 3);
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- epilogue (both CTAs, own 64 rows)
        static_assert(NT == 128, "epilogue mapping below assumes 128-wide tiles: 64 columns per lane, 32 per warp");
        const int ew = warp - 4;                       // 0..7
        const int q = ew & 3;                          // TMEM lane quadrant (== warp % 4)
        const int sub = ew >> 2;                       // which 32-column chunk of this lane's 64 columns
        const int r = (q & 1) * 32 + lane;             // row inside this CTA's 64
        const int half = q >> 1;                       // lanes 0-63: columns [0, 64), lanes 64-127: columns [64, 128)
        const int cbase = half * 64 + sub * 32;        // first tile column of this thread
        const int ep_tid = ew * 32 + lane;             // 0..255
        const int grow = row0 + r;
        const bool row_ok = grow < p.rows;
        const float LOG2E = 1.4426950408889634f;
        const float r_s = row_ok ? __ldg(p.a_inv_stu + grow) : 0.f;
        const float r_t = (kTeacher && row_ok) ? __ldg(p.a_inv_tea + grow) : 0.f;
        const float k1 = r_s * LOG2E, k1t = r_s * LOG2E * p.inv_temp, k2t = r_t * LOG2E * p.inv_temp;
        const float n1 = -LOG2E, n1t = -LOG2E * p.inv_temp;
        float up_h = 1.f, up_s = 1.f, gbound;
        float xc1 = 0.f, xc2 = 0.f;                       // kExtra: up_cos / (B (B - 1)) and 2 up_mse / B^2
        if (p.bounds) {
            clip_load_upstream(p.up, up_h, up_s);
            gbound = clip_grad_bound(p.bounds, up_h, up_s);
            if constexpr (kExtra) {
                float up_c, up_m;
                clip_load_upstream_extra(p.up, up_c, up_m);
                gbound += clip_extra_bound(up_c, up_m, p.inv_batch, p.inv_pairs);
                xc1 = up_c * p.inv_pairs;
                xc2 = 2.f * up_m * p.inv_batch * p.inv_batch;
            }
        } else {
            gbound = __ldg(p.gmax_row) + __ldg(p.gmax_col);
        }
        const float ra = row_ok ? up_h * __ldg(p.coef_row + grow) : 0.f;
        const float rbeta = (kTeacher && row_ok) ? up_s * __ldg(p.coef_row + p.rows + grow) : 0.f;
        const float rg = (kTeacher && row_ok) ? up_s * __ldg(p.coef_row + 2 * (size_t)p.rows + grow) : 0.f;
        const float gscale = pair_tile_scale(gbound);
        // column scales / coefficients of a tile: fetched one tile ahead into registers (the ~1000 clk of global-load latency
        // sat on the per-tile critical path, trace r01q), parked in shared memory at the top of the tile that uses them
        float pre0 = 0.f, pre1 = 0.f, pre2 = 0.f;
        auto fetch_scales = [&](int t_next) {
            pre0 = pre1 = pre2 = 0.f;
            if (t_next >= n_tiles) return;
            const int c = ep_tid < NT ? ep_tid : ep_tid - NT;
            const int gc = (tile_begin + t_next) * NT + c;
            if (gc >= p.cols) return;
            if (ep_tid < NT) {
                pre0 = __ldg(p.b_inv_stu + gc);
                pre1 = up_h * __ldg(p.coef_col + gc);
            } else if (kTeacher) {
                pre0 = __ldg(p.b_inv_tea + gc);
                pre1 = up_s * __ldg(p.coef_col + p.cols + gc);
                pre2 = up_s * __ldg(p.coef_col + 2 * (size_t)p.cols + gc);
            }
        };
        fetch_scales(0);
        for (int t = 0; t < n_tiles; ++t) {
            const int as = t % ST;
            const int col0 = (tile_begin + t) * NT;
            float* sc = scale_buf + (t & 1) * 5 * NT;      // [c_stu][c_tea][alpha'][beta'][gamma'] x NT
            const bool tr = ep_tid == 0;
            if (tr) DCB_TRACE(t, 8);
            if (ep_tid < NT) {
                sc[ep_tid] = pre0;
                sc[2 * NT + ep_tid] = pre1;
            } else if (kTeacher) {
                const int c = ep_tid - NT;
                sc[NT + c] = pre0;
                sc[3 * NT + c] = pre1;
                sc[4 * NT + c] = pre2;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            fetch_scales(t + 1);
            if (tr) DCB_TRACE(t, 9);
            mbar_wait(bar_stfull + 8 * as, (t / ST) & 1);
            if (tr) DCB_TRACE(t, 10);
            tc_fence_after_sync();
            const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * C::kStCols + sub * 32;
            uint32_t packed[16];
            float sv[32], tv[32];
            if (tr) DCB_TRACE(t, 16);
            tmem_ld_32x32(lane_addr, sv);
            if (kTeacher) tmem_ld_32x32(lane_addr + NT / 2, tv);
            tmem_ld_wait();
            if (tr) DCB_TRACE(t, 17);
            // S/T columns of this warp are in registers: release the accumulator stage as early as possible
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(l_stempty + 8 * as);
            if (tr) DCB_TRACE(t, 11);
            if (p.dump_s) {                                     // tests only; kept out of the hot loop
                for (int c = 0; c < 32; ++c)
                    if (row_ok && col0 + cbase + c < p.cols)
                        p.dump_s[(size_t)grow * p.cols + col0 + cbase + c] = sv[c] * sc[cbase + c] * r_s;
            }
            const float* s_cs = sc + cbase;
            const float* s_ct = sc + NT + cbase;
            const float* s_a = sc + 2 * NT + cbase;
            const float* s_b = sc + 3 * NT + cbase;
            const float* s_g = sc + 4 * NT + cbase;
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
                float g2[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float u = sv[c + e] * s_cs[c + e];
                    float g = ex2p(fmaf(u, k1, n1)) * (ra + s_a[c + e]);
                    if (kTeacher) {
                        const float v = tv[c + e] * s_ct[c + e];
                        g = fmaf(ex2p(fmaf(u, k1t, n1t)), rbeta + s_b[c + e], g);
                        g = fmaf(-ex2p(fmaf(v, k2t, n1t)), rg + s_g[c + e], g);
                        if constexpr (kExtra) {
                            const float d = fmaf(u, r_s, -(v * r_t));                  // S_ij - T_ij
                            const bool off_diag = col0 + cbase + c + e != p.diag0 + grow;
                            g = fmaf(xc2, d, g);                                       // logits_mse.py:9-10
                            g += (d > 0.f && off_diag) ? xc1 : 0.f;                    // clip_cos_diff.py:21-23 (relu'(0) = 0)
                        }
                    }
                    g2[e] = g * gscale;
                }
                packed[c >> 1] = pack2<__half>(g2[0], g2[1]);
            }
            if (tr) DCB_TRACE(t, 12);
            if (p.g_out && row_ok) {          // 64 contiguous bytes per thread; the b-side gradient GEMM reads them back (clip_bwd_gt.cu)
                __half* dst = p.g_out + (size_t)grow * p.g_ld + col0 + cbase;
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8)
                    if (col0 + cbase + 8 * c8 < p.g_ld)
                        reinterpret_cast<uint4*>(dst)[c8] = make_uint4(packed[4 * c8], packed[4 * c8 + 1], packed[4 * c8 + 2], packed[4 * c8 + 3]);
            }
            mbar_wait(bar_gempty, (t & 1) ^ 1);
            if (tr) DCB_TRACE(t, 13);
            // columns [cbase, cbase + 32) of row r -> K-major SW128 sub-tile `half`, 16-byte chunks sub*4 .. sub*4+3
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                uint4 w = make_uint4(packed[4 * c8], packed[4 * c8 + 1], packed[4 * c8 + 2], packed[4 * c8 + 3]);
                *reinterpret_cast<uint4*>(g_gen + half * (kRowsPerCta * kBK * 2) + sw128_chunk_offset(r, sub * 4 + c8)) = w;
            }
            if (tr) DCB_TRACE(t, 18);
            fence_proxy_async_smem();
            if (tr) DCB_TRACE(t, 19);
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(l_gfull);
            if (tr) DCB_TRACE(t, 14);
        }
        // gradient accumulator -> global partial buffer.  Slice sl: lanes 0-63 hold d in [256 sl, 256 sl + 128),
        // lanes 64-127 hold d in [256 sl + 128, 256 sl + 256); 128 TMEM columns per slice, split between the two warps
        // of a quadrant.
        mbar_wait(bar_accfull, 0);
        tc_fence_after_sync();
        const uint32_t acc_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + C::kAccCol;
        float* out = p.acc + ((size_t)sp * p.rows + (row_ok ? grow : 0)) * p.dim;
        for (int sl = 0; sl < p.slices; ++sl) {
#pragma unroll 1
            for (int ch = 2 * sub; ch < 2 * sub + 2; ++ch) {
                const int d0 = sl * 256 + half * 128 + ch * 32;
                if (d0 >= p.dim) break;                                  // warp-uniform
                float v[32];
                tmem_ld_32x32(acc_addr + sl * 128 + ch * 32, v);
                tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        const int d = d0 + c;
                        if (d + 3 < p.dim) {
                            *reinterpret_cast<float4*>(out + d) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                        } else {
                            for (int e = 0; e < 4; ++e)
                                if (d + e < p.dim) out[d + e] = v[c + e];
                        }
                    }
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                  // the peer may still be reading operands this CTA's MMAs depend on / arriving remotely
    if (warp == 1) {
        __syncwarp();
        tc_fence_after_sync();
        tmem_dealloc_pair(tmem_base, kTmemCols);
    }
}

static int clip_bwd_pair_nt(int64_t) { return 128; }

static int clip_bwd_pair_splits(int64_t rows, int64_t cols, int64_t dim) {
    const int64_t row_blocks = (rows + 127) / 128;
    const int64_t col_tiles = (cols + clip_bwd_pair_nt(dim) - 1) / clip_bwd_pair_nt(dim);
    const int64_t slots = kNumSMs / 2;                   // clusters resident at once
    if (const char* e = getenv("DCB_DEBUG_SPLITS")) return atoi(e) > 0 ? atoi(e) : 1;     // profiling experiments only
    int64_t best = 1;
    double best_cost = 1e30;
    for (int64_t n = 1; n <= 32 && n <= col_tiles; ++n) {
        const int64_t waves = (row_blocks * n + slots - 1) / slots;
        const int64_t tiles = (col_tiles + n - 1) / n;
        const double cost = (double)waves * ((double)tiles + 1.5) + 0.02 * (double)n;     // 1.5 tiles of prologue/epilogue per CTA
        if (cost < best_cost) { best_cost = cost; best = n; }
    }
    return (int)best;
}

template <bool kTeacher, int NT, int ST, bool kExtra = false>
static int launch_pair(dim3 grid, int smem, cudaStream_t st, const CUtensorMap& ma_s, const CUtensorMap& mb_s,
                       const CUtensorMap& ma_t, const CUtensorMap& mb_t, const CUtensorMap& mbt,
                       const ClipBwdPairParams& p, uint32_t idesc_st, uint32_t idesc_grad) {
    static int max_set = 0;
    if (smem > max_set) {
        DCB_CUDA_OK(cudaFuncSetAttribute(clip_bwd_pair_kernel<kTeacher, NT, ST, kExtra>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        max_set = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(bwdp::kThreads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DCB_CUDA_OK(cudaLaunchKernelEx(&cfg, clip_bwd_pair_kernel<kTeacher, NT, ST, kExtra>, ma_s, mb_s, ma_t, mb_t, mbt, p, idesc_st, idesc_grad));
    return 0;
}

}  // namespace dcb

extern "C" int dcb_clip_pair_supported(int64_t dim) { return dim >= 8 && dim <= 768 && dim % 8 == 0; }
extern "C" int dcb_clip_pair_splits(int64_t rows_local, int64_t cols, int64_t dim) {
    return dcb::clip_bwd_pair_splits(rows_local, cols, dim);
}

namespace dcb {
static int clip_pair_launch(const void* stu_a, const void* stu_b, const void* tea_a, const void* tea_b,
                            const void* stu_b_t, int64_t bt_pitch_elems, int64_t bt_block_cols,
                            const float* stu_a_inv, const float* stu_b_inv, const float* tea_a_inv,
                            const float* tea_b_inv, const float* coef_row, const float* coef_col,
                            const float* gmax_row, const float* gmax_col, const float* bounds, const ClipUpstream& up,
                            int extra, int64_t row_offset, int64_t global_batch, int64_t rows_local, int64_t cols,
                            int64_t dim, int dtype, float temperature, float* acc_parts, void* g_out,
                            int64_t g_pitch_elems, float* dump_s, long long* trace, void* stream) {
    DCB_REQUIRE(stu_a && stu_b && stu_b_t && stu_a_inv && stu_b_inv && coef_row && coef_col && acc_parts &&
                    (bounds || (gmax_row && gmax_col)), "NULL pointer argument");
    if (bt_block_cols <= 0 || bt_block_cols >= cols) bt_block_cols = cols;
    DCB_REQUIRE(bt_block_cols == cols || (bt_block_cols % 128 == 0 && cols % bt_block_cols == 0),
                "b_hatT blocks must hold a multiple of 128 columns and divide the column count");
    DCB_REQUIRE(dtype == DCB_BF16 || dtype == DCB_F16, "the fused contrastive kernel takes bf16 or fp16 embeddings");
    DCB_REQUIRE(dcb_clip_pair_supported(dim), "pair kernel supports 8 <= dim <= 768, dim %% 8 == 0 (got %lld)", (long long)dim);
    DCB_REQUIRE(rows_local >= 1 && cols >= 1, "bad shape");
    DCB_REQUIRE(bt_pitch_elems >= bt_block_cols && bt_pitch_elems % 8 == 0, "bT pitch must be >= the block width and a multiple of 8 elements");
    DCB_REQUIRE(!g_out || (g_pitch_elems >= cols && g_pitch_elems % 8 == 0 && reinterpret_cast<uintptr_t>(g_out) % 16 == 0),
                "G scratch: pitch must be >= cols and a multiple of 8 elements, base 16-byte aligned");
    const bool teacher = tea_a != nullptr;
    if (teacher) DCB_REQUIRE(tea_b && tea_a_inv && tea_b_inv && temperature > 0.f, "teacher arguments incomplete");
    const int nt = clip_bwd_pair_nt(dim);
    CUtensorMap ma_s, mb_s, ma_t, mb_t, mbt;
    const uint64_t pitch = (uint64_t)dim * 2;
    if (tc::encode_tile_map_16bit(&ma_s, stu_a, rows_local, dim, pitch, bwdp::kRowsPerCta)) return 1;
    if (tc::encode_tile_map_16bit(&mb_s, stu_b, cols, dim, pitch, nt / 2)) return 1;
    if (teacher) {
        if (tc::encode_tile_map_16bit(&ma_t, tea_a, rows_local, dim, pitch, bwdp::kRowsPerCta)) return 1;
        if (tc::encode_tile_map_16bit(&mb_t, tea_b, cols, dim, pitch, nt / 2)) return 1;
    } else {
        ma_t = ma_s;
        mb_t = mb_s;
    }
    const int64_t n_blocks = cols / bt_block_cols;
    if (tc::encode_tile_map_16bit(&mbt, stu_b_t, dim * n_blocks, bt_block_cols, (uint64_t)bt_pitch_elems * 2, bwdp::kSliceRows)) return 1;
    ClipBwdPairParams p{};
    p.a_inv_stu = stu_a_inv;
    p.b_inv_stu = stu_b_inv;
    p.a_inv_tea = tea_a_inv;
    p.b_inv_tea = tea_b_inv;
    p.coef_row = coef_row;
    p.coef_col = coef_col;
    p.gmax_row = gmax_row;
    p.gmax_col = gmax_col;
    p.bounds = bounds;
    p.up = up;
    p.bt_block_cols = (int)bt_block_cols;
    p.bt_block_rows = (int)dim;
    p.extra = extra;
    p.diag0 = (int)row_offset;
    if (extra) {
        DCB_REQUIRE(teacher && bounds && global_batch >= 1, "the cos_diff / logits_mse terms need the teacher and the pipeline mode");
        p.inv_batch = 1.0f / (float)global_batch;
        p.inv_pairs = global_batch > 1 ? (float)(1.0 / ((double)global_batch * (double)(global_batch - 1))) : 0.f;
    }
    p.acc = acc_parts;
    p.g_out = static_cast<__half*>(g_out);
    p.g_ld = g_pitch_elems;
    p.dump_s = dump_s;
    p.trace = trace;
    p.rows = (int)rows_local;
    p.cols = (int)cols;
    p.dim = (int)dim;
    p.slices = (int)((dim + 255) / 256);
    p.n_split = clip_bwd_pair_splits(rows_local, cols, dim);
    p.col_tiles = (int)((cols + nt - 1) / nt);
    p.inv_temp = teacher ? 1.0f / temperature : 1.0f;
    const int row_blocks = (int)((rows_local + 127) / 128);
    const uint32_t idesc_st = tc::umma_idesc_f16(128, nt, dtype == DCB_BF16 ? 1 : 0);
    const uint32_t idesc_grad = tc::umma_idesc_f16(128, 256, 0);          // fp16 G x fp16 b_hatT, 256-row slices
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((unsigned)(2 * row_blocks * p.n_split));
    if (extra) {
        if (dim <= 512) return launch_pair<true, 128, 2, true>(grid, bwdp::Cfg<128, 2>::smem_bytes(), st, ma_s, mb_s, ma_t, mb_t, mbt, p, idesc_st, idesc_grad);
        return launch_pair<true, 128, 1, true>(grid, bwdp::Cfg<128, 1>::smem_bytes(), st, ma_s, mb_s, ma_t, mb_t, mbt, p, idesc_st, idesc_grad);
    }
    if (dim <= 512) {       // S/T double buffered: 2 x 128 + D/2 <= 512 TMEM columns
        const int smem = bwdp::Cfg<128, 2>::smem_bytes();
        return teacher ? launch_pair<true, 128, 2>(grid, smem, st, ma_s, mb_s, ma_t, mb_t, mbt, p, idesc_st, idesc_grad)
                       : launch_pair<false, 128, 2>(grid, smem, st, ma_s, mb_s, ma_t, mb_t, mbt, p, idesc_st, idesc_grad);
    }
    const int smem = bwdp::Cfg<128, 1>::smem_bytes();   // D <= 768: 128 + 384 TMEM columns, S/T single buffered
    return teacher ? launch_pair<true, 128, 1>(grid, smem, st, ma_s, mb_s, ma_t, mb_t, mbt, p, idesc_st, idesc_grad)
                   : launch_pair<false, 128, 1>(grid, smem, st, ma_s, mb_s, ma_t, mb_t, mbt, p, idesc_st, idesc_grad);
}
}  // namespace dcb

extern "C" int dcb_clip_row_grads_pair(const void* stu_a, const void* stu_b, const void* tea_a, const void* tea_b,
                                       const void* stu_b_t, int64_t bt_pitch_elems,
                                       const float* stu_a_inv, const float* stu_b_inv, const float* tea_a_inv,
                                       const float* tea_b_inv, const float* coef_row, const float* coef_col,
                                       const float* gmax_row, const float* gmax_col, int64_t rows_local, int64_t cols,
                                       int64_t dim, int dtype, float temperature, float* acc_parts, void* g_out,
                                       int64_t g_pitch_elems, float* dump_s, long long* trace, void* stream) {
    return dcb::clip_pair_launch(stu_a, stu_b, tea_a, tea_b, stu_b_t, bt_pitch_elems, 0, stu_a_inv, stu_b_inv, tea_a_inv, tea_b_inv,
                                 coef_row, coef_col, gmax_row, gmax_col, nullptr, dcb::ClipUpstream{}, 0, 0, 0, rows_local, cols, dim, dtype,
                                 temperature, acc_parts, g_out, g_pitch_elems, dump_s, trace, stream);
}

// Pipeline flavour: UNIT coefficients (dcb_clip_post1 / dcb_clip_post2) times the upstream gradients read from the device, the
// fp16 scale from bounds[6], and b_hatT stored as one [dim, bt_block_cols] block per source rank (bt_block_cols = B / R).
extern "C" int dcb_clip_pair_bwd(const void* stu_a, const void* stu_b, const void* tea_a, const void* tea_b,
                                 const void* stu_b_t, int64_t bt_pitch_elems, int64_t bt_block_cols,
                                 const float* stu_a_inv, const float* stu_b_inv, const float* tea_a_inv,
                                 const float* tea_b_inv, const float* coef_row, const float* coef_col, const float* bounds,
                                 const float* const* g5, const float* w8, int extra, int64_t row_offset, int64_t global_batch,
                                 int64_t rows_local, int64_t cols, int64_t dim, int dtype,
                                 float temperature, float* acc_parts, void* g_out, int64_t g_pitch_elems, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(bounds && g5 && w8, "NULL bounds / upstream description");
    return clip_pair_launch(stu_a, stu_b, tea_a, tea_b, stu_b_t, bt_pitch_elems, bt_block_cols, stu_a_inv, stu_b_inv, tea_a_inv,
                            tea_b_inv, coef_row, coef_col, nullptr, nullptr, bounds, clip_upstream_from(g5, w8), extra, row_offset,
                            global_batch, rows_local, cols, dim, dtype, temperature, acc_parts, g_out, g_pitch_elems, nullptr, nullptr,
                            stream);
}
