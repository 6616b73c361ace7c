// Backward of the fused contrastive / logit-KL losses, SPLIT flow, step 1: recompute the logit tiles once and store the scaled
// fp16 gradient tiles  G_ij 2^k = dL/dS_ij 2^k  (both softmax directions folded in; the definition of G_ij and the reference
// lines are in clip_bwd.cu / clip_bwd_pair.cu: autograd through clip_model.py:36-44, hard_label.py:10-12, soft_label.py:11-16,
// and with kExtra clip_cos_diff.py:16-23, logits_mse.py:9-10).  Both towers' gradients are then plain tcgen05 GEMMs over the
// stored tiles (clip_bwd_gt.cu):  acc_a = G b_hat  (A = G K-major)  and  acc_b = G^T a_hat  (A = G^T MN-major).
//
// Why split: the fused pair kernel (clip_bwd_pair.cu) keeps the [rows x D] gradient accumulator in TMEM next to the S/T tiles,
// which caps the cluster tile at 128 rows x 128 columns (TMEM is exactly full at D = 768) -- 32 KiB of operands per K chunk
// for 64 x 128 logits per SM, single-buffered S/T accumulators -- and it ran at 58 % tensor-pipe activity, TMA-ingest bound.
// Without the accumulator the recompute has the forward kernel's shape: 256 rows per CTA pair (tcgen05.mma.cta_group::2,
// M = 256, N = 128), 48 KiB per K chunk for 128 x 128 logits per SM, S/T accumulators double-buffered in TMEM (2 x 256
// columns), 16 epilogue warps; and the a-side gradient GEMM runs at the G^T GEMM's ~85 % instead of inside a 58 % kernel.
// Cost: the 2 B_local B bytes of G scratch are read twice instead of once (HBM time hidden under the GEMMs).
//
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM allocator (+ MMA issuer in the leader), warps 2-17 = epilogue:
// warp w owns TMEM lanes 32 (w % 4) .. (one row per thread) and columns [32 sub, 32 sub + 32) of the tile, sub = (w - 2) / 4,
// 16 columns per tcgen05.ld; every thread stores 2 x 32 contiguous bytes of its row per tile.
#include "tc_common.cuh"
#include "clip_shared.cuh"

namespace dcb {

namespace gtl {
constexpr int kBM = 128, kBN = 128, kBK = 64, kUmmaK = 16;
constexpr int kStages = 4;
constexpr int kTileBytes = kBM * kBK * 2;                 // 16 KiB: one [128 x 64] 16-bit a-side tile
constexpr int kBHalfRows = kBN / 2;                       // b rows staged per CTA
constexpr int kBTileBytes = kBHalfRows * kBK * 2;         // 8 KiB
constexpr int kStageBytes = 2 * kTileBytes + 2 * kBTileBytes;   // a_stu, a_tea, b_stu half, b_tea half = 48 KiB
constexpr int kThreads = 576;                             // 2 control warps + 16 epilogue warps (4 per scheduler)
constexpr int kTmemCols = 512;
constexpr int kScaleBytes = 2 * 5 * kBN * 4;              // [tile parity][c_stu, c_tea, alpha', beta', gamma'][column] fp32
constexpr int kSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + kScaleBytes + 256;
}  // namespace gtl

struct ClipGTilesParams {
    const float* a_inv_stu;   // [rows]
    const float* b_inv_stu;   // [cols]
    const float* a_inv_tea;
    const float* b_inv_tea;
    const float* coef_row;    // [3][rows] UNIT coefficients (dcb_clip_post1)
    const float* coef_col;    // [3][cols] UNIT coefficients (dcb_clip_post2)
    const float* bounds;      // [6] -> the fp16 scale 2^k (clip_shared.cuh)
    ClipUpstream up;          // upstream gradients, read from the device
    int diag0;                // global column index of local row 0 (the diagonal carries no off-diagonal cos_diff term)
    float inv_batch, inv_pairs;   // 1/B, 1/(B (B - 1))
    __half* g_out;            // [rows][g_ld] fp16
    long long g_ld;
    int rows, cols, dim;
    int n_split, col_tiles;
    float inv_temp;
};

__device__ __forceinline__ float ex2g(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool kTeacher, bool kExtra>
__global__ void __launch_bounds__(gtl::kThreads, 1)
clip_g_tiles_kernel(const __grid_constant__ CUtensorMap map_a_stu, const __grid_constant__ CUtensorMap map_b_stu,
                    const __grid_constant__ CUtensorMap map_a_tea, const __grid_constant__ CUtensorMap map_b_tea,
                    const __grid_constant__ ClipGTilesParams p, const uint32_t idesc) {
    using namespace gtl;
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t ring = smem_base;
    float* scale_buf = reinterpret_cast<float*>(smem_gen + kStages * kStageBytes);   // [2][5][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_gen + kStages * kStageBytes + kScaleBytes);
    const uint32_t bar_full = smem_u32(bars);                   // [kStages] leader: TMA bytes of both CTAs
    const uint32_t bar_empty = bar_full + 8 * kStages;          // [kStages] each CTA: slot free (multicast commit)
    const uint32_t bar_tfull = bar_empty + 8 * kStages;         // [2] each CTA: accumulators complete (multicast commit)
    const uint32_t bar_tempty = bar_tfull + 16;                 // [2] leader: 16 epilogue warps of both CTAs drained them
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int rb = (cluster_id / p.n_split) * 2 + (int)rank, sp = cluster_id % p.n_split;
    const int tile_begin = (int)(((long long)sp * p.col_tiles) / p.n_split);
    const int tile_end = (int)(((long long)(sp + 1) * p.col_tiles) / p.n_split);
    const int n_tiles = tile_end - tile_begin;
    const int n_kc = (p.dim + kBK - 1) / kBK;
    const int row0 = rb * kBM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_tfull + 8 * s, 1);
            mbar_init(bar_tempty + 8 * s, 32);    // one arrive per epilogue warp of both CTAs
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(smem_u32(tmem_slot), kTmemCols);
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                                          // peer barriers initialised, both TMEM allocations done
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t l_tempty = map_to_cta(bar_tempty, 0);         // the leader's barriers as seen from this CTA

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (elect_one()) {
            tma_prefetch_desc(&map_a_stu);
            tma_prefetch_desc(&map_b_stu);
            if (kTeacher) {
                tma_prefetch_desc(&map_a_tea);
                tma_prefetch_desc(&map_b_tea);
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int t = 0; t < n_tiles; ++t) {
                const int col0 = (tile_begin + t) * kBN + (int)rank * kBHalfRows;       // this CTA's half of the b tile
                for (int kc = 0; kc < n_kc; ++kc) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t dst = ring + stage * kStageBytes;
                    if (leader) mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * (kTeacher ? 2 : 1) * (kTileBytes + kBTileBytes));
                    const uint32_t full = map_to_cta(bar_full + 8 * stage, 0);
                    tma_load_2d_pair(dst, &map_a_stu, full, kc * kBK, row0);
                    tma_load_2d_pair(dst + 2 * kTileBytes, &map_b_stu, full, kc * kBK, col0);
                    if (kTeacher) {
                        tma_load_2d_pair(dst + kTileBytes, &map_a_tea, full, kc * kBK, row0);
                        tma_load_2d_pair(dst + 2 * kTileBytes + kBTileBytes, &map_b_tea, full, kc * kBK, col0);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer (leader CTA): warp-uniform waits, one elected lane issues
        int stage = 0;
        uint32_t phase = 0;
        for (int t = 0; leader && t < n_tiles; ++t) {
            const int as = t & 1;                       // accumulator stage
            const uint32_t aphase = (t >> 1) & 1;
            mbar_wait(bar_tempty + 8 * as, aphase ^ 1);
            tc_fence_after_sync();
            const uint32_t acc_s = tmem_base + as * 256;
            const uint32_t acc_t = acc_s + 128;
            for (int kc = 0; kc < n_kc; ++kc) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after_sync();
                const uint32_t src = ring + stage * kStageBytes;
                if (elect_one()) {
                    const uint64_t da_s = umma_desc_k_sw128(src), da_t = umma_desc_k_sw128(src + kTileBytes);
                    const uint64_t db_s = umma_desc_k_sw128(src + 2 * kTileBytes);
                    const uint64_t db_t = umma_desc_k_sw128(src + 2 * kTileBytes + kBTileBytes);
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        const uint32_t accum = (kc > 0 || k > 0) ? 1u : 0u;
                        umma_f16_pair(acc_s, da_s + 2 * k, db_s + 2 * k, idesc, accum);     // +32 B per K=16 step
                        if (kTeacher) umma_f16_pair(acc_t, da_t + 2 * k, db_t + 2 * k, idesc, accum);
                    }
                    umma_commit_pair(bar_empty + 8 * stage, 3);     // frees the slot in both CTAs once these MMAs retire
                    if (kc == n_kc - 1) umma_commit_pair(bar_tfull + 8 * as, 3);     // accumulators of this tile are complete
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue: 16 warps, one row per thread
        const int q = warp & 3;                             // TMEM lane quadrant this warp may read
        const int sub = (warp - 2) >> 2;                    // column quarter of the tile handled by this warp
        const int r = q * 32 + lane;                        // row inside the block
        const int ep_tid = (warp - 2) * 32 + lane;          // 0..511, used to stage the column scales / coefficients
        const int grow = row0 + r;                          // local row
        const bool row_ok = grow < p.rows;
        const float LOG2E = 1.4426950408889634f;
        const float r_s = row_ok ? __ldg(p.a_inv_stu + grow) : 0.f;
        const float r_t = (kTeacher && row_ok) ? __ldg(p.a_inv_tea + grow) : 0.f;
        const float k1 = r_s * LOG2E, k1t = r_s * LOG2E * p.inv_temp, k2t = r_t * LOG2E * p.inv_temp;
        const float n1 = -LOG2E, n1t = -LOG2E * p.inv_temp;
        float up_h, up_s;
        clip_load_upstream(p.up, up_h, up_s);
        float gbound = clip_grad_bound(p.bounds, up_h, up_s);
        float xc1 = 0.f, xc2 = 0.f;                         // kExtra: 2^k up_cos / (B (B - 1)) and 2^k 2 up_mse / B^2
        if constexpr (kExtra) {
            float up_c, up_m;
            clip_load_upstream_extra(p.up, up_c, up_m);
            gbound += clip_extra_bound(up_c, up_m, p.inv_batch, p.inv_pairs);
            xc1 = up_c * p.inv_pairs;
            xc2 = 2.f * up_m * p.inv_batch * p.inv_batch;
        }
        // the power-of-two tile scale is folded into the coefficients (exact), not applied per element
        const float gscale = clip_tile_scale(gbound);
        const float uh = up_h * gscale, us = up_s * gscale;
        xc1 *= gscale;
        xc2 *= gscale;
        const float ra = row_ok ? uh * __ldg(p.coef_row + grow) : 0.f;
        const float rbeta = (kTeacher && row_ok) ? us * __ldg(p.coef_row + p.rows + grow) : 0.f;
        const float rg = (kTeacher && row_ok) ? us * __ldg(p.coef_row + 2 * (size_t)p.rows + grow) : 0.f;
        // column scales / coefficients: thread (k, c) = (ep_tid / 128, ep_tid % 128) fetches array k of column c one tile ahead
        // (threads of array 0 also fetch array 4), parked in shared memory at the top of the tile that uses them
        const int sk = ep_tid >> 7, scol = ep_tid & (kBN - 1);
        float pre0 = 0.f, pre1 = 0.f;
        auto fetch_scales = [&](int t_next) {
            pre0 = pre1 = 0.f;
            if (t_next >= n_tiles) return;
            const int gc = (tile_begin + t_next) * kBN + scol;
            if (gc >= p.cols) return;
            if (sk == 0) {
                pre0 = __ldg(p.b_inv_stu + gc);
                if (kTeacher) pre1 = us * __ldg(p.coef_col + 2 * (size_t)p.cols + gc);
            } else if (sk == 1) {
                if (kTeacher) pre0 = __ldg(p.b_inv_tea + gc);
            } else if (sk == 2) {
                pre0 = uh * __ldg(p.coef_col + gc);
            } else if (kTeacher) {
                pre0 = us * __ldg(p.coef_col + p.cols + gc);
            }
        };
        fetch_scales(0);
        __half* g_row = p.g_out + (size_t)(row_ok ? grow : 0) * p.g_ld;
        for (int t = 0; t < n_tiles; ++t) {
            const int as = t & 1;
            const uint32_t aphase = (t >> 1) & 1;
            const int col0 = (tile_begin + t) * kBN;
            float* sc = scale_buf + as * 5 * kBN;           // [c_stu][c_tea][alpha'][beta'][gamma'] x 128
            sc[sk * kBN + scol] = pre0;
            if (sk == 0) sc[4 * kBN + scol] = pre1;
            asm volatile("bar.sync 1, 512;" ::: "memory");      // scales staged; every warp is done with tile t-1
            fetch_scales(t + 1);
            mbar_wait(bar_tfull + 8 * as, aphase);
            tc_fence_after_sync();
            const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 256;
#pragma unroll 1
            for (int hc = 0; hc < 2; ++hc) {
                float sv[16], tv[16];
                const int cbase = sub * 32 + hc * 16;
                tmem_ld_32x16(lane_addr + cbase, sv);
                if (kTeacher) tmem_ld_32x16(lane_addr + 128 + cbase, tv);
                tmem_ld_wait();
                if (hc == 1) {                              // last TMEM read of this tile: release the accumulator stage
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(l_tempty + 8 * as);
                }
                const float* s_cs = sc + cbase;
                const float* s_ct = sc + kBN + cbase;
                const float* s_a = sc + 2 * kBN + cbase;
                const float* s_b = sc + 3 * kBN + cbase;
                const float* s_g = sc + 4 * kBN + cbase;
                uint32_t packed[8];
#pragma unroll
                for (int c = 0; c < 16; c += 2) {
                    float g2[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float u = sv[c + e] * s_cs[c + e];
                        float g = ex2g(fmaf(u, k1, n1)) * (ra + s_a[c + e]);
                        if (kTeacher) {
                            const float v = tv[c + e] * s_ct[c + e];
                            g = fmaf(ex2g(fmaf(u, k1t, n1t)), rbeta + s_b[c + e], g);
                            g = fmaf(-ex2g(fmaf(v, k2t, n1t)), rg + s_g[c + e], g);
                            if constexpr (kExtra) {
                                const float d = fmaf(u, r_s, -(v * r_t));                  // S_ij - T_ij
                                const bool off_diag = col0 + cbase + c + e != p.diag0 + grow;
                                g = fmaf(xc2, d, g);                                       // logits_mse.py:9-10
                                g += (d > 0.f && off_diag) ? xc1 : 0.f;                    // clip_cos_diff.py:21-23 (relu'(0) = 0)
                            }
                        }
                        g2[e] = g;
                    }
                    packed[c >> 1] = pack2<__half>(g2[0], g2[1]);
                }
                if (row_ok) {          // 32 contiguous bytes per thread; the two gradient GEMMs read them back (clip_bwd_gt.cu)
                    __half* dst = g_row + col0 + cbase;
#pragma unroll
                    for (int c8 = 0; c8 < 2; ++c8)
                        if (col0 + cbase + 8 * c8 < p.g_ld)
                            reinterpret_cast<uint4*>(dst)[c8] = make_uint4(packed[4 * c8], packed[4 * c8 + 1], packed[4 * c8 + 2], packed[4 * c8 + 3]);
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

// Column ranges per pair of row blocks: minimise (waves of 74 clusters) x (tiles per cluster + fixed per-cluster cost)
static int clip_g_tiles_splits(int64_t rows, int64_t cols) {
    const int64_t row_blocks = (rows + gtl::kBM - 1) / gtl::kBM;
    const int64_t col_tiles = (cols + gtl::kBN - 1) / gtl::kBN;
    if (const char* e = getenv("DCB_DEBUG_SPLITS")) return atoi(e) > 0 ? atoi(e) : 1;     // profiling experiments only
    int64_t best = 1;
    double best_cost = 1e30;
    for (int64_t n = 1; n <= 64 && n <= col_tiles; ++n) {
        const int64_t waves = ((row_blocks + 1) / 2 * n + kNumSMs / 2 - 1) / (kNumSMs / 2);
        const int64_t tiles = (col_tiles + n - 1) / n;
        const double cost = (double)waves * ((double)tiles + 1.0) + 0.01 * (double)n;
        if (cost < best_cost) { best_cost = cost; best = n; }
    }
    return (int)best;
}

}  // namespace dcb

extern "C" int dcb_clip_g_tiles(const void* stu_a, const void* stu_b, const void* tea_a, const void* tea_b,
                                const float* stu_a_inv, const float* stu_b_inv, const float* tea_a_inv,
                                const float* tea_b_inv, const float* coef_row, const float* coef_col, const float* bounds,
                                const float* const* g5, const float* w8, int extra, int64_t row_offset, int64_t global_batch,
                                int64_t rows_local, int64_t cols, int64_t dim, int dtype, float temperature, void* g_out,
                                int64_t g_pitch_elems, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(stu_a && stu_b && stu_a_inv && stu_b_inv && coef_row && coef_col && bounds && g5 && w8 && g_out, "NULL pointer argument");
    DCB_REQUIRE(dtype == DCB_BF16 || dtype == DCB_F16, "the fused contrastive kernel takes bf16 or fp16 embeddings");
    DCB_REQUIRE(rows_local >= 1 && cols >= 1 && dim >= 8 && dim % 8 == 0, "bad shape rows=%lld cols=%lld dim=%lld (dim %% 8 == 0)",
                (long long)rows_local, (long long)cols, (long long)dim);
    DCB_REQUIRE(rows_local < (1ll << 30) && cols < (1ll << 30), "batch too large");
    DCB_REQUIRE(g_pitch_elems >= cols && g_pitch_elems % 8 == 0 && reinterpret_cast<uintptr_t>(g_out) % 16 == 0,
                "G scratch: pitch must be >= cols and a multiple of 8 elements, base 16-byte aligned");
    const bool teacher = tea_a != nullptr;
    if (teacher) DCB_REQUIRE(tea_b && tea_a_inv && tea_b_inv && temperature > 0.f, "teacher arguments incomplete");
    CUtensorMap ma_s, mb_s, ma_t, mb_t;
    const uint64_t pitch = (uint64_t)dim * 2;
    if (tc::encode_tile_map_16bit(&ma_s, stu_a, rows_local, dim, pitch, gtl::kBM)) return 1;
    if (tc::encode_tile_map_16bit(&mb_s, stu_b, cols, dim, pitch, gtl::kBHalfRows)) return 1;
    if (teacher) {
        if (tc::encode_tile_map_16bit(&ma_t, tea_a, rows_local, dim, pitch, gtl::kBM)) return 1;
        if (tc::encode_tile_map_16bit(&mb_t, tea_b, cols, dim, pitch, gtl::kBHalfRows)) return 1;
    } else {
        ma_t = ma_s;
        mb_t = mb_s;
    }
    ClipGTilesParams p{};
    p.a_inv_stu = stu_a_inv;
    p.b_inv_stu = stu_b_inv;
    p.a_inv_tea = tea_a_inv;
    p.b_inv_tea = tea_b_inv;
    p.coef_row = coef_row;
    p.coef_col = coef_col;
    p.bounds = bounds;
    p.up = clip_upstream_from(g5, w8);
    p.diag0 = (int)row_offset;
    if (extra) {
        DCB_REQUIRE(teacher && global_batch >= 1, "the cos_diff / logits_mse terms need the teacher");
        p.inv_batch = 1.0f / (float)global_batch;
        p.inv_pairs = global_batch > 1 ? (float)(1.0 / ((double)global_batch * (double)(global_batch - 1))) : 0.f;
    }
    p.g_out = static_cast<__half*>(g_out);
    p.g_ld = g_pitch_elems;
    p.rows = (int)rows_local;
    p.cols = (int)cols;
    p.dim = (int)dim;
    p.n_split = clip_g_tiles_splits(rows_local, cols);
    p.col_tiles = (int)((cols + gtl::kBN - 1) / gtl::kBN);
    p.inv_temp = teacher ? 1.0f / temperature : 1.0f;
    const int row_blocks = (int)((rows_local + gtl::kBM - 1) / gtl::kBM);
    const uint32_t idesc = tc::umma_idesc_f16(2 * gtl::kBM, gtl::kBN, dtype == DCB_BF16 ? 1 : 0);     // M = 256 over the CTA pair
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((unsigned)(2 * ((row_blocks + 1) / 2) * p.n_split));
#define DCB_LAUNCH_GT(TEA, EXTRA)                                                                                      \
    {                                                                                                                  \
        static const cudaError_t attr_ = cudaFuncSetAttribute(clip_g_tiles_kernel<TEA, EXTRA>,                         \
                                                              cudaFuncAttributeMaxDynamicSharedMemorySize, gtl::kSmemBytes); \
        DCB_CUDA_OK(attr_);     /* set once per process (not a stream operation; kept out of graph captures) */         \
        cudaLaunchConfig_t cfg_{};                                                                                     \
        cfg_.gridDim = grid;                                                                                           \
        cfg_.blockDim = dim3(gtl::kThreads);                                                                           \
        cfg_.dynamicSmemBytes = gtl::kSmemBytes;                                                                       \
        cfg_.stream = st;                                                                                              \
        cudaLaunchAttribute attr2_[1];                                                                                 \
        attr2_[0].id = cudaLaunchAttributeClusterDimension;                                                            \
        attr2_[0].val.clusterDim.x = 2;                                                                                \
        attr2_[0].val.clusterDim.y = 1;                                                                                \
        attr2_[0].val.clusterDim.z = 1;                                                                                \
        cfg_.attrs = attr2_;                                                                                           \
        cfg_.numAttrs = 1;                                                                                             \
        DCB_CUDA_OK(cudaLaunchKernelEx(&cfg_, clip_g_tiles_kernel<TEA, EXTRA>, ma_s, mb_s, ma_t, mb_t, p, idesc));     \
    }
    if (teacher && extra) DCB_LAUNCH_GT(true, true)
    else if (teacher) DCB_LAUNCH_GT(true, false)
    else DCB_LAUNCH_GT(false, false)
#undef DCB_LAUNCH_GT
    DCB_CUDA_OK(cudaGetLastError());
    return 0;
}
