// Second half of the single-recompute backward of the fused contrastive / logit-KL losses.
//
// clip_bwd_pair_kernel (clip_bwd_pair.cu) recomputes the logits once, forms the scaled fp16 gradient tile
// G_ij 2^k = dL/dS_ij 2^k (both softmax directions already folded in, see clip_bwd.cu), multiplies it into the a-side
// gradient acc_a = G b_hat, and -- new -- stores the tile to HBM.  This kernel finishes the b side from the stored tiles,
//        acc_b[j, :] = sum_i G[i, j] a_hat[i, :]                      (autograd through clip_model.py:37-44 for the other tower)
// as a plain tcgen05 GEMM, so the executed work of the backward equals its algorithmic minimum
// (recompute 4 B^2 D + two gradient GEMMs 4 B^2 D) instead of recomputing the logits once per direction.
// 2 B^2 bytes of G scratch are written once and read once: 2 x 2 GiB at B = 32768, ~0.7 ms of HBM time hidden under
// ~5 ms of tensor work it replaces ~2.9 ms of.
//
// Mapping: cluster of 2 CTAs, tcgen05.mma.cta_group::2 with M = 256 (128 j per CTA), N = one chunk of the embedding
// dimension (<= 384 columns, two instructions of N/2 when N > 256), K = i in chunks of 64.  A = G^T is read straight
// from the row-major [i, j] scratch as an MN-major operand: a TMA box {64 j, 64 i} with the 128-byte swizzle IS the
// canonical MN-major SW128 layout ((8,8,m),(8,k)):((1,8,LBO),(64,SBO)) in 16-bit elements, LBO = 8 KiB between the
// two 64-wide M blocks, SBO = 1 KiB between 8-row K groups; a K = 16 step advances the start address by 2 KiB.
// B = a_hat^T [D, rows] fp16 (K-major, the same layout the a-side GEMM uses for b_hat^T), each CTA stages half of
// the N rows.  K can be split over clusters (fp32 partial buffers, summed by clip_grad_finish).
// Warps: 0 = TMA, 1 = TMEM alloc + MMA issue (leader), 2-5 = epilogue (TMEM -> fp32 partial buffer).
//
// kAK = true instantiation (split backward, clip_bwd_g.cu): the OTHER tower's gradient from the same stored tiles,
//        acc_a[i, :] = sum_j G[i, j] b_hat[j, :]
// -- M = i, K = j, so A = G is an ordinary K-major operand (one TMA box {64 j, 128 i} per CTA and K chunk) and B = b_hat^T
// [D, cols] fp16 K-major, stored as one [D, cols / R] block per source rank (`b_block_cols`).  Same ring, MMA shape, K split
// and epilogue; in the code below "rows" is always the K extent and "cols" the M extent of the launch.
#include "tc_common.cuh"

namespace dcb {

namespace gt {
constexpr int kBK = 64, kUmmaK = 16;
constexpr int kThreads = 192;
constexpr int kATile = 128 * kBK * 2;                  // 16 KiB: [2 M blocks][64 k][128 B]
constexpr int kMaxChunk = 384;                         // N columns per cluster
constexpr int kSmemBudget = 200 * 1024;
}  // namespace gt

constexpr int kGtMaxDest = 16;

struct ClipGtParams {
    float* acc;               // [k_split][cols][dim] fp32 (single destination)
    // Fused reduce-scatter over peer memory: output row j belongs to rank j / rows_per_dest and is stored straight into THAT
    // rank's partial buffer dest[rank] (a peer mapping: NVLink stores from the epilogue), slot src_slot * k_split + ks of
    // [n_dest * k_split][rows_per_dest][dim]; the owner sums the slots in a fixed order (clip_grad_finish).  n_dest == 0: off.
    float* dest[kGtMaxDest];
    int n_dest, src_slot;
    long long rows_per_dest;
    int rows, cols, dim;      // rows = i (K), cols = j (M)
    int chunk;                // N columns per cluster (multiple of 32 when > 256, of 16 otherwise)
    int pieces;               // MMA instructions per K step (1 or 2), each chunk / pieces wide
    int n_chunks, m_tiles, k_split;
    int stages, stage_bytes;
    int b_block_cols;         // kAK: B is stored as row blocks [n_blocks][b_block_rows][b_block_cols] over K (0 = one block)
    int b_block_rows;
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);   // start address
    d |= static_cast<uint64_t>(8192 >> 4) << 16;               // LBO: next 64-element block along M
    d |= static_cast<uint64_t>(1024 >> 4) << 32;               // SBO: next group of 8 K rows
    d |= static_cast<uint64_t>(1) << 46;                       // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;                       // SWIZZLE_128B
    return d;
}

template <bool kAK>
__global__ void __launch_bounds__(gt::kThreads, 1)
clip_gt_gemm_kernel(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_at,
                    const __grid_constant__ ClipGtParams p, const uint32_t idesc) {
    using namespace gt;
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t ring = smem_base;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_gen + p.stages * p.stage_bytes);
    const uint32_t bar_full = smem_u32(bars);                    // [stages] leader: bytes of both CTAs landed
    const uint32_t bar_empty = bar_full + 8 * p.stages;          // [stages] each CTA: slot free (multicast commit)
    const uint32_t bar_accfull = bar_empty + 8 * p.stages;       // each CTA: accumulator final
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    int unit = blockIdx.x >> 1;
    const int nc = unit % p.n_chunks;
    unit /= p.n_chunks;
    const int mt = unit % p.m_tiles, ks = unit / p.m_tiles;
    const int total_kc = (p.rows + kBK - 1) / kBK;
    const int kc_begin = (int)(((long long)ks * total_kc) / p.k_split);
    const int kc_end = (int)(((long long)(ks + 1) * total_kc) / p.k_split);
    const int j0 = mt * 256 + (int)rank * 128;                   // first M row (j) of this CTA
    const int n0 = nc * p.chunk;                                 // first N column (d) of this cluster
    const int piece_n = p.chunk / p.pieces;                      // N of one MMA
    const int half_n = piece_n / 2;                              // B rows staged per CTA per piece

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_accfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(smem_u32(tmem_slot), 512);
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            tma_prefetch_desc(&map_g);
            tma_prefetch_desc(&map_at);
            const uint32_t bytes_per_cta = kATile + (uint32_t)p.pieces * half_n * kBK * 2;
            int stage = 0;
            uint32_t phase = 0;
            for (int kc = kc_begin; kc < kc_end; ++kc) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                const uint32_t dst = ring + stage * p.stage_bytes;
                if (leader) mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * bytes_per_cta);
                const uint32_t full = map_to_cta(bar_full + 8 * stage, 0);
                int bx = kc * kBK, by = n0 + (int)rank * half_n;
                if constexpr (kAK) {
                    tma_load_2d_pair(dst, &map_g, full, kc * kBK, j0);                       // K-major: 128 M rows x 64 k
                    if (p.b_block_cols > 0) {                                                // K chunks never straddle a block
                        const int blk = bx / p.b_block_cols;
                        bx -= blk * p.b_block_cols;
                        by += blk * p.b_block_rows;
                    }
                } else {
                    tma_load_2d_pair(dst, &map_g, full, j0, kc * kBK);                       // M block 0: j0 .. j0+63
                    tma_load_2d_pair(dst + 64 * kBK * 2, &map_g, full, j0 + 64, kc * kBK);   // M block 1
                }
                for (int pc = 0; pc < p.pieces; ++pc)
                    tma_load_2d_pair(dst + kATile + pc * half_n * kBK * 2, &map_at, full, bx, by + pc * piece_n);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kc = kc_begin; kc < kc_end; ++kc) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after_sync();
                const uint32_t src = ring + stage * p.stage_bytes;
                if (elect_one()) {
                    const uint64_t da = kAK ? umma_desc_k_sw128(src) : umma_desc_mn_sw128(src);
                    constexpr uint64_t kAStep = kAK ? 2 : (2048 >> 4);       // start-address step of A per K = 16
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        const uint32_t accum = (kc > kc_begin || k > 0) ? 1u : 0u;
                        for (int pc = 0; pc < p.pieces; ++pc) {
                            const uint64_t db = umma_desc_k_sw128(src + kATile + pc * half_n * kBK * 2);
                            umma_f16_pair(tmem_base + pc * piece_n, da + (uint64_t)k * kAStep, db + 2 * k, idesc, accum);
                        }
                    }
                    umma_commit_pair(bar_empty + 8 * stage, 3);
                    if (kc == kc_end - 1) umma_commit_pair(bar_accfull, 3);
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: TMEM lane = j, columns = d.
        // TMEM -> registers (thread = row) -> shared memory (the operand ring is free once the accumulator is final) ->
        // global with one ROW per store instruction: 512 contiguous bytes per warp store instead of 32 scattered 16-byte
        // pieces, which is what makes the stores to a peer's buffer over NVLink efficient (and helps the local ones too).
        const int q = warp & 3;
        constexpr int kStageLd = 132;                                   // floats per staged row (128 + 4: conflict-free)
        float* stage = reinterpret_cast<float*>(smem_gen) + q * (32 * kStageLd);
        const bool have_acc = kc_end > kc_begin;
        if (have_acc) {
            mbar_wait(bar_accfull, 0);
            tc_fence_after_sync();
        }
        const int rpd = (int)p.rows_per_dest;
        for (int g0 = 0; g0 < p.chunk; g0 += 128) {
            if (n0 + g0 >= p.dim) break;
            const int ccount = p.chunk - g0 < 128 ? p.chunk - g0 : 128;
            for (int c0 = 0; c0 < ccount; c0 += 32) {
                float v[32];
                if (have_acc) {
                    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g0 + c0, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] = 0.f;
                }
                float* dst = stage + lane * kStageLd + c0;
#pragma unroll
                for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
            }
            __syncwarp();
            const int d = n0 + g0 + 4 * lane;
            const bool col_ok = 4 * lane < ccount && d + 3 < p.dim;      // dim % 8 == 0: a float4 is all in or all out
            for (int r = 0; r < 32; ++r) {
                const int j = j0 + q * 32 + r;
                if (j >= p.cols) break;                                     // warp-uniform
                float* out_row;
                if (p.n_dest > 0) {
                    const int dst_rank = j / rpd;
                    out_row = p.dest[dst_rank] + (((size_t)p.src_slot * p.k_split + ks) * rpd + (j - dst_rank * rpd)) * p.dim;
                } else {
                    out_row = p.acc + ((size_t)ks * p.cols + j) * p.dim;
                }
                if (col_ok) *reinterpret_cast<float4*>(out_row + d) = *reinterpret_cast<const float4*>(stage + r * kStageLd + 4 * lane);
            }
            __syncwarp();
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after_sync();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

struct GtPlan {
    int chunk, pieces, n_chunks, m_tiles, k_split, stages, stage_bytes;
};

// scatter: the partial buffers live in the owners' memory (NVLink stores) -- every extra K split is another [cols, dim] fp32
// image over the links and another slot for the finish kernel to sum, which costs far more than the wave it balances
static GtPlan clip_gt_plan(int64_t rows, int64_t cols, int64_t dim, bool scatter = false) {
    GtPlan g{};
    g.n_chunks = (int)((dim + gt::kMaxChunk - 1) / gt::kMaxChunk);
    const int64_t per = (dim + g.n_chunks - 1) / g.n_chunks;
    g.chunk = (int)((per + 31) / 32 * 32);                       // multiple of 32: halves stay 8-row swizzle groups
    g.pieces = g.chunk > 256 ? 2 : 1;
    g.m_tiles = (int)((cols + 255) / 256);
    g.stage_bytes = gt::kATile + g.chunk / 2 * gt::kBK * 2;
    g.stages = gt::kSmemBudget / g.stage_bytes;
    if (g.stages > 8) g.stages = 8;
    const int64_t kcs = (rows + gt::kBK - 1) / gt::kBK;
    const int64_t slots = kNumSMs / 2;
    int64_t best = 1;
    double best_cost = 1e30;
    for (int64_t n = 1; n <= 16 && n <= kcs; ++n) {
        const int64_t waves = ((int64_t)g.m_tiles * g.n_chunks * n + slots - 1) / slots;
        // unit = one 64-deep K step of a 256 x chunk tile pair (~0.4 us); writing + re-reading one fp32 image of the result
        // costs ~cols * dim * 8 B / 6 TB/s locally, ~cols * dim * 4 B / 0.6 TB/s over NVLink
        const double image_us = (double)cols * (double)dim * (scatter ? 4.0 / 0.6e6 : 8.0 / 6.0e6);
        const double cost = (double)waves * ((double)((kcs + n - 1) / n) + 12.0) + (double)n * image_us / 0.4;   // ~12 chunks of fill/drain
        if (cost < best_cost) { best_cost = cost; best = n; }
    }
    g.k_split = (int)best;
    return g;
}

}  // namespace dcb

extern "C" int dcb_clip_gt_splits(int64_t rows, int64_t cols, int64_t dim) {
    return dcb::clip_gt_plan(rows, cols, dim).k_split;
}
extern "C" int dcb_clip_gt_splits_scatter(int64_t rows, int64_t cols, int64_t dim) {
    return dcb::clip_gt_plan(rows, cols, dim, true).k_split;
}

namespace dcb {
// rows = K extent, cols = M extent.  a_kmajor = false: G is [rows][cols] (A = G^T MN-major), B = a_hat^T [dim][rows].
// a_kmajor = true: G is [cols][rows] (A = G K-major), B = b_hat^T in blocks of b_block_cols K columns ([dim * n_blocks][b_block_cols]).
static int clip_gt_launch(const void* g, int64_t g_pitch_elems, const void* a_hat_t, int64_t at_pitch_elems, int64_t rows,
                          int64_t cols, int64_t dim, float* acc_parts, void* const* dest, int n_dest, int src_slot,
                          void* stream, bool a_kmajor = false, int64_t b_block_cols = 0) {
    DCB_REQUIRE(g && a_hat_t && (acc_parts || n_dest > 0), "NULL pointer argument");
    DCB_REQUIRE(rows >= 1 && cols >= 1 && dim >= 8 && dim % 8 == 0, "bad shape");
    DCB_REQUIRE(g_pitch_elems >= (a_kmajor ? rows : cols) && g_pitch_elems % 8 == 0, "G pitch must be >= its column count and a multiple of 8 elements");
    if (b_block_cols <= 0 || b_block_cols >= rows) b_block_cols = rows;
    DCB_REQUIRE(b_block_cols == rows || (a_kmajor && b_block_cols % 64 == 0 && rows % b_block_cols == 0),
                "B blocks must hold a multiple of 64 K columns and divide the K extent");
    DCB_REQUIRE(at_pitch_elems >= b_block_cols && at_pitch_elems % 8 == 0, "B^T pitch must be >= its column count and a multiple of 8 elements");
    const GtPlan plan = clip_gt_plan(rows, cols, dim, n_dest > 0);
    CUtensorMap map_g, map_at;
    if (a_kmajor) {
        if (tc::encode_tile_map_16bit(&map_g, g, cols, rows, (uint64_t)g_pitch_elems * 2, 128)) return 1;
    } else {
        if (tc::encode_tile_map_16bit(&map_g, g, rows, cols, (uint64_t)g_pitch_elems * 2, 64)) return 1;
    }
    const int64_t n_blocks = rows / b_block_cols;
    if (tc::encode_tile_map_16bit(&map_at, a_hat_t, dim * n_blocks, b_block_cols, (uint64_t)at_pitch_elems * 2, plan.chunk / plan.pieces / 2)) return 1;
    ClipGtParams p{};
    p.b_block_cols = n_blocks > 1 ? (int)b_block_cols : 0;
    p.b_block_rows = (int)dim;
    p.acc = acc_parts;
    p.n_dest = n_dest;
    p.src_slot = src_slot;
    p.rows_per_dest = n_dest > 0 ? cols / n_dest : 0;
    for (int i = 0; i < n_dest; ++i) p.dest[i] = static_cast<float*>(dest[i]);
    p.rows = (int)rows;
    p.cols = (int)cols;
    p.dim = (int)dim;
    p.chunk = plan.chunk;
    p.pieces = plan.pieces;
    p.n_chunks = plan.n_chunks;
    p.m_tiles = plan.m_tiles;
    p.k_split = plan.k_split;
    p.stages = plan.stages;
    p.stage_bytes = plan.stage_bytes;
    // fp16 x fp16 -> fp32, M = 256 over the pair, A (= G^T) MN-major (bit 15), B K-major
    const uint32_t idesc = tc::umma_idesc_f16(256, plan.chunk / plan.pieces, 0) | (a_kmajor ? 0u : (1u << 15));
    const int smem = 1024 + plan.stages * plan.stage_bytes + 8 * (2 * plan.stages + 1) + 16;
    static int max_set[2] = {0, 0};
    if (smem > max_set[a_kmajor]) {
        if (a_kmajor) DCB_CUDA_OK(cudaFuncSetAttribute(clip_gt_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        else DCB_CUDA_OK(cudaFuncSetAttribute(clip_gt_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        max_set[a_kmajor] = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * plan.m_tiles * plan.n_chunks * plan.k_split));
    cfg.blockDim = dim3(gt::kThreads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (a_kmajor) DCB_CUDA_OK(cudaLaunchKernelEx(&cfg, clip_gt_gemm_kernel<true>, map_g, map_at, p, idesc));
    else DCB_CUDA_OK(cudaLaunchKernelEx(&cfg, clip_gt_gemm_kernel<false>, map_g, map_at, p, idesc));
    return 0;
}
}  // namespace dcb

// Split backward, a side: acc_parts[s][i, :] = sum_{j in K split s} G[i, j] 2^k b_hat[j, :] from the tiles dcb_clip_g_tiles stored.
// b_hat_t: fp16 [n_blocks][dim][bt_pitch_elems] with bt_block_cols valid columns per block (one block per source rank; 0 = one
// block of `cols` columns).  acc_parts: dcb_clip_rg_splits(...) buffers of [rows_local, dim] fp32 (same 2^k scale).
extern "C" int dcb_clip_rg_splits(int64_t rows_local, int64_t cols, int64_t dim) {
    return dcb::clip_gt_plan(cols, rows_local, dim).k_split;
}
extern "C" int dcb_clip_row_grads_from_g(const void* g, int64_t g_pitch_elems, const void* b_hat_t, int64_t bt_pitch_elems,
                                         int64_t bt_block_cols, int64_t rows_local, int64_t cols, int64_t dim, float* acc_parts,
                                         void* stream) {
    return dcb::clip_gt_launch(g, g_pitch_elems, b_hat_t, bt_pitch_elems, cols, rows_local, dim, acc_parts, nullptr, 0, 0, stream,
                               true, bt_block_cols);
}

extern "C" int dcb_clip_col_grads_from_g(const void* g, int64_t g_pitch_elems, const void* a_hat_t, int64_t at_pitch_elems,
                                         int64_t rows, int64_t cols, int64_t dim, float* acc_parts, void* stream) {
    return dcb::clip_gt_launch(g, g_pitch_elems, a_hat_t, at_pitch_elems, rows, cols, dim, acc_parts, nullptr, 0, 0, stream);
}

extern "C" int dcb_clip_col_grads_scatter(const void* g, int64_t g_pitch_elems, const void* a_hat_t, int64_t at_pitch_elems,
                                          int64_t rows, int64_t cols, int64_t dim, void* const* dest_parts, int n_dest,
                                          int src_slot, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(dest_parts && n_dest >= 1 && n_dest <= kGtMaxDest && cols % n_dest == 0 && src_slot >= 0 && src_slot < n_dest,
                "scatter: 1..%d destinations owning equal row ranges, src_slot in range", kGtMaxDest);
    for (int i = 0; i < n_dest; ++i) DCB_REQUIRE(dest_parts[i], "scatter: NULL destination %d", i);
    return clip_gt_launch(g, g_pitch_elems, a_hat_t, at_pitch_elems, rows, cols, dim, nullptr, dest_parts, n_dest, src_slot, stream);
}
