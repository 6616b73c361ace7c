// Fused similarity GEMM + softmax statistics for one direction of the CLIP contrastive / logit-KL losses.
//
// Replaces, without ever writing the B x B logits to HBM:
//   CLIPModel.forward's L2-normalise + `image_feature @ text_feature.t()`   (reference model/component/clip_model.py:36-44)
//   HardLabel.forward   CrossEntropy(mean) vs arange(B)                      (reference model/loss_component/hard_label.py:10-12)
//   SoftLabel.forward   softmax(./T), .log(), KLDiv(sum) * T^2               (reference model/loss_component/soft_label.py:11-16)
//
// A cluster of two CTAs (one TPC) owns 256 "a-side" rows, 128 per CTA, and walks a range of 128-wide "b-side" column
// tiles.  Per tile the raw bf16/fp16 embeddings stream through a TMA -> shared-memory ring (64-wide K chunks, 128 B
// swizzle) into tcgen05.mma.cta_group::2 (M=256, N=128, K=16): each CTA stages its own 128 a-rows and only HALF of the
// b-side tile (64 rows), so a K chunk costs 48 KiB of TMA traffic per SM instead of 64 -- one SM's TMA engine pulls
// at most ~78 B/clk (scripts/probe/tma_probe.cu) against the 125 B/clk a 128x128 single-CTA tile needs to keep the
// tensor pipe busy, and the kernel was bound exactly there (r01q).  The leader CTA issues the MMAs; student and teacher
// similarity accumulators sit side by side in each CTA's TMEM, double-buffered (2 x (128 + 128) = 512 columns), so the
// MMAs of tile n+1 overlap the epilogue of tile n.
// The epilogue warps read the accumulators with tcgen05.ld (one row per thread), apply the fp32 inverse norms
// (logit = acc * r_i * c_j, so normalised embeddings are never rounded to bf16) and accumulate per row
//   A  = sum_j exp(S_ij - 1)            Q  = sum_j [es_ij - et_ij + et_ij (T_ij - S_ij)/T]
//   Zt = sum_j et_ij                    W  = sum_j et_ij (T_ij - S_ij)                    and S_ii,
// es = exp((S - 1)/T), et = exp((T - 1)/T).  Q is the SECOND-ORDER part of Zs - Zt (Zs = Zt + Q - W/T): the KL of a row,
//   KL_i / T^2 = W/(T Zt) + log(Zs/Zt) = -m + log1p(m + Q/Zt),  m = -W/(T Zt),
// is a second-order quantity in (S - T) whose first-order terms cancel; carrying Q instead of Zs keeps fp32 rounding
// relative to that small quantity rather than to O(1) sums (a nearly converged student at T = 4 lost 1.4e-4 otherwise).
// Cosine logits are bounded by 1, so 1 is a valid softmax shift for every row: the sums of different column
// ranges simply add (no running max, no rescaling), which is what lets a row be split over CTAs and ranks.
//
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM allocator (+ MMA issuer in the leader), warps 2-17 = epilogue.
// Epilogue warp w reads TMEM lanes 32 (w % 4) .. and the columns [32 sub, 32 sub + 32) of the tile, sub = (w - 2) / 4,
// 16 columns at a time: four warps per scheduler hide the MUFU / FMA / shuffle latency (one warp per scheduler issued
// every ~3 cycles, ncu r01d; two reached 48 % issue utilisation with the epilogue setting the tile time, r01q).
// The four column quarters of a row are accumulated separately and added by the combine kernel.
#include <type_traits>

#include "tc_common.cuh"

namespace dcb {

namespace fwd {
constexpr int kBM = 128, kBN = 128, kBK = 64, kUmmaK = 16;
constexpr int kStages = 4;
constexpr int kTileBytes = kBM * kBK * 2;                 // 16 KiB: one [128 x 64] 16-bit a-side tile
constexpr int kBHalfRows = kBN / 2;                       // b rows staged per CTA
constexpr int kBTileBytes = kBHalfRows * kBK * 2;         // 8 KiB
constexpr int kStageBytes = 2 * kTileBytes + 2 * kBTileBytes;   // a_stu, a_tea, b_stu half, b_tea half = 48 KiB
constexpr int kThreads = 576;                            // 2 control warps + 16 epilogue warps (4 per scheduler)
constexpr int kEpiThreads = kThreads - 64;
constexpr int kSubs = 4;                                 // column quarters of a tile, one per epilogue warp of a lane quadrant
constexpr int kTmemCols = 512;
constexpr int kColBufBytes = 2 * 4 * 4 * kBN * 4;          // [tile parity][lane quadrant][stat][column] fp32
constexpr int kSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + 2 * 2 * kBN * 4 + 256 + kColBufBytes;
}  // namespace fwd

struct ClipFwdParams {
    const float* a_inv_stu;   // [rows]   1/||a_i||  student
    const float* b_inv_stu;   // [cols]
    const float* a_inv_tea;
    const float* b_inv_tea;
    float* ws;                // [kSubs * n_split][4][rows] partial sums (x kSubs: the epilogue warps of a row)
    float* diag;              // [rows] S_ii
    float* col_part;          // [row_blocks][4][cols] column sums of the same four statistics over each block of 128 rows
                              // (= the row statistics of the OPPOSITE direction, reduced later), or nullptr
    const float* rank_ref;    // optional [rows]: count the logits of row i that are > rank_ref[i] into statistics slot 1 (retrieval
                              // rank of the label for top-k accuracy; hard-label-only instantiation without column sums)
    float* dump_s;            // optional [rows, cols] raw logits (tests only), else nullptr
    float* dump_t;
    float* ws_extra;          // kExtra: [kSubs * n_split][2][rows] partial sums of relu(S - T) and (S - T)^2 over the row
                              // (CLIPCosDiff clip_cos_diff.py:16-23 and LogitsMSE logits_mse.py:9-10 from the same tiles)
    float* diag_t;            // kExtra: [rows] teacher logit T_ii
    int rows, cols, dim;
    int row_offset;           // column (relative to b-side row 0 of THIS launch) holding the label of local row 0: labels are
                              // arange(B) over the global batch, so this is (global row offset) - (first column of the chunk)
    int col_part_ld;          // columns per statistic row of col_part (= cols unless the launch covers a chunk of the columns)
    int n_split, col_tiles;
    float inv_temp;           // 1/T
};

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool kTeacher, bool kCols, bool kExtra = false>
__global__ void __launch_bounds__(fwd::kThreads, 1)
clip_fwd_kernel(const __grid_constant__ CUtensorMap map_a_stu, const __grid_constant__ CUtensorMap map_b_stu,
                const __grid_constant__ CUtensorMap map_a_tea, const __grid_constant__ CUtensorMap map_b_tea,
                const __grid_constant__ ClipFwdParams p, const uint32_t idesc) {
    using namespace fwd;
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t ring = smem_base;
    float* scale_buf = reinterpret_cast<float*>(smem_gen + kStages * kStageBytes);   // [2 buf][2 stu/tea][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_gen + kStages * kStageBytes + 2 * 2 * kBN * 4);
    float* col_buf = reinterpret_cast<float*>(smem_gen + kStages * kStageBytes + 2 * 2 * kBN * 4 + 256);   // [2][4][4][128]
    const uint32_t bar_full = smem_u32(bars);                   // [kStages] leader: TMA bytes of both CTAs
    const uint32_t bar_empty = bar_full + 8 * kStages;          // [kStages] each CTA: slot free (multicast commit)
    const uint32_t bar_tfull = bar_empty + 8 * kStages;         // [2] each CTA: accumulators complete (multicast commit)
    const uint32_t bar_tempty = bar_tfull + 16;                 // [2] leader: 8 epilogue warps of both CTAs drained them
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
        // ---------------------------------------------------------------- epilogue: 16 warps; warp (q, sub) owns TMEM lanes
        // 32 q .. 32 q + 31 (rows) and columns [32 sub, 32 sub + 32) of the tile.  The accumulators are read with the 16x256b
        // shape (the mma fragment layout, scripts/probe/tmem_layout_probe.cu): thread t = 4 g + m of the warp holds the FOUR rows
        // 16 h + 8 rr + g (h, rr in {0, 1}) and the column pairs 8 rep + 2 m + {0, 1} -- so the column sums (the statistics of the
        // opposite direction) are three in-thread additions plus a recursive halving over 8 lane groups, and the row sums a
        // halving over 4 lanes once per tile, instead of a 32-lane butterfly per statistic and 16 columns (which was half of
        // this kernel's instructions: 45 -> 25 per logit pair).
        const int q = warp & 3;                             // TMEM lane quadrant this warp may read
        const int sub = (warp - 2) >> 2;                    // column quarter of the tile handled by this warp
        const int g = lane >> 2, m = lane & 3;
        const int ep_tid = (warp - 2) * 32 + lane;          // 0..511, used to stage column scales
        constexpr int NR = kTeacher ? (kExtra ? 6 : 4) : 2; // row sums carried per row: A, Q (or the rank count), Zt, W, relu(S-T), (S-T)^2
        constexpr int NC = kTeacher ? 4 : 1;                // column sums
        const float LOG2E = 1.4426950408889634f;
        const float k1 = LOG2E, n1 = -LOG2E;
        const float2 k1t2 = make_float2(LOG2E * p.inv_temp, LOG2E * p.inv_temp), n1t2 = make_float2(-LOG2E * p.inv_temp, -LOG2E * p.inv_temp);
        const float2 neg2 = make_float2(-1.f, -1.f), invt2 = make_float2(p.inv_temp, p.inv_temp);
        const bool ranking = !kTeacher && !kCols && p.rank_ref != nullptr;
        // the four rows this thread computes on (index 2 h + rr)
        float r_s[4], r_t[4], rank_ref[4];
        bool rok[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int grow_i = row0 + q * 32 + 16 * (i >> 1) + 8 * (i & 1) + g;
            rok[i] = grow_i < p.rows;
            r_s[i] = rok[i] ? __ldg(p.a_inv_stu + grow_i) : 0.f;
            r_t[i] = (kTeacher && rok[i]) ? __ldg(p.a_inv_tea + grow_i) : 0.f;
            rank_ref[i] = (ranking && rok[i]) ? __ldg(p.rank_ref + grow_i) : 0.f;
        }
        // the one row this thread OWNS the totals of (after the per-tile halving over the 4 lanes of a group): index m
        const int grow = row0 + q * 32 + 16 * (m >> 1) + 8 * (m & 1) + g;
        const bool row_ok = grow < p.rows;
        // Row totals with Kahan compensation: the sums of each tile quarter start from zero (small magnitudes) and are folded
        // into the running totals with an error term, so Zt, W and Q keep ~1e-7 relative accuracy even for B = 32768
        float tot[NR], comp[NR];
#pragma unroll
        for (int k = 0; k < NR; ++k) tot[k] = comp[k] = 0.f;
        auto kahan = [](float& sum, float& c, float x) {
            const float y = x - c;
            const float t = sum + y;
            c = (t - sum) - y;
            sum = t;
        };
        // column partial sums of the previous tile: [4 quadrants][4 stats][128 columns] -> global, 2 values per thread
        auto flush_cols = [&](int t_prev) {
            const float* cb = col_buf + (t_prev & 1) * (4 * 4 * kBN);
            const int pcol0 = (tile_begin + t_prev) * kBN;
            if (row0 >= p.rows) return;                       // second CTA of the last cluster when the row blocks are odd
#pragma unroll
            for (int o = ep_tid; o < 4 * kBN; o += kEpiThreads) {
                const int st = o / kBN, c = o % kBN;
                if (pcol0 + c < p.cols && (kTeacher || st == 0))
                    p.col_part[((size_t)rb * 4 + st) * p.col_part_ld + pcol0 + c] =
                        (cb[(0 * 4 + st) * kBN + c] + cb[(1 * 4 + st) * kBN + c]) + (cb[(2 * 4 + st) * kBN + c] + cb[(3 * 4 + st) * kBN + c]);
            }
        };
        for (int t = 0; t < n_tiles; ++t) {
            const int as = t & 1;
            const uint32_t aphase = (t >> 1) & 1;
            const int col0 = (tile_begin + t) * kBN;
            float* sc = scale_buf + as * 2 * kBN;           // [stu 128][tea 128]
            if (ep_tid < kBN) {
                const int c = col0 + ep_tid;
                sc[ep_tid] = c < p.cols ? __ldg(p.b_inv_stu + c) : 0.f;
            } else if (kTeacher && ep_tid < 2 * kBN) {
                const int c = col0 + ep_tid - kBN;
                sc[ep_tid] = c < p.cols ? __ldg(p.b_inv_tea + c) : 0.f;
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");      // scales staged; every warp is done with tile t-1
            if (kCols && t > 0) flush_cols(t - 1);
            mbar_wait(bar_tfull + 8 * as, aphase);
            tc_fence_after_sync();
            const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 256;
            // masks are only needed where the tile meets the matrix edge, the diagonal, or rows beyond the batch (CTA-uniform)
            const int dlo = p.row_offset + row0;                // label columns of this CTA's rows: [dlo, dlo + 128)
            const bool edge = (col0 + kBN > p.cols) || (dlo < col0 + kBN && dlo + kBM > col0) ||
                              (row0 + kBM > p.rows) || (p.dump_s != nullptr) || (p.dump_t != nullptr);
            float* cb = col_buf + (t & 1) * (4 * 4 * kBN) + q * (4 * kBN);     // this quadrant's [4 stats][128 columns]
            float rowp[4][NR];                              // this tile quarter's sums of the four rows
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < NR; ++k) rowp[i][k] = 0.f;
            // 8 columns at a time: this thread's column pair is cbase + 2 m + {0, 1}.  Two instantiations of the body: the masked one
            // only for tiles on the matrix edge / the diagonal (a per-element `if (edge)` gets if-converted into predicated
            // address arithmetic for every element)
            auto piece = [&](int pc, auto edge_tag) {
                constexpr bool kEdge = decltype(edge_tag)::value;
                const int cbase = sub * 32 + pc * 8;
                const float2 cs2 = *reinterpret_cast<const float2*>(sc + cbase + 2 * m);
                const float2 ct2 = kTeacher ? *reinterpret_cast<const float2*>(sc + kBN + cbase + 2 * m) : make_float2(0.f, 0.f);
                float2 colacc[NC];                          // sums over this thread's four rows, per statistic and column
#pragma unroll
                for (int k = 0; k < NC; ++k) colacc[k] = make_float2(0.f, 0.f);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float sv[4], tv[4];                      // register 2 rr + e
                    const uint32_t a = lane_addr + (static_cast<uint32_t>(16 * h) << 16) + cbase;
                    tmem_ld_16x256b_x1(a, sv);
                    if (kTeacher) tmem_ld_16x256b_x1(a + 128, tv);
                    tmem_ld_wait();
                    if (pc == 3 && h == 1) {                 // last TMEM read of this tile: release the accumulator stage
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(l_tempty + 8 * as);
                    }
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        // the two columns of a row go through packed fp32x2 arithmetic (FMUL2 / FFMA2 / FADD2: half the issue slots)
                        const int ri = 2 * h + rr;
                        const float2 u2 = __fmul2_rn(make_float2(sv[2 * rr], sv[2 * rr + 1]), cs2);
                        const float2 s2 = make_float2(u2.x * r_s[ri], u2.y * r_s[ri]);                  // S_ij
                        float2 e1 = make_float2(ex2(fmaf(s2.x, k1, n1)), ex2(fmaf(s2.y, k1, n1)));      // exp(S - 1)
                        float2 et = make_float2(0.f, 0.f), f = et, qv = et, xr = et, xm = et, cnt = et, tl2 = et;
                        if (kTeacher) {
                            const float2 v2 = __fmul2_rn(make_float2(tv[2 * rr], tv[2 * rr + 1]), ct2);
                            tl2 = make_float2(v2.x * r_t[ri], v2.y * r_t[ri]);                          // T_ij
                            const float2 as2 = __ffma2_rn(s2, k1t2, n1t2), at2 = __ffma2_rn(tl2, k1t2, n1t2);
                            const float2 es = make_float2(ex2(as2.x), ex2(as2.y));
                            et = make_float2(ex2(at2.x), ex2(at2.y));
                            const float2 d = __ffma2_rn(s2, neg2, tl2);                                 // T - S
                            f = __fmul2_rn(et, d);
                            qv = __ffma2_rn(f, invt2, __ffma2_rn(et, neg2, es));                        // es - et + f / T
                            if constexpr (kExtra) {
                                xr = make_float2(fmaxf(-d.x, 0.f), fmaxf(-d.y, 0.f));
                                xm = __fmul2_rn(d, d);
                            }
                        } else if (ranking) {
                            // same expression as the diagonal: S_ii compares equal to itself
                            cnt = make_float2(s2.x > rank_ref[ri] ? 1.f : 0.f, s2.y > rank_ref[ri] ? 1.f : 0.f);
                        }
                        if constexpr (kEdge) {
                            const int grow_i = row0 + q * 32 + 16 * h + 8 * rr + g;
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int gc = col0 + cbase + 2 * m + e;
                                const bool ok = gc < p.cols && rok[ri];
                                const float s = e ? s2.y : s2.x, tl = e ? tl2.y : tl2.x;
                                if (ok && gc == p.row_offset + grow_i) {
                                    p.diag[grow_i] = s;
                                    if constexpr (kExtra) p.diag_t[grow_i] = tl;
                                }
                                if (ok && p.dump_s) p.dump_s[(size_t)grow_i * p.cols + gc] = s;
                                if (kTeacher && ok && p.dump_t) p.dump_t[(size_t)grow_i * p.cols + gc] = tl;
                                if (!ok) {
                                    if (e) e1.y = et.y = f.y = qv.y = xr.y = xm.y = cnt.y = 0.f;
                                    else e1.x = et.x = f.x = qv.x = xr.x = xm.x = cnt.x = 0.f;
                                }
                            }
                        }
                        rowp[ri][0] += e1.x + e1.y;
                        colacc[0] = __fadd2_rn(colacc[0], e1);
                        if (kTeacher) {
                            rowp[ri][1] += qv.x + qv.y;
                            rowp[ri][2] += et.x + et.y;
                            rowp[ri][3] += f.x + f.y;
                            if constexpr (kExtra) {
                                rowp[ri][4] += xr.x + xr.y;
                                rowp[ri][5] += xm.x + xm.y;
                            }
                            if constexpr (NC == 4) {
                                colacc[1] = __fadd2_rn(colacc[1], qv);
                                colacc[2] = __fadd2_rn(colacc[2], et);
                                colacc[3] = __fadd2_rn(colacc[3], f);
                            }
                        } else {
                            rowp[ri][1] += cnt.x + cnt.y;
                        }
                    }
                }
                if (kCols) {
                    // sum over the 8 lane groups g (lane bits 4, 3, 2): one halving step, then two plain exchanges.  Afterwards
                    // lane holds the total of column cbase + 2 m + (lane >> 4 & 1).
                    const bool up16 = lane & 16;
#pragma unroll
                    for (int k = 0; k < NC; ++k) {
                        const float keep = up16 ? colacc[k].y : colacc[k].x;
                        const float send = up16 ? colacc[k].x : colacc[k].y;
                        float w = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                        w += __shfl_xor_sync(0xffffffffu, w, 8);
                        w += __shfl_xor_sync(0xffffffffu, w, 4);
                        if (!(lane & 12)) cb[k * kBN + cbase + 2 * m + ((lane >> 4) & 1)] = w;
                    }
                }
            };
            if (edge) {
#pragma unroll 1
                for (int pc = 0; pc < 4; ++pc) piece(pc, std::true_type{});
            } else {
#pragma unroll 1
                for (int pc = 0; pc < 4; ++pc) piece(pc, std::false_type{});
            }
            // row sums of this tile quarter: halving over the 4 lanes of a group (lane bits 0, 1); lane m ends with row index m
            {
                const bool up1 = lane & 1, up2 = lane & 2;
#pragma unroll
                for (int k = 0; k < NR; ++k) {
                    float v[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float keep = up1 ? rowp[2 * h + 1][k] : rowp[2 * h][k];
                        const float send = up1 ? rowp[2 * h][k] : rowp[2 * h + 1][k];
                        v[h] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
                    }
                    const float keep = up2 ? v[1] : v[0];
                    const float send = up2 ? v[0] : v[1];
                    kahan(tot[k], comp[k], keep + __shfl_xor_sync(0xffffffffu, send, 2));
                }
            }
        }
        if (kCols && n_tiles > 0) {
            asm volatile("bar.sync 1, 512;" ::: "memory");
            flush_cols(n_tiles - 1);
        }
        if (row_ok) {
            float* w = p.ws + (size_t)(sp * kSubs + sub) * 4 * p.rows + grow;
            w[0] = tot[0];
            w[(size_t)p.rows] = tot[1];                      // Q, or the rank count of the ranking instantiation
            w[(size_t)2 * p.rows] = kTeacher ? tot[kTeacher ? 2 : 0] : 0.f;
            w[(size_t)3 * p.rows] = kTeacher ? tot[kTeacher ? 3 : 0] : 0.f;
            if constexpr (kExtra) {
                float* x = p.ws_extra + (size_t)(sp * kSubs + sub) * 2 * p.rows + grow;
                x[0] = tot[4];
                x[(size_t)p.rows] = tot[5];
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

// KL_i / T^2 from the row sums (Q, Zt, W): -m + log1p(m + Q/Zt), m = -W/(T Zt)  (see the file header)
__device__ __forceinline__ double clip_row_kl(double q, double zt, double w, double temperature) {
    const double m = -w / (temperature * zt);
    return -m + log1p(m + q / zt);
}

// stats[k][r] = sum over splits (fixed order -> deterministic), stats[4][r] = S_rr, and the row's loss terms in double
// (double precision; slot 1 of the statistics is Q, not Zs):
//   rowloss[0][r] = CE_r = 1 + log A_r - S_rr       rowloss[1][r] = KL_r / T^2 = clip_row_kl(Q_r, Zt_r, W_r)
// Four threads per row (one per statistic, 4 loads in flight each): the kernel is pure latency at B = 4096.
__global__ void __launch_bounds__(128) clip_combine_kernel(const float* __restrict__ ws, const float* __restrict__ diag,
                                                           float* __restrict__ stats, double* __restrict__ rowloss,
                                                           int rows, int n_split, float temperature, int has_teacher) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = t >> 2, k = t & 3;
    const bool ok = r < rows;
    double acc = 0.0;
    if (ok) {
        const float* __restrict__ src = ws + (size_t)k * rows + r;
        const size_t stride = (size_t)4 * rows;
        int s = 0;
        for (; s + 4 <= n_split; s += 4) {                      // fixed order, four independent loads at a time
            const float a0 = src[(size_t)s * stride], a1 = src[(size_t)(s + 1) * stride];
            const float a2 = src[(size_t)(s + 2) * stride], a3 = src[(size_t)(s + 3) * stride];
            acc += (double)a0;
            acc += (double)a1;
            acc += (double)a2;
            acc += (double)a3;
        }
        for (; s < n_split; ++s) acc += (double)src[(size_t)s * stride];
        stats[(size_t)k * rows + r] = (float)acc;
    }
    const float mine = (float)acc;                              // consumers see the fp32 statistics; so does the loss
    const unsigned quad = threadIdx.x & 28u;
    const float v0 = __shfl_sync(0xffffffffu, mine, quad), v1 = __shfl_sync(0xffffffffu, mine, quad + 1);
    const float v2 = __shfl_sync(0xffffffffu, mine, quad + 2), v3 = __shfl_sync(0xffffffffu, mine, quad + 3);
    if (!ok) return;
    if (k == 0) {
        const float dg = diag[r];
        stats[(size_t)4 * rows + r] = dg;
        rowloss[r] = 1.0 + log((double)v0) - (double)dg;
    } else if (k == 1) {
        rowloss[(size_t)rows + r] = has_teacher ? clip_row_kl((double)v1, (double)v2, (double)v3, (double)temperature) : 0.0;
    }
}

// col_stats[k][j] = sum over row blocks of col_part[rb][k][j]  (double accumulation, fixed order)
__global__ void __launch_bounds__(256) clip_colreduce_kernel(const float* __restrict__ col_part, float* __restrict__ col_stats,
                                                             int cols, int row_blocks, int n_stats) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (j >= cols) return;
    double acc = 0.0;
    if (k < n_stats)
        for (int rb = 0; rb < row_blocks; ++rb) acc += (double)col_part[((size_t)rb * 4 + k) * cols + j];
    col_stats[(size_t)k * cols + j] = (float)acc;
}

// Opposite-direction statistics of this rank's rows from the (all-reduced) column statistics:
//   stats[k][i] = col_stats[k][row_offset + i] (k < 4), stats[4][i] = S_ii, and the per-row losses as in clip_combine_kernel
__global__ void __launch_bounds__(128) clip_colfinish_kernel(const float* __restrict__ col_stats, int cols_total,
                                                             const float* __restrict__ diag, int row_offset, int rows,
                                                             float temperature, int has_teacher, float* __restrict__ stats,
                                                             double* __restrict__ rowloss) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = col_stats[(size_t)k * cols_total + row_offset + r];
        stats[(size_t)k * rows + r] = v[k];
    }
    const float dg = diag[r];
    stats[(size_t)4 * rows + r] = dg;
    rowloss[r] = 1.0 + log((double)v[0]) - (double)dg;
    rowloss[(size_t)rows + r] = has_teacher ? clip_row_kl((double)v[1], (double)v[2], (double)v[3], (double)temperature) : 0.0;
}

// Column ranges per row block: minimise (waves of 148 CTAs) x (tiles per CTA + fixed per-CTA cost)
static int clip_fwd_splits(int64_t rows, int64_t cols) {
    const int64_t row_blocks = (rows + fwd::kBM - 1) / fwd::kBM;
    const int64_t col_tiles = (cols + fwd::kBN - 1) / fwd::kBN;
    if (const char* e = getenv("DCB_DEBUG_SPLITS")) return atoi(e) > 0 ? atoi(e) : 1;     // profiling experiments only
    int64_t best = 1;
    double best_cost = 1e30;
    for (int64_t n = 1; n <= 64 && n <= col_tiles; ++n) {
        const int64_t waves = ((row_blocks + 1) / 2 * n + kNumSMs / 2 - 1) / (kNumSMs / 2);     // clusters of two row blocks
        const int64_t tiles = (col_tiles + n - 1) / n;
        const double cost = (double)waves * ((double)tiles + 1.0) + 0.01 * (double)n;
        if (cost < best_cost) { best_cost = cost; best = n; }
    }
    return (int)best;
}

}  // namespace dcb

extern "C" int64_t dcb_clip_workspace_bytes(int64_t rows_local, int64_t cols) {
    if (rows_local < 1 || cols < 1) return 0;
    const int64_t row_blocks = (rows_local + dcb::fwd::kBM - 1) / dcb::fwd::kBM;
    return (((int64_t)dcb::clip_fwd_splits(rows_local, cols) * dcb::fwd::kSubs * 4 + 1) * rows_local + row_blocks * 4 * cols) * (int64_t)sizeof(float);
}

namespace dcb {
static int clip_row_stats_impl(const void* stu_a, const void* stu_b, const void* tea_a, const void* tea_b,
                               const float* stu_a_inv, const float* stu_b_inv, const float* tea_a_inv,
                               const float* tea_b_inv, int64_t rows_local, int64_t row_offset, int64_t cols,
                               int64_t dim, int dtype, float temperature, float* stats, double* rowloss,
                               float* col_stats, void* workspace, float* dump_s, float* dump_t, const float* rank_ref,
                               void* stream, float* chunk_ws = nullptr, float* chunk_diag = nullptr, float* chunk_col_part = nullptr,
                               int64_t chunk_col_part_ld = 0, float* chunk_ws_extra = nullptr, float* chunk_diag_t = nullptr) {
    const bool chunk_mode = chunk_ws != nullptr;     // tiles only: partial sums into caller-owned buffers, no combine / colreduce
    DCB_REQUIRE(stu_a && stu_b && stu_a_inv && stu_b_inv && (chunk_mode || (stats && rowloss && workspace)), "NULL pointer argument");
    DCB_REQUIRE(dtype == DCB_BF16 || dtype == DCB_F16, "the fused contrastive kernel takes bf16 or fp16 embeddings");
    DCB_REQUIRE(rows_local >= 1 && cols >= 1 && dim >= 8 && dim % 8 == 0, "bad shape rows=%lld cols=%lld dim=%lld (dim %% 8 == 0)",
                (long long)rows_local, (long long)cols, (long long)dim);
    DCB_REQUIRE(rows_local < (1ll << 30) && cols < (1ll << 30), "batch too large");
    DCB_REQUIRE(row_offset > -(1ll << 30) && row_offset < (1ll << 30), "row offset out of range");
    const bool teacher = tea_a != nullptr;
    if (teacher) {
        DCB_REQUIRE(tea_b && tea_a_inv && tea_b_inv, "teacher pointers must be all set or all NULL");
        DCB_REQUIRE(temperature > 0.f, "temperature must be positive");
    }
    CUtensorMap ma_s, mb_s, ma_t, mb_t;
    const uint64_t pitch = (uint64_t)dim * 2;
    if (tc::encode_tile_map_16bit(&ma_s, stu_a, rows_local, dim, pitch, fwd::kBM)) return 1;
    if (tc::encode_tile_map_16bit(&mb_s, stu_b, cols, dim, pitch, fwd::kBHalfRows)) return 1;
    if (teacher) {
        if (tc::encode_tile_map_16bit(&ma_t, tea_a, rows_local, dim, pitch, fwd::kBM)) return 1;
        if (tc::encode_tile_map_16bit(&mb_t, tea_b, cols, dim, pitch, fwd::kBHalfRows)) return 1;
    } else {
        ma_t = ma_s;
        mb_t = mb_s;
    }
    ClipFwdParams p{};
    p.a_inv_stu = stu_a_inv;
    p.b_inv_stu = stu_b_inv;
    p.a_inv_tea = tea_a_inv;
    p.b_inv_tea = tea_b_inv;
    p.rows = (int)rows_local;
    p.cols = (int)cols;
    p.dim = (int)dim;
    p.row_offset = (int)row_offset;
    p.n_split = clip_fwd_splits(rows_local, cols);
    p.col_tiles = (int)((cols + fwd::kBN - 1) / fwd::kBN);
    if (chunk_mode) {
        p.ws = chunk_ws;
        p.diag = chunk_diag;
        p.col_part = chunk_col_part;
        p.col_part_ld = (int)chunk_col_part_ld;
        p.ws_extra = chunk_ws_extra;
        p.diag_t = chunk_diag_t;
        DCB_REQUIRE(!chunk_ws_extra || (teacher && chunk_diag_t), "the cos_diff / logits_mse sums need the teacher and a T_ii buffer");
        col_stats = chunk_col_part;                 // selects the column-sum instantiation below
    } else {
        p.ws = static_cast<float*>(workspace);
        p.diag = p.ws + (size_t)p.n_split * fwd::kSubs * 4 * rows_local;
        p.col_part = col_stats ? p.diag + rows_local : nullptr;
        p.col_part_ld = (int)cols;
    }
    p.dump_s = dump_s;
    p.dump_t = dump_t;
    p.rank_ref = rank_ref;
    p.inv_temp = teacher ? 1.0f / temperature : 1.0f;
    const int row_blocks = (int)((rows_local + fwd::kBM - 1) / fwd::kBM);
    const uint32_t idesc = tc::umma_idesc_f16(2 * fwd::kBM, fwd::kBN, dtype == DCB_BF16 ? 1 : 0);     // M = 256 over the CTA pair
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((unsigned)(2 * ((row_blocks + 1) / 2) * p.n_split));
#define DCB_LAUNCH_FWD(TEA, COLS, ...)                                                                                       \
    {                                                                                                                  \
        static const cudaError_t attr_ = cudaFuncSetAttribute(clip_fwd_kernel<TEA, COLS __VA_ARGS__>,                              \
                                                              cudaFuncAttributeMaxDynamicSharedMemorySize, fwd::kSmemBytes); \
        DCB_CUDA_OK(attr_);     /* set once per process (not a stream operation; kept out of graph captures) */         \
        cudaLaunchConfig_t cfg_{};                                                                                     \
        cfg_.gridDim = grid;                                                                                           \
        cfg_.blockDim = dim3(fwd::kThreads);                                                                           \
        cfg_.dynamicSmemBytes = fwd::kSmemBytes;                                                                       \
        cfg_.stream = st;                                                                                              \
        cudaLaunchAttribute attr2_[1];                                                                                 \
        attr2_[0].id = cudaLaunchAttributeClusterDimension;                                                            \
        attr2_[0].val.clusterDim.x = 2;                                                                                \
        attr2_[0].val.clusterDim.y = 1;                                                                                \
        attr2_[0].val.clusterDim.z = 1;                                                                                \
        cfg_.attrs = attr2_;                                                                                           \
        cfg_.numAttrs = 1;                                                                                             \
        DCB_CUDA_OK(cudaLaunchKernelEx(&cfg_, clip_fwd_kernel<TEA, COLS __VA_ARGS__>, ma_s, mb_s, ma_t, mb_t, p, idesc));           \
    }
    if (teacher && col_stats && p.ws_extra) DCB_LAUNCH_FWD(true, true, , true)
    else if (teacher && col_stats) DCB_LAUNCH_FWD(true, true)
    else if (teacher) DCB_LAUNCH_FWD(true, false)
    else if (col_stats) DCB_LAUNCH_FWD(false, true)
    else DCB_LAUNCH_FWD(false, false)
#undef DCB_LAUNCH_FWD
    DCB_CUDA_OK(cudaGetLastError());
    if (chunk_mode) return 0;
    clip_combine_kernel<<<(unsigned)((rows_local * 4 + 127) / 128), 128, 0, st>>>(p.ws, p.diag, stats, rowloss, (int)rows_local,
                                                                               fwd::kSubs * p.n_split, temperature, teacher ? 1 : 0);
    DCB_CUDA_OK(cudaGetLastError());
    if (col_stats) {
        dim3 g2((unsigned)((cols + 255) / 256), 4);
        clip_colreduce_kernel<<<g2, 256, 0, st>>>(p.col_part, col_stats, (int)cols, row_blocks, teacher ? 4 : 1);
        DCB_CUDA_OK(cudaGetLastError());
    }
    return 0;
}
}  // namespace dcb

extern "C" int dcb_clip_row_stats(const void* stu_a, const void* stu_b, const void* tea_a, const void* tea_b,
                                  const float* stu_a_inv, const float* stu_b_inv, const float* tea_a_inv,
                                  const float* tea_b_inv, int64_t rows_local, int64_t row_offset, int64_t cols,
                                  int64_t dim, int dtype, float temperature, float* stats, double* rowloss,
                                  float* col_stats, void* workspace, float* dump_s, float* dump_t, void* stream) {
    return dcb::clip_row_stats_impl(stu_a, stu_b, tea_a, tea_b, stu_a_inv, stu_b_inv, tea_a_inv, tea_b_inv, rows_local, row_offset,
                                    cols, dim, dtype, temperature, stats, rowloss, col_stats, workspace, dump_s, dump_t, nullptr,
                                    stream);
}

// Tiles only, for a chunk of the columns (the b-side rows of one source rank): partial row sums of the chunk into
// ws_chunk[parts][4][rows] (parts = dcb_clip_fwd_chunk_parts), S_ii into diag[rows] where the label falls inside the chunk,
// column sums of the chunk into col_part_chunk[row_blocks][4][col_part_ld] (pointer already advanced to the chunk's first
// column).  label_col0 = (global index of local row 0) - (global index of the chunk's first column).  dcb_clip_post1 reduces.
// ws_extra_chunk (optional, [parts][2][rows]) + diag_t[rows]: also the row sums of relu(S - T) and (S - T)^2 and T_ii, for
// CLIPCosDiff / LogitsMSE from the same tiles.
extern "C" int dcb_clip_fwd_chunk_parts(int64_t rows_local, int64_t cols_chunk) {
    return dcb::clip_fwd_splits(rows_local, cols_chunk) * dcb::fwd::kSubs;
}
extern "C" int dcb_clip_fwd_chunk(const void* stu_a, const void* stu_b, const void* tea_a, const void* tea_b,
                                  const float* stu_a_inv, const float* stu_b_inv, const float* tea_a_inv,
                                  const float* tea_b_inv, int64_t rows_local, int64_t label_col0, int64_t cols_chunk,
                                  int64_t dim, int dtype, float temperature, float* ws_chunk, float* diag,
                                  float* col_part_chunk, int64_t col_part_ld, float* ws_extra_chunk, float* diag_t,
                                  void* stream) {
    using namespace dcb;
    DCB_REQUIRE(ws_chunk && diag && col_part_chunk && col_part_ld >= cols_chunk, "NULL / bad chunk buffers");
    return clip_row_stats_impl(stu_a, stu_b, tea_a, tea_b, stu_a_inv, stu_b_inv, tea_a_inv, tea_b_inv, rows_local, label_col0,
                               cols_chunk, dim, dtype, temperature, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                               stream, ws_chunk, diag, col_part_chunk, col_part_ld, ws_extra_chunk, diag_t);
}

extern "C" int dcb_clip_rank_counts(const void* a, const void* b, const float* a_inv, const float* b_inv, int64_t rows_local,
                                    int64_t row_offset, int64_t cols, int64_t dim, int dtype, const float* ref, float* stats,
                                    double* rowloss, void* workspace, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(ref, "NULL reference scores");
    return clip_row_stats_impl(a, b, nullptr, nullptr, a_inv, b_inv, nullptr, nullptr, rows_local, row_offset, cols, dim, dtype,
                               1.0f, stats, rowloss, nullptr, workspace, nullptr, nullptr, ref, stream);
}

extern "C" int dcb_clip_col_finish(const float* col_stats, int64_t cols_total, const float* diag_local, int64_t row_offset,
                                   int64_t rows_local, float temperature, int has_teacher, float* stats, double* rowloss,
                                   void* stream) {
    using namespace dcb;
    DCB_REQUIRE(col_stats && diag_local && stats && rowloss && rows_local >= 1 && row_offset >= 0 &&
                    row_offset + rows_local <= cols_total, "bad arguments");
    clip_colfinish_kernel<<<(unsigned)((rows_local + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        col_stats, (int)cols_total, diag_local, (int)row_offset, (int)rows_local, temperature, has_teacher, stats, rowloss);
    DCB_CUDA_OK(cudaGetLastError());
    return 0;
}
