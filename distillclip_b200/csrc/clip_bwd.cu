// Backward of the fused contrastive / logit-KL losses for one direction: gradient w.r.t. the normalised
// a-side rows, recomputing the logits tile by tile (they are never stored).
//
// Autograd equivalent in the reference: backward through SoftLabel / HardLabel (model/loss_component/soft_label.py:11-16,
// hard_label.py:10-12), the 0.5*(i2t + t2i) sum (model/_loss.py:130-137) and `image_feature @ text_feature.t()`
// (model/component/clip_model.py:40).  With row statistics (A, Zs, Zt) of this direction and of the opposite
// direction (the "column" statistics of the same logit matrix), d total / d S_ij is
//   G_ij = e1_ij (alpha_i + alpha'_j) + es_ij (beta_i + beta'_j) - et_ij (gamma_i + gamma'_j)  - [i == j] gh / B
//   alpha = gh/(2 B A)   beta = gs T/(2 Zs)   gamma = gs T/(2 Zt)      (prepared by clip_coef_kernel)
// and d total / d a_hat_i = sum_j G_ij b_hat_j.  The delta term is added in fp32 by dcb_clip_grad_finish.
//
// Per CTA: a block of 128 a-side rows and one chunk (<= 256 columns) of the embedding dimension.  Loop over
// 64-wide column tiles:  tcgen05.mma S,T (M=128,N=64) -> epilogue warps turn the accumulators into the fp16 tile
// G_ij * 2^k, written to shared memory in the K-major 128B-swizzle layout -> tcgen05.mma acc[128 x Dc] += G * b_hatT
// (b_hatT = normalised student b-side, transposed to fp16 once per call, so both operands are K-major).  fp16 keeps
// 11 significant bits of G (bf16: 8 -> ~1.3e-3 gradient error, measured); the power-of-two scale 2^k, derived on the
// device from the largest possible |G| (max row coefficient + max column coefficient), keeps G in fp16's normal range.
// S/T accumulators are double buffered in TMEM (2 x 128 columns) next to the gradient accumulator (<= 256 columns).
//
// Warp roles (256 threads): warp 0 = TMA ring producer, warp 1 = TMEM alloc + MMA issuer, warp 2 = bT producer,
// warp 3 idle, warps 4-7 = epilogue.
#include "tc_common.cuh"

namespace dcb {

namespace bwd {
constexpr int kBM = 128, kBN = 64, kBK = 64, kUmmaK = 16;
constexpr int kStages = 3;
constexpr int kATile = kBM * kBK * 2;      // 16 KiB
constexpr int kBTile = kBN * kBK * 2;      //  8 KiB
constexpr int kStageBytes = 2 * kATile + 2 * kBTile;   // 48 KiB
constexpr int kGBytes = kBM * kBN * 2;     // 16 KiB
constexpr int kMaxDc = 256;
constexpr int kBtBytes = kMaxDc * kBN * 2;  // 32 KiB
constexpr int kThreads = 256;
constexpr int kTmemCols = 512;
constexpr int kAccCol = 256;
constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kGBytes + kBtBytes + 2 * 5 * kBN * 4 + 256;
}  // namespace bwd

struct ClipBwdParams {
    const float* a_inv_stu;
    const float* b_inv_stu;
    const float* a_inv_tea;
    const float* b_inv_tea;
    const float* coef_row;    // [3][rows]  alpha, beta, gamma of this direction's rows
    const float* coef_col;    // [3][cols]  alpha', beta', gamma' (opposite direction, all columns)
    const float* gmax_row;    // [1] max_i (|alpha_i| + |beta_i| + |gamma_i|) over this direction's rows (all ranks)
    const float* gmax_col;    // [1] same for the opposite direction
    float* acc;               // [n_split][rows][dim] fp32 partial gradients w.r.t. a_hat
    int rows, cols, dim;
    int dc;                   // columns of the embedding dimension per CTA (multiple of 16, <= 256)
    int n_split, col_tiles;
    float inv_temp;
};

// 2^k with 2^k * gmax <= 2^14: G * 2^k stays inside fp16's normal range with headroom for the accumulate
__device__ __forceinline__ float grad_tile_scale(float gmax) {
    if (!(gmax > 0.f) || !isfinite(gmax)) return 1.f;
    int e;
    frexpf(gmax, &e);                 // gmax = m * 2^e, m in [0.5, 1)
    return ldexpf(1.f, 14 - e);
}

__device__ __forceinline__ float ex2b(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool kTeacher>
__global__ void __launch_bounds__(bwd::kThreads, 1)
clip_bwd_kernel(const __grid_constant__ CUtensorMap map_a_stu, const __grid_constant__ CUtensorMap map_b_stu,
                const __grid_constant__ CUtensorMap map_a_tea, const __grid_constant__ CUtensorMap map_b_tea,
                const __grid_constant__ CUtensorMap map_bt, const __grid_constant__ ClipBwdParams p,
                const uint32_t idesc_st, const uint32_t idesc_grad) {
    using namespace bwd;
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t ring = smem_base;
    const uint32_t g_smem = ring + kStages * kStageBytes;
    uint8_t* g_gen = smem_gen + kStages * kStageBytes;
    const uint32_t bt_smem = g_smem + kGBytes;
    float* scale_buf = reinterpret_cast<float*>(smem_gen + kStages * kStageBytes + kGBytes + kBtBytes);   // [2][5][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(scale_buf) + 2 * 5 * kBN * 4);
    const uint32_t bar_full = smem_u32(bars);                  // [kStages]
    const uint32_t bar_empty = bar_full + 8 * kStages;         // [kStages]
    const uint32_t bar_stfull = bar_empty + 8 * kStages;       // [2]
    const uint32_t bar_stempty = bar_stfull + 16;              // [2]
    const uint32_t bar_gfull = bar_stempty + 16;
    const uint32_t bar_gempty = bar_gfull + 8;
    const uint32_t bar_btfull = bar_gempty + 8;
    const uint32_t bar_accfull = bar_btfull + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rb = blockIdx.x / p.n_split, sp = blockIdx.x % p.n_split;
    const int d0 = blockIdx.y * p.dc;
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
            mbar_init(bar_stfull + 8 * s, 1);
            mbar_init(bar_stempty + 8 * s, 4);
        }
        mbar_init(bar_gfull, 4);
        mbar_init(bar_gempty, 1);
        mbar_init(bar_btfull, 1);
        mbar_init(bar_accfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------------------------------------------------------- operand ring for the S / T recompute
        if (elect_one()) {
            tma_prefetch_desc(&map_a_stu);
            tma_prefetch_desc(&map_b_stu);
            int stage = 0;
            uint32_t phase = 0;
            for (int t = 0; t < n_tiles; ++t) {
                const int col0 = (tile_begin + t) * kBN;
                for (int kc = 0; kc < n_kc; ++kc) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t dst = ring + stage * kStageBytes;
                    const uint32_t full = bar_full + 8 * stage;
                    mbar_arrive_expect_tx(full, kTeacher ? 2 * (kATile + kBTile) : (kATile + kBTile));
                    tma_load_2d(dst, &map_a_stu, full, kc * kBK, row0);
                    tma_load_2d(dst + kATile, &map_b_stu, full, kc * kBK, col0);
                    if (kTeacher) {
                        tma_load_2d(dst + kATile + kBTile, &map_a_tea, full, kc * kBK, row0);
                        tma_load_2d(dst + 2 * kATile + kBTile, &map_b_tea, full, kc * kBK, col0);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 2) {
        // ---------------------------------------------------------------- bT tiles for the gradient GEMM
        if (elect_one()) {
            tma_prefetch_desc(&map_bt);
            for (int t = 0; t < n_tiles; ++t) {
                mbar_wait(bar_gempty, (t & 1) ^ 1);           // previous gradient MMAs have consumed the buffer
                mbar_arrive_expect_tx(bar_btfull, p.dc * kBN * 2);
                tma_load_2d(bt_smem, &map_bt, bar_btfull, (tile_begin + t) * kBN, d0);
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer: warp-uniform waits, one elected lane issues
        int stage = 0;
        uint32_t phase = 0;
        auto issue_st = [&](int t) {
            const int as = t & 1;
            mbar_wait(bar_stempty + 8 * as, ((t >> 1) & 1) ^ 1);
            tc_fence_after_sync();
            const uint32_t acc_s = tmem_base + as * 128, acc_t = acc_s + 64;
            for (int kc = 0; kc < n_kc; ++kc) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after_sync();
                const uint32_t src = ring + stage * kStageBytes;
                if (elect_one()) {
                    const uint64_t da_s = umma_desc_k_sw128(src), db_s = umma_desc_k_sw128(src + kATile);
                    const uint64_t da_t = umma_desc_k_sw128(src + kATile + kBTile);
                    const uint64_t db_t = umma_desc_k_sw128(src + 2 * kATile + kBTile);
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        const uint32_t accum = (kc > 0 || k > 0) ? 1u : 0u;
                        umma_f16(acc_s, da_s + 2 * k, db_s + 2 * k, idesc_st, accum);
                        if (kTeacher) umma_f16(acc_t, da_t + 2 * k, db_t + 2 * k, idesc_st, accum);
                    }
                    umma_commit(bar_empty + 8 * stage);
                    if (kc == n_kc - 1) umma_commit(bar_stfull + 8 * as);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        };
        issue_st(0);
        for (int t = 0; t < n_tiles; ++t) {
            if (t + 1 < n_tiles) issue_st(t + 1);          // overlaps the epilogue of tile t
            mbar_wait(bar_gfull, t & 1);
            mbar_wait(bar_btfull, t & 1);
            tc_fence_after_sync();
            if (elect_one()) {
                const uint64_t dg = umma_desc_k_sw128(g_smem), dbt = umma_desc_k_sw128(bt_smem);
#pragma unroll
                for (int k = 0; k < kBN / kUmmaK; ++k)
                    umma_f16(tmem_base + kAccCol, dg + 2 * k, dbt + 2 * k, idesc_grad, (t > 0 || k > 0) ? 1u : 0u);
                umma_commit(bar_gempty);
                if (t == n_tiles - 1) umma_commit(bar_accfull);
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- epilogue
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int grow = row0 + r;
        const bool row_ok = grow < p.rows;
        const float LOG2E = 1.4426950408889634f;
        const float r_s = row_ok ? __ldg(p.a_inv_stu + grow) : 0.f;
        const float r_t = (kTeacher && row_ok) ? __ldg(p.a_inv_tea + grow) : 0.f;
        const float k1 = r_s * LOG2E, k1t = r_s * LOG2E * p.inv_temp, k2t = r_t * LOG2E * p.inv_temp;
        const float n1 = -LOG2E, n1t = -LOG2E * p.inv_temp;
        const float ra = row_ok ? __ldg(p.coef_row + grow) : 0.f;
        const float rbeta = (kTeacher && row_ok) ? __ldg(p.coef_row + p.rows + grow) : 0.f;
        const float rg = (kTeacher && row_ok) ? __ldg(p.coef_row + 2 * (size_t)p.rows + grow) : 0.f;
        const float gscale = grad_tile_scale(__ldg(p.gmax_row) + __ldg(p.gmax_col));
        for (int t = 0; t < n_tiles; ++t) {
            const int as = t & 1;
            const int col0 = (tile_begin + t) * kBN;
            float* sc = scale_buf + as * 5 * kBN;          // [c_stu][c_tea][alpha'][beta'][gamma'] x 64
            if (r < kBN) {
                const int c = col0 + r;
                const bool ok = c < p.cols;
                sc[r] = ok ? __ldg(p.b_inv_stu + c) : 0.f;
                sc[2 * kBN + r] = ok ? __ldg(p.coef_col + c) : 0.f;
                if (kTeacher) {
                    sc[kBN + r] = ok ? __ldg(p.b_inv_tea + c) : 0.f;
                    sc[3 * kBN + r] = ok ? __ldg(p.coef_col + p.cols + c) : 0.f;
                    sc[4 * kBN + r] = ok ? __ldg(p.coef_col + 2 * (size_t)p.cols + c) : 0.f;
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait(bar_stfull + 8 * as, (t >> 1) & 1);
            tc_fence_after_sync();
            const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 128;
            uint32_t packed[kBN / 2];
#pragma unroll
            for (int ch = 0; ch < kBN / 32; ++ch) {
                float sv[32], tv[32];
                tmem_ld_32x32(lane_addr + ch * 32, sv);
                if (kTeacher) tmem_ld_32x32(lane_addr + 64 + ch * 32, tv);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    float g2[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int cc = ch * 32 + c + e;
                        const float cs = sc[cc];
                        const float u = sv[c + e] * cs;
                        float g = ex2b(fmaf(u, k1, n1)) * (ra + sc[2 * kBN + cc]);
                        if (kTeacher) {
                            const float v = tv[c + e] * sc[kBN + cc];
                            g = fmaf(ex2b(fmaf(u, k1t, n1t)), rbeta + sc[3 * kBN + cc], g);
                            g = fmaf(-ex2b(fmaf(v, k2t, n1t)), rg + sc[4 * kBN + cc], g);
                        }
                        g2[e] = g * gscale;      // columns >= cols meet all-zero rows of b_hatT, so they need no mask
                    }
                    packed[(ch * 32 + c) >> 1] = pack2<__half>(g2[0], g2[1]);
                }
            }
            // S/T accumulators of this stage are in registers now
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_stempty + 8 * as);
            // wait until the gradient MMAs of the previous tile have finished reading the G buffer
            mbar_wait(bar_gempty, (t & 1) ^ 1);
#pragma unroll
            for (int c8 = 0; c8 < kBN / 8; ++c8) {
                uint4 w = make_uint4(packed[4 * c8], packed[4 * c8 + 1], packed[4 * c8 + 2], packed[4 * c8 + 3]);
                *reinterpret_cast<uint4*>(g_gen + sw128_chunk_offset(r, c8)) = w;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_gfull);
        }
        // gradient accumulator -> global partial buffer
        mbar_wait(bar_accfull, 0);
        tc_fence_after_sync();
        const uint32_t acc_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kAccCol;
        float* out = p.acc + ((size_t)sp * p.rows + (row_ok ? grow : 0)) * p.dim + d0;
        for (int ch = 0; ch * 32 < p.dc; ++ch) {
            float v[32];
            tmem_ld_32x32(acc_addr + ch * 32, v);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    const int d = d0 + ch * 32 + c;
                    if (d + 3 < p.dim) {
                        *reinterpret_cast<float4*>(out + ch * 32 + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                    } else {
                        for (int e = 0; e < 4; ++e)
                            if (d + e < p.dim) out[ch * 32 + c + e] = v[c + e];
                    }
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after_sync();
        tc::tmem_dealloc(tmem_base, kTmemCols);
    }
}

static int clip_bwd_dc(int64_t dim) {
    int64_t d16 = (dim + 15) / 16 * 16;
    return (int)(d16 < bwd::kMaxDc ? d16 : bwd::kMaxDc);
}
static int clip_bwd_splits(int64_t rows, int64_t cols, int64_t dim) {
    const int64_t row_blocks = (rows + bwd::kBM - 1) / bwd::kBM;
    const int64_t col_tiles = (cols + bwd::kBN - 1) / bwd::kBN;
    const int dc = clip_bwd_dc(dim);
    const int64_t chunks = (dim + dc - 1) / dc;
    int64_t n = (2 * kNumSMs + row_blocks * chunks - 1) / (row_blocks * chunks);
    if (n > col_tiles) n = col_tiles;
    if (n > 16) n = 16;
    if (n < 1) n = 1;
    return (int)n;
}

}  // namespace dcb

extern "C" int64_t dcb_clip_grad_workspace_bytes(int64_t rows_local, int64_t cols, int64_t dim) {
    if (rows_local < 1 || cols < 1 || dim < 1) return 0;
    return (int64_t)dcb::clip_bwd_splits(rows_local, cols, dim) * rows_local * dim * (int64_t)sizeof(float);
}
extern "C" int dcb_clip_grad_splits(int64_t rows_local, int64_t cols, int64_t dim) {
    return dcb::clip_bwd_splits(rows_local, cols, dim);
}

extern "C" int dcb_clip_row_grads(const void* stu_a, const void* stu_b, const void* tea_a, const void* tea_b,
                                  const void* stu_b_t, int64_t bt_pitch_elems,
                                  const float* stu_a_inv, const float* stu_b_inv, const float* tea_a_inv,
                                  const float* tea_b_inv, const float* coef_row, const float* coef_col,
                                  const float* gmax_row, const float* gmax_col, int64_t rows_local, int64_t cols, int64_t dim, int dtype, float temperature,
                                  float* acc_parts, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(stu_a && stu_b && stu_b_t && stu_a_inv && stu_b_inv && coef_row && coef_col && gmax_row && gmax_col && acc_parts,
                "NULL pointer argument");
    DCB_REQUIRE(dtype == DCB_BF16 || dtype == DCB_F16, "the fused contrastive kernel takes bf16 or fp16 embeddings");
    DCB_REQUIRE(rows_local >= 1 && cols >= 1 && dim >= 8 && dim % 8 == 0, "bad shape");
    DCB_REQUIRE(bt_pitch_elems >= cols && bt_pitch_elems % 8 == 0, "bT pitch must be >= cols and a multiple of 8 elements");
    const bool teacher = tea_a != nullptr;
    if (teacher) DCB_REQUIRE(tea_b && tea_a_inv && tea_b_inv && temperature > 0.f, "teacher arguments incomplete");
    const int dc = clip_bwd_dc(dim);
    CUtensorMap ma_s, mb_s, ma_t, mb_t, mbt;
    const uint64_t pitch = (uint64_t)dim * 2;
    if (tc::encode_tile_map_16bit(&ma_s, stu_a, rows_local, dim, pitch, bwd::kBM)) return 1;
    if (tc::encode_tile_map_16bit(&mb_s, stu_b, cols, dim, pitch, bwd::kBN)) return 1;
    if (teacher) {
        if (tc::encode_tile_map_16bit(&ma_t, tea_a, rows_local, dim, pitch, bwd::kBM)) return 1;
        if (tc::encode_tile_map_16bit(&mb_t, tea_b, cols, dim, pitch, bwd::kBN)) return 1;
    } else {
        ma_t = ma_s;
        mb_t = mb_s;
    }
    if (tc::encode_tile_map_16bit(&mbt, stu_b_t, dim, cols, (uint64_t)bt_pitch_elems * 2, dc)) return 1;
    ClipBwdParams p{};
    p.a_inv_stu = stu_a_inv;
    p.b_inv_stu = stu_b_inv;
    p.a_inv_tea = tea_a_inv;
    p.b_inv_tea = tea_b_inv;
    p.coef_row = coef_row;
    p.coef_col = coef_col;
    p.gmax_row = gmax_row;
    p.gmax_col = gmax_col;
    p.acc = acc_parts;
    p.rows = (int)rows_local;
    p.cols = (int)cols;
    p.dim = (int)dim;
    p.dc = dc;
    p.n_split = clip_bwd_splits(rows_local, cols, dim);
    p.col_tiles = (int)((cols + bwd::kBN - 1) / bwd::kBN);
    p.inv_temp = teacher ? 1.0f / temperature : 1.0f;
    const int row_blocks = (int)((rows_local + bwd::kBM - 1) / bwd::kBM);
    const int chunks = (int)((dim + dc - 1) / dc);
    const uint32_t idesc_st = tc::umma_idesc_f16(bwd::kBM, bwd::kBN, dtype == DCB_BF16 ? 1 : 0);
    const uint32_t idesc_grad = tc::umma_idesc_f16(bwd::kBM, dc, 0);     // G and b_hatT are always fp16
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((unsigned)(row_blocks * p.n_split), (unsigned)chunks);
    if (teacher) {
        static const cudaError_t attr_true = cudaFuncSetAttribute(clip_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::kSmemBytes);
        DCB_CUDA_OK(attr_true);     // set once per process (not a stream operation; kept out of graph captures)
        clip_bwd_kernel<true><<<grid, bwd::kThreads, bwd::kSmemBytes, st>>>(ma_s, mb_s, ma_t, mb_t, mbt, p, idesc_st, idesc_grad);
    } else {
        static const cudaError_t attr_false = cudaFuncSetAttribute(clip_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::kSmemBytes);
        DCB_CUDA_OK(attr_false);     // set once per process (not a stream operation; kept out of graph captures)
        clip_bwd_kernel<false><<<grid, bwd::kThreads, bwd::kSmemBytes, st>>>(ma_s, mb_s, ma_t, mb_t, mbt, p, idesc_st, idesc_grad);
    }
    DCB_CUDA_OK(cudaGetLastError());
    return 0;
}
