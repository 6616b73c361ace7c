// The small kernels of the fused contrastive pipeline, one launch per stage instead of one per quantity:
//
//   clip_prep_kernel     inverse L2 norms of up to four [rows, D] embedding matrices (reference clip_model.py:37-38), the raw
//                        rows copied into the exchange buffer peers read from, and the normalised fp16 transposes the
//                        gradient GEMMs take as K-major operands -- one launch (was: inv_norm + 2 x transpose_norm).
//   clip_post1_kernel    after the similarity tiles: row statistics of the local rows from the per-split partial sums,
//                        per-row losses, unit gradient coefficients, and this rank's column sums -- written straight into the
//                        statistics slot of EVERY rank (peer-mapped pointers, NVLink stores): the all-reduce of the column
//                        statistics without a collective (was: combine + colreduce + NCCL all-reduce + all-gather).
//   clip_post2_kernel    after one cross-rank barrier: sums the per-source slots in a fixed order (deterministic, identical
//                        on every rank), finishes the t2i direction for all columns, the two loss values of the GLOBAL batch
//                        with the scale / percent weighting of _loss.py:231-234, the column coefficients and the bounds that
//                        fix the fp16 scale of the gradient tiles (was: colfinish + loss + 2 scalar all-reduces + 2 x coef).
//   clip_finish2_kernel  both towers' normalisation Jacobians and label terms in one launch (was: 2 x grad_finish).
//
// The gradient coefficients are stored WITHOUT the upstream gradients ("unit" coefficients x_i = 1/(2 B A_i),
// y_i = T/(2 Zs_i), z_i = T/(2 Zt_i)); the backward kernels read the upstream scalars from the device (clip_shared.cuh),
// so nothing between forward and backward depends on them and the backward needs no exchange before its first GEMM.
#include "clip_shared.cuh"

namespace dcb {

// ---------------------------------------------------------------------------------------------
// prep
// ---------------------------------------------------------------------------------------------
struct ClipPrepParams {
    const void* x[4];        // [rows, dim] 16-bit
    float* inv[4];           // [rows]
    void* copy[4];           // optional raw copy [rows, dim]
    __half* tr[4];           // optional fp16 [dim, tr_pitch]: (x * inv).T
    long long tr_pitch[4];
    int rows, dim;
};

// grid = (d splits, matrices, row blocks of 64): every CTA computes the norms of its 64 rows and transposes the 64 x 64 tiles
// d0 = 64 (blockIdx.x + k gridDim.x).  One split (each CTA walks all d tiles, rows read once) when the row blocks alone fill
// the machine; small batches split the d tiles over several CTAs of the same rows -- scheduled back to back (x is the fastest
// grid dimension), so the re-read rows come from L2 -- to get enough CTAs (B_local = 1024: 64 CTAs took 40 us; and one split
// per d tile at B = 32768 re-read the rows 12 times from DRAM: 0.27 ms, ncu r02).  Split 0 writes the norms and the raw copy.
template <typename T>
__global__ void __launch_bounds__(256) clip_prep_kernel(const __grid_constant__ ClipPrepParams p) {
    __shared__ float rinv[64];
    __shared__ float tile[64][65];
    const int m = blockIdx.y;
    const int r0 = blockIdx.z * 64;
    const bool first = blockIdx.x == 0;
    __half* __restrict__ out = p.tr[m];
    if (!first && (!out || (int)blockIdx.x * 64 >= p.dim)) return;             // block-uniform
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const T* __restrict__ x = static_cast<const T*>(p.x[m]);
    T* __restrict__ cp = first ? static_cast<T*>(p.copy[m]) : nullptr;
    for (int rr = warp; rr < 64; rr += 8) {
        const int row = r0 + rr;
        float acc = 0.f;
        if (row < p.rows) {
            const T* __restrict__ src = x + (size_t)row * p.dim;
            for (int d = lane * 8; d < p.dim; d += 256) {
                uint4 raw = *reinterpret_cast<const uint4*>(src + d);
                float a, b;
                unpack2<T>(raw.x, a, b); acc = fmaf(a, a, acc); acc = fmaf(b, b, acc);
                unpack2<T>(raw.y, a, b); acc = fmaf(a, a, acc); acc = fmaf(b, b, acc);
                unpack2<T>(raw.z, a, b); acc = fmaf(a, a, acc); acc = fmaf(b, b, acc);
                unpack2<T>(raw.w, a, b); acc = fmaf(a, a, acc); acc = fmaf(b, b, acc);
                if (cp) *reinterpret_cast<uint4*>(cp + (size_t)row * p.dim + d) = raw;
            }
        }
        acc = warp_sum(acc);
        const float r = 1.0f / sqrtf(acc);
        if (lane == 0) {
            rinv[rr] = r;
            if (first && row < p.rows) p.inv[m][row] = r;
        }
    }
    if (!out) return;                                        // block-uniform
    __syncthreads();
    const long long pitch = p.tr_pitch[m];
    const int tx = lane, ty = warp;                          // 32 x 8
    for (int d0 = blockIdx.x * 64; d0 < p.dim; d0 += gridDim.x * 64) {
        for (int k = ty; k < 64; k += 8) {
            const int j = r0 + k, d = d0 + 2 * tx;
            float v[2] = {0.f, 0.f};
            if (j < p.rows && d < p.dim) {                    // dim even
                load_vec<T, 2>(x + (size_t)j * p.dim + d, v);
                v[0] *= rinv[k];
                v[1] *= rinv[k];
            }
            tile[k][2 * tx] = v[0];
            tile[k][2 * tx + 1] = v[1];
        }
        __syncthreads();
        for (int k = ty; k < 64; k += 8) {
            const int d = d0 + k, j = r0 + 2 * tx;
            if (d < p.dim && j < pitch) {                     // pitch even; columns >= rows are written as zeros
                const float o[2] = {j < p.rows ? tile[2 * tx][k] : 0.f, j + 1 < p.rows ? tile[2 * tx + 1][k] : 0.f};
                store_vec<__half, 2>(out + (size_t)d * pitch + j, o);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// statistics slot exchanged between ranks (floats): [4][cols] column sums | [rows_per_rank] S_ii of the source's rows |
// (8-byte aligned) 5 doubles {sum CE_i2t, sum KL_i2t / T^2, sum relu(T_ii - S_ii), sum_{i != j} relu(S_ij - T_ij),
// sum (S_ij - T_ij)^2 over the source's rows} | 4 floats {max x, max y, max z, 0} | 2 floats padding
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline long long clip_slot_tail(long long cols, long long rows_per_rank) {
    return (4 * cols + rows_per_rank + 3) / 4 * 4;
}
__host__ __device__ inline long long clip_slot_floats_(long long cols, long long rows_per_rank) {
    return clip_slot_tail(cols, rows_per_rank) + 16;
}

constexpr int kMaxRanks = 16;

struct ClipPost1Params {
    const float* ws;          // [n_part][4][rows]
    const float* diag;        // [rows]
    const float* col_part;    // [row_blocks][4][cols]
    const float* ws_extra;    // optional [n_part][2][rows]: relu(S - T) and (S - T)^2 row sums (cos_diff / logits_mse)
    const float* diag_t;      // optional [rows] T_ii
    float* stats;             // [5][rows]
    float* coef_row;          // [4][rows] unit coefficients x, y, z and the flag [T_ii > S_ii] of the cos_diff label term
    float* dest[kMaxRanks];   // this source's slot in every destination rank's buffer
    double* block_part;       // [grid][5]
    float* block_max;         // [grid][4]
    unsigned int* ticket;
    int n_dest, rows, cols, n_part, row_blocks, n_stats;
    long long tail_off;
    float temperature, inv_batch;
    int has_teacher;
};

__device__ __forceinline__ float block_max_f(float v, float* smem8) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) smem8[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? smem8[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
    }
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(256) clip_post1_kernel(const __grid_constant__ ClipPost1Params p) {
    __shared__ float smax[8];
    __shared__ bool is_last;
    const int tid = threadIdx.x;
    const long long t = (long long)blockIdx.x * 256 + tid;
    // ---- rows: four threads per row (one per statistic), fixed summation order over the partial sets
    const long long r = t >> 2;
    const int k = (int)(t & 3);
    const bool ok = r < p.rows;
    double acc = 0.0;
    if (ok) {
        const float* __restrict__ src = p.ws + (size_t)k * p.rows + r;
        const size_t stride = (size_t)4 * p.rows;
        int s = 0;
        for (; s + 4 <= p.n_part; s += 4) {
            const float a0 = src[(size_t)s * stride], a1 = src[(size_t)(s + 1) * stride];
            const float a2 = src[(size_t)(s + 2) * stride], a3 = src[(size_t)(s + 3) * stride];
            acc += (double)a0;
            acc += (double)a1;
            acc += (double)a2;
            acc += (double)a3;
        }
        for (; s < p.n_part; ++s) acc += (double)src[(size_t)s * stride];
        p.stats[(size_t)k * p.rows + r] = (float)acc;
    }
    const float mine = (float)acc;
    const unsigned quad = tid & 28u;
    const float v0 = __shfl_sync(0xffffffffu, mine, quad), v1 = __shfl_sync(0xffffffffu, mine, quad + 1);
    const float v2 = __shfl_sync(0xffffffffu, mine, quad + 2), v3 = __shfl_sync(0xffffffffu, mine, quad + 3);
    double ce = 0.0, kl = 0.0, pos = 0.0, neg = 0.0, mse = 0.0;
    float mx = 0.f, my = 0.f, mz = 0.f;
    if (ok && k == 0) {
        const float dg = p.diag[r];
        p.stats[(size_t)4 * p.rows + r] = dg;
        ce = 1.0 + log((double)v0) - (double)dg;
        mx = 0.5f * p.inv_batch / v0;
        p.coef_row[r] = mx;
        for (int d = 0; d < p.n_dest; ++d) p.dest[d][4 * (size_t)p.cols + r] = dg;
    } else if (ok && (k == 2 || k == 3)) {
        // cos_diff / logits_mse: threads 2 and 3 of the row sum the relu(S - T) resp. (S - T)^2 partial sets (fixed order)
        float flag = 0.f;
        if (p.ws_extra) {
            const float* __restrict__ src = p.ws_extra + (size_t)(k - 2) * p.rows + r;
            const size_t stride = (size_t)2 * p.rows;
            double a = 0.0;
            for (int s = 0; s < p.n_part; ++s) a += (double)src[(size_t)s * stride];
            if (k == 2) {
                const float ds = p.diag[r], dt = p.diag_t[r];
                pos = (double)fmaxf(dt - ds, 0.f);                           // clip_cos_diff.py:17-18
                neg = a - (double)fmaxf(ds - dt, 0.f);                       // get_neg_element drops exactly the diagonal (:5-8)
                flag = dt > ds ? 1.f : 0.f;
            } else {
                mse = a;
            }
        }
        if (k == 2) p.coef_row[(size_t)3 * p.rows + r] = flag;
    }
    if (ok && k == 1) {
        if (p.has_teacher) {
            kl = clip_row_kl_d((double)v1, (double)v2, (double)v3, (double)p.temperature);
            const float zs = v2 + v1 - v3 / p.temperature;          // Zs = Zt + Q - W/T (slot 1 carries Q, clip_fwd.cu)
            my = 0.5f * p.temperature / zs;
            mz = 0.5f * p.temperature / v2;
        }
        p.coef_row[(size_t)p.rows + r] = my;
        p.coef_row[(size_t)2 * p.rows + r] = mz;
    }
    // ---- columns: this rank's share of the column sums, to every rank's slot
    if (t < p.cols) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            double a = 0.0;
            if (kk < p.n_stats) {
                const float* __restrict__ src = p.col_part + (size_t)kk * p.cols + t;
                const size_t stride = (size_t)4 * p.cols;
                int rb = 0;
                for (; rb + 4 <= p.row_blocks; rb += 4) {
                    const float a0 = src[(size_t)rb * stride], a1 = src[(size_t)(rb + 1) * stride];
                    const float a2 = src[(size_t)(rb + 2) * stride], a3 = src[(size_t)(rb + 3) * stride];
                    a += (double)a0;
                    a += (double)a1;
                    a += (double)a2;
                    a += (double)a3;
                }
                for (; rb < p.row_blocks; ++rb) a += (double)src[(size_t)rb * stride];
            }
            const float f = (float)a;
            for (int d = 0; d < p.n_dest; ++d) p.dest[d][(size_t)kk * p.cols + t] = f;
        }
    }
    // ---- block partial sums -> last block -> tail of every slot
    double sums[5] = {ce, kl, pos, neg, mse};
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        sums[q] = block_sum(sums[q]);
        __syncthreads();
    }
    mx = block_max_f(mx, smax);
    my = block_max_f(my, smax);
    mz = block_max_f(mz, smax);
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < 5; ++q) p.block_part[5 * blockIdx.x + q] = sums[q];
        p.block_max[4 * blockIdx.x] = mx;
        p.block_max[4 * blockIdx.x + 1] = my;
        p.block_max[4 * blockIdx.x + 2] = mz;
        __threadfence();
        is_last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const volatile double* bp = p.block_part;
    const volatile float* bm = p.block_max;
    double t2[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    float x2 = 0.f, y2 = 0.f, z2 = 0.f;
    for (unsigned i = tid; i < gridDim.x; i += 256) {          // fixed assignment, fixed tree
#pragma unroll
        for (int q = 0; q < 5; ++q) t2[q] += bp[5 * i + q];
        x2 = fmaxf(x2, bm[4 * i]);
        y2 = fmaxf(y2, bm[4 * i + 1]);
        z2 = fmaxf(z2, bm[4 * i + 2]);
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        t2[q] = block_sum(t2[q]);
        __syncthreads();
    }
    x2 = block_max_f(x2, smax);
    y2 = block_max_f(y2, smax);
    z2 = block_max_f(z2, smax);
    if (tid == 0) {
        for (int d = 0; d < p.n_dest; ++d) {
            double* td = reinterpret_cast<double*>(p.dest[d] + p.tail_off);
#pragma unroll
            for (int q = 0; q < 5; ++q) td[q] = t2[q];
            float* tm = p.dest[d] + p.tail_off + 10;
            tm[0] = x2;
            tm[1] = y2;
            tm[2] = z2;
            tm[3] = 0.f;
        }
        *p.ticket = 0u;
    }
}

struct ClipPost2Params {
    const float* slots;       // [n_src][slot_floats]
    long long slot_floats, tail_off;
    float* col_stats;         // [4][cols]
    float* coef_col;          // [3][cols]
    float* bounds;            // [6]
    float* out;               // [9] = hard, soft, hard s_hard, soft s_soft, total, cos_diff, logits_mse, cos s_cos, mse s_mse;
                              // total = p_hard out[2] + p_soft out[3] + p_cos out[7] + p_mse out[8]
    double* block_part;
    float* block_max;
    unsigned int* ticket;
    int n_src, rows_per_src, cols;
    float temperature, inv_batch;
    int has_teacher;
    float p_hard, p_soft, s_hard, s_soft, p_cos, p_mse, s_cos, s_mse;
};

__global__ void __launch_bounds__(256) clip_post2_kernel(const __grid_constant__ ClipPost2Params p) {
    __shared__ float smax[8];
    __shared__ bool is_last;
    const int tid = threadIdx.x;
    const long long j = (long long)blockIdx.x * 256 + tid;
    double ce = 0.0, kl = 0.0;
    float mx = 0.f, my = 0.f, mz = 0.f;
    if (j < p.cols) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double a = 0.0;
            for (int s = 0; s < p.n_src; ++s) a += (double)p.slots[(size_t)s * p.slot_floats + (size_t)k * p.cols + j];
            v[k] = (float)a;
            p.col_stats[(size_t)k * p.cols + j] = v[k];
        }
        const long long src = j / p.rows_per_src;
        const float dg = p.slots[(size_t)src * p.slot_floats + 4 * (size_t)p.cols + (j - src * p.rows_per_src)];
        ce = 1.0 + log((double)v[0]) - (double)dg;
        mx = 0.5f * p.inv_batch / v[0];
        if (p.has_teacher) {
            kl = clip_row_kl_d((double)v[1], (double)v[2], (double)v[3], (double)p.temperature);
            const float zs = v[2] + v[1] - v[3] / p.temperature;
            my = 0.5f * p.temperature / zs;
            mz = 0.5f * p.temperature / v[2];
        }
        p.coef_col[j] = mx;
        p.coef_col[(size_t)p.cols + j] = my;
        p.coef_col[(size_t)2 * p.cols + j] = mz;
    }
    ce = block_sum(ce);
    __syncthreads();
    kl = block_sum(kl);
    __syncthreads();
    mx = block_max_f(mx, smax);
    my = block_max_f(my, smax);
    mz = block_max_f(mz, smax);
    if (tid == 0) {
        p.block_part[2 * blockIdx.x] = ce;
        p.block_part[2 * blockIdx.x + 1] = kl;
        p.block_max[4 * blockIdx.x] = mx;
        p.block_max[4 * blockIdx.x + 1] = my;
        p.block_max[4 * blockIdx.x + 2] = mz;
        __threadfence();
        is_last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const volatile double* bp = p.block_part;
    const volatile float* bm = p.block_max;
    double c2 = 0.0, k2 = 0.0;
    float x2 = 0.f, y2 = 0.f, z2 = 0.f;
    for (unsigned i = tid; i < gridDim.x; i += 256) {
        c2 += bp[2 * i];
        k2 += bp[2 * i + 1];
        x2 = fmaxf(x2, bm[4 * i]);
        y2 = fmaxf(y2, bm[4 * i + 1]);
        z2 = fmaxf(z2, bm[4 * i + 2]);
    }
    c2 = block_sum(c2);
    __syncthreads();
    k2 = block_sum(k2);
    __syncthreads();
    x2 = block_max_f(x2, smax);
    y2 = block_max_f(y2, smax);
    z2 = block_max_f(z2, smax);
    if (tid == 0) {
        double ce_rows = 0.0, kl_rows = 0.0, pos = 0.0, neg = 0.0, mse = 0.0;
        float rx = 0.f, ry = 0.f, rz = 0.f;
        for (int s = 0; s < p.n_src; ++s) {                    // the i2t sums of every rank's rows, fixed order
            const double* td = reinterpret_cast<const double*>(p.slots + (size_t)s * p.slot_floats + p.tail_off);
            ce_rows += td[0];
            kl_rows += td[1];
            pos += td[2];
            neg += td[3];
            mse += td[4];
            const float* tm = p.slots + (size_t)s * p.slot_floats + p.tail_off + 10;
            rx = fmaxf(rx, tm[0]);
            ry = fmaxf(ry, tm[1]);
            rz = fmaxf(rz, tm[2]);
        }
        const double t2 = (double)p.temperature * (double)p.temperature;
        const float hard = (float)(0.5 * (ce_rows + c2) * (double)p.inv_batch);      // _loss.py:131
        const float soft = (float)(0.5 * (kl_rows * t2 + k2 * t2));                  // _loss.py:135-136
        p.out[0] = hard;
        p.out[1] = soft;
        p.out[2] = hard * p.s_hard;                                                  // _loss.py:233
        p.out[3] = soft * p.s_soft;
        // CLIPCosDiff and LogitsMSE are invariant under transposition: 0.5 (loss(i2t) + loss(t2i)) = loss(i2t) (_loss.py:138-145)
        const double nb = (double)p.cols;
        const float cosd = (float)(pos / nb + (p.cols > 1 ? neg / (nb * (nb - 1.0)) : 0.0));   // clip_cos_diff.py:16-23
        const float lmse = (float)(mse / (nb * nb));                                            // logits_mse.py:9-10
        p.out[5] = cosd;
        p.out[6] = lmse;
        p.out[7] = cosd * p.s_cos;
        p.out[8] = lmse * p.s_mse;
        p.out[4] = p.out[2] * p.p_hard + p.out[3] * p.p_soft + p.out[7] * p.p_cos + p.out[8] * p.p_mse;   // _loss.py:234
        p.bounds[0] = rx;
        p.bounds[1] = ry;
        p.bounds[2] = rz;
        p.bounds[3] = x2;
        p.bounds[4] = y2;
        p.bounds[5] = z2;
        *p.ticket = 0u;
    }
}

// ---------------------------------------------------------------------------------------------
// finish: grad[i,:] = r_i (acc_i - x_hat_i (x_hat_i . acc_i)),  acc_i = 2^-k sum_splits parts - (up_hard / B) y_hat_{label(i)}
// (Jacobian of x / ||x||, reference clip_model.py:37-38, plus the -delta_ij label term of the cross entropy), both towers
// in one launch (blockIdx.y = side).  16-byte accesses: dim % 8 == 0, dim <= 1024.
// ---------------------------------------------------------------------------------------------
struct ClipFinishSide {
    const float* acc;         // [n_split][rows][dim] fp32 partial sums (scaled by 2^k)
    const void* x;            // [rows][dim] this side's embeddings
    const float* x_inv;
    const void* y;            // the other side's embeddings holding the label rows
    const float* y_inv;
    void* grad;               // [rows][dim]
    long long rows, label_rows, label_offset, split_stride;
    int n_split;
    const float* cos_flag;    // optional [rows]: [T_ii > S_ii], the diagonal term of CLIPCosDiff (same pairing as the label)
};
struct ClipFinishParams {
    ClipFinishSide side[2];
    ClipUpstream up;
    const float* bounds;
    int dim;
    float inv_batch, inv_pairs;
    int extra;                // cos_diff / logits_mse terms are part of the gradient tiles
};

template <typename T, typename G, int kGroups>
__global__ void __launch_bounds__(256) clip_finish2_kernel(const __grid_constant__ ClipFinishParams p) {
    const ClipFinishSide& sd = p.side[blockIdx.y];
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= sd.rows || sd.grad == nullptr) return;
    const int lane = threadIdx.x & 31;
    float up_h, up_s;
    clip_load_upstream(p.up, up_h, up_s);
    float gbound = clip_grad_bound(p.bounds, up_h, up_s);
    if (p.extra) {
        float up_c, up_m;
        clip_load_upstream_extra(p.up, up_c, up_m);
        gbound += clip_extra_bound(up_c, up_m, p.inv_batch, p.inv_pairs);
    }
    const float unscale = 1.0f / clip_tile_scale(gbound);
    const float r = sd.x_inv[row];
    const long long gi = sd.label_offset + row;
    const bool has_label = gi < sd.label_rows;
    float lab_w = up_h;
    if (sd.cos_flag) {
        float up_c, up_m;
        clip_load_upstream_extra(p.up, up_c, up_m);
        lab_w += up_c * sd.cos_flag[row];             // d/dS_ii of relu(T_ii - S_ii) / B = -[T_ii > S_ii] / B, like the -1/B of the label
    }
    const float lab = has_label ? lab_w * p.inv_batch * sd.y_inv[gi] : 0.f;
    const int dim = p.dim;
    const T* __restrict__ xp = static_cast<const T*>(sd.x) + row * dim;
    const T* __restrict__ yp = static_cast<const T*>(sd.y) + (has_label ? gi : 0) * dim;
    G* __restrict__ gp = static_cast<G*>(sd.grad) + row * dim;
    const float* __restrict__ accp = sd.acc + (size_t)row * dim;
    float v[kGroups][8];
#pragma unroll
    for (int g = 0; g < kGroups; ++g)
#pragma unroll
        for (int e = 0; e < 8; ++e) v[g][e] = 0.f;
    for (int s = 0; s < sd.n_split; ++s) {
        const float* __restrict__ src = accp + (size_t)s * sd.split_stride;
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
            const int d = (lane + 32 * g) * 8;
            if (d < dim) {
                float lo[4], hi[4];
                load_vec<float, 4>(src + d, lo);
                load_vec<float, 4>(src + d + 4, hi);
#pragma unroll
                for (int e = 0; e < 4; ++e) { v[g][e] += lo[e]; v[g][4 + e] += hi[e]; }
            }
        }
    }
    float xv[kGroups][8];
    float dot = 0.f;
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
        const int d = (lane + 32 * g) * 8;
        if (d < dim) {
            float yv[8];
            load_vec<T, 8>(xp + d, xv[g]);
            load_vec<T, 8>(yp + d, yv);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                xv[g][e] *= r;
                v[g][e] = v[g][e] * unscale - lab * yv[e];
                dot = fmaf(xv[g][e], v[g][e], dot);
            }
        }
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
        const int d = (lane + 32 * g) * 8;
        if (d < dim) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = r * (v[g][e] - xv[g][e] * dot);
            store_vec<G, 8>(gp + d, o);
        }
    }
}

}  // namespace dcb

extern "C" {

int64_t dcb_clip_slot_floats(int64_t cols, int64_t rows_per_rank) { return dcb::clip_slot_floats_(cols, rows_per_rank); }

int dcb_clip_prep(int n_mats, const void* const* mats, float* const* inv_norm, void* const* copy_out, void* const* tr_out,
                  const int64_t* tr_pitch_elems, int64_t rows, int64_t dim, int dtype, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(n_mats >= 1 && n_mats <= 4 && rows >= 1 && dim >= 8 && dim % 8 == 0, "bad arguments");
    DCB_REQUIRE(dtype == DCB_BF16 || dtype == DCB_F16, "bf16 or fp16 embeddings only");
    ClipPrepParams p{};
    p.rows = (int)rows;
    p.dim = (int)dim;
    for (int k = 0; k < n_mats; ++k) {
        DCB_REQUIRE(mats[k] && inv_norm[k], "matrix %d: NULL pointer", k);
        DCB_REQUIRE(reinterpret_cast<uintptr_t>(mats[k]) % 16 == 0, "matrix %d: base must be 16-byte aligned", k);
        p.x[k] = mats[k];
        p.inv[k] = inv_norm[k];
        p.copy[k] = copy_out ? copy_out[k] : nullptr;
        DCB_REQUIRE(reinterpret_cast<uintptr_t>(p.copy[k]) % 16 == 0, "matrix %d: copy target must be 16-byte aligned", k);
        p.tr[k] = tr_out ? static_cast<__half*>(tr_out[k]) : nullptr;
        p.tr_pitch[k] = tr_pitch_elems ? tr_pitch_elems[k] : 0;
        if (p.tr[k])
            DCB_REQUIRE(p.tr_pitch[k] >= rows && p.tr_pitch[k] % 2 == 0 && reinterpret_cast<uintptr_t>(p.tr[k]) % 4 == 0,
                        "matrix %d: transpose pitch must be even and >= rows", k);
    }
    bool any_tr = false;
    for (int k = 0; k < n_mats; ++k) any_tr = any_tr || p.tr[k] != nullptr;
    const long long row_blocks = (rows + 63) / 64, d_tiles = (dim + 63) / 64;
    DCB_REQUIRE(row_blocks <= 65535, "too many rows for one prep launch");
    long long split = any_tr ? (2LL * kNumSMs + row_blocks * n_mats - 1) / (row_blocks * n_mats) : 1;      // ~2 CTAs per SM in total
    split = split < 1 ? 1 : (split > d_tiles ? d_tiles : split);
    const dim3 grid((unsigned)split, (unsigned)n_mats, (unsigned)row_blocks);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == DCB_BF16) clip_prep_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p);
    else clip_prep_kernel<__half><<<grid, 256, 0, st>>>(p);
    DCB_CUDA_OK(cudaGetLastError());
    return 0;
}

int dcb_clip_post1(const float* ws, int n_part, const float* diag, const float* col_part, int row_blocks, int64_t rows,
                   int64_t cols, float temperature, int has_teacher, int64_t global_batch, float* stats, float* coef_row,
                   void* const* dest_slots, int n_dest, const float* ws_extra, const float* diag_t, void* scratch, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(ws && diag && col_part && stats && coef_row && dest_slots && scratch, "NULL pointer argument");
    DCB_REQUIRE(n_dest >= 1 && n_dest <= kMaxRanks && rows >= 1 && cols >= 1 && n_part >= 1 && row_blocks >= 1, "bad arguments");
    ClipPost1Params p{};
    p.ws = ws;
    p.diag = diag;
    p.col_part = col_part;
    p.ws_extra = ws_extra;
    p.diag_t = diag_t;
    DCB_REQUIRE(!ws_extra || diag_t, "the cos_diff sums need T_ii");
    p.stats = stats;
    p.coef_row = coef_row;
    for (int d = 0; d < n_dest; ++d) {
        DCB_REQUIRE(dest_slots[d], "NULL destination slot %d", d);
        p.dest[d] = static_cast<float*>(dest_slots[d]);
    }
    p.n_dest = n_dest;
    p.rows = (int)rows;
    p.cols = (int)cols;
    p.n_part = n_part;
    p.row_blocks = row_blocks;
    p.n_stats = has_teacher ? 4 : 1;
    p.tail_off = clip_slot_tail(cols, rows);
    p.temperature = temperature;
    p.inv_batch = 1.0f / (float)global_batch;
    p.has_teacher = has_teacher;
    const long long work = 4 * rows > cols ? 4 * rows : cols;
    const unsigned grid = (unsigned)((work + 255) / 256);
    // scratch: [ticket (16 B)] [grid x 5 doubles] [grid x 4 floats]
    p.ticket = static_cast<unsigned int*>(scratch);
    p.block_part = reinterpret_cast<double*>(static_cast<char*>(scratch) + 16);
    p.block_max = reinterpret_cast<float*>(static_cast<char*>(scratch) + 16 + (size_t)grid * 40);
    clip_post1_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    DCB_CUDA_OK(cudaGetLastError());
    return 0;
}

int64_t dcb_clip_post_scratch_bytes(int64_t rows, int64_t cols) {
    const long long work = 4 * rows > cols ? 4 * rows : cols;
    return 16 + ((work + 255) / 256) * 56;
}

int dcb_clip_post2(const float* slots, int n_src, int64_t rows_per_src, int64_t cols, float temperature, int has_teacher,
                   int64_t global_batch, const float* weights8, float* col_stats,
                   float* coef_col, float* bounds, float* out, void* scratch, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(slots && col_stats && coef_col && bounds && out && scratch, "NULL pointer argument");
    DCB_REQUIRE(n_src >= 1 && n_src <= kMaxRanks && rows_per_src >= 1 && cols >= 1 && cols == n_src * rows_per_src, "bad arguments");
    ClipPost2Params p{};
    p.slots = slots;
    p.slot_floats = clip_slot_floats_(cols, rows_per_src);
    p.tail_off = clip_slot_tail(cols, rows_per_src);
    p.col_stats = col_stats;
    p.coef_col = coef_col;
    p.bounds = bounds;
    p.out = out;
    p.n_src = n_src;
    p.rows_per_src = (int)rows_per_src;
    p.cols = (int)cols;
    p.temperature = temperature;
    p.inv_batch = 1.0f / (float)global_batch;
    p.has_teacher = has_teacher;
    DCB_REQUIRE(weights8, "NULL weights");       // host array {p_hard, p_soft, s_hard, s_soft, p_cos, p_mse, s_cos, s_mse}
    p.p_hard = weights8[0];
    p.p_soft = weights8[1];
    p.s_hard = weights8[2];
    p.s_soft = weights8[3];
    p.p_cos = weights8[4];
    p.p_mse = weights8[5];
    p.s_cos = weights8[6];
    p.s_mse = weights8[7];
    const unsigned grid = (unsigned)((cols + 255) / 256);
    p.ticket = static_cast<unsigned int*>(scratch);
    p.block_part = reinterpret_cast<double*>(static_cast<char*>(scratch) + 16);
    p.block_max = reinterpret_cast<float*>(static_cast<char*>(scratch) + 16 + (size_t)grid * 16);
    clip_post2_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    DCB_CUDA_OK(cudaGetLastError());
    return 0;
}

int dcb_clip_finish2(const float* acc_a, int n_split_a, int64_t split_stride_a, const void* a, const float* a_inv, void* grad_a,
                     int64_t rows_a, const void* a_label, const float* a_label_inv, int64_t a_label_rows, int64_t a_label_offset,
                     const float* acc_b, int n_split_b, int64_t split_stride_b, const void* b, const float* b_inv, void* grad_b,
                     int64_t rows_b, const void* b_label, const float* b_label_inv, int64_t b_label_rows, int64_t b_label_offset,
                     int64_t dim, int64_t global_batch, const float* const* g5, const float* w8, const float* cos_flag,
                     const float* bounds, int in_dtype, int grad_dtype, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(bounds && dim >= 8 && dim % 8 == 0 && dim <= 1024 && global_batch >= 1, "bad arguments (dim %% 8 == 0, dim <= 1024)");
    DCB_REQUIRE(in_dtype == DCB_BF16 || in_dtype == DCB_F16, "bf16 or fp16 embeddings only");
    ClipFinishParams p{};
    DCB_REQUIRE(g5 && w8, "NULL upstream description");
    p.side[0] = ClipFinishSide{acc_a, a, a_inv, a_label, a_label_inv, grad_a, rows_a, a_label_rows, a_label_offset, split_stride_a, n_split_a, cos_flag};
    p.side[1] = ClipFinishSide{acc_b, b, b_inv, b_label, b_label_inv, grad_b, rows_b, b_label_rows, b_label_offset, split_stride_b, n_split_b, cos_flag};
    long long max_rows = 0;
    for (int s = 0; s < 2; ++s) {
        const ClipFinishSide& sd = p.side[s];
        if (!sd.grad) continue;
        DCB_REQUIRE(sd.acc && sd.x && sd.x_inv && sd.y && sd.y_inv && sd.rows >= 1 && sd.n_split >= 1, "side %d: bad arguments", s);
        DCB_REQUIRE(((reinterpret_cast<uintptr_t>(sd.acc) | reinterpret_cast<uintptr_t>(sd.x) | reinterpret_cast<uintptr_t>(sd.y) |
                      reinterpret_cast<uintptr_t>(sd.grad)) % 16) == 0 && sd.split_stride % 4 == 0, "side %d: 16-byte alignment required", s);
        max_rows = sd.rows > max_rows ? sd.rows : max_rows;
    }
    DCB_REQUIRE(max_rows >= 1, "nothing to do");
    p.up = clip_upstream_from(g5, w8);
    p.bounds = bounds;
    p.dim = (int)dim;
    p.inv_batch = 1.0f / (float)global_batch;
    p.inv_pairs = global_batch > 1 ? (float)(1.0 / ((double)global_batch * (double)(global_batch - 1))) : 0.f;
    p.extra = cos_flag != nullptr;
    const dim3 grid((unsigned)((max_rows + 7) / 8), 2);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int groups = (int)((dim / 8 + 31) / 32);
    return dispatch_in_grad(in_dtype, grad_dtype, [&](auto tt, auto gg) -> int {
        using T = decltype(tt);
        using G = decltype(gg);
        if constexpr (sizeof(T) == 2) {
            if (groups == 1) clip_finish2_kernel<T, G, 1><<<grid, 256, 0, st>>>(p);
            else if (groups == 2) clip_finish2_kernel<T, G, 2><<<grid, 256, 0, st>>>(p);
            else if (groups == 3) clip_finish2_kernel<T, G, 3><<<grid, 256, 0, st>>>(p);
            else clip_finish2_kernel<T, G, 4><<<grid, 256, 0, st>>>(p);
            DCB_CUDA_OK(cudaGetLastError());
            return 0;
        }
        return fail("finish2: 16-bit embeddings only");
    });
}

}  // extern "C"
