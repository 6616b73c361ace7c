// Small kernels around the fused contrastive path: inverse L2 norms, the one-off normalised b-side transpose,
// loss values from the row statistics, gradient coefficients, and the normalisation Jacobian.
#include "common.cuh"

namespace dcb {

// ---------------------------------------------------------------------------------------------
// r_i = 1 / ||x_i||_2   (reference model/component/clip_model.py:37-38 divides by x.norm(dim=1))
// ---------------------------------------------------------------------------------------------
struct InvNormParams {
    const void* x[8];
    float* out[8];
    long long rows[8];
};

// one launch for up to 8 matrices (blockIdx.y selects the matrix); 16-byte loads when the rows allow it
template <typename T>
__global__ void __launch_bounds__(256) inv_norm_kernel(const __grid_constant__ InvNormParams p, int dim, int vec_ok) {
    const long long rows = p.rows[blockIdx.y];
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const T* __restrict__ x = static_cast<const T*>(p.x[blockIdx.y]) + row * dim;
    float acc = 0.f;
    constexpr int kPer = 16 / (int)sizeof(T);
    if (vec_ok) {
        for (int d = lane * kPer; d < dim; d += 32 * kPer) {
            float v[kPer];
            load_vec<T, kPer>(x + d, v);
#pragma unroll
            for (int e = 0; e < kPer; ++e) acc = fmaf(v[e], v[e], acc);
        }
    } else {
        for (int d = lane; d < dim; d += 32) {
            const float v = Elem<T>::to_f(x[d]);
            acc = fmaf(v, v, acc);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) p.out[blockIdx.y][row] = 1.0f / sqrtf(acc);
}

// ---------------------------------------------------------------------------------------------
// out[d][j] = fp16(in[j][d] * inv_norm[j])   in: [rows, dim] row-major, out: [dim, pitch] row-major (pitch >= rows)
// (normalised entries are bounded by 1, so fp16 cannot overflow and keeps 11 significant bits)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) transpose_norm_f16_kernel(const T* __restrict__ in, const float* __restrict__ inv_norm,
                                                                 __half* __restrict__ out, int rows, int dim, long long pitch) {
    __shared__ float tile[32][33];
    const int j0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        const int j = j0 + k, d = d0 + tx;
        tile[k][tx] = (j < rows && d < dim) ? Elem<T>::to_f(in[(long long)j * dim + d]) * inv_norm[j] : 0.f;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int d = d0 + k, j = j0 + tx;
        if (d < dim && j < pitch) out[(long long)d * pitch + j] = __float2half_rn(j < rows ? tile[tx][k] : 0.f);
    }
}

// 64 x 64 tiles, two elements per access (dim even): a warp reads / writes 128 contiguous bytes per instruction
template <typename T>
__global__ void __launch_bounds__(256) transpose_norm_f16_x2_kernel(const T* __restrict__ in, const float* __restrict__ inv_norm,
                                                                    __half* __restrict__ out, int rows, int dim, long long pitch) {
    __shared__ float tile[64][65];
    const int j0 = blockIdx.x * 64, d0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    for (int k = ty; k < 64; k += 8) {
        const int j = j0 + k, d = d0 + 2 * tx;
        float v[2] = {0.f, 0.f};
        if (j < rows && d < dim) {                               // dim even: d + 1 < dim as well
            load_vec<T, 2>(in + (long long)j * dim + d, v);
            const float r = inv_norm[j];
            v[0] *= r;
            v[1] *= r;
        }
        tile[k][2 * tx] = v[0];
        tile[k][2 * tx + 1] = v[1];
    }
    __syncthreads();
    for (int k = ty; k < 64; k += 8) {
        const int d = d0 + k, j = j0 + 2 * tx;
        if (d < dim && j < pitch) {                              // pitch even: j + 1 < pitch as well; columns >= rows are zero
            const float o[2] = {j < rows ? tile[2 * tx][k] : 0.f, j + 1 < rows ? tile[2 * tx + 1][k] : 0.f};
            store_vec<__half, 2>(out + (long long)d * pitch + j, o);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Loss values of both directions from the per-row losses written by the forward kernel's combine step
// (rowloss[0][i] = CE_i = 1 + log A_i - S_ii, rowloss[1][i] = KL_i / T^2 = W_i/(T Zt_i) + log(Zs_i/Zt_i), double):
// sums[0..3] (double) = {sum CE i2t, sum CE t2i, T^2 sum KL i2t, T^2 sum KL t2i} over this rank's rows;
// out[0] = 0.5 (sums0 + sums1) / global_batch, out[1] = 0.5 (sums2 + sums3)   (_loss.py:131,135-136)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) clip_loss_kernel(const double* __restrict__ rl_i2t, const double* __restrict__ rl_t2i,
                                                         int rows_i2t, int rows_t2i, float temperature, int has_teacher,
                                                         double inv_batch, double* __restrict__ sums, float* __restrict__ out) {
    __shared__ double res[4];
    for (int dir = 0; dir < 2; ++dir) {
        const double* rl = dir == 0 ? rl_i2t : rl_t2i;
        const int rows = dir == 0 ? rows_i2t : rows_t2i;
        double ce = 0.0, kl = 0.0;
        int i = threadIdx.x;
        for (; i + 3 * (int)blockDim.x < rows; i += 4 * blockDim.x) {       // four independent loads in flight (fixed order)
            const double c0 = rl[i], c1 = rl[i + blockDim.x], c2 = rl[i + 2 * blockDim.x], c3 = rl[i + 3 * blockDim.x];
            ce += c0; ce += c1; ce += c2; ce += c3;
            if (has_teacher) {
                const double* k = rl + (size_t)rows + i;
                const double k0 = k[0], k1 = k[blockDim.x], k2 = k[2 * blockDim.x], k3 = k[3 * blockDim.x];
                kl += k0; kl += k1; kl += k2; kl += k3;
            }
        }
        for (; i < rows; i += blockDim.x) {
            ce += rl[i];
            if (has_teacher) kl += rl[(size_t)rows + i];
        }
        ce = block_sum(ce);
        __syncthreads();
        kl = block_sum(kl);
        if (threadIdx.x == 0) {
            res[dir] = ce;
            res[2 + dir] = kl * (double)temperature * (double)temperature;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        for (int k = 0; k < 4; ++k) sums[k] = res[k];
        out[0] = (float)(0.5 * (res[0] + res[1]) * inv_batch);
        out[1] = (float)(0.5 * (res[2] + res[3]));
    }
}

// ---------------------------------------------------------------------------------------------
// coef[0][i] = gh / (2 B A_i)   coef[1][i] = gs T / (2 Zs_i)   coef[2][i] = gs T / (2 Zt_i)
// upstream = {gh, gs} on the device (no host sync in backward)
// ---------------------------------------------------------------------------------------------
// gmax[0] = max_i (|coef0| + |coef1| + |coef2|): bounds |G_ij| <= gmax(rows) + gmax(cols) since every exp term is <= 1
__global__ void __launch_bounds__(256) clip_coef_kernel(const float* __restrict__ stats, int rows, float temperature,
                                                        int has_teacher, float inv_batch, const float* __restrict__ upstream,
                                                        float* __restrict__ coef, unsigned int* __restrict__ gmax) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.f;
    if (i < rows) {
        const float gh = upstream[0], gs = upstream[1];
        const float a = 0.5f * gh * inv_batch / stats[i];
        // stats[1] carries Q, the second-order part of Zs - Zt (clip_fwd.cu): Zs = Zt + Q - W/T
        const float zs = has_teacher ? stats[(size_t)2 * rows + i] + stats[(size_t)rows + i] - stats[(size_t)3 * rows + i] / temperature : 1.f;
        const float b = has_teacher ? 0.5f * gs * temperature / zs : 0.f;
        const float c = has_teacher ? 0.5f * gs * temperature / stats[(size_t)2 * rows + i] : 0.f;
        coef[i] = a;
        coef[(size_t)rows + i] = b;
        coef[(size_t)2 * rows + i] = c;
        m = fabsf(a) + fabsf(b) + fabsf(c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(gmax, __float_as_uint(m));     // non-negative floats order like uints
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float clip_grad_tile_scale(float gmax) {     // must match grad_tile_scale in clip_bwd.cu
    if (!(gmax > 0.f) || !isfinite(gmax)) return 1.f;
    int e;
    frexpf(gmax, &e);
    return ldexpf(1.f, 14 - e);
}

// grad_a[i,:] = r_i (acc_i - a_hat_i (a_hat_i . acc_i)),  acc_i = sum_splits acc_parts / 2^k - (gh/B) b_hat_{offset+i}
// (Jacobian of x / ||x||, reference clip_model.py:37-38, plus the -delta_ij label term of cross entropy)
// ---------------------------------------------------------------------------------------------
constexpr int kFinishMaxPerLane = 24;     // rows up to 768 wide stay in registers (one pass over acc_parts)

// 16-byte version for 16-bit embeddings with dim % 8 == 0, dim <= 1024: lane l owns the 8-element groups l, l + 32, ...
// (one uint4 of a, one of the label row, two float4 per accumulator split, one 16-byte / two float4 stores per group).
template <typename T, typename G, int kGroups>
__global__ void __launch_bounds__(256) clip_grad_finish_vec_kernel(const float* __restrict__ acc_parts, int n_split,
                                                                   const T* __restrict__ a, const float* __restrict__ a_inv,
                                                                   const T* __restrict__ b, const float* __restrict__ b_inv,
                                                                   long long rows, long long cols, int dim, long long row_offset,
                                                                   float inv_batch, const float* __restrict__ upstream,
                                                                   const float* __restrict__ gmax_row, const float* __restrict__ gmax_col,
                                                                   G* __restrict__ grad) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float r = a_inv[row];
    const float unscale = 1.0f / clip_grad_tile_scale(gmax_row[0] + gmax_col[0]);
    const long long gi = row_offset + row;
    const bool has_label = gi < cols;
    const float lab = has_label ? upstream[0] * inv_batch * b_inv[gi] : 0.f;
    const T* __restrict__ ap = a + row * dim;
    const T* __restrict__ bp = b + (has_label ? gi : 0) * dim;
    G* __restrict__ gp = grad + row * dim;
    const size_t split_stride = (size_t)rows * dim;
    const float* __restrict__ accp = acc_parts + (size_t)row * dim;
    float v[kGroups][8];
#pragma unroll
    for (int g = 0; g < kGroups; ++g)
#pragma unroll
        for (int e = 0; e < 8; ++e) v[g][e] = 0.f;
    for (int s = 0; s < n_split; ++s) {
        const float* __restrict__ src = accp + (size_t)s * split_stride;
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
    float av[kGroups][8];
    float dot = 0.f;
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
        const int d = (lane + 32 * g) * 8;
        if (d < dim) {
            float bv[8];
            load_vec<T, 8>(ap + d, av[g]);
            load_vec<T, 8>(bp + d, bv);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                av[g][e] *= r;
                v[g][e] = v[g][e] * unscale - lab * bv[e];
                dot = fmaf(av[g][e], v[g][e], dot);
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
            for (int e = 0; e < 8; ++e) o[e] = r * (v[g][e] - av[g][e] * dot);
            store_vec<G, 8>(gp + d, o);
        }
    }
}

template <typename T, typename G>
__global__ void __launch_bounds__(256) clip_grad_finish_kernel(const float* __restrict__ acc_parts, int n_split,
                                                               const T* __restrict__ a, const float* __restrict__ a_inv,
                                                               const T* __restrict__ b, const float* __restrict__ b_inv,
                                                               long long rows, long long cols, int dim, long long row_offset,
                                                               float inv_batch, const float* __restrict__ upstream,
                                                               const float* __restrict__ gmax_row, const float* __restrict__ gmax_col,
                                                               G* __restrict__ grad) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float r = a_inv[row];
    const float unscale = 1.0f / clip_grad_tile_scale(gmax_row[0] + gmax_col[0]);   // the MMA accumulated G * 2^k
    const long long gi = row_offset + row;
    const bool has_label = gi < cols;
    const float lab = has_label ? upstream[0] * inv_batch * b_inv[gi] : 0.f;
    const T* __restrict__ ap = a + row * dim;
    const T* __restrict__ bp = b + (has_label ? gi : 0) * dim;
    G* __restrict__ gp = grad + row * dim;
    const size_t split_stride = (size_t)rows * dim;
    const float* __restrict__ accp = acc_parts + (size_t)row * dim;
    if (dim <= 32 * kFinishMaxPerLane) {
        float v[kFinishMaxPerLane], av[kFinishMaxPerLane];
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < kFinishMaxPerLane; ++k) v[k] = 0.f;
        for (int s = 0; s < n_split; ++s) {               // all loads of one split in flight together
            const float* __restrict__ src = accp + (size_t)s * split_stride;
#pragma unroll
            for (int k = 0; k < kFinishMaxPerLane; ++k) {
                const int d = lane + 32 * k;
                if (d < dim) v[k] += src[d];
            }
        }
#pragma unroll
        for (int k = 0; k < kFinishMaxPerLane; ++k) {
            const int d = lane + 32 * k;
            av[k] = 0.f;
            if (d < dim) {
                av[k] = Elem<T>::to_f(ap[d]) * r;
                v[k] = v[k] * unscale - lab * Elem<T>::to_f(bp[d]);
                dot = fmaf(av[k], v[k], dot);
            }
        }
        dot = warp_sum(dot);
#pragma unroll
        for (int k = 0; k < kFinishMaxPerLane; ++k) {
            const int d = lane + 32 * k;
            if (d < dim) gp[d] = Elem<G>::from_f(r * (v[k] - av[k] * dot));
        }
        return;
    }
    float dot = 0.f;
    for (int d = lane; d < dim; d += 32) {
        float v = 0.f;
        for (int s = 0; s < n_split; ++s) v += accp[(size_t)s * split_stride + d];
        v = v * unscale - lab * Elem<T>::to_f(bp[d]);
        dot = fmaf(Elem<T>::to_f(ap[d]) * r, v, dot);
    }
    dot = warp_sum(dot);
    for (int d = lane; d < dim; d += 32) {
        float v = 0.f;
        for (int s = 0; s < n_split; ++s) v += accp[(size_t)s * split_stride + d];
        v = v * unscale - lab * Elem<T>::to_f(bp[d]);
        gp[d] = Elem<G>::from_f(r * (v - Elem<T>::to_f(ap[d]) * r * dot));
    }
}

}  // namespace dcb

extern "C" {

int dcb_row_inv_norm(int n_mats, const void* const* mats, float* const* inv_norm, const int64_t* rows, int64_t dim,
                     int dtype, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(n_mats >= 1 && n_mats <= 8 && dim >= 1, "bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    InvNormParams p{};
    long long max_rows = 0;
    bool vec_ok = (dim * dtype_size(dtype)) % 16 == 0;
    for (int k = 0; k < n_mats; ++k) {
        DCB_REQUIRE(mats[k] && inv_norm[k] && rows[k] >= 1, "matrix %d: bad arguments", k);
        p.x[k] = mats[k];
        p.out[k] = inv_norm[k];
        p.rows[k] = rows[k];
        max_rows = rows[k] > max_rows ? rows[k] : max_rows;
        vec_ok = vec_ok && reinterpret_cast<uintptr_t>(mats[k]) % 16 == 0;
    }
    const dim3 grid((unsigned)((max_rows + 7) / 8), (unsigned)n_mats);
    switch (dtype) {
        case DCB_BF16: inv_norm_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p, (int)dim, vec_ok ? 1 : 0); break;
        case DCB_F16: inv_norm_kernel<__half><<<grid, 256, 0, st>>>(p, (int)dim, vec_ok ? 1 : 0); break;
        case DCB_F32: inv_norm_kernel<float><<<grid, 256, 0, st>>>(p, (int)dim, vec_ok ? 1 : 0); break;
        default: return fail("unknown dtype %d", dtype);
    }
    DCB_CUDA_OK(cudaGetLastError());
    return 0;
}

int dcb_transpose_norm_f16(const void* in, const float* inv_norm, void* out, int64_t rows, int64_t dim,
                           int64_t out_pitch_elems, int dtype, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(in && inv_norm && out && rows >= 1 && dim >= 1 && out_pitch_elems >= rows, "bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    __half* o = static_cast<__half*>(out);
    if (dim % 2 == 0 && out_pitch_elems % 2 == 0 && reinterpret_cast<uintptr_t>(in) % 4 == 0 && reinterpret_cast<uintptr_t>(out) % 4 == 0) {
        dim3 grid2((unsigned)((out_pitch_elems + 63) / 64), (unsigned)((dim + 63) / 64));
        switch (dtype) {
            case DCB_BF16: transpose_norm_f16_x2_kernel<__nv_bfloat16><<<grid2, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(in), inv_norm, o, (int)rows, (int)dim, out_pitch_elems); break;
            case DCB_F16: transpose_norm_f16_x2_kernel<__half><<<grid2, 256, 0, st>>>(static_cast<const __half*>(in), inv_norm, o, (int)rows, (int)dim, out_pitch_elems); break;
            default: return fail("transpose: bf16 or fp16 input only");
        }
        DCB_CUDA_OK(cudaGetLastError());
        return 0;
    }
    dim3 grid((unsigned)((out_pitch_elems + 31) / 32), (unsigned)((dim + 31) / 32));
    switch (dtype) {
        case DCB_BF16: transpose_norm_f16_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(in), inv_norm, o, (int)rows, (int)dim, out_pitch_elems); break;
        case DCB_F16: transpose_norm_f16_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(in), inv_norm, o, (int)rows, (int)dim, out_pitch_elems); break;
        default: return fail("transpose: bf16 or fp16 input only");
    }
    DCB_CUDA_OK(cudaGetLastError());
    return 0;
}

int dcb_clip_losses(const double* rowloss_i2t, const double* rowloss_t2i, int64_t rows_i2t, int64_t rows_t2i,
                    int64_t global_batch, float temperature, int has_teacher, double* sums, float* out, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(rowloss_i2t && rowloss_t2i && sums && out && global_batch >= 1, "bad arguments");
    clip_loss_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(rowloss_i2t, rowloss_t2i, (int)rows_i2t, (int)rows_t2i,
                                                                        temperature, has_teacher, 1.0 / (double)global_batch,
                                                                        sums, out);
    DCB_CUDA_OK(cudaGetLastError());
    return 0;
}

int dcb_clip_grad_coef(const float* stats, int64_t rows, int64_t global_batch, float temperature, int has_teacher,
                       const float* upstream, float* coef, float* gmax, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(stats && upstream && coef && gmax && rows >= 1 && global_batch >= 1, "bad arguments");
    DCB_CUDA_OK(cudaMemsetAsync(gmax, 0, sizeof(float), static_cast<cudaStream_t>(stream)));
    clip_coef_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        stats, (int)rows, temperature, has_teacher, 1.0f / (float)global_batch, upstream, coef,
        reinterpret_cast<unsigned int*>(gmax));
    DCB_CUDA_OK(cudaGetLastError());
    return 0;
}

int dcb_clip_grad_finish(const float* acc_parts, int n_split, const void* stu_a, const float* stu_a_inv,
                         const void* stu_b, const float* stu_b_inv, int64_t rows, int64_t cols, int64_t dim,
                         int64_t row_offset, int64_t global_batch, const float* upstream, const float* gmax_row,
                         const float* gmax_col, int in_dtype, void* grad_a, int grad_dtype, void* stream) {
    using namespace dcb;
    DCB_REQUIRE(acc_parts && stu_a && stu_a_inv && stu_b && stu_b_inv && upstream && gmax_row && gmax_col && grad_a,
                "NULL pointer argument");
    DCB_REQUIRE(n_split >= 1 && rows >= 1 && dim >= 1 && global_batch >= 1, "bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)((rows + 7) / 8);
    const float inv_b = 1.0f / (float)global_batch;
    const bool aligned = ((reinterpret_cast<uintptr_t>(acc_parts) | reinterpret_cast<uintptr_t>(stu_a) |
                           reinterpret_cast<uintptr_t>(stu_b) | reinterpret_cast<uintptr_t>(grad_a)) % 16) == 0;
    return dispatch_in_grad(in_dtype, grad_dtype, [&](auto tt, auto gg) -> int {
        using T = decltype(tt);
        using G = decltype(gg);
        if constexpr (sizeof(T) == 2) {
            if (aligned && dim % 8 == 0 && dim <= 1024) {
                const int groups = (int)((dim / 8 + 31) / 32);
#define DCB_FINISH_VEC(K)                                                                                                      \
    clip_grad_finish_vec_kernel<T, G, K><<<grid, 256, 0, st>>>(acc_parts, n_split, static_cast<const T*>(stu_a), stu_a_inv,   \
                                                              static_cast<const T*>(stu_b), stu_b_inv, rows, cols, (int)dim,  \
                                                              row_offset, inv_b, upstream, gmax_row, gmax_col, static_cast<G*>(grad_a))
                if (groups == 1) DCB_FINISH_VEC(1);
                else if (groups == 2) DCB_FINISH_VEC(2);
                else if (groups == 3) DCB_FINISH_VEC(3);
                else DCB_FINISH_VEC(4);
#undef DCB_FINISH_VEC
                DCB_CUDA_OK(cudaGetLastError());
                return 0;
            }
        }
        clip_grad_finish_kernel<T, G><<<grid, 256, 0, st>>>(acc_parts, n_split, static_cast<const T*>(stu_a), stu_a_inv,
                                                           static_cast<const T*>(stu_b), stu_b_inv, rows, cols, (int)dim,
                                                           row_offset, inv_b, upstream, gmax_row, gmax_col, static_cast<G*>(grad_a));
        DCB_CUDA_OK(cudaGetLastError());
        return 0;
    });
}

}  // extern "C"
