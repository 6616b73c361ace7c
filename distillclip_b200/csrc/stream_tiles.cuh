// Tile bodies of the streaming losses, shared by the per-family kernels (mse_stream.cu, attn_kl.cu) and the
// single-launch tower kernel (tower_stream.cu).  One tile = what one 256-thread CTA processes per loop iteration.
#pragma once
#include "common.cuh"

namespace dcb {

constexpr int kStreamThreads = 256;
constexpr int kMseUnroll = 4;

// ---------------------------------------------------------------------------------------------
// Elementwise-difference tile: kStreamThreads * kMseUnroll * VEC consecutive elements.
//   L1 == false (nn.MSELoss):  returns sum (s-t)^2,  writes g = (s-t) * grad_coef
//   L1 == true  (nn.L1Loss):   returns sum |s-t|,    writes g = sign(s-t) * grad_coef   (sign(0) = 0, as ATen)
// ---------------------------------------------------------------------------------------------
template <bool L1> __device__ __forceinline__ void diff_op(float d, float gc, float& acc, float& g) {
    if constexpr (L1) {
        acc += fabsf(d);
        g = d > 0.f ? gc : (d < 0.f ? -gc : (d == 0.f ? 0.f : d));     // NaN propagates
    } else {
        acc = fmaf(d, d, acc);
        g = d * gc;
    }
}

template <typename T, typename G, int VEC, bool L1 = false>
__device__ __forceinline__ float mse_tile(const T* __restrict__ s, const T* __restrict__ t, G* __restrict__ g,
                                          long long rem, float gc, int tid) {
    constexpr int kTile = kStreamThreads * kMseUnroll * VEC;
    float acc = 0.f;
    if (rem >= kTile) {
        float sv[kMseUnroll][VEC], tv[kMseUnroll][VEC];
#pragma unroll
        for (int u = 0; u < kMseUnroll; ++u) load_vec<T, VEC>(s + (u * kStreamThreads + tid) * VEC, sv[u]);
#pragma unroll
        for (int u = 0; u < kMseUnroll; ++u) load_vec<T, VEC>(t + (u * kStreamThreads + tid) * VEC, tv[u]);
#pragma unroll
        for (int u = 0; u < kMseUnroll; ++u) {
            float gv[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) diff_op<L1>(sv[u][e] - tv[u][e], gc, acc, gv[e]);
            if (g) store_vec<G, VEC>(g + (u * kStreamThreads + tid) * VEC, gv);
        }
    } else {
#pragma unroll 1
        for (int u = 0; u < kMseUnroll; ++u) {
            const long long i0 = (long long)(u * kStreamThreads + tid) * VEC;
            if (i0 + VEC <= rem) {
                float sv[VEC], tv[VEC], gv[VEC];
                load_vec<T, VEC>(s + i0, sv);
                load_vec<T, VEC>(t + i0, tv);
#pragma unroll
                for (int e = 0; e < VEC; ++e) diff_op<L1>(sv[e] - tv[e], gc, acc, gv[e]);
                if (g) store_vec<G, VEC>(g + i0, gv);
            } else {
                for (long long i = i0; i < rem; ++i) {
                    float gv1;
                    diff_op<L1>(Elem<T>::to_f(s[i]) - Elem<T>::to_f(t[i]), gc, acc, gv1);
                    if (g) g[i] = Elem<G>::from_f(gv1);
                }
            }
        }
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// Attention-KL tile: kStreamThreads groups of VEC consecutive positions of one sample, all heads.
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC, int H>
__device__ __forceinline__ void head_sum(const T* __restrict__ p, long long stride, int h_rt, float (&acc)[VEC]) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
    if constexpr (H > 0) {
        float v[H][VEC];
#pragma unroll
        for (int h = 0; h < H; ++h) load_vec<T, VEC>(p + h * stride, v[h]);
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] += v[h][e];
    } else {
#pragma unroll 4
        for (int h = 0; h < h_rt; ++h) {
            float v[VEC];
            load_vec<T, VEC>(p + h * stride, v);
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] += v[e];
        }
    }
}

struct AttnShape {
    long long groups;        // batch * positions / VEC
    long long groups_per_b;  // positions / VEC
    long long positions;
    int hs, ht;
    float inv_hs, inv_ht;
};

// H = compile-time head count of both maps (0 = runtime head counts).  gi = group index handled by this thread.
//   MSE == false: KL(sum) of the head means (attention_probs_kl.py:15-20)
//   MSE == true : squared error of the head means (attention_probs_mse.py:13-20, attention_score_mse.py:13-20);
//                 value = sum (s_mean - t_mean)^2, gradient = (s_mean - t_mean) * gc for every student head
template <typename T, typename G, int VEC, int H, bool MSE = false>
__device__ __forceinline__ float attn_tile(const T* __restrict__ s_base, const T* __restrict__ t_base, G* __restrict__ g_base,
                                           const AttnShape& sh, long long gi, float gc) {
    if (gi >= sh.groups) return 0.f;
    const long long P = sh.positions;
    const long long b = gi / sh.groups_per_b;
    const long long pos = (gi - b * sh.groups_per_b) * VEC;
    const T* __restrict__ s = s_base + (b * sh.hs) * P + pos;
    const T* __restrict__ t = t_base + (b * sh.ht) * P + pos;
    float ssum[VEC], tsum[VEC], gv[VEC];
    head_sum<T, VEC, H>(s, P, sh.hs, ssum);
    head_sum<T, VEC, H>(t, P, sh.ht, tsum);
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        const float sm = ssum[e] * sh.inv_hs;
        const float tm = tsum[e] * sh.inv_ht;
        if constexpr (MSE) {
            const float d = sm - tm;
            acc = fmaf(d, d, acc);
            gv[e] = d * gc;
        } else {
            const float tl = (tm == 0.f) ? 0.f : tm * logf(tm);     // xlogy(t, t)
            acc += tl - tm * logf(sm);                               // 0 * -inf -> NaN like the reference
            gv[e] = -gc * (tm / sm);
        }
    }
    if (g_base) {
        G* __restrict__ g = g_base + (b * sh.hs) * P + pos;
        if constexpr (H > 0) {
#pragma unroll
            for (int h = 0; h < H; ++h) store_vec<G, VEC>(g + h * P, gv);
        } else {
#pragma unroll 4
            for (int h = 0; h < sh.hs; ++h) store_vec<G, VEC>(g + h * P, gv);
        }
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// Cosine rows tile (nn.CosineEmbeddingLoss with target 1, out_cos.py:10-11): one warp per row of [rows, dim].
//   cos = <s,t> / sqrt((<s,s> + eps)(<t,t> + eps)), eps = 1e-12 (ATen EPSILON);  value = sum_rows (1 - cos)
//   d value / d s = -( t / sqrt(..) - cos * s / (<s,s> + eps) )      (times gc)
// Returns the row's (1 - cos) in lane 0 (0 elsewhere).
// ---------------------------------------------------------------------------------------------
template <typename T, typename G>
__device__ __forceinline__ float cos_row_tile(const T* __restrict__ s_base, const T* __restrict__ t_base, G* __restrict__ g_base,
                                              long long rows, int dim, long long row, float gc, int lane) {
    if (row >= rows) return 0.f;
    const T* __restrict__ s = s_base + row * dim;
    const T* __restrict__ t = t_base + row * dim;
    float st = 0.f, ss = 0.f, tt = 0.f;
    for (int d = lane; d < dim; d += 32) {
        const float a = Elem<T>::to_f(s[d]), b = Elem<T>::to_f(t[d]);
        st = fmaf(a, b, st);
        ss = fmaf(a, a, ss);
        tt = fmaf(b, b, tt);
    }
    st = warp_sum(st);
    ss = warp_sum(ss) + 1e-12f;
    tt = warp_sum(tt) + 1e-12f;
    const float inv = rsqrtf(ss) * rsqrtf(tt);
    const float c = st * inv;
    if (g_base) {
        G* __restrict__ g = g_base + row * dim;
        const float k = c / ss;
        for (int d = lane; d < dim; d += 32)
            g[d] = Elem<G>::from_f(-gc * (Elem<T>::to_f(t[d]) * inv - k * Elem<T>::to_f(s[d])));
    }
    return lane == 0 ? 1.f - c : 0.f;
}

}  // namespace dcb
