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

// group index -> sample: a 64-bit division costs ~120 instructions per thread (it was a third of the text-stage tower
// kernel's issued instructions, ncu r02); group counts fit 32 bits for every real shape
__device__ __forceinline__ long long div_groups(long long n, long long d) {
    if ((static_cast<unsigned long long>(n | d) >> 32) == 0) return static_cast<long long>(static_cast<unsigned>(n) / static_cast<unsigned>(d));
    return n / d;
}
// One term of KLDiv(sum, log_target=False): xlogy(t, t) - t log(s) = t log(t / s) with the quotient the gradient needs
// anyway (one logf instead of two); t == 0 keeps the reference's corner cases: 0 for s > 0, NaN for s == 0
__device__ __forceinline__ float kl_term(float tm, float sm, float ratio) {
    return tm == 0.f ? 0.f * logf(sm) : tm * logf(ratio);
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
    const long long b = div_groups(gi, sh.groups_per_b);
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
            const float ratio = tm / sm;
            acc += kl_term(tm, sm, ratio);
            gv[e] = -gc * ratio;
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

// GPT groups per thread (group k of the thread = gi0 + k * stride) with EVERY load issued before the first use: a thread
// keeps 2 * H * GPT independent loads in flight.  With VEC = 1 (odd map sizes such as N = 77: 2-byte loads) one group per
// thread leaves only ~32 bytes per thread outstanding, which is what bounded the attention tiles (0.60 of HBM, r01).
template <typename T, typename G, int VEC, int H, bool MSE, int GPT>
__device__ __forceinline__ float attn_tile_multi(const T* __restrict__ s_base, const T* __restrict__ t_base, G* __restrict__ g_base,
                                                 const AttnShape& sh, long long gi0, long long stride, float gc) {
    static_assert(H > 0, "the multi-group tile needs compile-time head counts");
    const long long P = sh.positions;
    float sv[GPT][H][VEC], tv[GPT][H][VEC];
    long long off_s[GPT], off_t[GPT];
    bool ok[GPT];
#pragma unroll
    for (int k = 0; k < GPT; ++k) {
        const long long gi = gi0 + k * stride;
        ok[k] = gi < sh.groups;
        const long long gq = ok[k] ? gi : 0;
        const long long b = div_groups(gq, sh.groups_per_b);
        const long long pos = (gq - b * sh.groups_per_b) * VEC;
        off_s[k] = (b * sh.hs) * P + pos;
        off_t[k] = (b * sh.ht) * P + pos;
    }
#pragma unroll
    for (int k = 0; k < GPT; ++k)
#pragma unroll
        for (int h = 0; h < H; ++h) load_vec<T, VEC>(s_base + off_s[k] + h * P, sv[k][h]);
#pragma unroll
    for (int k = 0; k < GPT; ++k)
#pragma unroll
        for (int h = 0; h < H; ++h) load_vec<T, VEC>(t_base + off_t[k] + h * P, tv[k][h]);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < GPT; ++k) {
        float gv[VEC];
        float part = 0.f;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            float ss = 0.f, ts = 0.f;
#pragma unroll
            for (int h = 0; h < H; ++h) {
                ss += sv[k][h][e];
                ts += tv[k][h][e];
            }
            const float sm = ss * sh.inv_hs, tm = ts * sh.inv_ht;
            if constexpr (MSE) {
                const float d = sm - tm;
                part = fmaf(d, d, part);
                gv[e] = d * gc;
            } else {
                const float ratio = tm / sm;
                part += kl_term(tm, sm, ratio);
                gv[e] = -gc * ratio;
            }
        }
        if (ok[k]) {
            acc += part;
            if (g_base) {
                G* __restrict__ g = g_base + off_s[k];
#pragma unroll
                for (int h = 0; h < H; ++h) store_vec<G, VEC>(g + h * P, gv);
            }
        }
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// Attention tile on ALIGNED 16-byte accesses for 16-bit maps whose head rows are not 16-byte aligned (N = 77: 5929
// positions per head, N = 50: 2500): a thread owns 8 consecutive positions [p0, p0 + 8), p0 % 8 == 0, of one sample.  For
// head h the 8 elements start at element e = (b H + h) P + p0, i.e. s = e & 7 elements into an aligned vector: the thread
// loads the two aligned vectors around them (the second one is the first one of the next thread: an L1 hit) and realigns
// in registers (select network + funnel shift).  The gradient is the same for every head; each thread receives the packed
// gradient of the PREVIOUS thread through shared memory and stores, per head, the aligned vector made of the previous
// thread's last s values and its own first 8 - s.  Row / sample / CTA edges fall back to 2-byte stores (a few threads in
// a thousand).  2-byte accesses moved 64 useful bytes per warp instruction (0.60 of HBM on the text stage, r01).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg128(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// 8 consecutive 16-bit elements starting at element s (0..7) of the 16-element sequence lo || hi
__device__ __forceinline__ uint4 realign16(const uint4 lo, const uint4 hi, const int s) {
    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint32_t t[6], u[5];
#pragma unroll
    for (int j = 0; j < 6; ++j) t[j] = (s & 4) ? w[j + 2] : w[j];
#pragma unroll
    for (int j = 0; j < 5; ++j) u[j] = (s & 2) ? t[j + 1] : t[j];
    const uint32_t sh = (s & 1) * 16;
    return make_uint4(__funnelshift_r(u[0], u[1], sh), __funnelshift_r(u[1], u[2], sh), __funnelshift_r(u[2], u[3], sh),
                      __funnelshift_r(u[3], u[4], sh));
}

struct AttnShape8 {
    long long groups;        // batch * groups_per_b
    long long groups_per_b;  // ceil(positions / 8)
    long long positions;
    long long total_s, total_t;   // elements of the student / teacher tensor (bounds of the second vector load)
    int hs, ht;
    float inv_hs, inv_ht;
};

// one chunk of up to 4 head rows: all loads first, then realign + accumulate
template <typename T>
__device__ __forceinline__ void head_chunk_aligned(const T* __restrict__ base, long long row0, long long P, int h0, int n, long long total,
                                                   float (&acc)[8]) {
    constexpr int kChunk = 4;
    uint4 v0[kChunk], v1[kChunk];
    int sft[kChunk];
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
        if (h0 + i < n) {
            const long long e = row0 + (long long)(h0 + i) * P;
            sft[i] = (int)(e & 7);
            const long long a0 = e - sft[i];
            v0[i] = ldg128(base + a0);
            if (a0 + 16 <= total) {
                v1[i] = ldg128(base + a0 + 8);
            } else {                                   // last vector of the tensor: never read past its end
                const unsigned short* raw = reinterpret_cast<const unsigned short*>(base);
                uint32_t w[4] = {0u, 0u, 0u, 0u};
                for (int k = 0; k < 8; ++k)
                    if (a0 + 8 + k < total) w[k >> 1] |= (uint32_t)raw[a0 + 8 + k] << (16 * (k & 1));
                v1[i] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
        if (h0 + i < n) {
            const uint4 r = realign16(v0[i], v1[i], sft[i]);
            float a, b;
            unpack2<T>(r.x, a, b); acc[0] += a; acc[1] += b;
            unpack2<T>(r.y, a, b); acc[2] += a; acc[3] += b;
            unpack2<T>(r.z, a, b); acc[4] += a; acc[5] += b;
            unpack2<T>(r.w, a, b); acc[6] += a; acc[7] += b;
        }
    }
}

// acc[0..7] += sum over the `heads` rows of the 8 elements starting at element row0 + h * P
template <typename T, int H>
__device__ __forceinline__ void head_sum_aligned(const T* __restrict__ base, long long row0, long long P, int heads, long long total,
                                                 float (&acc)[8]) {
    if constexpr (H > 0) {
#pragma unroll
        for (int h0 = 0; h0 < H; h0 += 4) head_chunk_aligned<T>(base, row0, P, h0, H, total, acc);
    } else {
#pragma unroll 1
        for (int h0 = 0; h0 < heads; h0 += 4) head_chunk_aligned<T>(base, row0, P, h0, heads, total, acc);
    }
}

// Must be called by ALL threads of the CTA (two __syncthreads inside); xchg: kStreamThreads uint4 of shared memory.
template <typename T, typename G, int H, bool MSE>
__device__ __forceinline__ float attn_tile_aligned(const T* __restrict__ s_base, const T* __restrict__ t_base, G* __restrict__ g_base,
                                                   const AttnShape8& sh, long long gi, float gc, uint4* xchg, int tid) {
    static_assert(sizeof(T) == 2 && sizeof(G) == 2, "aligned attention tile: 16-bit maps and gradients");
    const bool active = gi < sh.groups;
    const long long P = sh.positions;
    const long long b = active ? div_groups(gi, sh.groups_per_b) : 0;
    const long long p0 = active ? (gi - b * sh.groups_per_b) * 8 : 0;
    const int valid = (int)(P - p0 < 8 ? P - p0 : 8);
    float acc = 0.f;
    uint4 mine = make_uint4(0u, 0u, 0u, 0u);
    if (active) {
        float ss[8], ts[8], gv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ss[j] = ts[j] = 0.f;
        head_sum_aligned<T, H>(s_base, (b * sh.hs) * P + p0, P, sh.hs, sh.total_s, ss);
        head_sum_aligned<T, H>(t_base, (b * sh.ht) * P + p0, P, sh.ht, sh.total_t, ts);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float sm = ss[j] * sh.inv_hs, tm = ts[j] * sh.inv_ht;
            float part, g;
            if constexpr (MSE) {
                const float d = sm - tm;
                part = d * d;
                g = d * gc;
            } else {
                const float ratio = tm / sm;
                part = kl_term(tm, sm, ratio);
                g = -gc * ratio;
            }
            acc += j < valid ? part : 0.f;
            gv[j] = j < valid ? g : 0.f;
        }
        mine = make_uint4(pack2<G>(gv[0], gv[1]), pack2<G>(gv[2], gv[3]), pack2<G>(gv[4], gv[5]), pack2<G>(gv[6], gv[7]));
    }
    if (g_base == nullptr) return acc;                     // block-uniform
    xchg[tid] = mine;
    __syncthreads();
    const uint4 prev = tid > 0 ? xchg[tid - 1] : make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (!active) return acc;
    const bool has_prev = tid > 0 && p0 > 0;                                   // previous thread = previous 8 positions, same sample
    const bool next_full = tid < kStreamThreads - 1 && p0 + 16 <= P;           // the next thread will store my last s values
    const uint32_t mw[4] = {mine.x, mine.y, mine.z, mine.w};
    unsigned short* __restrict__ g16 = reinterpret_cast<unsigned short*>(g_base);
    const int n_heads = H > 0 ? H : sh.hs;
#pragma unroll 4
    for (int h = 0; h < n_heads; ++h) {
        const long long e = (b * sh.hs + h) * P + p0;
        const int s = (int)(e & 7);
        if (valid == 8 && (s == 0 || has_prev)) {
            const uint4 o = s == 0 ? mine : realign16(prev, mine, 8 - s);
            *reinterpret_cast<uint4*>(g16 + (e - s)) = o;
        } else {
            const int lim = valid < 8 - s ? valid : 8 - s;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < lim) g16[e + j] = (unsigned short)(mw[j >> 1] >> (16 * (j & 1)));
        }
        if (s > 0 && !next_full) {
#pragma unroll
            for (int j = 1; j < 8; ++j)
                if (j >= 8 - s && j < valid) g16[e + j] = (unsigned short)(mw[j >> 1] >> (16 * (j & 1)));
        }
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// STAGED attention tile (16- and 32-bit maps whose head rows are off the 16-byte grid): the bytes in flight live in shared
// memory, not in registers.  One tile = W consecutive positions of one sample, all student and teacher heads.  Every head
// row's byte range is rounded out to 16-byte chunks and fetched with cp.async (16 B per copy, any row alignment), one commit
// group per tile, the NEXT tile of the CTA in flight while the current one is computed (2 buffers).  Threads then read their
// positions from shared memory with element-sized loads (a row's start offset inside its first chunk is row_off).
// Why: with per-thread 2-byte loads a thread holds 2 useful bytes per register in flight; 4 CTAs x 256 threads x 16 loads keep
// ~48 KB per SM outstanding, below what the HBM latency needs (ncu r02: DRAM traffic = algorithmic, 60 % issue slots, 5.7
// warps per issue stalled on loads).  Two staged tiles per CTA keep 2 x 17 KB x 4 CTAs = 136 KB per SM in flight.
// ---------------------------------------------------------------------------------------------
struct AttnStaged {
    long long positions;     // P
    long long tiles_per_b;   // ceil(P / W)
    long long total_s, total_t;   // elements of the student / teacher tensor
    int W, row_bytes;        // positions per tile; staged bytes per head row = W * sizeof(T) + 32
    int hs, ht;
    float inv_hs, inv_ht;
};

__device__ __forceinline__ unsigned smem_addr_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

// all threads of the CTA; ends with a commit_group (possibly empty)
template <typename T>
__device__ __forceinline__ void attn_stage_issue(const T* __restrict__ s_base, const T* __restrict__ t_base, const AttnStaged& a,
                                                 long long tile, unsigned char* buf, short* row_off, int tid) {
    constexpr long long ES = sizeof(T);
    const long long b = div_groups(tile, a.tiles_per_b);
    const long long p0 = (tile - b * a.tiles_per_b) * a.W;
    const int rows = a.hs + a.ht;
    const int cpr = a.row_bytes >> 4;
    for (int c = tid; c < rows * cpr; c += kStreamThreads) {
        const int row = c / cpr, j = c - row * cpr;
        const bool stu = row < a.hs;
        const long long e0 = (b * (stu ? a.hs : a.ht) + (stu ? row : row - a.hs)) * a.positions + p0;
        const long long first = (e0 * ES) & ~15ll;
        const long long src = first + 16ll * j;
        const long long end = (stu ? a.total_s : a.total_t) * ES;
        if (j == 0) row_off[row] = (short)(e0 * ES - first);
        const long long left = end - src;
        if (left > 0) {
            const char* base = reinterpret_cast<const char*>(stu ? s_base : t_base);
            const unsigned dst = smem_addr_u32(buf + (size_t)row * a.row_bytes + 16 * j);
            const unsigned n = left >= 16 ? 16u : (unsigned)left;          // the tensor's last chunk: zero-fill past its end
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(base + src), "r"(n) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <typename T, typename G, int H, bool MSE>
__device__ __forceinline__ float attn_stage_compute(const unsigned char* buf, const short* row_off, G* __restrict__ g_base,
                                                    const AttnStaged& a, long long tile, float gc, int tid) {
    const long long b = div_groups(tile, a.tiles_per_b);
    const long long p0 = (tile - b * a.tiles_per_b) * a.W;
    const int hs = H > 0 ? H : a.hs, ht = H > 0 ? H : a.ht;
    float acc = 0.f;
    for (int q = tid; q < a.W; q += kStreamThreads) {
        if (p0 + q >= a.positions) break;
        float ss = 0.f, ts = 0.f;
        const unsigned char* col = buf + (size_t)q * sizeof(T);
#pragma unroll 4
        for (int h = 0; h < hs; ++h) ss += Elem<T>::to_f(*reinterpret_cast<const T*>(col + (size_t)h * a.row_bytes + row_off[h]));
#pragma unroll 4
        for (int h = 0; h < ht; ++h) ts += Elem<T>::to_f(*reinterpret_cast<const T*>(col + (size_t)(hs + h) * a.row_bytes + row_off[hs + h]));
        const float sm = ss * a.inv_hs, tm = ts * a.inv_ht;
        float g;
        if constexpr (MSE) {
            const float d = sm - tm;
            acc = fmaf(d, d, acc);
            g = d * gc;
        } else {
            const float ratio = tm / sm;
            acc += kl_term(tm, sm, ratio);
            g = -gc * ratio;
        }
        if (g_base) {
            G* __restrict__ gp = g_base + (b * hs) * a.positions + p0 + q;
            const G gval = Elem<G>::from_f(g);
#pragma unroll 4
            for (int h = 0; h < hs; ++h) gp[(long long)h * a.positions] = gval;
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
