// Pieces shared by the kernels of the fused contrastive pipeline (clip_pipeline.cu, clip_bwd_pair.cu): how the upstream
// gradients reach the device, and the power-of-two scale of the fp16 gradient tiles.
#pragma once
#include "common.cuh"

namespace dcb {

// d(total)/d(hard) and d(total)/d(soft) without any host arithmetic or torch glue kernels:
//   outputs of the autograd node are (hard * s_hard, soft * s_soft, total = p_hard * hard * s_hard + p_soft * soft * s_soft)
//   (reference model/_loss.py:231-234: `cal_res[n] *= scale[n]; loss += cal_res[n] * percent[n]`), so
//   up_hard = g_total * w_hard + g_hard * s_hard,  w_hard = p_hard * s_hard   (same for soft);
// the g_* are 0-dim fp32 device tensors handed over by autograd (nullptr = that output received no gradient).
struct ClipUpstream {
    const float* g_total;
    const float* g_hard;
    const float* g_soft;
    float w_hard, w_soft, s_hard, s_soft;
    // the two other logit losses that ride on the same tiles (all zero / NULL when they are not requested):
    // CLIPCosDiff (clip_cos_diff.py:16-23) and LogitsMSE (logits_mse.py:9-10), each 0.5 (i2t + t2i) = one direction
    const float* g_cos;
    const float* g_mse;
    float w_cos, w_mse, s_cos, s_mse;
};

// g5 = {g_total, g_hard, g_soft, g_cos, g_mse} device scalars (NULL = no gradient for that output);
// w8 = {w_hard, w_soft, s_hard, s_soft, w_cos, w_mse, s_cos, s_mse} host floats (w = percent * scale)
inline ClipUpstream clip_upstream_from(const float* const* g5, const float* w8) {
    ClipUpstream u{};
    u.g_total = g5[0];
    u.g_hard = g5[1];
    u.g_soft = g5[2];
    u.g_cos = g5[3];
    u.g_mse = g5[4];
    u.w_hard = w8[0];
    u.w_soft = w8[1];
    u.s_hard = w8[2];
    u.s_soft = w8[3];
    u.w_cos = w8[4];
    u.w_mse = w8[5];
    u.s_cos = w8[6];
    u.s_mse = w8[7];
    return u;
}

__device__ __forceinline__ void clip_load_upstream(const ClipUpstream& u, float& up_hard, float& up_soft) {
    const float gt = u.g_total ? __ldg(u.g_total) : 0.f;
    up_hard = gt * u.w_hard + (u.g_hard ? __ldg(u.g_hard) * u.s_hard : 0.f);
    up_soft = gt * u.w_soft + (u.g_soft ? __ldg(u.g_soft) * u.s_soft : 0.f);
}
__device__ __forceinline__ void clip_load_upstream_extra(const ClipUpstream& u, float& up_cos, float& up_mse) {
    const float gt = u.g_total ? __ldg(u.g_total) : 0.f;
    up_cos = gt * u.w_cos + (u.g_cos ? __ldg(u.g_cos) * u.s_cos : 0.f);
    up_mse = gt * u.w_mse + (u.g_mse ? __ldg(u.g_mse) * u.s_mse : 0.f);
}
// dL/dS_ij of the two extra losses, off the diagonal:  up_cos [S_ij > T_ij] / (B (B - 1)) + up_mse 2 (S_ij - T_ij) / B^2
// (the diagonal's -up_cos [T_ii > S_ii] / B is a label-like term, added in fp32 by the finish kernel); |S - T| <= 2
__device__ __forceinline__ float clip_extra_bound(float up_cos, float up_mse, float inv_batch, float inv_pairs) {
    return fabsf(up_cos) * inv_pairs + fabsf(up_mse) * 4.f * inv_batch * inv_batch;
}

// bounds[0..2] = max over ALL rows (every rank) of the unit coefficients {1/(2 B A_i), T/(2 Zs_i), T/(2 Zt_i)},
// bounds[3..5] = the same over all columns.  Every exp term of G_ij is <= 1 (cosine logits, shift 1), hence
//   |G_ij| <= |up_hard| (x_i + x'_j) + |up_soft| (y_i + z_i + y'_j + z'_j) <= this bound.
__device__ __forceinline__ float clip_grad_bound(const float* __restrict__ bounds, float up_hard, float up_soft) {
    return fabsf(up_hard) * (__ldg(bounds) + __ldg(bounds + 3)) +
           fabsf(up_soft) * ((__ldg(bounds + 1) + __ldg(bounds + 2)) + (__ldg(bounds + 4) + __ldg(bounds + 5)));
}

// 2^k with 2^k * gmax in [2^13, 2^14): the gradient tiles are stored / multiplied as fp16 (11 significant bits)
__device__ __forceinline__ float clip_tile_scale(float gmax) {
    if (!(gmax > 0.f) || !isfinite(gmax)) return 1.f;
    int e;
    frexpf(gmax, &e);
    return ldexpf(1.f, 14 - e);
}

// KL_i / T^2 from the row sums (Q, Zt, W): -m + log1p(m + Q/Zt), m = -W/(T Zt)  (see clip_fwd.cu)
__device__ __forceinline__ double clip_row_kl_d(double q, double zt, double w, double temperature) {
    const double m = -w / (temperature * zt);
    return -m + log1p(m + q / zt);
}

}  // namespace dcb
