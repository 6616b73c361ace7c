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
};

__device__ __forceinline__ void clip_load_upstream(const ClipUpstream& u, float& up_hard, float& up_soft) {
    const float gt = u.g_total ? __ldg(u.g_total) : 0.f;
    up_hard = gt * u.w_hard + (u.g_hard ? __ldg(u.g_hard) * u.s_hard : 0.f);
    up_soft = gt * u.w_soft + (u.g_soft ? __ldg(u.g_soft) * u.s_soft : 0.f);
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
