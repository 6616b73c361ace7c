"""Fused tcgen05 contrastive + logit-KL path (and the per-module logits kernels) against the reference-generated golden
fixtures and the float64 oracle.  All calls go through the C ABI.

Tolerances (BASELINE.json north_star): loss rel <= 1e-4; gradients rel-L2 <= 1e-3, asserted on the fp32 gradient output;
the bf16-stored gradient (what autograd returns for bf16 inputs) carries ~1.6e-3 of pure storage rounding and is
asserted at 4e-3.  Labels are the global row indices (exact: the diagonal statistic is compared to the oracle logits)."""
import numpy as np
import pytest
import torch

from conftest import golden, numbered, rel_l2
from oracle import closed_form as cf

pytestmark = pytest.mark.gpu

LOSS_RTOL, GRAD_RTOL, GRAD_BF16_STORAGE_RTOL = 1e-4, 1e-3, 4e-3
CLIP = ["clip_b24_d32_t2", "clip_b40_d64_t4", "clip_b130_d72_t1"]


def dev(x, dtype=torch.bfloat16, grad=False):
    return torch.tensor(np.asarray(x), device="cuda").to(dtype).requires_grad_(grad)


def synth(b, d, seed, dtype=torch.bfloat16):
    gen = torch.Generator().manual_seed(seed)
    ti = torch.randn(b, d, generator=gen)
    tt = ti * 0.6 + 0.8 * torch.randn(b, d, generator=gen)
    si = ti + 0.5 * torch.randn(b, d, generator=gen)
    st = tt + 0.5 * torch.randn(b, d, generator=gen)
    return [x.to(dtype) for x in (si, st, ti, tt)]


@pytest.mark.parametrize("b,d", [(24, 32), (130, 72), (256, 512), (300, 200)])
def test_similarity_tiles_match_matmul(cuda_device, b, d):
    """TMA + UMMA descriptors + TMEM read-back: the logits the kernel sees (dumped for this test only) must equal the
    normalised matmul of clip_model.py:37-40, and S_ii must sit at label arange(B)."""
    from distillclip_b200.contrastive import CudaEngine
    si, st, ti, tt = [x.cuda() for x in synth(b, d, 1)]
    eng = CudaEngine()
    inv = eng.inv_norms([si, st, ti, tt])
    dump = (torch.zeros(b, b, device="cuda"), torch.zeros(b, b, device="cuda"))
    stats, _ = eng.row_stats(si, st, ti, tt, inv[0], inv[1], inv[2], inv[3], 0, 2.0, dump=dump)
    torch.cuda.synchronize()
    s_ref, _ = cf.clip_logits(si.float().cpu().numpy(), st.float().cpu().numpy())
    t_ref, _ = cf.clip_logits(ti.float().cpu().numpy(), tt.float().cpu().numpy())
    assert np.abs(dump[0].cpu().numpy() - s_ref).max() <= 2e-6
    assert np.abs(dump[1].cpu().numpy() - t_ref).max() <= 2e-6
    assert np.abs(stats[4].cpu().numpy() - np.diag(s_ref)).max() <= 2e-6
    assert rel_l2(stats[0].cpu().numpy(), np.exp(s_ref - 1).sum(1)) <= 1e-5
    assert rel_l2(stats[2].cpu().numpy(), np.exp((t_ref - 1) / 2.0).sum(1)) <= 1e-5
    # one pass, both directions: the column sums equal the row statistics of the transposed (t2i) problem
    st1, _, col = eng.row_stats(si, st, ti, tt, inv[0], inv[1], inv[2], inv[3], 0, 2.0, with_cols=True)
    st2, _ = eng.row_stats(st, si, tt, ti, inv[1], inv[0], inv[3], inv[2], 0, 2.0)
    assert torch.equal(st1, stats)
    for k in (0, 2, 3):
        assert rel_l2(col[k].cpu().numpy(), st2[k].double().cpu().numpy()) <= 2e-6, k
    et, es = np.exp((t_ref - 1) / 2.0), np.exp((s_ref - 1) / 2.0)
    assert rel_l2(col[3].cpu().numpy(), (et * (t_ref - s_ref)).sum(0)) <= 1e-4
    # slot 1 = Q, the second-order part of Zs - Zt: a sum of per-element differences, accurate relative to ITS size
    q_ref = es - et + et * (t_ref - s_ref) / 2.0
    assert rel_l2(stats[1].cpu().numpy(), q_ref.sum(1)) <= 1e-3 and rel_l2(col[1].cpu().numpy(), q_ref.sum(0)) <= 1e-3
    zs = (stats[2] + stats[1] - stats[3] / 2.0).cpu().numpy()
    assert rel_l2(zs, es.sum(1)) <= 1e-5


@pytest.fixture(params=["single_pass", "pair", "chunk"])
def bwd_kernel(request):
    """The three backward routes: CTA-pair kernel storing its gradient tiles + G^T GEMM for the other side (default,
    D <= 768), one CTA-pair pass per side, and the single-CTA D-chunked kernel."""
    from distillclip_b200 import contrastive as ct
    old = ct.CudaEngine.use_pair_kernel, ct.CudaEngine.single_pass_backward
    ct.CudaEngine.use_pair_kernel = request.param != "chunk"
    ct.CudaEngine.single_pass_backward = request.param == "single_pass"
    yield request.param
    ct.CudaEngine.use_pair_kernel, ct.CudaEngine.single_pass_backward = old


@pytest.mark.parametrize("rows,cols,d", [(64, 256, 32), (130, 130, 72), (300, 520, 512), (1030, 700, 768), (257, 64, 8),
                                         (2048, 1024, 1024)])
def test_gt_gemm_matches_matmul(cuda_device, rows, cols, d):
    """clip_gt_gemm_kernel alone: acc[j, :] = sum_i G[i, j] a_hatT[:, i] with G read as an MN-major tcgen05 operand
    (TMA box {64 j, 64 i}), ragged edges zero-filled, K split over clusters.  fp16 products are exact in fp32, so the
    only difference to a float64 matmul is fp32 summation order."""
    from distillclip_b200.contrastive import CudaEngine
    gen = torch.Generator().manual_seed(rows + cols + d)
    eng = CudaEngine()
    g = eng.alloc_g(rows, cols, "cuda")
    g.fill_(float("nan"))                                        # pitch padding must never be read
    g[:, :cols] = torch.randn(rows, cols, generator=gen).to(torch.float16).cuda()
    pitch = (rows + 7) // 8 * 8
    a_hat_t = torch.full((d, pitch), float("nan"), dtype=torch.float16, device="cuda")
    a_hat_t[:, :rows] = (torch.randn(d, rows, generator=gen) / 8).to(torch.float16).cuda()
    acc = eng.col_acc_from_g(g, a_hat_t, rows, cols, d)
    ref = g[:, :cols].double().t() @ a_hat_t[:, :rows].double().t()
    got = acc.double().sum(0)
    assert got.shape == ref.shape
    assert torch.isfinite(got).all()
    assert float((got - ref).abs().max()) <= 2e-5 * float(ref.abs().max()) + 1e-6


@pytest.mark.parametrize("b,d", [(130, 72), (256, 512), (300, 200), (384, 768)])
def test_pair_kernel_sees_the_right_logits(cuda_device, b, d):
    """cta_group::2 plumbing (cluster launch, per-CTA operand halves, the 2x2 TMEM accumulator layout): the logits the
    pair kernel's epilogue reconstructs must equal the normalised matmul."""
    from distillclip_b200 import contrastive as ct
    si, st, ti, tt = [x.cuda() for x in synth(b, d, 5)]
    eng = ct.CudaEngine()
    eng.use_pair_kernel = True
    eng.dump_pair_logits = torch.zeros(b, b, device="cuda")
    out, saved = ct.contrastive_forward(eng, si, st, ti, tt, 2.0, None)
    up = torch.tensor([1.0, 1.0], device="cuda")
    ct.contrastive_backward(eng, saved, up, want_img=True, want_txt=False)
    torch.cuda.synchronize()
    s_ref, _ = cf.clip_logits(si.float().cpu().numpy(), st.float().cpu().numpy())
    assert np.abs(eng.dump_pair_logits.cpu().numpy() - s_ref).max() <= 2e-6


def _fused(si, st, ti, tt, T, w_hard, w_soft, grad_dtype=None):
    from distillclip_b200 import contrastive as ct
    eng = ct.CudaEngine()
    out, saved = ct.contrastive_forward(eng, si, st, ti, tt, T, None)
    up = torch.tensor([w_hard, w_soft], dtype=torch.float32, device="cuda")
    gi, gt = ct.contrastive_backward(eng, saved, up, grad_dtype=grad_dtype)
    return out, gi, gt


@pytest.mark.parametrize("name", CLIP)
def test_fused_contrastive_golden(cuda_device, bwd_kernel, name):
    g = golden(name)
    T = float(g["temperature"])
    si, st, ti, tt = dev(g["stu_img"]), dev(g["stu_txt"]), dev(g["tea_img"]), dev(g["tea_txt"])
    out, gi, gt = _fused(si, st, ti, tt, T, 1.0, 0.0, torch.float32)
    assert float(out[0]) == pytest.approx(float(g["hard_f64"]), rel=LOSS_RTOL)
    assert float(out[1]) == pytest.approx(float(g["soft_f64"]), rel=LOSS_RTOL)
    assert rel_l2(gi.cpu().numpy(), g["dhard_img_f64"]) <= GRAD_RTOL
    assert rel_l2(gt.cpu().numpy(), g["dhard_txt_f64"]) <= GRAD_RTOL
    out, gi, gt = _fused(si, st, ti, tt, T, 0.0, 1.0, torch.float32)
    assert rel_l2(gi.cpu().numpy(), g["dsoft_img_f64"]) <= GRAD_RTOL
    assert rel_l2(gt.cpu().numpy(), g["dsoft_txt_f64"]) <= GRAD_RTOL


@pytest.mark.parametrize("b,d,T,dtype", [(256, 512, 2.0, torch.bfloat16), (512, 512, 1.0, torch.bfloat16),
                                         (384, 768, 4.0, torch.bfloat16), (200, 136, 0.5, torch.bfloat16),
                                         (256, 512, 2.0, torch.float16), (256, 1024, 2.0, torch.bfloat16)])
def test_fused_contrastive_random_vs_oracle(cuda_device, bwd_kernel, b, d, T, dtype):
    si, st, ti, tt = synth(b, d, b + d, dtype)
    ref = cf.contrastive_from_embeddings(*[x.float().numpy() for x in (si, st, ti, tt)], T, w_hard=0.6, w_soft=0.4)
    out, gi, gt = _fused(si.cuda(), st.cuda(), ti.cuda(), tt.cuda(), T, 0.6, 0.4, torch.float32)
    assert float(out[0]) == pytest.approx(ref["hard"], rel=LOSS_RTOL)
    assert float(out[1]) == pytest.approx(ref["soft"], rel=LOSS_RTOL)
    assert rel_l2(gi.cpu().numpy(), ref["d_img"]) <= GRAD_RTOL
    assert rel_l2(gt.cpu().numpy(), ref["d_txt"]) <= GRAD_RTOL


@pytest.mark.parametrize("T", [0.05, 0.3, 16.0])
def test_fused_contrastive_temperature_range(cuda_device, T):
    """Sharp (T = 0.05: exp((S-1)/T) spans 17 decades, just above MIN_FUSED_TEMPERATURE) and flat (T = 16: the row KL is a
    ~1e-6 second-order quantity) softmaxes against the f64 oracle."""
    si, st, ti, tt = synth(256, 128, 77)
    ref = cf.contrastive_from_embeddings(*[x.float().numpy() for x in (si, st, ti, tt)], T, w_hard=0.5, w_soft=0.5)
    out, gi, gt = _fused(si.cuda(), st.cuda(), ti.cuda(), tt.cuda(), T, 0.5, 0.5, torch.float32)
    assert float(out[0]) == pytest.approx(ref["hard"], rel=LOSS_RTOL)
    assert float(out[1]) == pytest.approx(ref["soft"], rel=LOSS_RTOL)
    assert rel_l2(gi.cpu().numpy(), ref["d_img"]) <= GRAD_RTOL
    assert rel_l2(gt.cpu().numpy(), ref["d_txt"]) <= GRAD_RTOL


def test_hard_label_only_autograd(cuda_device, bwd_kernel):
    from distillclip_b200.contrastive import clip_contrastive
    si, st, _, _ = synth(320, 256, 9)
    ref = cf.contrastive_from_embeddings(si.float().numpy(), st.float().numpy(), w_hard=1.0)
    a, b = si.cuda().requires_grad_(True), st.cuda().requires_grad_(True)
    res = clip_contrastive(a, b, want_hard=True, want_soft=False)
    assert set(res) == {"hard_label"}
    res["hard_label"].backward()
    assert float(res["hard_label"]) == pytest.approx(ref["hard"], rel=LOSS_RTOL)
    assert a.grad.dtype == torch.bfloat16
    assert rel_l2(a.grad.float().cpu().numpy(), ref["d_img"]) <= GRAD_BF16_STORAGE_RTOL
    assert rel_l2(b.grad.float().cpu().numpy(), ref["d_txt"]) <= GRAD_BF16_STORAGE_RTOL


def test_row_sharded_virtual_ranks(cuda_device, bwd_kernel):
    """R virtual ranks on one GPU.  Each rank's kernel call sees its row slice (row_offset = r*B/R)
    against all columns; per-rank sums add up to the single-process global-batch oracle and the per-rank gradients
    concatenate to the oracle gradient.  Labels for local row i are r*B/R + i (checked through stats[4])."""
    from distillclip_b200 import contrastive as ct
    R, b, d, T = 4, 512, 256, 2.0
    si, st, ti, tt = [x.cuda() for x in synth(b, d, 21)]
    ref = cf.contrastive_from_embeddings(*[x.float().cpu().numpy() for x in (si, st, ti, tt)], T, w_hard=1.0, w_soft=1.0)
    s_ref, _ = cf.clip_logits(si.float().cpu().numpy(), st.float().cpu().numpy())
    eng = ct.CudaEngine()
    inv = eng.inv_norms([si, st, ti, tt])
    bl = b // R
    sums = torch.zeros(4, dtype=torch.float64, device="cuda")
    stats_i, cols, diags = [], torch.zeros(4, b, device="cuda"), []
    for r in range(R):
        loc = slice(r * bl, (r + 1) * bl)
        s1, l1, col = eng.row_stats(si[loc], st, ti[loc], tt, inv[0][loc], inv[1], inv[2][loc], inv[3], r * bl, T, with_cols=True)
        assert np.abs(s1[4].cpu().numpy() - np.diag(s_ref)[loc]).max() <= 2e-6       # global labels, exact position
        cols += col                                                                   # the all-reduce of the real thing
        stats_i.append((s1, l1))
    stats_t = []
    for r in range(R):
        s2, l2 = eng.col_finish(cols, stats_i[r][0][4].contiguous(), r * bl, T, True)
        sums += eng.losses(stats_i[r][1], l2, b, T, True)[0]
        stats_t.append(s2)
    stats_i = [x[0] for x in stats_i]
    assert float(0.5 * (sums[0] + sums[1]) / b) == pytest.approx(ref["hard"], rel=LOSS_RTOL)
    assert float(0.5 * (sums[2] + sums[3])) == pytest.approx(ref["soft"], rel=LOSS_RTOL)
    up = torch.tensor([1.0, 1.0], device="cuda")
    ci_all, gm_i = eng.coef(torch.cat(stats_i, 1).contiguous(), b, T, True, up)
    ct_all, gm_t = eng.coef(torch.cat(stats_t, 1).contiguous(), b, T, True, up)
    st_t, si_t = eng.transpose_norm(st, inv[1]), eng.transpose_norm(si, inv[0])
    gi, gt = [], []
    for r in range(R):
        loc = slice(r * bl, (r + 1) * bl)
        gi.append(eng.row_grads(si[loc], st, ti[loc], tt, st_t, inv[0][loc], inv[1], inv[2][loc], inv[3],
                                ci_all[:, loc].contiguous(), ct_all, gm_i, gm_t, r * bl, b, T, up, torch.float32))
        gt.append(eng.row_grads(st[loc], si, tt[loc], ti, si_t, inv[1][loc], inv[0], inv[3][loc], inv[2],
                                ct_all[:, loc].contiguous(), ci_all, gm_t, gm_i, r * bl, b, T, up, torch.float32))
    assert rel_l2(torch.cat(gi).cpu().numpy(), ref["d_img"]) <= GRAD_RTOL
    assert rel_l2(torch.cat(gt).cpu().numpy(), ref["d_txt"]) <= GRAD_RTOL


def _tower(g, prefix, cls, grad):
    kw = {}
    for f in ("last_representation", "embedding"):
        kw[f] = dev(g[f"{prefix}.{f}"], grad=grad)
    for f in ("attention_probs", "representations"):
        kw[f] = [dev(x, grad=grad) for x in numbered(g, f"{prefix}.{f}.")]
    return cls(**kw)


def _leaves(t):
    return [t.last_representation, *t.attention_probs, *t.representations, t.embedding]


@pytest.mark.parametrize("name,kwargs", [
    ("calc_lclip_stage", dict(loss_name=["hard_label", "soft_label", "hidden_rep_mse"], temperature=2.0,
                              loss_scale={"soft_label": 0.25},
                              percent={"hard_label": 0.5, "soft_label": 0.25, "hidden_rep_mse": 0.25})),
    ("calc_lclip_logits_only", dict(loss_name=["hard_label", "soft_label"], temperature=3.0))])
@pytest.mark.parametrize("fused", [True, False])
def test_loss_calculator_two_tower_golden(cuda_device, name, kwargs, fused):
    """LossCalculator(...)(stu, tea, 'all') as DualDistillModel calls it (reference model/dual_distill_model.py:79-80,124).
    fused=True: from `last_representation` (logits never built); fused=False: HardLabel/SoftLabel on materialised logits."""
    from distillclip_b200.model import (CLIPOutput, LossCalculator, TextTransformerOutput, VisionTransformerOutput)
    g = golden(name)
    sv, stx = _tower(g, "stu.visual", VisionTransformerOutput, True), _tower(g, "stu.text", TextTransformerOutput, True)
    tv, ttx = _tower(g, "tea.visual", VisionTransformerOutput, False), _tower(g, "tea.text", TextTransformerOutput, False)

    def clip_out(v, x):
        if fused:
            return CLIPOutput(visual_output=v, text_output=x)
        a = v.last_representation.float()
        b = x.last_representation.float()
        lg = (a / a.norm(dim=1, keepdim=True)) @ (b / b.norm(dim=1, keepdim=True)).t()      # the caller's CLIPModel.forward
        return CLIPOutput(visual_output=v, text_output=x, i2t_logits=lg, t2i_logits=lg.T)
    calc = LossCalculator(**kwargs)
    calc.fused_contrastive = fused
    loss, res = calc(clip_out(sv, stx), clip_out(tv, ttx), "all")
    loss.backward()
    assert float(loss) == pytest.approx(float(g["loss_f64"]), rel=LOSS_RTOL)
    for k, v in res.items():
        assert float(v) == pytest.approx(float(g[f"res.{k}_f64"]), rel=LOSS_RTOL), k
    assert set(res) == {k[4:-4] for k in g if k.startswith("res.") and k.endswith("_f64")}
    for i, leaf in enumerate(_leaves(sv) + _leaves(stx)):
        key = f"grad{i}_f64"
        if key in g:
            assert rel_l2(leaf.grad.float().cpu().numpy(), g[key]) <= GRAD_BF16_STORAGE_RTOL, key


@pytest.mark.parametrize("name", CLIP)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_logits_modules_golden(cuda_device, name, dtype):
    """HardLabel()(logits) / SoftLabel(T)(stu, tea) on materialised logits and on the `.T` view."""
    from distillclip_b200.model import HardLabel, SoftLabel
    g = golden(name)
    T = float(g["temperature"])
    s_np, t_np = g["i2t_logits_f64"], g["tea_i2t_logits_f64"]
    s = torch.tensor(s_np, device="cuda").to(dtype).requires_grad_(True)
    t = torch.tensor(t_np, device="cuda").to(dtype)
    s_used, t_used = s.detach().double().cpu().numpy(), t.double().cpu().numpy()     # the values the kernel actually saw
    l_ref, dl_ref = cf.hard_label(s_used)
    loss = HardLabel()(s)
    loss.backward()
    assert float(loss) == pytest.approx(l_ref, rel=LOSS_RTOL)
    assert rel_l2(s.grad.float().cpu().numpy(), dl_ref) <= (GRAD_RTOL if dtype == torch.float32 else GRAD_BF16_STORAGE_RTOL)
    s.grad = None
    l_ref, dl_ref = cf.soft_label(s_used, t_used, T)
    loss = SoftLabel(T)(s, t)
    loss.backward()
    assert float(loss) == pytest.approx(l_ref, rel=LOSS_RTOL, abs=1e-7)
    assert rel_l2(s.grad.float().cpu().numpy(), dl_ref) <= (GRAD_RTOL if dtype == torch.float32 else GRAD_BF16_STORAGE_RTOL)
    s.grad = None
    l_ref, dl_ref = cf.hard_label(s_used.T)
    loss = HardLabel()(s.T)
    loss.backward()
    assert float(loss) == pytest.approx(l_ref, rel=LOSS_RTOL)
    assert rel_l2(s.grad.float().cpu().numpy(), dl_ref.T) <= (GRAD_RTOL if dtype == torch.float32 else GRAD_BF16_STORAGE_RTOL)


def test_full_size_lclip_properties(cuda_device):
    """BASELINE configs[3] size (B=4096, D=512).  Size-independent properties instead of the O(B^2) oracle:
    identical student and teacher -> KL = 0 and zero soft gradient; hard loss of perfectly aligned pairs < log(B);
    gradients are orthogonal to their embedding rows (Jacobian of x/||x||); value symmetric under image<->text swap."""
    from distillclip_b200 import contrastive as ct
    b, d, T = 4096, 512, 2.0
    si, st, ti, tt = [x.cuda() for x in synth(b, d, 33)]
    out, gi, gt = _fused(ti, tt, ti, tt, T, 0.0, 1.0, torch.float32)
    assert abs(float(out[1])) <= 1e-4 * b and float(gi.abs().max()) <= 1e-6
    out, gi, gt = _fused(si, st, ti, tt, T, 1.0, 1.0, torch.float32)
    out2, gi2, gt2 = _fused(st, si, tt, ti, T, 1.0, 1.0, torch.float32)
    # row sums (Kahan over 16-column pieces) against column sums (butterfly over rows, double across row blocks) of the same
    # tiles: both within the 1e-4 loss budget of the oracle, a few 1e-5 apart on the KL (a small difference of logs)
    assert float(out[0]) == pytest.approx(float(out2[0]), rel=1e-6) and float(out[1]) == pytest.approx(float(out2[1]), rel=5e-5)
    ref = cf.contrastive_from_embeddings(*[x.float().cpu().numpy() for x in (si, st, ti, tt)], T, w_hard=1.0, w_soft=1.0)
    assert float(out[0]) == pytest.approx(ref["hard"], rel=LOSS_RTOL) and float(out[1]) == pytest.approx(ref["soft"], rel=LOSS_RTOL)
    assert rel_l2(gi.cpu().numpy(), ref["d_img"]) <= GRAD_RTOL and rel_l2(gt.cpu().numpy(), ref["d_txt"]) <= GRAD_RTOL
    # row sums (per-thread, Kahan) and column sums (warp butterfly + per-row-block reduce) round differently: ~2e-5
    assert rel_l2(gi.cpu().numpy(), gt2.cpu().numpy()) <= 1e-4
    assert 0.0 < float(out[0]) < np.log(b)
    radial = (gi * si.float()).sum(1).abs().max() / (gi.norm(dim=1).max() * si.float().norm(dim=1).max())
    assert float(radial) <= 1e-4
