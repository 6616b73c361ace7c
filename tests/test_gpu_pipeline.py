"""The fused contrastive PIPELINE (distillclip_b200/pipeline.py) on the GPU: 7-launch single-GPU flow, device-side
scale / percent weighting and upstream routing, and the multi-rank code paths of every kernel (column chunks per source
rank with negative label offsets, statistics slots stored into every rank's buffer, b_hat^T blocks per rank, the G^T GEMM
scattering into the owners' buffers) driven by R VIRTUAL ranks on one GPU.  Oracle: oracle/closed_form.py (float64)."""
import numpy as np
import pytest
import torch

from conftest import golden, rel_l2
from oracle import closed_form as cf

pytestmark = pytest.mark.gpu
LOSS_RTOL, GRAD_RTOL, GRAD_BF16_STORAGE_RTOL = 1e-4, 1e-3, 4e-3


@pytest.fixture(params=["pair", "split"], autouse=True)
def backward_flow(request):
    """Every test of this module runs with both backward flows: the fused pair kernel (recompute + image-side GEMM) and the
    split flow (recompute -> fp16 G tiles, one GEMM per tower over the stored tiles; the default above 4096 tiles)."""
    from distillclip_b200 import contrastive as ct
    eng = type(ct._ENGINE)
    old = eng.split_backward
    eng.split_backward = "1" if request.param == "split" else "0"
    yield request.param
    eng.split_backward = old


def synth(b, d, seed, dtype=torch.bfloat16):
    gen = torch.Generator().manual_seed(seed)
    ti = torch.randn(b, d, generator=gen)
    tt = ti * 0.6 + 0.8 * torch.randn(b, d, generator=gen)
    si = ti + 0.5 * torch.randn(b, d, generator=gen)
    st = tt + 0.5 * torch.randn(b, d, generator=gen)
    return [x.to(dtype) for x in (si, st, ti, tt)]


def _one(v=1.0):
    return torch.tensor(float(v), dtype=torch.float32, device="cuda")


@pytest.mark.parametrize("b,d,T,dtype", [(256, 512, 2.0, torch.bfloat16), (384, 768, 4.0, torch.bfloat16),
                                         (200, 136, 0.5, torch.bfloat16), (130, 72, 1.0, torch.bfloat16),
                                         (256, 512, 2.0, torch.float16), (1000, 64, 2.0, torch.bfloat16),
                                         (384, 1024, 2.0, torch.bfloat16), (200, 896, 1.0, torch.bfloat16)])
def test_pipeline_single_gpu_vs_oracle(cuda_device, b, d, T, dtype):
    from distillclip_b200 import contrastive as ct, pipeline as pl
    si, st, ti, tt = synth(b, d, b + d, dtype)
    ref = cf.contrastive_from_embeddings(*[x.float().numpy() for x in (si, st, ti, tt)], T, w_hard=0.6, w_soft=0.4)
    out, saved = pl.pipeline_forward(ct._ENGINE, pl.LocalExchange(), si.cuda(), st.cuda(), ti.cuda(), tt.cuda(), T, (0.6, 0.4, 1.0, 1.0))
    gi, gt = pl.pipeline_backward(ct._ENGINE, saved, (_one(), None, None), grad_dtype=torch.float32)
    assert float(out[0]) == pytest.approx(ref["hard"], rel=LOSS_RTOL)
    assert float(out[1]) == pytest.approx(ref["soft"], rel=LOSS_RTOL)
    assert float(out[4]) == pytest.approx(0.6 * ref["hard"] + 0.4 * ref["soft"], rel=LOSS_RTOL)
    assert rel_l2(gi.cpu().numpy(), ref["d_img"]) <= GRAD_RTOL
    assert rel_l2(gt.cpu().numpy(), ref["d_txt"]) <= GRAD_RTOL


@pytest.mark.parametrize("name", ["clip_b24_d32_t2", "clip_b40_d64_t4", "clip_b130_d72_t1"])
def test_pipeline_golden(cuda_device, name):
    from distillclip_b200 import contrastive as ct, pipeline as pl
    g = golden(name)
    T = float(g["temperature"])
    emb = [torch.tensor(g[k], device="cuda").to(torch.bfloat16) for k in ("stu_img", "stu_txt", "tea_img", "tea_txt")]
    for w, keys in (((1.0, 0.0, 1.0, 1.0), ("dhard_img_f64", "dhard_txt_f64")), ((0.0, 1.0, 1.0, 1.0), ("dsoft_img_f64", "dsoft_txt_f64"))):
        out, saved = pl.pipeline_forward(ct._ENGINE, pl.LocalExchange(), *emb, T, w)
        gi, gt = pl.pipeline_backward(ct._ENGINE, saved, (_one(), None, None), grad_dtype=torch.float32)
        assert float(out[0]) == pytest.approx(float(g["hard_f64"]), rel=LOSS_RTOL)
        assert float(out[1]) == pytest.approx(float(g["soft_f64"]), rel=LOSS_RTOL)
        assert rel_l2(gi.cpu().numpy(), g[keys[0]]) <= GRAD_RTOL
        assert rel_l2(gt.cpu().numpy(), g[keys[1]]) <= GRAD_RTOL


def test_public_call_weighting_and_upstream_routes(cuda_device):
    """clip_contrastive(percent=, scale=): 'total' and the scaled terms come from the device (reference _loss.py:231-234);
    gradients for (3 * total - 0.5 * soft_label).backward() equal the oracle with the combined weights; hard-only; no_grad;
    a second backward raises a clear error."""
    from distillclip_b200.contrastive import clip_contrastive
    b, d, T = 320, 256, 2.0
    si, st, ti, tt = synth(b, d, 9)
    p_h, p_s, s_h, s_s = 0.3, 0.7, 2.0, 0.25
    a, c = si.cuda().requires_grad_(True), st.cuda().requires_grad_(True)
    res = clip_contrastive(a, c, ti.cuda(), tt.cuda(), T, True, True, percent=(p_h, p_s), scale=(s_h, s_s))
    ref0 = cf.contrastive_from_embeddings(*[x.float().numpy() for x in (si, st, ti, tt)], T)
    assert float(res["hard_label"]) == pytest.approx(s_h * ref0["hard"], rel=LOSS_RTOL)
    assert float(res["soft_label"]) == pytest.approx(s_s * ref0["soft"], rel=LOSS_RTOL)
    assert float(res["total"]) == pytest.approx(p_h * s_h * ref0["hard"] + p_s * s_s * ref0["soft"], rel=LOSS_RTOL)
    loss = 3.0 * res["total"] - 0.5 * res["soft_label"]
    loss.backward(retain_graph=True)
    ref = cf.contrastive_from_embeddings(*[x.float().numpy() for x in (si, st, ti, tt)], T,
                                         w_hard=3.0 * p_h * s_h, w_soft=3.0 * p_s * s_s - 0.5 * s_s)
    assert rel_l2(a.grad.float().cpu().numpy(), ref["d_img"]) <= GRAD_BF16_STORAGE_RTOL
    assert rel_l2(c.grad.float().cpu().numpy(), ref["d_txt"]) <= GRAD_BF16_STORAGE_RTOL
    with pytest.raises(RuntimeError, match="already ran"):
        loss.backward()
    # hard label only (no teacher), default weights
    a2 = si.cuda().requires_grad_(True)
    r2 = clip_contrastive(a2, st.cuda(), want_hard=True, want_soft=False)
    assert set(r2) == {"hard_label"}
    r2["hard_label"].backward()
    refh = cf.contrastive_from_embeddings(si.float().numpy(), st.float().numpy(), w_hard=1.0)
    assert float(r2["hard_label"]) == pytest.approx(refh["hard"], rel=LOSS_RTOL)
    assert rel_l2(a2.grad.float().cpu().numpy(), refh["d_img"]) <= GRAD_BF16_STORAGE_RTOL
    # no_grad: buffers are released at once (many forwards in a row must not trip the in-flight limit)
    with torch.no_grad():
        vals = [float(clip_contrastive(si.cuda(), st.cuda(), ti.cuda(), tt.cuda(), T, True, True)["soft_label"]) for _ in range(5)]
    assert all(v == vals[0] for v in vals) and vals[0] == pytest.approx(ref0["soft"], rel=LOSS_RTOL)


def _extras_oracle(si, st, ti, tt, T, w):
    """values and gradients of p_h hard + p_s soft + p_c s_c cos_diff + p_m s_m logits_mse from the float64 closed forms."""
    f = lambda x: x.float().cpu().numpy()
    s, _ = cf.clip_logits(f(si), f(st))
    t, _ = cf.clip_logits(f(ti), f(tt))
    cd, g_cd = cf.cos_diff(s, t)
    cd2, g_cd2 = cf.cos_diff(s.T, t.T)
    lm, g_lm = cf.logits_mse(s, t)
    p_h, p_s, s_h, s_s, p_c, p_m, s_c, s_m = w
    ref = cf.contrastive_from_embeddings(f(si), f(st), f(ti), f(tt), T, w_hard=p_h * s_h, w_soft=p_s * s_s)
    dl = ref["d_logits"] + p_c * s_c * 0.5 * (g_cd + g_cd2.T) + p_m * s_m * 0.5 * (g_lm + g_lm.T.T)
    d_img, d_txt = cf.clip_logits_backward(f(si), f(st), dl)
    return dict(hard=ref["hard"], soft=ref["soft"], cos=0.5 * (cd + cd2), mse=lm, d_img=d_img, d_txt=d_txt)


@pytest.mark.parametrize("b,d,T", [(256, 512, 2.0), (384, 768, 4.0), (130, 72, 1.0), (520, 64, 2.0)])
def test_pipeline_cos_diff_logits_mse_vs_oracle(cuda_device, b, d, T):
    """CLIPCosDiff (reference clip_cos_diff.py:5-23: exact off-diagonal set, diagonal term) and LogitsMSE (logits_mse.py:9-10)
    from the embeddings inside the fused kernels, alone and mixed with hard / soft label."""
    from distillclip_b200 import contrastive as ct, pipeline as pl
    si, st, ti, tt = [x.cuda() for x in synth(b, d, b + d)]
    for w in ((0.0, 0.0, 1.0, 1.0, 1.0, 0.0, 1.0, 1.0), (0.0, 0.0, 1.0, 1.0, 0.0, 1.0, 1.0, 1.0), (0.3, 0.2, 1.0, 2.0, 0.4, 0.1, 0.5, 3.0)):
        ref = _extras_oracle(si, st, ti, tt, T, w)
        out, saved = pl.pipeline_forward(ct._ENGINE, pl.LocalExchange(), si, st, ti, tt, T, w, extra=True)
        gi, gt = pl.pipeline_backward(ct._ENGINE, saved, (_one(), None, None, None, None), grad_dtype=torch.float32)
        assert float(out[5]) == pytest.approx(ref["cos"], rel=LOSS_RTOL)
        assert float(out[6]) == pytest.approx(ref["mse"], rel=LOSS_RTOL)
        want = w[0] * w[2] * ref["hard"] + w[1] * w[3] * ref["soft"] + w[4] * w[6] * ref["cos"] + w[5] * w[7] * ref["mse"]
        assert float(out[4]) == pytest.approx(want, rel=LOSS_RTOL)
        # the indicator [S_ij > T_ij] can flip where |S - T| is at rounding level: budget 2e-3 for the pure cos_diff gradient
        tol = 2e-3 if w[4] else GRAD_RTOL
        assert rel_l2(gi.cpu().numpy(), ref["d_img"]) <= tol
        assert rel_l2(gt.cpu().numpy(), ref["d_txt"]) <= tol


def test_public_call_cos_diff_only(cuda_device):
    """The shipped stage-3 loss list has cos_diff and no hard / soft label: clip_contrastive(want_cos_diff=True) alone."""
    from distillclip_b200.contrastive import clip_contrastive
    si, st, ti, tt = [x.cuda() for x in synth(256, 128, 3)]
    a, c = si.clone().requires_grad_(True), st.clone().requires_grad_(True)
    res = clip_contrastive(a, c, ti, tt, None, want_hard=False, want_soft=False, want_cos_diff=True, want_logits_mse=True)
    assert set(res) == {"cos_diff", "logits_mse"}
    (res["cos_diff"] + 2.0 * res["logits_mse"]).backward()
    ref = _extras_oracle(si, st, ti, tt, 1.0, (0.0, 0.0, 1.0, 1.0, 1.0, 2.0, 1.0, 1.0))
    assert float(res["cos_diff"]) == pytest.approx(ref["cos"], rel=LOSS_RTOL)
    assert float(res["logits_mse"]) == pytest.approx(ref["mse"], rel=LOSS_RTOL)
    assert rel_l2(a.grad.float().cpu().numpy(), ref["d_img"]) <= GRAD_BF16_STORAGE_RTOL
    assert rel_l2(c.grad.float().cpu().numpy(), ref["d_txt"]) <= GRAD_BF16_STORAGE_RTOL


class VirtualExchange:
    """R ranks inside one process (TEST INFRASTRUCTURE): the text-side buffers are shared, every virtual rank owns its
    statistics slots and text-gradient partial buffers, the 'peers' are reached through ordinary tensors.  The caller runs
    each pipeline stage for all ranks before the next stage (that is what the barriers of the real exchange enforce)."""

    chunked = True             # one tile launch per source rank, as the symmetric-memory exchange does

    def __init__(self, world, rank, shared, chunked=True):
        self.world, self.rank, self.shared, self.chunked = world, rank, shared, chunked

    def acquire(self, b_local, dim, dtype, has_teacher, k_split, device, aux):
        from distillclip_b200 import pipeline as pl
        sh = self.shared
        if "st_all" not in sh:
            b, pitch = self.world * b_local, (b_local + 7) // 8 * 8
            sh["st_all"], sh["tt_all"] = torch.empty(b, dim, dtype=dtype, device=device), torch.empty(b, dim, dtype=dtype, device=device)
            sh["st_inv_all"], sh["tt_inv_all"] = torch.empty(b, device=device), torch.empty(b, device=device)
            sh["bt_all"] = torch.full((self.world, dim, pitch), float("nan"), dtype=torch.float16, device=device)
            sh["slots"] = [torch.full((self.world, pl.slot_floats(b, b_local)), float("nan"), device=device) for _ in range(self.world)]
            sh["gt_parts"] = [torch.full((self.world * k_split, b_local, dim), float("nan"), device=device) for _ in range(self.world)]
        s = pl.ExchangeSet()
        for k in ("st_all", "tt_all", "st_inv_all", "tt_inv_all", "bt_all"):
            setattr(s, k, sh[k])
        if not has_teacher:
            s.tt_all = s.tt_inv_all = None
        s.slots, s.gt_parts = sh["slots"][self.rank], sh["gt_parts"][self.rank]
        return s

    def start_gather(self, s):
        pass

    def wait_chunk(self, s, src):
        pass

    def wait_all(self, s):
        pass

    def tile_streams(self):
        return None

    def slot_targets(self, s):
        return [self.shared["slots"][d][self.rank] for d in range(self.world)]

    def exchange_slots(self, s):
        return s.slots

    def gt_targets(self, s):
        return self.shared["gt_parts"]

    def after_scatter(self, s):
        pass

    def release(self, s):
        pass


@pytest.mark.parametrize("chunked", [True, False])
@pytest.mark.parametrize("R,b,d,T,teacher", [(4, 512, 256, 2.0, True), (2, 768, 768, 1.0, True), (8, 1024, 64, 2.0, True), (2, 512, 1024, 2.0, True),
                                             (3, 384, 512, 2.0, False)])
def test_pipeline_virtual_ranks(cuda_device, R, b, d, T, teacher, chunked):
    """Every kernel's multi-rank path on one GPU; the result must equal the single-process global-batch oracle
    and the one-rank pipeline on the same data."""
    from distillclip_b200 import contrastive as ct, pipeline as pl
    eng = ct._ENGINE
    si, st, ti, tt = [x.cuda() for x in synth(b, d, 21 + R)]
    if not teacher:
        ti = tt = None
    w = (0.75, 0.5, 1.0, 1.0, 0.3, 0.2, 1.0, 1.0) if teacher else (0.75, 0.0, 1.0, 1.0)
    if teacher:
        ref = _extras_oracle(si, st, ti, tt, T, w)
    else:
        ref = cf.contrastive_from_embeddings(si.float().cpu().numpy(), st.float().cpu().numpy(), w_hard=w[0])
    n = b // R
    shared = {}
    xcs = [VirtualExchange(R, r, shared, chunked) for r in range(R)]
    loc = [slice(r * n, (r + 1) * n) for r in range(R)]
    vs = [pl.forward_prep(eng, xcs[r], si[loc[r]].contiguous(), st[loc[r]].contiguous(), ti[loc[r]].contiguous() if teacher else None,
                          tt[loc[r]].contiguous() if teacher else None, T if teacher else None, w, teacher) for r in range(R)]
    for r in range(R):
        pl.forward_tiles(eng, vs[r])
    outs = [pl.forward_finish(eng, vs[r]) for r in range(R)]
    torch.cuda.synchronize()
    for o in outs:
        assert torch.equal(o, outs[0])                         # bit-identical global values on every rank
    assert float(outs[0][0]) == pytest.approx(ref["hard"], rel=LOSS_RTOL)
    if teacher:
        assert float(outs[0][1]) == pytest.approx(ref["soft"], rel=LOSS_RTOL)
        assert float(outs[0][5]) == pytest.approx(ref["cos"], rel=LOSS_RTOL)
        assert float(outs[0][6]) == pytest.approx(ref["mse"], rel=LOSS_RTOL)
    ups = (_one(), None, None)
    for r in range(R):
        pl.backward_gemms(eng, vs[r], ups)
    grads = [pl.backward_finish(eng, vs[r], ups, grad_dtype=torch.float32) for r in range(R)]
    gi, gt = torch.cat([g[0] for g in grads]), torch.cat([g[1] for g in grads])
    assert torch.isfinite(gi).all() and torch.isfinite(gt).all()
    assert rel_l2(gi.cpu().numpy(), ref["d_img"]) <= GRAD_RTOL
    assert rel_l2(gt.cpu().numpy(), ref["d_txt"]) <= GRAD_RTOL
    # labels: S_ii of local row i sits at global column r n + i (exact position, also in chunks with a negative offset)
    s_ref, _ = cf.clip_logits(si.float().cpu().numpy(), st.float().cpu().numpy())
    diag = torch.cat([v["stats_i2t"][4] for v in vs]).cpu().numpy()
    assert np.abs(diag - np.diag(s_ref)).max() <= 2e-6
