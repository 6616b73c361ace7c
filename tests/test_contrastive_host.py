"""Host-side logic of the fused contrastive path on CPU: the kernel decomposition (via the float64 engine double)
against the oracle / golden fixtures, and the row-sharded path under gloo with world_size 2."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import golden, rel_l2
from engine_double import DoubleEngine
from oracle import closed_form as cf
from distillclip_b200 import contrastive as ct

CLIP = ["clip_b24_d32_t2", "clip_b40_d64_t4", "clip_b130_d72_t1"]


def _run(g, w_hard, w_soft, group=None, rows=slice(None), single_pass=True):
    T = float(g["temperature"])
    si, st = torch.tensor(g["stu_img"])[rows], torch.tensor(g["stu_txt"])[rows]
    ti, tt = torch.tensor(g["tea_img"])[rows], torch.tensor(g["tea_txt"])[rows]
    eng = DoubleEngine()
    eng.single_pass_backward = single_pass      # stored gradient tiles + G^T GEMM (+ reduce-scatter) vs one recompute per side
    out, saved = ct.contrastive_forward(eng, si, st, ti, tt, T, group)
    up = torch.tensor([w_hard, w_soft], dtype=torch.float32)
    gi, gt = ct.contrastive_backward(eng, saved, up)
    return out, gi, gt


@pytest.mark.parametrize("single_pass", [True, False])
@pytest.mark.parametrize("name", CLIP)
def test_decomposition_matches_reference_golden(name, single_pass):
    g = golden(name)
    out, gi, gt = _run(g, 1.0, 0.0, single_pass=single_pass)
    assert float(out[0]) == pytest.approx(float(g["hard_f64"]), rel=1e-10)
    assert float(out[1]) == pytest.approx(float(g["soft_f64"]), rel=1e-9)
    assert rel_l2(gi.numpy(), g["dhard_img_f64"]) <= 1e-9
    assert rel_l2(gt.numpy(), g["dhard_txt_f64"]) <= 1e-9
    out, gi, gt = _run(g, 0.0, 1.0, single_pass=single_pass)
    assert rel_l2(gi.numpy(), g["dsoft_img_f64"]) <= 1e-8
    assert rel_l2(gt.numpy(), g["dsoft_txt_f64"]) <= 1e-8


def test_hard_only_without_teacher():
    g = golden(CLIP[1])
    eng = DoubleEngine()
    si, st = torch.tensor(g["stu_img"]), torch.tensor(g["stu_txt"])
    out, saved = ct.contrastive_forward(eng, si, st, None, None, None, None)
    gi, gt = ct.contrastive_backward(eng, saved, torch.tensor([1.0, 0.0]))
    assert float(out[0]) == pytest.approx(float(g["hard_f64"]), rel=1e-10)
    assert rel_l2(gi.numpy(), g["dhard_img_f64"]) <= 1e-9


def _worker(rank, world, port, name, single_pass, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = golden(name)
    b = g["stu_img"].shape[0] // world
    out, gi, gt = _run(g, 0.75, 0.5, group=dist.group.WORLD, rows=slice(rank * b, (rank + 1) * b), single_pass=single_pass)
    q.put((rank, out.numpy(), gi.numpy(), gt.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,single_pass", [("clip_b24_d32_t2", True), ("clip_b40_d64_t4", True), ("clip_b40_d64_t4", False)])
def test_row_sharded_world2_gloo(name, single_pass):
    """Each of 2 ranks holds half the rows; the result must equal the single-process global-batch oracle
    (the oracle of the sharded path is the reference on the concatenated batch)."""
    g = golden(name)
    T = float(g["temperature"])
    ref = cf.contrastive_from_embeddings(g["stu_img"], g["stu_txt"], g["tea_img"], g["tea_txt"], T, w_hard=0.75, w_soft=0.5)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, single_pass, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out, gi, gt in res:
        assert float(out[0]) == pytest.approx(ref["hard"], rel=1e-6)       # fp32 output; same global value on every rank
        assert float(out[1]) == pytest.approx(ref["soft"], rel=1e-6)
    gi = np.concatenate([r[2] for r in res])
    gt = np.concatenate([r[3] for r in res])
    assert rel_l2(gi, ref["d_img"]) <= 1e-8
    assert rel_l2(gt, ref["d_txt"]) <= 1e-8


def _metrics_worker(rank, world, port, name, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from distillclip_b200.metrics import retrieval_metrics
    g = golden(name)
    b = g["img"].shape[0] // world
    rows = slice(rank * b, (rank + 1) * b)
    res = retrieval_metrics(torch.tensor(g["img"])[rows].double(), torch.tensor(g["txt"])[rows].double(), group=dist.group.WORLD,
                            engine=DoubleEngine())
    q.put((rank, {k: float(v) for k, v in res.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_retrieval_metrics_sharded_world2_gloo():
    """Validation metrics with the rows split over 2 ranks (text rows gathered, per-rank sums all-reduced) = the
    reference-generated values for the whole batch."""
    name = "retrieval_b200_d64"
    g = golden(name)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 11) % 2000
    procs = [ctx.Process(target=_metrics_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, got in res:
        for k, v in got.items():
            assert v == pytest.approx(float(g[f"{k}_f64"]), rel=1e-6, abs=1e-9), k


# ----------------------------------------------------------------------------------------------
# the pipeline flow (distillclip_b200/pipeline.py): stage boundaries, slot layout, exchanges, buffer-set bookkeeping
# ----------------------------------------------------------------------------------------------
@pytest.fixture(params=["pair", "split"])
def flow(request):
    """Both backward flows of pipeline.backward_gemms: the fused pair kernel, and the split flow (gradient tiles first, then the
    text-side GEMM, then the image-side GEMM) -- with the float64 engine double."""
    DoubleEngine.split = request.param == "split"
    yield request.param
    DoubleEngine.split = False


def _run_pipeline(g, weights, ups, group=None, rows=slice(None), hard_only=False, extra=False):
    from distillclip_b200 import pipeline as pl
    T = float(g["temperature"])
    si, st = torch.tensor(g["stu_img"])[rows].double(), torch.tensor(g["stu_txt"])[rows].double()
    ti, tt = (None, None) if hard_only else (torch.tensor(g["tea_img"])[rows].double(), torch.tensor(g["tea_txt"])[rows].double())
    eng = DoubleEngine()
    xc = pl.LocalExchange() if group is None else pl.CollectiveExchange(group)
    out, saved = pl.pipeline_forward(eng, xc, si, st, ti, tt, None if hard_only else T, weights, extra)
    gi, gt = pl.pipeline_backward(eng, saved, ups)
    return out, gi, gt, xc, saved


@pytest.mark.parametrize("name", CLIP)
def test_pipeline_cos_diff_and_logits_mse_from_embeddings(name, flow):
    """CLIPCosDiff (clip_cos_diff.py:5-23) and LogitsMSE (logits_mse.py:9-10) from the same tiles: values = the reference on
    materialised logits (0.5 (i2t + t2i), _loss.py:138-145); gradients incl. the diagonal term and exact off-diagonal set."""
    g = golden(name)
    T = float(g["temperature"])
    s, _ = cf.clip_logits(g["stu_img"], g["stu_txt"])
    t, _ = cf.clip_logits(g["tea_img"], g["tea_txt"])
    cd, g_cd = cf.cos_diff(s, t)
    cd2, g_cd2 = cf.cos_diff(s.T, t.T)
    lm, g_lm = cf.logits_mse(s, t)
    w = (0.2, 0.1, 1.0, 1.0, 0.4, 0.3, 2.0, 0.5)        # p_hard, p_soft, s_hard, s_soft, p_cos, p_mse, s_cos, s_mse
    out, gi, gt, _, _ = _run_pipeline(g, w, (torch.tensor(1.0), None, None, torch.tensor(0.25), None), extra=True)
    assert float(out[5]) == pytest.approx(0.5 * (cd + cd2), rel=1e-10) and float(out[6]) == pytest.approx(lm, rel=1e-10)
    assert float(out[7]) == pytest.approx(2.0 * 0.5 * (cd + cd2), rel=1e-10)
    hard, soft = float(g["hard_f64"]), float(g["soft_f64"])
    assert float(out[4]) == pytest.approx(0.2 * hard + 0.1 * soft + 0.4 * 2.0 * 0.5 * (cd + cd2) + 0.3 * 0.5 * lm, rel=1e-9)
    ref = cf.contrastive_from_embeddings(g["stu_img"], g["stu_txt"], g["tea_img"], g["tea_txt"], T, w_hard=0.2, w_soft=0.1)
    w_cos = 0.4 * 2.0 + 0.25 * 2.0                          # g_total * p * s + g_cos * s
    dl = ref["d_logits"] + w_cos * 0.5 * (g_cd + g_cd2.T) + 0.3 * 0.5 * g_lm
    d_img, d_txt = cf.clip_logits_backward(g["stu_img"], g["stu_txt"], dl)
    assert rel_l2(gi.numpy(), d_img) <= 1e-8 and rel_l2(gt.numpy(), d_txt) <= 1e-8


@pytest.mark.parametrize("name", CLIP)
def test_pipeline_decomposition_matches_reference_golden(name, flow):
    g = golden(name)
    one = torch.tensor(1.0)
    out, gi, gt, _, _ = _run_pipeline(g, (1.0, 0.0, 1.0, 1.0), (one, None, None))
    assert float(out[0]) == pytest.approx(float(g["hard_f64"]), rel=1e-10)
    assert float(out[1]) == pytest.approx(float(g["soft_f64"]), rel=1e-9)
    assert rel_l2(gi.numpy(), g["dhard_img_f64"]) <= 1e-9
    assert rel_l2(gt.numpy(), g["dhard_txt_f64"]) <= 1e-9
    out, gi, gt, _, _ = _run_pipeline(g, (0.0, 1.0, 1.0, 1.0), (one, None, None))
    assert rel_l2(gi.numpy(), g["dsoft_img_f64"]) <= 1e-8
    assert rel_l2(gt.numpy(), g["dsoft_txt_f64"]) <= 1e-8


def test_pipeline_weighting_and_upstream_routes(flow):
    """out = {hard, soft, hard s_h, soft s_s, p_h hard s_h + p_s soft s_s} (reference _loss.py:231-234); gradients for an
    upstream on `total` plus one on the scaled soft term = the oracle with the combined weights."""
    g = golden(CLIP[1])
    T = float(g["temperature"])
    p_h, p_s, s_h, s_s = 0.3, 0.7, 2.0, 0.25
    g_total, g_soft = torch.tensor(1.5), torch.tensor(-0.5)
    out, gi, gt, _, _ = _run_pipeline(g, (p_h, p_s, s_h, s_s), (g_total, None, g_soft))
    hard, soft = float(g["hard_f64"]), float(g["soft_f64"])
    assert out[:5].numpy() == pytest.approx([hard, soft, hard * s_h, soft * s_s, p_h * hard * s_h + p_s * soft * s_s], rel=1e-9)
    ref = cf.contrastive_from_embeddings(g["stu_img"], g["stu_txt"], g["tea_img"], g["tea_txt"], T,
                                         w_hard=1.5 * p_h * s_h, w_soft=1.5 * p_s * s_s - 0.5 * s_s)
    assert rel_l2(gi.numpy(), ref["d_img"]) <= 1e-8 and rel_l2(gt.numpy(), ref["d_txt"]) <= 1e-8


def test_pipeline_hard_only_and_buffer_sets(flow):
    from distillclip_b200 import pipeline as pl
    g = golden(CLIP[0])
    out, gi, gt, xc, saved = _run_pipeline(g, (1.0, 0.0, 1.0, 1.0), (torch.tensor(1.0), None, None), hard_only=True)
    assert float(out[0]) == pytest.approx(float(g["hard_f64"]), rel=1e-10) and float(out[1]) == 0.0
    assert rel_l2(gi.numpy(), g["dhard_img_f64"]) <= 1e-9
    with pytest.raises(RuntimeError, match="already ran"):                 # buffers were released by the first backward
        pl.pipeline_backward(DoubleEngine(), saved, (torch.tensor(1.0), None, None))
    # more than MAX_SETS forwards in flight: the oldest is recycled and its backward refuses to run on overwritten buffers
    eng, xc = DoubleEngine(), pl.LocalExchange()
    si, st = torch.tensor(g["stu_img"]).double(), torch.tensor(g["stu_txt"]).double()
    fw = [pl.pipeline_forward(eng, xc, si, st, None, None, None)[1] for _ in range(pl.MAX_SETS + 1)]
    with pytest.raises(RuntimeError, match="recycled"):
        pl.pipeline_backward(eng, fw[0], (torch.tensor(1.0), None, None))
    gi2, _ = pl.pipeline_backward(eng, fw[-1], (torch.tensor(1.0), None, None))
    assert rel_l2(gi2.numpy(), g["dhard_img_f64"]) <= 1e-9


def _pipeline_worker(rank, world, port, name, q, split=False):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    DoubleEngine.split = split
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from distillclip_b200 import pipeline as pl
    g = golden(name)
    b = g["stu_img"].shape[0] // world
    pl.check_equal_batches(dist.group.WORLD, b, torch.device("cpu"))
    res = []
    for step in range(3):                          # three steps: buffer sets are reused, results must not change
        out, gi, gt, _, _ = _run_pipeline(g, (0.75, 0.5, 1.0, 1.0, 0.25, 0.5, 1.0, 1.0), (torch.tensor(1.0), None, None),
                                          group=dist.group.WORLD, rows=slice(rank * b, (rank + 1) * b), extra=True)
        res.append((out.numpy(), gi.numpy(), gt.numpy()))
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,world,split", [("clip_b24_d32_t2", 2, False), ("clip_b40_d64_t4", 2, True), ("clip_b24_d32_t2", 3, True)])
def test_pipeline_row_sharded_gloo(name, world, split):
    """The pipeline's three exchanges (text rows, statistics slots, text-gradient partial sums) under gloo with the
    collective-based exchange: every rank obtains the global losses; gradients concatenate to the oracle's."""
    g = golden(name)
    T = float(g["temperature"])
    ref = cf.contrastive_from_embeddings(g["stu_img"], g["stu_txt"], g["tea_img"], g["tea_txt"], T, w_hard=0.75, w_soft=0.5)
    s_l, _ = cf.clip_logits(g["stu_img"], g["stu_txt"])
    t_l, _ = cf.clip_logits(g["tea_img"], g["tea_txt"])
    cd, g_cd = cf.cos_diff(s_l, t_l)
    lm, g_lm = cf.logits_mse(s_l, t_l)
    _, g_cd2 = cf.cos_diff(s_l.T, t_l.T)
    d_img, d_txt = cf.clip_logits_backward(g["stu_img"], g["stu_txt"], ref["d_logits"] + 0.25 * 0.5 * (g_cd + g_cd2.T) + 0.5 * g_lm)
    ref = dict(ref, d_img=d_img, d_txt=d_txt, cos=cd, mse=lm)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 31 * world) % 2000
    procs = [ctx.Process(target=_pipeline_worker, args=(r, world, port, name, q, split)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for step in range(3):
        for rank, steps in res:
            out = steps[step][0]
            assert float(out[0]) == pytest.approx(ref["hard"], rel=1e-10)
            assert float(out[1]) == pytest.approx(ref["soft"], rel=1e-9)
            assert float(out[4]) == pytest.approx(0.75 * ref["hard"] + 0.5 * ref["soft"] + 0.25 * ref["cos"] + 0.5 * ref["mse"], rel=1e-9)
            assert float(out[5]) == pytest.approx(ref["cos"], rel=1e-10) and float(out[6]) == pytest.approx(ref["mse"], rel=1e-10)
        assert rel_l2(np.concatenate([r[1][step][1] for r in res]), ref["d_img"]) <= 1e-8
        assert rel_l2(np.concatenate([r[1][step][2] for r in res]), ref["d_txt"]) <= 1e-8


def test_unequal_batches_are_rejected_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 77) % 2000
    procs = [ctx.Process(target=_unequal_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all("same per-rank batch" in r for r in res)


def _unequal_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from distillclip_b200 import pipeline as pl
    try:
        pl._CHECKED_BATCH.clear()
        # both ranks see a NEW (group, batch) pair, so both reach the check
        pl.check_equal_batches(dist.group.WORLD, 8 + rank, torch.device("cpu"))
        q.put("no error")
    except ValueError as e:
        q.put(str(e))
    dist.barrier()
    dist.destroy_process_group()
