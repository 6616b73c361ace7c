#!/usr/bin/env python
"""Benchmark of the distillation-loss hot path (BASELINE.json metric: distill-loss fwd+bwd samples/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--no-extras]

A "step" is one fwd+bwd pass of the loss stack over one batch of synthetic input (seed 2022).
Primary workload (default): BASELINE configs[1], the image-encoder stage -- attention-map KL + hidden MSE over L=4
layer pairs, batch 256, 50 tokens, 12 heads, width 768, bf16.  It does not shard (per-sample sums, no exchange step),
so with --gpus N every rank runs an independent replica ("scaling": "weak").  The path that does shard -- the fused
InfoNCE + logit-KL kernels, rows split over ranks with an embedding all-gather -- is timed in the same run and reported
under "contrastive" (BASELINE configs[3] and [4]), so every line carries both rooflines.

Timing: CUDA events on the launching stream per step, summed over exactly K steps after W warm-ups, barrier +
synchronize on both sides, max over ranks.  Inputs larger than L2 are simply re-read; for working sets below 126 MB
the L2 is flushed (256 MiB memset) before every timed step, outside the event pair.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: kind, params
    "image_stage": dict(kind="tower", model_type="image", batch=256, tokens=50, heads=12, width=768, layers=4,
                        names=["attention_probs_kl", "hidden_rep_mse"],
                        desc="BASELINE configs[1]: attn-map KL + hidden MSE, L=4, B=256, N=50, H=12, W=768, bf16"),
    "text_stage": dict(kind="tower", model_type="text", batch=512, tokens=77, heads=8, width=512, layers=4,
                       names=["attention_probs_kl", "hidden_rep_mse", "embedding_mse"],
                       desc="BASELINE configs[2]: attn-map KL + hidden MSE + embedding MSE, L=4, B=512, N=77, H=8, W=512, bf16"),
    "lclip": dict(kind="clip", batch=4096, dim=512, temperature=2.0,
                  desc="BASELINE configs[3]: global-batch InfoNCE + teacher logit KL, B=4096, D=512, bf16, rows sharded over ranks"),
    "sweep": dict(kind="clip", batch=32768, dim=768, temperature=2.0,
                  desc="BASELINE configs[4]: InfoNCE + logit KL, B=32768, D=768, bf16, logits never materialised, rows sharded over ranks"),
}
L2_BYTES = 126 * 1024 * 1024


def ncu_traffic(csv_name, kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` extract (bytes), or None."""
    import csv
    path = os.path.join(ROOT, "profiles", csv_name)
    if not os.path.exists(path):
        return None
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        vals = [float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]] for r in rows[2:] if kernel_substr in r[ik]]
        return int(sum(vals) / len(vals)) if vals else None
    except (ValueError, KeyError, IndexError):
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi in the background during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) != 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = max(mx, float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = sorted(sm)[len(sm) // 2:]                      # upper half = samples taken under load
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------
def make_tower(cfg, device, gen):
    b, n, h, w, layers = cfg["batch"], cfg["tokens"], cfg["heads"], cfg["width"], cfg["layers"]

    def attn():
        return torch.softmax(torch.randn(b, h, n, n, device=device, generator=gen), dim=-1).to(torch.bfloat16)

    def hid():
        return torch.randn(b, n, w, device=device, generator=gen).to(torch.bfloat16)
    d = dict(attention_probs=[attn() for _ in range(layers)], representations=[hid() for _ in range(layers)])
    if "embedding_mse" in cfg["names"]:
        d["embedding"] = hid()
    return d


def tower_elements(cfg):
    b, n, h, w, layers = cfg["batch"], cfg["tokens"], cfg["heads"], cfg["width"], cfg["layers"]
    el = {"attention_probs_kl": layers * b * h * n * n, "hidden_rep_mse": layers * b * n * w, "embedding_mse": b * n * w}
    return {k: el[k] for k in cfg["names"]}


def make_clip(cfg, device, gen, rows, offset):
    """Rank-local rows [offset, offset+rows) of one global seeded batch (every rank draws the same global tensors)."""
    b, d = cfg["batch"], cfg["dim"]
    ti = torch.randn(b, d, device=device, generator=gen)
    tt = ti * 0.6 + 0.8 * torch.randn(b, d, device=device, generator=gen)
    si = ti + 0.5 * torch.randn(b, d, device=device, generator=gen)
    st = tt + 0.5 * torch.randn(b, d, device=device, generator=gen)
    loc = slice(offset, offset + rows)
    return [x[loc].to(torch.bfloat16).contiguous() for x in (si, st, ti, tt)]


# ------------------------------------------------------------------------------------------------
# timing helpers
# ------------------------------------------------------------------------------------------------
class Timer:
    """Times exactly `steps` executions of `fn` with one CUDA-event pair per step (summed), after `warmup` untimed
    executions, barrier + synchronize on both sides, max over ranks.  With graph=True the step is captured once into
    a CUDA graph (the C ABI only enqueues on the current stream and never allocates or synchronises) and each timed
    step is one replay, so the number is device throughput rather than Python launch overhead; if capture fails the
    step runs eagerly and `mode` says so."""

    def __init__(self, device, flush: bool):
        self.flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device) if flush else None
        self.mode = "eager"

    def _capture(self, fn):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g

    def run(self, fn, steps, warmup, dist=None, graph=True):
        run_step = fn
        if graph:
            try:
                g = self._capture(fn)
                run_step, self.mode = g.replay, "cuda-graph replay"
            except Exception as e:                                  # noqa: BLE001 -- report and fall back to eager
                torch.cuda.synchronize()
                self.mode = f"eager (graph capture failed: {type(e).__name__})"
        for _ in range(warmup):
            run_step()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        pairs = []
        for _ in range(steps):
            if self.flush_buf is not None:
                self.flush_buf.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            run_step()
            e.record()
            pairs.append((s, e))
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        total_ms = sum(s.elapsed_time(e) for s, e in pairs)
        if dist is not None:
            t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms


def time_kernel(fn, iters, device):
    """Average duration of the launch(es) in `fn`, measured on the launching stream with one event pair per launch.
    `fn` is captured into a CUDA graph and replays alternate with a 256 MiB L2 flush, all enqueued ahead of the GPU,
    so each pair brackets exactly one cold-L2 execution and no host time."""
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    pairs = []
    for _ in range(iters + 2):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        e.record()
        pairs.append((s, e))
    torch.cuda.synchronize()
    times = [s.elapsed_time(e) for s, e in pairs[2:]]
    return sum(times) / len(times)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def bench_tower(cfg, name, args, device, dist, world, pk, with_cpu):
    from distillclip_b200 import _lib, ops
    from distillclip_b200.model import LossCalculator, TextTransformerOutput, VisionTransformerOutput
    gen = torch.Generator(device=device).manual_seed(2022)
    stu, tea = make_tower(cfg, device, gen), make_tower(cfg, device, gen)
    cls = VisionTransformerOutput if cfg["model_type"] == "image" else TextTransformerOutput
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        calc = LossCalculator(cfg["names"]).to(device)

    leaves = []

    def wrap(d, grad):
        out = {}
        for k, v in d.items():
            if isinstance(v, list):
                out[k] = [x.detach().requires_grad_(grad) for x in v]
                leaves.extend(out[k] if grad else [])
            else:
                out[k] = v.detach().requires_grad_(grad)
                if grad:
                    leaves.append(out[k])
        return cls(**out)
    stu_out, tea_out = wrap(stu, True), wrap(tea, False)

    def step():
        for x in leaves:
            x.grad = None
        loss, _ = calc(stu_out, tea_out, cfg["model_type"])
        loss.backward()
        return loss

    elements = tower_elements(cfg)
    total_el = sum(elements.values())
    algo_bytes = 6 * total_el                                   # read s, read t, write ds; bf16
    in_bytes = 4 * total_el
    timer = Timer(device, flush=in_bytes < L2_BYTES)
    step()
    _lib.LAUNCHES = 0
    step()
    launches = _lib.LAUNCHES * args.steps                         # our kernels per step x timed steps
    total_ms = timer.run(step, args.steps, args.warmup, dist)
    ms = total_ms / args.steps
    value = world * cfg["batch"] / (ms * 1e-3)

    # dominant kernel = the single tower launch (all layers of all terms + weighting), timed alone through the raw C-ABI
    # call on the same inputs and stream with preallocated outputs; the per-family kernels are timed the same way for
    # reference (they serve the per-module API).
    fields = {"hidden_rep_mse": (ops.KIND_MSE, "representations"), "attention_probs_kl": (ops.KIND_ATTN_KL, "attention_probs"),
              "embedding_mse": (ops.KIND_MSE, "embedding")}
    entries = []
    for nm in cfg["names"]:
        kind, field = fields[nm]
        sv, tv = stu[field], tea[field]
        sv, tv = (sv if isinstance(sv, list) else [sv]), (tv if isinstance(tv, list) else [tv])
        entries.append((kind, len(sv), [x.detach() for x in sv], tv, [True] * len(sv), 1.0 / len(cfg["names"])))
    w = [1.0] * len(entries)
    pct = [1.0 / len(entries)] * len(entries)
    bufs = ops.launch_tower(entries, w, pct)
    kernels = {"tower_stream_kernel": (lambda: ops.launch_tower(entries, w, pct, out=bufs), algo_bytes)}
    s_h, t_h = [x.detach() for x in stu["representations"]], tea["representations"]
    p_h, _, g_h = ops.launch_mse(s_h, t_h, len(s_h), 1.0, [True] * len(s_h))
    kernels["mse_stream_kernel"] = (lambda: ops.launch_mse(s_h, t_h, len(s_h), 1.0, [True] * len(s_h), out=(p_h, g_h)),
                                    6 * elements["hidden_rep_mse"])
    s_a, t_a = [x.detach() for x in stu["attention_probs"]], tea["attention_probs"]
    p_a, _, g_a = ops.launch_attn_kl(s_a, t_a, len(s_a), 1.0, [True] * len(s_a))
    kernels["attn_kl_kernel"] = (lambda: ops.launch_attn_kl(s_a, t_a, len(s_a), 1.0, [True] * len(s_a), out=(p_a, g_a)),
                                 6 * elements["attention_probs_kl"])
    kres = {}
    for kname, (fn, nbytes) in kernels.items():
        k_ms = time_kernel(fn, 20, device)
        kres[kname] = {"ms": round(k_ms, 5), "algorithmic_bytes": nbytes, "gbs": round(nbytes / k_ms / 1e6, 1),
                       "frac": round(nbytes / k_ms / 1e6 / pk["hbm"], 4)}
    dom = "tower_stream_kernel"
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kres[dom]["gbs"], "peak": pk["hbm"], "unit": "GB/s",
                "frac": kres[dom]["frac"],
                "traffic": ncu_traffic("r01_ncu_full_tower_stream.csv", "tower_stream") if name == "image_stage" else None,
                "traffic_note": "dram bytes read + written inside one isolated launch (ncu --set full, profiles/); part of the "
                                "gradient writes is still dirty in L2 when the kernel exits",
                "peak_source": pk["source"],
                "step_frac": round(algo_bytes / ms / 1e6 / pk["hbm"], 4), "kernels": kres}

    # end to end through the module API with HOST (pinned) buffers: H2D of every input + D2H of the loss inside the timer
    host = {k: ([x.cpu().pin_memory() for x in v] if isinstance(v, list) else v.cpu().pin_memory()) for k, v in stu.items()}
    host_t = {k: ([x.cpu().pin_memory() for x in v] if isinstance(v, list) else v.cpu().pin_memory()) for k, v in tea.items()}
    h2d = sum(x.numel() * 2 for d in (host, host_t) for v in d.values() for x in (v if isinstance(v, list) else [v]))
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def e2e_step():
        def up(d, grad):
            o = {}
            for k, v in d.items():
                o[k] = ([x.to(device, non_blocking=True).requires_grad_(grad) for x in v] if isinstance(v, list)
                        else v.to(device, non_blocking=True).requires_grad_(grad))
            return cls(**o)
        loss, _ = calc(up(host, True), up(host_t, False), cfg["model_type"])
        loss.backward()
        loss_host.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e2e_ms = Timer(device, flush=False).run(e2e_step, max(3, args.steps // 4), 2, dist, graph=False) / max(3, args.steps // 4)
    e2e = {"value": round(world * cfg["batch"] / (e2e_ms * 1e-3), 1), "unit": "samples/s", "ms_per_step": round(e2e_ms, 4),
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}

    cpu = cpu_tower(cfg, stu, tea, budget_s=20.0) if with_cpu else None
    return dict(value=value, ms=ms, launches=launches, roofline=roofline, e2e=e2e, cpu=cpu,
                flush=timer.flush_buf is not None, algo_bytes=algo_bytes, mode=timer.mode)


def cpu_tower(cfg, stu, tea, budget_s):
    """The oracle's torch port of the reference path on the host cores, fp32 copies, bounded time."""
    from oracle import torch_port as tp
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    names = cfg["names"]

    def to_cpu(d, grad):
        return {k: ([x.float().cpu().requires_grad_(grad) for x in v] if isinstance(v, list)
                    else v.float().cpu().requires_grad_(grad)) for k, v in d.items()}
    s, t = to_cpu(stu, True), to_cpu(tea, False)
    times, t_end = [], time.perf_counter() + budget_s
    while len(times) < 5 and (time.perf_counter() < t_end or not times):
        for v in s.values():
            for x in (v if isinstance(v, list) else [v]):
                x.grad = None
        t0 = time.perf_counter()
        tp.stage_step_cpu(names, s, t, threads=threads)
        times.append(time.perf_counter() - t0)
    best = min(times)
    return {"value": round(cfg["batch"] / best, 1), "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": f"{len(times)} full steps of the same workload (best of), oracle/torch_port.py fp32, {threads} threads"}


def cpu_clip(cfg, budget_s):
    """Oracle torch port of the reference path (normalise, matmul, HardLabel + SoftLabel both directions, autograd) on the
    host cores; the B x B fp32 logits bound the sample to 4096 rows."""
    from oracle import torch_port as tp
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    b = min(cfg["batch"], 4096)
    gen = torch.Generator().manual_seed(2022)
    si, st, ti, tt = [x.float() for x in make_clip(dict(cfg, batch=b), "cpu", gen, b, 0)]
    si.requires_grad_(True)
    st.requires_grad_(True)
    times, t_end = [], time.perf_counter() + budget_s
    while len(times) < 3 and (time.perf_counter() < t_end or not times):
        si.grad = st.grad = None
        t0 = time.perf_counter()
        tp.stage_step_cpu(["hard_label", "soft_label"], {"visual": {"last_representation": si}, "text": {"last_representation": st}},
                          {"visual": {"last_representation": ti}, "text": {"last_representation": tt}},
                          temperature=cfg["temperature"], two=True, threads=threads)
        times.append(time.perf_counter() - t0)
    return {"value": round(b / min(times), 1), "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": f"{len(times)} fwd+bwd steps on a {b}-row batch (best of), oracle/torch_port.py fp32, {threads} threads"}


def bench_clip(cfg, args, device, dist, rank, world, pk, steps, warmup):
    from distillclip_b200 import _lib
    from distillclip_b200.contrastive import clip_contrastive
    b, d, T = cfg["batch"], cfg["dim"], cfg["temperature"]
    rows = b // world
    gen = torch.Generator(device=device).manual_seed(2022)
    si, st, ti, tt = make_clip(cfg, device, gen, rows, rank * rows)
    si.requires_grad_(True)
    st.requires_grad_(True)
    group = dist.group.WORLD if (dist is not None and world > 1) else None

    def step():
        si.grad = None
        st.grad = None
        res = clip_contrastive(si, st, ti, tt, T, want_hard=True, want_soft=True, group=group)
        (0.5 * res["hard_label"] + 0.5 * res["soft_label"]).backward()
    timer = Timer(device, flush=4 * b * d * 2 < L2_BYTES)
    step()
    _lib.LAUNCHES = 0
    step()
    launches = _lib.LAUNCHES * steps
    total_ms = timer.run(step, steps, warmup, dist, graph=True)      # NCCL collectives are graph-capturable; falls back to eager
    ms = total_ms / steps
    flops = 12.0 * b * b * d                                     # credited, whole job
    tf = flops / (ms * 1e-3) / 1e12
    # end to end: pinned host embeddings -> device, fused fwd+bwd through the public call, loss back to the host
    host = [x.detach().cpu().pin_memory() for x in (si, st, ti, tt)]
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def e2e_step():
        a, c_, e, f = [h.to(device, non_blocking=True) for h in host]
        a.requires_grad_(True)
        c_.requires_grad_(True)
        res = clip_contrastive(a, c_, e, f, T, want_hard=True, want_soft=True, group=group)
        loss = 0.5 * res["hard_label"] + 0.5 * res["soft_label"]
        loss.backward()
        loss_host.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    n_e2e = max(3, steps // 4)
    e2e_ms = Timer(device, flush=False).run(e2e_step, n_e2e, 2, dist, graph=False) / n_e2e
    e2e = {"value": round(b / (e2e_ms * 1e-3), 1), "unit": "samples/s", "ms_per_step": round(e2e_ms, 4),
           "h2d_bytes_per_step": sum(h.numel() * 2 for h in host), "d2h_bytes_per_step": 4}
    return {"workload": cfg["desc"], "global_batch": b, "dim": d, "temperature": T, "n_gpus": world,
            "value": round(b / (ms * 1e-3), 1), "unit": "samples/s", "ms_per_step": round(ms, 4), "steps": steps,
            "scaling": "strong", "gpu_launches": launches, "l2_flush": timer.flush_buf is not None, "timing": timer.mode,
            "e2e": e2e,
            "roofline": {"bound": "tensor", "kernel": "clip_fwd_kernel + clip_bwd_pair_kernel + clip_gt_gemm_kernel (fused tcgen05)",
                         "achieved": round(tf, 2), "peak": pk["tf_burst"] * world, "unit": "TFLOP/s",
                         "frac": round(tf / (pk["tf_burst"] * world), 4),
                         "frac_of_sustained": round(tf / (pk["tf_sustained"] * world), 4),
                         "credited_flops": flops,
                         "traffic": (sum(ncu_traffic("r01_ncu_full_clip_sweep.csv", k) or 0 for k in
                                         ("clip_fwd_kernel", "clip_bwd_pair_kernel", "clip_gt_gemm_kernel")) or None)
                         if (b, d, world) == (32768, 768, 1) else None,
                         "traffic_note": "dram bytes of the three tcgen05 kernels of one step (ncu --set full, profiles/), 4.3 GB of "
                                         "which are the fp16 gradient tiles written once and read once",
                         "peak_source": pk["source"]}}


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=device)
        dist = dist_mod
    import __graft_entry__
    __graft_entry__.build()
    pk = peaks()
    name = args.workload
    cfg = WORKLOADS[name]
    line = {}
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.__enter__()                                           # nvidia-smi -lms 100 for the whole measurement phase
    if cfg["kind"] == "tower":
        r = bench_tower(cfg, name, args, device, dist, world, pk, with_cpu=False)
        line = {"metric": "distill-loss fwd+bwd samples/sec", "value": round(r["value"], 1), "unit": "samples/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["ms"], 5),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 in, fp32 accumulate",
                "data": "synthetic (seed 2022)",
                "config": {"workload": cfg["desc"], "parallelism": f"{world} independent replica(s); this path has no exchange step",
                           "l2": "flushed before every timed step" if r["flush"] else "inputs (student+teacher) larger than the 126 MB L2",
                           "algorithmic_bytes_per_step": r["algo_bytes"], "timing": r["mode"]},
                "roofline": r["roofline"], "e2e": r["e2e"], "gpu_launches": r["launches"]}
        extras = []
        if not args.no_extras:
            for cname in ("lclip", "sweep"):
                ccfg = WORKLOADS[cname]
                if ccfg["batch"] % (128 * world):
                    continue
                steps = max(3, min(args.steps, 20 if cname == "lclip" else 5))
                extras.append(bench_clip(ccfg, args, device, dist, rank, world, pk, steps, 3))
            if world == 1:
                t = bench_tower(WORKLOADS["text_stage"], "text_stage", args, device, dist, world, pk, with_cpu=False)
                line["text_stage"] = {"workload": WORKLOADS["text_stage"]["desc"], "value": round(t["value"], 1),
                                      "unit": "samples/s", "ms_per_step": round(t["ms"], 5), "roofline": t["roofline"],
                                      "e2e": t["e2e"], "gpu_launches": t["launches"], "timing": t["mode"]}
        line["contrastive"] = extras
        sampler.__exit__(None, None, None)
        line["clocks"] = sampler.summary()
        if rank == 0 and world == 1 and not args.no_cpu:             # CPU leg after the GPU clocks have been sampled
            gen = torch.Generator(device=device).manual_seed(2022)
            line["cpu_baseline"] = cpu_tower(cfg, make_tower(cfg, device, gen), make_tower(cfg, device, gen), budget_s=20.0)
    else:
        c = bench_clip(cfg, args, device, dist, rank, world, pk, args.steps, args.warmup)
        sampler.__exit__(None, None, None)
        line = {"metric": "distill-loss fwd+bwd samples/sec", "value": c["value"], "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": c["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16 in, fp32 accumulate", "data": "synthetic (seed 2022)",
                "config": {"workload": cfg["desc"], "parallelism": f"rows sharded over {world} rank(s), embedding all-gather",
                           "l2": "flushed before every timed step" if c["l2_flush"] else "inputs larger than L2"},
                "roofline": c["roofline"], "e2e": c["e2e"], "gpu_launches": c["gpu_launches"], "clocks": sampler.summary()}
        if rank == 0 and world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_clip(cfg, budget_s=20.0)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the same path (oracle port; the reference tree does not travel)
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import torch_port as tp
    cfg = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    gen = torch.Generator().manual_seed(2022)
    if cfg["kind"] == "tower":
        stu, tea = make_tower(cfg, "cpu", gen), make_tower(cfg, "cpu", gen)

        def to32(d, grad):
            return {k: ([x.float().requires_grad_(grad) for x in v] if isinstance(v, list) else v.float().requires_grad_(grad))
                    for k, v in d.items()}
        s, t = to32(stu, True), to32(tea, False)

        def step():
            for v in s.values():
                for x in (v if isinstance(v, list) else [v]):
                    x.grad = None
            tp.stage_step_cpu(cfg["names"], s, t, threads=threads)
        batch, sample = cfg["batch"], "each step = one full fwd+bwd of the same workload"
    else:
        b = min(cfg["batch"], 4096)                               # B x B fp32 logits: bounded sample of rows
        sub = dict(cfg, batch=b)
        si, st, ti, tt = [x.float() for x in make_clip(sub, "cpu", gen, b, 0)]
        si.requires_grad_(True)
        st.requires_grad_(True)
        names = ["hard_label", "soft_label"]

        def step():
            si.grad = None
            st.grad = None
            stu = {"visual": {"last_representation": si}, "text": {"last_representation": st}}
            tea = {"visual": {"last_representation": ti}, "text": {"last_representation": tt}}
            tp.stage_step_cpu(names, stu, tea, temperature=cfg["temperature"], two=True, threads=threads)
        batch, sample = b, f"each step = full fwd+bwd on a {b}-row sub-batch (B x B fp32 logits do not fit the time budget at B={cfg['batch']})"
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    ms = (time.perf_counter() - t0) / steps * 1e3
    value = round(batch / (ms * 1e-3), 1)
    line = {"impl": "reference", "metric": "distill-loss fwd+bwd samples/sec", "value": value, "unit": "samples/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": warmup, "ms_per_step": round(ms, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32 (bf16 inputs upcast)",
            "data": "synthetic (seed 2022)", "config": {"workload": cfg["desc"]},
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="image_stage")
    ap.add_argument("--no-extras", action="store_true", help="skip the contrastive / text-stage sub-benchmarks")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
