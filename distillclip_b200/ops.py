"""Host side of the streaming losses: launches through the C ABI + custom autograd.

The one-pass kernels write the student gradients during the forward pass, pre-multiplied by the
upstream gradient the caller is assumed to send (`percent * scale` when called from LossCalculator, else 1;
times a GradScaler's device-side scale when one is registered).  `backward` launches the same kernel in
"regrad" mode: it compares the real upstream scalar on the device and exits without touching HBM when
the assumption held -- fwd+bwd then costs read s + read t + write ds and nothing more -- and otherwise
recomputes the gradients from the inputs with the true value (one rounding, at the true magnitude).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

_DT = {torch.bfloat16: _lib.BF16, torch.float16: _lib.F16, torch.float32: _lib.F32}


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"distillclip_b200: unsupported dtype {t.dtype} (bf16 / fp16 / fp32 only)") from None


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise _lib.DistillClipB200Error(
            f"{what} is on {t.device}: distillclip_b200 runs on CUDA (sm_100a) only and has no CPU fallback")


def _prep_pair(stu: Sequence[torch.Tensor], tea: Sequence[torch.Tensor]):
    """zip()-truncate like the reference loops do, make contiguous, unify dtypes."""
    n = min(len(stu), len(tea))
    s_out, t_out = [], []
    for s, t in zip(stu[:n], tea[:n]):
        _require_cuda(s, "student tensor")
        _require_cuda(t, "teacher tensor")
        dtype_code(s)
        if t.dtype != s.dtype:
            t = t.to(s.dtype)
        s_out.append(s if s.is_contiguous() else s.contiguous())
        t_out.append(t.detach() if t.is_contiguous() else t.detach().contiguous())
    if n and any(s.dtype != s_out[0].dtype for s in s_out):
        raise TypeError("distillclip_b200: all student layers of one loss must share a dtype")
    return s_out, t_out


# ----------------------------------------------------------------------------------------------
# raw launches (no autograd)
# ----------------------------------------------------------------------------------------------
def _alloc_partials(nchunk: int, dev):
    if nchunk == 1:
        return torch.empty(_lib.MAX_PARTIALS, dtype=torch.float64, device=dev)
    return torch.zeros(nchunk * _lib.MAX_PARTIALS, dtype=torch.float64, device=dev)


def launch_mse(stu: List[torch.Tensor], tea: List[torch.Tensor], divisor: int, grad_scale: float,
               need_grad: Sequence[bool], grad_dtype: Optional[torch.dtype] = None, out=None):
    """-> (partials, count, grads). Shapes must match pairwise (same rule as nn.MSELoss without broadcasting).
    `out=(partials, grads)` reuses buffers from an earlier call (no allocation: used for kernel-only timing)."""
    dev = stu[0].device
    nchunk = (len(stu) + _lib.MAX_LAYERS - 1) // _lib.MAX_LAYERS
    if out is not None:
        partials, grads = out
    else:
        grads = []
        for s, t, ng in zip(stu, tea, need_grad):
            if s.shape != t.shape:
                raise ValueError(f"MSE: student {tuple(s.shape)} and teacher {tuple(t.shape)} shapes differ")
            grads.append(torch.empty_like(s, dtype=grad_dtype or s.dtype) if ng else None)
        partials = _alloc_partials(nchunk, dev)
    gd = _DT[grad_dtype] if grad_dtype is not None else dtype_code(stu[0])
    count = C.c_int(0)
    for c in range(nchunk):
        sl = slice(c * _lib.MAX_LAYERS, (c + 1) * _lib.MAX_LAYERS)
        ss, tt, gg = stu[sl], tea[sl], grads[sl]
        _lib.call("dcb_mse_fwd_bwd", len(ss), _lib.ptr_array([x.data_ptr() for x in ss]),
                  _lib.ptr_array([x.data_ptr() for x in tt]),
                  _lib.ptr_array([g.data_ptr() if g is not None else 0 for g in gg]),
                  _lib.i64_array([x.numel() for x in ss]), dtype_code(ss[0]), gd, int(divisor), float(grad_scale),
                  C.c_void_p(partials.data_ptr() + 8 * c * _lib.MAX_PARTIALS), C.byref(count), _stream_ptr())
    n_part = count.value if nchunk == 1 else nchunk * _lib.MAX_PARTIALS
    return partials, n_part, grads


def launch_attn_kl(stu: List[torch.Tensor], tea: List[torch.Tensor], divisor: int, grad_scale: float,
                   need_grad: Sequence[bool], grad_dtype: Optional[torch.dtype] = None, out=None):
    dev = stu[0].device
    grads: List[Optional[torch.Tensor]] = []
    batch, hs, ht, pos = [], [], [], []
    for s, t, ng in zip(stu, tea, need_grad):
        if s.dim() < 3 or t.dim() != s.dim() or s.shape[0] != t.shape[0] or s.shape[2:] != t.shape[2:]:
            raise ValueError(f"attention maps must be [B, H, ...] with equal B and map size: "
                             f"student {tuple(s.shape)} vs teacher {tuple(t.shape)}")
        batch.append(s.shape[0])
        hs.append(s.shape[1])
        ht.append(t.shape[1])
        pos.append(s.numel() // (s.shape[0] * s.shape[1]))
        if out is None:
            grads.append(torch.empty_like(s, dtype=grad_dtype or s.dtype) if ng else None)
    gd = _DT[grad_dtype] if grad_dtype is not None else dtype_code(stu[0])
    nchunk = (len(stu) + _lib.MAX_LAYERS - 1) // _lib.MAX_LAYERS
    if out is not None:
        partials, grads = out
    else:
        partials = _alloc_partials(nchunk, dev)
    count = C.c_int(0)
    for c in range(nchunk):
        sl = slice(c * _lib.MAX_LAYERS, (c + 1) * _lib.MAX_LAYERS)
        ss, tt, gg = stu[sl], tea[sl], grads[sl]
        _lib.call("dcb_attn_kl_fwd_bwd", len(ss), _lib.ptr_array([x.data_ptr() for x in ss]),
                  _lib.ptr_array([x.data_ptr() for x in tt]),
                  _lib.ptr_array([g.data_ptr() if g is not None else 0 for g in gg]),
                  _lib.i64_array(batch[sl]), _lib.i32_array(hs[sl]), _lib.i32_array(ht[sl]), _lib.i64_array(pos[sl]),
                  dtype_code(ss[0]), gd, int(divisor), float(grad_scale),
                  C.c_void_p(partials.data_ptr() + 8 * c * _lib.MAX_PARTIALS), C.byref(count), _stream_ptr())
    n_part = count.value if nchunk == 1 else nchunk * _lib.MAX_PARTIALS
    return partials, n_part, grads


_TICKETS = {}


def _ticket(dev: torch.device) -> torch.Tensor:
    """Zero-initialised uint32 the tower kernel uses to find its last CTA; the kernel resets it, so one per
    (device, stream) serves every launch on that stream."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    t = _TICKETS.get(key)
    if t is None:
        t = _TICKETS[key] = torch.zeros(1, dtype=torch.int32, device=dev)
    return t


def _seg_shape(kind, s, t):
    if kind in (KIND_MSE, KIND_L1):
        if s.shape != t.shape:
            raise ValueError(f"{kind}: student {tuple(s.shape)} and teacher {tuple(t.shape)} shapes differ")
        return 1, 1, 1, 1
    if kind == KIND_COS:
        if s.shape != t.shape or s.dim() != 2:
            raise ValueError(f"cosine loss takes equal [B, D] tensors: student {tuple(s.shape)}, teacher {tuple(t.shape)}")
        return s.shape[0], 1, 1, s.shape[1]
    if s.dim() < 3 or t.dim() != s.dim() or s.shape[0] != t.shape[0] or s.shape[2:] != t.shape[2:]:
        raise ValueError(f"attention maps must be [B, H, ...] with equal B and map size: "
                         f"student {tuple(s.shape)} vs teacher {tuple(t.shape)}")
    return s.shape[0], s.shape[1], t.shape[1], s.numel() // (s.shape[0] * s.shape[1])


def launch_tower(entries, scale: Sequence[float], percent: Sequence[float], grad_dtype: Optional[torch.dtype] = None,
                 out=None, fwd_mult: Optional[torch.Tensor] = None, regrad=None):
    """All streaming losses of one tower -- every layer of every term, the deterministic reduction and the weighting --
    in ONE launch of the tower kernel (csrc/tower_stream.cu).
    entries: [(kind, divisor, stu list, tea list, need_grad list, grad_scale)], one per loss term.
    -> (out[n_terms + 1] = scaled term values + weighted total, grads per entry, partials).  `out=(out, grads, partials)`
    reuses buffers (kernel-only timing).  fwd_mult: optional device scalar multiplied into every gradient (AMP loss scale).
    regrad = [(upstream 0-dim fp32 device tensor, up_mult, expected)] per entry: gradients only, into out[1], recomputed
    with the true upstream gradient unless it equals what the forward assumed (see include/distillclip_b200.h)."""
    lib = _lib.load()
    dev = entries[0][2][0].device
    in_dt = dtype_code(entries[0][2][0])
    gd = _DT[grad_dtype] if grad_dtype is not None else in_dt
    n_terms = len(entries)
    stride = lib.dcb_tower_grid()
    d = dict(kinds=[], terms=[], stu=[], tea=[], grad=[], numel=[], batch=[], hs=[], ht=[], pos=[], div=[], gsc=[],
             up=[], upm=[], exp=[])
    all_grads = [] if out is None else out[1]
    for ti, (kind, divisor, stu, tea, need, pre) in enumerate(entries):
        grads = [] if out is None else all_grads[ti]
        for li, (s, t, ng) in enumerate(zip(stu, tea, need)):
            b_, hs_, ht_, pos_ = _seg_shape(kind, s, t)
            if out is None:
                grads.append(torch.empty_like(s, dtype=grad_dtype or s.dtype) if ng else None)
            g = grads[li]
            d["kinds"].append(_KIND_CODE[kind]), d["terms"].append(ti)
            d["stu"].append(s.data_ptr()), d["tea"].append(t.data_ptr()), d["grad"].append(g.data_ptr() if g is not None else 0)
            d["numel"].append(s.numel()), d["batch"].append(b_), d["hs"].append(hs_), d["ht"].append(ht_)
            d["pos"].append(pos_), d["div"].append(int(divisor)), d["gsc"].append(float(pre))
            if regrad is not None:
                up, upm, exp = regrad[ti]
                d["up"].append(up.data_ptr()), d["upm"].append(float(upm)), d["exp"].append(float(exp))
        if out is None:
            all_grads.append(grads)
    res = partials = None
    if regrad is None:
        if out is None:
            res = torch.empty(n_terms + 1, dtype=torch.float32, device=dev)
            partials = torch.empty(n_terms * stride, dtype=torch.float64, device=dev)
        else:
            res, partials = out[0], out[2]
    n = len(d["kinds"])
    pad = (lambda v: v if v else [0])
    _lib.call("dcb_tower_fwd_bwd", n, _lib.i32_array(pad(d["kinds"])), _lib.i32_array(pad(d["terms"])), _lib.ptr_array(pad(d["stu"])),
              _lib.ptr_array(pad(d["tea"])), _lib.ptr_array(pad(d["grad"])), _lib.i64_array(pad(d["numel"])), _lib.i64_array(pad(d["batch"])),
              _lib.i32_array(pad(d["hs"])), _lib.i32_array(pad(d["ht"])), _lib.i64_array(pad(d["pos"])), _lib.i32_array(pad(d["div"])),
              _lib.f32_array(pad(d["gsc"])), n_terms, _lib.f32_array(scale), _lib.f32_array(percent), in_dt, gd,
              C.c_void_p(partials.data_ptr()) if partials is not None else None, stride, 0, 0,
              C.c_void_p(_ticket(dev).data_ptr()) if regrad is None else None,
              C.c_void_p(res.data_ptr()) if res is not None else None,
              C.c_void_p(fwd_mult.data_ptr()) if fwd_mult is not None else None,
              _lib.ptr_array(pad(d["up"])) if regrad is not None else None,
              _lib.f32_array(pad(d["upm"])) if regrad is not None else None,
              _lib.f32_array(pad(d["exp"])) if regrad is not None else None, _stream_ptr())
    return res, all_grads, partials


def finalize(terms: Sequence[Tuple[torch.Tensor, int]], scale: Sequence[float], percent: Sequence[float]):
    """out[k] = scale[k]*sum(partials_k) ; out[-1] = sum_k percent[k]*out[k]   (fp32, device)."""
    out = torch.empty(len(terms) + 1, dtype=torch.float32, device=terms[0][0].device)
    _lib.call("dcb_finalize", len(terms), _lib.ptr_array([p.data_ptr() for p, _ in terms]),
              _lib.i32_array([c for _, c in terms]), _lib.f32_array(scale), _lib.f32_array(percent),
              C.c_void_p(out.data_ptr()), _stream_ptr())
    return out


def _as_upstream(g: Optional[torch.Tensor], like: torch.Tensor) -> torch.Tensor:
    if g is None:
        return torch.zeros((), dtype=torch.float32, device=like.device)
    if g.dtype != torch.float32 or not g.is_contiguous():
        g = g.to(torch.float32).contiguous()
    return g


# ----------------------------------------------------------------------------------------------
# autograd
# ----------------------------------------------------------------------------------------------
#: Upstream gradient the one-pass kernels assume when they write the student gradients during the forward pass.
#: 1.0 is right for plain `loss.backward()`.  A different upstream value is ALWAYS honoured exactly: backward compares on
#: the device and, on a mismatch, recomputes the gradients from the inputs with the true value (one more pass over the
#: inputs, a single rounding at the true magnitude -- fp16 AMP is safe without any configuration).  To avoid that extra
#: pass under a GradScaler, hand the scaler to `LossCalculator.grad_scaler` (its device scale tensor is read by the
#: forward kernel) or set this / `LossCalculator.expected_grad_scale` to `scaler.get_scale()`.
EXPECTED_GRAD_SCALE = 1.0

KIND_MSE, KIND_ATTN_KL, KIND_L1, KIND_COS, KIND_ATTN_MSE = "mse", "attn_kl", "l1", "cos", "attn_mse"
_LAUNCH = {KIND_MSE: launch_mse, KIND_ATTN_KL: launch_attn_kl}        # per-family kernels (raw launches, kernel timing)
_KIND_CODE = {KIND_MSE: 0, KIND_ATTN_KL: 1, KIND_L1: 2, KIND_COS: 3, KIND_ATTN_MSE: 4}
_NAN = float("nan")


def grad_scaler_mult(scaler) -> Optional[torch.Tensor]:
    """Device scalar holding a torch.amp.GradScaler's current scale (None when there is none / it is disabled)."""
    if scaler is None or not getattr(scaler, "is_enabled", lambda: True)():
        return None
    t = getattr(scaler, "_scale", None)
    if t is None and hasattr(scaler, "_lazy_init_scale_growth_tracker"):
        scaler._lazy_init_scale_growth_tracker(torch.device("cuda", torch.cuda.current_device()))
        t = getattr(scaler, "_scale", None)
    if t is None:
        return None
    return t if t.dtype == torch.float32 else t.float()


class TowerLossFn(torch.autograd.Function):
    """spec: list of (kind, divisor, n_layers, scale, percent); tensors: for each entry stu_0.., tea_0..
    Returns (total, res_0, ..., res_{k-1}) with res_k = raw_k * scale_k and total = sum res_k * percent_k
    (reference model/_loss.py:195-200).  `expected` = upstream gradient of `total` assumed in the forward pass (times the
    device scalar `fwd_mult` when given).  Terms may differ in dtype (fp16 AMP: softmax outputs are fp32, linear outputs
    fp16): one launch per dtype."""

    @staticmethod
    def forward(ctx, spec, expected, fwd_mult, *tensors):
        ctx.set_materialize_grads(False)
        entries, layout = [], []
        off = 0
        for kind, divisor, n, scale, percent in spec:
            stu, tea = list(tensors[off:off + n]), list(tensors[off + n:off + 2 * n])
            need = [bool(x) for x in ctx.needs_input_grad[3 + off:3 + off + n]]
            w = float(np.float32(percent) * np.float32(scale))
            pre = w * expected if w != 0.0 else 1.0
            entries.append((kind, divisor, stu, tea, need, pre))
            layout.append((off, n, w, pre))
            off += 2 * n
        # one launch per dtype (terms keep their order inside a group)
        groups = {}
        for i, e in enumerate(entries):
            groups.setdefault(e[2][0].dtype, []).append(i)
        for idx in groups.values():
            if len(idx) > _lib.TOWER_MAX_TERMS or sum(len(entries[i][2]) for i in idx) > _lib.TOWER_MAX_SEG:
                raise _lib.DistillClipB200Error("one tower launch takes <= 8 loss terms and <= 40 layer pairs per dtype")
        outs = [None] * len(spec)
        all_grads = [None] * len(spec)
        total = None
        for idx in groups.values():
            res, grads, _ = launch_tower([entries[i] for i in idx], [spec[i][3] for i in idx], [spec[i][4] for i in idx],
                                         fwd_mult=fwd_mult)
            for j, i in enumerate(idx):
                outs[i], all_grads[i] = res[j], grads[j]
            total = res[len(idx)] if total is None else total + res[len(idx)]
        ctx.all_grads, ctx.layout, ctx.spec, ctx.n_in, ctx.groups = all_grads, layout, spec, len(tensors), list(groups.values())
        ctx.expected, ctx.fwd_mult = expected, fwd_mult
        ctx.save_for_backward(*tensors)
        return (total, *outs)

    @staticmethod
    def backward(ctx, g_total, *g_res):
        tensors = ctx.saved_tensors
        ret: List[Optional[torch.Tensor]] = [None] * ctx.n_in
        # First backward: the forward-time buffers are checked on the device and handed over WITHOUT keeping a reference, so
        # AccumulateGrad adopts them as `.grad` (no clone, no extra HBM pass).  Any later backward of the same graph
        # (retain_graph=True) recomputes into fresh buffers: the adopted ones belong to the caller now.
        first = ctx.all_grads is not None
        all_grads, ctx.all_grads = ctx.all_grads, None
        dev = tensors[0].device
        zero = None
        for idx in ctx.groups:
            entries, regrad, bufs = [], [], []
            for i in idx:
                kind, divisor, n, scale, percent = ctx.spec[i]
                off, _, w, pre = ctx.layout[i]
                stu, tea = list(tensors[off:off + n]), list(tensors[off + n:off + 2 * n])
                need = [bool(x) for x in ctx.needs_input_grad[3 + off:3 + off + n]]
                if first:
                    grads = all_grads[i]
                else:
                    grads = [torch.empty_like(s) if ng else None for s, ng in zip(stu, need)]
                if not any(need):
                    continue
                if g_res[i] is None and g_total is not None and pre == w * ctx.expected:
                    # common path, no torch kernels: true multiplier = g_total * w, assumed = expected (* fwd_mult) * w
                    regrad.append((_as_upstream(g_total, stu[0]), w, ctx.expected if first else _NAN))
                else:
                    if g_total is None and g_res[i] is None:
                        if zero is None:
                            zero = torch.zeros((), dtype=torch.float32, device=dev)
                        up = zero
                    else:
                        up = _as_upstream(g_total, stu[0]) * w if g_total is not None else None
                        if g_res[i] is not None:
                            r = _as_upstream(g_res[i], stu[0]) * float(scale)
                            up = r if up is None else up + r
                    regrad.append((up.contiguous(), 1.0, _NAN))          # NaN never compares equal: always recompute
                entries.append((kind, divisor, stu, tea, need, 1.0))
                bufs.append(grads)
                ret[off:off + n] = grads
            if entries:
                launch_tower(entries, [1.0] * len(entries), [1.0] * len(entries), out=(None, bufs, None),
                             fwd_mult=ctx.fwd_mult, regrad=regrad)
        del all_grads
        return (None, None, None, *ret)


def stream_loss(kind: str, stu: Sequence[torch.Tensor], tea: Sequence[torch.Tensor]):
    """Reference list semantics: zip-truncation, divisor len(stu), ZeroDivisionError on empty (F8).  One term through the
    tower kernel (value + gradients in one pass, backward-time check / recompute as for a whole tower)."""
    divisor = len(stu)
    if divisor == 0:
        raise ZeroDivisionError("division by zero")
    s, t = _prep_pair(stu, tea)
    if not s:
        return 0.0          # reference: `res_loss = 0; res_loss /= len(stu)` -> python float
    return TowerLossFn.apply([(kind, divisor, len(s), 1.0, 1.0)], float(EXPECTED_GRAD_SCALE), None, *s, *t)[0]


# ----------------------------------------------------------------------------------------------
# row-softmax losses on pooled outputs [B, D]: OutKLLoss / OutCELoss
# ----------------------------------------------------------------------------------------------
class RowSoftmaxLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, stu, tea, temperature, mode):
        rows, cols = stu.shape
        saved = torch.empty(rows, 4, dtype=torch.float32, device=stu.device)
        rowloss = torch.empty(rows, dtype=torch.float64, device=stu.device)
        _lib.call("dcb_row_softmax_stats", C.c_void_p(stu.data_ptr()), C.c_void_p(tea.data_ptr()), rows, cols, dtype_code(stu),
                  float(temperature or 1.0), mode, C.c_void_p(saved.data_ptr()), C.c_void_p(rowloss.data_ptr()), _stream_ptr())
        out = finalize([(rowloss, rows)], [float(temperature) ** 2 if mode == 0 else 1.0 / rows], [1.0])
        ctx.save_for_backward(stu, tea, saved)
        ctx.meta = (temperature, mode)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        stu, tea, saved = ctx.saved_tensors
        temperature, mode = ctx.meta
        up = g.to(torch.float32).reshape(1).contiguous()
        grad = torch.empty_like(stu)
        _lib.call("dcb_row_softmax_grads", C.c_void_p(stu.data_ptr()), C.c_void_p(tea.data_ptr()), stu.shape[0], stu.shape[1],
                  dtype_code(stu), float(temperature or 1.0), mode, C.c_void_p(saved.data_ptr()), C.c_void_p(up.data_ptr()),
                  C.c_void_p(grad.data_ptr()), _DT[grad.dtype], _stream_ptr())
        return grad, None, None, None


def row_softmax_loss(stu: torch.Tensor, tea: torch.Tensor, temperature, mode: int):
    """mode 0 = OutKLLoss (out_kl.py:12-16), mode 1 = OutCELoss (out_ce.py:9-13); [B, D] inputs, softmax over dim 1."""
    _require_cuda(stu, "student output")
    _require_cuda(tea, "teacher output")
    dtype_code(stu)
    if stu.dim() != 2 or stu.shape != tea.shape:
        raise ValueError(f"expected equal [B, D] tensors, got {tuple(stu.shape)} and {tuple(tea.shape)}")
    tea = tea.detach()
    if tea.dtype != stu.dtype:
        tea = tea.to(stu.dtype)
    return RowSoftmaxLossFn.apply(stu.contiguous(), tea.contiguous(), temperature, mode)


# ----------------------------------------------------------------------------------------------
# LastValueMapKL: KL over the head axis of [B, H, N, N] value-relation maps
# ----------------------------------------------------------------------------------------------
class ValueMapKLFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, stu, tea, expected, fwd_mult):
        grad = torch.empty_like(stu) if ctx.needs_input_grad[0] else None
        out = ValueMapKLFn._launch(stu, tea, grad, float(expected), fwd_mult, None, 0.0)
        ctx.grad, ctx.expected, ctx.fwd_mult = grad, expected, fwd_mult
        ctx.save_for_backward(stu, tea)
        return out

    @staticmethod
    def _launch(stu, tea, grad, grad_scale, fwd_mult, upstream, expected):
        b, h = stu.shape[0], stu.shape[1]
        vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        if upstream is None:
            partials = torch.empty(_lib.MAX_PARTIALS, dtype=torch.float64, device=stu.device)
            count = C.c_int(0)
            _lib.call("dcb_value_map_kl_fwd_bwd", vp(stu), vp(tea), vp(grad), b, h, stu.numel() // (b * h), dtype_code(stu),
                      dtype_code(stu), grad_scale, vp(partials), C.byref(count), vp(fwd_mult), None, 0.0, _stream_ptr())
            return finalize([(partials, count.value)], [1.0], [1.0])[0]
        _lib.call("dcb_value_map_kl_fwd_bwd", vp(stu), vp(tea), vp(grad), b, h, stu.numel() // (b * h), dtype_code(stu),
                  dtype_code(stu), grad_scale, None, None, vp(fwd_mult), vp(upstream), float(expected), _stream_ptr())
        return None

    @staticmethod
    def backward(ctx, g):
        stu, tea = ctx.saved_tensors
        if not ctx.needs_input_grad[0] or g is None:
            return None, None, None, None
        first = ctx.grad is not None
        grad, ctx.grad = ctx.grad, None                     # hand over (AccumulateGrad adopts it); later backwards recompute
        if not first:
            grad = torch.empty_like(stu)
        ValueMapKLFn._launch(stu, tea, grad, 1.0, ctx.fwd_mult, _as_upstream(g, stu), ctx.expected if first else _NAN)
        return grad, None, None, None


def value_map_kl(stu: torch.Tensor, tea: torch.Tensor, expected=None, fwd_mult=None):
    """LastValueMapKL (last_value_map_kl.py:10-14): softmax over dim=1 (heads) of both maps, KLDiv(sum)."""
    _require_cuda(stu, "student value map")
    _require_cuda(tea, "teacher value map")
    dtype_code(stu)
    if stu.dim() < 3 or stu.shape != tea.shape:
        raise ValueError(f"value maps must be equal [B, H, ...] tensors, got {tuple(stu.shape)} and {tuple(tea.shape)}")
    tea = tea.detach()
    if tea.dtype != stu.dtype:
        tea = tea.to(stu.dtype)
    return ValueMapKLFn.apply(stu.contiguous(), tea.contiguous(), float(EXPECTED_GRAD_SCALE if expected is None else expected), fwd_mult)
