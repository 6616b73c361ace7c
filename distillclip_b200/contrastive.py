"""Host side of the contrastive (InfoNCE) + teacher/student logit-KL path.

Two entry levels, both backed only by the CUDA library (no eager fallback):

* `clip_contrastive(...)` -- the fused path used by `LossCalculator.cal_tow_tower_loss`: losses and gradients
  straight from the embeddings (`last_representation`), the B x B logits never exist in HBM.  It replaces
  `CLIPModel.forward`'s normalise + matmul (reference model/component/clip_model.py:36-44), `HardLabel`
  (model/loss_component/hard_label.py:10-12), `SoftLabel` (soft_label.py:11-16) and the 0.5*(i2t+t2i) sums of
  model/_loss.py:130-137.  With a process group it is row-sharded: every rank owns B/R rows of each embedding
  matrix, all-gathers the opposite side, and computes its row slice of the global logits (labels are the global
  row indices) and the gradients of its own rows; see DESIGN.md section "multi-GPU".

* `hard_label_from_logits / soft_label_from_logits` -- the per-module API on materialised logits (including the
  `.T` view), for callers that use `HardLabel` / `SoftLabel` directly.

The sharding / collective logic is written against a small "engine" interface so that it can be exercised on CPU
(gloo, world_size 2) with a test double; the product engine is `CudaEngine` below.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import torch

from . import _lib, ops

MIN_FUSED_TEMPERATURE = 0.025       # exp((S-1)/T) with S in [-1, 1] stays a normal fp32 number above this
_HALF = (torch.bfloat16, torch.float16)


def _vp(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else None


# ==============================================================================================
# engine: one call per kernel family of the C ABI
# ==============================================================================================
class CudaEngine:
    """Thin wrapper over the C ABI (include/distillclip_b200.h)."""
    #: backward kernel choice: CTA-pair kernel for D <= 768 (default) or the single-CTA D-chunked kernel
    use_pair_kernel = os.environ.get("DCB_BWD_KERNEL", "pair") != "chunk"
    #: one recompute of the logits for BOTH sides' gradients (G tiles stored in fp16, second side = one GEMM); "0" = one
    #: pair-kernel pass per side (no O(B^2) scratch)
    single_pass_backward = os.environ.get("DCB_BWD_SINGLE_PASS", "1") != "0"
    #: tests only: [rows, cols] fp32 buffer that receives the logits the pair kernel's epilogue sees
    dump_pair_logits = None
    #: profiling only: int64 [2, 64, 16] buffer for clock64() stamps of the pair kernel's first cluster
    trace_pair = None

    def inv_norms(self, mats: Sequence[torch.Tensor]):
        outs = [torch.empty(m.shape[0], dtype=torch.float32, device=m.device) for m in mats]
        _lib.call("dcb_row_inv_norm", len(mats), _lib.ptr_array([m.data_ptr() for m in mats]),
                  _lib.ptr_array([o.data_ptr() for o in outs]), _lib.i64_array([m.shape[0] for m in mats]),
                  mats[0].shape[1], ops.dtype_code(mats[0]), ops._stream_ptr())
        return outs

    def row_stats(self, a_s, b_s, a_t, b_t, a_s_inv, b_s_inv, a_t_inv, b_t_inv, row_offset, temperature,
                  dump=None, with_cols=False):
        """One pass over the logit tiles S = a b^T (and T): row statistics [5, rows] + per-row losses [2, rows] of the
        a-side rows and, with_cols=True, the column sums [4, cols] of the same terms over these rows (this call's
        contribution to the row statistics of the opposite direction)."""
        rows, dim = a_s.shape
        cols = b_s.shape[0]
        stats = torch.empty(5, rows, dtype=torch.float32, device=a_s.device)
        rowloss = torch.empty(2, rows, dtype=torch.float64, device=a_s.device)
        col_stats = torch.empty(4, cols, dtype=torch.float32, device=a_s.device) if with_cols else None
        ws = torch.empty(max(1, _lib.load().dcb_clip_workspace_bytes(rows, cols)), dtype=torch.uint8, device=a_s.device)
        _lib.call("dcb_clip_row_stats", _vp(a_s), _vp(b_s), _vp(a_t), _vp(b_t), _vp(a_s_inv), _vp(b_s_inv),
                  _vp(a_t_inv), _vp(b_t_inv), rows, int(row_offset), cols, dim, ops.dtype_code(a_s),
                  float(temperature or 1.0), _vp(stats), _vp(rowloss), _vp(col_stats), _vp(ws),
                  _vp(dump[0]) if dump else None, _vp(dump[1]) if dump and a_t is not None else None, ops._stream_ptr())
        if not with_cols:
            _lib.LAUNCHES -= 1          # no column-reduce launch
            return stats, rowloss
        return stats, rowloss, col_stats

    def rank_counts(self, a, b, a_inv, b_inv, row_offset, ref):
        """counts[i] = #{j : S_ij > ref[i]} for S = normalised a b^T (exact integers in fp32): the retrieval rank of the
        label when ref = S_ii of row_stats (validation top-k accuracy, distillclip_b200/metrics.py)."""
        rows, dim = a.shape
        cols = b.shape[0]
        stats = torch.empty(5, rows, dtype=torch.float32, device=a.device)
        rowloss = torch.empty(2, rows, dtype=torch.float64, device=a.device)
        ws = torch.empty(max(1, _lib.load().dcb_clip_workspace_bytes(rows, cols)), dtype=torch.uint8, device=a.device)
        _lib.call("dcb_clip_rank_counts", _vp(a), _vp(b), _vp(a_inv), _vp(b_inv), rows, int(row_offset), cols, dim,
                  ops.dtype_code(a), _vp(ref), _vp(stats), _vp(rowloss), _vp(ws), ops._stream_ptr())
        return stats[1]

    def col_finish(self, col_stats, diag_local, row_offset, temperature, has_teacher):
        """Opposite-direction statistics [5, rows_local] and per-row losses of this rank's rows from complete column sums."""
        rows = diag_local.shape[0]
        stats = torch.empty(5, rows, dtype=torch.float32, device=col_stats.device)
        rowloss = torch.empty(2, rows, dtype=torch.float64, device=col_stats.device)
        _lib.call("dcb_clip_col_finish", _vp(col_stats), col_stats.shape[1], _vp(diag_local), int(row_offset), rows,
                  float(temperature or 1.0), int(has_teacher), _vp(stats), _vp(rowloss), ops._stream_ptr())
        return stats, rowloss

    def losses(self, rowloss_i2t, rowloss_t2i, global_batch, temperature, has_teacher):
        dev = rowloss_i2t.device
        sums = torch.empty(4, dtype=torch.float64, device=dev)
        out = torch.empty(2, dtype=torch.float32, device=dev)
        _lib.call("dcb_clip_losses", _vp(rowloss_i2t), _vp(rowloss_t2i), rowloss_i2t.shape[1], rowloss_t2i.shape[1],
                  int(global_batch), float(temperature or 1.0), int(has_teacher), _vp(sums), _vp(out), ops._stream_ptr())
        return sums, out

    def coef(self, stats, global_batch, temperature, has_teacher, upstream):
        """-> (coef [3, rows], gmax [1]): gradient coefficients and the bound max_i sum_k |coef_k,i|."""
        rows = stats.shape[1]
        coef = torch.empty(3, rows, dtype=torch.float32, device=stats.device)
        gmax = torch.empty(1, dtype=torch.float32, device=stats.device)
        _lib.call("dcb_clip_grad_coef", _vp(stats), rows, int(global_batch), float(temperature or 1.0),
                  int(has_teacher), _vp(upstream), _vp(coef), _vp(gmax), ops._stream_ptr())
        return coef, gmax

    def transpose_norm(self, b, b_inv):
        """fp16 [D, pitch] = (b * b_inv[:, None]).T, the K-major b-side operand of the gradient GEMM."""
        rows, dim = b.shape
        pitch = (rows + 7) // 8 * 8
        out = torch.empty(dim, pitch, dtype=torch.float16, device=b.device)
        _lib.call("dcb_transpose_norm_f16", _vp(b), _vp(b_inv), _vp(out), rows, dim, pitch, ops.dtype_code(b),
                  ops._stream_ptr())
        return out

    def single_pass_supported(self, dim: int) -> bool:
        """Single-recompute backward: the pair kernel stores its fp16 gradient tiles and the other side's gradient is
        one GEMM over them (csrc/clip_bwd_gt.cu)."""
        return bool(self.single_pass_backward and self.use_pair_kernel and _lib.load().dcb_clip_pair_supported(dim))

    def alloc_g(self, rows: int, cols: int, device):
        return torch.empty(rows, (cols + 7) // 8 * 8, dtype=torch.float16, device=device)

    def row_acc(self, a_s, b_s, a_t, b_t, b_s_t, a_s_inv, b_s_inv, a_t_inv, b_t_inv, coef_row, coef_col,
                gmax_row, gmax_col, temperature, g_out=None):
        """acc[s, i, :] = 2^k sum_{j in split s} G_ij b_hat_j (fp32 partial buffers); optionally G 2^k -> g_out (fp16)."""
        rows, dim = a_s.shape
        cols = b_s.shape[0]
        lib = _lib.load()
        if self.use_pair_kernel and lib.dcb_clip_pair_supported(dim):
            # CTA-pair kernel (cta_group::2): whole embedding dimension in one pass, logits recomputed once
            n_split = lib.dcb_clip_pair_splits(rows, cols, dim)
            acc = torch.empty(n_split, rows, dim, dtype=torch.float32, device=a_s.device)
            _lib.call("dcb_clip_row_grads_pair", _vp(a_s), _vp(b_s), _vp(a_t), _vp(b_t), _vp(b_s_t), b_s_t.shape[1],
                      _vp(a_s_inv), _vp(b_s_inv), _vp(a_t_inv), _vp(b_t_inv), _vp(coef_row), _vp(coef_col),
                      _vp(gmax_row), _vp(gmax_col), rows, cols, dim, ops.dtype_code(a_s), float(temperature or 1.0),
                      _vp(acc), _vp(g_out), g_out.shape[1] if g_out is not None else 0,
                      _vp(self.dump_pair_logits), _vp(self.trace_pair), ops._stream_ptr())
        else:
            if g_out is not None:
                raise _lib.DistillClipB200Error("the gradient tiles are only stored by the CTA-pair kernel")
            n_split = lib.dcb_clip_grad_splits(rows, cols, dim)
            acc = torch.empty(n_split, rows, dim, dtype=torch.float32, device=a_s.device)
            _lib.call("dcb_clip_row_grads", _vp(a_s), _vp(b_s), _vp(a_t), _vp(b_t), _vp(b_s_t), b_s_t.shape[1],
                      _vp(a_s_inv), _vp(b_s_inv), _vp(a_t_inv), _vp(b_t_inv), _vp(coef_row), _vp(coef_col),
                      _vp(gmax_row), _vp(gmax_col), rows, cols, dim, ops.dtype_code(a_s), float(temperature or 1.0),
                      _vp(acc), ops._stream_ptr())
        return acc

    def col_acc_from_g(self, g, a_hat_t, rows: int, cols: int, dim: int):
        """acc[s, j, :] = sum_{i in split s} G[i, j] 2^k a_hat[i, :]  -- tcgen05 GEMM with A = G^T read MN-major."""
        n_split = _lib.load().dcb_clip_gt_splits(rows, cols, dim)
        acc = torch.empty(n_split, cols, dim, dtype=torch.float32, device=g.device)
        _lib.call("dcb_clip_col_grads_from_g", _vp(g), g.shape[1], _vp(a_hat_t), a_hat_t.shape[1], rows, cols, dim,
                  _vp(acc), ops._stream_ptr())
        return acc

    def col_acc_scatter(self, g, a_hat_t, rows: int, cols: int, dim: int, dest_ptrs, src_slot: int):
        """The G^T GEMM with its epilogue storing every output row into the partial buffer of the rank that owns it
        (peer-mapped pointers `dest_ptrs`, one per rank): the reduce-scatter of the row-sharded backward without a collective."""
        _lib.call("dcb_clip_col_grads_scatter", _vp(g), g.shape[1], _vp(a_hat_t), a_hat_t.shape[1], rows, cols, dim,
                  _lib.ptr_array([d if isinstance(d, int) else d.data_ptr() for d in dest_ptrs]), len(dest_ptrs), int(src_slot),
                  ops._stream_ptr())

    def finish_grads(self, acc, a_s, a_s_inv, b_s, b_s_inv, gmax_row, gmax_col, row_offset, global_batch, upstream,
                     grad_dtype):
        """2^-k sum_s acc[s] - label term, then the x/||x|| Jacobian (dcb_clip_grad_finish)."""
        rows, dim = a_s.shape
        grad = torch.empty(rows, dim, dtype=grad_dtype, device=a_s.device)
        _lib.call("dcb_clip_grad_finish", _vp(acc), acc.shape[0], _vp(a_s), _vp(a_s_inv), _vp(b_s), _vp(b_s_inv), rows,
                  b_s.shape[0], dim, int(row_offset), int(global_batch), _vp(upstream), _vp(gmax_row), _vp(gmax_col),
                  ops.dtype_code(a_s), _vp(grad), ops._DT[grad_dtype], ops._stream_ptr())
        return grad

    def row_grads(self, a_s, b_s, a_t, b_t, b_s_t, a_s_inv, b_s_inv, a_t_inv, b_t_inv, coef_row, coef_col,
                  gmax_row, gmax_col, row_offset, global_batch, temperature, upstream, grad_dtype, g_out=None):
        acc = self.row_acc(a_s, b_s, a_t, b_t, b_s_t, a_s_inv, b_s_inv, a_t_inv, b_t_inv, coef_row, coef_col,
                           gmax_row, gmax_col, temperature, g_out)
        return self.finish_grads(acc, a_s, a_s_inv, b_s, b_s_inv, gmax_row, gmax_col, row_offset, global_batch,
                                 upstream, grad_dtype)

    # ------------------------------------------------------------------------------------------
    # pipeline flavour (distillclip_b200/pipeline.py): fused small kernels, unit coefficients, exchange-friendly outputs
    # ------------------------------------------------------------------------------------------
    slot_dtype = stat_dtype = torch.float32
    tr_dtype = torch.float16
    _scratch: Dict = {}

    def rg_splits(self, rows, cols, dim):
        return _lib.load().dcb_clip_rg_splits(rows, cols, dim)

    def gt_splits(self, rows, cols, dim, scatter=False):
        lib = _lib.load()
        return lib.dcb_clip_gt_splits_scatter(rows, cols, dim) if scatter else lib.dcb_clip_gt_splits(rows, cols, dim)

    def fwd_parts(self, rows, cols_chunk):
        return _lib.load().dcb_clip_fwd_chunk_parts(rows, cols_chunk)

    @classmethod
    def _post_scratch(cls, rows, cols, dev, which):
        """Zero-initialised ticket + block partials of the post kernels (they reset the ticket): one per device, stream, kernel."""
        need = _lib.load().dcb_clip_post_scratch_bytes(rows, cols)
        key = (dev.index, torch.cuda.current_stream(dev).cuda_stream, which)
        buf = cls._scratch.get(key)
        if buf is None or buf.numel() < need:
            buf = cls._scratch[key] = torch.zeros(max(need, 4096), dtype=torch.uint8, device=dev)
        return buf

    def prep(self, mats, invs, copies, trs):
        rows, dim = mats[0].shape
        _lib.call("dcb_clip_prep", len(mats), _lib.ptr_array([m.data_ptr() for m in mats]),
                  _lib.ptr_array([o.data_ptr() for o in invs]),
                  _lib.ptr_array([c.data_ptr() if c is not None else 0 for c in copies]),
                  _lib.ptr_array([t.data_ptr() if t is not None else 0 for t in trs]),
                  _lib.i64_array([t.stride(0) if t is not None else 0 for t in trs]), rows, dim, ops.dtype_code(mats[0]),
                  ops._stream_ptr())

    def fwd_chunk(self, a_s, b_s, a_t, b_t, a_s_inv, b_s_inv, a_t_inv, b_t_inv, label_col0, temperature, ws_chunk, diag,
                  col_part_chunk, col_part_ld, ws_extra=None, diag_t=None):
        _lib.call("dcb_clip_fwd_chunk", _vp(a_s), _vp(b_s), _vp(a_t), _vp(b_t), _vp(a_s_inv), _vp(b_s_inv), _vp(a_t_inv),
                  _vp(b_t_inv), a_s.shape[0], int(label_col0), b_s.shape[0], a_s.shape[1], ops.dtype_code(a_s),
                  float(temperature or 1.0), _vp(ws_chunk), _vp(diag), _vp(col_part_chunk), int(col_part_ld), _vp(ws_extra),
                  _vp(diag_t), ops._stream_ptr())

    def post1(self, ws, diag, col_part, temperature, has_teacher, global_batch, stats, coef_row, dests, ws_extra=None, diag_t=None):
        rows, cols = diag.shape[0], col_part.shape[2]
        _lib.call("dcb_clip_post1", _vp(ws), ws.shape[0], _vp(diag), _vp(col_part), col_part.shape[0], rows, cols,
                  float(temperature or 1.0), int(has_teacher), int(global_batch), _vp(stats), _vp(coef_row),
                  _lib.ptr_array([d.data_ptr() for d in dests]), len(dests), _vp(ws_extra), _vp(diag_t),
                  _vp(self._post_scratch(rows, cols, diag.device, 1)), ops._stream_ptr())

    def post2(self, slots, rows_per_src, cols, temperature, has_teacher, weights):
        dev = slots.device
        col_stats = torch.empty(4, cols, dtype=torch.float32, device=dev)
        coef_col = torch.empty(3, cols, dtype=torch.float32, device=dev)
        bounds = torch.empty(6, dtype=torch.float32, device=dev)
        out = torch.empty(9, dtype=torch.float32, device=dev)
        _lib.call("dcb_clip_post2", _vp(slots), slots.shape[0], int(rows_per_src), int(cols), float(temperature or 1.0),
                  int(has_teacher), int(cols), _lib.f32_array(weights), _vp(col_stats), _vp(coef_col), _vp(bounds), _vp(out),
                  _vp(self._post_scratch(rows_per_src, cols, dev, 2)), ops._stream_ptr())
        return col_stats, coef_col, bounds, out

    @staticmethod
    def _up_args(up):
        """up = (g5 tensors-or-None, w8 floats) -> ctypes arrays"""
        g5, w8 = up
        return _lib.ptr_array([g.data_ptr() if g is not None else 0 for g in g5]), _lib.f32_array(w8)

    def pair_bwd(self, a_s, b_s, a_t, b_t, bt_all, a_s_inv, b_s_inv, a_t_inv, b_t_inv, coef_row, coef_col, bounds, up,
                 temperature, g_out, extra=False, row_offset=0):
        rows, dim = a_s.shape
        cols = b_s.shape[0]
        n_split = _lib.load().dcb_clip_pair_splits(rows, cols, dim)
        acc = torch.empty(n_split, rows, dim, dtype=torch.float32, device=a_s.device)
        g5, w8 = self._up_args(up)
        _lib.call("dcb_clip_pair_bwd", _vp(a_s), _vp(b_s), _vp(a_t), _vp(b_t), _vp(bt_all), bt_all.stride(1), cols // bt_all.shape[0],
                  _vp(a_s_inv), _vp(b_s_inv), _vp(a_t_inv), _vp(b_t_inv), _vp(coef_row), _vp(coef_col), _vp(bounds),
                  g5, w8, int(bool(extra)), int(row_offset), cols, rows, cols, dim,
                  ops.dtype_code(a_s), float(temperature or 1.0), _vp(acc), _vp(g_out), g_out.shape[1] if g_out is not None else 0,
                  ops._stream_ptr())
        return acc

    #: split backward (csrc/clip_bwd_g.cu): recompute -> fp16 G tiles, then one tcgen05 GEMM per tower over the stored tiles,
    #: instead of the pair kernel that fuses the image-side GEMM into the recompute.  "auto": when the local block of the
    #: logit matrix has at least `split_min_tiles` 128 x 128 tiles (measured at D = 512: 4096 x 4096 4 % faster, 2048 x 4096 8 %, 1024 x 4096 4 %;
    #: below that the fused kernel saves a launch)
    split_backward = os.environ.get("DCB_BWD_SPLIT", "auto")
    split_min_tiles = int(os.environ.get("DCB_BWD_SPLIT_MIN_TILES", "256"))

    def split_supported(self, dim: int) -> bool:
        """The split flow has no TMEM-resident gradient accumulator, so it also covers 768 < D <= 1024 (the finish kernel's row
        registers end there); it shares the stored-tiles switch with the pair flow."""
        return bool(self.single_pass_backward and self.use_pair_kernel and dim % 8 == 0 and 8 <= dim <= 1024)

    def use_split(self, rows: int, cols: int) -> bool:
        if self.split_backward in ("0", "1"):
            return self.split_backward == "1"
        return ((rows + 127) // 128) * ((cols + 127) // 128) >= self.split_min_tiles

    def g_tiles(self, a_s, b_s, a_t, b_t, a_s_inv, b_s_inv, a_t_inv, b_t_inv, coef_row, coef_col, bounds, up, temperature,
                g_out, extra=False, row_offset=0):
        rows, dim = a_s.shape
        cols = b_s.shape[0]
        g5, w8 = self._up_args(up)
        _lib.call("dcb_clip_g_tiles", _vp(a_s), _vp(b_s), _vp(a_t), _vp(b_t), _vp(a_s_inv), _vp(b_s_inv), _vp(a_t_inv), _vp(b_t_inv),
                  _vp(coef_row), _vp(coef_col), _vp(bounds), g5, w8, int(bool(extra)), int(row_offset), cols, rows, cols, dim,
                  ops.dtype_code(a_s), float(temperature or 1.0), _vp(g_out), g_out.shape[1], ops._stream_ptr())

    def row_acc_from_g(self, g, bt_all, rows: int, cols: int, dim: int, out=None):
        """acc[s, i, :] = sum_{j in split s} G[i, j] 2^k b_hat[j, :]  -- tcgen05 GEMM with A = G read K-major.  `out`: the
        accumulator allocated by the caller (on the stream that will consume it) when this launch goes to a side stream."""
        n_split = _lib.load().dcb_clip_rg_splits(rows, cols, dim)
        acc = out if out is not None else torch.empty(n_split, rows, dim, dtype=torch.float32, device=g.device)
        assert acc.shape == (n_split, rows, dim)
        _lib.call("dcb_clip_row_grads_from_g", _vp(g), g.shape[1], _vp(bt_all), bt_all.stride(1), cols // bt_all.shape[0],
                  rows, cols, dim, _vp(acc), ops._stream_ptr())
        return acc

    def finish2(self, side_a, side_b, global_batch, up, bounds, grad_dtype, cos_flag=None):
        """side = dict(acc [n_split, rows, D], x [rows, D], x_inv, y (label rows of the other tower), y_inv, label_offset) or None."""
        args, grads, ref = [], [], side_a or side_b
        dim = ref["x"].shape[1]
        for sd in (side_a, side_b):
            if sd is None:
                args += [None, 0, 0, None, None, None, 0, None, None, 0, 0]
                grads.append(None)
                continue
            x, acc = sd["x"], sd["acc"]
            g = torch.empty(x.shape, dtype=grad_dtype, device=x.device)
            grads.append(g)
            args += [_vp(acc), acc.shape[0], acc.stride(0), _vp(x), _vp(sd["x_inv"]), _vp(g), x.shape[0], _vp(sd["y"]),
                     _vp(sd["y_inv"]), sd["y"].shape[0], int(sd["label_offset"])]
        g5, w8 = self._up_args(up)
        _lib.call("dcb_clip_finish2", *args, dim, int(global_batch), g5, w8, _vp(cos_flag), _vp(bounds),
                  ops.dtype_code(ref["x"]), ops._DT[grad_dtype], ops._stream_ptr())
        return grads[0], grads[1]


# ==============================================================================================
# sharding logic (engine-agnostic; collectives through torch.distributed)
# ==============================================================================================
def _shard_info(group):
    """-> (rank, world). group=None means "not sharded" even if torch.distributed is initialised."""
    if group is None:
        return 0, 1
    import torch.distributed as dist
    return dist.get_rank(group), dist.get_world_size(group)


def _all_gather_rows(x: torch.Tensor, group, world: int) -> torch.Tensor:
    """[n, ...] per rank -> [world * n, ...] in rank order (equal n on every rank)."""
    if world == 1:
        return x
    import torch.distributed as dist
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def _reduce_scatter_rows(x: torch.Tensor, group, world: int, rank: int) -> torch.Tensor:
    """[k, world * n, d] partial sums on every rank -> [k, n, d]: this rank's rows summed over ranks."""
    if world == 1:
        return x
    import torch.distributed as dist
    n = x.shape[1] // world
    if x.shape[0] > 1:
        x = x.sum(0, keepdim=True)               # K-split partials: one collective of [B, D] instead of one per split
    if dist.get_backend(group) == "nccl":
        out = torch.empty(x.shape[0], n, x.shape[2], dtype=x.dtype, device=x.device)
        for k in range(x.shape[0]):
            dist.reduce_scatter_tensor(out[k], x[k], group=group)
        return out
    dist.all_reduce(x, group=group)                  # gloo (CPU tests) has no reduce_scatter
    return x[:, rank * n:(rank + 1) * n].contiguous()


class PeerScatter:
    """Symmetric (peer-mapped) partial buffers for the fused G^T GEMM + reduce-scatter: rank r owns
    [world * k_split][B/R][D] fp32; every rank's GEMM epilogue stores the rows r owns into slot block `src rank` of r's
    buffer over NVLink.  One stream-ordered cross-rank barrier after the GEMM tells the owner that all stores have landed;
    the buffers alternate between two copies, so a rank can only overwrite partials its owner read two steps ago -- and the
    barrier of the step in between already ordered every rank behind that read.  torch symmetric memory supplies the mapping
    and the barrier; set DCB_PEER_SCATTER=0 (or `enabled = False`) to use the NCCL reduce-scatter instead."""
    enabled = os.environ.get("DCB_PEER_SCATTER", "1") != "0"
    _cache: Dict = {}
    _broken = False
    _flip: Dict = {}

    @classmethod
    def get(cls, group, world: int, slots: int, rows: int, dim: int, device):
        """-> (handle, local buffer [slots, rows, dim], peer pointers) or None when symmetric memory is unavailable."""
        if not cls.enabled or cls._broken or world > 16:
            return None
        import torch.distributed as dist
        if dist.get_backend(group) != "nccl":
            return None
        base = (group.group_name, slots, rows, dim, device.index)
        cls._flip[base] = cls._flip.get(base, 0) ^ 1        # alternate per shape
        key = base + (cls._flip[base],)
        if key not in cls._cache:
            try:
                import torch.distributed._symmetric_memory as symm
                buf = symm.empty(slots * rows * dim, dtype=torch.float32, device=device)
                hdl = symm.rendezvous(buf, group.group_name)
                cls._cache[key] = (hdl, buf.view(slots, rows, dim), [int(p) for p in hdl.buffer_ptrs])
            except Exception as exc:          # no P2P mapping on this machine: fall back for the rest of the process
                cls._broken = True
                import warnings
                warnings.warn(f"distillclip_b200: symmetric memory unavailable ({exc}); using the NCCL reduce-scatter")
                return None
        return cls._cache[key]


def _all_gather_cols(x: torch.Tensor, group, world: int) -> torch.Tensor:
    """[k, n] per rank -> [k, world * n]."""
    if world == 1:
        return x
    g = _all_gather_rows(x.t().contiguous(), group, world)
    return g.t().contiguous()


def contrastive_forward(engine, si, st, ti, tt, temperature, group=None):
    """Row statistics + loss values.  si/st/ti/tt: this rank's rows [B_local, D] (tt/ti None = hard label only).
    Returns (out[2] = {hard, soft} for the GLOBAL batch, saved state for the backward)."""
    rank, world = _shard_info(group)
    has_teacher = ti is not None
    b_local = si.shape[0]
    b_global = b_local * world
    offset = rank * b_local
    # exchange step: every rank needs all TEXT rows (student and teacher).  The image rows stay local: the t2i direction
    # comes from column sums / the G^T GEMM over this rank's own image rows (only the two-pass backward gathers them)
    st_all = _all_gather_rows(st, group, world)
    tt_all = _all_gather_rows(tt, group, world) if has_teacher else None
    inv = engine.inv_norms([si, st_all] + ([ti, tt_all] if has_teacher else []))
    si_inv, st_inv_all = inv[0], inv[1]
    ti_inv, tt_inv_all = (inv[2], inv[3]) if has_teacher else (None, None)
    # ONE pass over the logits of this rank's image rows against all text rows: row sums = i2t statistics of the local
    # rows, column sums = this rank's share of the t2i statistics of ALL text rows (t2i logits are the transpose)
    stats_i2t, rl_i2t, col = engine.row_stats(si, st_all, ti, tt_all, si_inv, st_inv_all, ti_inv, tt_inv_all, offset,
                                              temperature, with_cols=True)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(col, group=group)                 # exchange step 2: complete the column sums
    stats_t2i, rl_t2i = engine.col_finish(col, stats_i2t[4], offset, temperature, has_teacher)
    sums, out = engine.losses(rl_i2t, rl_t2i, b_global, temperature, has_teacher)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(sums, group=group)
        out = torch.stack([0.5 * (sums[0] + sums[1]) / b_global, 0.5 * (sums[2] + sums[3])]).to(torch.float32)
    saved = dict(si=si, st=st, ti=ti, tt=tt, st_all=st_all, tt_all=tt_all,
                 si_inv=si_inv, st_inv_all=st_inv_all, ti_inv=ti_inv, tt_inv_all=tt_inv_all,
                 stats_i2t=stats_i2t, stats_t2i=stats_t2i, col_stats=col, offset=offset, b_global=b_global, world=world,
                 group=group, rank=rank, temperature=temperature, has_teacher=has_teacher)
    return out, saved


def contrastive_backward(engine, saved, upstream, want_img=True, want_txt=True, grad_dtype=None):
    """upstream: device float32[2] = {d total/d hard, d total/d soft}.  Returns (grad_si, grad_st) for the local rows."""
    s = saved
    world, group = s["world"], s["group"]
    b_global, T, has_teacher, offset = s["b_global"], s["temperature"], s["has_teacher"], s["offset"]
    b_local = s["si"].shape[0]
    loc = slice(offset, offset + b_local)
    # column softmax statistics of a direction = row statistics of the opposite direction, for ALL rows: the t2i ones
    # are the (already complete) column sums of the forward pass, the i2t ones are gathered
    stats_i2t_all = _all_gather_cols(s["stats_i2t"], group, world)
    coef_i2t_all, gmax_i2t = engine.coef(stats_i2t_all, b_global, T, has_teacher, upstream)
    coef_t2i_all, gmax_t2i = engine.coef(s["col_stats"], b_global, T, has_teacher, upstream)
    coef_i2t = coef_i2t_all[:, loc].contiguous() if world > 1 else coef_i2t_all
    coef_t2i = coef_t2i_all[:, loc].contiguous() if world > 1 else coef_t2i_all

    def local(x):
        return None if x is None else x[loc]
    g_img = g_txt = None
    dim = s["si"].shape[1]
    if want_img and want_txt and engine.single_pass_supported(dim):
        # one recompute: the image-side pass also stores G 2^k [local image rows, all text columns]; the text-side
        # accumulator is G^T a_hat over the LOCAL image rows, summed across ranks (reduce-scatter to the local text rows);
        # its label term pairs local text row i with local image row i
        g_tiles = engine.alloc_g(b_local, b_global, s["si"].device)
        g_img = engine.row_grads(s["si"], s["st_all"], s["ti"], s["tt_all"],
                                 engine.transpose_norm(s["st_all"], s["st_inv_all"]),
                                 s["si_inv"], s["st_inv_all"], s["ti_inv"], s["tt_inv_all"],
                                 coef_i2t, coef_t2i_all, gmax_i2t, gmax_t2i, offset, b_global, T, upstream,
                                 grad_dtype or s["si"].dtype, g_out=g_tiles)
        a_hat_t = engine.transpose_norm(s["si"], s["si_inv"])
        peer = None
        if world > 1 and hasattr(engine, "col_acc_scatter"):
            k_split = _lib.load().dcb_clip_gt_splits(b_local, b_global, dim)
            peer = PeerScatter.get(group, world, world * k_split, b_local, dim, s["si"].device)
        if peer is not None:
            # GEMM -> reduce-scatter in ONE kernel: the epilogue stores each text row's partial into its owner's buffer
            hdl, acc, ptrs = peer
            engine.col_acc_scatter(g_tiles, a_hat_t, b_local, b_global, dim, ptrs, s["rank"])
            hdl.barrier(channel=0)              # every rank's stores have landed
        else:
            acc = engine.col_acc_from_g(g_tiles, a_hat_t, b_local, b_global, dim)
            acc = _reduce_scatter_rows(acc, group, world, s["rank"])
        g_txt = engine.finish_grads(acc, s["st"], local(s["st_inv_all"]), s["si"], s["si_inv"], gmax_t2i, gmax_i2t,
                                    0, b_global, upstream, grad_dtype or s["st"].dtype)
        return g_img, g_txt
    if want_img:
        g_img = engine.row_grads(s["si"], s["st_all"], s["ti"], s["tt_all"],
                                 engine.transpose_norm(s["st_all"], s["st_inv_all"]),
                                 s["si_inv"], s["st_inv_all"], s["ti_inv"], s["tt_inv_all"],
                                 coef_i2t, coef_t2i_all, gmax_i2t, gmax_t2i, offset, b_global, T, upstream, grad_dtype or s["si"].dtype)
    if want_txt:
        # two-pass route: a second recompute over this rank's text rows against ALL image rows (gathered only here)
        si_all = _all_gather_rows(s["si"], group, world)
        ti_all = _all_gather_rows(s["ti"], group, world) if has_teacher else None
        si_inv_all = _all_gather_rows(s["si_inv"], group, world)
        ti_inv_all = _all_gather_rows(s["ti_inv"], group, world) if has_teacher else None
        g_txt = engine.row_grads(s["st"], si_all, s["tt"], ti_all,
                                 engine.transpose_norm(si_all, si_inv_all),
                                 local(s["st_inv_all"]), si_inv_all, local(s["tt_inv_all"]), ti_inv_all,
                                 coef_t2i, coef_i2t_all, gmax_t2i, gmax_i2t, offset, b_global, T, upstream, grad_dtype or s["st"].dtype)
    return g_img, g_txt


# ==============================================================================================
# autograd
# ==============================================================================================
_ENGINE = CudaEngine()


class ClipContrastiveFn(torch.autograd.Function):
    """(stu_img, stu_txt, tea_img, tea_txt) -> (hard, soft): 0.5*(i2t + t2i) of CrossEntropy(mean) and of
    T^2 KL(sum) (reference _loss.py:130-137), global batch when `group` is given."""

    @staticmethod
    def forward(ctx, si, st, ti, tt, temperature, group):
        out, saved = contrastive_forward(_ENGINE, si, st, ti, tt, temperature, group)
        ctx.saved = saved
        ctx.set_materialize_grads(False)
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_hard, g_soft):
        dev = ctx.saved["si"].device
        zero = torch.zeros((), dtype=torch.float32, device=dev)
        up = torch.stack([(g_hard if g_hard is not None else zero).to(torch.float32).reshape(()),
                          (g_soft if g_soft is not None else zero).to(torch.float32).reshape(())]).contiguous()
        g_img, g_txt = contrastive_backward(_ENGINE, ctx.saved, up, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return g_img, g_txt, None, None, None, None


def _as_scalar(g: Optional[torch.Tensor]):
    if g is None:
        return None
    if g.dtype != torch.float32 or not g.is_contiguous():
        g = g.to(torch.float32).contiguous()
    return g


class ClipPipelineFn(torch.autograd.Function):
    """(stu_img, stu_txt, tea_img, tea_txt) -> (hard s_hard, soft s_soft, cos_diff s_cos, logits_mse s_mse, total): the four
    logit losses of reference _loss.py:130-145 -- 0.5*(i2t + t2i) of CrossEntropy(mean), T^2 KL(sum), CLIPCosDiff and
    LogitsMSE -- with the scale / percent weighting of _loss.py:231-234 (total = sum percent * scaled value), for the global
    batch of the exchange `xc` (distillclip_b200/pipeline.py)."""

    @staticmethod
    def forward(ctx, si, st, ti, tt, temperature, xc, weights, extra):
        from . import pipeline
        out, saved = pipeline.pipeline_forward(_ENGINE, xc, si, st, ti, tt, temperature, weights, extra)
        ctx.set_materialize_grads(False)
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            ctx.saved = saved
        else:
            xc.release(saved["set"])                   # nothing will come back for these buffers
        return out[2], out[3], out[7], out[8], out[4]

    @staticmethod
    def backward(ctx, g_hard, g_soft, g_cos, g_mse, g_total):
        from . import pipeline
        ups = tuple(_as_scalar(g) for g in (g_total, g_hard, g_soft, g_cos, g_mse))
        g_img, g_txt = pipeline.pipeline_backward(_ENGINE, ctx.saved, ups, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return g_img, g_txt, None, None, None, None, None, None


def fused_supported(stu_img: torch.Tensor, stu_txt: torch.Tensor, temperature=None, tea_img=None, tea_txt=None) -> bool:
    """The tcgen05 path takes bf16/fp16 [B, D] embeddings with D % 8 == 0, T >= MIN_FUSED_TEMPERATURE and (for the soft
    label) teacher embeddings of the student's shape."""
    ok = (stu_img is not None and stu_txt is not None and stu_img.is_cuda and stu_img.dim() == 2
          and stu_img.shape == stu_txt.shape and stu_img.dtype in _HALF
          and stu_txt.dtype == stu_img.dtype and stu_img.shape[1] % 8 == 0)
    if temperature is not None:
        ok = ok and float(temperature) >= MIN_FUSED_TEMPERATURE
    for t in (tea_img, tea_txt):
        if t is not None:
            ok = ok and t.is_cuda and t.shape == stu_img.shape
    return bool(ok)


def _prep(x: Optional[torch.Tensor], dtype, what: str):
    if x is None:
        return None
    ops._require_cuda(x, what)
    if x.dtype != dtype:
        x = x.to(dtype)
    if not x.is_contiguous() or x.data_ptr() % 16:
        x = x.contiguous() if not x.is_contiguous() else x.clone()
    return x


def clip_contrastive(stu_img, stu_txt, tea_img=None, tea_txt=None, temperature=None, want_hard=True,
                     want_soft=False, group=None, percent=None, scale=None, want_cos_diff=False,
                     want_logits_mse=False) -> Dict[str, torch.Tensor]:
    """Fused logit losses from embeddings.  Returns {'hard_label', 'soft_label', 'cos_diff', 'logits_mse'} (only the requested
    keys; multiplied by `scale` when given) and, with `percent`, also 'total' = sum percent * scaled value computed on the
    device.  `percent` / `scale`: 2-tuples (hard, soft) or 4-tuples (hard, soft, cos_diff, logits_mse).  Values are 0-dim fp32
    tensors on the autograd graph of the student embeddings.  With a process group the batch is the GLOBAL batch (rows of
    all ranks).  cos_diff / logits_mse (reference clip_cos_diff.py:5-23, logits_mse.py:9-10) need the teacher embeddings."""
    from . import pipeline
    extra = bool(want_cos_diff or want_logits_mse)
    need_teacher = bool(want_soft or extra)
    if not fused_supported(stu_img, stu_txt, temperature if want_soft else None):
        raise _lib.DistillClipB200Error(
            "fused contrastive path needs CUDA bf16/fp16 [B, D] embeddings with D % 8 == 0 and "
            f"temperature >= {MIN_FUSED_TEMPERATURE}; use HardLabel / SoftLabel on logits otherwise")
    dt = stu_img.dtype
    si, st = _prep(stu_img, dt, "student image embedding"), _prep(stu_txt, dt, "student text embedding")
    ti = tt = None
    if need_teacher:
        if tea_img is None or tea_txt is None:
            raise ValueError("soft_label / cos_diff / logits_mse need the teacher embeddings")
        ti, tt = _prep(tea_img.detach(), dt, "teacher image embedding"), _prep(tea_txt.detach(), dt, "teacher text embedding")
        if ti.shape != si.shape or tt.shape != st.shape:
            raise ValueError("teacher and student embeddings must have the same [B, D] shape on the fused path "
                             f"(student {tuple(si.shape)}, teacher {tuple(ti.shape)})")
    sc = [float(x) for x in scale] if scale is not None else [1.0, 1.0]
    pc = [float(x) for x in percent] if percent is not None else [0.0, 0.0]
    sc, pc = sc + [1.0] * (4 - len(sc)), pc + [0.0] * (4 - len(pc))
    # the soft-label statistics need a temperature; when only cos_diff / logits_mse use the teacher any valid one will do
    T = float(temperature) if (want_soft or (need_teacher and temperature)) else (1.0 if need_teacher else None)
    xc = pipeline.exchange_for(group)
    if xc.world > 1:
        pipeline.check_equal_batches(group, si.shape[0], si.device)
    res = {}
    if USE_PIPELINE and pipeline.pipeline_supported(_ENGINE, xc, si.shape[0], si.shape[1]):
        weights = (pc[0], pc[1], sc[0], sc[1], pc[2], pc[3], sc[2], sc[3])
        hard, soft, cosd, lmse, total = ClipPipelineFn.apply(si, st, ti, tt, T, xc, weights, extra)
    else:
        if extra:
            raise _lib.DistillClipB200Error("cos_diff / logits_mse from embeddings need the pipeline path (D <= 1024, D % 8 == 0, "
                                            "per-rank batch a multiple of 128 when sharded); use the logits modules otherwise")
        hard, soft = ClipContrastiveFn.apply(si, st, ti, tt, T, group)
        hard, soft = hard * sc[0], soft * sc[1]
        total = hard * pc[0] + soft * pc[1]
        cosd = lmse = None
    if want_hard:
        res["hard_label"] = hard
    if want_soft:
        res["soft_label"] = soft
    if want_cos_diff:
        res["cos_diff"] = cosd
    if want_logits_mse:
        res["logits_mse"] = lmse
    if percent is not None:
        res["total"] = total
    return res


def extras_supported(stu_img, group=None) -> bool:
    """cos_diff / logits_mse from embeddings ride on the pipeline kernels only."""
    from . import pipeline
    if not USE_PIPELINE or stu_img is None or stu_img.dim() != 2:
        return False
    return pipeline.pipeline_supported(_ENGINE, pipeline.exchange_for(group), stu_img.shape[0], stu_img.shape[1])


#: False (or DCB_PIPELINE=0): the round-1 flow (one launch per quantity, NCCL collectives) for every shape
USE_PIPELINE = os.environ.get("DCB_PIPELINE", "1") != "0"


# ==============================================================================================
# per-module API on materialised logits
# ==============================================================================================
class LogitsLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, stu, tea, temperature, mode):
        n = stu.shape[0]
        dev = stu.device
        saved = torch.empty(n, 4, dtype=torch.float32, device=dev)
        rowloss = torch.empty(n, dtype=torch.float64, device=dev)
        _lib.call("dcb_logits_row_stats", _vp(stu), stu.stride(0), stu.stride(1), _vp(tea),
                  tea.stride(0) if tea is not None else 0, tea.stride(1) if tea is not None else 0, n,
                  ops.dtype_code(stu), float(temperature or 1.0), mode, _vp(saved), _vp(rowloss), ops._stream_ptr())
        scale = float(temperature) ** 2 if mode == 1 else (1.0 / n if mode == 0 else 1.0)
        out = ops.finalize([(rowloss, n)], [scale], [1.0])
        ctx.save_for_backward(stu, tea if tea is not None else stu, saved)
        ctx.meta = (temperature, mode, tea is not None)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        stu, tea, saved = ctx.saved_tensors
        temperature, mode, has_tea = ctx.meta
        n = stu.shape[0]
        up = g.to(torch.float32).reshape(1).contiguous()
        grad = torch.empty(n, n, dtype=stu.dtype, device=stu.device)
        _lib.call("dcb_logits_row_grads", _vp(stu), stu.stride(0), stu.stride(1), _vp(tea) if has_tea else None,
                  tea.stride(0) if has_tea else 0, tea.stride(1) if has_tea else 0, n, ops.dtype_code(stu),
                  float(temperature or 1.0), mode, _vp(saved), _vp(up), _vp(grad), ops._DT[grad.dtype],
                  ops._stream_ptr())
        return grad, None, None, None


def _check_logits(x: torch.Tensor, what: str):
    ops._require_cuda(x, what)
    ops.dtype_code(x)
    if x.dim() != 2 or x.shape[0] != x.shape[1]:
        raise ValueError(f"{what} must be a square [B, B] matrix, got {tuple(x.shape)}")


def hard_label_from_logits(stu_logits: torch.Tensor) -> torch.Tensor:
    """CrossEntropy(mean)(logits, arange(B)) -- reference hard_label.py:10-12.  Strided views are read in place."""
    _check_logits(stu_logits, "stu_logits")
    return LogitsLossFn.apply(stu_logits, None, None, 0)


def soft_label_from_logits(stu_logits: torch.Tensor, tea_logits: torch.Tensor, temperature) -> torch.Tensor:
    """KLDiv(sum)(softmax(stu/T).log(), softmax(tea/T)) * T^2 -- reference soft_label.py:11-16."""
    _check_logits(stu_logits, "stu_logits")
    _check_logits(tea_logits, "tea_logits")
    if tea_logits.shape != stu_logits.shape:
        raise ValueError("student and teacher logits must have the same shape")
    if tea_logits.dtype != stu_logits.dtype:
        tea_logits = tea_logits.to(stu_logits.dtype)
    return LogitsLossFn.apply(stu_logits, tea_logits.detach(), float(temperature), 1)


def cos_diff_from_logits(stu_logits: torch.Tensor, tea_logits: torch.Tensor) -> torch.Tensor:
    """mean relu(tea_ii - stu_ii) + mean_{i != j} relu(stu_ij - tea_ij) -- reference clip_cos_diff.py:12-23."""
    _check_logits(stu_logits, "stu_logits")
    _check_logits(tea_logits, "tea_logits")
    if tea_logits.shape != stu_logits.shape:
        raise ValueError("student and teacher logits must have the same shape")
    if tea_logits.dtype != stu_logits.dtype:
        tea_logits = tea_logits.to(stu_logits.dtype)
    return LogitsLossFn.apply(stu_logits, tea_logits.detach(), None, 2)
