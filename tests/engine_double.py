"""CPU test double of `distillclip_b200.contrastive.CudaEngine` (TEST INFRASTRUCTURE ONLY).

It restates, in float64 torch on the CPU, exactly the decomposition the CUDA kernels implement (shift-1 softmax
statistics, loss from statistics, gradient coefficients, G tile, label term, normalisation Jacobian), so that
(a) the formulas can be checked against the oracle without a GPU and (b) the sharding / collective logic of
`contrastive.contrastive_forward/backward` can run under gloo with world_size 2.  The product never imports it."""
import torch


class DoubleEngine:
    def inv_norms(self, mats):
        return [1.0 / m.double().norm(dim=1) for m in mats]

    def _logits(self, a, b, a_inv, b_inv):
        return (a.double() @ b.double().t()) * a_inv[:, None] * b_inv[None, :]

    def row_stats(self, a_s, b_s, a_t, b_t, a_s_inv, b_s_inv, a_t_inv, b_t_inv, row_offset, temperature, dump=None,
                  with_cols=False):
        S = self._logits(a_s, b_s, a_s_inv, b_s_inv)
        rows, cols = S.shape
        idx = torch.arange(rows)
        st = torch.zeros(5, rows, dtype=torch.float64)
        col = torch.zeros(4, cols, dtype=torch.float64)
        e1 = torch.exp(S - 1)
        st[0], col[0] = e1.sum(1), e1.sum(0)
        st[4] = S[idx, row_offset + idx]
        if a_t is not None:
            T = self._logits(a_t, b_t, a_t_inv, b_t_inv)
            et, es = torch.exp((T - 1) / temperature), torch.exp((S - 1) / temperature)
            q = es - et + et * (T - S) / temperature          # second-order part of Zs - Zt, as the kernel carries it
            st[1], col[1] = q.sum(1), q.sum(0)
            st[2], col[2] = et.sum(1), et.sum(0)
            st[3], col[3] = (et * (T - S)).sum(1), (et * (T - S)).sum(0)
        rl = self._rowloss(st, temperature, a_t is not None)
        return (st, rl, col) if with_cols else (st, rl)

    @staticmethod
    def _rowloss(st, temperature, has_teacher):
        rl = torch.zeros(2, st.shape[1], dtype=torch.float64)
        rl[0] = 1 + torch.log(st[0]) - st[4]
        if has_teacher:
            m = -st[3] / (temperature * st[2])
            rl[1] = -m + torch.log1p(m + st[1] / st[2])
        return rl

    def rank_counts(self, a, b, a_inv, b_inv, row_offset, ref):
        return (self._logits(a, b, a_inv, b_inv) > ref[:, None]).sum(1).double()

    def col_finish(self, col_stats, diag_local, row_offset, temperature, has_teacher):
        rows = diag_local.shape[0]
        st = torch.zeros(5, rows, dtype=torch.float64)
        st[:4] = col_stats[:, row_offset:row_offset + rows]
        st[4] = diag_local
        return st, self._rowloss(st, temperature, has_teacher)

    def losses(self, rl_i2t, rl_t2i, global_batch, temperature, has_teacher):
        sums = torch.zeros(4, dtype=torch.float64)
        for d, rl in enumerate((rl_i2t, rl_t2i)):
            sums[d] = rl[0].sum()
            if has_teacher:
                sums[2 + d] = temperature ** 2 * rl[1].sum()
        out = torch.stack([0.5 * (sums[0] + sums[1]) / global_batch, 0.5 * (sums[2] + sums[3])])
        return sums, out

    def coef(self, stats, global_batch, temperature, has_teacher, upstream):
        gh, gs = upstream.double()
        c = torch.zeros(3, stats.shape[1], dtype=torch.float64)
        c[0] = 0.5 * gh / (global_batch * stats[0])
        if has_teacher:
            c[1] = 0.5 * gs * temperature / (stats[2] + stats[1] - stats[3] / temperature)       # Zs = Zt + Q - W/T
            c[2] = 0.5 * gs * temperature / stats[2]
        return c, c.abs().sum(0).max().reshape(1)

    def transpose_norm(self, b, b_inv):
        return (b.double() * b_inv[:, None]).t().contiguous()

    single_pass_backward = True

    def single_pass_supported(self, dim):
        return self.single_pass_backward

    def alloc_g(self, rows, cols, device):
        return torch.zeros(rows, (cols + 7) // 8 * 8, dtype=torch.float64)

    @staticmethod
    def _scale(gmax_row, gmax_col):
        return 2.0 ** (14 - torch.frexp(gmax_row + gmax_col)[1].item())     # the fp16 tile scale of the kernel

    def row_acc(self, a_s, b_s, a_t, b_t, b_s_t, a_s_inv, b_s_inv, a_t_inv, b_t_inv, coef_row, coef_col,
                gmax_row, gmax_col, temperature, g_out=None):
        S = self._logits(a_s, b_s, a_s_inv, b_s_inv)
        G = torch.exp(S - 1) * (coef_row[0][:, None] + coef_col[0][None, :])
        if a_t is not None:
            T = self._logits(a_t, b_t, a_t_inv, b_t_inv)
            G = G + torch.exp((S - 1) / temperature) * (coef_row[1][:, None] + coef_col[1][None, :])
            G = G - torch.exp((T - 1) / temperature) * (coef_row[2][:, None] + coef_col[2][None, :])
        scale = self._scale(gmax_row, gmax_col)
        assert float((G.abs() * scale).max()) <= 2.0 ** 14
        if g_out is not None:
            g_out[:, :G.shape[1]] = G * scale
        return ((G * scale) @ b_s_t.t())[None]                       # 2^k sum_j G_ij b_hat_j, one split

    def col_acc_from_g(self, g, a_hat_t, rows, cols, dim):
        return (g[:rows, :cols].t() @ a_hat_t.t()[:rows])[None]      # 2^k sum_i G_ij a_hat_i

    def finish_grads(self, acc, a_s, a_s_inv, b_s, b_s_inv, gmax_row, gmax_col, row_offset, global_batch, upstream,
                     grad_dtype):
        acc = acc.sum(0) / self._scale(gmax_row, gmax_col)
        rows = a_s.shape[0]
        gi = row_offset + torch.arange(rows)
        acc = acc - (upstream[0].double() / global_batch) * b_s_inv[gi][:, None] * b_s.double()[gi]
        a_hat = a_s.double() * a_s_inv[:, None]
        return a_s_inv[:, None] * (acc - a_hat * (a_hat * acc).sum(1, keepdim=True))

    def row_grads(self, a_s, b_s, a_t, b_t, b_s_t, a_s_inv, b_s_inv, a_t_inv, b_t_inv, coef_row, coef_col,
                  gmax_row, gmax_col, row_offset, global_batch, temperature, upstream, grad_dtype, g_out=None):
        acc = self.row_acc(a_s, b_s, a_t, b_t, b_s_t, a_s_inv, b_s_inv, a_t_inv, b_t_inv, coef_row, coef_col,
                           gmax_row, gmax_col, temperature, g_out)
        return self.finish_grads(acc, a_s, a_s_inv, b_s, b_s_inv, gmax_row, gmax_col, row_offset, global_batch,
                                 upstream, grad_dtype)
