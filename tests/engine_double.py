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

    # ------------------------------------------------------------------------------------------
    # pipeline flavour (distillclip_b200/pipeline.py): same stage boundaries and slot layout as the CUDA kernels
    # ------------------------------------------------------------------------------------------
    slot_dtype = stat_dtype = tr_dtype = torch.float64

    def gt_splits(self, rows, cols, dim, scatter=False):
        return 1

    def fwd_parts(self, rows, cols_chunk):
        return 1

    def prep(self, mats, invs, copies, trs):
        for m, inv, c, t in zip(mats, invs, copies, trs):
            r = 1.0 / m.double().norm(dim=1)
            inv.copy_(r)
            if c is not None:
                c.copy_(m)
            if t is not None:
                t.zero_()
                t[:, :m.shape[0]] = (m.double() * r[:, None]).t()

    def fwd_chunk(self, a_s, b_s, a_t, b_t, a_s_inv, b_s_inv, a_t_inv, b_t_inv, label_col0, temperature, ws_chunk, diag,
                  col_part_chunk, col_part_ld, ws_extra=None, diag_t=None):
        S = self._logits(a_s, b_s, a_s_inv, b_s_inv)
        rows, cols = S.shape
        ws_chunk.zero_()
        col_part_chunk.zero_()
        e1 = torch.exp(S - 1)
        ws_chunk[0, 0], col_part_chunk[0, 0] = e1.sum(1), e1.sum(0)
        lab = label_col0 + torch.arange(rows)
        hit = (lab >= 0) & (lab < cols)
        diag[hit] = S[torch.arange(rows)[hit], lab[hit]]
        if a_t is not None:
            T = self._logits(a_t, b_t, a_t_inv, b_t_inv)
            et, es = torch.exp((T - 1) / temperature), torch.exp((S - 1) / temperature)
            q = es - et + et * (T - S) / temperature
            ws_chunk[0, 1], col_part_chunk[0, 1] = q.sum(1), q.sum(0)
            ws_chunk[0, 2], col_part_chunk[0, 2] = et.sum(1), et.sum(0)
            ws_chunk[0, 3], col_part_chunk[0, 3] = (et * (T - S)).sum(1), (et * (T - S)).sum(0)
            if ws_extra is not None:
                ws_extra.zero_()
                ws_extra[0, 0], ws_extra[0, 1] = torch.relu(S - T).sum(1), ((S - T) ** 2).sum(1)
                diag_t[hit] = T[torch.arange(rows)[hit], lab[hit]]

    @staticmethod
    def _unit_coefs(st, global_batch, temperature, has_teacher):
        c = torch.zeros(3, st.shape[1], dtype=torch.float64)
        c[0] = 0.5 / (global_batch * st[0])
        if has_teacher:
            c[1] = 0.5 * temperature / (st[2] + st[1] - st[3] / temperature)
            c[2] = 0.5 * temperature / st[2]
        return c

    def post1(self, ws, diag, col_part, temperature, has_teacher, global_batch, stats, coef_row, dests, ws_extra=None, diag_t=None):
        from distillclip_b200.pipeline import slot_floats, slot_tail
        rows, cols = diag.shape[0], col_part.shape[2]
        stats[:4] = ws.sum(0)
        stats[4] = diag
        rl = self._rowloss(stats, temperature, has_teacher)
        coef_row.zero_()
        coef_row[:3] = self._unit_coefs(stats, global_batch, temperature, has_teacher)
        slot = torch.zeros(slot_floats(cols, rows), dtype=torch.float64)
        slot[:4 * cols] = col_part.sum(0).reshape(-1)
        slot[4 * cols:4 * cols + rows] = diag
        tail = slot_tail(cols, rows)
        slot[tail], slot[tail + 1] = rl[0].sum(), rl[1].sum()
        if ws_extra is not None:
            x = ws_extra.sum(0)
            slot[tail + 2] = torch.relu(diag_t - diag).sum()
            slot[tail + 3] = x[0].sum() - torch.relu(diag - diag_t).sum()
            slot[tail + 4] = x[1].sum()
            coef_row[3] = (diag_t > diag).double()
        slot[tail + 10:tail + 13] = coef_row[:3].max(1).values
        for d in dests:
            d.copy_(slot)

    def post2(self, slots, rows_per_src, cols, temperature, has_teacher, weights):
        from distillclip_b200.pipeline import slot_tail
        n_src = slots.shape[0]
        tail = slot_tail(cols, rows_per_src)
        st = torch.zeros(5, cols, dtype=torch.float64)
        st[:4] = slots[:, :4 * cols].reshape(n_src, 4, cols).sum(0)
        st[4] = slots[:, 4 * cols:4 * cols + rows_per_src].reshape(-1)
        rl = self._rowloss(st, temperature, has_teacher)
        coef_col = self._unit_coefs(st, cols, temperature, has_teacher)
        hard = 0.5 * (slots[:, tail].sum() + rl[0].sum()) / cols
        soft = 0.5 * (temperature or 1.0) ** 2 * (slots[:, tail + 1].sum() + rl[1].sum())
        cosd = slots[:, tail + 2].sum() / cols + (slots[:, tail + 3].sum() / (cols * (cols - 1.0)) if cols > 1 else 0.0)
        lmse = slots[:, tail + 4].sum() / (float(cols) * cols)
        p_h, p_s, s_h, s_s, p_c, p_m, s_c, s_m = weights
        out = torch.stack([hard, soft, hard * s_h, soft * s_s,
                           hard * s_h * p_h + soft * s_s * p_s + cosd * s_c * p_c + lmse * s_m * p_m,
                           cosd, lmse, cosd * s_c, lmse * s_m])
        bounds = torch.cat([slots[:, tail + 10:tail + 13].max(0).values, coef_col.max(1).values])
        return st[:4].clone(), coef_col, bounds, out

    @staticmethod
    def _ups(up):
        g5, w8 = up
        f = lambda g: 0.0 if g is None else float(g)
        gt = f(g5[0])
        return (gt * w8[0] + f(g5[1]) * w8[2], gt * w8[1] + f(g5[2]) * w8[3], gt * w8[4] + f(g5[3]) * w8[6], gt * w8[5] + f(g5[4]) * w8[7])

    @staticmethod
    def _scale2(bounds, ups, batch, extra):
        up_h, up_s, up_c, up_m = ups
        gmax = abs(up_h) * (bounds[0] + bounds[3]) + abs(up_s) * (bounds[1] + bounds[2] + bounds[4] + bounds[5])
        if extra:
            gmax = gmax + abs(up_c) / (batch * (batch - 1.0)) + abs(up_m) * 4.0 / (batch * batch)
        gmax = torch.as_tensor(gmax, dtype=torch.float64)
        return 2.0 ** (14 - torch.frexp(gmax)[1].item()) if float(gmax) > 0 else 1.0

    #: True: pipeline.backward_gemms takes the split flow (g_tiles -> text-side GEMM -> image-side GEMM) with this double too
    split = False

    def use_split(self, rows, cols):
        return self.split

    def rg_splits(self, rows, cols, dim):
        return 1

    def g_tiles(self, a_s, b_s, a_t, b_t, a_s_inv, b_s_inv, a_t_inv, b_t_inv, coef_row, coef_col, bounds, up, temperature,
                g_out, extra=False, row_offset=0):
        """The recompute kernel of the split backward: only the scaled gradient tiles."""
        dummy_bt = torch.zeros(1, a_s.shape[1], b_s.shape[0], dtype=torch.float64)
        self.pair_bwd(a_s, b_s, a_t, b_t, dummy_bt, a_s_inv, b_s_inv, a_t_inv, b_t_inv, coef_row, coef_col, bounds, up,
                      temperature, g_out, extra=extra, row_offset=row_offset)

    def row_acc_from_g(self, g, bt_all, rows, cols, dim, out=None):
        n = cols // bt_all.shape[0]
        b_hat_t = torch.cat([bt_all[r][:, :n] for r in range(bt_all.shape[0])], dim=1)        # [D, B]
        acc = (g[:rows, :cols].double() @ b_hat_t.t())[None]                                      # 2^k sum_j G_ij b_hat_j
        if out is not None:
            out.copy_(acc)
            return out
        return acc

    def pair_bwd(self, a_s, b_s, a_t, b_t, bt_all, a_s_inv, b_s_inv, a_t_inv, b_t_inv, coef_row, coef_col, bounds, up,
                 temperature, g_out, extra=False, row_offset=0):
        ups = self._ups(up)
        up_h, up_s, up_c, up_m = ups
        S = self._logits(a_s, b_s, a_s_inv, b_s_inv)
        batch = S.shape[1]
        G = torch.exp(S - 1) * up_h * (coef_row[0][:, None] + coef_col[0][None, :])
        if a_t is not None:
            T = self._logits(a_t, b_t, a_t_inv, b_t_inv)
            G = G + torch.exp((S - 1) / temperature) * up_s * (coef_row[1][:, None] + coef_col[1][None, :])
            G = G - torch.exp((T - 1) / temperature) * up_s * (coef_row[2][:, None] + coef_col[2][None, :])
            if extra:
                off = torch.ones_like(S, dtype=torch.bool)
                idx = torch.arange(S.shape[0])
                off[idx, row_offset + idx] = False
                G = G + up_c / (batch * (batch - 1.0)) * ((S > T) & off).double() + 2.0 * up_m / (batch * batch) * (S - T)
        scale = self._scale2(bounds, ups, batch, extra)
        assert float((G.abs() * scale).max()) <= 2.0 ** 14
        if g_out is not None:
            g_out[:, :G.shape[1]] = G * scale
        n = b_s.shape[0] // bt_all.shape[0]
        b_hat_t = torch.cat([bt_all[r][:, :n] for r in range(bt_all.shape[0])], dim=1)        # [D, B]
        return ((G * scale) @ b_hat_t.t())[None]

    def finish2(self, side_a, side_b, global_batch, up, bounds, grad_dtype, cos_flag=None):
        ups = self._ups(up)
        up_h, up_s, up_c, up_m = ups
        scale = self._scale2(bounds, ups, global_batch, cos_flag is not None)
        out = []
        for sd in (side_a, side_b):
            if sd is None:
                out.append(None)
                continue
            x, x_inv = sd["x"].double(), sd["x_inv"]
            acc = sd["acc"].sum(0) / scale
            gi = sd["label_offset"] + torch.arange(x.shape[0])
            lab = up_h + (up_c * cos_flag if cos_flag is not None else 0.0)
            acc = acc - (lab / global_batch)[:, None] * sd["y_inv"][gi][:, None] * sd["y"].double()[gi] if cos_flag is not None else \
                acc - (up_h / global_batch) * sd["y_inv"][gi][:, None] * sd["y"].double()[gi]
            x_hat = x * x_inv[:, None]
            out.append(x_inv[:, None] * (acc - x_hat * (x_hat * acc).sum(1, keepdim=True)))
        return out[0], out[1]
