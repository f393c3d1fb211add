"""Vocab-parallel logits for the multi-GPU path: one process per GPU, batch data-parallel
encoder, the full-catalogue GEMM split by item rows across ranks.

  training  : all-gather `out` -> every rank runs the fused tcgen05 CE-partial kernel on ITS table rows for
              ALL rows -> all-gather of the per-row (max, sum-exp) pairs -> local combine (log-sum-exp);
              backward: shard-local G^T, d_out partial = G.E_shard, reduce-scatter(sum) of d_out,
              dE_shard = G^T.out written into the owner's rows of the gradient buffer (no traffic).
  evaluation: all-gather `out` -> fused top-k on the shard (indices offset by the shard start, column 0
              skipped on shard 0) -> all-gather of the partial lists -> merge of the rank's own rows.

Round-1 storage note: the table itself is still replicated (the K1 gather reads local rows); only the
logits/CE/top-k work and its HBM traffic are sharded.  Sharded storage + peer-memory gather is the next step.

The collective plumbing is backend-agnostic: `compute` supplies the five local kernels (CUDA ops by default;
the gloo CPU tests inject torch restatements), so world_size-2 tests run without a GPU.
"""
import torch
import torch.distributed as dist


class CudaCompute(object):
    """the local kernels, bound to libacsr.so through ops.py"""

    def __init__(self, passes=3):
        from . import ops
        self.ops, self.passes = ops, passes

    def num_chunks(self, rows, shard_rows):
        return self.ops.logits_num_chunks(rows, shard_rows) if shard_rows > 0 else 1

    def ce_partial(self, out, table):
        return self.ops.ce_partial(out, table, self.passes)

    def ce_grad_t(self, out, table, lse, target_local, row_scale):
        return self.ops.ce_grad_matrix_t(out, table, lse, target_local, row_scale, self.passes)

    def gt_times_table(self, Gt, table):
        return self.ops.linear_wgrad(Gt, table, want_bias=False)[0]

    def topk_partial(self, out, table, k, idx_offset, skip_col0):
        return self.ops.logits_topk_partial(out, table, k, idx_offset, skip_col0, self.passes)

    def topk_merge(self, pv, pi, k, positive):
        return self.ops.topk_merge(pv, pi, k, positive)


class VocabParallel(object):
    def __init__(self, n_items, group=None, compute=None, align=64):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        per = (n_items + self.world - 1) // self.world
        per = (per + align - 1) // align * align                 # whole 64-row tiles per shard
        self.V, self.per = n_items, per
        self.lo = min(n_items, self.rank * per)
        self.hi = min(n_items, self.lo + per)
        self.compute = compute or CudaCompute()
        self.shard_rows = [max(0, min(n_items, (r + 1) * per) - min(n_items, r * per)) for r in range(self.world)]

    def _ncmax(self, rows):
        """largest per-shard chunk count, computed on the host from the shard sizes (no sync, graph-capturable)"""
        return max(self.compute.num_chunks(rows, sr) for sr in self.shard_rows)

    # ------------------------------------------------------------------------------------------
    def _all_gather(self, t):
        t = t.contiguous()
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return out.view((self.world,) + tuple(t.shape))

    def shard(self, table):
        return table[self.lo:self.hi]

    # ------------------------------------------------------------------------------------------
    def ce_forward(self, out_local, table, target_local, n_groups):
        """-> (loss [n_groups] over this rank's rows, saved state for ce_backward)."""
        R, d = out_local.shape
        out_all = self._all_gather(out_local).view(self.world * R, d)
        tgt_all = self._all_gather(target_local).view(self.world * R)
        E_s = self.shard(table)
        if E_s.shape[0] > 0:
            part = self.compute.ce_partial(out_all, E_s)                     # [W*R, nc, 2]
        else:
            part = torch.tensor([-float('inf'), 0.0], dtype=out_local.dtype, device=out_local.device).repeat(self.world * R, 1, 1)
        ncmax = self._ncmax(self.world * R)
        if part.shape[1] < ncmax:                                            # ragged chunk counts: pad with (-inf, 0)
            pad = part.new_zeros((part.shape[0], ncmax - part.shape[1], 2))
            pad[..., 0] = -float('inf')
            part = torch.cat((part, pad), 1)
        mine = self._all_gather(part)[:, self.rank * R:(self.rank + 1) * R]   # [W, R, ncmax, 2]: my rows on every shard
        mine = mine.permute(1, 0, 2, 3).reshape(R, self.world * ncmax, 2)
        m = mine[..., 0].max(dim=1).values
        s = (mine[..., 1] * torch.exp(mine[..., 0] - m.unsqueeze(1))).sum(1)
        lse = m + torch.log(s)
        tgt_logit = (out_local * table[target_local]).sum(1)                 # the replica holds every row
        row_loss = lse - tgt_logit
        loss = row_loss.view(n_groups, R // n_groups).mean(1)
        return loss, dict(out_all=out_all, tgt_all=tgt_all, lse=lse, R=R)

    def ce_backward(self, st, table, row_scale_local, table_grad=None, table_half=None, n_groups=2):
        """row_scale_local [R]: d loss / d row_loss of this rank's rows.  -> d_out_local [R,d]; when table_grad is
        given, dE of the owned rows is accumulated into table_grad[lo:hi] from the row group `table_half`
        (of n_groups equal groups per rank) that trains the table -- the calibrated rows."""
        R = st['R']
        lse_all = self._all_gather(st['lse']).view(-1)
        scale_all = self._all_gather(row_scale_local).view(-1)
        E_s = self.shard(table)
        d = table.shape[1]
        if E_s.shape[0] > 0:
            Gt = self.compute.ce_grad_t(st['out_all'], E_s, lse_all, (st['tgt_all'] - self.lo).contiguous(), scale_all)
            d_out_all = self.compute.gt_times_table(Gt, E_s)                 # [W*R, d] partial over my item rows
            if table_grad is not None:
                W, per = self.world, R // n_groups
                Gs = Gt.view(Gt.shape[0], W, n_groups, per)[:, :, table_half].reshape(Gt.shape[0], W * per)
                os_ = st['out_all'].view(W, n_groups, per, d)[:, table_half].reshape(W * per, d)
                table_grad[self.lo:self.hi].addmm_(Gs, os_)
        else:
            d_out_all = torch.zeros((self.world * R, d), dtype=table.dtype, device=table.device)
        if dist.get_backend(self.group) == 'nccl':
            d_out = torch.empty((R, d), dtype=table.dtype, device=table.device)
            dist.reduce_scatter_tensor(d_out, d_out_all.contiguous(), op=dist.ReduceOp.SUM, group=self.group)
            return d_out
        dist.all_reduce(d_out_all, group=self.group)                         # gloo (CPU tests): no reduce-scatter
        return d_out_all[self.rank * R:(self.rank + 1) * R].clone()

    # ------------------------------------------------------------------------------------------
    def full_sort_topk(self, out_local, table, k, positive_local=None):
        """all-gather of partial top-k lists (north_star eval path) -> (val [R,k], idx [R,k], rec|None)."""
        R, d = out_local.shape
        out_all = self._all_gather(out_local).view(self.world * R, d)
        E_s = self.shard(table)
        if E_s.shape[0] > 0:
            pv, pi = self.compute.topk_partial(out_all, E_s, k, self.lo, self.lo == 0)   # [W*R, nc, k]
        else:
            pv = torch.full((self.world * R, 1, k), -float('inf'), dtype=out_local.dtype, device=out_local.device)
            pi = torch.full((self.world * R, 1, k), -1, dtype=torch.int64, device=out_local.device)
        ncmax = self._ncmax(self.world * R)
        if pv.shape[1] < ncmax:
            padn = ncmax - pv.shape[1]
            pv = torch.cat((pv, pv.new_full((pv.shape[0], padn, k), -float('inf'))), 1)
            pi = torch.cat((pi, pi.new_full((pi.shape[0], padn, k), -1)), 1)
        sl = slice(self.rank * R, (self.rank + 1) * R)
        mv = self._all_gather(pv)[:, sl].permute(1, 0, 2, 3).reshape(R, self.world * ncmax, k).contiguous()
        mi = self._all_gather(pi)[:, sl].permute(1, 0, 2, 3).reshape(R, self.world * ncmax, k).contiguous()
        return self.compute.topk_merge(mv, mi, k, positive_local)
