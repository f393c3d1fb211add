"""world_size-2 gloo tests (CPU) of the vocab-parallel collective plumbing (ac-tsr_b200/dist.py): with a torch
restatement of the five local kernels injected, the sharded CE forward/backward and the sharded top-k must equal
the unsharded computation on the concatenated batch."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class TorchCompute(object):
    def num_chunks(self, rows, shard_rows):
        return 1

    def ce_partial(self, out, table):
        logits = out @ table.t()
        m = logits.max(1).values
        return torch.stack((m, torch.exp(logits - m[:, None]).sum(1)), -1).unsqueeze(1)

    def ce_backward_local(self, out_all, table, lse, target_local, row_scale, row_begin, n_rows, table_grad):
        G = torch.exp(out_all @ table.t() - lse[:, None])
        ok = (target_local >= 0) & (target_local < table.shape[0])
        rows = torch.nonzero(ok).view(-1)
        G[rows, target_local[rows]] -= 1.0
        G = G * row_scale[:, None]
        if table_grad is not None:
            table_grad += G[row_begin:row_begin + n_rows].t() @ out_all[row_begin:row_begin + n_rows]
        return G @ table

    def gather_rows(self, ids, shard, lo, hi):
        own = (ids >= lo) & (ids < hi)
        return shard[(ids - lo).clamp(0, shard.shape[0] - 1)] * own.to(shard.dtype)[:, None]

    def scatter_add_rows(self, ids, rows, lo, hi, shard_grad):
        own = (ids >= lo) & (ids < hi) & (ids != 0)
        shard_grad.index_add_(0, (ids - lo)[own], rows[own])

    def topk_partial(self, out, table, k, idx_offset, skip_col0):
        s = out @ table.t()
        if skip_col0:
            s[:, 0] = -float('inf')
        kk = min(k, s.shape[1])
        v, i = torch.topk(s, kk, dim=1)
        pv = torch.full((s.shape[0], 1, k), -float('inf'), dtype=s.dtype)
        pi = torch.full((s.shape[0], 1, k), -1, dtype=torch.int64)
        pv[:, 0, :kk], pi[:, 0, :kk] = v, i + idx_offset
        pi[pv == -float('inf')] = -1
        return pv, pi

    def topk_merge(self, pv, pi, k, positive):
        R = pv.shape[0]
        v, order = torch.sort(pv.reshape(R, -1), dim=1, descending=True, stable=True)
        idx = torch.gather(pi.reshape(R, -1), 1, order)[:, :k]
        rec = None
        if positive is not None:
            rec = torch.cat(((idx == positive.view(-1, 1)).int(), torch.ones(R, 1, dtype=torch.int32)), 1)
        return v[:, :k], idx, rec


def _worker(rank, world, port, V, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import ac_tsr_b200 as A
        from ac_tsr_b200 import dist as D        # noqa: F401
        torch.manual_seed(0)
        d, B, k = 16, 6, 10
        E = torch.randn(V, d, dtype=torch.float64) * 0.3
        out_all = torch.randn(world, 2 * B, d, dtype=torch.float64)
        tgt_all = torch.randint(0, V, (world, 2 * B))
        vp = A.dist.VocabParallel(V, compute=TorchCompute(), align=8)
        out, tgt = out_all[rank].clone(), tgt_all[rank].clone()
        loss, st = vp.ce_forward(out, E, tgt, 2)
        # reference: plain CE on this rank's rows
        logits = out @ E.t()
        rl = torch.logsumexp(logits, 1) - logits[torch.arange(2 * B), tgt]
        want = rl.view(2, B).mean(1)
        assert torch.allclose(loss, want, atol=1e-10), (loss, want)
        # backward: loss_total = mean over ranks of (cal - att): row scale +1/B on the second half, -1/B on the first
        scale = torch.cat((torch.full((B,), -1.0 / B), torch.full((B,), 1.0 / B))).double()
        gE = torch.zeros_like(E)
        d_out = vp.ce_backward(st, E, scale, gE, table_half=1, n_groups=2)
        o = out.clone().requires_grad_(True)
        lg = o @ E.t()
        r2 = torch.logsumexp(lg, 1) - lg[torch.arange(2 * B), tgt]
        (r2 * scale).sum().backward()
        assert torch.allclose(d_out, o.grad, atol=1e-10)
        # table gradient: owner rows hold the sum over ALL ranks' calibrated rows
        Eg = E.clone().requires_grad_(True)
        tot = 0
        for r in range(world):
            lg = out_all[r, B:] @ Eg.t()
            tot = tot + ((torch.logsumexp(lg, 1) - lg[torch.arange(B), tgt_all[r, B:]]) / B).sum()
        tot.backward()
        assert torch.allclose(gE[vp.lo:vp.hi], Eg.grad[vp.lo:vp.hi], atol=1e-10)
        assert float(gE[:vp.lo].abs().sum() + gE[vp.hi:].abs().sum()) == 0.0
        # eval: sharded top-k == top-k of the full scores with column 0 removed
        pos = tgt[:B]
        val, idx, rec = vp.full_sort_topk(out[:B], E, k, pos)
        s = out[:B] @ E.t()
        s[:, 0] = -float('inf')
        rv, ri = torch.topk(s, k, dim=1)
        assert torch.equal(idx, ri) and torch.allclose(val, rv)
        assert torch.equal(rec[:, :-1].bool(), ri == pos.view(-1, 1))
        # ---- sharded STORAGE: every rank holds only its rows; same losses / gradients / top-k, plus the row exchange ----
        vs = A.dist.VocabParallel(V, compute=TorchCompute(), align=8, sharded=True)
        shard = vs.make_shard(E)
        assert shard.shape[0] == vs.per and torch.equal(vs.gather_full(shard), E)
        loss_s, st_s = vs.ce_forward(out, shard, tgt, 2)
        assert torch.allclose(loss_s, want, atol=1e-10)
        g_shard = torch.zeros_like(shard)
        d_out_s = vs.ce_backward(st_s, shard, scale, g_shard, table_half=1, n_groups=2)
        assert torch.allclose(d_out_s, o.grad, atol=1e-10)
        assert torch.allclose(g_shard[:vs.hi - vs.lo], Eg.grad[vs.lo:vs.hi], atol=1e-10)
        val_s, idx_s, rec_s = vs.full_sort_topk(out[:B], shard, k, pos)
        assert torch.equal(idx_s, ri) and torch.allclose(val_s, rv)
        # embedding rows of my tokens come from their owners; gradient rows go back to them (id 0 = padding gets none)
        T = 11
        ids_all_ref = torch.randint(0, V, (world, T))
        ids_all_ref[:, 0] = 0
        rows, ids_all = vs.fetch_rows(ids_all_ref[rank], shard)
        assert torch.equal(ids_all, ids_all_ref) and torch.equal(rows, E[ids_all_ref[rank]])
        d_rows_all = torch.randn(world, T, d, dtype=torch.float64)
        g2 = torch.zeros_like(shard)
        vs.scatter_grad_rows(ids_all, d_rows_all[rank], g2)
        full = torch.zeros_like(E)
        keep = ids_all_ref.reshape(-1) != 0
        full.index_add_(0, ids_all_ref.reshape(-1)[keep], d_rows_all.reshape(-1, d)[keep])
        assert torch.allclose(g2[:vs.hi - vs.lo], full[vs.lo:vs.hi], atol=1e-12)
        ret[rank] = 'ok'
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('V', [97, 40, 12])
def test_vocab_parallel_matches_unsharded(V):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + (os.getpid() + V) % 300
    mp.spawn(_worker, args=(world, port, V, ret), nprocs=world, join=True)
    assert dict(ret) == {0: 'ok', 1: 'ok'}
