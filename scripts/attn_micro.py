"""Micro-benchmark of the fused attention kernels alone (CUDA events, L2-resident inputs).
usage: python scripts/attn_micro.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
import ac_tsr_b200 as A
from ac_tsr_b200 import ops

dev = torch.device('cuda')
ORDER = [None]
CTXR = [None]
KEEP = []
USE_ORDER = os.environ.get('ATTN_ORDER', '1') == '1'
H, dh, L = 2, 32, 50
d = H * dh


def run(B, lens, p=0.5, need_att=True, iters=30, bwd=False, last=True, last_fwd=False):
    g = torch.Generator().manual_seed(0)
    t = [torch.randn(B, L, d, generator=g).to(dev) * 0.5 for _ in range(5)]
    gate = torch.randn(B, L, L, generator=g).to(dev)
    seq = torch.zeros(B, L, dtype=torch.int64)
    for b in range(B):
        seq[b, :lens[b]] = 1 + b % 7
    seq = seq.to(dev)
    ow, ob = torch.randn(2 * dh).to(dev) * .1, torch.randn(1).to(dev) * .1
    dw, db, sc = torch.randn(2 * dh).to(dev) * .1, torch.randn(1).to(dev) * .1, torch.randn(1).to(dev)
    rng = ops.DeviceRng(1, dev)
    order = torch.empty(B, dtype=torch.int32, device=dev)
    A.LIB.call('acsr_seq_order', seq.data_ptr(), B, L, order.data_ptr(), ops._stream())
    ORDER[0] = order.data_ptr() if USE_ORDER else None
    lens_dev = torch.tensor(lens, dtype=torch.int64, device=dev)
    KEEP.append(lens_dev)
    CTXR[0] = lens_dev.data_ptr() if os.environ.get('ATTN_CTXROWS', '1') == '1' else None
    ca, cc = torch.empty(B, L, d, device=dev), torch.empty(B, L, d, device=dev)
    pen = torch.zeros(1, dtype=torch.float64, device=dev)
    P = ops._p
    st = None
    shared = (P(t[0]), P(t[1]), P(t[2]), P(t[3]), P(t[4]), P(gate), P(seq, torch.int64), P(ow), P(ob), P(dw), P(db), P(sc),
              B, L, H, dh, 1, 0, 0.0, 0, None, p, None, None, None, None, rng.ptr, 16)
    T = B * L
    if bwd:
        dc = torch.zeros(2 * T, d, device=dev)
        lens_t = torch.tensor(lens)
        idx = torch.arange(B) * L + lens_t - 1
        if last:
            dc[idx] = torch.randn(B, d).to(dev)
            dc[T + idx] = torch.randn(B, d).to(dev)
        else:
            m = (torch.arange(L)[None, :] < lens_t[:, None]).reshape(-1).to(dev)
            dc[:T][m] = torch.randn(int(m.sum()), d, device=dev)
            dc[T:] = torch.randn(T, d, device=dev)
        outs = [torch.empty(2 * T, d, device=dev) for _ in range(5)]
        dgl = torch.zeros(2 * T, L, device=dev)
        pg = [torch.zeros(2 * dh, device=dev), torch.zeros(1, device=dev), torch.zeros(2 * dh, device=dev), torch.zeros(1, device=dev), torch.zeros(1, device=dev)]
        dpen = torch.tensor([1e-3], device=dev)
        att1, cal1 = (dc[T:], None) if last else (None, dc[T:])

        def f():
            A.LIB.call('acsr_attn_calib_bwd2', P(dc[:T]), None, P(att1), P(cal1), P(dpen), *shared, *[P(o) for o in outs], P(dgl),
                       P(pg[0]), P(pg[1]), P(pg[2]), P(pg[3]), P(pg[4]), None, ORDER[0], CTXR[0] if last else None, ops._stream())
    else:
        def f():
            A.LIB.call('acsr_attn_calib_fwd', *shared, P(ca) if need_att else None, P(cc), pen.data_ptr(), None, ORDER[0], CTXR[0] if last_fwd else None, ops._stream())
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        st = ops._stream()
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            st = ops._stream()
            for _ in range(iters):
                f()
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for B in (8, 37, 74, 148, 256, 512):
    seqs, lnn, _ = A.data.synth_sequences(B, L, 1000, seed=42)
    lens = lnn.tolist()
    full = [L] * B
    short = [5] * B
    print('B=%4d CTAs=%4d  fwd us: lognormal %.1f  all-50 %.1f  all-5 %.1f | fwd last-layer %.1f | bwd2 last %.1f / %.1f  lower %.1f / %.1f' % (
        B, B * H, run(B, lens), run(B, full), run(B, short), run(B, lens, last_fwd=True),
        run(B, lens, bwd=True, last=True), run(B, full, bwd=True, last=True),
        run(B, lens, bwd=True, last=False), run(B, full, bwd=True, last=False)))
