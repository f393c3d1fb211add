"""time acsr_logits_topk_partial + acsr_topk_merge alone (CUDA events, L2 flushed): python scripts/topk_micro.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ac_tsr_b200 as A

def t(fn, n=20):
    flush = torch.empty(64 << 20, dtype=torch.float32, device='cuda')
    for _ in range(3):
        fn()
    ev = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return sum(x.elapsed_time(y) for x, y in ev) / n * 1e3

for M, V, d in ((256, 12102, 64), (256, 20034, 64), (256, 1000001, 64), (256, 12102, 128)):
    g = torch.Generator().manual_seed(0)
    out = torch.randn(M, d, generator=g).cuda()
    E = (torch.randn(V, d, generator=g) * 0.5).cuda()
    pos = torch.randint(1, V, (M,), generator=g).cuda()
    pv, pi = A.ops.logits_topk_partial(out, E, 50)
    t1 = t(lambda: A.ops.logits_topk_partial(out, E, 50))
    t2 = t(lambda: A.ops.topk_merge(pv, pi, 50, pos))
    t3 = t(lambda: A.ops.logits_scores(out, E))
    t4 = 0.0
    if V < 50000:
        sc = A.ops.logits_scores(out, E)
        t4 = t(lambda: A.ops.topk_select(sc, 50, pos))
    print('M=%d V=%d d=%d slots=%d: topk_partial %.1f us, merge %.1f us, store %.1f us, select %.1f us' % (M, V, d, pv.shape[1], t1, t2, t3, t4))
