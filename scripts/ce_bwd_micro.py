"""micro-benchmark of the CE forward / backward kernels at a given catalogue size.  usage: python scripts/ce_bwd_micro.py [V] [M]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
import ac_tsr_b200 as A

V = int(sys.argv[1]) if len(sys.argv) > 1 else 1000001
M = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device('cuda', 0)
g = torch.Generator().manual_seed(1)
out = (torch.randn(M, 64, generator=g) * 2).to(dev)
E = (torch.randn(V, 64, generator=g) * 0.3).to(dev)
tgt = torch.randint(0, V, (M,), generator=g).to(dev)
scale = torch.full((M,), 1.0 / M, device=dev)
part = A.ops.ce_partial(out, E, 3)
lse, _, _, _ = A.ops.ce_finalize(part, out, E, tgt, 2)
d_out = torch.zeros(M, 64, device=dev)
d_E = torch.zeros(V, 64, device=dev)


def timeit(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for passes in (3, 1):
    print('V=%d M=%d passes=%d' % (V, M, passes))
    print('  ce_partial   %8.1f us' % timeit(lambda: A.ops.ce_partial(out, E, passes)))
    print('  ce_bwd_dout  %8.1f us' % timeit(lambda: A.ops.ce_bwd_dout(out, E, lse, tgt, scale, d_out, passes)))
    print('  ce_bwd_dtable%8.1f us (M=%d rows)' % (timeit(lambda: A.ops.ce_bwd_dtable(out[:M // 2], E, lse[:M // 2], tgt[:M // 2], scale[:M // 2], d_E, passes)), M // 2))
