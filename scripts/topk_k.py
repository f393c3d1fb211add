import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ac_tsr_b200 as A
def t(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
V, d = 1000001, 64
g = torch.Generator().manual_seed(0)
E = (torch.randn(V, d, generator=g) * 0.5).cuda()
for M in (256, 128):
    out = torch.randn(M, d, generator=g).cuda()
    for k in (1, 10, 50):
        for sb in (True, False):
            print('M=%d k=%d share_bound=%s: %.1f us' % (M, k, sb, t(lambda: A.ops.logits_topk_partial(out, E, k, 0, True, 3, share_bound=sb))))
    print('M=%d ce_partial (same stream, 8 epilogue warps): %.1f us' % (M, t(lambda: A.ops.ce_partial(out, E, 3))))
