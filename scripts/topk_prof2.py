import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ac_tsr_b200 as A
M, V, d = 256, 1000001, 64
g = torch.Generator().manual_seed(0)
out = torch.randn(M, d, generator=g).cuda()
E = (torch.randn(V, d, generator=g) * 0.5).cuda()
for _ in range(2):
    pv, pi = A.ops.logits_topk_partial(out, E, 50)
torch.cuda.synchronize()
