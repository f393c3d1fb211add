"""two launches of the fused top-k kernel for ncu: python scripts/topk_prof.py V"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ac_tsr_b200 as A
V = int(sys.argv[1]) if len(sys.argv) > 1 else 1000001
g = torch.Generator().manual_seed(0)
out = torch.randn(256, 64, generator=g).cuda()
E = (torch.randn(V, 64, generator=g) * 0.5).cuda()
for _ in range(2):
    A.ops.logits_topk_partial(out, E, 50)
torch.cuda.synchronize()
print('done')
