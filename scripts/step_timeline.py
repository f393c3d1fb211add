"""Cumulative timeline of the fused training step: CUDA-graph replays of step prefixes (forward, +CE, full) and of the
full step with the weight-gradient stream disabled / PDL off.  usage: python scripts/step_timeline.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as BN

dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
import __graft_entry__ as ge
ge.build()


def measure(stop=None, overlap=True, pdl=True, adam=True, iters=30):
    A, cfg, config, model, trainer = BN.build(dev, 0, 1, cuda_graph=True)
    B, L, V = BN.WORKLOAD['B'], BN.WORKLOAD['L'], BN.WORKLOAD['V']
    seq, ln, tgt = A.data.synth_sequences(B, L, V, seed=42)
    b = A.Interaction({'item_id_list': seq, 'item_length': ln, 'item_id': tgt}).to(dev)
    model.train()
    f = trainer.fused
    f._stop_after, f.overlap_wgrad, f.pdl = stop, overlap, pdl
    if not adam:
        trainer.optimizer.step = lambda: None
    for _ in range(3):
        trainer.graphed_step(b)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); trainer.graphed_step(b); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


print('forward only            %.1f us' % measure('fwd', adam=False))
print('forward + CE            %.1f us' % measure('ce', adam=False))
print('full step               %.1f us' % measure())
print('full, no Adam           %.1f us' % measure(adam=False))
print('full, wgrad in-line     %.1f us' % measure(overlap=False))
print('full, PDL off           %.1f us' % measure(pdl=False))
