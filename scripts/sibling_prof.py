"""Per-kernel device time of a training step of a sibling model (ACSSEPT / ACTiSASRec) at the C2 shape: every C-ABI launch of
five eager steps is bracketed by CUDA events (ac-tsr_b200/_lib.py: KernelTimer; an event pair adds ~12 us to a short launch, so
read the shares, not the absolute sum).  The autograd path frees its temporaries, so the launches cannot be replayed in a graph.
    python scripts/sibling_prof.py ACTiSASRec"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ac_tsr_b200 as A
from ac_tsr_b200._lib import KernelTimer
import bench

name = sys.argv[1] if len(sys.argv) > 1 else 'ACTiSASRec'
dev = torch.device('cuda:0')
B, L, V = 256, 50, 12102
base = bench.model_cfg()
base.update(n_layers=2, n_heads=2, hidden_size=64, inner_size=256, MAX_ITEM_LIST_LENGTH=L, USER_ID_FIELD='user_id', ITEM_ID_FIELD='item_id',
            LIST_SUFFIX='_list', ITEM_LIST_LENGTH_FIELD='item_length', NEG_PREFIX='neg_', TIME_FIELD='timestamp', device=dev, seed=42,
            learning_rate=1e-4, epochs=1, train_batch_size=B, eval_batch_size=B, topk=[10, 50], metrics=['Hit'], valid_metric='Hit@10',
            checkpoint_dir='/tmp/acsr_bench_ckpt', cuda_graph=False, user_hidden_size=32, item_hidden_size=32, time_span=256)
config = A.Config(model=name, config_dict=base)
config['model'] = name
ds = A.data.SyntheticSequentialDataset(config, B, V, seed=77, pin=False)
torch.manual_seed(42)
model = getattr(A, name)(config, ds).to(dev)
trainer = getattr(A, name + 'Trainer')(config, model)
batch = A.Interaction({k: v.to(dev) for k, v in ds.inter_feat.interaction.items()})
model.train()
for _ in range(3):
    trainer.train_step(batch)
torch.cuda.synchronize()
t = KernelTimer()
A.LIB.timer = t
n_steps = 5
for _ in range(n_steps):
    trainer.train_step(batch)
A.LIB.timer = None
res = t.summary()
tot = sum(v[1] for v in res.values())
print('%s: %d launches of libacsr kernels per step, %.3f ms event-bracketed device time per step' % (name, t.launches // n_steps, tot / n_steps))
for k, v in sorted(res.items(), key=lambda kv: -kv[1][1]):
    print('  %-30s calls/step %5.1f  %8.1f us/step  (%.1f us/launch)  share %.3f' % (k, v[0] / n_steps, v[1] * 1e3 / n_steps, v[1] * 1e3 / v[0], v[1] / tot))
