"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/<case>.npz.  Nothing is written into /root/reference.

What is recorded per case: the reference's state_dict, the batch, every dropout
mask in call order (mapped to the names oracle.Rand uses), the CPU attack noise
(layers.py:917), and the reference's outputs: forward pair, losses, the
per-parameter .grad after the two routed backward passes of trainer.py:672-686,
full-sort scores / top-k / hit flags, predict scores.
"""
import os
import sys
import types
import logging

os.environ.setdefault('PYTHONDONTWRITEBYTECODE', '1')
sys.dont_write_bytecode = True
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def import_reference():
    for n in ('colorlog', 'colorama', 'thop'):
        sys.modules[n] = types.ModuleType(n)
    sys.modules['colorlog'].ColoredFormatter = lambda fmt, datefmt, log_colors=None: logging.Formatter(
        fmt.replace('%(log_color)s', ''), datefmt)
    sys.modules['colorama'].init = lambda **k: None
    sys.modules['thop'].profile = None
    if not hasattr(np, 'float'):
        np.float = float
    sys.path.insert(0, '/root/reference')
    from recbole.model.sequential_recommender.acsasrec import ACSASRec
    from recbole.data.interaction import Interaction
    torch.autograd.set_detect_anomaly(False)      # sine.py:25 switches it on at import; arithmetic unaffected
    return ACSASRec, Interaction


class FakeDataset:
    def __init__(self, n):
        self.n = n

    def num(self, field):
        return self.n


class Recorder:
    """Replaces nn.Dropout.forward and torch.randn while the reference runs."""

    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)
        self.masks, self.noises = [], []

    def __enter__(self):
        rec = self
        self._fwd = torch.nn.Dropout.forward
        self._randn = torch.randn

        def fwd(mod, x):
            if not mod.training or mod.p <= 0:
                return x
            keep = (torch.rand(x.shape, generator=rec.g) >= mod.p)
            rec.masks.append(keep)
            return x * keep.to(x.dtype) / (1.0 - mod.p)

        def randn(*shape, **kw):
            n = rec._randn(*shape, generator=rec.g)
            rec.noises.append(n)
            return n
        torch.nn.Dropout.forward = fwd
        torch.randn = randn
        return self

    def __exit__(self, *a):
        torch.nn.Dropout.forward = self._fwd
        torch.randn = self._randn


def base_config(**kw):
    from oracle.acsr_oracle import default_cfg
    c = default_cfg(**kw)
    c.update(USER_ID_FIELD='user_id', ITEM_ID_FIELD='item_id', LIST_SUFFIX='_list',
             ITEM_LIST_LENGTH_FIELD='item_length', NEG_PREFIX='neg_', device='cpu')
    return c


def make_case(name, V, B, seed, train, k=50, **cfgkw):
    from oracle.acsr_oracle import synth_batch
    ACSASRec, Interaction = import_reference()
    cfg = base_config(**cfgkw)
    L = cfg['MAX_ITEM_LIST_LENGTH']
    torch.manual_seed(seed)
    model = ACSASRec(cfg, FakeDataset(V))
    # make biases / LN non-trivial so the test exercises them (the init zeroes them)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith('.bias'):
                p.add_(torch.randn(p.shape, generator=g) * 0.02)
            if 'LayerNorm.weight' in n:
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
    seq, ln, pos = synth_batch(B, L, V, seed=seed + 2)
    ln[0] = L                                   # one full-length row
    seq[0] = torch.randint(1, V, (L,), generator=g)
    ln[1] = 1                                   # one minimal row
    seq[1, 1:] = 0
    fields = {'item_id_list': seq, 'item_length': ln, 'item_id': pos}
    if cfg['loss_type'] == 'BPR':                # one sampled negative per row (acsasrec.py:110), never the positive
        neg = torch.randint(1, V - 1, (B,), generator=g)
        neg = neg + (neg >= pos).long()
        fields['neg_item_id'] = neg
    inter = Interaction(fields)
    out = {'V': V, 'B': B, 'k': min(k, V - 1), 'train': int(train), 'seed': seed,
           'item_id_list': seq.numpy(), 'item_length': ln.numpy(), 'item_id': pos.numpy()}
    if 'neg_item_id' in fields:
        out['neg_item_id'] = fields['neg_item_id'].numpy()
    for kk, vv in cfgkw.items():
        out['cfg.' + kk] = np.array(vv)
    for n, p in model.state_dict().items():
        out['param.' + n] = p.detach().numpy().copy()
    model.train(train)
    N = cfg['n_layers']
    with Recorder(seed + 3) as rec:
        if train:
            l_att, l_cal = model.calculate_loss(inter)
        else:
            att, cal, Ms = model.forward(seq, ln)
    # map recorded randomness to names
    if train and (cfg['hidden_dropout_prob'] > 0 or cfg['attn_dropout_prob'] > 0):
        assert cfg['hidden_dropout_prob'] > 0 and cfg['attn_dropout_prob'] > 0
        assert len(rec.masks) == 1 + 7 * N, len(rec.masks)
        out['rand.emb'] = rec.masks[0].numpy().astype(np.uint8)
        for l in range(N):
            for j, key in enumerate(('D1', 'D2', 'D3', 'D4', 'D5', 'D6', 'D7')):
                out['rand.%d.%s' % (l, key)] = rec.masks[1 + 7 * l + j].numpy().astype(np.uint8)
    assert len(rec.noises) == N
    for l in range(N):
        out['rand.%d.noise' % l] = rec.noises[l].numpy()
    if train:
        out['loss_att'] = l_att.detach().numpy()
        out['loss_cal'] = l_cal.detach().numpy()
        # trainer.py:672-686
        def is_attack(n):
            return 'attack_key_transform' in n or 'attack_query_transform' in n
        for n, p in model.named_parameters():
            p.requires_grad = not is_attack(n)
        l_cal.backward(retain_graph=True)
        for n, p in model.named_parameters():
            p.requires_grad = is_attack(n)
        l_att.backward()
        for n, p in model.named_parameters():
            p.requires_grad = True
            gr = p.grad if p.grad is not None else torch.zeros_like(p)
            out['grad.' + n] = gr.detach().numpy().copy()
    else:
        out['out_att'] = att.detach().numpy()
        out['out_cal'] = cal.detach().numpy()
        for l, M in enumerate(Ms):
            out['pen_sq.%d' % l] = torch.sum((1 - M) ** 2).detach().numpy()
        with Recorder(seed + 3):                  # same noise again
            _, scores = model.full_sort_predict(inter)
        scores = scores.detach().clone()
        out['scores'] = scores.numpy().copy()
        scores[:, 0] = -np.inf                    # trainer.py:942
        _, idx = torch.topk(scores, out['k'], dim=-1)
        out['topk_idx'] = idx.numpy()
        pm = torch.zeros_like(scores, dtype=torch.int)
        pm[torch.arange(B), pos] = 1              # collector.py:148-152
        out['rec_topk'] = torch.cat((torch.gather(pm, 1, idx), pm.sum(1, keepdim=True)), 1).numpy()
        with Recorder(seed + 3):
            a_s, c_s = model.predict(inter)
        out['predict_att'] = a_s.detach().numpy()
        out['predict_cal'] = c_s.detach().numpy()
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print(name, os.path.getsize(path) // 1024, 'KB')


def make_bert_case(name, V, B, seed, train, k=50, **cfgkw):
    """AcBERT4Rec (acbert4rec.py): bidirectional mask, masked-item CE.  The host-side masking of reconstruct_train_data
    (python `random`) is recorded, so the case pins both the masking procedure (random.seed given) and the arithmetic."""
    import random
    from oracle.acsr_oracle import synth_batch
    import_reference()
    from recbole.model.sequential_recommender.acbert4rec import AcBERT4Rec
    from recbole.data.interaction import Interaction
    cfg = base_config(**cfgkw)
    cfg['mask_ratio'] = cfgkw.get('mask_ratio', 0.2)
    L = cfg['MAX_ITEM_LIST_LENGTH']
    torch.manual_seed(seed)
    model = AcBERT4Rec(cfg, FakeDataset(V))
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith('.bias'):
                p.add_(torch.randn(p.shape, generator=g) * 0.02)
            if 'LayerNorm.weight' in n:
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
    seq, ln, pos = synth_batch(B, L, V, seed=seed + 2)
    ln[0] = L - 1                               # eval appends one position: keep room for the mask token (acbert4rec.py:152-160)
    seq[0] = torch.randint(1, V, (L,), generator=g)
    seq[0, L - 1] = 0
    ln = ln.clamp(max=L - 1)
    seq[torch.arange(L).view(1, L) >= ln.view(B, 1)] = 0
    inter = Interaction({'item_id_list': seq, 'item_length': ln, 'item_id': pos})
    out = {'V': V, 'B': B, 'k': min(k, V - 1), 'train': int(train), 'seed': seed, 'bert': 1,
           'item_id_list': seq.numpy(), 'item_length': ln.numpy(), 'item_id': pos.numpy()}
    for kk, vv in cfgkw.items():
        out['cfg.' + kk] = np.array(vv)
    out['cfg.mask_ratio'] = np.array(cfg['mask_ratio'])
    for n, p in model.state_dict().items():
        out['param.' + n] = p.detach().numpy().copy()
    model.train(train)
    N = cfg['n_layers']
    if train:
        captured = []
        orig = model.reconstruct_train_data

        def wrap(x):
            r = orig(x)
            captured.append(r)
            return r
        model.reconstruct_train_data = wrap
        random.seed(seed + 5)
        with Recorder(seed + 3) as rec:
            l_att, l_cal = model.calculate_loss(inter)
        masked_seq, pos_items, neg_items, masked_index = captured[0]
        out['masked_seq'] = masked_seq.numpy(); out['pos_items'] = pos_items.numpy()
        out['neg_items'] = neg_items.numpy(); out['masked_index'] = masked_index.numpy()
        out['random_seed'] = seed + 5
        assert len(rec.masks) == 1 + 7 * N, len(rec.masks)
        out['rand.emb'] = rec.masks[0].numpy().astype(np.uint8)
        for l in range(N):
            for j, key in enumerate(('D1', 'D2', 'D3', 'D4', 'D5', 'D6', 'D7')):
                out['rand.%d.%s' % (l, key)] = rec.masks[1 + 7 * l + j].numpy().astype(np.uint8)
        for l in range(N):
            out['rand.%d.noise' % l] = rec.noises[l].numpy()
        out['loss_att'] = l_att.detach().numpy()
        out['loss_cal'] = l_cal.detach().numpy()

        def is_attack(n):
            return 'attack_key_transform' in n or 'attack_query_transform' in n
        for n, p in model.named_parameters():
            p.requires_grad = not is_attack(n)
        l_cal.backward(retain_graph=True)
        for n, p in model.named_parameters():
            p.requires_grad = is_attack(n)
        l_att.backward()
        for n, p in model.named_parameters():
            gr = p.grad if p.grad is not None else torch.zeros_like(p)
            out['grad.' + n] = gr.detach().numpy().copy()
    else:
        with Recorder(seed + 3) as rec:
            a_scores, scores = model.full_sort_predict(inter)
        for l in range(N):
            out['rand.%d.noise' % l] = rec.noises[l].numpy()
        scores = scores.detach().clone()
        out['scores'] = scores.numpy().copy()
        out['scores_att'] = a_scores.detach().numpy().copy()
        scores[:, 0] = -np.inf
        _, idx = torch.topk(scores, out['k'], dim=-1)
        out['topk_idx'] = idx.numpy()
        with Recorder(seed + 3):
            a_s, c_s = model.predict(inter)
        out['predict_att'] = a_s.detach().numpy()
        out['predict_cal'] = c_s.detach().numpy()
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print(name, os.path.getsize(path) // 1024, 'KB')


def make_ssept_case(name, V, U, B, seed, train, k=50, **cfgkw):
    """ACSSEPT (acssept.py on transformer_layers.py:742-953): user embedding concatenated to every position, the encoder
    variant without the re-normalising softmaxes.  The reference registers no ACSSEPTTrainer; the recorded gradients are those
    of the AC training step the model's tuple API is written for (AttackSASRecTrainer, trainer.py:672-686)."""
    from oracle.acsr_oracle import synth_batch
    import_reference()
    from recbole.model.sequential_recommender.acssept import ACSSEPT
    from recbole.data.interaction import Interaction
    cfg = base_config(**cfgkw)
    cfg.setdefault('user_hidden_size', 32)
    cfg.setdefault('item_hidden_size', 32)
    L = cfg['MAX_ITEM_LIST_LENGTH']

    class DS:
        def num(self, field):
            return U if field == 'user_id' else V
    torch.manual_seed(seed)
    model = ACSSEPT(cfg, DS())
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith('.bias'):
                p.add_(torch.randn(p.shape, generator=g) * 0.02)
            if 'LayerNorm.weight' in n:
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
    seq, ln, pos = synth_batch(B, L, V, seed=seed + 2)
    ln[0] = L
    seq[0] = torch.randint(1, V, (L,), generator=g)
    ln[1] = 1
    seq[1, 1:] = 0
    user = torch.randint(1, U, (B,), generator=g)
    fields = {'item_id_list': seq, 'item_length': ln, 'item_id': pos, 'user_id': user}
    if cfg['loss_type'] == 'BPR':
        neg = torch.randint(1, V - 1, (B,), generator=g)
        neg = neg + (neg >= pos).long()
        fields['neg_item_id'] = neg
    inter = Interaction(fields)
    out = {'V': V, 'U': U, 'B': B, 'k': min(k, V - 1), 'train': int(train), 'seed': seed, 'ssept': 1,
           'item_id_list': seq.numpy(), 'item_length': ln.numpy(), 'item_id': pos.numpy(), 'user_id': user.numpy()}
    if 'neg_item_id' in fields:
        out['neg_item_id'] = fields['neg_item_id'].numpy()
    for kk, vv in cfgkw.items():
        out['cfg.' + kk] = np.array(vv)
    out['cfg.user_hidden_size'] = np.array(cfg['user_hidden_size'])
    out['cfg.item_hidden_size'] = np.array(cfg['item_hidden_size'])
    for n, p in model.state_dict().items():
        out['param.' + n] = p.detach().numpy().copy()
    model.train(train)
    N = cfg['n_layers']
    if train:
        with Recorder(seed + 3) as rec:
            l_att, l_cal = model.calculate_loss(inter)
        assert len(rec.masks) == 1 + 7 * N, len(rec.masks)
        out['rand.emb'] = rec.masks[0].numpy().astype(np.uint8)
        for l in range(N):
            for j, key in enumerate(('D1', 'D2', 'D3', 'D4', 'D5', 'D6', 'D7')):
                out['rand.%d.%s' % (l, key)] = rec.masks[1 + 7 * l + j].numpy().astype(np.uint8)
        for l in range(N):
            out['rand.%d.noise' % l] = rec.noises[l].numpy()
        out['loss_att'] = l_att.detach().numpy()
        out['loss_cal'] = l_cal.detach().numpy()

        def is_attack(n):
            return 'attack_key_transform' in n or 'attack_query_transform' in n
        for n, p in model.named_parameters():
            p.requires_grad = not is_attack(n)
        l_cal.backward(retain_graph=True)
        for n, p in model.named_parameters():
            p.requires_grad = is_attack(n)
        l_att.backward()
        for n, p in model.named_parameters():
            gr = p.grad if p.grad is not None else torch.zeros_like(p)
            out['grad.' + n] = gr.detach().numpy().copy()
    else:
        with Recorder(seed + 3) as rec:
            a_scores, scores = model.full_sort_predict(inter)
        for l in range(N):
            out['rand.%d.noise' % l] = rec.noises[l].numpy()
        with Recorder(seed + 3):
            att, cal, Ms = model.forward(seq, ln, user)
        out['out_att'] = att.detach().numpy()
        out['out_cal'] = cal.detach().numpy()
        for l, M in enumerate(Ms):
            out['pen_sq.%d' % l] = torch.sum((1 - M) ** 2).detach().numpy()
        scores = scores.detach().clone()
        out['scores'] = scores.numpy().copy()
        out['scores_att'] = a_scores.detach().numpy().copy()
        scores[:, 0] = -np.inf
        _, idx = torch.topk(scores, out['k'], dim=-1)
        out['topk_idx'] = idx.numpy()
        with Recorder(seed + 3):
            a_s, c_s = model.predict(inter)
        out['predict_att'] = a_s.detach().numpy()
        out['predict_cal'] = c_s.detach().numpy()
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print(name, os.path.getsize(path) // 1024, 'KB')


def make_ti_case(name, V, B, seed, train, k=50, **cfgkw):
    """ACTiSASRec (actisasrec.py on transformer_layers.py:1010-1327): time-interval aware keys / values.  As for ACSSEPT the
    reference registers no trainer of that name; the recorded gradients are those of the AC step (trainer.py:672-686)."""
    from oracle.acsr_oracle import synth_batch
    import_reference()
    from recbole.model.sequential_recommender.actisasrec import ACTiSASRec
    from recbole.data.interaction import Interaction
    cfg = base_config(**cfgkw)
    cfg.setdefault('time_span', 16)
    cfg['TIME_FIELD'] = 'timestamp'
    L = cfg['MAX_ITEM_LIST_LENGTH']
    torch.manual_seed(seed)
    model = ACTiSASRec(cfg, FakeDataset(V))
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith('.bias'):
                p.add_(torch.randn(p.shape, generator=g) * 0.02)
            if 'LayerNorm.weight' in n:
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
    seq, ln, pos = synth_batch(B, L, V, seed=seed + 2)
    ln[0] = L
    seq[0] = torch.randint(1, V, (L,), generator=g)
    ln[1] = 1
    seq[1, 1:] = 0
    # increasing integer-valued timestamps (float field, like the atomic files), 0 on the padding (sequential_dataset.py:128-132)
    gaps = torch.randint(0, 7, (B, L), generator=g)
    ts = (torch.cumsum(gaps, 1) + 1000).float()
    ts[seq == 0] = 0.0
    fields = {'item_id_list': seq, 'item_length': ln, 'item_id': pos, 'timestamp_list': ts}
    if cfg['loss_type'] == 'BPR':
        neg = torch.randint(1, V - 1, (B,), generator=g)
        neg = neg + (neg >= pos).long()
        fields['neg_item_id'] = neg
    inter = Interaction(fields)
    out = {'V': V, 'B': B, 'k': min(k, V - 1), 'train': int(train), 'seed': seed, 'ti': 1,
           'item_id_list': seq.numpy(), 'item_length': ln.numpy(), 'item_id': pos.numpy(), 'timestamp_list': ts.numpy()}
    if 'neg_item_id' in fields:
        out['neg_item_id'] = fields['neg_item_id'].numpy()
    for kk, vv in cfgkw.items():
        out['cfg.' + kk] = np.array(vv)
    out['cfg.time_span'] = np.array(cfg['time_span'])
    for n, p in model.state_dict().items():
        out['param.' + n] = p.detach().numpy().copy()
    model.train(train)
    N = cfg['n_layers']
    if train:
        with Recorder(seed + 3) as rec:
            l_att, l_cal = model.calculate_loss(inter)
        assert len(rec.masks) == 5 + 7 * N, len(rec.masks)
        for j, key in enumerate(('emb', 'posK', 'posV', 'timeK', 'timeV')):      # actisasrec.py:118-124
            out['rand.' + key] = rec.masks[j].numpy().astype(np.uint8)
        for l in range(N):
            for j, key in enumerate(('D1', 'D2', 'D3', 'D4', 'D5', 'D6', 'D7')):
                out['rand.%d.%s' % (l, key)] = rec.masks[5 + 7 * l + j].numpy().astype(np.uint8)
        for l in range(N):
            out['rand.%d.noise' % l] = rec.noises[l].numpy()
        out['loss_att'] = l_att.detach().numpy()
        out['loss_cal'] = l_cal.detach().numpy()

        def is_attack(n):
            return 'attack_key_transform' in n or 'attack_query_transform' in n
        for n, p in model.named_parameters():
            p.requires_grad = not is_attack(n)
        l_cal.backward(retain_graph=True)
        for n, p in model.named_parameters():
            p.requires_grad = is_attack(n)
        l_att.backward()
        for n, p in model.named_parameters():
            gr = p.grad if p.grad is not None else torch.zeros_like(p)
            out['grad.' + n] = gr.detach().numpy().copy()
    else:
        with Recorder(seed + 3) as rec:
            a_scores, scores = model.full_sort_predict(inter)
        for l in range(N):
            out['rand.%d.noise' % l] = rec.noises[l].numpy()
        with Recorder(seed + 3):
            att, cal, Ms = model.forward(seq, ln, model.get_time_matrix(ts))
        out['out_att'] = att.detach().numpy()
        out['out_cal'] = cal.detach().numpy()
        for l, M in enumerate(Ms):
            out['pen_sq.%d' % l] = torch.sum((1 - M) ** 2).detach().numpy()
        scores = scores.detach().clone()
        out['scores'] = scores.numpy().copy()
        out['scores_att'] = a_scores.detach().numpy().copy()
        scores[:, 0] = -np.inf
        _, idx = torch.topk(scores, out['k'], dim=-1)
        out['topk_idx'] = idx.numpy()
        with Recorder(seed + 3):
            a_s, c_s = model.predict(inter)
        out['predict_att'] = a_s.detach().numpy()
        out['predict_cal'] = c_s.detach().numpy()
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print(name, os.path.getsize(path) // 1024, 'KB')


TI_CASES = [
    # ACTiSASRec
    ('ti_gate_train', 131, 3, 51, True, dict(n_layers=2, hidden_size=32, inner_size=64)),
    ('ti_fixed_train', 131, 3, 52, True, dict(n_layers=1, n_heads=4, hidden_size=32, inner_size=64, combine_option='fixed',
                                              two_level=False, rich_calibrated_combine='fixed')),
    ('ti_gate_eval', 131, 3, 53, False, dict(n_layers=2, hidden_size=32, inner_size=64)),
]

# (the reference's ACSSEPT only runs with user_hidden_size == item_hidden_size: acssept.py:127 expands the user vector "as" the item
# embeddings; the implementation here takes any pair of widths)
# more configuration branches of the two transformer_layers.py models, used by tests/test_oracle_golden.py ONLY (they pin the
# oracle restatement; the CUDA path's own branch coverage is the kernel test matrix): file names start with `oracleonly_`
ORACLE_ONLY_TI = [
    ('oracleonly_ti_bpr_noorder_train', 101, 3, 61, True, dict(n_layers=1, hidden_size=32, inner_size=64, loss_type='BPR', use_order=False,
                                                              trainable_mask_loss_weight=True)),
    ('oracleonly_ti_nodist_onelevel_train', 101, 3, 62, True, dict(n_layers=2, hidden_size=32, inner_size=64, use_distance=False, two_level=False,
                                                                  rich_calibrated_combine='trainable', time_span=8)),
]
ORACLE_ONLY_SSEPT = [
    ('oracleonly_ssept_noorder_tw_train', 101, 13, 3, 63, True, dict(n_layers=2, n_heads=4, use_order=False, trainable_mask_loss_weight=True,
                                                                    item_hidden_size=16, user_hidden_size=16)),
    ('oracleonly_ssept_relu_eval', 101, 13, 3, 64, False, dict(n_layers=1, hidden_act='relu', combine_option='fixed', item_hidden_size=24,
                                                              user_hidden_size=24)),
]

SSEPT_CASES = [
    # ACSSEPT.  hidden = item_hidden_size + user_hidden_size
    ('ssept_gate_train', 151, 23, 3, 41, True, dict(n_layers=2)),
    ('ssept_fixed_onelevel_train', 131, 17, 3, 42, True, dict(n_layers=1, n_heads=4, combine_option='fixed', two_level=False,
                                                             rich_calibrated_combine='fixed', use_position_embedding=True)),
    ('ssept_bpr_train', 131, 17, 3, 44, True, dict(n_layers=1, loss_type='BPR', item_hidden_size=64, user_hidden_size=64)),
    ('ssept_gate_eval', 151, 23, 3, 43, False, dict(n_layers=2)),
]

BERT_CASES = [
    # AcBERT4Rec.  `gate` trains in the reference but cannot be evaluated there (the gate is Linear(d, 50)
    # and evaluation runs on L+1 = 51 positions, layers.py:878/887), so the eval case uses `fixed`.
    ('bert_fixed_train', 151, 3, 31, True, dict(n_layers=2, combine_option='fixed')),
    ('bert_gate_train', 151, 3, 32, True, dict(n_layers=1, n_heads=4)),
    ('bert_fixed_eval', 151, 3, 33, False, dict(n_layers=2, combine_option='fixed')),
]

CASES = [
    # name, V, B, seed, train, cfg overrides
    ('c1_eval', 301, 4, 11, False, {}),
    ('c1_train', 301, 4, 12, True, {}),
    ('c1_train_p0', 301, 4, 13, True, dict(hidden_dropout_prob=0.0, attn_dropout_prob=0.0)),
    ('beauty_train', 257, 3, 14, True, dict(n_layers=3, n_heads=4, inner_size=128)),
    ('beauty_eval', 257, 3, 15, False, dict(n_layers=3, n_heads=4, inner_size=128)),
    ('pos_tw_train', 131, 3, 16, True, dict(n_layers=1, use_position_embedding=True,
                                           trainable_mask_loss_weight=True)),
    ('noorder_train', 131, 3, 17, True, dict(n_layers=1, use_order=False)),
    ('nodist_train', 131, 3, 18, True, dict(n_layers=1, use_distance=False)),
    ('fixed_onelevel_train', 131, 3, 19, True, dict(n_layers=2, combine_option='fixed', two_level=False,
                                                   rich_calibrated_combine='fixed')),
    ('relu_h8_eval', 131, 3, 20, False, dict(n_layers=1, n_heads=8, hidden_size=128, inner_size=64,
                                             hidden_act='relu')),
    # long sequences (L > 64: attn_long.cu) at a hidden size off the tensor-core logits path (logits_simt.cu); the shape family
    # of BASELINE configuration #5 (L=200, d=256, 4 heads) at a size that stays a small fixture
    # (the reference's gate is Linear(d, 50) whatever the sequence length -- acsasrec.py never passes seq_length, layers.py:878 --
    # so the unmodified reference only runs L != 50 with combine_option fixed / annealing)
    # loss_type BPR (acsasrec.py:109-116, loss.py:21-47): one sampled negative per row
    ('bpr_train', 131, 4, 23, True, dict(n_layers=2, loss_type='BPR')),
    ('long_train', 151, 2, 21, True, dict(n_layers=2, n_heads=4, hidden_size=128, inner_size=256, MAX_ITEM_LIST_LENGTH=100,
                                          combine_option='fixed')),
    ('long_eval', 151, 2, 22, False, dict(n_layers=2, n_heads=4, hidden_size=128, inner_size=256, MAX_ITEM_LIST_LENGTH=100,
                                          combine_option='fixed', hidden_act='swish')),
]

if __name__ == '__main__':
    only = set(sys.argv[1:])
    for name, V, B, seed, train, kw in CASES:
        if only and name not in only:
            continue
        make_case(name, V, B, seed, train, **kw)
    for name, V, B, seed, train, kw in BERT_CASES:
        if only and name not in only:
            continue
        make_bert_case(name, V, B, seed, train, **kw)
    for name, V, B, seed, train, kw in TI_CASES + ORACLE_ONLY_TI:
        if only and name not in only:
            continue
        make_ti_case(name, V, B, seed, train, **kw)
    for name, V, U, B, seed, train, kw in SSEPT_CASES + ORACLE_ONLY_SSEPT:
        if only and name not in only:
            continue
        make_ssept_case(name, V, U, B, seed, train, **kw)
