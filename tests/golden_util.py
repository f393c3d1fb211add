"""Load tests/golden/*.npz into the shapes the oracle and the CUDA path take."""
import glob
import os
from collections import OrderedDict

import numpy as np
import torch

from oracle import acsr_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def case_names(train=None):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, '*.npz'))):
        n = os.path.basename(p)[:-4]
        if train is None or n.endswith('_train') == train or (train and '_train' in n) and not n.endswith('_eval'):
            out.append(n)
    return out


def load_case(name, dtype=torch.float32):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    kw = {}
    for k in z.files:
        if k.startswith('cfg.'):
            v = z[k]
            kw[k[4:]] = v.item() if v.shape == () else v.tolist()
    cfg = O.default_cfg(**kw)
    params = OrderedDict((k[6:], torch.from_numpy(z[k]).to(dtype)) for k in z.files if k.startswith('param.'))
    grads = OrderedDict((k[5:], torch.from_numpy(z[k])) for k in z.files if k.startswith('grad.'))
    ph, pa = cfg['hidden_dropout_prob'], cfg['attn_dropout_prob']
    r = {}
    for k in z.files:
        if not k.startswith('rand.'):
            continue
        parts = k.split('.')
        a = torch.from_numpy(z[k])
        if parts[1] in ('emb', 'posK', 'posV', 'timeK', 'timeV'):      # hidden-dropout masks drawn once per forward
            r[parts[1]] = a.to(dtype) / (1 - ph)
        elif parts[2] == 'noise':
            r[(int(parts[1]), 'noise')] = a.to(dtype)
        else:
            p = pa if parts[2] in ('D1', 'D2', 'D3') else ph
            r[(int(parts[1]), parts[2])] = a.to(dtype) / (1 - p)
    batch = dict(item_seq=torch.from_numpy(z['item_id_list']), item_len=torch.from_numpy(z['item_length']),
                 pos=torch.from_numpy(z['item_id']))
    for key in ('masked_seq', 'pos_items', 'neg_items', 'masked_index'):        # AcBERT4Rec: the recorded host-side masking
        if key in z.files:
            batch[key] = torch.from_numpy(z[key])
    if 'timestamp_list' in z.files:               # ACTiSASRec: the time stamp of every position
        batch['time'] = torch.from_numpy(z['timestamp_list'])
    if 'user_id' in z.files:                      # ACSSEPT: the user of every row
        batch['user'] = torch.from_numpy(z['user_id'])
    if 'neg_item_id' in z.files:                  # loss_type BPR: the sampled negative of every row
        batch['neg'] = torch.from_numpy(z['neg_item_id'])
    return dict(z=z, cfg=cfg, bert=('bert' in z.files), ssept=('ssept' in z.files), ti=('ti' in z.files), params=params, grads=grads, rand=O.Rand(r), batch=batch,
                train=bool(z['train']), V=int(z['V']), k=int(z['k']))
