"""Atomic file -> SequentialDataset -> leave-one-out loaders (ac-tsr_b200/dataset.py) against digests of the tensors the
UNMODIFIED reference pipeline produces on the same ml-100k file (tests/golden/make_dataset_golden.py).  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import ac_tsr_b200 as A

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
REF = json.load(open(os.path.join(GOLD, 'ml100k_dataset.json')))


def make_config(extra):
    cd = dict(data_path=GOLD + '/', load_col={'inter': ['user_id', 'item_id', 'rating', 'timestamp']},
              eval_args={'split': {'LS': 'valid_and_test'}, 'group_by': 'user', 'order': 'TO', 'mode': 'full'},
              device=torch.device('cpu'), train_batch_size=256, eval_batch_size=256)
    cd.update(extra)
    return A.Config(model='ACSASRec', dataset='ml-100k', config_dict=cd)


def digest(t):
    a = t.contiguous().numpy()
    return {'shape': list(a.shape), 'dtype': str(a.dtype), 'sha256': hashlib.sha256(a.tobytes()).hexdigest()}


@pytest.mark.parametrize('name', sorted(REF))
def test_timestamp_list_matches_reference_pipeline(name):
    """ACTiSASRec reads `timestamp_list` (actisasrec.py:35, 177): emitted for that model only, bit-identical to the reference's
    augmentation (sequential_dataset.py:112-135)"""
    ref = REF[name]
    config = make_config(ref['config'])
    assert 'timestamp_list' not in A.create_dataset(config).build()[0].inter_feat
    config['model'] = 'ACTiSASRec'
    parts = dict(zip(('train', 'valid', 'test'), A.create_dataset(config).build()))
    for part, d in parts.items():
        assert digest(d.inter_feat['timestamp_list']) == ref['%s.timestamp_list' % part], part
        assert digest(d.inter_feat['item_id_list']) == ref['%s.item_id_list' % part], part


@pytest.mark.parametrize('name', sorted(REF))
def test_dataset_matches_reference_pipeline(name):
    ref = REF[name]
    config = make_config(ref['config'])
    ds = A.create_dataset(config)
    assert ds.user_num == ref['user_num'] and ds.item_num == ref['item_num']
    assert ds.num('item_id') == ref['item_num'] and ds.field2id_token['item_id'][0] == '[PAD]'
    assert hashlib.sha256('\n'.join(ds.field2id_token['item_id']).encode()).hexdigest() == ref['item_token_sha256']
    parts = dict(zip(('train', 'valid', 'test'), ds.build()))
    for part, d in parts.items():
        for f in ('user_id', 'item_id', 'item_length', 'item_id_list'):
            assert digest(d.inter_feat[f]) == ref['%s.%s' % (part, f)], (part, f)
    # structural properties of the augmentation / split
    tr, va, te = parts['train'].inter_feat, parts['valid'].inter_feat, parts['test'].inter_feat
    L = config['MAX_ITEM_LIST_LENGTH']
    for d in (tr, va, te):
        if len(d) == 0:                                                # valid_only / test_only leave one part empty
            continue
        ln, seq = d['item_length'], d['item_id_list']
        assert int(ln.min()) >= 1 and int(ln.max()) <= L and seq.shape[1] == L
        assert bool(((seq != 0).sum(1) == ln).all())                       # right-padded with 0, no 0 inside the prefix
        assert bool((seq[:, 0] != 0).all())
    assert len(set(va['user_id'].tolist())) == len(va) and len(set(te['user_id'].tolist())) == len(te)
    # the test row of a user extends its valid row by exactly the valid target
    vmap = {int(u): i for i, u in enumerate(va['user_id'].tolist())}
    for j in range(0, len(te) if len(va) else 0, 37):
        u = int(te['user_id'][j])
        i = vmap[u]
        lv = int(va['item_length'][i])
        if lv < L:
            assert int(te['item_length'][j]) == lv + 1
            assert torch.equal(te['item_id_list'][j, :lv], va['item_id_list'][i, :lv])
            assert int(te['item_id_list'][j, lv]) == int(va['item_id'][i])


def test_loaders_and_errors():
    config = make_config({})
    ds = A.create_dataset(config)
    train_data, valid_data, test_data = A.data_preparation(config, ds)
    assert len(train_data) == int(np.ceil(97171 / 256)) and len(valid_data) == int(np.ceil(943 / 256))
    batch = next(iter(train_data))
    assert batch['item_id_list'].shape == (256, 50) and batch['item_id_list'].dtype == torch.int64
    inter, hist, pu, pi = next(iter(valid_data))
    assert hist is None and torch.equal(pu, torch.arange(256)) and torch.equal(pi, inter['item_id'])
    with pytest.raises(ValueError):
        A.SequentialDataset(make_config({'eval_args': {'split': {'LS': 'valid_and_test'}, 'order': 'RO', 'group_by': 'user', 'mode': 'full'}})).build()
    with pytest.raises(ValueError):
        A.SequentialDataset(A.Config(model='ACSASRec', dataset='nope', config_dict=dict(data_path=GOLD + '/', device=torch.device('cpu'))))
    with pytest.raises(ValueError):
        A.quick_start.get_model('SASRec')


def test_train_loader_epochs_cover_every_row_once():
    """two epochs over the real training split: every augmented row exactly once per epoch, full batches arrive packed (one
    buffer per epoch, refilled in place), the ragged tail as a plain Interaction."""
    config = make_config({})
    ds = A.create_dataset(config)
    train_data, _, _ = A.data_preparation(config, ds)
    n = len(train_data.dataset)
    key = lambda it: (it['item_id'] * 100003 + it['item_length'] * 7 + it['item_id_list'].sum(1))      # row fingerprint
    want = torch.sort(key(train_data.dataset.inter_feat)).values
    ptrs = []
    for epoch in range(2):
        torch.manual_seed(epoch)
        rows, packed = [], 0
        for batch in train_data:
            rows.append(key(batch))
            if isinstance(batch, A.compat.PackedInteraction):
                packed += 1
                assert batch['item_id_list'].shape == (256, 50)
        assert packed == n // 256 and len(rows) == int(np.ceil(n / 256))
        assert torch.equal(torch.sort(torch.cat(rows)).values, want)
        ptrs.append(train_data._packed.buf.data_ptr())
    assert ptrs[0] == ptrs[1]
