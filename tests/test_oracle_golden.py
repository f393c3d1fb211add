"""The oracle (oracle/acsr_oracle.py) against every golden vector made from the real reference."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import acsr_oracle as O
from golden_util import GOLDEN_DIR, load_case

EVERY = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))
ALL = [n for n in EVERY if not n.startswith(('bert_', 'ssept_', 'ti_', 'oracleonly_'))]
TI = [n for n in EVERY if n.startswith(('ti_', 'oracleonly_ti_'))]              # ACTiSASRec cases (actisasrec.py on transformer_layers.py)
SSEPT = [n for n in EVERY if n.startswith(('ssept_', 'oracleonly_ssept_'))]      # oracleonly_*: more config branches, pin the oracle only        # ACSSEPT cases (acssept.py on transformer_layers.py)
BERT = [n for n in EVERY if n.startswith('bert_')]          # AcBERT4Rec cases (acbert4rec.py)
TRAIN = [n for n in ALL if '_train' in n]
EVAL = [n for n in ALL if '_eval' in n]


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def test_fixture_inventory():
    assert len(TRAIN) >= 6 and len(EVAL) >= 3


@pytest.mark.parametrize('name', EVAL)
def test_eval_forward_scores_topk(name):
    c = load_case(name)
    b = c['batch']
    att, cal, Ms = O.forward(c['params'], c['cfg'], b['item_seq'], b['item_len'], c['rand'])
    z = c['z']
    assert rel(att, z['out_att']) < 2e-5
    assert rel(cal, z['out_cal']) < 2e-5
    for l, M in enumerate(Ms):
        assert rel(torch.sum((1 - M) ** 2), z['pen_sq.%d' % l]) < 1e-5
    scores = O.full_sort_scores(c['params'], c['cfg'], b['item_seq'], b['item_len'], c['rand'])
    assert rel(scores, z['scores']) < 2e-5          # north_star: logits within 1e-3 relative
    _, idx = O.full_sort_topk(scores, c['k'])
    ok, nbad = O.topk_equal_modulo_ties(idx, torch.from_numpy(z['topk_idx']), torch.from_numpy(z['scores']))
    assert ok, nbad
    flags = O.hit_flags(torch.from_numpy(z['topk_idx']), b['pos'])
    assert np.array_equal(flags.numpy(), z['rec_topk'][:, :-1])
    assert (z['rec_topk'][:, -1] == 1).all()
    pa, pc = O.predict(c['params'], c['cfg'], b['item_seq'], b['item_len'], b['pos'], c['rand'])
    assert rel(pa, z['predict_att']) < 2e-5 and rel(pc, z['predict_cal']) < 2e-5


@pytest.mark.parametrize('name', TRAIN)
def test_train_losses_and_routed_grads(name):
    c = load_case(name)
    b = c['batch']
    l_att, l_cal, grads = O.train_grads(c['params'], c['cfg'], b['item_seq'], b['item_len'], b['pos'], c['rand'], neg_items=b.get('neg'))
    z = c['z']
    assert abs(float(l_att) - float(z['loss_att'])) < 1e-5 * abs(float(z['loss_att']))
    assert abs(float(l_cal) - float(z['loss_cal'])) < 1e-5 * abs(float(z['loss_cal']))
    assert set(grads) == set(c['grads'])
    for n, g in c['grads'].items():
        scale = float(g.abs().max())
        err = float((grads[n] - g).abs().max())
        assert err <= 2e-4 * scale + 1e-9, (n, err, scale)


def test_metrics_against_reference_formulas():
    rng = np.random.RandomState(0)
    pos = np.zeros((64, 50), dtype=bool)
    for r in range(64):
        if rng.rand() < 0.7:
            pos[r, rng.randint(50)] = True
    m = O.topk_metrics(pos, np.ones(64, dtype=np.int64))
    for k in (1, 3, 5, 10, 20, 50):
        hit = pos[:, :k].any(1).mean()
        assert m['hit@%d' % k] == round(float(hit), 4) == m['recall@%d' % k]
        rank = np.where(pos.any(1), pos.argmax(1), 10 ** 6)
        assert m['mrr@%d' % k] == round(float(np.where(rank < k, 1.0 / (rank + 1), 0).mean()), 4)
        assert m['ndcg@%d' % k] == round(float(np.where(rank < k, 1.0 / np.log2(rank + 2), 0).mean()), 4)


@pytest.mark.parametrize('name', [n for n in BERT if '_train' in n])
def test_bert_train_losses_and_routed_grads(name):
    """AcBERT4Rec (acbert4rec.py:207-245): bidirectional mask + masked-item CE, against the reference's losses and routed .grad"""
    c = load_case(name)
    b, z = c['batch'], c['z']
    l_att, l_cal, grads = O.bert_train_grads(c['params'], c['cfg'], b['masked_seq'], b['pos_items'], b['masked_index'], c['rand'])
    assert abs(float(l_att) - float(z['loss_att'])) < 1e-5 * abs(float(z['loss_att']))
    assert abs(float(l_cal) - float(z['loss_cal'])) < 1e-5 * abs(float(z['loss_cal']))
    assert set(grads) == set(c['grads'])
    for n, g in c['grads'].items():
        scale = float(g.abs().max())
        err = float((grads[n] - g).abs().max())
        assert err <= 2e-4 * scale + 1e-9, (n, err, scale)


@pytest.mark.parametrize('name', [n for n in BERT if '_eval' in n])
def test_bert_eval_scores(name):
    c = load_case(name)
    b, z = c['batch'], c['z']
    sa, sc = O.bert_full_sort_scores(c['params'], c['cfg'], b['item_seq'], b['item_len'], c['rand'])
    assert rel(sc, z['scores']) < 2e-5 and rel(sa, z['scores_att']) < 2e-5
    _, idx = O.full_sort_topk(sc, c['k'])
    ok, nbad = O.topk_equal_modulo_ties(idx, torch.from_numpy(z['topk_idx']), torch.from_numpy(z['scores']))
    assert ok, nbad


@pytest.mark.parametrize('name', [n for n in SSEPT if '_train' in n])
def test_ssept_train_losses_and_routed_grads(name):
    """ACSSEPT (acssept.py:174-190) on the transformer_layers.py encoder (no re-normalising softmaxes): losses + routed .grad"""
    c = load_case(name)
    b, z = c['batch'], c['z']
    l_att, l_cal, grads = O.ssept_train_grads(c['params'], c['cfg'], b['item_seq'], b['item_len'], b['user'], b['pos'], c['rand'],
                                              neg_items=b.get('neg'))
    assert abs(float(l_att) - float(z['loss_att'])) < 1e-5 * abs(float(z['loss_att']))
    assert abs(float(l_cal) - float(z['loss_cal'])) < 1e-5 * abs(float(z['loss_cal']))
    assert set(grads) == set(c['grads'])
    for n, g in c['grads'].items():
        scale = float(g.abs().max())
        err = float((grads[n] - g).abs().max())
        assert err <= 2e-4 * scale + 2e-9, (n, err, scale)


@pytest.mark.parametrize('name', [n for n in SSEPT if '_eval' in n])
def test_ssept_eval_scores(name):
    c = load_case(name)
    b, z = c['batch'], c['z']
    att, cal, Ms = O.ssept_forward(c['params'], c['cfg'], b['item_seq'], b['item_len'], b['user'], c['rand'])
    assert rel(att, z['out_att']) < 2e-5 and rel(cal, z['out_cal']) < 2e-5
    for l, M in enumerate(Ms):
        assert rel(torch.sum((1 - M) ** 2), z['pen_sq.%d' % l]) < 1e-5
    sa, sc = O.ssept_full_sort_scores(c['params'], c['cfg'], b['item_seq'], b['item_len'], b['user'], c['rand'])
    assert rel(sc, z['scores']) < 2e-5 and rel(sa, z['scores_att']) < 2e-5
    _, idx = O.full_sort_topk(sc, c['k'])
    ok, nbad = O.topk_equal_modulo_ties(idx, torch.from_numpy(z['topk_idx']), torch.from_numpy(z['scores']))
    assert ok, nbad
    pa, pc = O.ssept_predict(c['params'], c['cfg'], b['item_seq'], b['item_len'], b['user'], b['pos'], c['rand'])
    assert rel(pa, z['predict_att']) < 2e-5 and rel(pc, z['predict_cal']) < 2e-5


@pytest.mark.parametrize('name', [n for n in TI if '_train' in n])
def test_ti_train_losses_and_routed_grads(name):
    """ACTiSASRec (actisasrec.py:173-193): time-interval aware keys / values on the transformer_layers.py layer"""
    c = load_case(name)
    b, z = c['batch'], c['z']
    l_att, l_cal, grads = O.ti_train_grads(c['params'], c['cfg'], b['item_seq'], b['item_len'], b['time'], b['pos'], c['rand'],
                                           neg_items=b.get('neg'))
    assert abs(float(l_att) - float(z['loss_att'])) < 1e-5 * abs(float(z['loss_att']))
    assert abs(float(l_cal) - float(z['loss_cal'])) < 1e-5 * abs(float(z['loss_cal']))
    assert set(grads) == set(c['grads'])
    for n, g in c['grads'].items():
        scale = float(g.abs().max())
        err = float((grads[n] - g).abs().max())
        assert err <= 2e-4 * scale + 2e-9, (n, err, scale)


@pytest.mark.parametrize('name', [n for n in TI if '_eval' in n])
def test_ti_eval_scores(name):
    c = load_case(name)
    b, z = c['batch'], c['z']
    att, cal, Ms = O.ti_forward(c['params'], c['cfg'], b['item_seq'], b['item_len'], b['time'], c['rand'])
    assert rel(att, z['out_att']) < 2e-5 and rel(cal, z['out_cal']) < 2e-5
    for l, M in enumerate(Ms):
        assert rel(torch.sum((1 - M) ** 2), z['pen_sq.%d' % l]) < 1e-5
    sa, sc = O.ti_full_sort_scores(c['params'], c['cfg'], b['item_seq'], b['item_len'], b['time'], c['rand'])
    assert rel(sc, z['scores']) < 2e-5 and rel(sa, z['scores_att']) < 2e-5
    _, idx = O.full_sort_topk(sc, c['k'])
    ok, nbad = O.topk_equal_modulo_ties(idx, torch.from_numpy(z['topk_idx']), torch.from_numpy(z['scores']))
    assert ok, nbad
    pa, pc = O.ti_predict(c['params'], c['cfg'], b['item_seq'], b['item_len'], b['time'], b['pos'], c['rand'])
    assert rel(pa, z['predict_att']) < 2e-5 and rel(pc, z['predict_cal']) < 2e-5
