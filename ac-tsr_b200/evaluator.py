"""Top-k metrics on the [n_users, k+1] int matrix the collector builds (hit flags + pos_len):
mirrors recbole/evaluator/metrics.py:39-202, base_metric.py:43-80, evaluator.py:16-42.
Host side, numpy fp64, vectorised (the reference loops over users in MRR/NDCG)."""
from collections import OrderedDict

import numpy as np


def _hit(pos, pos_len):
    return (np.cumsum(pos, axis=1) > 0).astype(int)


def _mrr(pos, pos_len):
    idx = pos.argmax(axis=1)
    has = pos[np.arange(pos.shape[0]), idx] > 0
    rr = np.where(has, 1.0 / (idx + 1), 0.0)
    k = pos.shape[1]
    return np.where(np.arange(k)[None, :] >= idx[:, None], rr[:, None], 0.0)


def _recall(pos, pos_len):
    return np.cumsum(pos, axis=1) / pos_len.reshape(-1, 1)


def _precision(pos, pos_len):
    return pos.cumsum(axis=1) / np.arange(1, pos.shape[1] + 1)


def _ndcg(pos, pos_len):
    k = pos.shape[1]
    ranks = np.arange(1, k + 1, dtype=np.float64)
    disc = 1.0 / np.log2(ranks + 1)
    idcg_full = np.cumsum(disc)
    idcg_len = np.minimum(pos_len, k).astype(np.int64)
    col = np.minimum(np.arange(k)[None, :], idcg_len[:, None] - 1)      # idcg[row, idx:] = idcg[row, idx-1]
    idcg = idcg_full[col]
    dcg = np.cumsum(np.where(pos, disc[None, :], 0.0), axis=1)
    return dcg / idcg


def _map(pos, pos_len):
    k = pos.shape[1]
    pre = pos.cumsum(axis=1) / np.arange(1, k + 1)
    sum_pre = np.cumsum(pre * pos.astype(np.float64), axis=1)
    actual = np.minimum(pos_len, k).astype(np.int64)
    ranges = np.minimum(np.arange(1, k + 1)[None, :], actual[:, None])
    return sum_pre / ranges


METRICS = {'hit': _hit, 'mrr': _mrr, 'recall': _recall, 'ndcg': _ndcg, 'precision': _precision, 'map': _map}


class Evaluator(object):
    """evaluator.py:16-42: evaluate(rec_topk) -> OrderedDict('metric@k' -> rounded mean over users)."""

    def __init__(self, config):
        self.metrics = [m.lower() for m in config['metrics']]
        for m in self.metrics:
            if m not in METRICS:
                raise NotImplementedError('metric [%s] is outside the AC-SASRec hot path (top-k ranking metrics only)' % m)
        self.topk = list(config['topk'])
        self.decimal_place = config['metric_decimal_place'] if config['metric_decimal_place'] is not None else 4

    def evaluate(self, rec_topk):
        rec = np.asarray(rec_topk)
        pos = rec[:, :-1].astype(bool)
        pos_len = rec[:, -1].astype(np.int64)
        out = OrderedDict()
        for m in self.metrics:
            avg = METRICS[m](pos, pos_len).mean(axis=0)
            for k in self.topk:
                out['%s@%d' % (m, k)] = round(float(avg[k - 1]), self.decimal_place)
        return out
