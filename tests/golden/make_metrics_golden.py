"""Run the UNMODIFIED reference evaluator (recbole/evaluator/{evaluator,metrics,collector}.py) on seeded `rec.topk` matrices and
record its output: tests/golden/metrics_golden.json.   python tests/golden/make_metrics_golden.py"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference            # noqa: E402


def main():
    import_reference()
    from recbole.evaluator import Evaluator
    from recbole.evaluator.collector import DataStruct
    out = {}
    for name, n, k, seed, pmax in (('one_pos', 943, 50, 1, 1), ('multi_pos', 500, 20, 2, 7), ('tiny', 3, 10, 3, 2)):
        g = torch.Generator().manual_seed(seed)
        pos_len = torch.randint(1, pmax + 1, (n, 1), generator=g)
        flags = (torch.rand(n, k, generator=g) < 0.08).int()
        # no more hits than positives per user
        over = flags.cumsum(1) > pos_len
        flags[over] = 0
        rec = torch.cat((flags, pos_len.int()), dim=1)
        topk = [1, 3, 5, 10, k] if k >= 10 else [1, k]
        config = {'metrics': ['Hit', 'MRR', 'NDCG', 'Recall', 'Precision', 'MAP'], 'topk': topk, 'metric_decimal_place': 4}
        ds = DataStruct()
        ds.set('rec.topk', rec)
        res = Evaluator(config).evaluate(ds)
        out[name] = {'rec_topk': rec.tolist(), 'topk': topk, 'result': {k_: float(v) for k_, v in res.items()}}
        print(name, list(res.items())[:4])
    json.dump(out, open(os.path.join(HERE, 'metrics_golden.json'), 'w'))


if __name__ == '__main__':
    main()
