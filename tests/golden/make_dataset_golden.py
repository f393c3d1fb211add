"""Run the UNMODIFIED reference dataset pipeline (create_dataset + data_preparation) on ml-100k in the build container and
record digests of the train / valid / test tensors the model reads: tests/golden/ml100k_dataset.json.
    python tests/golden/make_dataset_golden.py
Nothing is written into /root/reference (the .inter file is read from tests/golden/ml-100k through a temp copy)."""
import gzip
import hashlib
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference            # noqa: E402  (stubs colorlog / colorama / thop, np.float)


def digest(t):
    a = t.contiguous().numpy()
    return {'shape': list(a.shape), 'dtype': str(a.dtype), 'sha256': hashlib.sha256(a.tobytes()).hexdigest()}


def main():
    import_reference()
    from recbole.config import Config
    from recbole.data import create_dataset, data_preparation
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, 'ml-100k'))
    with gzip.open(os.path.join(HERE, 'ml-100k', 'ml-100k.inter.gz'), 'rb') as src, open(os.path.join(tmp, 'ml-100k', 'ml-100k.inter'), 'wb') as dst:
        shutil.copyfileobj(src, dst)
    out = {}
    cases = {
        'plain': dict(),
        'kcore_dedup': dict(rm_dup_inter='first', user_inter_num_interval='[30,inf)', item_inter_num_interval='[20,inf)',
                            val_interval={'rating': '[3,inf)'}, MAX_ITEM_LIST_LENGTH=20),
    }
    cases['window5_valid_only'] = dict(MAX_ITEM_LIST_LENGTH=5, eval_args={'split': {'LS': 'valid_only'}, 'group_by': 'user', 'order': 'TO',
                                                                            'mode': 'full'})
    cases['dedup_last_test_only'] = dict(rm_dup_inter='last', item_inter_num_interval='[50,300]', MAX_ITEM_LIST_LENGTH=10,
                                         eval_args={'split': {'LS': 'test_only'}, 'group_by': 'user', 'order': 'TO', 'mode': 'full'})
    for name, extra in cases.items():
        cd = dict(data_path=tmp + '/', load_col={'inter': ['user_id', 'item_id', 'rating', 'timestamp']},
                  eval_args={'split': {'LS': 'valid_and_test'}, 'group_by': 'user', 'order': 'TO', 'mode': 'full'},
                  neg_sampling=None, use_gpu=False, n_layers=2, n_heads=2, hidden_size=64, inner_size=256, hidden_dropout_prob=0.5,
                  attn_dropout_prob=0.5, hidden_act='gelu', layer_norm_eps=1e-12, initializer_range=0.02, loss_type='CE',
                  train_batch_size=256, eval_batch_size=256, save_dataset=False, save_dataloaders=False)
        cd.update(extra)
        config = Config(model='ACSASRec', dataset='ml-100k', config_dict=cd)
        dataset = create_dataset(config)
        train, valid, test = dataset.build()
        rec = {'user_num': int(dataset.user_num), 'item_num': int(dataset.item_num), 'config': extra}
        for part, ds in (('train', train), ('valid', valid), ('test', test)):
            for f in ('user_id', 'item_id', 'item_length', 'item_id_list', 'timestamp_list'):     # timestamp_list: read by ACTiSASRec only
                rec['%s.%s' % (part, f)] = digest(ds.inter_feat[f])
        rec['item_token_sha256'] = hashlib.sha256('\n'.join(dataset.field2id_token['item_id']).encode()).hexdigest()
        out[name] = rec
        print(name, rec['user_num'], rec['item_num'], rec['train.item_id']['shape'])
    json.dump(out, open(os.path.join(HERE, 'ml100k_dataset.json'), 'w'), indent=1)
    shutil.rmtree(tmp)


if __name__ == '__main__':
    main()
