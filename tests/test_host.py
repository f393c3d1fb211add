"""CPU-side checks: the C-ABI library loads and exports every symbol include/acsr.h declares, the
host mirror keeps the reference's names / shapes / error behaviour, and nothing silently falls
back to the CPU."""
import ctypes
import glob
import os

import numpy as np
import pytest
import torch

from oracle import acsr_oracle as O
from golden_util import GOLDEN_DIR, load_case

import ac_tsr_b200 as A


class DS:
    def __init__(self, n):
        self.n = self.item_num = n

    def num(self, field):
        return self.n


def cfg_for(**kw):
    d = O.default_cfg(**kw)
    d.update(device=torch.device('cpu'), seed=42, learning_rate=1e-3, epochs=1, eval_batch_size=8, train_batch_size=8,
             topk=[1, 5, 10], metrics=['Hit', 'MRR', 'NDCG', 'Recall'], valid_metric='Hit@10', checkpoint_dir='/tmp/acsr_ckpt')
    return A.Config(model='ACSASRec', config_dict=d)


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    protos = A._lib.parse_header()
    assert len(protos) >= 22
    dll = ctypes.CDLL(A._lib.LIB_PATH)
    for name in protos:
        assert hasattr(dll, name), name
    A.LIB.load()
    assert A.LIB.query('acsr_version') == 1
    assert A.LIB.query('acsr_num_sms') == 148
    # grid planning is host-side: persistent grid never exceeds the SM count for the benchmark shapes
    for M, V in ((256, 12102), (512, 12102), (512, 1000001), (256, 1683)):
        nc = A.LIB.query('acsr_logits_num_chunks', M, V)
        assert 1 <= nc and ((M + 127) // 128) * nc <= 148


def test_abi_rejects_bad_arguments_without_a_gpu():
    A.LIB.load()
    with pytest.raises(A.AcsrError, match='NULL'):
        A.LIB.call('acsr_rng_advance', None, None)
    with pytest.raises(A.AcsrError):
        A.LIB.call('acsr_topk_merge', None, None, 1, 1, 1, None, None, None, None, None)


def test_state_dict_names_match_reference_checkpoints():
    for name in ('c1_train', 'beauty_train', 'pos_tw_train', 'noorder_train', 'nodist_train', 'fixed_onelevel_train', 'relu_h8_eval'):
        c = load_case(name)
        cfg = cfg_for(**{k: c['cfg'][k] for k in c['cfg']})
        model = A.ACSASRec(cfg, DS(c['V']))
        own = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        ref = {k: tuple(v.shape) for k, v in c['params'].items()}
        assert own == ref, name
        model.load_state_dict(c['params'], strict=True)
    assert model.type == A.ModelType.SEQUENTIAL and 'Trainable parameters' in str(model)


def test_init_follows_reference_rules():
    torch.manual_seed(0)
    m = A.ACSASRec(cfg_for(n_layers=3), DS(1000))
    sd = m.state_dict()
    assert abs(float(sd['item_embedding.weight'].std()) - 0.02) < 2e-3
    assert float(sd['item_embedding.weight'][0].abs().sum()) > 0          # pad row is re-initialised (acsasrec.py:76-79)
    assert float(sd['trm_encoder.layer.0.attack_attention.query.bias'].abs().sum()) == 0
    assert float((sd['LayerNorm.weight'] - 1).abs().sum()) == 0
    # `scalar` is drawn once and deep-copied into every layer, never re-initialised (layers.py:640,1095)
    s = [float(sd['trm_encoder.layer.%d.attack_attention.scalar' % l]) for l in range(3)]
    assert s[0] == s[1] == s[2]
    assert not torch.equal(sd['trm_encoder.layer.0.attack_attention.query.weight'], sd['trm_encoder.layer.1.attack_attention.query.weight'])
    assert sd['trm_encoder.layer.0.gate.weight'].shape == (50, 64)


def test_constructor_errors_match_reference():
    with pytest.raises(ValueError, match='not a multiple'):
        A.ACSASRec(cfg_for(n_heads=3), DS(10))
    with pytest.raises(NotImplementedError):
        A.ACSASRec(cfg_for(loss_type='XX'), DS(10))
    with pytest.raises(KeyError):
        A.layers.FeedForward(64, 64, 0.5, 'nope', 1e-12)


def test_no_cpu_fallback():
    m = A.ACSASRec(cfg_for(), DS(100))
    seq, ln, pos = O.synth_batch(4, 50, 100)
    inter = A.Interaction({'item_id_list': seq, 'item_length': ln, 'item_id': pos})
    for fn in (m.calculate_loss, m.predict, m.full_sort_predict):
        with pytest.raises(A.AcsrError, match='CUDA'):
            fn(inter)


def test_config_priority_and_yaml_floats(tmp_path):
    f = tmp_path / 'a.yaml'
    f.write_text('layer_norm_eps: 1e-12\nn_layers: 3\ntopk: [1,3]\nlearning_rate: 0.0001\nmask_loss_weight: 0.03\n')
    c = A.Config(model='ACSASRec', dataset='x', config_file_list=[str(f)], config_dict={'n_layers': 4},
                 cmd_args=['--learning_rate=0.01'])
    assert c['layer_norm_eps'] == 1e-12 and isinstance(c['layer_norm_eps'], float)
    assert c['n_layers'] == 4 and c['learning_rate'] == 0.01 and c['topk'] == [1, 3]
    assert c['nonexistent'] is None and 'n_layers' in c and c['model'] == 'ACSASRec'
    ref_yaml = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'config', 'amazon-beauty.yaml')
    if os.path.exists(ref_yaml):
        c2 = A.Config(model='ACSASRec', config_file_list=[ref_yaml])
        assert c2['n_heads'] == 4 and c2['combine_option'] == 'gate' and c2['layer_norm_eps'] == 1e-12


def test_interaction_semantics():
    it = A.Interaction({'a': torch.arange(10), 'b': np.arange(20).reshape(10, 2)})
    assert len(it) == 10 and it['a'][3] == 3 and it.b.shape == (10, 2)
    sub = it[2:5]
    assert len(sub) == 3 and sub['b'][0, 0] == 4
    with pytest.raises(ValueError):
        A.Interaction({'a': 'x'})
    assert set(it.to('cpu').columns) == {'a', 'b'}


def test_evaluator_matches_oracle_metrics():
    rng = np.random.RandomState(1)
    k = 50
    pos = np.zeros((200, k), dtype=np.int32)
    for r in range(200):
        if rng.rand() < 0.6:
            pos[r, rng.randint(k)] = 1
    rec = np.concatenate([pos, np.ones((200, 1), dtype=np.int32)], 1)
    cfg = cfg_for()
    cfg['topk'] = [1, 3, 5, 10, 20, 50]
    cfg['metrics'] = ['Hit', 'MRR', 'NDCG', 'Recall']
    got = A.evaluator.Evaluator(cfg).evaluate(rec)
    want = O.topk_metrics(pos, np.ones(200, dtype=np.int64))
    assert got == want
    cfg['metrics'] = ['AUC']
    with pytest.raises(NotImplementedError):
        A.evaluator.Evaluator(cfg)


def test_loaders_shapes_and_eval_tuple():
    cfg = cfg_for()
    ds = A.data.SyntheticSequentialDataset(cfg, 21, 300, seed=3, pin=False)
    assert ds.num('item_id') == 300 and len(ds) == 21
    tl = A.data.TrainDataLoader(cfg, ds, shuffle=True)
    sizes = [len(b) for b in tl]
    assert sizes == [8, 8, 5] and len(tl) == 3
    b = next(iter(tl))
    assert b['item_id_list'].shape == (8, 50) and b['item_id_list'].dtype == torch.int64
    ln = b['item_length']
    assert ((b['item_id_list'] != 0).sum(1) == ln).all()                 # right-padded with 0
    el = A.data.FullSortEvalDataLoader(cfg, ds)
    inter, hist, pu, pi = next(iter(el))
    assert hist is None and torch.equal(pu, torch.arange(8)) and torch.equal(pi, inter['item_id'])


def test_early_stopping_rule():
    es = A.trainer.early_stopping
    assert es(0.5, 0.4, 3, 10) == (0.5, 0, False, True)
    assert es(0.3, 0.4, 10, 10) == (0.4, 11, True, False)
    assert es(0.3, 0.4, 0, 10, bigger=False) == (0.3, 0, False, True)
    assert A.trainer.is_attack_param('trm_encoder.layer.0.attack_attention.attack_key_transform.weight')
    assert not A.trainer.is_attack_param('trm_encoder.layer.0.attack_attention.key.weight')


def test_packed_batches_equal_sliced_batches():
    """loaders hand out the model's three fields as views of one buffer (a single host->device copy per batch); same values
    as slicing the dataset, ragged tail unpacked, layout stable across shuffles."""
    import ac_tsr_b200 as A
    cfg = A.Config(model='ACSASRec', config_dict=dict(train_batch_size=64, eval_batch_size=64, device=torch.device('cpu')))
    ds = A.data.SyntheticSequentialDataset(cfg, 200, 301, seed=3, pin=False)
    ev = A.data.FullSortEvalDataLoader(cfg, ds)
    batches = list(ev)
    assert len(batches) == 4
    for i, (inter, hist, pu, pi) in enumerate(batches):
        ref = ds.inter_feat[i * 64:(i + 1) * 64]
        for k in ('item_id_list', 'item_length', 'item_id'):
            assert torch.equal(inter[k], ref[k]) and inter[k].is_contiguous()
        assert pi is inter.interaction['item_id'] and hist is None
        if i < 3:
            assert isinstance(inter, A.compat.PackedInteraction) and inter.packed.numel() == 64 * 52
            assert [f for f, _, _ in inter.layout] == ['item_id_list', 'item_length', 'item_id']
            moved = inter.to(torch.device('cpu'))
            assert moved.layout == inter.layout and torch.equal(moved['item_id_list'], inter['item_id_list'])
        else:
            assert not isinstance(inter, A.compat.PackedInteraction) and len(inter) == 8
    tr = A.data.TrainDataLoader(cfg, ds, shuffle=True)
    torch.manual_seed(0)
    seen = torch.cat([b['item_id'] for b in tr])
    assert sorted(seen.tolist()) == sorted(ds.inter_feat['item_id'].tolist()) and len(seen) == 200
    first = next(iter(tr))
    assert torch.equal(first['item_id_list'], ds.inter_feat['item_id_list'][:64])      # packed view of the reshuffled rows
    tr.pr = 0
    with pytest.raises(ValueError):
        A.Interaction({'x': torch.zeros(3)}).pack(['x'])


def test_evaluator_matches_reference_evaluator_outputs():
    """metric@k values of the UNMODIFIED reference Evaluator on seeded rec.topk matrices (tests/golden/make_metrics_golden.py):
    Hit / MRR / NDCG / Recall / Precision / MAP, one or several positives per user, rounding to 4 places included."""
    import json
    gold = json.load(open(os.path.join(GOLDEN_DIR, 'metrics_golden.json')))
    for name, c in gold.items():
        cfg = {'metrics': ['Hit', 'MRR', 'NDCG', 'Recall', 'Precision', 'MAP'], 'topk': c['topk'], 'metric_decimal_place': 4}
        got = A.evaluator.Evaluator(cfg).evaluate(np.array(c['rec_topk']))
        assert set(got) == set(c['result']), name
        for k, v in c['result'].items():
            assert abs(got[k] - v) <= 1.0001e-4, (name, k, got[k], v)          # north_star: Recall@10 / NDCG@10 within 1e-4


def test_flat_adam_state_dict_is_torch_adam_format():
    """FlatAdam <-> torch.optim.Adam checkpoints (trainer.py:718-728 saves optimizer.state_dict(), :733-761 loads it):
    a reference-written 'optimizer' entry resumes here and ours resumes in the reference trainer."""
    torch.manual_seed(0)
    model = A.ACSASRec(cfg_for(n_layers=1, trainable_mask_loss_weight=True), DS(37))
    ref_model = A.ACSASRec(cfg_for(n_layers=1, trainable_mask_loss_weight=True), DS(37))
    ref_model.load_state_dict(model.state_dict())
    # a reference-side optimizer with real state: two torch Adam steps on random gradients (mask_loss_weight gets none)
    adam = torch.optim.Adam(ref_model.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(1)
    for _ in range(2):
        for n, p in ref_model.named_parameters():
            p.grad = None if n == 'mask_loss_weight' else torch.randn(p.shape, generator=g)
        adam.step()
    sd = adam.state_dict()
    flat = A.FlatAdam(model, lr=5e-4)
    flat.load_state_dict(sd)                                  # reference format in
    assert int(flat.step_count.item()) == 2 and flat.lr == 1e-3
    for i, p in enumerate(model.parameters()):
        if i in sd['state']:
            assert torch.equal(flat._view(flat.exp_avg, p), sd['state'][i]['exp_avg'])
            assert torch.equal(flat._view(flat.exp_avg_sq, p), sd['state'][i]['exp_avg_sq'])
        else:
            assert float(flat._view(flat.exp_avg, p).abs().max()) == 0.0
    out = flat.state_dict()                                   # reference format out: torch's own Adam accepts it
    adam2 = torch.optim.Adam(ref_model.parameters(), lr=1.0)
    adam2.load_state_dict(out)
    assert adam2.param_groups[0]['lr'] == 1e-3
    names = [n for n, _ in ref_model.named_parameters()]
    for i, p in enumerate(ref_model.parameters()):
        if names[i] == 'mask_loss_weight':
            continue
        assert torch.equal(adam2.state[p]['exp_avg'], adam.state[p]['exp_avg'])
        assert float(adam2.state[p]['step']) == 2.0
    # shape / count mismatches are errors, not silent mis-assignment
    bad = {'state': dict(sd['state']), 'param_groups': [dict(sd['param_groups'][0], params=sd['param_groups'][0]['params'][:-1])]}
    with pytest.raises(ValueError):
        flat.load_state_dict(bad)
    bad2 = {'state': {k: dict(v) for k, v in sd['state'].items()}, 'param_groups': sd['param_groups']}
    bad2['state'][sorted(bad2['state'])[0]]['exp_avg'] = torch.zeros(3)
    with pytest.raises(ValueError):
        flat.load_state_dict(bad2)
    with pytest.raises(ValueError):
        flat.load_state_dict({'something': 1})


def test_config_gpu_id_selects_the_device():
    c = cfg_for()
    assert c['device'].type == 'cpu'
    d = O.default_cfg()
    d.update(seed=42, learning_rate=1e-3, epochs=1, eval_batch_size=8, train_batch_size=8, topk=[10], metrics=['Hit'], gpu_id=0, use_gpu=False)
    assert A.Config(model='ACSASRec', config_dict=d)['device'].type == 'cpu'


def test_bert_masking_follows_the_reference_procedure():
    """AcBERT4Rec.reconstruct_train_data / reconstruct_test_data (acbert4rec.py:86-160): with the reference's random.seed the
    masked sequence, positives, negatives and masked indices are the ones the reference produced (tests/golden/bert_*_train)."""
    import random
    c = load_case('bert_fixed_train')
    z, b = c['z'], c['batch']
    cfg = dict(c['cfg'])
    cfg.update(device=torch.device('cpu'), seed=42, learning_rate=1e-3, epochs=1, eval_batch_size=8, train_batch_size=8,
               topk=[10], metrics=['Hit'], valid_metric='Hit@10', checkpoint_dir='/tmp/acsr_ckpt')
    model = A.AcBERT4Rec(A.Config(model='AcBERT4Rec', config_dict=cfg), DS(c['V']))
    assert model.mask_token == c['V'] and model.mask_item_length == int(cfg['mask_ratio'] * 50)
    random.seed(int(z['random_seed']))
    masked, pos, neg, index = model.reconstruct_train_data(b['item_seq'])
    assert torch.equal(masked, b['masked_seq']) and torch.equal(pos, b['pos_items'])
    assert torch.equal(neg, b['neg_items']) and torch.equal(index, b['masked_index'])
    seq = model.reconstruct_test_data(b['item_seq'].clone(), b['item_len'])
    assert seq.shape[1] == b['item_seq'].shape[1] + 1
    assert torch.equal(seq, O.bert_test_sequence(b['item_seq'], b['item_len'], c['V']))
    assert set(model.state_dict()) == set(c['params'])


def test_sibling_models_keep_reference_state_dicts_and_have_no_cpu_path():
    """ACSSEPT / ACTiSASRec: parameter names and shapes are those of the reference's checkpoints (goldens made by
    the real reference), get_model / get_trainer style lookup by name works, and CPU tensors raise instead of falling back"""
    from ac_tsr_b200.quick_start import _MODELS

    class DSU:
        def __init__(self, n_items, n_users):
            self.n_items, self.n_users = n_items, n_users

        def num(self, field):
            return self.n_users if field == 'user_id' else self.n_items
    for name in sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz'))):
        if not name.startswith(('ssept_', 'ti_')):
            continue
        c = load_case(name)
        cls_name = 'ACSSEPT' if c['ssept'] else 'ACTiSASRec'
        Model, Trainer = _MODELS[cls_name]
        assert Model.__name__ == cls_name and Trainer.__name__ == cls_name + 'Trainer'
        cfg = cfg_for(**{k: c['cfg'][k] for k in c['cfg']})
        cfg['TIME_FIELD'] = 'timestamp'
        model = Model(cfg, DSU(c['V'], int(c['z']['U']) if c['ssept'] else 0))
        own = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        ref = {k: tuple(v.shape) for k, v in c['params'].items()}
        assert own == ref, name
        model.load_state_dict(c['params'], strict=True)
        b = c['batch']
        f = {'item_id_list': b['item_seq'], 'item_length': b['item_len'], 'item_id': b['pos']}
        if c['ssept']:
            f['user_id'] = b['user']
        else:
            f['timestamp_list'] = b['time']
        with pytest.raises(A.AcsrError, match='CUDA'):
            model.full_sort_predict(A.Interaction(f))
    # the gate of the transformer_layers.py encoder is 50 wide whatever the sequence length (transformer_layers.py:891)
    enc = A.transformer_layers.AttackRTransformerEncoder(n_layers=1, hidden_size=32, inner_size=32, combine_option='gate')
    assert enc.layer[0].gate.out_features == 50 and enc.layer[0].plain_variant
    assert not A.layers.AttackRTransformerEncoder(n_layers=1, hidden_size=32, inner_size=32).layer[0].plain_variant


def test_backward_gradient_routing_helpers():
    """ops._wants_grad / _param_grad_buffers / direct_param_grads (host logic of the autograd path): a leaf whose requires_grad is
    off at backward time gets no buffer (the reference's routed double backward would drop that gradient, trainer.py:672-686);
    inside direct_param_grads a leaf that owns a .grad accumulates in place and nothing is returned to autograd; everything else
    shares one zero-filled block"""
    ops = A.ops
    a = torch.nn.Parameter(torch.ones(3, 5))
    b = torch.nn.Parameter(torch.ones(7))
    c = torch.ones(4, requires_grad=True) * 2.0            # non-leaf: always wanted
    a.grad, b.grad = torch.zeros_like(a), torch.zeros_like(b)
    assert ops._wants_grad(a) and ops._wants_grad(c) and not ops._wants_grad(None)
    b.requires_grad = False
    assert not ops._wants_grad(b)
    bufs, rets = ops._param_grad_buffers([a, b, None, c])
    assert bufs[1] is None and rets[1] is None and bufs[2] is None
    assert bufs[0] is rets[0] and bufs[3] is rets[3] and bufs[0].shape == a.shape and bufs[3].shape == c.shape
    assert float(bufs[0].abs().sum()) == 0.0 and bufs[0].data_ptr() != a.grad.data_ptr()
    assert bufs[0].untyped_storage().data_ptr() == bufs[3].untyped_storage().data_ptr()       # one shared block
    assert bufs[3].data_ptr() % 16 == 0                                                        # 4-float aligned slots
    with ops.direct_param_grads():
        bufs, rets = ops._param_grad_buffers([a, b, None, c])
        assert bufs[0] is a.grad and rets[0] is None           # in place, nothing for autograd to add
        assert bufs[1] is None and bufs[3] is rets[3]
        (bW, bb), (dW, db) = ops._linear_grad_buffers(a, None)
        assert bW is a.grad and dW is None and bb is None
        b.requires_grad = True
        (bW, bb), (dW, db) = ops._linear_grad_buffers(a, b)
        assert bW is a.grad and bb is b.grad and dW is None and db is None
    assert ops._grad_sink(a) is None                           # outside the context gradients go back to autograd
    (bW, bb), (dW, db) = ops._linear_grad_buffers(a, b)
    assert bW is dW and bb is db and float(dW.abs().sum()) == 0.0


def test_model_fields_and_bert_prepare_batch():
    """the fields a loader ships per model, and AcBERT4Rec.prepare_batch (the host half of its graphed step) == one
    reconstruct_train_data call on the same python `random` stream"""
    import random
    from ac_tsr_b200.data import _model_fields
    cfg = cfg_for()
    assert _model_fields(cfg) == ['item_id_list', 'item_length', 'item_id']
    cfg['model'] = 'ACSSEPT'
    assert _model_fields(cfg)[-1] == 'user_id'
    cfg['model'] = 'ACTiSASRec'
    cfg['TIME_FIELD'] = 'timestamp'
    assert _model_fields(cfg)[-1] == 'timestamp_list'
    ds = A.data.SyntheticSequentialDataset(cfg, 16, 50, seed=3, pin=False)
    ts = ds.inter_feat['timestamp_list']
    assert ts.dtype == torch.float32 and bool((ts[ds.inter_feat['item_id_list'] == 0] == 0).all())
    loader = A.data.TrainDataLoader(cfg, ds, shuffle=False, batch_size=8)
    batch = next(iter(loader))                                  # a float field: the batch is not packed into the int64 buffer
    assert batch['timestamp_list'].dtype == torch.float32 and batch['item_id_list'].shape == (8, 50)
    c = load_case('bert_fixed_train')
    bcfg = cfg_for(**{k: c['cfg'][k] for k in c['cfg']})
    model = A.AcBERT4Rec(bcfg, DS(c['V']))
    b = c['batch']
    inter = A.Interaction({'item_id_list': b['item_seq'], 'item_length': b['item_len'], 'item_id': b['pos']})
    random.seed(5)
    want = model.reconstruct_train_data(b['item_seq'])
    random.seed(5)
    prepared = model.prepare_batch(inter)
    for k, w in zip(model.EXTRA_FIELDS, want):
        assert torch.equal(prepared[k], w), k
    assert model.GRAPH_SAFE_STEP and all(k in prepared for k in ('item_id_list', 'item_length', 'item_id'))


@pytest.mark.parametrize('name,yaml_file,fields', [
    ('ACSSEPT', 'ml-100k-acssept.yaml', ['item_id_list', 'item_length', 'item_id', 'user_id']),
    ('ACTiSASRec', 'ml-100k-actisasrec.yaml', ['item_id_list', 'item_length', 'item_id', 'timestamp_list']),
    ('AcBERT4Rec', 'ml-100k-acbert4rec.yaml', ['item_id_list', 'item_length', 'item_id']),
])
def test_sibling_run_recbole_pipeline_up_to_the_first_batch(name, yaml_file, fields):
    """run_recbole.py --model=<sibling> --config_files=config/<yaml>: Config -> dataset -> loaders -> model, everything short of the
    first kernel (no GPU here): the batches carry the fields the model reads, the models take their sizes from the dataset, and
    only AC-SASRec trains from the HBM-resident loader"""
    from ac_tsr_b200.dataset import device_resident_training
    from ac_tsr_b200.quick_start import get_model, get_trainer
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    config = A.Config(model=name, dataset='ml-100k', config_file_list=[os.path.join(root, 'config', yaml_file)],
                      config_dict={'data_path': os.path.join(root, 'tests', 'golden') + '/', 'device': torch.device('cpu')})
    assert config['model'] == name
    ds = A.create_dataset(config)
    train, valid, test = A.data_preparation(config, ds)
    assert type(train).__name__ == 'TrainDataLoader'
    batch = next(iter(train))
    for f in fields:
        assert f in batch and len(batch[f]) == config['train_batch_size'], f
    ev, hist, pu, pi = next(iter(valid))
    for f in fields:
        assert f in ev, f
    model = get_model(name)(config, train.dataset)
    assert get_trainer(None, name).__name__ == name + 'Trainer'
    assert model.n_items == ds.item_num
    if name == 'ACSSEPT':
        assert model.n_users == ds.user_num and model.hidden_size == 64
    if name == 'ACTiSASRec':
        assert batch['timestamp_list'].dtype == torch.float32 and model.time_matrix_emb_K_embedding.num_embeddings == 257
    # loader choice on a GPU box
    config['device'] = torch.device('cuda')
    assert not device_resident_training(config)
    config['model'] = 'ACSASRec'
    assert device_resident_training(config)
    config['device_resident_data'] = False
    assert not device_resident_training(config)
