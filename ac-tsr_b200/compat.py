"""Minimal host-side mirrors of the RecBole objects the hot path touches: ModelType,
Interaction (recbole/data/interaction.py:43-345), a YAML/dict Config with RecBole's priority
(cmd line > config_dict > config files > defaults; recbole/config/configurator.py:167-172) and
the SequentialRecommender base (recbole/model/abstract_recommender.py:25-143).
Same names, argument meaning and error behaviour; only what AC-SASRec needs.
"""
import os
import re
from enum import Enum
from logging import getLogger

import numpy as np
import torch
import torch.nn as nn
import yaml


class ModelType(Enum):
    GENERAL = 1
    SEQUENTIAL = 2
    CONTEXT = 3
    KNOWLEDGE = 4
    TRADITIONAL = 5
    DECISIONTREE = 6


class EvaluatorType(Enum):
    RANKING = 1
    VALUE = 2


class Interaction(object):
    """dict of equally long tensors; `interaction[field]`, `.to(device)`, slicing, `len`."""

    def __init__(self, interaction):
        self.interaction = dict()
        if not isinstance(interaction, dict):
            raise ValueError(f'[{type(interaction)}] is not supported for initialize `Interaction`!')
        for key, value in interaction.items():
            if isinstance(value, (list, np.ndarray)):
                self.interaction[key] = torch.as_tensor(np.asarray(value))
            elif isinstance(value, torch.Tensor):
                self.interaction[key] = value
            else:
                raise ValueError(f'The type of {key}[{type(value)}] is not supported!')
        self.length = -1
        for k in self.interaction:
            self.length = max(self.length, self.interaction[k].unsqueeze(-1).shape[0])

    def __iter__(self):
        return self.interaction.__iter__()

    def __getattr__(self, item):
        if 'interaction' not in self.__dict__:
            raise AttributeError("'Interaction' object has no attribute 'interaction'")
        if item in self.interaction:
            return self.interaction[item]
        raise AttributeError(f"'Interaction' object has no attribute '{item}'")

    def __getitem__(self, index):
        if isinstance(index, str):
            return self.interaction[index]
        return Interaction({k: v[index] for k, v in self.interaction.items()})

    def __setitem__(self, key, value):
        if not isinstance(key, str):
            raise KeyError(f'{type(key)} object does not support item assigment')
        self.interaction[key] = value

    def __contains__(self, item):
        return item in self.interaction

    def __len__(self):
        return self.length

    @property
    def columns(self):
        return list(self.interaction.keys())

    def to(self, device, selected_field=None):
        """interaction.py:174-200.  Pinned host tensors are copied with non_blocking=True."""
        ret = {}
        if isinstance(selected_field, str):
            selected_field = [selected_field]
        sel = set(selected_field) if selected_field is not None else None
        for k, v in self.interaction.items():
            if sel is None or k in sel:
                ret[k] = v.to(device, non_blocking=True)
            else:
                ret[k] = v
        return Interaction(ret)

    def cpu(self):
        return Interaction({k: v.cpu() for k, v in self.interaction.items()})

    def numpy(self):
        return {k: v.numpy() for k, v in self.interaction.items()}

    def shuffle(self):
        index = torch.randperm(self.length)
        for k in self.interaction:
            self.interaction[k] = self.interaction[k][index]

    def update(self, new_inter):
        for k in new_inter.interaction:
            self.interaction[k] = new_inter.interaction[k]

    def pack(self, fields=None, out=None, pin=True):
        """-> PackedInteraction: the int64 `fields` laid out back to back in ONE (pinned) buffer, each field a contiguous view of it,
        so that a batch crosses PCIe / NVLink-C2C as a single copy instead of one per field."""
        fields = list(fields) if fields is not None else [k for k, v in self.interaction.items() if v.dtype == torch.int64]
        return PackedInteraction.from_fields({k: self.interaction[k] for k in fields}, out=out, pin=pin)


class PackedInteraction(Interaction):
    """Interaction whose fields are views of one flat int64 buffer (`packed`); `layout` = [(field, offset, shape)]."""

    def __init__(self, packed, layout):
        self.packed, self.layout = packed, list(layout)
        super().__init__({k: packed[off:off + int(np.prod(shape))].view(shape) for k, off, shape in self.layout})

    @staticmethod
    def layout_of(fields):
        lay, off = [], 0
        for k, v in fields.items():
            if v.dtype != torch.int64:
                raise ValueError('packed batches hold int64 fields only (%s is %s)' % (k, v.dtype))
            lay.append((k, off, tuple(v.shape)))
            off += v.numel()
        return lay, off

    @classmethod
    def from_fields(cls, fields, out=None, pin=True):
        lay, total = cls.layout_of(fields)
        if out is None:
            out = torch.empty(total, dtype=torch.int64)
            if pin and torch.cuda.is_available():
                out = out.pin_memory()
        for k, off, shape in lay:
            out[off:off + fields[k].numel()].view(shape).copy_(fields[k])
        return cls(out[:total], lay)

    def to(self, device, selected_field=None):
        if selected_field is not None:
            return super().to(device, selected_field)
        return PackedInteraction(self.packed.to(device, non_blocking=True), self.layout)


# ---------------------------------------------------------------------------------------------
_DEFAULTS = dict(
    # overall.yaml / dataset defaults the hot path reads
    gpu_id=0, use_gpu=True, seed=2020, state='INFO', reproducibility=True, data_path='dataset/',
    checkpoint_dir='saved', show_progress=False, save_dataset=False, save_dataloaders=False,
    epochs=300, train_batch_size=2048, learner='adam', learning_rate=0.001, eval_step=1, stopping_step=10,
    clip_grad_norm=None, weight_decay=0.0, loss_decimal_place=4, reg_weight=None,
    eval_args={'split': {'RS': [0.8, 0.1, 0.1]}, 'group_by': 'user', 'order': 'RO', 'mode': 'full'},
    repeatable=False, metrics=['Recall', 'MRR', 'NDCG', 'Hit', 'Precision'], topk=[10], valid_metric='MRR@10',
    valid_metric_bigger=True, eval_batch_size=4096, metric_decimal_place=4,
    USER_ID_FIELD='user_id', ITEM_ID_FIELD='item_id', RATING_FIELD='rating', TIME_FIELD='timestamp',
    LIST_SUFFIX='_list', MAX_ITEM_LIST_LENGTH=50, ITEM_LIST_LENGTH_FIELD='item_length', NEG_PREFIX='neg_',
    POSITION_FIELD='position_id', load_col={'inter': ['user_id', 'item_id']}, field_separator='\t', seq_separator=' ',
    rm_dup_inter=None, val_interval=None, filter_inter_by_user_or_item=True, user_inter_num_interval=None,
    item_inter_num_interval=None, neg_sampling=None, benchmark_filename=None, seq_len=None,
    # B200 build additions (ignored by the reference)
    logits_passes=3, cuda_graph=True, fused_topk=True, fused_step=True,
)

_yaml_loader = yaml.FullLoader
_yaml_loader.add_implicit_resolver(      # configurator.py:90-104: floats like 1e-12
    u'tag:yaml.org,2002:float',
    re.compile(u'''^(?:
     [-+]?(?:[0-9][0-9_]*)\\.[0-9_]*(?:[eE][-+]?[0-9]+)?
    |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
    |\\.[0-9_]+(?:[eE][-+][0-9]+)?
    |[-+]?[0-9][0-9_]*(?::[0-5]?[0-9])+\\.[0-9_]*
    |[-+]?\\.(?:inf|Inf|INF)
    |\\.(?:nan|NaN|NAN))$''', re.X), list(u'-+0123456789.'))


class Config(object):
    """`Config(model, dataset, config_file_list, config_dict)`; dict-style access; missing key -> None
    (configurator.py:413-417)."""

    def __init__(self, model=None, dataset=None, config_file_list=None, config_dict=None, cmd_args=None):
        final = dict(_DEFAULTS)
        for f in config_file_list or []:
            with open(f, 'r', encoding='utf-8') as fh:
                final.update(yaml.load(fh.read(), Loader=_yaml_loader) or {})
        final.update(config_dict or {})
        final.update(self._parse_cmd(cmd_args))
        if model is not None:
            final['model'] = model if isinstance(model, str) else model.__name__
        if dataset is not None:
            final['dataset'] = dataset
        self.final_config_dict = final
        self._set_derived()

    @staticmethod
    def _parse_cmd(args):
        out = {}
        for a in args or []:
            if a.startswith('--') and '=' in a:
                k, v = a[2:].split('=', 1)
                try:
                    v = yaml.load(v, Loader=_yaml_loader)
                except Exception:
                    pass
                out[k] = v
        return out

    def _set_derived(self):
        c = self.final_config_dict
        c['MODEL_TYPE'] = ModelType.SEQUENTIAL
        c['eval_type'] = EvaluatorType.RANKING
        if isinstance(c.get('topk'), int):
            c['topk'] = [c['topk']]
        if isinstance(c.get('metrics'), str):
            c['metrics'] = [c['metrics']]
        if 'device' not in c or c['device'] is None:
            use = c.get('use_gpu', True) and torch.cuda.is_available()
            # configurator.py:344-348 exports gpu_id as CUDA_VISIBLE_DEVICES before CUDA is initialised; here CUDA may already
            # be up (one process per GPU sets its own device), so gpu_id selects the device index directly
            gpu_id = c.get('gpu_id', 0)
            try:
                gpu_id = int(str(gpu_id).split(',')[0])
            except (TypeError, ValueError):
                gpu_id = 0
            if use and gpu_id >= torch.cuda.device_count():
                raise ValueError('gpu_id %d: only %d CUDA device(s) visible' % (gpu_id, torch.cuda.device_count()))
            c['device'] = torch.device('cuda:%d' % gpu_id if use else 'cpu')
        c.setdefault('data_path', 'dataset/')
        if c.get('dataset') and not str(c['data_path']).rstrip('/').endswith(str(c['dataset'])):
            c['data_path'] = os.path.join(c['data_path'], c['dataset'])

    def __setitem__(self, key, value):
        if not isinstance(key, str):
            raise TypeError("index must be a str.")
        self.final_config_dict[key] = value

    def __getitem__(self, item):
        return self.final_config_dict.get(item, None)

    def __contains__(self, key):
        if not isinstance(key, str):
            raise TypeError("index must be a str.")
        return key in self.final_config_dict

    def get(self, key, default=None):
        return self.final_config_dict.get(key, default)

    def __str__(self):
        return '\n'.join('%s = %s' % kv for kv in sorted(self.final_config_dict.items(), key=lambda x: x[0]))

    __repr__ = __str__


def cfg_get(config, key, default=None):
    """config[key] that works for dict, Config and RecBole's Config (missing -> default)."""
    try:
        v = config[key]
    except (KeyError, TypeError):
        v = None
    return default if v is None else v


# ---------------------------------------------------------------------------------------------
class AbstractRecommender(nn.Module):
    """abstract_recommender.py:25-98."""

    def __init__(self):
        self.logger = getLogger()
        super().__init__()

    def calculate_loss(self, interaction):
        raise NotImplementedError

    def predict(self, interaction):
        raise NotImplementedError

    def full_sort_predict(self, interaction):
        raise NotImplementedError

    def other_parameter(self):
        if hasattr(self, 'other_parameter_name'):
            return {key: getattr(self, key) for key in self.other_parameter_name}
        return dict()

    def load_other_parameter(self, para):
        if para is None:
            return
        for key, value in para.items():
            setattr(self, key, value)

    def __str__(self):
        params = sum(int(np.prod(p.size())) for p in self.parameters() if p.requires_grad)
        return super().__str__() + '\nTrainable parameters' + f': {params}'


class SequentialRecommender(AbstractRecommender):
    """abstract_recommender.py:108-143."""
    type = ModelType.SEQUENTIAL

    def __init__(self, config, dataset):
        super().__init__()
        self.USER_ID = config['USER_ID_FIELD']
        self.ITEM_ID = config['ITEM_ID_FIELD']
        self.ITEM_SEQ = self.ITEM_ID + config['LIST_SUFFIX']
        self.ITEM_SEQ_LEN = config['ITEM_LIST_LENGTH_FIELD']
        self.POS_ITEM_ID = self.ITEM_ID
        self.NEG_ITEM_ID = config['NEG_PREFIX'] + self.ITEM_ID
        self.max_seq_length = config['MAX_ITEM_LIST_LENGTH']
        self.n_items = dataset.num(self.ITEM_ID)
        self.device = config['device']

    def gather_indexes(self, output, gather_index):
        """abstract_recommender.py:130-134 (index glue; the model's own path uses the fused K9 kernel)."""
        gather_index = gather_index.view(-1, 1, 1).expand(-1, -1, output.shape[-1])
        return output.gather(dim=1, index=gather_index).squeeze(1)

    def get_attention_mask(self, item_seq, bidirectional=False):
        """abstract_recommender.py:136-143.  Kept for API compatibility; the fused attention kernel
        derives the same mask from item_seq and never reads this tensor."""
        attention_mask = (item_seq != 0)
        ext = attention_mask.unsqueeze(1).unsqueeze(2)
        if not bidirectional:
            ext = torch.tril(ext.expand((-1, -1, item_seq.size(-1), -1)))
        return torch.where(ext, 0., -10000.)
