"""Host-side mirror of the reference's AttackR* layers (recbole/model/layers.py:614-798, 859-951,
1070-1131): same class names, constructor arguments, parameter names / shapes (so reference
checkpoints load unchanged) and error behaviour.  The arithmetic does not run in PyTorch: each
layer issues a handful of GEMMs (cuBLAS, plain library GEMMs) and the fused sm_100a kernels of
libacsr.so through ops.py.  The [B,H,L,L] intermediates of the reference are never materialised
unless a caller asks for the attention probabilities.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


class AttackMask:
    """What the encoder hands back per layer in place of the reference's [B,H,L,L] attack mask:
    pen_sq = sum((1 - M)^2) (differentiable; acsasrec.py:135 takes its sqrt) and, only when
    attention probabilities were requested, the detached mask itself."""

    def __init__(self, pen_sq, mask=None):
        self.pen_sq = pen_sq
        self.mask = mask

    def penalty(self):
        return torch.sqrt(self.pen_sq[0])


class Runtime:
    """Per-call execution context threaded from the model down to the kernels."""

    def __init__(self, rng=None, rand=None, attacked_last_only=False, bidirectional=False):
        self.rng = rng                  # ops.DeviceRng or None
        self.rand = rand                # explicit masks/noise {key: tensor} (parity tests) or None
        self.attacked_last_only = attacked_last_only
        self.bidirectional = bidirectional      # get_attention_mask(bidirectional=True) of AcBERT4Rec (abstract_recommender.py:136-143)

    def mask(self, key):
        return None if self.rand is None else self.rand.get(key)


_DEFAULT_RNG = {}


def default_runtime(device):
    """Runtime for callers that use a layer stand-alone (no model): Philox state seeded from torch's seed."""
    key = str(device)
    if key not in _DEFAULT_RNG:
        _DEFAULT_RNG[key] = ops.DeviceRng(torch.initial_seed(), device)
    _DEFAULT_RNG[key].advance()
    return Runtime(rng=_DEFAULT_RNG[key])


def _stream_base(layer_idx):
    return 16 * (layer_idx + 1)


class FeedForward(nn.Module):
    """layers.py:745-798.  dense_1 -> act -> dense_2 -> dropout -> LN(. + input)."""

    def __init__(self, hidden_size, inner_size, hidden_dropout_prob, hidden_act, layer_norm_eps):
        super().__init__()
        self.dense_1 = nn.Linear(hidden_size, inner_size)
        if hidden_act not in ops.ACT_IDS:
            raise KeyError(hidden_act)
        self.hidden_act = hidden_act
        self.dense_2 = nn.Linear(inner_size, hidden_size)
        self.LayerNorm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)
        self.dropout = nn.Dropout(hidden_dropout_prob)

    def forward(self, input_tensor, rt=None, mask_key=None, rng_stream=0):
        rt = rt or default_runtime(input_tensor.device)
        p = self.dropout.p if self.training else 0.0
        z = ops.linear(input_tensor, self.dense_1.weight)
        z = ops.BiasActFn.apply(z, self.dense_1.bias, ops.ACT_IDS[self.hidden_act])
        z = ops.linear(z, self.dense_2.weight)
        return ops.BiasDropoutResLnFn.apply(z, self.dense_2.bias, input_tensor, self.LayerNorm.weight, self.LayerNorm.bias,
                                            self.LayerNorm.eps, p, rt.mask(mask_key) if p > 0 else None, rt.rng, rng_stream)


class AttackRMultiHeadAttention(nn.Module):
    """Parameter container + projections of layers.py:614-742; the attention itself is the fused kernel."""

    def __init__(self, n_heads, hidden_size, hidden_dropout_prob, attn_dropout_prob, layer_norm_eps, use_order, use_distance):
        super().__init__()
        if hidden_size % n_heads != 0:
            raise ValueError(
                "The hidden size (%d) is not a multiple of the number of attention "
                "heads (%d)" % (hidden_size, n_heads))
        self.num_attention_heads = n_heads
        self.attention_head_size = int(hidden_size / n_heads)
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.sqrt_attention_head_size = math.sqrt(self.attention_head_size)
        self.query = nn.Linear(hidden_size, self.all_head_size)
        self.key = nn.Linear(hidden_size, self.all_head_size)
        self.value = nn.Linear(hidden_size, self.all_head_size)
        self.use_order = use_order
        self.use_distance = use_distance
        if self.use_order:
            self.order_affine = nn.Linear(2 * self.attention_head_size, 1)
        if self.use_distance:
            self.distance_affine = nn.Linear(2 * self.attention_head_size, 1)
            self.scalar = nn.Parameter(torch.randn(1))
        self.attack_query_transform = nn.Linear(self.all_head_size, self.all_head_size)
        self.attack_key_transform = nn.Linear(self.all_head_size, self.all_head_size)
        self.attn_dropout = nn.Dropout(attn_dropout_prob)
        self.dense = nn.Linear(hidden_size, hidden_size)
        self.LayerNorm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)
        self.out_dropout = nn.Dropout(hidden_dropout_prob)

    def cal_adjusted_outputs(self, context_layer, input_tensor, rt=None, mask_key=None, rng_stream=0):
        """layers.py:676-684 after probs.V: dense -> dropout -> LN(. + input)."""
        rt = rt or default_runtime(input_tensor.device)
        p = self.out_dropout.p if self.training else 0.0
        h = ops.linear(context_layer, self.dense.weight)
        return ops.BiasDropoutResLnFn.apply(h, self.dense.bias, input_tensor, self.LayerNorm.weight, self.LayerNorm.bias,
                                            self.LayerNorm.eps, p, rt.mask(mask_key) if p > 0 else None, rt.rng, rng_stream)


def key_ids_from_mask(attention_mask):
    """Accept the reference's additive mask [B,1,L,L] (0 / -10000, causal + padding) or an
    item_seq-like [B,L] id tensor; return int64 [B,L] with 0 == padded key."""
    if attention_mask.dim() == 2:
        return attention_mask if attention_mask.dtype == torch.int64 else attention_mask.to(torch.int64)
    if attention_mask.dim() != 4:
        raise ValueError('attention_mask must be [B,1,L,L], [B,1,1,L] or [B,L]')
    return (attention_mask[:, 0, -1, :] == 0).to(torch.int64)


class AttackRTransformerLayer(nn.Module):
    """layers.py:859-951."""

    plain_variant = False       # True in transformer_layers.py's subclass: no re-normalising softmaxes (ACSR_ATTN_PLAIN)

    def __init__(self, n_heads, hidden_size, intermediate_size, hidden_dropout_prob, attn_dropout_prob, hidden_act,
                 layer_norm_eps, combine_option='fixed', use_order=True, use_distance=True, two_level=True,
                 rich_calibrated_combine='fixed', seq_length=50):
        super().__init__()
        self.hidden_size = hidden_size
        self.attack_attention = AttackRMultiHeadAttention(n_heads, hidden_size, hidden_dropout_prob, attn_dropout_prob,
                                                          layer_norm_eps, use_order=use_order, use_distance=use_distance)
        self.two_level = two_level
        self.rich_calibrated_combine = rich_calibrated_combine
        if self.rich_calibrated_combine == 'trainable':
            self.rich_calibrated_combine_ratio = torch.nn.Parameter(torch.FloatTensor([0.5]), requires_grad=True)
        self.combine_option = combine_option
        if self.combine_option == 'gate':
            self.gate = torch.nn.Linear(hidden_size, seq_length)
        self.combine_ratio = 0.5
        self.feed_forward = FeedForward(hidden_size, intermediate_size, hidden_dropout_prob, hidden_act, layer_norm_eps)
        self.anneal_step = 0

    def _dense_pair(self, ctx_cal, ctx_att, x, rt, layer_idx, base):
        """cal_adjusted_outputs + feed_forward (layers.py:676-684, 790-798) of the calibrated and the attacked stream, the two
        streams sharing each GEMM launch; same dropout keys / Philox streams as the one-stream methods -> (cal_out, att_out)"""
        aa, ff = self.attack_attention, self.feed_forward
        training = self.training
        p_out = aa.out_dropout.p if training else 0.0
        p_ff = ff.dropout.p if training else 0.0
        act = ops.ACT_IDS[ff.hidden_act]

        def bdrl(h, bias, res, ln, p, key, stream):
            return ops.BiasDropoutResLnFn.apply(h, bias, res, ln.weight, ln.bias, ln.eps, p, rt.mask((layer_idx, key)) if p > 0 else None,
                                                rt.rng, stream)
        h_cal, h_att = ops.multi_linear([ctx_cal, ctx_att], [aa.dense.weight, aa.dense.weight])
        a_cal = bdrl(h_cal, aa.dense.bias, x, aa.LayerNorm, p_out, 'D5', base + 3)
        a_att = bdrl(h_att, aa.dense.bias, x, aa.LayerNorm, p_out, 'D4', base + 2)
        z_cal, z_att = ops.multi_linear([a_cal, a_att], [ff.dense_1.weight, ff.dense_1.weight])
        z_cal = ops.BiasActFn.apply(z_cal, ff.dense_1.bias, act)
        z_att = ops.BiasActFn.apply(z_att, ff.dense_1.bias, act)
        o_cal, o_att = ops.multi_linear([z_cal, z_att], [ff.dense_2.weight, ff.dense_2.weight])
        return (bdrl(o_cal, ff.dense_2.bias, a_cal, ff.LayerNorm, p_ff, 'D7', base + 5),
                bdrl(o_att, ff.dense_2.bias, a_att, ff.LayerNorm, p_ff, 'D6', base + 4))

    def forward(self, hidden_states, attention_mask, return_attention_prob=False, return_all_attention_prob=False,
                rt=None, layer_idx=0, need_attacked=True, time_terms=None):
        rt = rt or default_runtime(hidden_states.device)
        if self.combine_option not in ops.COMBINE_IDS:
            raise KeyError(self.combine_option)
        if not self.two_level and self.rich_calibrated_combine not in ('fixed', 'trainable'):
            raise KeyError(self.rich_calibrated_combine)
        aa = self.attack_attention
        x = hidden_states
        B, L, d = x.shape
        key_ids = key_ids_from_mask(attention_mask)
        # the three projections of x are one launch, the attack pair (+ the gate) of mixed_q / mixed_k another (ops.MultiLinearFn)
        mq, mk, mv = ops.multi_linear([x, x, x], [aa.query.weight, aa.key.weight, aa.value.weight],
                                      [aa.query.bias, aa.key.bias, aa.value.bias])
        gate_logit, comb_scalar = None, 0.0
        if self.combine_option == 'gate':
            if self.gate.out_features != L:
                raise ValueError('gate width %d != sequence length %d (layers.py:878/887)' % (self.gate.out_features, L))
            aq, ak, gate_logit = ops.multi_linear(
                [mq, mk, mq], [aa.attack_query_transform.weight, aa.attack_key_transform.weight, self.gate.weight],
                [aa.attack_query_transform.bias, aa.attack_key_transform.bias, self.gate.bias])
        else:
            aq, ak = ops.multi_linear([mq, mk], [aa.attack_query_transform.weight, aa.attack_key_transform.weight],
                                      [aa.attack_query_transform.bias, aa.attack_key_transform.bias])
        if self.combine_option == 'annealing':
            comb_scalar = math.exp(-self.anneal_step / 100000)      # layers.py:889-891
            self.anneal_step += 1
        base = _stream_base(layer_idx)
        training = self.training
        p_attn = aa.attn_dropout.p if training else 0.0
        rand = None
        if rt.rand is not None:
            rand = {k: rt.rand.get((layer_idx, k)) for k in ('D1', 'D2', 'D3', 'noise')}
            if p_attn == 0.0:
                rand['D1'] = rand['D2'] = rand['D3'] = None
        want_probs = bool(return_attention_prob or return_all_attention_prob)
        opts = ops.AttnOpts(aa.num_attention_heads, self.two_level, self.combine_option,
                            self.rich_calibrated_combine if not self.two_level else 'none', p_attn,
                            bidirectional=bool(getattr(rt, 'bidirectional', False)), plain=self.plain_variant)
        calib_args = (
            mq, mk, mv, aq, ak, gate_logit, key_ids,
            aa.order_affine.weight if aa.use_order else None, aa.order_affine.bias if aa.use_order else None,
            aa.distance_affine.weight if aa.use_distance else None, aa.distance_affine.bias if aa.use_distance else None,
            aa.scalar if aa.use_distance else None,
            getattr(self, 'rich_calibrated_combine_ratio', None) if not self.two_level else None,
            opts, comb_scalar, p_attn, rand, rt.rng, base, need_attacked)
        if time_terms is None:
            ctx_att, ctx_cal, pen_sq, probs = ops.AttnCalibFn.apply(*calib_args, want_probs)
        else:
            # ACTiSASRec (transformer_layers.py:1116-1134, 1085-1091): q.posK + q.timeK[t_ij] join the raw scores, and
            # probs.posV + probs.timeV[t_ij] the context of both streams
            if want_probs:
                raise NotImplementedError('attention probabilities of the time-aware layer are not exported')
            s_bias = time_terms.score_bias(mq)
            ctx_att, ctx_cal, pen_sq, prob_att, prob_cal = ops.AttnCalibTiFn.apply(s_bias, *calib_args)
            ctx_cal = ctx_cal + time_terms.context(prob_cal)
            if need_attacked:
                ctx_att = ctx_att + time_terms.context(prob_att)
            probs = None
        att_out = None
        if need_attacked:
            # both streams go through the same out-projection / feed-forward weights: one launch per GEMM for the pair
            cal_out, att_out = self._dense_pair(ctx_cal, ctx_att, x, rt, layer_idx, base)
        else:
            cal_att_out = aa.cal_adjusted_outputs(ctx_cal, x, rt, (layer_idx, 'D5'), base + 3)
            cal_out = self.feed_forward(cal_att_out, rt, (layer_idx, 'D7'), base + 5)
        attack_mask = AttackMask(pen_sq, probs[2] if probs is not None else None)
        combined = probs[5] if probs is not None else None
        if return_all_attention_prob:
            all_prob = {'before_spatial': probs[0], 'after_spatial': probs[1], 'perturbed_mask': probs[2],
                        'perturbed_attention': probs[3], 'calibrated_attention': probs[5]}
            return att_out, cal_out, attack_mask, combined, all_prob
        return att_out, cal_out, attack_mask, combined


class AttackRTransformerEncoder(nn.Module):
    """layers.py:1070-1131: n identical layers chained on the calibrated stream."""

    layer_class = AttackRTransformerLayer

    def __init__(self, n_layers=2, n_heads=2, hidden_size=64, inner_size=256, hidden_dropout_prob=0.5,
                 attn_dropout_prob=0.5, hidden_act='gelu', layer_norm_eps=1e-12, combine_option='fixed', use_order=True,
                 use_distance=True, two_level=True, rich_calibrated_combine='fixed', seq_length=50):
        super().__init__()
        self.layer = nn.ModuleList([
            self.layer_class(n_heads, hidden_size, inner_size, hidden_dropout_prob, attn_dropout_prob, hidden_act,
                                    layer_norm_eps, combine_option, use_order=use_order, use_distance=use_distance,
                                    two_level=two_level, rich_calibrated_combine=rich_calibrated_combine,
                                    seq_length=seq_length)
            for _ in range(n_layers)])
        # the reference deep-copies ONE layer (layers.py:1095): all layers start from identical weights
        for l in self.layer[1:]:
            l.load_state_dict(self.layer[0].state_dict())

    def forward(self, hidden_states, attention_mask, output_all_encoded_layers=True, return_attention_prob=False,
                return_all_attention_prob=False, rt=None):
        rt = rt or default_runtime(hidden_states.device)
        all_encoder_layers, all_attack_masks = [], []
        all_attention_prob = [] if return_attention_prob else None
        all_probs = [] if return_all_attention_prob else None
        att = cal = None
        n = len(self.layer)
        for layer_idx, layer_module in enumerate(self.layer):
            need_att = (layer_idx == n - 1) or not rt.attacked_last_only
            res = layer_module(hidden_states, attention_mask, return_attention_prob, return_all_attention_prob,
                               rt=rt, layer_idx=layer_idx, need_attacked=need_att)
            att, cal, attack_mask, combined = res[:4]
            hidden_states = cal
            all_attack_masks.append(attack_mask)
            if output_all_encoded_layers:
                all_encoder_layers.append((att, cal))
            if return_attention_prob:
                all_attention_prob.append(combined)
            if return_all_attention_prob:
                all_probs.append(res[4])
        if not output_all_encoded_layers:
            all_encoder_layers.append((att, cal))
        if return_all_attention_prob:
            return all_encoder_layers, all_attack_masks, all_probs
        if return_attention_prob:
            return all_encoder_layers, all_attack_masks, all_attention_prob
        return all_encoder_layers, all_attack_masks
