"""Host-side mirror of the AttackR* classes of recbole/model/transformer_layers.py:742-1007 -- the encoder ACSSEPT builds
(acssept.py:15, 61-75).  It has the same parameters, projections, calibrators and combine options as the layers.py encoder
(layers.py:614-951) but a different attention: the attacked, calibrated and combined attention matrices are used as computed
(transformer_layers.py:919-927) instead of being pushed through one more masked softmax each (layers.py:917-925), and the gate
is always Linear(hidden, 50) (transformer_layers.py:891).  On the device that is the same fused kernel with ACSR_ATTN_PLAIN.
"""
from .layers import AttackRMultiHeadAttention, FeedForward, Runtime          # noqa: F401  (same classes in both reference files)
from . import layers as _layers


class AttackRTransformerLayer(_layers.AttackRTransformerLayer):
    """transformer_layers.py:873-953."""

    plain_variant = True

    def __init__(self, n_heads, hidden_size, intermediate_size, hidden_dropout_prob, attn_dropout_prob, hidden_act,
                 layer_norm_eps, combine_option='fixed', use_order=True, use_distance=True, two_level=True,
                 rich_calibrated_combine='fixed', seq_length=50):
        super().__init__(n_heads, hidden_size, intermediate_size, hidden_dropout_prob, attn_dropout_prob, hidden_act,
                         layer_norm_eps, combine_option, use_order=use_order, use_distance=use_distance, two_level=two_level,
                         rich_calibrated_combine=rich_calibrated_combine, seq_length=seq_length)


class AttackRTransformerEncoder(_layers.AttackRTransformerEncoder):
    """transformer_layers.py:955-1007 (no seq_length argument: the gate is 50 wide, transformer_layers.py:891)."""

    layer_class = AttackRTransformerLayer

    def __init__(self, n_layers=2, n_heads=2, hidden_size=64, inner_size=256, hidden_dropout_prob=0.5, attn_dropout_prob=0.5,
                 hidden_act='gelu', layer_norm_eps=1e-12, combine_option='fixed', use_order=True, use_distance=True,
                 two_level=True, rich_calibrated_combine='fixed'):
        super().__init__(n_layers, n_heads, hidden_size, inner_size, hidden_dropout_prob, attn_dropout_prob, hidden_act,
                         layer_norm_eps, combine_option, use_order=use_order, use_distance=use_distance, two_level=two_level,
                         rich_calibrated_combine=rich_calibrated_combine, seq_length=50)


class TimeTerms:
    """The time-interval aware key / value embeddings of one ACTiSASRec forward (actisasrec.py:104-124), as the layers use them:
    tables + interval matrix + dropout streams instead of the reference's gathered and dropped [B,L,L,d] tensors.  One instance
    is shared by all layers (the reference draws the four dropout masks once per forward)."""

    def __init__(self, pos_k, pos_v, time_k, time_v, tmat, n_heads, p, rt):
        from . import ops
        self.ops = ops
        self.pos_k, self.pos_v, self.time_k, self.time_v = pos_k, pos_v, time_k, time_v
        m = (lambda k: rt.mask(k)) if p > 0 else (lambda k: None)
        self.spec_k = ops.PairSpec(tmat, n_heads, p, m('posK'), m('timeK'), rt.rng, 2, 4)
        self.spec_v = ops.PairSpec(tmat, n_heads, p, m('posV'), m('timeV'), rt.rng, 3, 5)

    def score_bias(self, mq):
        return self.ops.PairScoreFn.apply(mq, self.pos_k, self.time_k, self.spec_k, 1)

    def context(self, prob):
        return self.ops.PairContextFn.apply(prob, self.pos_v, self.time_v, self.spec_v)


class ACTimeAwareMultiHeadAttention(AttackRMultiHeadAttention):
    """transformer_layers.py:1010-1177: the parameter set is that of AttackRMultiHeadAttention"""


class ACTimeAwareTransformerLayer(AttackRTransformerLayer):
    """transformer_layers.py:1232-1327.  forward takes a TimeTerms object where the reference takes the four gathered tensors
    (absolute_pos_K, absolute_pos_V, time_matrix_emb_K, time_matrix_emb_V)."""

    def forward(self, hidden_states, attention_mask, time_terms, return_attention_prob=False, rt=None, layer_idx=0,
                need_attacked=True):
        return super().forward(hidden_states, attention_mask, return_attention_prob, False, rt=rt, layer_idx=layer_idx,
                               need_attacked=need_attacked, time_terms=time_terms)


class ACTimeAwareTransformerEncoder(AttackRTransformerEncoder):
    """transformer_layers.py:1330-1448."""

    layer_class = ACTimeAwareTransformerLayer

    def forward(self, hidden_states, attention_mask, time_terms, output_all_encoded_layers=True, rt=None):
        from .layers import default_runtime
        rt = rt or default_runtime(hidden_states.device)
        all_encoder_layers, all_attack_masks = [], []
        att = cal = None
        n = len(self.layer)
        for layer_idx, layer_module in enumerate(self.layer):
            need_att = (layer_idx == n - 1) or not rt.attacked_last_only
            att, cal, attack_mask, _ = layer_module(hidden_states, attention_mask, time_terms, rt=rt, layer_idx=layer_idx,
                                                    need_attacked=need_att)
            hidden_states = cal
            all_attack_masks.append(attack_mask)
            if output_all_encoded_layers:
                all_encoder_layers.append((att, cal))
        if not output_all_encoded_layers:
            all_encoder_layers.append((att, cal))
        return all_encoder_layers, all_attack_masks
