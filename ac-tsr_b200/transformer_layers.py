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
