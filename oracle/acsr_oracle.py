"""CPU oracle for the AC-SASRec hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (not a copy) of the reference algorithm in plain
torch-on-CPU tensor algebra.  It is the checker for the CUDA path: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg
may import it.  Nothing under ``ac-tsr_b200/`` imports it and the product path
raises when the CUDA library is missing -- there is no CPU fallback.

Parity pinning: the reference ships no tests / golden vectors, so
the oracle is pinned against outputs of the reference itself, generated in the
build container by ``tests/golden/make_golden.py`` (imports /root/reference,
records every dropout mask and the attack noise, saves inputs + outputs +
gradients) and committed as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py``
checks the oracle against all of them.

Reference lines each function follows (relative to /root/reference/recbole):
  embed_ln_dropout   model/sequential_recommender/acsasrec.py:86-95
  additive_mask      model/abstract_recommender.py:136-143
  attn_calib         model/layers.py:657-674 (attack mask), 686-742 (spatial calibrator),
                     883-896 (combine), 917-936 (attacked / calibrated / combined attention)
  adjusted_output    model/layers.py:676-684
  feed_forward       model/layers.py:776-798
  ac_layer           model/layers.py:898-951
  encoder            model/layers.py:1097-1131
  forward            model/sequential_recommender/acsasrec.py:86-104
  calculate_loss     model/sequential_recommender/acsasrec.py:107-144 (CE and BPR branches; model/loss.py:21-47)
  predict/full_sort  model/sequential_recommender/acsasrec.py:146-164
  bert_*             model/sequential_recommender/acbert4rec.py:152-160, 162-179, 198-245, 260-267 (AcBERT4Rec)
  ti_*               model/sequential_recommender/actisasrec.py:104-224 (ACTiSASRec) on model/transformer_layers.py:1010-1327
  ssept_*            model/sequential_recommender/acssept.py:124-228 (ACSSEPT) on model/transformer_layers.py:742-953,
                     the encoder variant without the re-normalising softmaxes (cfg['attn_variant'] == 'transformer_layers')
  train_grads        trainer/trainer.py:660-687 (two backward passes routed by name)
  full_sort_topk     trainer/trainer.py:941-942, evaluator/collector.py:145-153
  metrics            evaluator/metrics.py:62-64,88-96,159-160,186-202; base_metric.py:65-80

The spatial-calibrator affine over cat(q_i, k_j) (layers.py:706-727) is restated
in its rank-1 form  w[:dh].q_i + w[dh:].k_j + b ; the [B,H,L,L,2dh] tensor is
never built.  Everything else keeps the reference's operation order.
"""
import math
from collections import OrderedDict

import numpy as np
import torch

MASK_NEG = -10000.0          # abstract_recommender.py:142
ORDER_EPS = 1e-24            # layers.py:719


# --------------------------------------------------------------------------- #
# randomness injection
# --------------------------------------------------------------------------- #
class Rand:
    """Holds the multiplicative dropout masks (0 or 1/(1-p)) and the attack noise.

    keys: 'emb' ; (l,'D1') P-dropout ; (l,'D2') P0-dropout ; (l,'D3') M-dropout ;
    (l,'noise') ; (l,'D4') attacked out-proj ; (l,'D5') calibrated out-proj ;
    (l,'D6') attacked FFN ; (l,'D7') calibrated FFN.  Missing key == identity
    (or zero noise)."""

    def __init__(self, d=None):
        self.d = dict(d or {})

    def mask(self, key, x):
        m = self.d.get(key)
        return x if m is None else x * m.to(x.dtype)

    def noise(self, l, like):
        n = self.d.get((l, 'noise'))
        return torch.zeros_like(like) if n is None else n.to(like.dtype)


def draw_rand(cfg, B, L, seed, train=True, dtype=torch.float32):
    """Seeded masks/noise in the shapes the kernels' explicit mode takes."""
    g = torch.Generator().manual_seed(seed)
    H, d, N = cfg['n_heads'], cfg['hidden_size'], cfg['n_layers']
    ph, pa = cfg['hidden_dropout_prob'], cfg['attn_dropout_prob']

    def dm(shape, p):
        if not train or p <= 0:
            return None
        return (torch.rand(shape, generator=g) >= p).to(dtype) / (1.0 - p)
    r = {}
    r['emb'] = dm((B, L, d), ph)
    for l in range(N):
        r[(l, 'D1')] = dm((B, H, L, L), pa)
        r[(l, 'D2')] = dm((B, H, L, L), pa)
        r[(l, 'D3')] = dm((B, H, L, L), pa)
        r[(l, 'noise')] = torch.randn((B, H, L, L), generator=g).to(dtype)
        for k in ('D4', 'D5', 'D6', 'D7'):
            r[(l, k)] = dm((B, L, d), ph)
    return Rand({k: v for k, v in r.items() if v is not None})


# --------------------------------------------------------------------------- #
# building blocks
# --------------------------------------------------------------------------- #
def layer_norm(x, w, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def linear(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


def act_fn(name):
    if name == 'gelu':
        return lambda x: x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))   # layers.py:776-785
    if name == 'relu':
        return torch.relu
    if name == 'swish':
        return lambda x: x * torch.sigmoid(x)
    if name == 'tanh':
        return torch.tanh
    if name == 'sigmoid':
        return torch.sigmoid
    raise KeyError(name)


def additive_mask(item_seq, bidirectional=False):
    """[B,1,L,L]: 0 where key j is not padding and (j<=i, unless bidirectional), else -10000  (abstract_recommender.py:136-143)."""
    B, L = item_seq.shape
    keep = (item_seq != 0).view(B, 1, 1, L).expand(B, 1, L, L)
    if not bidirectional:
        keep = torch.tril(keep)
    return torch.where(keep, 0.0, MASK_NEG)


def embed_ln_dropout(item_seq, E, ln_w, ln_b, eps, rnd, pos_emb=None):
    x = torch.nn.functional.embedding(item_seq, E, padding_idx=0)   # acsasrec.py:37: no gradient to row 0 via the gather
    if pos_emb is not None:
        x = x + pos_emb[: item_seq.shape[1]].unsqueeze(0)
    x = layer_norm(x, ln_w, ln_b, eps)
    return rnd.mask('emb', x)


def split_heads(x, H):
    B, L, d = x.shape
    return x.view(B, L, H, d // H).permute(0, 2, 1, 3)      # [B,H,L,dh]


def merge_heads(x):
    B, H, L, dh = x.shape
    return x.permute(0, 2, 1, 3).reshape(B, L, H * dh)


def attn_calib(mq, mk, mv, aq, ak, gate_logit, mask, lp, cfg, l, rnd, anneal_rate=None, s_bias=None, ctx_extra=None):
    """Core of one AC layer on already-projected tensors.

    mq,mk,mv : x.Wq+bq etc. [B,L,d];  aq,ak : attack transforms of mq,mk [B,L,d]
    gate_logit : mq.Wg^T+bg [B,L,L] (combine_option 'gate') or None
    Returns dict with P0,P,M,A,C,R (all [B,H,L,L]), ctx_att, ctx_cal [B,L,d], pen_sq (sum (1-M)^2).
    """
    H = cfg['n_heads']
    dh = mq.shape[-1] // H
    L = mq.shape[1]
    sq = math.sqrt(dh)
    q, k, v = split_heads(mq, H), split_heads(mk, H), split_heads(mv, H)
    S = q @ k.transpose(-1, -2)
    if s_bias is not None:          # ACTiSASRec: q.posK + q.timeK[t_ij] join the raw scores (transformer_layers.py:1128-1134)
        S = S + s_bias
    dt = S.dtype
    e_o = torch.zeros_like(S)
    e_d = torch.zeros_like(S)
    idx = torch.arange(L, device=S.device)
    if cfg['use_order']:
        wo, bo = lp['order_affine.weight'][0], lp['order_affine.bias'][0]
        u = (q @ wo[:dh]).unsqueeze(-1) + (k @ wo[dh:]).unsqueeze(-2) + bo
        gd = (idx.view(1, L) > idx.view(L, 1)).to(dt)            # triu(diagonal=1): 1 where j>i
        pr = torch.sigmoid(u)
        e_o = torch.log(pr + ORDER_EPS) * gd + torch.log(1 - pr + ORDER_EPS) * (1 - gd)
    if cfg['use_distance']:
        wd, bd = lp['distance_affine.weight'][0], lp['distance_affine.bias'][0]
        t = (q @ wd[:dh]).unsqueeze(-1) + (k @ wd[dh:]).unsqueeze(-2) + bd
        gdist = torch.log((idx.view(1, L) - idx.view(L, 1)).abs().to(torch.float32) + 1).to(dt)
        e_d = -torch.square(gdist - t) * torch.square(lp['scalar'][0]) / 2
    P = rnd.mask((l, 'D1'), torch.softmax((S + e_o + e_d) / sq + mask, -1))
    P0 = rnd.mask((l, 'D2'), torch.softmax(S / sq + mask, -1))
    origin = P if cfg['two_level'] else P0
    Sa = split_heads(aq, H) @ split_heads(ak, H).transpose(-1, -2)
    M = rnd.mask((l, 'D3'), torch.softmax(Sa / sq + mask, -1))
    n = rnd.noise(l, M)
    # transformer_layers.py:919-927 (ACSSEPT / ACTiSASRec) is the same layer WITHOUT the three re-normalising softmaxes of
    # layers.py:917-925: the attacked "probabilities" are origin*M + noise*(1-M) as they come (pure noise on masked keys)
    plain = cfg.get('attn_variant', 'layers') == 'transformer_layers'
    renorm = (lambda z: z) if plain else (lambda z: torch.softmax(z + mask, -1))
    A = renorm(origin * M + n * (1 - M))
    C = renorm(origin * torch.exp(1 - M))
    opt = cfg['combine_option']
    if opt == 'gate':
        g = torch.sigmoid(gate_logit).unsqueeze(1)
        comb = g * origin + (1 - g) * C
    elif opt == 'fixed':
        comb = torch.softmax(origin + 0.5 * C, -1)
    elif opt == 'annealing':
        comb = anneal_rate * origin + (1 - anneal_rate) * C
    else:
        raise KeyError(opt)
    R = renorm(comb)
    if not cfg['two_level']:
        rc = cfg['rich_calibrated_combine']
        if rc == 'fixed':
            R = (R + P) / 2
        elif rc == 'trainable':
            ratio = lp['rich_calibrated_combine_ratio']
            R = ratio * R + (1 - ratio) * P
        else:
            raise KeyError(rc)
    ca, cc = A @ v, R @ v
    if ctx_extra is not None:       # ACTiSASRec: probs.posV + probs.timeV[t_ij] join the context (transformer_layers.py:1088-1091)
        ca, cc = ca + ctx_extra(A), cc + ctx_extra(R)
    return dict(P0=P0, P=P, M=M, A=A, C=C, R=R, ctx_att=merge_heads(ca), ctx_cal=merge_heads(cc),
                pen_sq=torch.sum((1 - M) ** 2))


def sub(params, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in params.items() if k.startswith(prefix)}


def adjusted_output(ctx, x, ap, eps, rnd, key):
    h = linear(ctx, ap['dense.weight'], ap['dense.bias'])
    return layer_norm(rnd.mask(key, h) + x, ap['LayerNorm.weight'], ap['LayerNorm.bias'], eps)


def feed_forward(h, fp, cfg, rnd, key):
    z = act_fn(cfg['hidden_act'])(linear(h, fp['dense_1.weight'], fp['dense_1.bias']))
    z = linear(z, fp['dense_2.weight'], fp['dense_2.bias'])
    return layer_norm(rnd.mask(key, z) + h, fp['LayerNorm.weight'], fp['LayerNorm.bias'], cfg['layer_norm_eps'])


def ac_layer(x, mask, params, cfg, l, rnd, anneal_rate=None, want_probs=False, time_terms=None, prefix='trm_encoder'):
    lp_all = sub(params, '%s.layer.%d.' % (prefix, l))
    ap = sub(lp_all, 'attack_attention.')
    mq = linear(x, ap['query.weight'], ap['query.bias'])
    mk = linear(x, ap['key.weight'], ap['key.bias'])
    mv = linear(x, ap['value.weight'], ap['value.bias'])
    aq = linear(mq, ap['attack_query_transform.weight'], ap['attack_query_transform.bias'])
    ak = linear(mk, ap['attack_key_transform.weight'], ap['attack_key_transform.bias'])
    gl = None
    if cfg['combine_option'] == 'gate':
        gl = linear(mq, lp_all['gate.weight'], lp_all['gate.bias'])
    ap2 = dict(ap)
    if 'rich_calibrated_combine_ratio' in lp_all:
        ap2['rich_calibrated_combine_ratio'] = lp_all['rich_calibrated_combine_ratio']
    s_bias = ctx_extra = None
    if time_terms is not None:
        s_bias, ctx_extra = time_terms(mq)
    r = attn_calib(mq, mk, mv, aq, ak, gl, mask, ap2, cfg, l, rnd, anneal_rate, s_bias, ctx_extra)
    eps = cfg['layer_norm_eps']
    h_att = adjusted_output(r['ctx_att'], x, ap, eps, rnd, (l, 'D4'))
    h_cal = adjusted_output(r['ctx_cal'], x, ap, eps, rnd, (l, 'D5'))
    fp = sub(lp_all, 'feed_forward.')
    out_att = feed_forward(h_att, fp, cfg, rnd, (l, 'D6'))
    out_cal = feed_forward(h_cal, fp, cfg, rnd, (l, 'D7'))
    if want_probs:
        return out_att, out_cal, r
    return out_att, out_cal, r['M']


def encode(params, cfg, item_seq, rnd=None, anneal_rates=None, want_probs=False, bidirectional=False):
    """-> attacked[B,L,d], calibrated[B,L,d] of the last layer, [M_l]: embedding + LayerNorm + dropout + the encoder stack."""
    rnd = rnd or Rand()
    pos = params.get('position_embedding.weight') if cfg.get('use_position_embedding') else None
    x = embed_ln_dropout(item_seq, params['item_embedding.weight'], params['LayerNorm.weight'],
                         params['LayerNorm.bias'], cfg['layer_norm_eps'], rnd, pos)
    mask = additive_mask(item_seq, bidirectional).to(x.dtype)
    Ms, att = [], None
    for l in range(cfg['n_layers']):
        ar = None if anneal_rates is None else anneal_rates[l]
        att, x, M = ac_layer(x, mask, params, cfg, l, rnd, ar, want_probs)
        Ms.append(M)
    return att, x, Ms


def forward(params, cfg, item_seq, item_len, rnd=None, anneal_rates=None, want_probs=False):
    """-> attacked[B,d], calibrated[B,d], [M_l] (acsasrec.py:86-104)."""
    att, x, Ms = encode(params, cfg, item_seq, rnd, anneal_rates, want_probs)
    B = item_seq.shape[0]
    rows = torch.arange(B, device=item_seq.device)
    return att[rows, item_len - 1], x[rows, item_len - 1], Ms


def cross_entropy(out, E, target):
    logits = out @ E.t()
    lse = torch.logsumexp(logits, -1)
    return (lse - logits[torch.arange(out.shape[0], device=out.device), target]).mean()


def bpr_loss(out, E, pos_items, neg_items, gamma=1e-10):
    """acsasrec.py:109-116 + loss.py:21-47: -log(gamma + sigmoid(out.E[pos] - out.E[neg])) averaged over the batch."""
    pos_score = (out * E[pos_items]).sum(-1)
    neg_score = (out * E[neg_items]).sum(-1)
    return -torch.log(gamma + torch.sigmoid(pos_score - neg_score)).mean()


def calculate_loss(params, cfg, item_seq, item_len, pos_items, rnd=None, anneal_rates=None, neg_items=None):
    """-> (final_attacked_loss, calibrated_loss)  acsasrec.py:123-144 (CE branch, or BPR when cfg['loss_type'] == 'BPR')."""
    att, cal, Ms = forward(params, cfg, item_seq, item_len, rnd, anneal_rates)
    E = params['item_embedding.weight']
    pens = [torch.sqrt(torch.sum((1 - (M['M'] if isinstance(M, dict) else M)) ** 2)) for M in Ms]
    pen = torch.stack(pens).mean()
    w = params['mask_loss_weight'][0] if cfg.get('trainable_mask_loss_weight') else cfg['mask_loss_weight']
    if cfg.get('loss_type', 'CE') == 'BPR':
        l_att = -bpr_loss(att, E, pos_items, neg_items) + pen * w
        l_cal = bpr_loss(cal, E, pos_items, neg_items)
        return l_att, l_cal
    l_att = -cross_entropy(att, E, pos_items) + pen * w
    l_cal = cross_entropy(cal, E, pos_items)
    return l_att, l_cal


# ---- AcBERT4Rec (acbert4rec.py:11-267): bidirectional mask, masked-item CE over table[:n_items] ---------------- #
def bert_calculate_loss(params, cfg, masked_seq, pos_items, masked_index, rnd=None, anneal_rates=None):
    """acbert4rec.py:207-245 given the masked sequence of reconstruct_train_data -> (final_attacked_loss, calibrated_loss)."""
    att, cal, Ms = encode(params, cfg, masked_seq, rnd, anneal_rates, bidirectional=True)
    B, L = masked_seq.shape
    rows = torch.arange(B, device=masked_seq.device).view(B, 1)
    E = params['item_embedding.weight'][:-1]                       # the last row is the mask token (acbert4rec.py:200)
    targets = (masked_index > 0).to(att.dtype).view(-1)

    def loss_of(h):
        seq_out = h[rows, masked_index].reshape(-1, h.shape[-1])   # one-hot bmm of acbert4rec.py:214-222 == a gather
        logits = seq_out @ E.t()
        ce = torch.logsumexp(logits, -1) - logits[torch.arange(seq_out.shape[0], device=h.device), pos_items.reshape(-1)]
        return (ce * targets).sum() / targets.sum()
    pens = [torch.sqrt(torch.sum((1 - (M['M'] if isinstance(M, dict) else M)) ** 2)) for M in Ms]
    w = params['mask_loss_weight'][0] if cfg.get('trainable_mask_loss_weight') else cfg['mask_loss_weight']
    return -loss_of(att) + torch.stack(pens).mean() * w, loss_of(cal)


def bert_test_sequence(item_seq, item_len, mask_token):
    """acbert4rec.py:152-160: one more position, the mask token at index item_len."""
    seq = torch.cat((item_seq, torch.zeros(item_seq.shape[0], 1, dtype=item_seq.dtype)), 1)
    seq[torch.arange(seq.shape[0]), item_len] = mask_token
    return seq


def bert_full_sort_scores(params, cfg, item_seq, item_len, rnd=None):
    """acbert4rec.py:260-267 -> (attacked_scores, scores) [B, n_items]."""
    E = params['item_embedding.weight']
    seq = bert_test_sequence(item_seq, item_len, E.shape[0] - 1)
    att, cal, _ = encode(params, cfg, seq, rnd, bidirectional=True)
    rows = torch.arange(seq.shape[0])
    return att[rows, item_len] @ E[:-1].t(), cal[rows, item_len] @ E[:-1].t()


def bert_train_grads(params, cfg, masked_seq, pos_items, masked_index, rnd=None, anneal_rates=None):
    """the two routed backward passes of trainer.py:672-686 for AcBERT4Rec -> (l_att, l_cal, {name: grad})."""
    p = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in params.items())
    l_att, l_cal = bert_calculate_loss(p, cfg, masked_seq, pos_items, masked_index, rnd, anneal_rates)
    names = list(p)
    g_cal = torch.autograd.grad(l_cal, [p[n] for n in names], retain_graph=True, allow_unused=True)
    g_att = torch.autograd.grad(l_att, [p[n] for n in names], allow_unused=True)
    grads = {}
    for n, gc, ga in zip(names, g_cal, g_att):
        g = ga if any(s in n for s in ATTACK_KEYS) else gc
        if n == 'mask_loss_weight':
            g = None
        grads[n] = torch.zeros_like(p[n]) if g is None else g.detach()
    return l_att.detach(), l_cal.detach(), grads


ATTACK_KEYS = ('attack_key_transform', 'attack_query_transform')   # trainer.py:673


# ---- ACTiSASRec (actisasrec.py:20-224): time-interval aware keys / values on the transformer_layers.py layer ---------- #
def ti_time_matrix(time_seq, time_span):
    """actisasrec.py:146-155: |t_i - t_j| clipped to time_span, as integers"""
    tm = (time_seq.unsqueeze(-1) - time_seq.unsqueeze(1)).abs()
    return torch.where(tm > time_span, torch.full_like(tm, time_span), tm).int()


def ti_forward(params, cfg, item_seq, item_len, time_seq, rnd=None, anneal_rates=None):
    """actisasrec.py:104-144 + transformer_layers.py:1010-1327 -> attacked[B,d], calibrated[B,d], [M_l].
    Dropout keys beyond the per-layer ones: 'emb', 'posK', 'posV' ([B,L,d]) and 'timeK', 'timeV' ([B,L,L,d]), drawn once per
    forward and shared by all layers (actisasrec.py:120-124)."""
    rnd = rnd or Rand()
    c = dict(cfg)
    c['attn_variant'] = 'transformer_layers'
    B, L = item_seq.shape
    H = cfg['n_heads']
    tmat = ti_time_matrix(time_seq, cfg['time_span']).long()
    emb = lambda name, idx: torch.nn.functional.embedding(idx, params[name + '.weight'], padding_idx=0)
    pos_ids = torch.arange(L).unsqueeze(0).expand(B, L)
    x = rnd.mask('emb', layer_norm(emb('item_embedding', item_seq), params['LayerNorm.weight'], params['LayerNorm.bias'],
                                   cfg['layer_norm_eps']))
    pK = rnd.mask('posK', emb('absolute_pos_K_embedding', pos_ids))            # [B,L,d]
    pV = rnd.mask('posV', emb('absolute_pos_V_embedding', pos_ids))
    tK = rnd.mask('timeK', emb('time_matrix_emb_K_embedding', tmat))           # [B,L,L,d]
    tV = rnd.mask('timeV', emb('time_matrix_emb_V_embedding', tmat))
    d = x.shape[-1]
    dh = d // H
    tKh = tK.view(B, L, L, H, dh).permute(0, 3, 1, 2, 4)                       # [B,H,L,L,dh]
    tVh = tV.view(B, L, L, H, dh).permute(0, 3, 1, 2, 4)

    def time_terms(mq):
        q = split_heads(mq, H)
        s_bias = q @ split_heads(pK, H).transpose(-1, -2) + (tKh @ q.unsqueeze(-1)).squeeze(-1)

        def ctx_extra(prob):
            return prob @ split_heads(pV, H) + (prob.unsqueeze(-2) @ tVh).squeeze(-2)
        return s_bias, ctx_extra
    mask = additive_mask(item_seq).to(x.dtype)
    Ms, att = [], None
    for l in range(cfg['n_layers']):
        ar = None if anneal_rates is None else anneal_rates[l]
        att, x, M = ac_layer(x, mask, params, c, l, rnd, ar, time_terms=time_terms, prefix='ti_trm_encoder')
        Ms.append(M)
    rows = torch.arange(B)
    return att[rows, item_len - 1], x[rows, item_len - 1], Ms


def ti_calculate_loss(params, cfg, item_seq, item_len, time_seq, pos_items, rnd=None, anneal_rates=None, neg_items=None):
    """actisasrec.py:173-193 -> (final_attacked_loss, calibrated_loss)"""
    att, cal, Ms = ti_forward(params, cfg, item_seq, item_len, time_seq, rnd, anneal_rates)
    E = params['item_embedding.weight']
    pen = torch.stack([torch.sqrt(torch.sum((1 - M) ** 2)) for M in Ms]).mean()
    w = params['mask_loss_weight'][0] if cfg.get('trainable_mask_loss_weight') else cfg['mask_loss_weight']
    if cfg.get('loss_type', 'CE') == 'BPR':
        return -bpr_loss(att, E, pos_items, neg_items) + pen * w, bpr_loss(cal, E, pos_items, neg_items)
    return -cross_entropy(att, E, pos_items) + pen * w, cross_entropy(cal, E, pos_items)


def ti_full_sort_scores(params, cfg, item_seq, item_len, time_seq, rnd=None):
    """actisasrec.py:211-224 -> (attacked_scores, scores) [B, n_items]"""
    att, cal, _ = ti_forward(params, cfg, item_seq, item_len, time_seq, rnd)
    E = params['item_embedding.weight']
    return att @ E.t(), cal @ E.t()


def ti_predict(params, cfg, item_seq, item_len, time_seq, test_item, rnd=None):
    att, cal, _ = ti_forward(params, cfg, item_seq, item_len, time_seq, rnd)
    e = params['item_embedding.weight'][test_item]
    return (att * e).sum(1), (cal * e).sum(1)


def ti_train_grads(params, cfg, item_seq, item_len, time_seq, pos_items, rnd=None, anneal_rates=None, neg_items=None):
    """the two routed backward passes of trainer.py:672-686 for ACTiSASRec -> (l_att, l_cal, {name: grad})"""
    p = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in params.items())
    l_att, l_cal = ti_calculate_loss(p, cfg, item_seq, item_len, time_seq, pos_items, rnd, anneal_rates, neg_items)
    names = list(p)
    g_cal = torch.autograd.grad(l_cal, [p[n] for n in names], retain_graph=True, allow_unused=True)
    g_att = torch.autograd.grad(l_att, [p[n] for n in names], allow_unused=True)
    grads = {}
    for n, gc, ga in zip(names, g_cal, g_att):
        g = ga if any(s in n for s in ATTACK_KEYS) else gc
        if n == 'mask_loss_weight':
            g = None
        grads[n] = torch.zeros_like(p[n]) if g is None else g.detach()
    return l_att.detach(), l_cal.detach(), grads


# ---- ACSSEPT (acssept.py:21-228): user embedding concatenated to every position, transformer_layers.py encoder ---- #
def ssept_cfg(cfg):
    """the encoder ACSSEPT builds: hidden = item_hidden_size + user_hidden_size, transformer_layers.py:873-953 layers"""
    c = dict(cfg)
    c['hidden_size'] = cfg['item_hidden_size'] + cfg['user_hidden_size']
    c['attn_variant'] = 'transformer_layers'
    return c


def ssept_forward(params, cfg, item_seq, item_len, user_id, rnd=None, anneal_rates=None):
    """acssept.py:124-145 -> attacked[B,d], calibrated[B,d], [M_l]"""
    rnd = rnd or Rand()
    c = ssept_cfg(cfg)
    B, L = item_seq.shape
    item_emb = torch.nn.functional.embedding(item_seq, params['item_embedding.weight'], padding_idx=0)
    user_emb = torch.nn.functional.embedding(user_id, params['user_embedding.weight'], padding_idx=0)
    x = torch.cat((item_emb, user_emb.unsqueeze(1).expand(B, L, user_emb.shape[-1])), -1)
    if cfg.get('use_position_embedding'):
        x = x + params['position_embedding.weight'][:L].unsqueeze(0)
    x = rnd.mask('emb', layer_norm(x, params['LayerNorm.weight'], params['LayerNorm.bias'], cfg['layer_norm_eps']))
    mask = additive_mask(item_seq).to(x.dtype)
    Ms, att = [], None
    for l in range(cfg['n_layers']):
        ar = None if anneal_rates is None else anneal_rates[l]
        att, x, M = ac_layer(x, mask, params, c, l, rnd, ar)
        Ms.append(M)
    rows = torch.arange(B, device=item_seq.device)
    return att[rows, item_len - 1], x[rows, item_len - 1], Ms


def ssept_logits(out, E, user_vec):
    """acssept.py:164-171 / 212-224: every candidate row is cat(E[v], user_vec[b]) -> out_item.E^T + (out_user.user_vec)"""
    di = E.shape[1]
    return out[:, :di] @ E.t() + (out[:, di:] * user_vec).sum(-1, keepdim=True)


def ssept_calculate_loss(params, cfg, item_seq, item_len, user_id, pos_items, rnd=None, anneal_rates=None, neg_items=None):
    """acssept.py:174-190 -> (final_attacked_loss, calibrated_loss); the loss uses user_embedding (acssept.py:150,166)"""
    att, cal, Ms = ssept_forward(params, cfg, item_seq, item_len, user_id, rnd, anneal_rates)
    E = params['item_embedding.weight']
    u = torch.nn.functional.embedding(user_id, params['user_embedding.weight'], padding_idx=0)
    rows = torch.arange(att.shape[0])

    def loss_of(out):
        if cfg.get('loss_type', 'CE') == 'BPR':
            pe, ne = torch.cat((E[pos_items], u), -1), torch.cat((E[neg_items], u), -1)
            return -torch.log(1e-10 + torch.sigmoid((out * pe).sum(-1) - (out * ne).sum(-1))).mean()
        logits = ssept_logits(out, E, u)
        return (torch.logsumexp(logits, -1) - logits[rows, pos_items]).mean()
    pen = torch.stack([torch.sqrt(torch.sum((1 - M) ** 2)) for M in Ms]).mean()
    w = params['mask_loss_weight'][0] if cfg.get('trainable_mask_loss_weight') else cfg['mask_loss_weight']
    return -loss_of(att) + pen * w, loss_of(cal)


def ssept_full_sort_scores(params, cfg, item_seq, item_len, user_id, rnd=None):
    """acssept.py:209-228 -> (attacked_scores, scores) [B, n_items]; scoring uses user_TEST_embedding (acssept.py:213)"""
    att, cal, _ = ssept_forward(params, cfg, item_seq, item_len, user_id, rnd)
    ut = torch.nn.functional.embedding(user_id, params['user_test_embedding.weight'], padding_idx=0)
    E = params['item_embedding.weight']
    return ssept_logits(att, E, ut), ssept_logits(cal, E, ut)


def ssept_predict(params, cfg, item_seq, item_len, user_id, test_item, rnd=None):
    """acssept.py:192-207"""
    att, cal, _ = ssept_forward(params, cfg, item_seq, item_len, user_id, rnd)
    ut = torch.nn.functional.embedding(user_id, params['user_test_embedding.weight'], padding_idx=0)
    e = torch.cat((params['item_embedding.weight'][test_item], ut), -1)
    return (att * e).sum(1), (cal * e).sum(1)


def ssept_train_grads(params, cfg, item_seq, item_len, user_id, pos_items, rnd=None, anneal_rates=None, neg_items=None):
    """the two routed backward passes of trainer.py:672-686 (AttackSASRecTrainer) for ACSSEPT -> (l_att, l_cal, {name: grad})"""
    p = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in params.items())
    l_att, l_cal = ssept_calculate_loss(p, cfg, item_seq, item_len, user_id, pos_items, rnd, anneal_rates, neg_items)
    names = list(p)
    g_cal = torch.autograd.grad(l_cal, [p[n] for n in names], retain_graph=True, allow_unused=True)
    g_att = torch.autograd.grad(l_att, [p[n] for n in names], allow_unused=True)
    grads = {}
    for n, gc, ga in zip(names, g_cal, g_att):
        g = ga if any(s in n for s in ATTACK_KEYS) else gc
        if n == 'mask_loss_weight':
            g = None
        grads[n] = torch.zeros_like(p[n]) if g is None else g.detach()
    return l_att.detach(), l_cal.detach(), grads


def train_grads(params, cfg, item_seq, item_len, pos_items, rnd=None, anneal_rates=None, neg_items=None):
    """The two backward passes of trainer.py:672-686 -> (l_att, l_cal, {name: grad}).

    attack_{query,key}_transform get d l_att, every other parameter gets d l_cal;
    a trainable mask_loss_weight gets nothing (it only enters l_att)."""
    p = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in params.items())
    l_att, l_cal = calculate_loss(p, cfg, item_seq, item_len, pos_items, rnd, anneal_rates, neg_items)
    names = list(p)
    g_cal = torch.autograd.grad(l_cal, [p[n] for n in names], retain_graph=True, allow_unused=True)
    g_att = torch.autograd.grad(l_att, [p[n] for n in names], allow_unused=True)
    grads = {}
    for n, gc, ga in zip(names, g_cal, g_att):
        g = ga if any(s in n for s in ATTACK_KEYS) else gc
        if n == 'mask_loss_weight':
            g = None
        grads[n] = torch.zeros_like(p[n]) if g is None else g.detach()
    return l_att.detach(), l_cal.detach(), grads


def adam_step(param, grad, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """torch.optim.Adam (non-amsgrad) single-tensor update; returns new (param, m, v)."""
    if weight_decay:
        grad = grad + weight_decay * param
    m = beta1 * m + (1 - beta1) * grad
    v = beta2 * v + (1 - beta2) * grad * grad
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return param - (lr / bc1) * m / denom, m, v


def predict(params, cfg, item_seq, item_len, test_item, rnd=None):
    att, cal, _ = forward(params, cfg, item_seq, item_len, rnd)
    e = params['item_embedding.weight'][test_item]
    return (att * e).sum(1), (cal * e).sum(1)


def full_sort_scores(params, cfg, item_seq, item_len, rnd=None):
    _, cal, _ = forward(params, cfg, item_seq, item_len, rnd)
    return cal @ params['item_embedding.weight'].t()


def full_sort_topk(scores, k):
    """scores[:,0] = -inf then top-k (trainer.py:941-942, collector.py:147)."""
    s = scores.clone()
    s[:, 0] = -float('inf')
    val, idx = torch.topk(s, k, dim=-1)
    return val, idx


def topk_equal_modulo_ties(idx_a, idx_b, scores, rtol=2e-6):
    """True when two top-k index lists agree except where the scores tie within
    rtol*max|score| (two fp32 GEMMs with different summation order can swap such pairs)."""
    scores = scores.double()
    tol = rtol * scores.abs().clamp(max=1e30)[:, 1:].max().item()
    sa = torch.gather(scores, 1, idx_a)
    sb = torch.gather(scores, 1, idx_b)
    bad = (idx_a != idx_b) & ((sa - sb).abs() > tol)
    return not bool(bad.any()), int(bad.sum())


def hit_flags(topk_idx, pos_items):
    """[B,k] int flags + pos_len (always 1)  collector.py:148-153."""
    return (topk_idx == pos_items.view(-1, 1)).to(torch.int32)


# ---- metrics on the [n_users,k] bool matrix (numpy, fp64) ------------------ #
def metric_hit(pos):
    return (np.cumsum(pos, axis=1) > 0).astype(int)


def metric_recall(pos, pos_len):
    return np.cumsum(pos, axis=1) / pos_len.reshape(-1, 1)


def metric_mrr(pos):
    idxs = pos.argmax(axis=1)
    res = np.zeros(pos.shape, dtype=np.float64)
    for r, i in enumerate(idxs):
        res[r, i:] = 1.0 / (i + 1) if pos[r, i] > 0 else 0.0
    return res


def metric_ndcg(pos, pos_len):
    k = pos.shape[1]
    idcg_len = np.minimum(pos_len, k)
    ranks = np.arange(1, k + 1, dtype=np.float64)
    idcg = np.tile(np.cumsum(1.0 / np.log2(ranks + 1)), (pos.shape[0], 1))
    for r, i in enumerate(idcg_len):
        idcg[r, i:] = idcg[r, i - 1]
    dcg = np.cumsum(np.where(pos, 1.0 / np.log2(ranks + 1), 0), axis=1)
    return dcg / idcg


def topk_metrics(pos, pos_len, topk=(1, 3, 5, 10, 20, 50), names=('hit', 'mrr', 'ndcg', 'recall'), decimals=4):
    pos = np.asarray(pos).astype(bool)
    pos_len = np.asarray(pos_len)
    fn = {'hit': lambda: metric_hit(pos), 'mrr': lambda: metric_mrr(pos),
          'ndcg': lambda: metric_ndcg(pos, pos_len), 'recall': lambda: metric_recall(pos, pos_len)}
    out = OrderedDict()
    for n in names:
        avg = fn[n]().mean(axis=0)
        for k in topk:
            out['%s@%d' % (n, k)] = round(float(avg[k - 1]), decimals)
    return out


# ---- synthetic inputs -------------------------------------- #
def synth_batch(B, L, V, seed=42, full_len=False):
    """lengths ~ clip(round(LogNormal(ln 7, .8)),1,L); items Zipf(1) on [1,V-1]; right-padded with 0."""
    g = torch.Generator().manual_seed(seed)
    if full_len:
        ln = torch.full((B,), L, dtype=torch.int64)
    else:
        ln = torch.exp(torch.randn(B, generator=g) * 0.8 + math.log(7.0)).round().clamp(1, L).to(torch.int64)
    w = 1.0 / torch.arange(1, V, dtype=torch.float64)
    items = torch.multinomial(w, B * (L + 1), replacement=True, generator=g).view(B, L + 1) + 1
    seq = items[:, :L].clone()
    seq[torch.arange(L).view(1, L) >= ln.view(B, 1)] = 0
    return seq, ln, items[:, L].clone()


def default_cfg(**kw):
    c = dict(n_layers=2, n_heads=2, hidden_size=64, inner_size=256, hidden_dropout_prob=0.5,
             attn_dropout_prob=0.5, hidden_act='gelu', layer_norm_eps=1e-12, initializer_range=0.02,
             loss_type='CE', combine_option='gate', rich_calibrated_combine='none', two_level=True,
             use_position_embedding=False, use_order=True, use_distance=True,
             trainable_mask_loss_weight=False, mask_loss_weight=0.03, MAX_ITEM_LIST_LENGTH=50)
    c.update(kw)
    return c


def init_params(cfg, n_items, seed=42, dtype=torch.float32):
    """Same parameter set / init distribution as acsasrec.py:36-84, layers.py:615-881 (own RNG order)."""
    g = torch.Generator().manual_seed(seed)
    d, I, L, H = cfg['hidden_size'], cfg['inner_size'], cfg['MAX_ITEM_LIST_LENGTH'], cfg['n_heads']
    dh, std = d // H, cfg['initializer_range']
    p = OrderedDict()

    def nrm(*s):
        return (torch.randn(*s, generator=g) * std).to(dtype)
    p['item_embedding.weight'] = nrm(n_items, d)
    if cfg.get('use_position_embedding'):
        p['position_embedding.weight'] = nrm(L, d)
    for l in range(cfg['n_layers']):
        a = 'trm_encoder.layer.%d.attack_attention.' % l
        if cfg['use_distance']:
            p[a + 'scalar'] = torch.randn(1, generator=g).to(dtype)
        for n in ('query', 'key', 'value'):
            p[a + n + '.weight'] = nrm(d, d); p[a + n + '.bias'] = torch.zeros(d, dtype=dtype)
        if cfg['use_order']:
            p[a + 'order_affine.weight'] = nrm(1, 2 * dh); p[a + 'order_affine.bias'] = torch.zeros(1, dtype=dtype)
        if cfg['use_distance']:
            p[a + 'distance_affine.weight'] = nrm(1, 2 * dh); p[a + 'distance_affine.bias'] = torch.zeros(1, dtype=dtype)
        for n in ('attack_query_transform', 'attack_key_transform', 'dense'):
            p[a + n + '.weight'] = nrm(d, d); p[a + n + '.bias'] = torch.zeros(d, dtype=dtype)
        p[a + 'LayerNorm.weight'] = torch.ones(d, dtype=dtype); p[a + 'LayerNorm.bias'] = torch.zeros(d, dtype=dtype)
        b = 'trm_encoder.layer.%d.' % l
        if cfg['combine_option'] == 'gate':
            p[b + 'gate.weight'] = nrm(L, d); p[b + 'gate.bias'] = torch.zeros(L, dtype=dtype)
        f = b + 'feed_forward.'
        p[f + 'dense_1.weight'] = nrm(I, d); p[f + 'dense_1.bias'] = torch.zeros(I, dtype=dtype)
        p[f + 'dense_2.weight'] = nrm(d, I); p[f + 'dense_2.bias'] = torch.zeros(d, dtype=dtype)
        p[f + 'LayerNorm.weight'] = torch.ones(d, dtype=dtype); p[f + 'LayerNorm.bias'] = torch.zeros(d, dtype=dtype)
    p['LayerNorm.weight'] = torch.ones(d, dtype=dtype); p['LayerNorm.bias'] = torch.zeros(d, dtype=dtype)
    if cfg.get('trainable_mask_loss_weight'):
        p['mask_loss_weight'] = torch.tensor([0.3], dtype=dtype)
    return p
