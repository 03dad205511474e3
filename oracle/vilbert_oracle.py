"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the two-stream ViLBERT hot path.

A functional restatement (plain dense tensor algebra on the host, fp32 or fp64, no modules, no CUDA) of
what ``/root/reference`` computes for ``Lily.forward`` + ``get_loss_correct`` in ``eval()`` mode
(dropout = identity).  It is the checker for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the product package
(``youtube-vln_b200/``) never does.

Pinning: the reference has no tests or golden vectors of its own, so the oracle
is pinned against outputs of the *real* reference imported in the build container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``; checked by ``tests/test_oracle_golden.py``).

Weights are a flat ``{state_dict key: tensor}`` mapping with the reference's ``Lily`` key names
(``bert.*``, ``cls.*``, ``vil_logit.*``, ``judge.*``).  Gradients come from autograd over this
restatement (every op below is differentiable), so ``oracle_grads`` is also the backward oracle.

Reference lines followed (all under /root/reference):
  gelu                      vilbert/vilbert.py:113-119
  BertLayerNorm             vilbert/vilbert.py:204-217  (biased variance, eps inside sqrt)
  BertEmbeddings            vilbert/vilbert.py:240-256
  BertImageEmbeddings       vilbert/vilbert.py:1356-1370
  Bert(Image)SelfAttention  vilbert/vilbert.py:284-311, 413-440
  Bert(Image)SelfOutput     vilbert/vilbert.py:321-325, 449-453
  Bert(Image)Intermediate   vilbert/vilbert.py:351-354, 478-481
  Bert(Image)Output         vilbert/vilbert.py:364-368, 491-495
  BertBiAttention           vilbert/vilbert.py:552-618  (co_attention_mask term is dead: :736)
  BertBiOutput              vilbert/vilbert.py:638-650
  BertConnectionLayer       vilbert/vilbert.py:665-679
  BertEncoder schedule      vilbert/vilbert.py:737-816
  poolers                   vilbert/vilbert.py:827-848
  prediction heads          vilbert/vilbert.py:863-867, 882-886, 904-907, 939-954, 966-969
  BertModel masks           vilbert/vilbert.py:1254-1287
  Lily tail                 lily.py:93-127
  losses                    utils/utils_init.py:108-164, utils/dataset/common.py:21-26
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

LN_EPS = 1e-12

#: dropout probability applied at the reference's 94 ``nn.Dropout`` sites; 0 = eval mode (every parity check).  Only
#: bench.py's "stock PyTorch ops on the GPU" comparison switches it on (``oracle_step(..., train_dropout=True)``) so
#: that the proxy pays for the same bernoulli kernels the reference's train-mode step issues.
_DROP = {"p": 0.0}


def _drop(x):
    return F.dropout(x, _DROP["p"], True) if _DROP["p"] > 0.0 else x


def gelu_erf(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def layer_norm(x, w, b):
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return w * ((x - u) / torch.sqrt(s + LN_EPS)) + b


def linear(x, sd, prefix):
    y = x @ sd[prefix + ".weight"].transpose(0, 1)
    b = sd.get(prefix + ".bias")
    return y if b is None else y + b


def _heads(x, n_heads):
    n, s, h = x.shape
    return x.view(n, s, n_heads, h // n_heads).permute(0, 2, 1, 3)


def _merge(x):
    n, h, s, d = x.shape
    return x.permute(0, 2, 1, 3).reshape(n, s, h * d)


def attention_core(q, k, v, add_mask, n_heads):
    """softmax(q k^T / sqrt(dh) + mask) v with [N,1,1,S_k] additive mask."""
    qh, kh, vh = _heads(q, n_heads), _heads(k, n_heads), _heads(v, n_heads)
    dh = qh.shape[-1]
    scores = qh @ kh.transpose(-1, -2) / math.sqrt(dh) + add_mask
    probs = _drop(torch.softmax(scores, dim=-1))
    return _merge(probs @ vh), probs


def self_attention(x, add_mask, sd, p, n_heads):
    q = linear(x, sd, p + ".query")
    k = linear(x, sd, p + ".key")
    v = linear(x, sd, p + ".value")
    return attention_core(q, k, v, add_mask, n_heads)[0]


def dense_res_ln(x, res, sd, p, ln="LayerNorm", dense="dense"):
    return layer_norm(_drop(linear(x, sd, f"{p}.{dense}")) + res, sd[f"{p}.{ln}.weight"], sd[f"{p}.{ln}.bias"])


def transformer_layer(x, add_mask, sd, p, n_heads):
    ctx = self_attention(x, add_mask, sd, p + ".attention.self", n_heads)
    a = dense_res_ln(ctx, x, sd, p + ".attention.output")
    inter = gelu_erf(linear(a, sd, p + ".intermediate.dense"))
    return dense_res_ln(inter, a, sd, p + ".output")


def connection_layer(v, v_mask, t, t_mask, sd, p, n_heads):
    b = p + ".biattention"
    q1, k1, v1 = (linear(v, sd, f"{b}.{n}1") for n in ("query", "key", "value"))
    q2, k2, v2 = (linear(t, sd, f"{b}.{n}2") for n in ("query", "key", "value"))
    ctx1 = attention_core(q2, k1, v1, v_mask, n_heads)[0]          # text attends vision  [N,T,bi]
    ctx2 = attention_core(q1, k2, v2, t_mask, n_heads)[0]          # vision attends text  [N,V,bi]
    o = p + ".biOutput"
    v_a = layer_norm(_drop(linear(ctx2, sd, o + ".dense1")) + v, sd[o + ".LayerNorm1.weight"], sd[o + ".LayerNorm1.bias"])
    t_a = layer_norm(_drop(linear(ctx1, sd, o + ".dense2")) + t, sd[o + ".LayerNorm2.weight"], sd[o + ".LayerNorm2.bias"])
    v_o = dense_res_ln(gelu_erf(linear(v_a, sd, p + ".v_intermediate.dense")), v_a, sd, p + ".v_output")
    t_o = dense_res_ln(gelu_erf(linear(t_a, sd, p + ".t_intermediate.dense")), t_a, sd, p + ".t_output")
    return v_o, t_o


def text_embeddings(tokens, segs, sd, p="bert.embeddings"):
    pos = torch.arange(tokens.shape[1], device=tokens.device)
    # nn.Embedding(..., padding_idx=0) in the reference (vilbert/vilbert.py:225-227): token 0 receives no gradient
    e = F.embedding(tokens, sd[p + ".word_embeddings.weight"], padding_idx=0) \
        + sd[p + ".position_embeddings.weight"][pos][None] \
        + sd[p + ".token_type_embeddings.weight"][segs]
    return _drop(layer_norm(e, sd[p + ".LayerNorm.weight"], sd[p + ".LayerNorm.bias"]))


def image_embeddings(feat, loc, sd, p="bert.v_embeddings"):
    e = linear(feat, sd, p + ".image_embeddings")
    e = e + linear(loc[..., :5], sd, p + ".image_location_embeddings") \
        + linear(loc[..., 5:9], sd, p + ".image_orientation_embeddings") \
        + linear(loc[..., 9:11], sd, p + ".image_next_orientation_embeddings") \
        + sd[p + ".image_sequence_embeddings.weight"][loc[..., 11].long()]
    return _drop(layer_norm(e, sd[p + ".LayerNorm.weight"], sd[p + ".LayerNorm.bias"]))


def bert_model(sd, cfg, tokens, feat, loc, segs=None, tmask=None, vmask=None):
    """BertModel.forward (vilbert/vilbert.py:1242-1337) -> (seq_t, seq_v, pooled_t, pooled_v)."""
    dt = sd["bert.embeddings.LayerNorm.weight"].dtype
    if tmask is None:
        tmask = torch.ones_like(tokens)
    if segs is None:
        segs = torch.zeros_like(tokens)
    if vmask is None:
        vmask = torch.ones(feat.shape[0], feat.shape[1], dtype=tokens.dtype, device=tokens.device)
    t_add = ((1.0 - tmask.to(dt)) * -10000.0)[:, None, None, :]
    v_add = ((1.0 - vmask.to(dt)) * -10000.0)[:, None, None, :]
    t = text_embeddings(tokens, segs, sd)
    v = image_embeddings(feat.to(dt), loc.to(dt), sd)
    th, vh, bh = cfg["num_attention_heads"], cfg["v_num_attention_heads"], cfg["bi_num_attention_heads"]
    v_start = t_start = 0
    for count, (v_end, t_end) in enumerate(zip(cfg["v_biattention_id"], cfg["t_biattention_id"])):
        for i in range(v_start, v_end):
            v = transformer_layer(v, v_add, sd, f"bert.encoder.v_layer.{i}", vh)
        for i in range(t_start, t_end):
            t = transformer_layer(t, t_add, sd, f"bert.encoder.layer.{i}", th)
        if cfg.get("with_coattention", True):
            v, t = connection_layer(v, v_add, t, t_add, sd, f"bert.encoder.c_layer.{count}", bh)
        v_start, t_start = v_end, t_end
    for i in range(v_start, cfg["v_num_hidden_layers"]):
        v = transformer_layer(v, v_add, sd, f"bert.encoder.v_layer.{i}", vh)
    for i in range(t_start, cfg["num_hidden_layers"]):
        t = transformer_layer(t, t_add, sd, f"bert.encoder.layer.{i}", th)
    pooled_t = torch.relu(linear(t[:, 0], sd, "bert.t_pooler.dense"))
    pooled_v = torch.relu(linear(v[:, 0], sd, "bert.v_pooler.dense"))
    return t, v, pooled_t, pooled_v


def pretraining_heads(sd, seq_t, seq_v, pooled_t, pooled_v, fusion="mul"):
    """BertPreTrainingHeads.forward (vilbert/vilbert.py:939-954), eval mode."""
    p = "cls.predictions"
    h = gelu_erf(linear(seq_t, sd, p + ".transform.dense"))
    h = layer_norm(h, sd[p + ".transform.LayerNorm.weight"], sd[p + ".transform.LayerNorm.bias"])
    lm = h @ sd["bert.embeddings.word_embeddings.weight"].transpose(0, 1) + sd[p + ".bias"]
    q = "cls.imagePredictions"
    g = gelu_erf(linear(seq_v, sd, q + ".transform.dense"))
    g = layer_norm(g, sd[q + ".transform.LayerNorm.weight"], sd[q + ".transform.LayerNorm.bias"])
    vis = linear(g, sd, q + ".decoder")
    pooled = _drop(pooled_t * pooled_v if fusion == "mul" else pooled_t + pooled_v)
    rel = linear(pooled, sd, "cls.bi_seq_relationship")
    return lm, vis, rel


def lily_forward(sd, cfg, args, tokens, feat, loc, segs=None, tmask=None, vmask=None) -> Dict[str, torch.Tensor]:
    """Lily.forward (lily.py:58-129), eval mode (dropout = identity)."""
    seq_t, seq_v, pt, pv = bert_model(sd, cfg, tokens, feat, loc, segs, tmask, vmask)
    lm, vis, _ = pretraining_heads(sd, seq_t, seq_v, pt, pv, cfg.get("fusion_method", "mul"))
    pooled = _drop(pt * pv if cfg.get("fusion_method", "mul") == "mul" else pt + pv)
    out = {}
    if args.ranking:
        out["ranking"] = linear(pooled, sd, "vil_logit")
    if args.traj_judge:
        out["traj"] = linear(pooled, sd, "judge")
    if args.masked_vision:
        out["vision"] = vis
    if args.masked_language:
        out["language"] = lm
    return out


def pad_packed(t, mask):
    """utils/dataset/common.py:21-26."""
    mask = mask.bool()
    out = torch.full(mask.shape, -float("inf"), dtype=t.dtype, device=t.device)
    out = out.masked_scatter(mask, t)
    return out


def losses(batch: List[torch.Tensor], outputs: Dict[str, torch.Tensor], args, training: bool = True):
    """get_loss_correct (utils/utils_init.py:108-164) for every active task -> dict of scalar losses."""
    opt_mask = batch[13]
    res = {}
    if "vision" in outputs:
        pred = outputs["vision"]
        pred = pred.reshape(-1, pred.shape[2])
        target = batch[4][opt_mask].flatten(0, 1).to(pred.dtype)
        tmask = batch[5][opt_mask].flatten()
        logp = torch.log_softmax(pred, dim=-1)
        kl = torch.where(target > 0, target * (torch.log(target.clamp_min(1e-300)) - logp), torch.zeros_like(logp))
        kl = kl * tmask.unsqueeze(-1).to(pred.dtype)
        res["vision"] = kl.sum() / max(1, int(tmask.sum().item()))
    if "language" in outputs:
        pred = outputs["language"]
        pred = pred.reshape(-1, pred.shape[-1])
        target = batch[8][opt_mask].flatten()
        keep = target != -1
        logp = torch.log_softmax(pred[keep], dim=-1)
        res["language"] = -logp.gather(1, target[keep][:, None]).mean()
    if "ranking" in outputs:
        pred = pad_packed(outputs["ranking"].squeeze(1), opt_mask)
        target = batch[0]
        if training:
            res["ranking"] = F.cross_entropy(pred, target, ignore_index=-1)
        else:
            res["ranking"] = F.binary_cross_entropy_with_logits(pred, target.to(pred.dtype))
    if "traj" in outputs:
        pred = pad_packed(outputs["traj"].squeeze(1), opt_mask)
        target = torch.zeros(pred.shape, dtype=torch.bool, device=pred.device)
        if not (args.ranking or args.not_traj_judge_data):
            target[:, 0] = 1
        elif args.pretrain:
            target[:, :(1 + args.num_negatives)] = 1
        else:
            target[:, :-args.num_negatives] = 1
        tf = target.to(pred.dtype)
        pos_weight = target.shape[1] / tf[0].sum() - 1
        # BCE-with-logits with pos_weight, written out: -(pw*y*log(sig(x)) + (1-y)*log(1-sig(x)))
        l = -(pos_weight * tf * F.logsigmoid(pred) + (1 - tf) * F.logsigmoid(-pred))
        res["traj"] = l.mean()
    return res


def total_loss(loss_dict, args):
    """train_epoch loss sum (utils/utils_init.py:217-224)."""
    tot = 0.0
    for k in ("vision", "language", "ranking"):
        if k in loss_dict:
            tot = tot + loss_dict[k]
    if "traj" in loss_dict:
        tot = tot + args.traj_loss_scale * loss_dict["traj"]
    return tot


def oracle_step(sd: Dict[str, torch.Tensor], cfg, args, batch, dtype=torch.float32, want_grads=True, clone=True,
                train_dropout=False):
    """Forward + all active losses (+ autograd backward).  Returns (outputs, loss_dict, total, grads).
    Runs on whatever device ``sd`` / ``batch`` live on (the host for parity checks; bench.py also times it on
    the GPU as the "stock PyTorch ops" comparison, there optionally with ``train_dropout``: p = 0.1 at all 94 sites)."""
    _DROP["p"] = float(cfg.get("hidden_dropout_prob", 0.1)) if train_dropout else 0.0
    try:
        return _oracle_step(sd, cfg, args, batch, dtype, want_grads, clone)
    finally:
        _DROP["p"] = 0.0


def _oracle_step(sd, cfg, args, batch, dtype, want_grads, clone):
    if clone:
        sd = {k: v.detach().to(dtype).clone().requires_grad_(want_grads) for k, v in sd.items()
              if not k.endswith("cls.predictions.decoder.weight")}
    else:
        for v in sd.values():
            v.grad = None
    m = batch[13]
    co = batch[11]
    tokens, feat, loc, segs, tmask, vmask = batch[6][m], batch[1][m], batch[2][m], batch[10][m], batch[7][m], batch[3][m]
    out = lily_forward(sd, cfg, args, tokens, feat.to(dtype), loc.to(dtype), segs, tmask, vmask)
    ld = losses(batch, out, args, training=True)
    tot = total_loss(ld, args)
    grads = {}
    if want_grads:
        tot.backward()
        grads = {k: v.grad for k, v in sd.items() if v.grad is not None}
    return {k: v.detach() for k, v in out.items()}, {k: v.detach() for k, v in ld.items()}, tot.detach(), grads
