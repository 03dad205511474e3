"""Deterministic synthetic configs, weights and batches for the ViLBERT hot path.

The recipe is SURVEY.md section 8(d): N(0,1) 2048-d region features, U(0,1) boxes with the frame index
in column 11 (reference layout: utils/dataset/all_dataset.py:309-317), random token ids with [CLS]=101
first, partial masks on odd rows, soft 1601-way region targets, BERT-style -1 ignore labels.  The batch
is the reference's 16-tuple (utils/utils_init.py:34-52, utils/dataset/all_dataset.py:275-292) so the
unmodified ``get_model_input`` / ``get_loss_correct`` can consume it.

Everything is generated on the CPU with explicit ``torch.Generator`` seeds so that the build container
(golden-vector generation with the real reference) and the GPU box (parity tests, bench) see identical
bits.  No file of the reference is needed at run time.
"""
from __future__ import annotations

import hashlib
import types
from typing import Dict, List, Tuple

import torch

# --------------------------------------------------------------------------------------------------
# model configs (JSON of data/config/bert_base_6_layer_6_connect.json is not in the reference repo;
# the structure is pinned by vilbert/vilbert.py:693-706 and :1331-1334)
# --------------------------------------------------------------------------------------------------
FULL_CONFIG = dict(
    vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
    intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1,
    attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=2,
    initializer_range=0.02, v_feature_size=2048, v_target_size=1601, v_hidden_size=1024,
    v_num_hidden_layers=6, v_num_attention_heads=8, v_intermediate_size=1024, bi_hidden_size=1024,
    bi_num_attention_heads=8, v_attention_probs_dropout_prob=0.1, v_hidden_act="gelu",
    v_hidden_dropout_prob=0.1, v_initializer_range=0.02,
    v_biattention_id=[0, 1, 2, 3, 4, 5], t_biattention_id=[6, 7, 8, 9, 10, 11],
    fusion_method="mul", with_coattention=True,
)

#: BASELINE.json configs[0]: full widths, 2 text + 2 vision + 2 connection layers
TINY_CONFIG = dict(FULL_CONFIG, num_hidden_layers=2, v_num_hidden_layers=2,
                   v_biattention_id=[0, 1], t_biattention_id=[0, 1])

#: narrow widths so that complete outputs and gradients fit in a committed fixture
MICRO_CONFIG = dict(
    vocab_size=320, hidden_size=64, num_hidden_layers=3, num_attention_heads=4,
    intermediate_size=128, hidden_act="gelu", hidden_dropout_prob=0.1,
    attention_probs_dropout_prob=0.1, max_position_embeddings=64, type_vocab_size=2,
    initializer_range=0.02, v_feature_size=96, v_target_size=40, v_hidden_size=128,
    v_num_hidden_layers=2, v_num_attention_heads=2, v_intermediate_size=128, bi_hidden_size=128,
    bi_num_attention_heads=2, v_attention_probs_dropout_prob=0.1, v_hidden_act="gelu",
    v_hidden_dropout_prob=0.1, v_initializer_range=0.02,
    v_biattention_id=[0, 1], t_biattention_id=[1, 2],
    fusion_method="mul", with_coattention=True,
)

CONFIGS = {"full": FULL_CONFIG, "tiny": TINY_CONFIG, "micro": MICRO_CONFIG}

#: named workloads: (model config, items bs, candidates C, frames P, regions/frame B, tokens T)
WORKLOADS = {
    "micro": dict(config="micro", bs=2, cands=4, frames=2, boxes=6, tokens=12),
    # degenerate masks: an instruction that is all padding (token id 0 = the embedding's padding_idx) except [CLS], a
    # fully padded trajectory, a pair with nothing masked
    "micro_pad": dict(config="micro", bs=2, cands=4, frames=2, boxes=6, tokens=12, degenerate=True),
    "cfg1": dict(config="tiny", bs=1, cands=2, frames=1, boxes=36, tokens=20, args=dict(pretrain=False)),
    "cfg2": dict(config="full", bs=2, cands=4, frames=8, boxes=36, tokens=80),
    "cfg3": dict(config="full", bs=4, cands=4, frames=8, boxes=36, tokens=80),
    "cfg4_p4": dict(config="full", bs=2, cands=4, frames=4, boxes=36, tokens=80),
    "cfg4_p16": dict(config="full", bs=2, cands=4, frames=16, boxes=36, tokens=80),
    "cfg4_p32": dict(config="full", bs=2, cands=4, frames=32, boxes=36, tokens=80),
    # 2-pair versions of the trajectory-length sweep: small enough for the host oracle inside a test
    "cfg4_p4_n2": dict(config="full", bs=1, cands=2, frames=4, boxes=36, tokens=80, args=dict(pretrain=False)),
    "cfg4_p16_n2": dict(config="full", bs=1, cands=2, frames=16, boxes=36, tokens=80, args=dict(pretrain=False)),
    "cfg4_p32_n2": dict(config="full", bs=1, cands=2, frames=32, boxes=36, tokens=80, args=dict(pretrain=False)),
    # fine-tune style: ranking objective only (BASELINE config 3 shape, 16 pairs)
    "cfg3_rank": dict(config="full", bs=4, cands=4, frames=8, boxes=36, tokens=80,
                      args=dict(pretrain=False, masked_vision=False, masked_language=False, traj_judge=False)),
}


def derive_workload(base: str, pairs: int) -> str:
    """Register (once) and name the variant of ``base`` with ``pairs`` trajectory-instruction pairs per batch: BASELINE
    configs 3 / 5 fix the GLOBAL batch (16 / 64 pairs), so the per-GPU batch depends on the number of ranks."""
    if pairs == WORKLOADS[base]["bs"] * WORKLOADS[base]["cands"]:
        return base
    name = f"{base}@{pairs}"
    if name not in WORKLOADS:
        cands = 4 if pairs % 4 == 0 else (2 if pairs % 2 == 0 else 1)
        w = dict(WORKLOADS[base], bs=pairs // cands, cands=cands)
        if cands < 3:                       # two candidates: fine-tune style traj target (as in cfg1)
            w["args"] = dict(w.get("args", {}), pretrain=False)
        WORKLOADS[name] = w
    return name


def workload_args(workload: str, **over) -> types.SimpleNamespace:
    """``make_args`` with the per-workload overrides (cfg1 has 2 candidates -> fine-tune style traj target)."""
    kw = dict(WORKLOADS[workload].get("args", {}))
    kw.update(over)
    return make_args(**kw)


def make_args(**over) -> types.SimpleNamespace:
    """The slice of the reference's argparse namespace that ``Lily`` and the losses read
    (lily.py:27-30,117-127; utils/utils_init.py:147-158; utils/cli.py defaults)."""
    a = dict(model_name="vilbert", ranking=True, traj_judge=True, masked_vision=True,
             masked_language=True, pretrain=True, num_negatives=1, not_traj_judge_data=False,
             traj_loss_scale=1.0, gradient_accumulation_steps=1, skip_all_reduce=True,
             local_rank=-1, shuffle_visual_features=False)
    a.update(over)
    return types.SimpleNamespace(**a)


def _seed_of(name: str, seed: int) -> int:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return int.from_bytes(h[:7], "little")


def make_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Name-keyed deterministic weights (independent of module construction order).

    Matrices/embeddings ~ N(0, 0.02) like ``init_bert_weights`` (vilbert/vilbert.py:991-1002) but biases
    and LayerNorm parameters are *not* the trivial 0 / 1 of a fresh init, so that every bias, gamma and
    beta path is exercised by the parity tests.
    """
    out = {}
    for name, shape in shapes.items():
        g = torch.Generator().manual_seed(_seed_of(name, seed))
        if "LayerNorm" in name and name.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif "LayerNorm" in name and name.endswith("bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif name.endswith("weight_g"):           # SimpleClassifier weight_norm scalars
            t = 1.0 + 0.1 * torch.rand(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        out[name] = t.float()
    return out


def load_synthetic_weights(model: torch.nn.Module, seed: int = 0) -> None:
    """Overwrite every parameter of ``model`` with ``make_state_dict`` values (tied tensors once)."""
    sd = model.state_dict()
    new = make_state_dict({k: tuple(v.shape) for k, v in sd.items()}, seed)
    # tied LM decoder (vilbert/vilbert.py:896-901): both keys must carry the embedding values
    for k in list(new):
        if k.endswith("cls.predictions.decoder.weight"):
            emb = k.replace("cls.predictions.decoder.weight", "bert.embeddings.word_embeddings.weight")
            if emb in new:
                new[k] = new[emb]
    with torch.no_grad():
        for k, v in sd.items():
            v.copy_(new[k])


def make_batch(workload: str, seed: int = 1, rank: int = 0) -> List[torch.Tensor]:
    """Return the reference's 16-tuple batch (CPU tensors) for a named workload."""
    w = WORKLOADS[workload]
    cfg = CONFIGS[w["config"]]
    bs, C, P, B, T = w["bs"], w["cands"], w["frames"], w["boxes"], w["tokens"]
    V = P * B
    N = bs * C
    g = torch.Generator().manual_seed(seed + rank)
    feat = torch.randn(N, V, cfg["v_feature_size"], generator=g)
    loc = torch.rand(N, V, 12, generator=g)
    loc[..., 11] = torch.arange(P).repeat_interleave(B).float()[None, :]
    tok = torch.randint(1, cfg["vocab_size"], (N, T), generator=g)
    tok[:, 0] = 101 % cfg["vocab_size"]
    seg = torch.zeros(N, T, dtype=torch.long)
    tmask = torch.ones(N, T, dtype=torch.long)
    vmask = torch.ones(N, V, dtype=torch.long)
    tail = max(1, T // 16)
    tmask[1::2, T - tail:] = 0
    tok[1::2, T - tail:] = 0
    if P > 1:
        vmask[1::2, V - B:] = 0
    else:
        vmask[1::2, V - max(1, B // 6):] = 0
    img_tgt = torch.softmax(torch.randn(N, V, cfg["v_target_size"], generator=g), dim=-1)
    img_tgt_mask = (torch.rand(N, V, generator=g) < 0.15).long() * vmask
    img_tgt_mask[:, 1] = vmask[:, 1]                      # at least one supervised region per pair
    lm_sel = torch.rand(N, T, generator=g) < 0.15
    lm_sel[:, 1] = True
    lm_sel &= tmask.bool()
    lm_tgt = torch.where(lm_sel, tok, torch.full_like(tok, -1))
    co = torch.zeros(bs, 2, V, T, dtype=torch.long)
    opt_mask = torch.ones(bs, C, dtype=torch.bool)
    target = torch.zeros(bs, dtype=torch.long)
    hl = torch.zeros(N, T, dtype=torch.long)

    def r(t):                                             # [N, ...] -> [bs, C, ...]
        return t.view(bs, C, *t.shape[1:])

    batch = [None] * 16
    batch[0] = target
    batch[1] = r(feat)
    batch[2] = r(loc)
    batch[3] = r(vmask)
    batch[4] = r(img_tgt)
    batch[5] = r(img_tgt_mask)
    batch[6] = r(tok)
    batch[7] = r(tmask)
    batch[8] = r(lm_tgt)
    batch[9] = r(hl)
    batch[10] = r(seg)
    batch[11] = co
    batch[12] = torch.zeros(bs, dtype=torch.long)
    batch[13] = opt_mask
    if WORKLOADS[workload].get("degenerate"):
        batch[7][0, 0, 1:] = 0                 # instr_mask: only [CLS] left ...
        batch[6][0, 0, 1:] = 0                 # ... and the tokens are padding (id 0)
        batch[3][0, 1, :] = 0                  # image_mask: whole trajectory padded ...
        batch[5][0, 1, :] = 0                  # ... so none of its regions is supervised
        batch[7][1, 0, :] = 1                  # nothing masked
        batch[3][1, 0, :] = 1
    batch[14] = torch.zeros(bs, dtype=torch.long)
    batch[15] = torch.zeros(1)
    return batch


def model_inputs(batch: List[torch.Tensor]):
    """Own restatement of ``get_model_input`` (utils/utils_init.py:34-77): boolean-mask flatten
    ``[bs, C, ...] -> [N, ...]``; returns the 9 positional arguments of ``Lily.forward``."""
    m = batch[13]
    co = batch[11]
    return (batch[6][m], batch[1][m], batch[2][m], batch[10][m], batch[7][m], batch[3][m],
            co.reshape(-1, co.size(2), co.size(3)), batch[9][m], batch[15])


def num_pairs(batch) -> int:
    return int(batch[13].sum().item())


def lily_param_shapes(cfg: Dict) -> Dict[str, Tuple[int, ...]]:
    """State-dict schema of the reference's ``Lily`` (lily.py:23-56 + vilbert/vilbert.py module tree)
    derived from a config alone: 542 keys for the full config (the tied decoder weight appears twice)."""
    H, I, Hv, Iv, Hb = (cfg["hidden_size"], cfg["intermediate_size"], cfg["v_hidden_size"],
                        cfg["v_intermediate_size"], cfg["bi_hidden_size"])
    s: Dict[str, Tuple[int, ...]] = {}

    def lin(p, o, i):
        s[p + ".weight"] = (o, i)
        s[p + ".bias"] = (o,)

    def ln(p, h):
        s[p + ".weight"] = (h,)
        s[p + ".bias"] = (h,)

    e = "bert.embeddings"
    s[e + ".word_embeddings.weight"] = (cfg["vocab_size"], H)
    s[e + ".position_embeddings.weight"] = (cfg["max_position_embeddings"], H)
    s[e + ".token_type_embeddings.weight"] = (cfg["type_vocab_size"], H)
    ln(e + ".LayerNorm", H)
    v = "bert.v_embeddings"
    lin(v + ".image_embeddings", Hv, cfg["v_feature_size"])
    lin(v + ".image_location_embeddings", Hv, 5)
    lin(v + ".image_orientation_embeddings", Hv, 4)
    lin(v + ".image_next_orientation_embeddings", Hv, 2)
    s[v + ".image_sequence_embeddings.weight"] = (32, Hv)
    ln(v + ".LayerNorm", Hv)

    def block(p, h, inter):
        for n in ("query", "key", "value"):
            lin(f"{p}.attention.self.{n}", h, h)
        lin(p + ".attention.output.dense", h, h)
        ln(p + ".attention.output.LayerNorm", h)
        lin(p + ".intermediate.dense", inter, h)
        lin(p + ".output.dense", h, inter)
        ln(p + ".output.LayerNorm", h)

    for i in range(cfg["num_hidden_layers"]):
        block(f"bert.encoder.layer.{i}", H, I)
    for i in range(cfg["v_num_hidden_layers"]):
        block(f"bert.encoder.v_layer.{i}", Hv, Iv)
    for i in range(len(cfg["v_biattention_id"])):
        p = f"bert.encoder.c_layer.{i}"
        for n in ("query", "key", "value"):
            lin(f"{p}.biattention.{n}1", Hb, Hv)
        for n in ("query", "key", "value"):
            lin(f"{p}.biattention.{n}2", Hb, H)
        lin(p + ".biOutput.dense1", Hv, Hb)
        ln(p + ".biOutput.LayerNorm1", Hv)
        lin(p + ".biOutput.q_dense1", Hv, Hb)
        lin(p + ".biOutput.dense2", H, Hb)
        ln(p + ".biOutput.LayerNorm2", H)
        lin(p + ".biOutput.q_dense2", H, Hb)
        lin(p + ".v_intermediate.dense", Iv, Hv)
        lin(p + ".v_output.dense", Hv, Iv)
        ln(p + ".v_output.LayerNorm", Hv)
        lin(p + ".t_intermediate.dense", I, H)
        lin(p + ".t_output.dense", H, I)
        ln(p + ".t_output.LayerNorm", H)
    lin("bert.t_pooler.dense", Hb, H)
    lin("bert.v_pooler.dense", Hb, Hv)
    s["cls.predictions.bias"] = (cfg["vocab_size"],)
    lin("cls.predictions.transform.dense", H, H)
    ln("cls.predictions.transform.LayerNorm", H)
    s["cls.predictions.decoder.weight"] = (cfg["vocab_size"], H)
    lin("cls.bi_seq_relationship", 2, Hb)
    lin("cls.imagePredictions.transform.dense", Hv, Hv)
    ln("cls.imagePredictions.transform.LayerNorm", Hv)
    lin("cls.imagePredictions.decoder", cfg["v_target_size"], Hv)
    lin("vil_logit", 1, Hb)
    lin("judge", 1, Hb)
    return s


def lily_state_dict(cfg: Dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic weights for the whole ``Lily`` schema, with the LM decoder tied to the word embeddings."""
    sd = make_state_dict(lily_param_shapes(cfg), seed)
    sd["cls.predictions.decoder.weight"] = sd["bert.embeddings.word_embeddings.weight"]
    return sd
