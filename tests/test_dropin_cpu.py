"""CPU (BASELINE config 1 plumbing): the drop-in ``vilbert.vilbert`` host path against the reference's golden
vectors and the oracle; surface / state-dict / error-behaviour checks."""
import os

import numpy as np
import pytest
import torch

from yvb200 import synth, losses
from yvb200.lily_compat import build_lily
import vilbert.vilbert as V
from test_oracle_golden import _check_grads, _check_outputs, _load


@pytest.mark.parametrize("wl", ["micro", "cfg1"])
def test_dropin_cpu_matches_reference_golden(golden_dir, wl):
    g = _load(golden_dir, wl)
    cfg = synth.CONFIGS[synth.WORKLOADS[wl]["config"]]
    args = synth.workload_args(wl)
    model = build_lily(cfg, args).eval()
    batch = synth.make_batch(wl, seed=1)
    out = model(*synth.model_inputs(batch))
    ld = losses.step_losses(batch, out, args, training=True)
    tot = losses.total_loss(ld, args)
    tot.backward()
    for k, v in ld.items():
        assert abs(float(v) - float(g[f"loss/{k}"])) <= 2e-5 * max(1.0, abs(float(g[f"loss/{k}"]))), k
    _check_outputs(g, {k: v.detach() for k, v in out.items()}, 2e-5)
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    assert _check_grads(g, grads, 2e-4) > 100
    dead = {k[len("nograd/"):] for k in g.files if k.startswith("nograd/")}
    assert dead == {n for n, p in model.named_parameters() if p.grad is None}


def test_state_dict_schema_matches_reference_layout():
    cfg = synth.FULL_CONFIG
    # build on the meta device: no 1 GB allocation needed to compare names and shapes
    with torch.device("meta"):
        config = V.BertConfig(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in cfg.items()})
        config.args = synth.make_args()
        from yvb200.lily_compat import Lily
        model = Lily(config)
    sd = model.state_dict()
    want = synth.lily_param_shapes(cfg)
    assert len(sd) == 542
    assert {k: tuple(v.shape) for k, v in sd.items()} == want
    assert sum(p.numel() for p in model.parameters()) == 250_087_039
    assert model.cls.predictions.decoder.weight is model.bert.embeddings.word_embeddings.weight


def test_public_surface_and_errors(tmp_path):
    names = ("BertConfig BertPreTrainedModel BertModel BertPreTrainingHeads VILBertForVLTasks BertLayerNorm "
             "BertEmbeddings BertSelfAttention BertSelfOutput BertAttention BertIntermediate BertOutput BertLayer "
             "BertImageSelfAttention BertImageSelfOutput BertImageAttention BertImageIntermediate BertImageOutput "
             "BertImageLayer BertImageEmbeddings BertImagePooler BertImagePredictionHead BertBiAttention BertBiOutput "
             "BertConnectionLayer BertEncoder BertTextPooler BertPredictionHeadTransform "
             "BertImgPredictionHeadTransform BertLMPredictionHead BertOnlyMLMHead BertOnlyNSPHead "
             "BertForMultiModalPreTraining SimpleClassifier gelu swish ACT2FN").split()
    for n in names:
        assert hasattr(V, n), n
    with pytest.raises(ValueError):
        V.BertSelfAttention(V.BertConfig(hidden_size=100, num_attention_heads=12))
    with pytest.raises(AssertionError):
        V.BertConfig(v_biattention_id=(0, 5), v_num_hidden_layers=3)
    c = V.BertConfig(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in synth.MICRO_CONFIG.items()})
    c.args = synth.make_args()
    assert "hidden_size" in c.to_dict() and "args" in c.to_json_string()
    p = tmp_path / "cfg.json"
    p.write_text(__import__("json").dumps(synth.MICRO_CONFIG))
    assert V.BertConfig.from_json_file(str(p)).hidden_size == 64
    with pytest.raises(RuntimeError):
        V.BertModel.from_pretrained(str(tmp_path / "missing.bin"), c)


def test_from_pretrained_roundtrip_and_7tuple(tmp_path):
    cfg = synth.MICRO_CONFIG
    args = synth.make_args()
    model = build_lily(cfg, args).eval()
    ck = tmp_path / "lily.bin"
    sd = model.state_dict()
    sd = {k.replace("LayerNorm.weight", "LayerNorm.gamma").replace("LayerNorm.bias", "LayerNorm.beta"): v
          for k, v in sd.items()}                           # TF-era names must be accepted
    torch.save({"model_state_dict": sd, "epoch": 3}, ck)
    from yvb200.lily_compat import Lily
    config = model.config
    m2 = Lily.from_pretrained(str(ck), config).eval()
    for (k1, v1), (k2, v2) in zip(model.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    # a bare BertModel loads the "bert."-prefixed trunk out of the same file
    trunk = V.BertModel.from_pretrained(str(ck), config)
    assert torch.equal(trunk.t_pooler.dense.weight, model.bert.t_pooler.dense.weight)
    vl = V.VILBertForVLTasks(config, num_labels=2).eval()
    batch = synth.make_batch("micro")
    tok, feat, loc, seg, tm, vm, co, _, _ = synth.model_inputs(batch)
    outs = vl(tok, feat, loc, seg, tm, vm.float(), None)
    n, vlen, tlen = feat.shape[0], feat.shape[1], tok.shape[1]
    assert [tuple(o.shape) for o in outs] == [(n, 2), (n, 1), (n, 2), (n, vlen, cfg["v_target_size"]), (n, vlen, 1),
                                              (n, tlen, cfg["vocab_size"]), (n, tlen, 1)]
