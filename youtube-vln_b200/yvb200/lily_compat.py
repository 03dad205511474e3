"""Host-side stand-in for the reference's task wrapper ``Lily`` (lily.py:23-129).

In a YouTube-VLN checkout the *unmodified* ``lily.py`` is used (it only needs ``vilbert.vilbert`` to resolve to the
drop-in).  The GPU box has no checkout, so tests and ``bench.py`` use this restatement: same sub-module names
(``bert``, ``cls``, ``vil_logit``, ``judge``), hence the same 542-key state dict, same forward signature and same
output dict.  The two ``Linear(1024, 1)`` heads on ``[N, 1024]`` are a few KFLOP and stay on ATen, as in lily.py.
"""
from typing import Dict

import torch

from vilbert.vilbert import BertModel, BertPreTrainedModel, BertPreTrainingHeads


class Lily(BertPreTrainedModel):
    def __init__(self, config, dropout_prob=0.1):
        super().__init__(config)
        self.args = config.args
        self.bert = BertModel(config)
        self.cls = BertPreTrainingHeads(config, self.bert.embeddings.word_embeddings.weight)
        self.vil_logit = torch.nn.Linear(config.bi_hidden_size, 1)
        self.judge = torch.nn.Linear(config.bi_hidden_size, 1)
        self.dropout = torch.nn.Dropout(dropout_prob)
        self.fusion_method = config.fusion_method
        self.apply(self.init_bert_weights)

    def forward(self, instr_tokens, image_features, image_locations, token_type_ids=None, attention_mask=None,
                image_attention_mask=None, co_attention_mask=None, highlight_tokens=None,
                order_atteneded_visual_feature=None) -> Dict[str, torch.Tensor]:
        seq_t, seq_v, pooled_t, pooled_v, _ = self.bert(
            input_txt=instr_tokens, input_imgs=image_features, image_loc=image_locations,
            token_type_ids=token_type_ids, attention_mask=attention_mask, image_attention_mask=image_attention_mask,
            co_attention_mask=co_attention_mask, output_all_encoded_layers=False)
        lang, vis, _ = self.cls(seq_t, seq_v, pooled_t, pooled_v)
        if self.fusion_method == "sum":
            pooled = pooled_t + pooled_v
        elif self.fusion_method == "mul":
            pooled = pooled_t * pooled_v
        else:
            assert False
        pooled = self.dropout(pooled)
        out = {}
        if self.args.ranking:
            out["ranking"] = self.vil_logit(pooled)
        if self.args.traj_judge:
            out["traj"] = self.judge(pooled)
        if self.args.masked_vision:
            out["vision"] = vis
        if self.args.masked_language:
            out["language"] = lang
        return out


def build_lily(cfg: dict, args, seed_weights: int = 0, device="cpu"):
    """``Lily`` on ``BertConfig(**cfg)`` with the name-keyed synthetic weights of ``yvb200.synth``."""
    from vilbert.vilbert import BertConfig
    from . import synth
    config = BertConfig(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in cfg.items()})
    config.args = args
    model = Lily(config)
    synth.load_synthetic_weights(model, seed=seed_weights)
    return model.to(device)
