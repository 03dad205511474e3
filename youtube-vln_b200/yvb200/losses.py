"""Losses of the training step (utils/utils_init.py:108-164, :217-224 of the reference).

``step_losses`` is the host restatement on ATen ops that consumes the logits of the drop-in exactly as the
reference's ``get_loss_correct`` does (used by the parity tests on both devices).  The fused CUDA versions
(``yvb200.fused``) produce the same numbers without materialising probability tensors.
"""
from typing import Dict, List

import torch
import torch.nn.functional as F


def pad_packed(t: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Scatter packed per-pair values back to ``[bs, C]`` with -inf at padded candidates (common.py:21-26)."""
    mask = mask.bool()
    out = torch.full(mask.shape, float("-inf"), dtype=t.dtype, device=t.device)
    return out.masked_scatter(mask, t)


def step_losses(batch: List[torch.Tensor], outputs: Dict[str, torch.Tensor], args, training: bool = True):
    opt_mask = batch[13]
    res = {}
    if "vision" in outputs:
        pred = outputs["vision"]
        pred = pred.reshape(-1, pred.shape[2])
        target = batch[4][opt_mask].flatten(0, 1)
        tmask = batch[5][opt_mask].flatten()
        loss = F.kl_div(F.log_softmax(pred, dim=-1), target, reduction="none") * tmask.unsqueeze(-1).float()
        res["vision"] = loss.sum() / torch.clamp(tmask.sum(), min=1)
    if "language" in outputs:
        pred = outputs["language"]
        res["language"] = F.cross_entropy(pred.reshape(-1, pred.shape[-1]), batch[8][opt_mask].flatten(), ignore_index=-1)
    if "ranking" in outputs:
        pred = pad_packed(outputs["ranking"].squeeze(1), opt_mask)
        if training:
            res["ranking"] = F.cross_entropy(pred, batch[0], ignore_index=-1)
        else:
            res["ranking"] = F.binary_cross_entropy_with_logits(pred, batch[0].float())
    if "traj" in outputs:
        pred = pad_packed(outputs["traj"].squeeze(1), opt_mask)
        target = torch.zeros(pred.shape, device=pred.device, dtype=torch.bool)
        if not (args.ranking or args.not_traj_judge_data):
            target[:, 0] = 1
        elif args.pretrain:
            target[:, :(1 + args.num_negatives)] = 1
        else:
            target[:, :-args.num_negatives] = 1
        n_pos = int(1 if not (args.ranking or args.not_traj_judge_data) else
                    ((1 + args.num_negatives) if args.pretrain else target.shape[1] - args.num_negatives))
        pos_weight = torch.tensor([target.shape[1] / n_pos - 1], device=pred.device)
        res["traj"] = F.binary_cross_entropy_with_logits(pred, target.float(), pos_weight=pos_weight)
    return res


def total_loss(loss_dict, args):
    tot = 0.0
    for k in ("vision", "language", "ranking"):
        if k in loss_dict:
            tot = tot + loss_dict[k]
    if "traj" in loss_dict:
        tot = tot + args.traj_loss_scale * loss_dict["traj"]
    return tot
