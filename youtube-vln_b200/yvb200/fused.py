"""Fused loss kernels of the training step (utils/utils_init.py:117-135 of the reference) as autograd Functions.

  language : cross_entropy(logits[R, vocab], target[R], ignore_index=-1)              (:129-135)
  vision   : sum(kl_div(log_softmax(logits), target) * mask) / max(1, sum(mask))      (:117-128)

One block per row computes the log-sum-exp and the loss term without materialising probabilities; the backward
kernel writes d(logits) directly (scaled by the upstream gradient and 1/count on the device -- no ``.item()``
sync, unlike utils/utils_init.py:127).  The ranking / traj losses act on ``[bs, C]`` logits (a few dozen floats)
and stay on ATen (``yvb200.losses``).
"""
from typing import Dict, List

import torch
from torch.autograd import Function

from . import lib as L
from . import losses as _host


class CELossFn(Function):
    @staticmethod
    def forward(ctx, logits, target):
        R, V = logits.shape
        assert logits.stride(1) == 1 and logits.dtype == torch.float32     # rows may be padded (stride(0) >= V)
        target = target.contiguous().long()
        acc = torch.zeros(2, dtype=torch.float32, device=logits.device)     # {loss_sum, count}
        L.ce_loss(logits, logits.stride(0), target, R, V, acc[0:1], acc[1:2])
        ctx.save_for_backward(logits, target, acc)
        return acc[0] / acc[1].clamp_min(1.0)

    @staticmethod
    def backward(ctx, g):
        logits, target, acc = ctx.saved_tensors
        R, V = logits.shape
        ld = logits.stride(0)
        dl = torch.empty((R, ld), dtype=torch.float32, device=logits.device)
        L.ce_grad(logits, ld, target, R, V, acc[1:2], g.contiguous().float().reshape(1), dl, None)
        return dl[:, :V], None


class KLLossFn(Function):
    @staticmethod
    def forward(ctx, logits, target, mask):
        R, Cc = logits.shape
        assert logits.stride(1) == 1 and logits.dtype == torch.float32
        target = target.contiguous().float()
        mask = mask.contiguous().long()
        acc = torch.zeros(2, dtype=torch.float32, device=logits.device)
        L.kl_loss(logits, logits.stride(0), target, Cc, mask, R, Cc, acc[0:1], acc[1:2])
        ctx.save_for_backward(logits, target, mask, acc)
        return acc[0] / acc[1].clamp_min(1.0)

    @staticmethod
    def backward(ctx, g):
        logits, target, mask, acc = ctx.saved_tensors
        R, Cc = logits.shape
        ld = logits.stride(0)
        dl = torch.empty((R, ld), dtype=torch.float32, device=logits.device)
        L.kl_grad(logits, ld, target, Cc, mask, R, Cc, acc[1:2], g.contiguous().float().reshape(1), dl, None)
        return dl[:, :Cc], None, None


def _rows(logits: torch.Tensor) -> torch.Tensor:
    """[..., V] -> 2-D view; keeps the padded row pitch of the drop-in's logits (no copy)."""
    if logits.dim() == 2:
        return logits
    try:
        v = logits.view(-1, logits.shape[-1])
        if v.stride(1) == 1:
            return v
    except RuntimeError:
        pass
    return logits.reshape(-1, logits.shape[-1]).contiguous()


def language_loss(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return CELossFn.apply(_rows(logits), target.reshape(-1))


def vision_loss(logits: torch.Tensor, target: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    return KLLossFn.apply(_rows(logits), target.reshape(-1, target.shape[-1]), mask.reshape(-1))


def step_losses(batch: List[torch.Tensor], outputs: Dict[str, torch.Tensor], args, training: bool = True,
                flat: bool = False):
    """Same dict as ``yvb200.losses.step_losses`` with the two big losses on the fused kernels.  ``flat=True``
    means every candidate is valid (opt_mask all ones) so the [bs, C] -> [N] flatten is a reshape, which keeps
    the step free of data-dependent shapes (CUDA-graph capturable)."""
    opt_mask = batch[13]
    sel = (lambda t: t.flatten(0, 1)) if flat else (lambda t: t[opt_mask])
    res = {}
    if "vision" in outputs:
        res["vision"] = vision_loss(outputs["vision"], sel(batch[4]), sel(batch[5]))
    if "language" in outputs:
        res["language"] = language_loss(outputs["language"], sel(batch[8]))
    small = {k: v for k, v in outputs.items() if k in ("ranking", "traj")}
    if flat and small:
        bs, C = opt_mask.shape
        small = {k: v for k, v in small.items()}
        fake = list(batch)
        res.update(_host_small_flat(fake, small, args, training, bs, C))
    else:
        res.update(_host.step_losses(batch, small, args, training))
    return res


def _host_small_flat(batch, outputs, args, training, bs, C):
    import torch.nn.functional as F
    res = {}
    if "ranking" in outputs:
        pred = outputs["ranking"].reshape(bs, C)
        res["ranking"] = (F.cross_entropy(pred, batch[0], ignore_index=-1) if training
                          else F.binary_cross_entropy_with_logits(pred, batch[0].float()))
    if "traj" in outputs:
        pred = outputs["traj"].reshape(bs, C)
        if not (args.ranking or args.not_traj_judge_data):
            n_pos = 1
        elif args.pretrain:
            n_pos = 1 + args.num_negatives
        else:
            n_pos = C - args.num_negatives
        target = torch.zeros(bs, C, device=pred.device)
        target[:, :n_pos] = 1
        pos_weight = torch.full((1,), C / n_pos - 1, device=pred.device)
        res["traj"] = F.binary_cross_entropy_with_logits(pred, target, pos_weight=pos_weight)
    return res
