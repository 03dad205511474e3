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


# ----------------------------------------------------------------------------------------------------
# prediction heads + losses over the supervised rows only
# ----------------------------------------------------------------------------------------------------
def head_capacity(rows: int) -> int:
    """Static row capacity of the compacted heads: 30 % of the rows, rounded up to whole 128-row GEMM tiles (masked-LM /
    masked-region objectives supervise ~15 % of the positions; utils/dataset/common.py:213-300)."""
    cap = (int(0.3 * rows) + 127) // 128 * 128
    return max(128, min(cap, (rows + 127) // 128 * 128)) if rows > 128 else rows


def compact_rows(sel: torch.Tensor, cap: int):
    """Indices of the selected rows first (stable), truncated to ``cap`` -- a static shape, so the step stays
    capturable.  Returns (idx [cap], overflow flag as a 0 / NaN scalar to add to the loss): rows beyond the selected
    ones are fillers whose targets are "ignore", so they change nothing; if more than ``cap`` rows are selected the loss
    is poisoned with NaN instead of silently dropping supervision."""
    order = torch.argsort((~sel).to(torch.uint8), stable=True)
    idx = order[:cap]
    over = sel.sum() > cap
    poison = torch.where(over, torch.full((), float("nan"), device=sel.device), torch.zeros((), device=sel.device))
    return idx, poison


def language_head_loss(head, seq_t: torch.Tensor, targets: torch.Tensor, cap: int = None) -> torch.Tensor:
    """``cross_entropy(BertLMPredictionHead(seq_t), targets, ignore_index=-1)`` (vilbert/vilbert.py:889-907 +
    utils/utils_init.py:129-135) evaluated on the supervised rows only: the transform, the 768 -> 30522 decoder, the
    softmax passes and all three backward products shrink from N*T rows to ``cap`` rows (d(logits) of an ignored row is
    exactly zero, so nothing is approximated).  ``head`` is the model's own ``cls.predictions`` module."""
    rows = seq_t.reshape(-1, seq_t.shape[-1])
    t = targets.reshape(-1)
    cap = cap or head_capacity(rows.shape[0])
    if cap >= rows.shape[0]:
        return CELossFn.apply(_rows(head(rows)), t)
    idx, poison = compact_rows(t != -1, cap)
    logits = head(rows.index_select(0, idx))
    return CELossFn.apply(_rows(logits), t.index_select(0, idx)) + poison


def vision_head_loss(head, seq_v: torch.Tensor, target: torch.Tensor, mask: torch.Tensor, cap: int = None) -> torch.Tensor:
    """Masked-region KL loss (vilbert/vilbert.py:957-969 + utils/utils_init.py:117-128) on the masked regions only."""
    rows = seq_v.reshape(-1, seq_v.shape[-1])
    tg = target.reshape(-1, target.shape[-1])
    m = mask.reshape(-1)
    cap = cap or head_capacity(rows.shape[0])
    if cap >= rows.shape[0]:
        return KLLossFn.apply(_rows(head(rows)), tg, m)
    idx, poison = compact_rows(m != 0, cap)
    logits = head(rows.index_select(0, idx))
    return KLLossFn.apply(_rows(logits), tg.index_select(0, idx), m.index_select(0, idx)) + poison


def small_losses(batch: List[torch.Tensor], outputs: Dict[str, torch.Tensor], args, training: bool = True,
                 correct: Dict[str, torch.Tensor] = None):
    """Ranking / traj losses on ``[bs, C]`` logits computed for EVERY candidate slot (static shapes): padded
    candidates are sent to -inf exactly like the reference's ``pad_packed`` (utils/dataset/common.py:21-26).
    ``correct`` (optional dict) receives the accuracy counters of utils/utils_init.py:140-162 as device scalars."""
    opt_mask = batch[13].bool()
    bs, C = opt_mask.shape
    outs = {}
    for k in ("ranking", "traj"):
        if k in outputs:
            outs[k] = outputs[k].reshape(bs, C).masked_fill(~opt_mask, float("-inf"))
    return _host_small_flat(batch, outs, args, training, bs, C, correct)


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


def _host_small_flat(batch, outputs, args, training, bs, C, correct=None):
    import torch.nn.functional as F
    res = {}
    if "ranking" in outputs:
        pred = outputs["ranking"].reshape(bs, C)
        if training:
            res["ranking"] = F.cross_entropy(pred, batch[0], ignore_index=-1)
            if correct is not None:
                correct["ranking"] = (pred.detach().argmax(1) == batch[0]).sum().float()
        else:
            res["ranking"] = F.binary_cross_entropy_with_logits(pred, batch[0].float())
            if correct is not None:
                correct["ranking"] = batch[0].gather(1, pred.detach().argmax(1).view(-1, 1)).sum().float()
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
        if correct is not None:
            correct["traj"] = ((pred.detach().sigmoid() > 0.5) == target.bool()).sum().float() / C
    return res
