"""Batch masking on the GPU (SURVEY.md 8f "next" #3) -- device twins of the reference's CPU data-path functions
``randomize_tokens`` / ``randomize_regions`` (utils/dataset/common.py:213-300, called per sample by the dataset at
utils/dataset/all_dataset.py:246-261).

``mask_tokens`` / ``mask_regions`` take the uniform draws as explicit tensors: with the draws of the reference they
reproduce its outputs bit for bit (tests/test_masking.py against golden vectors recorded from the reference).
``randomize_tokens`` / ``randomize_regions`` keep the reference's names, argument order and return values, draw their
numbers with torch's device generator, and work on whole batches already resident in HBM -- the per-sample CPU
masking and its float64 padding path no longer sit in front of a 10 ms GPU step.  CUDA tensors only: there is no host
fallback on this path (``lib.load()`` raises if the library is missing).
"""
from typing import Optional, Tuple

import numpy as np
import torch

from . import lib as L

ACTION_TOKENS = (2187, 2830, 2157)          # 'left', 'forward', 'right' (utils/dataset/common.py:215-222,236)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("yvb200.masking works on CUDA tensors only (no host fallback on this path)")


def mask_tokens(tokens: torch.Tensor, mask: torch.Tensor, p: torch.Tensor, random_tokens: torch.Tensor, mask_id: int,
                forced: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """In-place BERT masking of ``tokens`` (i64) given the draws; returns (tokens, targets) like the reference."""
    _need_cuda(tokens, mask, p, random_tokens, forced)
    L.load()
    if tokens.dtype != torch.int64 or not tokens.is_contiguous():
        raise RuntimeError("mask_tokens: tokens must be a contiguous int64 tensor")
    m8 = mask.to(torch.bool).contiguous().view(torch.uint8)
    f8 = None if forced is None else forced.to(torch.bool).contiguous().view(torch.uint8)
    targets = torch.empty_like(tokens)
    L.mask_tokens(tokens, m8, p.float().contiguous(), random_tokens.long().contiguous(), f8, int(mask_id), targets)
    return tokens, targets


def mask_regions(features: torch.Tensor, probs: torch.Tensor, mask: torch.Tensor, p: torch.Tensor):
    """In-place ViLBERT region masking; returns (features, targets, targets_mask) like the reference."""
    _need_cuda(features, probs, mask, p)
    L.load()
    if features.dtype != torch.float32 or not features.is_contiguous():
        raise RuntimeError("mask_regions: features must be a contiguous float32 tensor")
    F, Cc = features.shape[-1], probs.shape[-1]
    rows = features.numel() // F
    if probs.numel() // Cc != rows or mask.numel() != rows or p.numel() != rows:
        raise RuntimeError("mask_regions: features / probs / mask / p disagree on the number of regions")
    targets = torch.empty_like(probs, dtype=torch.float32)
    targets_mask = torch.empty_like(mask, dtype=torch.int64)
    L.mask_regions(features, probs.float().contiguous(), mask.long().contiguous(), p.float().contiguous(), targets,
                   targets_mask, rows, F, Cc)
    return features, targets, targets_mask.to(mask.dtype)


def randomize_tokens(tokens, mask, tokenizer, args):
    """Reference signature (utils/dataset/common.py:213): tokens randomly masked with the standard BERT probabilities."""
    _need_cuda(tokens, mask)
    p = torch.rand(tokens.shape, dtype=torch.float32, device=tokens.device)
    random_tokens = torch.randint(0, len(tokenizer.vocab), tokens.shape, dtype=torch.int64, device=tokens.device)
    forced = None
    rate = float(getattr(args, "mask_action_rate", 0.0) or 0.0)
    if rate > 0:
        # the reference picks action-word positions with np.random.choice (with replacement) on the host (:236-252)
        pos = torch.nonzero(torch.stack([tokens == a for a in ACTION_TOKENS]).any(0))
        # order the candidates as the reference does: all 'left' first, then 'forward', then 'right'
        if pos.numel():
            key = torch.zeros(pos.shape[0], dtype=torch.int64, device=tokens.device)
            tv = tokens[pos[:, 0], pos[:, 1]]
            for i, a in enumerate(ACTION_TOKENS):
                key[tv == a] = i
            pos = pos[torch.argsort(key, stable=True)]
            pick = np.random.choice(range(pos.shape[0]), int(rate * pos.shape[0]))
            forced = torch.zeros_like(tokens, dtype=torch.bool)
            if len(pick):
                sel = pos[torch.as_tensor(pick, device=tokens.device)]
                forced[sel[:, 0], sel[:, 1]] = True
    return mask_tokens(tokens, mask, p, random_tokens, tokenizer.vocab["[MASK]"], forced)


def randomize_regions(features, probs, mask):
    """Reference signature (utils/dataset/common.py:272): features masked with the ViLBERT probabilities."""
    _need_cuda(features, probs, mask)
    p = torch.rand(mask.shape, dtype=torch.float32, device=mask.device)
    return mask_regions(features, probs, mask, p)
