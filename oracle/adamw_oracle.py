"""TEST INFRASTRUCTURE ONLY -- host restatement of the reference optimizer update
(vilbert/optimization.py:141-187: bias-corrected Adam, eps added after the square root, decoupled weight decay
applied to the already-updated weight; param grouping of vilbert/vilbert_init.py:9-18).
Pinned against the real reference optimizer by ``oracle/make_golden.py adamw`` -> tests/golden/adamw.npz."""
import math

import torch

NO_DECAY = ("bias", "LayerNorm.weight", "LayerNorm.bias")


def weight_decay_of(name: str, wd: float) -> float:
    return 0.0 if any(nd in name for nd in NO_DECAY) else wd


def adamw_step(p, g, m, v, step, lr, wd, beta1=0.9, beta2=0.999, eps=1e-6, correct_bias=True):
    """One update; returns (p, m, v) as new tensors (plain dense algebra, fp32)."""
    m = m * beta1 + (1.0 - beta1) * g
    v = v * beta2 + (1.0 - beta2) * g * g
    denom = v.sqrt() + eps
    step_size = lr
    if correct_bias:
        step_size = step_size * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    p = p - step_size * (m / denom)
    if wd > 0.0:
        p = p - lr * wd * p
    return p, m, v
