"""Fused multi-tensor AdamW for the drop-in ("next" row 1).

Same update rule, hyper-parameters, param-group layout and ``state_dict`` keys (``step``, ``exp_avg``,
``exp_avg_sq``) as the reference's ``AdamW`` (vilbert/optimization.py:107-189), so ``get_optimization`` / resume
(vilbert/vilbert_init.py:7-70) keep working; ``step()`` is one kernel launch over all parameters instead of a Python
loop of ~10 ATen ops per tensor.  Parameters that are GEMM operands of the drop-in also get their bf16 hi/lo planes
refreshed by the same kernel, so the next forward skips its weight re-split pass.
CPU parameters use the plain per-tensor update (same arithmetic).
"""
import math
import struct
from typing import List

import torch
from torch.optim import Optimizer


class FusedAdamW(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[1]))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(eps))
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias))
        self._tables = {}

    def load_state_dict(self, state_dict):
        self._tables.clear()
        return super().load_state_dict(state_dict)

    def add_param_group(self, param_group):
        if hasattr(self, "_tables"):
            self._tables.clear()
        return super().add_param_group(param_group)

    def __setstate__(self, state):
        super().__setstate__(state)
        self._tables = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1
            cuda = [p for p in ps if p.is_cuda]
            host = [p for p in ps if not p.is_cuda]
            if cuda:
                self._step_cuda(gi, group, cuda)
            for p in host:
                self._step_host(group, p)
        return loss

    @staticmethod
    def _step_size(group, step):
        s = group["lr"]
        if group["correct_bias"]:
            b1, b2 = group["betas"]
            s = s * math.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
        return s

    def _step_host(self, group, p):
        st = self.state[p]
        b1, b2 = group["betas"]
        g = p.grad
        st["exp_avg"].mul_(b1).add_(g, alpha=1.0 - b1)
        st["exp_avg_sq"].mul_(b2).addcmul_(g, g, value=1.0 - b2)
        denom = st["exp_avg_sq"].sqrt().add_(group["eps"])
        p.addcdiv_(st["exp_avg"], denom, value=-self._step_size(group, st["step"]))
        if group["weight_decay"] > 0.0:
            p.add_(p, alpha=-group["lr"] * group["weight_decay"])

    def _step_cuda(self, gi, group, ps: List[torch.Tensor]):
        from . import lib, ops
        steps = {self.state[p]["step"] for p in ps}
        dev = ps[0].device
        by_step = {}
        for p in ps:
            by_step.setdefault(self.state[p]["step"], []).append(p)
        for step, plist in by_step.items():
            arena = ops.rt(dev).arena
            # every raw pointer the kernel will dereference is part of the key: a replaced state tensor
            # (load_state_dict), a re-allocated gradient or a new arena entry all rebuild the table
            key = (gi, step == max(steps), arena.generation,
                   tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                          self.state[p]["exp_avg_sq"].data_ptr()) for p in plist))
            tab = self._tables.get(key)
            if tab is None:
                planes = {}
                arena.prune()
                for e in arena.entries.values():          # GEMM weights: where their hi/lo planes live
                    off = 0
                    for q in e.params:
                        planes[q.data_ptr()] = (e.planes.addr + 2 * off, e.planes.addr + 2 * (off + e.planes.plane_stride))
                        off += q.numel()
                rows, blk = [], 0
                for p in plist:
                    st = self.state[p]
                    if not (p.is_contiguous() and p.grad.is_contiguous() and p.dtype == torch.float32):
                        raise RuntimeError("FusedAdamW needs contiguous fp32 parameters and gradients")
                    hi, lo = planes.get(p.data_ptr(), (0, 0))
                    rows.append(struct.pack("<QQQQQQqqfi", p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(),
                                            st["exp_avg_sq"].data_ptr(), hi, lo, p.numel(), blk,
                                            float(group["weight_decay"]), 0))
                    blk += (p.numel() + 2047) // 2048
                raw = torch.frombuffer(bytearray(b"".join(rows)), dtype=torch.uint8).to(dev)
                keep = [t for p in plist for t in (self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"])]
                tab = (raw, len(rows), blk, torch.empty(7, dtype=torch.float32, device=dev), (plist, keep))
                if len(self._tables) > 8:
                    self._tables.clear()
                self._tables[key] = tab
            raw, nseg, blocks, hyper, _ = tab
            b1, b2 = group["betas"]
            hyper.copy_(torch.tensor([group["lr"], self._step_size(group, step), b1, b2, group["eps"], 1.0 - b1, 1.0 - b2],
                                     dtype=torch.float32), non_blocking=True)
            lib.adamw_multi(raw, nseg, blocks, hyper)
            # the kernel writes weights and planes through raw pointers (tensor versions do not move); a parameter
            # without planes in this table (its arena entry did not exist yet) is re-split by the forced start-of-
            # forward refresh of BertModel.forward, like with any other optimizer
