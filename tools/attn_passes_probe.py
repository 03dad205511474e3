"""Developer probe (GPU box): how far do the outputs / losses / parameter gradients of the full step move when the
attention contractions (QK^T, PV and their backward products) run as single-pass bf16 instead of bf16x3?
Prints, per workload, the largest relative error against the golden vectors recorded from the reference.
    python tools/attn_passes_probe.py [workload ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "youtube-vln_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)

from yvb200 import losses, ops, synth  # noqa: E402
from yvb200.lily_compat import build_lily  # noqa: E402


def errors(g, out, grads):
    worst_out, worst_grad, worst_name = 0.0, 0.0, ""
    for k, v in out.items():
        a = v.float().numpy()
        if f"out/{k}" in g.files:
            ref = g[f"out/{k}"]
            err = np.linalg.norm(a - ref) / max(np.linalg.norm(ref), 1e-30)
        else:
            ref = g[f"outval/{k}"]
            got = a.reshape(-1)[g[f"outpos/{k}"]]
            err = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
        worst_out = max(worst_out, float(err))
    norms = [float(np.linalg.norm(g[k])) if k.startswith("grad/") else float(g[k])
             for k in g.files if k.startswith("grad/") or k.startswith("gradnorm/")]
    floor = 1e-6 * max(norms)
    for name, gr in grads.items():
        a = gr.float().numpy()
        if f"grad/{name}" in g.files:
            ref = g[f"grad/{name}"]
            if np.linalg.norm(ref) < floor:
                continue
            err = np.linalg.norm(a - ref) / max(np.linalg.norm(ref), 1e-30)
        elif f"gradval/{name}" in g.files:
            ref = g[f"gradval/{name}"]
            norm = float(g[f"gradnorm/{name}"])
            if norm < floor:
                continue
            got = a.reshape(-1)[g[f"gradpos/{name}"]]
            err = np.linalg.norm(got - ref) / max(np.sqrt(len(ref)) * norm / np.sqrt(a.size), 1e-30)
        else:
            continue
        if err > worst_grad:
            worst_grad, worst_name = float(err), name
    return worst_out, worst_grad, worst_name


def run(wl, attn_passes):
    r = ops.rt("cuda")
    r.set_precision("bf16x3")
    r.attn_passes = attn_passes
    cfg = synth.CONFIGS[synth.WORKLOADS[wl]["config"]]
    args = synth.workload_args(wl)
    model = build_lily(cfg, args, device="cuda").eval()
    batch = [t.cuda() if torch.is_tensor(t) else t for t in synth.make_batch(wl, seed=1)]
    out = model(*synth.model_inputs(batch))
    ld = losses.step_losses(batch, out, args, training=True)
    losses.total_loss(ld, args).backward()
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().cpu() for n, p in model.named_parameters() if p.grad is not None}
    return {k: v.detach().cpu() for k, v in out.items()}, grads


if __name__ == "__main__":
    for wl in (sys.argv[1:] or ["micro", "cfg1", "cfg2"]):
        g = np.load(os.path.join(ROOT, "tests", "golden", f"{wl}.npz"))
        for passes in (3, 1):
            out, grads = run(wl, passes)
            wo, wg, name = errors(g, out, grads)
            print(f"{wl}: attention passes={passes}: worst output err {wo:.2e}, worst gradient err {wg:.2e} ({name})",
                  flush=True)
