"""Developer tool: does the captured train-mode step reproduce itself?  Replays the cfg2 step several times from the same
counter-RNG state (and the same torch generator state for the task wrapper's pooled dropout) and prints the largest
relative difference of the loss, of the forward outputs and of every gradient between replays.  Differences at the 1e-7
level are reduction-order noise (split-K / slab sums); anything larger is a race."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
import torch  # noqa: E402

from yvb200 import ops, synth  # noqa: E402
from yvb200.lily_compat import build_lily  # noqa: E402
from yvb200.step import GraphedStep  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    cfg = synth.CONFIGS[synth.WORKLOADS[wl]["config"]]
    args = synth.workload_args(wl)
    model = build_lily(cfg, args, device="cuda").train()
    batch = synth.make_batch(wl, seed=1)
    r = ops.rt("cuda")
    st = GraphedStep(model, args, batch, use_graph=True)
    state = r.rng_state()
    runs = []
    for i in range(reps):
        r.set_rng_state(state)
        torch.cuda.manual_seed(1234)
        loss = st.run()
        torch.cuda.synchronize()
        runs.append((float(loss), {k: float(v) for k, v in st.losses.items()},
                     {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}))
    l0, ld0, g0 = runs[0]
    gmax = max(float(v.norm()) for v in g0.values())
    for i, (l, ld, g) in enumerate(runs[1:], 1):
        worst, where = 0.0, ""
        for n, v in g.items():
            d = float((v - g0[n]).norm()) / max(float(g0[n].norm()), 1e-6 * gmax)
            if d > worst:
                worst, where = d, n
        print(f"replay {i}: loss {l:.7f} vs {l0:.7f} (rel {abs(l - l0) / abs(l0):.2e}); per-task "
              + ", ".join(f"{k} {abs(ld[k] - ld0[k]) / max(abs(ld0[k]), 1e-30):.1e}" for k in ld0)
              + f"; worst gradient {worst:.2e} ({where})")


if __name__ == "__main__":
    main()
