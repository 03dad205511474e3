"""Developer tool: per-shape time of every yv_gemm launch in one training step (CUDA events, GPU busy-ahead)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
import torch
from yvb200 import synth, lib, ops
from yvb200.lily_compat import build_lily
from yvb200.step import GraphedStep
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
cfg = synth.CONFIGS[synth.WORKLOADS[wl]["config"]]
args = synth.workload_args(wl)
ops.rt("cuda").set_precision(mode)
model = build_lily(cfg, args, device="cuda").train()
batch = synth.make_batch(wl, seed=1)
st = GraphedStep(model, args, batch, use_graph=False, warmup=2)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import devtools
with devtools.trace_gemms() as tr:
    torch.cuda._sleep(int(6e8))
    st.run()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for name, shape, fl, e0, e1 in tr:
    if name != "yv_gemm":
        continue
    M, N, K, B, P = shape
    k = (M, N, K, B, P)
    c, t = agg.get(k, (0, 0.0))
    agg[k] = (c + 1, t + e0.elapsed_time(e1))
tot = sum(t for _, t in agg.values())
print(f"total gemm ms {tot:.3f} launches {len(tr)}")
for (M, N, K, B, P), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    fl = 2.0 * M * N * K * B * c
    tiles = ((M + 127) // 128) * ((N + 127) // 128) * B
    print(f"M={M:6d} N={N:6d} K={K:6d} B={B:4d} p={P} x{c:3d}  {t:8.3f} ms  {t/c*1e3:8.1f} us/launch  tiles={tiles:5d} "
          f"{fl/t/1e9:8.1f} algTF/s")
