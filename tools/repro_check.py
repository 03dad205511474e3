import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
import torch
from yvb200 import synth
from yvb200.lily_compat import build_lily
wl = "cfg2"
cfg = synth.CONFIGS[synth.WORKLOADS[wl]["config"]]
args = synth.workload_args(wl)
model = build_lily(cfg, args, device="cuda").eval()
batch = [t.cuda() for t in synth.make_batch(wl, seed=4)]
inp = list(synth.model_inputs(batch))
with torch.no_grad():
    outs = [model(*inp) for _ in range(4)]
for k in outs[0]:
    errs = [float((outs[i][k] - outs[0][k]).norm() / outs[0][k].norm()) for i in range(1, 4)]
    print(k, errs, "bitwise" if all(torch.equal(outs[i][k], outs[0][k]) for i in range(1, 4)) else "")
