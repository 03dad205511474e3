"""Developer tool: run the drop-in on CPU (host path) and CUDA (kernels) and print the first modules whose
outputs / gradients diverge.  usage: python tools/debug_compare.py [workload] [mode]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
import torch
from yvb200 import synth, losses, ops
from yvb200.lily_compat import build_lily

wl = sys.argv[1] if len(sys.argv) > 1 else "micro"
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
cfg = synth.CONFIGS[synth.WORKLOADS[wl]["config"]]
args = synth.workload_args(wl)
mc = build_lily(cfg, args).eval()
mg = build_lily(cfg, args, device="cuda").eval()
ops.rt("cuda").set_precision(mode)
batch = synth.make_batch(wl, seed=1)
bg = [t.cuda() for t in batch]


def flat(o):
    if torch.is_tensor(o):
        return [o]
    if isinstance(o, (tuple, list)):
        r = []
        for x in o:
            r += flat(x)
        return r
    if isinstance(o, dict):
        r = []
        for x in o.values():
            r += flat(x)
        return r
    return []


def hook(store):
    def f(name):
        def g(mod, inp, out):
            store.append((name, [t.detach().cpu().double() for t in flat(out) if t.is_floating_point()]))
        return g
    return f

sc, sg = [], []
for n, m in mc.named_modules():
    m.register_forward_hook(hook(sc)(n))
for n, m in mg.named_modules():
    m.register_forward_hook(hook(sg)(n))
oc = mc(*synth.model_inputs(batch))
og = mg(*synth.model_inputs(bg))
bad = 0
dg = dict(sg)
for n1, t1 in sc:
    if n1 not in dg:
        continue
    t2 = dg[n1]
    for i, (a, b) in enumerate(zip(t1, t2)):
        if a.shape != b.shape:
            print("SHAPE", n1, i, a.shape, b.shape)
            continue
        e = float((a - b).norm() / a.norm().clamp_min(1e-30))
        if e > 2e-5 or e != e:
            print(f"{n1:60s} out{i} rel={e:.3e} shape={tuple(a.shape)}")
            bad += 1
    if bad > 25:
        break
print("forward compare done, mismatches:", bad)
lc = losses.step_losses(batch, oc, args, True)
lg = losses.step_losses(bg, og, args, True)
print({k: (float(lc[k]), float(lg[k])) for k in lc})
losses.total_loss(lc, args).backward()
losses.total_loss(lg, args).backward()
gmax = max(float(p.grad.norm()) for p in mc.parameters() if p.grad is not None)
rows = []
for (n, pc), (_, pg) in zip(mc.named_parameters(), mg.named_parameters()):
    if pc.grad is None or pg.grad is None:
        if (pc.grad is None) != (pg.grad is None):
            print("GRAD PRESENCE", n, pc.grad is None, pg.grad is None)
        continue
    a, b = pc.grad.double(), pg.grad.cpu().double()
    if float(a.norm()) < 1e-6 * gmax:
        continue
    rows.append((float((a - b).norm() / a.norm()), n))
rows.sort(reverse=True)
for e, n in rows[:25]:
    print(f"grad {n:70s} rel={e:.3e}")
