"""Developer tool: every kernel of one graph replay of the cfg2 step inside a time window, in start order, with stream,
start, duration and the gap to the previous kernel on the same stream -- to read the dependency chain off the timeline.
    python tools/timeline_window.py <t0_ms> <t1_ms>"""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from yvb200 import synth
from yvb200.lily_compat import build_lily
from yvb200.step import GraphedStep
wins = [(float(sys.argv[i]), float(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)] or [(5.0, 5.6)]
wl = "cfg2"
cfg = synth.CONFIGS[synth.WORKLOADS[wl]["config"]]
args = synth.workload_args(wl)
model = build_lily(cfg, args, device="cuda").train()
st = GraphedStep(model, args, synth.make_batch(wl, seed=1), use_graph=True)
for _ in range(3):
    st.run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    st.run()
    torch.cuda.synchronize()
out = os.path.join(ROOT, "gpurun_out", "trace_w.json")
prof.export_chrome_trace(out)
ev = json.load(open(out))["traceEvents"]
os.remove(out)
ks = sorted((e for e in ev if e.get("cat") == "kernel"), key=lambda e: e["ts"])
t0 = ks[0]["ts"]
last_end = {}
streams = {}
print(f"span {(max(e['ts'] + e['dur'] for e in ks) - t0) / 1e3:.3f} ms, {len(ks)} kernels")
for e in ks:
    sid = e["args"].get("stream", -1)
    sname = streams.setdefault(sid, f"s{len(streams)}")
    s, d = (e["ts"] - t0) / 1e3, e["dur"]
    gap = (e["ts"] - last_end[sid]) if sid in last_end else float("nan")
    last_end[sid] = e["ts"] + e["dur"]
    if any(a <= s <= b for a, b in wins):
        m = re.search(r"(yv_[a-z_]+kernel|[a-z_0-9]+_kernel)", e["name"])
        grid = e["args"].get("grid", "")
        print(f"{s:8.3f} ms  {sname:>4s}  dur {d:6.1f} us  gap {gap:7.1f} us  grid {str(grid):14s} {m.group(1) if m else e['name'][:40]}")
