"""Developer tool: kernel timeline of graph replays of the training step via torch.profiler (CUPTI)."""
import os, sys, json, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from yvb200 import synth, lib, ops
from yvb200.lily_compat import build_lily
from yvb200.step import GraphedStep
wl = "cfg2"
cfg = synth.CONFIGS[synth.WORKLOADS[wl]["config"]]
args = synth.workload_args(wl)
model = build_lily(cfg, args, device="cuda").train()
batch = synth.make_batch(wl, seed=1)
st = GraphedStep(model, args, batch, use_graph=True)
for _ in range(3): st.run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    st.run()
    torch.cuda.synchronize()
out = os.path.join(ROOT, "gpurun_out", "trace.json")
prof.export_chrome_trace(out)
ev = json.load(open(out))["traceEvents"]
ks = [e for e in ev if e.get("cat") == "kernel"]
ks.sort(key=lambda e: e["ts"])
print("kernels", len(ks))
t0 = ks[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in ks)
print(f"span {(t1 - t0)/1e3:.3f} ms, sum of kernel durations {sum(e['dur'] for e in ks)/1e3:.3f} ms")
# busy time (union of intervals)
cur_s, cur_e, busy = None, None, 0.0
for e in ks:
    s, f = e["ts"], e["ts"] + e["dur"]
    if cur_e is None or s > cur_e:
        if cur_e is not None: busy += cur_e - cur_s
        cur_s, cur_e = s, f
    else:
        cur_e = max(cur_e, f)
busy += cur_e - cur_s
print(f"GPU busy (union) {busy/1e3:.3f} ms, idle gaps {(t1 - t0 - busy)/1e3:.3f} ms")
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ks:
    import re
    m = re.search(r"(yv_gemm_pair_kernel<\d>|yv_gemm_kernel<\d>|[a-z_0-9]+_kernel(<\d+>)?)", e["name"])
    n = m.group(1) if m else e["name"][:60]
    agg[n][0] += 1; agg[n][1] += e["dur"]
for n, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{d/1e3:8.3f} ms {c:5d}x {d/c:7.1f} us  {n}")
# concurrency histogram: time-weighted number of kernels in flight
pts = []
for e in ks:
    pts.append((e["ts"], 1)); pts.append((e["ts"] + e["dur"], -1))
pts.sort()
lvl, last, hist = 0, pts[0][0], collections.Counter()
for t, d in pts:
    hist[lvl] += t - last; last = t; lvl += d
print("time by #kernels in flight:", {k: round(v/1e3, 3) for k, v in sorted(hist.items())})

# per-kernel-family: time during which it is the ONLY kernel running (serial exposure)
ivs = sorted((e["ts"], e["ts"] + e["dur"], e["name"]) for e in ks)
import re, heapq
events = []
for s0, f0, nm in ivs:
    events.append((s0, 1, nm)); events.append((f0, -1, nm))
events.sort(key=lambda x: (x[0], x[1]))
active = collections.Counter(); last = events[0][0]; alone = collections.Counter()
def fam(nm):
    m = re.search(r"(yv_gemm_pair_kernel|yv_gemm_kernel|[a-z_0-9]+_kernel)", nm)
    return m.group(1) if m else nm[:40]
for t, d, nm in events:
    tot_active = sum(active.values())
    if tot_active == 1:
        only = [k for k, v in active.items() if v > 0][0]
        alone[only] += t - last
    last = t
    active[fam(nm)] += d
print("time running alone (ms):", {k: round(v / 1e3, 3) for k, v in alone.most_common(12)})

# per-stream view: kernels, busy time, and the gaps between consecutive kernels of the busiest streams
by_stream = collections.defaultdict(list)
for e in ks:
    by_stream[e["args"].get("stream", -1)].append(e)
print("streams:")
for sid, evs in sorted(by_stream.items(), key=lambda kv: -sum(e["dur"] for e in kv[1]))[:8]:
    evs.sort(key=lambda e: e["ts"])
    busy_s = sum(e["dur"] for e in evs)
    gaps = [evs[i + 1]["ts"] - (evs[i]["ts"] + evs[i]["dur"]) for i in range(len(evs) - 1)]
    small = [g for g in gaps if 0 <= g < 20]
    print(f"  stream {sid}: {len(evs):4d} kernels, busy {busy_s/1e3:7.3f} ms, first {(evs[0]['ts']-t0)/1e3:7.3f} last {(evs[-1]['ts']+evs[-1]['dur']-t0)/1e3:7.3f} ms, "
          f"median gap {sorted(gaps)[len(gaps)//2] if gaps else 0:.1f} us, sum of gaps<20us {sum(small)/1e3:.3f} ms")
