"""Condense the ncu artefacts of one GPU visit (gpurun_out/<tag>_*) into the small tracked files under profiles/.
usage: python tools/ncu_summarise.py <tag>"""
import collections, csv, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out"); P = os.environ.get("YV_PROFILE_OUT", os.path.join(ROOT, "profiles"))

# 1. launch list -> per-kernel share of the (eager, serialised, cold-cache) step
rows = list(csv.reader(open(os.path.join(G, f"{tag}_launches_eager.csv"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, gi, mi, ui = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= mi:
        continue
    v = float(r[mi].replace(",", "")) / (1000.0 if r[ui] == "ns" else 1.0)
    m = re.search(r"(yv_gemm_pair_kernel<\d>|yv_gemm_kernel<\d>|[A-Za-z_0-9]+_kernel(<\d+>)?)", r[ki])
    name = m.group(1) if m else r[ki][:60]
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, f"{tag}_launch_shares.csv"), "w") as fh:
    fh.write("kernel,launches,total_us,share_pct,avg_us\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write(f"\"{k}\",{c},{t:.1f},{100 * t / tot:.2f},{t / c:.2f}\n")
print(f"launch list: {len(data)} launches, {tot / 1e3:.2f} ms")

# 2. ncu --set full of the GEMM and of the fused attention kernels -> key metrics per captured launch
want = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x", "sm__cycles_elapsed.max",
        "sm__cycles_active.avg"]
for what in ("gemm", "attn"):
    rep = os.path.join(G, f"{tag}_{what}_full.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(os.path.join(P, f"{tag}_{what}_ncu_full_key_metrics.csv"), "w") as fh:
        w = csv.writer(fh)
        w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
        for r in data:
            w.writerow([r[i] for i in idx])
    print(f"{what} full: {len(data)} launches")
    if what == "gemm" and data:
        # DRAM traffic of the dominant launch (first capture: the 2304 x 3072 x 1024 vision projection) for bench.py
        import json
        r0 = data[0]
        def val(name, scale_of):
            i = hdr.index(name)
            u = units[i]
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            return float(r0[i].replace(",", "")) * mult
        tr = val("dram__bytes_read.sum", None) + val("dram__bytes_write.sum", None)
        json.dump({"shape": [2304, 3072, 1024, 3], "dram_bytes": tr,
                   "source": f"ncu --set full, launch 0 of tools/ncu_gemm.py, dram__bytes_read.sum + dram__bytes_write.sum ({tag})"},
                  open(os.path.join(P, "r2_dominant_launch_traffic.json"), "w"))
