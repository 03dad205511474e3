"""Developer tool: communication vs compute timeline of the data-parallel step (run under torchrun, one rank per GPU).
Rank 0 profiles one step (torch.profiler / CUPTI) and prints, per gradient-exchange segment, when its NCCL kernel started
and ended relative to the step, the union of compute-kernel time, and how long NCCL kernels ran with no compute kernel in
flight (= exposed communication).
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/exchange_timeline.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from yvb200 import synth  # noqa: E402
from yvb200.lily_compat import build_lily  # noqa: E402
from yvb200.step import GradientExchange, GraphedStep  # noqa: E402


def union(ivs):
    ivs = sorted(ivs)
    out, cs, ce = [], None, None
    for s, e in ivs:
        if ce is None or s > ce:
            if ce is not None:
                out.append((cs, ce))
            cs, ce = s, e
        else:
            ce = max(ce, e)
    if ce is not None:
        out.append((cs, ce))
    return out


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    wl = "cfg2"
    cfg = synth.CONFIGS[synth.WORKLOADS[wl]["config"]]
    args = synth.workload_args(wl)
    model = build_lily(cfg, args, device=dev).train()
    batch = synth.make_batch(wl, seed=1, rank=rank)
    warm = torch.ones(1, device=dev)
    dist.all_reduce(warm)
    torch.cuda.synchronize()
    spec = (sys.argv[1] if len(sys.argv) > 1 else "nccl:320").split(":")      # transport : largest MB [: smallest MB]
    os.environ["YVB200_EXCHANGE"] = spec[0]
    ex = GradientExchange(model, segment_mb=float(spec[1]) if len(spec) > 1 else None)
    if len(spec) > 2:
        ex.segment_min_bytes = int(float(spec[2]) * 2 ** 20)
    st = GraphedStep(model, args, batch, use_graph=True, exchange=ex)
    for _ in range(4):
        st.run()
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            st.run()
            torch.cuda.synchronize()
        out = os.path.join(ROOT, "gpurun_out", f"exchange_trace_{world}gpu.json")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        prof.export_chrome_trace(out)
        ev = json.load(open(out))["traceEvents"]
        ks = sorted((e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy")), key=lambda e: e["ts"])
        os.remove(out)
        t0 = ks[0]["ts"]
        # communication = NCCL kernels, or (copy-engine transport) everything on the streams that carry peer copies
        comm_streams = {e["args"].get("stream") for e in ks if e.get("cat") == "gpu_memcpy" and "PtoP" in e["name"].replace(" ", "")}
        def is_comm(e):
            return "nccl" in e["name"].lower() or (ex.transport == "ce" and e["args"].get("stream") in comm_streams)
        nccl = [e for e in ks if is_comm(e)]
        comp = [e for e in ks if not is_comm(e)]
        sizes = "/".join(f"{sum(g.numel() for g in grads) * 4 / 2 ** 20:.0f}" for _, grads in ex.segments)
        print(f"transport {ex.transport}, segments (MB) {sizes}")
        end_all = max(e["ts"] + e["dur"] for e in ks)
        end_comp = max(e["ts"] + e["dur"] for e in comp)
        print(f"{world} GPUs: step span {(end_all - t0) / 1e3:.3f} ms, compute ends at {(end_comp - t0) / 1e3:.3f} ms, "
              f"{len(ex.segments)} segments of <= {ex.segment_bytes / 2**20:.0f} MB")
        cu = union([(e["ts"], e["ts"] + e["dur"]) for e in comp])
        nu = union([(e["ts"], e["ts"] + e["dur"]) for e in nccl])
        busy_c = sum(e - s for s, e in cu)
        busy_n = sum(e - s for s, e in nu)
        # exposed = nccl time not covered by any compute interval
        exposed = 0.0
        for s, e in nu:
            cov = 0.0
            for cs, ce in cu:
                lo, hi = max(s, cs), min(e, ce)
                if hi > lo:
                    cov += hi - lo
            exposed += (e - s) - cov
        print(f"  compute busy (union) {busy_c / 1e3:.3f} ms, NCCL busy (union) {busy_n / 1e3:.3f} ms, "
              f"NCCL with no compute kernel in flight {exposed / 1e3:.3f} ms")
        for i, e in enumerate(nccl):
            if e["dur"] >= 20 or "nccl" in e["name"].lower():
                print(f"  comm[{i}] {e['name'][:48]:48s} start {(e['ts'] - t0) / 1e3:7.3f} ms  dur {e['dur'] / 1e3:7.3f} ms")
        # when did the gradients become final?  the gather copies of each segment start right after its events fire
        gath = [e for e in ks if "multi_tensor_apply" in e["name"] or "foreach" in e["name"].lower()]
        for e in gath[:40]:
            print(f"  gather {e['name'][:40]:40s} start {(e['ts'] - t0) / 1e3:7.3f} ms  dur {e['dur'] / 1e3:7.3f} ms")
    else:
        st.run()
        torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
