"""Developer tool (run under torchrun, one rank per GPU): step time of the data-parallel cfg2 step for several gradient-
exchange settings in ONE process per rank -- transport (nccl | ce | nvls) x segment size -- with the step split into
"graph done" (compute) and "exchange done" by CUDA events, max over ranks.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29521 tools/exchange_sweep.py \
        nccl:320 ce:320 ce:320:24 nccl:320:320      (transport : largest segment MB [: smallest segment MB])"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from yvb200 import ops, synth  # noqa: E402
from yvb200.lily_compat import build_lily  # noqa: E402
from yvb200.step import GradientExchange, GraphedStep  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    configs = [c for c in sys.argv[1:] if ":" in c] or ["nccl:320", "ce:320", "ce:96"]
    steps = 20
    wl = "cfg2"
    cfg = synth.CONFIGS[synth.WORKLOADS[wl]["config"]]
    args = synth.workload_args(wl)
    model = build_lily(cfg, args, device=dev).train()
    batch = synth.make_batch(wl, seed=1, rank=rank)
    warm = torch.ones(1, device=dev)
    dist.all_reduce(warm)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    state0 = chk0 = None
    for c in configs:
        transport, mb, *rest = c.split(":")
        os.environ["YVB200_EXCHANGE"] = transport
        direct = "direct" in rest               # ":direct" = weight gradients written straight into the flat buffer
        rest = [x for x in rest if x not in ("direct", "gather")]
        ex = GradientExchange(model, segment_mb=float(mb), direct=direct)
        if rest:                                # smallest segment towards the end of backward (default: no shrinking)
            ex.segment_min_bytes = int(float(rest[0]) * 2 ** 20)
        st = GraphedStep(model, args, batch, use_graph=True, exchange=ex)
        main_stream = torch.cuda.current_stream(dev)
        for _ in range(5):
            st.run()
        torch.cuda.synchronize()
        dist.barrier()
        evs = []
        for _ in range(steps):
            flush.fill_(1)
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            st.graph.replay()
            e1.record(main_stream)              # the captured step (compute) is done
            ex.exchange()
            e2.record(main_stream)              # the last segment has been averaged
            evs.append((e0, e1, e2))
        torch.cuda.synchronize()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b, _ in evs) / steps, sum(a.elapsed_time(c_) for a, _, c_ in evs) / steps],
                         dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        err = ex.verify(8)
        # the averaged gradients of one step from a fixed RNG state must not depend on the exchange settings
        r = ops.rt(dev)
        if state0 is None:
            state0 = r.rng_state()
        r.set_rng_state(state0)
        torch.cuda.manual_seed(1234)
        st.run()
        torch.cuda.synchronize()
        ps = [p for p in model.parameters() if p.grad is not None]
        chk = torch.stack([torch.stack([p.grad.double().sum(), p.grad.double().abs().sum()]) for p in ps[:: max(1, len(ps) // 96)]])
        if chk0 is None:
            chk0 = chk
        drift = float(((chk - chk0).abs() / chk0.abs().clamp_min(1e-12))[:, 1].max())
        if rank == 0:
            sizes = "/".join(f"{sum(g.numel() for g in grads) * 4 / 2 ** 20:.0f}" for _, grads in ex.segments)
            print(f"{world} GPUs  {ex.transport:5s} segments <= {float(mb):4.0f} MB, >= {ex.segment_min_bytes / 2 ** 20:3.0f} MB ({sizes}): graph done {float(t[0]):7.3f} ms, "
                  f"exchange done {float(t[1]):7.3f} ms  (tail {float(t[1] - t[0]):6.3f} ms)  mean error {err:.1e}  "
                  f"in place {ex.in_place_fraction():.2f}  gradient checksums vs first config {drift:.1e}", flush=True)
        ex.remove()
        del st, ex
        torch.cuda.synchronize()
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
