#!/usr/bin/env python
"""bench.py -- trajectory-instruction pairs/sec of the ViLBERT training step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (B200 kernels, one process per GPU)
  python bench.py --impl reference --gpus N ...            reference arm: the CPU oracle port on host cores

A step = forward + the active losses + backward on one batch (train mode, dropout on, weights re-split to bf16 planes
every step).  `value` times K CUDA-graph replays with inputs resident in HBM (CUDA events, L2 flushed between steps,
max over ranks); `e2e` adds the pinned-host -> device copy of the batch and the device -> host read of the loss inside
the timed region.

Workloads (BASELINE.json `configs`; default cfg2 = the configuration the metric is quoted on):
  cfg2            full pre-training step (vision + language + ranking + traj), 8 frames, 8 pairs per GPU   [weak scaling]
  cfg3            ranking-only fine-tune step, GLOBAL batch 16 pairs split over the ranks                  [strong]
  cfg4_p4/8/16/32 cfg2 with 4 / 8 / 16 / 32 frames (trajectory-length sweep), 8 pairs per GPU              [weak]
  cfg5            cfg2 objectives, GLOBAL batch 64 pairs split over the ranks                              [strong]
  --pairs-per-gpu / --global-batch override the batch of any of them.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))

import torch  # noqa: E402

METRIC = "trajectory-instruction pairs/sec"
UNIT = "pairs/s"
#: workload name -> (synth base workload, global batch or None for 8 pairs per GPU)
WORKLOAD_TABLE = {"cfg2": ("cfg2", None), "cfg3": ("cfg3_rank", 16), "cfg4_p4": ("cfg4_p4", None), "cfg4_p8": ("cfg2", None),
                  "cfg4_p16": ("cfg4_p16", None), "cfg4_p32": ("cfg4_p32", None), "cfg5": ("cfg2", 64)}


def mac_fwd_per_pair(cfg, frames, boxes, tokens, args):
    """multiply-accumulates of one forward pass per pair (bias / LN / softmax / GELU ignored).
    37.322 GMAC at cfg2; train FLOPs = 6 x this (forward 2, dgrad 2, wgrad 2)."""
    V, T = frames * boxes, tokens
    Ht, Ft, Lt = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"]
    Hv, Fv, Lv = cfg["v_hidden_size"], cfg["v_intermediate_size"], cfg["v_num_hidden_layers"]
    Hb, Lc = cfg["bi_hidden_size"], len(cfg["v_biattention_id"])
    mac = V * (cfg["v_feature_size"] + 11) * Hv
    mac += Lt * (T * (4 * Ht * Ht + 2 * Ht * Ft) + 2 * T * T * Ht)
    mac += Lv * (V * (4 * Hv * Hv + 2 * Hv * Fv) + 2 * V * V * Hv)
    mac += Lc * (V * 3 * Hv * Hb + T * 3 * Ht * Hb + 4 * T * V * Hb + V * Hb * Hv + T * Hb * Ht + V * 2 * Hv * Fv
                 + T * 2 * Ht * Ft)
    mac += (Ht + Hv) * Hb + 2 * Hb
    if args.masked_language:
        mac += T * (Ht * Ht + Ht * cfg["vocab_size"])
    if args.masked_vision:
        mac += V * (Hv * Hv + Hv * cfg["v_target_size"])
    return float(mac)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return (float(p.get("bf16_tflops_sustained", 1382.5)), float(p.get("bf16_tflops", 1590.0)),
                float(p.get("hbm_gbs", 6538.3)), "measured")
    return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        busy = [x for x in sm if mx and x > 0.3 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def resolve_workload(a, world):
    """-> (synth workload name, pairs per GPU, scaling, description fields)"""
    from yvb200 import synth
    base, global_batch = WORKLOAD_TABLE[a.workload]
    if a.global_batch:
        global_batch = a.global_batch
    if a.pairs_per_gpu:
        pairs, scaling = a.pairs_per_gpu, "weak"
    elif global_batch:
        if global_batch % world:
            raise SystemExit(f"global batch {global_batch} does not split over {world} ranks")
        pairs, scaling = global_batch // world, "strong"
    else:
        pairs, scaling = synth.WORKLOADS[base]["bs"] * synth.WORKLOADS[base]["cands"], "weak"
    if pairs < 2:
        raise SystemExit("at least 2 pairs (one item with 2 candidates) per GPU")
    return synth.derive_workload(base, pairs), pairs, scaling, global_batch


def run_cpu_oracle(wl_name, steps, warmup, threads=None):
    """fwd + losses + bwd of the oracle port on the host cores; returns (pairs/s, seconds per step, cores)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vilbert_oracle as O
    from yvb200 import synth
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = synth.CONFIGS[synth.WORKLOADS[wl_name]["config"]]
    args = synth.workload_args(wl_name)
    sd = synth.lily_state_dict(cfg, seed=0)
    batch = synth.make_batch(wl_name, seed=1)
    n = synth.num_pairs(batch)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.oracle_step(sd, cfg, args, batch, dtype=torch.float32, want_grads=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return n / sec, sec, cores


def timed_events(fn, reps, flush=None):
    evs = []
    for _ in range(reps):
        if flush is not None:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    return [x.elapsed_time(y) for x, y in evs]


def contraction_roofline(model, args, host_batch, dev, flush, peak_tf, peak_src, precision, reps):
    """Roofline of the dominant kernels.  The step is one CUDA graph, so single launches cannot be bracketed inside the
    timed region; instead the same step is captured a second time with every launch that is NOT a tensor-core contraction
    (yv_gemm, yv_attn_fwd, yv_attn_bwd) suppressed and all work on ONE stream: that graph holds exactly the step's
    contraction launches, back to back in program order.  Its replay time (CUDA events, L2 flushed between replays like
    the timed loop) is their summed duration; achieved = sum of their algorithmic FLOPs / that time."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import devtools
    from yvb200 import ops
    from yvb200.step import GraphedStep
    rt_ = ops.rt(dev)
    was_concurrent = rt_.concurrent
    rt_.concurrent = False
    try:
        with devtools.contraction_only() as trace:
            gonly = GraphedStep(model, args, host_batch, use_graph=True, warmup=1)
            launches = list(trace[-gonly.launches_per_step:]) if gonly.launches_per_step else list(trace)
    finally:
        rt_.concurrent = was_concurrent
    for _ in range(3):
        gonly.graph.replay()
    torch.cuda.synchronize(dev)
    g_ms = statistics.mean(timed_events(gonly.graph.replay, reps, flush))
    g_flop = sum(fl for _, _, fl, _, _ in launches)
    n_gemm = sum(1 for n, *_ in launches if n == "yv_gemm")
    n_attn = len(launches) - n_gemm
    attn_flop = sum(fl for n, _, fl, _, _ in launches if n != "yv_gemm")
    del gonly
    hw_mult = 3.0 if precision == "bf16x3" else 1.0
    achieved = g_flop / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    return {"bound": "tensor", "kernel": "yv_gemm_kernel + yv_attn_fwd/bwd_kernel (every tensor-core contraction launch of one "
                                         "step, replayed back to back as a contraction-only CUDA graph on one stream)",
            "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": None,
            "peak_source": f"{peak_src} bf16_tflops_sustained", "launches": len(launches), "gemm_launches": n_gemm,
            "attention_launches": n_attn, "attention_share_of_flops": attn_flop / g_flop if g_flop else 0.0,
            "ms_per_step_in_kernel": g_ms, "avg_launch_us": g_ms * 1e3 / max(1, len(launches)),
            "algorithmic_gflop_per_step": g_flop / 1e9, "tensor_pipe_flop_multiplier": hw_mult,
            "frac_of_tensor_pipe": achieved * hw_mult / peak_tf}


def dominant_launch(dev, pairs, wl, cfg, precision, burst_tf):
    """The single most expensive launch of the step on its own: the Q|K|V projection of the vision stream
    (2304 x 3072 x 1024 at 8 pairs x 8 frames), 50 launches back to back.  `traffic` = DRAM read + write of that launch
    from the committed ncu --set full capture of this round (profiles/r2_dominant_launch_traffic.json), cited, not
    re-measured: ncu cannot run inside the bench."""
    from yvb200 import lib
    Md, Nd, Kd = pairs * wl["frames"] * wl["boxes"], 3 * cfg["bi_hidden_size"], cfg["v_hidden_size"]
    pa = lib.split_planes(torch.randn(Md, Kd, device=dev))
    pb = lib.split_planes(torch.randn(Nd, Kd, device=dev) * 0.05)
    bias_d = torch.randn(Nd, device=dev)
    outp = lib.Planes.empty(Md, Nd, dev)
    passes = 3 if precision == "bf16x3" else 1

    def one():
        lib.gemm(Md, Nd, Kd, lib.op_of(pa), lib.op_of(pb), passes=passes, bias=bias_d, out_planes=outp.ptr(),
                 ld_pl=outp.ld, pl_plane_stride=outp.plane_stride)
    for _ in range(5):
        one()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        one()
    e1.record()
    torch.cuda.synchronize(dev)
    us = e0.elapsed_time(e1) / 50 * 1e3
    tf = 2.0 * Md * Nd * Kd / (us * 1e-6) / 1e12
    traffic, src = None, None
    tpath = os.path.join(ROOT, "profiles", "r2_dominant_launch_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            t = json.load(fh)
        if [Md, Nd, Kd, passes] == t.get("shape"):
            traffic, src = t.get("dram_bytes"), f"profiles/r2_dominant_launch_traffic.json ({t.get('source')})"
    return {"what": f"yv_gemm {Md}x{Nd}x{Kd} (vision Q|K|V projection), {passes} pass(es), plane output, 50 launches back to back",
            "us_per_launch": us, "achieved": tf, "peak": burst_tf, "unit": "TFLOP/s", "frac": tf / burst_tf,
            "frac_of_tensor_pipe": tf * passes / burst_tf, "peak_source": "bf16_tflops (burst: kernel timed alone)",
            "traffic": traffic, "traffic_source": src, "algorithmic_bytes": 2.0 * 2 * (Md * Kd + Nd * Kd) + 2.0 * 2 * Md * Nd}


def torch_ops_proxy(dev, wl_name, pairs):
    """What the reference's un-fused ATen path costs on this GPU.  The reference checkout is not on the GPU box (and is
    not committed), so this is a PROXY: the oracle restatement of vilbert/vilbert.py + get_loss_correct issued as eager
    fp32 torch ops on cuda:0 (TF32 off) -- (a) eval mode, (b) train mode with dropout at the reference's 94 sites (what
    pretrain.py actually runs), (c) the same under torch.autocast(bf16).  20 steps each, median."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vilbert_oracle as O
    from yvb200 import synth
    cfg = synth.CONFIGS[synth.WORKLOADS[wl_name]["config"]]
    args = synth.workload_args(wl_name)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd = {k: v.to(dev).requires_grad_(True) for k, v in synth.lily_state_dict(cfg, seed=0).items()
          if not k.endswith("cls.predictions.decoder.weight")}
    dbatch = [t.to(dev) if torch.is_tensor(t) else t for t in synth.make_batch(wl_name, seed=1)]
    res = {"what": "PROXY for the reference's stock PyTorch-CUDA path: oracle restatement as eager torch ops on cuda:0 "
                   "(fp32, TF32 off); the reference checkout itself is not available on the GPU box", "unit": UNIT}
    legs = (("eval_fp32", False, False, 5, 2), ("train_fp32_dropout", True, False, 20, 5),
            ("train_autocast_bf16_dropout", True, True, 20, 5))
    for name, train, amp, steps, warm in legs:
        def one():
            if amp:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    O.oracle_step(sd, cfg, args, dbatch, clone=False, train_dropout=train)
            else:
                O.oracle_step(sd, cfg, args, dbatch, clone=False, train_dropout=train)
        try:
            for _ in range(warm):
                one()
            torch.cuda.synchronize(dev)
            ms = statistics.median(timed_events(one, steps))
            res[name] = {"ms_per_step": ms, "value": pairs / (ms * 1e-3), "steps": steps}
        except Exception as e:  # informational only
            res[name] = {"error": repr(e)[:200]}
    return res


def golden_parity(model, args, wl_name, host_batch, dev):
    """One eval-mode step through the benched path (GraphedStep, graph on) against the total loss the reference itself
    produced for this workload (tests/golden/<workload>.npz, recorded by oracle/make_golden.py)."""
    import numpy as np
    from yvb200.step import GraphedStep
    path = os.path.join(ROOT, "tests", "golden", f"{wl_name}.npz")
    if not os.path.exists(path):
        return None
    ref = float(np.load(path)["total_loss"])
    was_training = model.training
    model.eval()
    try:
        st = GraphedStep(model, args, host_batch, use_graph=True, warmup=1, prefetch=False)
        got = float(st.run())
        del st
    finally:
        model.train(was_training)
    rel = abs(got - ref) / abs(ref)
    if not rel < 1e-3:
        raise SystemExit(f"bench.py: eval-mode loss {got} differs from the reference's {ref} (rel {rel:.2e}) -- refusing to time a wrong step")
    return {"eval_total_loss": got, "reference_total_loss": ref, "rel_err": rel, "tolerance": 1e-3,
            "golden": f"tests/golden/{wl_name}.npz"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOAD_TABLE))
    ap.add_argument("--pairs-per-gpu", type=int, default=0)
    ap.add_argument("--global-batch", type=int, default=0)
    ap.add_argument("--precision", default=os.environ.get("YVB200_PRECISION", "bf16x3"), choices=["bf16x3", "bf16"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="developer runs: only the timed step loops, no roofline / baselines")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from yvb200 import synth
    wl_name, pairs, scaling, global_batch = resolve_workload(a, world)
    wl = synth.WORKLOADS[wl_name]
    cfg = synth.CONFIGS[wl["config"]]
    args = synth.workload_args(wl_name)
    tasks = "+".join(k for k, on in (("vision", args.masked_vision), ("language", args.masked_language),
                                     ("ranking", args.ranking), ("traj", args.traj_judge)) if on)
    train_gflop_per_pair = 6.0 * mac_fwd_per_pair(cfg, wl["frames"], wl["boxes"], wl["tokens"], args) / 1e9
    config = {"workload": f"{a.workload}: ViLBERT {cfg['num_hidden_layers']}t/{cfg['v_num_hidden_layers']}v/"
                          f"{len(cfg['v_biattention_id'])}c training step ({tasks}), {wl['frames']} frames x {wl['boxes']} "
                          f"regions, {wl['tokens']} tokens, {pairs} pairs/GPU"
                          + (f" (global batch {global_batch})" if scaling == "strong" else ""),
              "pairs_per_gpu": pairs, "global_batch": pairs * world, "parallelism": f"dp{a.gpus}", "precision": a.precision,
              "train_gflop_per_pair": train_gflop_per_pair,
              "gradient_exchange": "none (1 GPU)" if world == 1 else
              "segments closed by external CUDA events inside the step graph; each segment's slice of one flat buffer is "
              "averaged on a communication stream while the rest of backward runs",
              "l2": "flushed between timed steps (256 MB write); per-step working set ~3 GB >> 126 MB L2"}

    if a.impl == "reference":
        if rank != 0:
            return
        k, w = max(1, min(a.steps, 5)), max(1, min(a.warmup, 1))
        v, sec, cores = run_cpu_oracle(wl_name, k, w)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": k, "warmup": w,
                "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{k} full {a.workload} steps ({pairs} pairs each, eval-mode dropout) of the CPU "
                                           "oracle restatement of vilbert/vilbert.py + get_loss_correct; the Python reference "
                                           "itself cannot travel to the GPU box"},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device for --impl ours (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    force_exchange = world == 1 and os.environ.get("YVB200_FORCE_EXCHANGE", "0") == "1"   # debugging aid
    if world > 1 or force_exchange:
        import torch.distributed as dist
        if force_exchange:
            dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29533", rank=0, world_size=1, device_id=dev)
        else:
            dist.init_process_group("nccl", device_id=dev)
    os.environ["YVB200_PRECISION"] = a.precision
    from yvb200 import lib, ops
    from yvb200.lily_compat import build_lily
    from yvb200.step import GraphedStep
    ops.rt(dev).set_precision(a.precision)
    model = build_lily(cfg, args, device=dev).train()
    host_batch = [t.pin_memory() if torch.is_tensor(t) else t for t in synth.make_batch(wl_name, seed=1, rank=rank)]
    parity = None
    if world == 1 and not a.quick:
        parity = golden_parity(model, args, wl_name, host_batch, dev)     # before timing: a wrong step is not timed
    exchange = None
    if world > 1 or force_exchange:
        from yvb200.step import GradientExchange
        warm = torch.ones(1, device=dev)
        dist.all_reduce(warm)                       # communicator set-up happens outside any capture
        torch.cuda.synchronize(dev)
        exchange = GradientExchange(model)          # segmented NCCL all-reduce overlapped with backward
    step = GraphedStep(model, args, host_batch, use_graph=not a.no_graph, exchange=exchange)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(a.warmup):
        step.run()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t_dev = sum(timed_events(step.run, a.steps, flush)) / 1e3
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the public step API: pinned host batch -> device, loss -> host, every step
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()

    def e2e_step():
        step.load(host_batch)
        loss = step.run()
        loss_host.copy_(loss.reshape(1), non_blocking=True)
    barrier()
    t_e2e = sum(timed_events(e2e_step, a.steps, flush)) / 1e3
    barrier()
    final_loss = float(loss_host[0])
    metrics = {k: {t: float(v) for t, v in d.items()} for k, d in step.metrics().items()}   # one packed all-reduce

    t = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
    exchange_check = exchange_mean_error = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # after the exchange every rank must hold the same averaged gradients (ranks see different batches)
        ps = [p for p in model.parameters() if p.grad is not None]
        chk = torch.stack([p.grad.double().abs().sum() for p in ps[:: max(1, len(ps) // 64)]])
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        exchange_check = float(max(((c - allc[0]).abs() / allc[0].clamp_min(1e-30)).max() for c in allc))
        exchange_mean_error = exchange.verify()      # vs all-gather + fp64 mean of the ranks' local gradients
        config["gradient_exchange"] += f" [transport: {exchange.transport}, payload: {exchange.payload}]"
    t_dev, t_e2e = float(t[0]), float(t[1])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()        # the other ranks are done: nothing below runs while they wait
    if rank != 0:
        return
    total_pairs = pairs * world * a.steps
    value = total_pairs / t_dev
    ms_per_step = t_dev / a.steps * 1e3
    if a.quick:
        print(json.dumps({"quick": True, "workload": a.workload, "ms_per_step": ms_per_step,
                          "e2e_ms_per_step": t_e2e / a.steps * 1e3, "value": value,
                          "gpu_launches_per_step": step.launches_per_step, "final_loss": final_loss,
                          "exchange_max_rank_mismatch": exchange_check, "exchange_mean_error": exchange_mean_error,
                          "exchange_transport": getattr(exchange, "transport", None),
                          "variant": os.environ.get("YVB200_GEMM_VARIANT", "auto")}))
        return

    peak_tf, burst_tf, peak_bw, peak_src = peaks()
    hw_mult = 3.0 if a.precision == "bf16x3" else 1.0
    step_tf = pairs * train_gflop_per_pair / ms_per_step          # algorithmic TFLOP/s per GPU over the whole step
    roofline, torch_gpu, optim_ms, cpu = None, None, None, None
    if world == 1:
        # single-GPU extras (rank 0 is the only process left: no rank waits in a collective while these run)
        if exchange is not None:
            exchange.remove()
        try:
            roofline = contraction_roofline(model, args, host_batch, dev, flush, peak_tf, peak_src, a.precision,
                                            max(3, min(a.steps, 10)))
        except Exception as e:
            roofline = {"bound": "tensor", "achieved": step_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": step_tf / peak_tf,
                        "traffic": None, "error": repr(e)[:200]}
        try:
            roofline["dominant_launch"] = dominant_launch(dev, pairs, wl, cfg, a.precision, burst_tf)
        except Exception as e:  # never let the extra evidence break the bench line
            roofline["dominant_launch"] = {"error": repr(e)[:200]}
        try:
            torch_gpu = torch_ops_proxy(dev, wl_name, pairs)
        except Exception as e:  # informational only
            torch_gpu = {"error": repr(e)[:200]}
        # optimizer step, reported separately from the metric: fused multi-tensor AdamW over all gradients
        try:
            from yvb200.optim import FusedAdamW
            nd = ("bias", "LayerNorm.weight", "LayerNorm.bias")
            named = list(model.named_parameters())
            opt = FusedAdamW([{"params": [p for n, p in named if any(x in n for x in nd)], "weight_decay": 0.0},
                              {"params": [p for n, p in named if not any(x in n for x in nd)], "weight_decay": 0.01}], lr=4e-5)
            for _ in range(2):
                opt.step()
            torch.cuda.synchronize(dev)
            optim_ms = statistics.mean(timed_events(opt.step, 5))
        except Exception as e:  # informational only
            optim_ms = repr(e)[:200]
        if not a.no_cpu_baseline:
            v, sec, cores = run_cpu_oracle(wl_name, 2, 1)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"2 full {a.workload} steps ({pairs} pairs each) after 1 warm-up, fwd+losses+bwd of the CPU oracle port"}
    else:
        roofline = {"bound": "tensor", "kernel": "whole step (per-kernel replay only runs at N=1)", "achieved": step_tf,
                    "peak": peak_tf, "unit": "TFLOP/s", "frac": step_tf / peak_tf, "traffic": None,
                    "peak_source": f"{peak_src} bf16_tflops_sustained", "tensor_pipe_flop_multiplier": hw_mult}
    roofline["step_algorithmic_tflops_per_gpu"] = step_tf
    roofline["step_frac"] = step_tf / peak_tf

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": a.precision, "data": "synthetic", "config": config,
            "e2e": {"value": total_pairs / t_e2e, "unit": UNIT, "h2d_bytes_per_step": step.h2d_bytes(),
                    "d2h_bytes_per_step": 4, "ms_per_step": t_e2e / a.steps * 1e3},
            "gpu_launches": step.launches_per_step * a.steps, "gpu_launches_per_step": step.launches_per_step,
            "cuda_graph": not a.no_graph, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "torch_ops_on_gpu_proxy": torch_gpu, "fused_adamw_ms_per_step": optim_ms,
            "exchange_max_rank_mismatch": exchange_check, "exchange_mean_error": exchange_mean_error,
            "train_tflops_algorithmic": value * train_gflop_per_pair / 1e3,
            "final_loss": final_loss, "step_metrics": metrics, "parity_check": parity}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
