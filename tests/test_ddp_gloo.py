"""CPU, world_size 2 over gloo: the N>1 path.  (a) GradientExchange (segmented overlapped all-reduce used by
bench.py / GraphedStep) averages exactly like a plain all-reduce; (b) the drop-in under the reference's own
wrapper DistributedDataParallel(find_unused_parameters=True) (utils/distributed.py:97-99) reproduces the
single-process gradient of the concatenated batch."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from yvb200 import synth, losses
    from yvb200.lily_compat import build_lily
    from yvb200.step import GradientExchange
    cfg = synth.CONFIGS["micro"]
    args = synth.workload_args("micro")

    def grads_of(model, batch):
        model.zero_grad()
        out = model(*synth.model_inputs(batch))
        losses.total_loss(losses.step_losses(batch, out, args, True), args).backward()
        return {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}

    batches = [synth.make_batch("micro", seed=11, rank=r) for r in range(world)]
    # single-process reference: mean of the per-rank gradients (each rank's loss is a mean over its own items)
    ref_model = build_lily(cfg, args).eval()
    per_rank = [grads_of(ref_model, b) for b in batches]
    want = {n: sum(g[n] for g in per_rank) / world for n in per_rank[0]}

    # (a) bucketed exchange, tiny buckets so several collectives are issued
    m1 = build_lily(cfg, args).eval()
    ex = GradientExchange(m1, segment_mb=0.05)
    m1.zero_grad()
    out = m1(*synth.model_inputs(batches[rank]))
    ex.begin()
    losses.total_loss(losses.step_losses(batches[rank], out, args, True), args).backward()
    ex.end()
    ex.exchange()
    err_a = max(float((p.grad - want[n]).abs().max() / (want[n].abs().max() + 1e-12))
                for n, p in m1.named_parameters() if p.grad is not None)
    n_buckets = ex.launched
    # second pass: the layout frozen after the first pass is followed (fixed views of one flat buffer per gradient)
    planned = ex.plan is not None
    m1.zero_grad()
    out = m1(*synth.model_inputs(batches[rank]))
    ex.begin()
    losses.total_loss(losses.step_losses(batches[rank], out, args, True), args).backward()
    ex.end()
    planned = planned and ex._pass_plan is ex.plan and len(ex.segments) * 2 == n_buckets * 2
    ex.exchange()
    err_a = max(err_a, max(float((p.grad - want[n]).abs().max() / (want[n].abs().max() + 1e-12))
                           for n, p in m1.named_parameters() if p.grad is not None))
    views = {id(v.untyped_storage()) for vs in ex.plan.views for v in vs} if planned else set()
    planned = planned and len(views) == 1 and all(p.grad.untyped_storage().data_ptr() == ex.plan.bucket.untyped_storage().data_ptr()
                                                  for p in m1.parameters() if p.grad is not None)
    # third pass with a parameter frozen: the arrival order no longer matches, the pass is observed and re-planned
    frozen = next(p for n, p in m1.named_parameters() if n.endswith("v_layer.0.attention.self.query.weight"))
    frozen.requires_grad_(False)
    m1.zero_grad()
    out = m1(*synth.model_inputs(batches[rank]))
    ex.begin()
    losses.total_loss(losses.step_losses(batches[rank], out, args, True), args).backward()
    ex.end()
    replanned = ex._pass_plan is None and ex.plan is not None and all(q is not frozen for q in ex.plan.order)
    ex.exchange()
    err_a = max(err_a, max(float((p.grad - want[n]).abs().max() / (want[n].abs().max() + 1e-12))
                           for n, p in m1.named_parameters() if p.grad is not None))
    if not (planned and replanned):
        err_a = float("inf")
    ex.remove()

    # (b) the reference's wrapper
    m2 = torch.nn.parallel.DistributedDataParallel(build_lily(cfg, args).eval(), find_unused_parameters=True)
    g2 = grads_of(m2, batches[rank])
    err_b = max(float((g2["module." + n] - want[n]).abs().max() / (want[n].abs().max() + 1e-12)) for n in want)
    q.put((rank, err_a, err_b, n_buckets))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_exchange_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    import queue as _queue
    import time as _time
    res, deadline = [], _time.time() + 300
    while len(res) < world and _time.time() < deadline:
        try:
            res.append(q.get(timeout=2))
        except _queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.terminate()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert len(res) == world
    for rank, err_a, err_b, n_buckets in res:
        assert err_a < 1e-5, (rank, err_a)
        assert err_b < 1e-5, (rank, err_b)
        assert n_buckets >= 3
