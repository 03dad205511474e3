"""Index arithmetic of the copy-engine gradient exchange (yvb200/step.py, ``YVB200_EXCHANGE=ce``): the three data phases
driven for several fake ranks in one process, phase by phase (the CUDA path puts a barrier between the phases), against
the plain mean.  The real peer-to-peer path is exercised by ``bench.py --gpus N`` (``exchange_mean_error``)."""
import pytest
import torch

from yvb200.step import GradientExchange as GE


@pytest.mark.parametrize("world", [2, 3, 8])
def test_ce_phases_average_every_slice(world):
    torch.manual_seed(world)
    quantum = world * 32
    lens = [quantum * 3, quantum, quantum * 5]                 # three segments
    total = sum(lens)
    buckets = [torch.randn(total) for _ in range(world)]
    want = torch.stack(buckets).double().mean(0)

    def get_buffer_of(_rank):
        def get_buffer(peer, sizes, dtype, offset):
            assert dtype == torch.float32
            return buckets[peer][offset:offset + sizes[0]]
        return get_buffer

    start = 0
    for ln in lens:
        n = ln // world
        stages = [torch.empty(world - 1, max(lens) // world)[:, :n] for _ in range(world)]
        slices = [b[start:start + ln] for b in buckets]
        for r in range(world):
            GE.ce_pull_chunks(r, world, get_buffer_of(r), start, n, stages[r])
        for r in range(world):
            GE.ce_reduce(r, world, slices[r], n, stages[r])
        for r in range(world):
            GE.ce_pull_reduced(r, world, get_buffer_of(r), start, n, slices[r])
        start += ln
    for r in range(world):
        assert torch.equal(buckets[r], buckets[0])             # bit-identical on every rank
    assert float((buckets[0].double() - want).abs().max()) < 1e-6


def _one_rank_group():
    import socket
    import torch.distributed as dist
    if dist.is_initialized():
        return False
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    return True


def test_plan_is_frozen_after_the_first_pass_and_followed():
    """One rank (gloo, CPU): the first pass is observed, the plan reproduces its segmentation, later passes land in
    fixed views of one flat buffer and leave the (one-rank) gradients unchanged; shrinking segments keep every
    parameter exactly once."""
    import torch.distributed as dist
    mine = _one_rank_group()
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(*[torch.nn.Linear(64, 64) for _ in range(12)])
        x = torch.randn(8, 64)
        ex = GE(net, segment_mb=64 * 64 * 4 * 2.5 / 2 ** 20)          # a segment closes every third weight or so

        def one_pass():
            net.zero_grad()
            ex.begin()
            net(x).square().mean().backward()
            ex.end()
            local = {n: p.grad.clone() for n, p in net.named_parameters()}
            ex.exchange()
            return local

        one_pass()
        assert ex._pass_plan is None and ex.plan is not None           # observed; plan frozen at its end
        observed = [len(g) for _, g in ex.segments]
        assert [len(s) for s in ex.plan.segments] == observed and len(observed) >= 4
        local = one_pass()
        assert ex._pass_plan is ex.plan
        base = ex.plan.bucket.untyped_storage().data_ptr()
        for n, p in net.named_parameters():
            assert p.grad.untyped_storage().data_ptr() == base
            assert torch.equal(p.grad, local[n])
        # shrinking rule: target = half of what is still to come, never below the minimum, never above the maximum
        ex.segment_min_bytes = 64 * 4
        segs = ex._planned_segments(ex.plan.order)
        sizes = [sum(q.numel() * 4 for q in s) for s in segs]
        assert sum(len(s) for s in segs) == len(ex.plan.order) and len(segs) > len(observed)
        assert sizes[0] >= sizes[len(sizes) // 2] and max(sizes) <= ex.segment_bytes + 64 * 64 * 4
        ex.remove()
    finally:
        if mine:
            dist.destroy_process_group()
