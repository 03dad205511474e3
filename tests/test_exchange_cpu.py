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
