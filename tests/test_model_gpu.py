"""GPU parity of the whole path (drop-in Lily on CUDA -> C ABI kernels) against
  (a) golden vectors recorded from the real reference (tests/golden/*.npz), and
  (b) the CPU oracle on the same seeded inputs.
Tolerance: 1e-3 relative (BASELINE.json north_star) on the four outputs, every loss and every parameter
gradient, in the default bf16x3 mode.  Gradient tensors whose reference norm is < 1e-6 of the largest one
(analytically zero key biases) are only required to stay that small."""
import numpy as np
import pytest
import torch

from yvb200 import synth, losses
from yvb200.lily_compat import build_lily
from test_oracle_golden import _check_grads, _check_outputs, _load

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _to_dev(batch, dev):
    return [t.to(dev) if torch.is_tensor(t) else t for t in batch]


def _run(wl, mode=None, train=False):
    from yvb200 import ops
    cfg = synth.CONFIGS[synth.WORKLOADS[wl]["config"]]
    args = synth.workload_args(wl)
    model = build_lily(cfg, args, device="cuda")
    model.train(train)
    if mode:
        ops.rt("cuda").set_precision(mode)
    batch = _to_dev(synth.make_batch(wl, seed=1), "cuda")
    out = model(*synth.model_inputs(batch))
    ld = losses.step_losses(batch, out, args, training=True)
    tot = losses.total_loss(ld, args)
    tot.backward()
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().cpu() for n, p in model.named_parameters() if p.grad is not None}
    return ({k: v.detach().cpu() for k, v in out.items()}, {k: float(v) for k, v in ld.items()}, float(tot), grads,
            model)


@pytest.mark.parametrize("wl", ["micro", "micro_pad", "cfg1", "cfg2", "cfg4_p16", "cfg4_p32"])
def test_cuda_path_matches_reference_golden(golden_dir, wl):
    """(cfg4_p16 / cfg4_p32: the trajectory-length sweep of BASELINE config 4 at its full batch of 8 pairs -- 576 and 1152
    vision tokens per pair, the KV-chunked regime of the fused attention kernels.)"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from yvb200 import lib
    n0 = lib.launch_count()
    g = _load(golden_dir, wl)
    out, ld, tot, grads, model = _run(wl, "bf16x3")
    assert lib.launch_count() - n0 > 50, "the CUDA path must run yvb200 kernels"
    for k, v in ld.items():
        ref = float(g[f"loss/{k}"])
        assert abs(v - ref) <= TOL * max(1.0, abs(ref)), (k, v, ref)
    assert abs(tot - float(g["total_loss"])) <= TOL * abs(float(g["total_loss"]))
    _check_outputs(g, out, TOL)
    n = _check_grads(g, grads, TOL)
    assert n > 100
    dead = {k[len("nograd/"):] for k in g.files if k.startswith("nograd/")}
    assert dead == {n_ for n_, p in model.named_parameters() if p.grad is None}


def test_cuda_path_matches_oracle_seeded(golden_dir):
    """Same comparison against the oracle run here on the host (different seed than the golden files)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import vilbert_oracle as O
    from yvb200 import ops
    wl = "cfg1"
    cfg = synth.CONFIGS[synth.WORKLOADS[wl]["config"]]
    args = synth.workload_args(wl)
    sd = synth.lily_state_dict(cfg, seed=0)
    batch = synth.make_batch(wl, seed=5)
    o_out, o_ld, o_tot, o_grads = O.oracle_step(sd, cfg, args, batch, dtype=torch.float32)
    model = build_lily(cfg, args, device="cuda").eval()
    ops.rt("cuda").set_precision("bf16x3")
    b = _to_dev(batch, "cuda")
    out = model(*synth.model_inputs(b))
    ld = losses.step_losses(b, out, args, training=True)
    losses.total_loss(ld, args).backward()
    for k in o_out:
        a, r = out[k].detach().cpu().double(), o_out[k].double()
        assert float((a - r).norm() / r.norm()) < TOL, k
    for k in o_ld:
        assert abs(float(ld[k]) - float(o_ld[k])) < TOL * max(1.0, abs(float(o_ld[k]))), k
    gmax = max(float(v.norm()) for v in o_grads.values())
    for n, p in model.named_parameters():
        if p.grad is None:
            continue
        r = o_grads[n].double()
        if float(r.norm()) < 1e-6 * gmax:
            continue
        assert float((p.grad.cpu().double() - r).norm() / r.norm()) < TOL, n


def test_plain_bf16_mode_documented_looser_bound(golden_dir):
    """YVB200_PRECISION=bf16 (single pass): SURVEY measured 3e-3..8e-3 end to end on the wide outputs; the scalar
    ranking / traj logits are cancellation-heavy sums and move by ~3e-2 with the split-K partition, so the gate of
    this opt-in mode is 6e-2 (the parity mode, bf16x3, is gated at 1e-3 in the tests above)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    g = _load(golden_dir, "cfg1")
    out, ld, tot, grads, _ = _run("cfg1", "bf16")
    from yvb200 import ops
    ops.rt("cuda").set_precision("bf16x3")
    _check_outputs(g, out, 6e-2)
    assert abs(tot - float(g["total_loss"])) <= 3e-2 * abs(float(g["total_loss"]))


def test_train_mode_dropout_statistics():
    """Train-mode parity with the reference's Philox stream is impossible; check instead that dropout is
    active, reproducible for a fixed RNG state and re-drawn every forward.  (Keep rate 0.9 +- 0.01 and the exact 1/0.9
    scaling of the kept elements -- i.e. unbiasedness -- are asserted at kernel level:
    test_kernels_gpu.test_gemm_dropout_is_deterministic_and_unbiased.)"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from yvb200 import ops
    r = ops.rt("cuda")
    cfg = synth.CONFIGS["micro"]
    args = synth.workload_args("micro")
    model = build_lily(cfg, args, device="cuda").train()
    batch = _to_dev(synth.make_batch("micro", seed=1), "cuda")
    inp = synth.model_inputs(batch)
    state = r.rng_state()
    o1 = model(*inp)["vision"].detach().clone()
    r.set_rng_state(state)
    o2 = model(*inp)["vision"].detach().clone()
    assert torch.equal(o1, o2)
    # every training forward advances the step counter by itself (the unmodified reference loop never touches the RNG):
    # two consecutive forwards draw different masks
    o3 = model(*inp)["vision"].detach().clone()
    assert not torch.equal(o1, o3)
    model.eval()
    oe = model(*inp)["vision"].detach()
    assert float((o1 - oe).norm() / oe.norm()) > 1e-3          # dropout really perturbs the output
    # backward runs in train mode and regenerates the same masks
    model.train()
    r.set_rng_state(state)
    torch.manual_seed(3)            # the [N,1024] pooled dropout of the task wrapper stays on torch's generator
    out = model(*inp)
    losses.total_loss(losses.step_losses(batch, out, args, True), args).backward()
    g1 = model.bert.encoder.layer[0].attention.self.query.weight.grad.clone()
    model.zero_grad()
    r.set_rng_state(state)
    torch.manual_seed(3)
    out = model(*inp)
    losses.total_loss(losses.step_losses(batch, out, args, True), args).backward()
    assert torch.allclose(g1, model.bert.encoder.layer[0].attention.self.query.weight.grad, rtol=1e-4, atol=1e-7)


def test_backward_regenerates_the_masks_of_its_own_forward():
    """Two training forwards before the first backward (e.g. two micro-batches summed into one loss): each backward must
    regenerate the dropout masks of ITS forward, not those of whichever forward ran last."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from yvb200 import ops
    r = ops.rt("cuda")
    cfg = synth.CONFIGS["micro"]
    args = synth.workload_args("micro")
    model = build_lily(cfg, args, device="cuda").train()
    batch = _to_dev(synth.make_batch("micro", seed=1), "cuda")
    inp = synth.model_inputs(batch)
    w = model.bert.encoder.layer[0].attention.self.query.weight
    state = r.rng_state()
    torch.manual_seed(3)
    out = model(*inp)
    losses.total_loss(losses.step_losses(batch, out, args, True), args).backward()
    want = w.grad.clone()
    model.zero_grad()
    r.set_rng_state(state)
    torch.manual_seed(3)
    out = model(*inp)
    loss = losses.total_loss(losses.step_losses(batch, out, args, True), args)
    with torch.no_grad():
        model(*inp)                      # an unrelated training-mode forward in between advances the counter
    loss.backward()
    assert torch.allclose(want, w.grad, rtol=1e-4, atol=1e-7)


def test_weight_planes_follow_in_place_data_updates():
    """The reference's AdamW updates ``p.data`` in place (vilbert/optimization.py:176-187), which does not bump the
    tensor version: the start-of-forward refresh must re-split the weights anyway."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg = synth.CONFIGS["micro"]
    args = synth.workload_args("micro")
    model = build_lily(cfg, args, device="cuda").eval()
    batch = _to_dev(synth.make_batch("micro", seed=1), "cuda")
    inp = synth.model_inputs(batch)
    o1 = model(*inp)["ranking"].detach().clone()
    w = model.bert.encoder.layer[0].attention.self.query.weight
    v0 = w._version
    w.data.mul_(1.5)                     # what an optimizer step through .data looks like
    assert w._version == v0
    o2 = model(*inp)["ranking"].detach().clone()          # autograd is on: every weight is re-split
    ref = build_lily(cfg, args, device="cuda").eval()
    ref.load_state_dict(model.state_dict())
    o3 = ref(*inp)["ranking"].detach()
    assert not torch.equal(o1, o2)
    assert torch.allclose(o2, o3, rtol=1e-5, atol=1e-6)
