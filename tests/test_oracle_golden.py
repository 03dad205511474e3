"""CPU: the oracle (oracle/vilbert_oracle.py) is pinned against golden vectors recorded from the real
reference (oracle/make_golden.py).  Tolerances: fp32 oracle vs fp32 reference, both on the host, differ
only by summation order -> 2e-5 relative on outputs/losses, 2e-4 relative-to-norm on gradients."""
import os

import numpy as np
import pytest
import torch

from yvb200 import synth
import vilbert_oracle as O


def _load(golden_dir, wl):
    path = os.path.join(golden_dir, f"{wl}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    return np.load(path)


def _check_outputs(g, outs, rtol):
    for k, v in outs.items():
        a = v.float().numpy()
        if f"out/{k}" in g.files:
            ref = g[f"out/{k}"]
            err = np.linalg.norm(a - ref) / max(np.linalg.norm(ref), 1e-30)
        else:
            ref = g[f"outval/{k}"]
            got = a.reshape(-1)[g[f"outpos/{k}"]]
            err = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
            nerr = abs(np.linalg.norm(a.astype(np.float64)) - float(g[f"outnorm/{k}"])) / float(g[f"outnorm/{k}"])
            assert nerr < rtol, (k, nerr)
        assert err < rtol, (k, err)


def _check_grads(g, grads, rtol):
    """Key-projection biases have an analytically zero gradient (softmax is shift invariant), so tensors
    whose reference norm is below 1e-6 of the largest gradient norm are only required to stay that small."""
    n = 0
    norms = [float(np.linalg.norm(g[k])) if k.startswith("grad/") else float(g[k])
             for k in g.files if k.startswith("grad/") or k.startswith("gradnorm/")]
    floor = 1e-6 * max(norms)
    for name, gr in grads.items():
        a = gr.float().numpy()
        refnorm = (float(np.linalg.norm(g[f"grad/{name}"])) if f"grad/{name}" in g.files
                   else float(g[f"gradnorm/{name}"]) if f"gradnorm/{name}" in g.files else None)
        if refnorm is not None and refnorm < floor:
            assert float(np.linalg.norm(a)) < 10 * floor, name
            continue
        if f"grad/{name}" in g.files:
            ref = g[f"grad/{name}"]
            err = np.linalg.norm(a - ref) / max(np.linalg.norm(ref), 1e-30)
        elif f"gradval/{name}" in g.files:
            ref = g[f"gradval/{name}"]
            norm = float(g[f"gradnorm/{name}"])
            got = a.reshape(-1)[g[f"gradpos/{name}"]]
            scale = norm / np.sqrt(a.size)           # rms magnitude of an entry
            err = np.linalg.norm(got - ref) / max(np.sqrt(len(ref)) * scale, 1e-30)
            nerr = abs(np.linalg.norm(a.astype(np.float64)) - norm) / max(norm, 1e-30)
            assert nerr < rtol, (name, nerr)
        else:
            # the reference produced no gradient for this tensor: ours must be absent or exactly zero
            assert f"nograd/{name}" in g.files, name
            assert float(np.abs(a).max()) == 0.0, name
            continue
        assert err < rtol, (name, err)
        n += 1
    return n


@pytest.mark.parametrize("wl", ["micro", "micro_pad", "cfg1", "cfg4_p32_n2", "cfg3_rank"])
def test_oracle_matches_reference_golden(golden_dir, wl):
    g = _load(golden_dir, wl)
    cfg = synth.CONFIGS[synth.WORKLOADS[wl]["config"]]
    args = synth.workload_args(wl)
    sd = synth.lily_state_dict(cfg, seed=0)
    batch = synth.make_batch(wl, seed=1)
    outs, ld, tot, grads = O.oracle_step(sd, cfg, args, batch, dtype=torch.float32)
    assert {f"loss/{k}" for k in ld} == {k for k in g.files if k.startswith("loss/")}
    for k, v in ld.items():
        assert abs(float(v) - float(g[f"loss/{k}"])) <= 2e-5 * max(1.0, abs(float(g[f"loss/{k}"]))), k
    assert abs(float(tot) - float(g["total_loss"])) <= 2e-5 * abs(float(g["total_loss"]))
    _check_outputs(g, outs, 2e-5)
    n = _check_grads(g, grads, 2e-4)
    assert n > 100
    # dead parameters of the reference (q_dense1/2, bi_seq_relationship) get no gradient there
    dead = [k[len("nograd/"):] for k in g.files if k.startswith("nograd/")]
    assert any("q_dense1" in d for d in dead) and any("bi_seq_relationship" in d for d in dead)


def test_schema_is_542_keys_for_full_config():
    shapes = synth.lily_param_shapes(synth.FULL_CONFIG)
    assert len(shapes) == 542
    numel = sum(int(np.prod(s)) for k, s in shapes.items() if k != "cls.predictions.decoder.weight")
    assert numel == 250_087_039
