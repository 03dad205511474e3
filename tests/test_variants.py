"""Less-travelled paths of the reference, pinned by ``tests/golden/variants.npz`` (recorded from the unmodified reference
by ``oracle/make_golden.py variants``, narrow ``micro`` model, complete tensors):

  fixed/ : frozen text prefix (``fixed_t_layer``, vilbert/vilbert.py:745-764)
  attn/  : ``output_all_attention_masks=True`` (vilbert/vilbert.py:1242-1337) -- every attention-probability map
  vl/    : ``VILBertForVLTasks`` (vilbert/vilbert.py:1457-1520) -- the 7-tuple and its gradients

Each is checked on the host path of the drop-in (CPU, 2e-5 / 2e-4) and on the CUDA kernels (1e-3).  A further CPU test
runs the UNMODIFIED reference ``lily.py`` + ``utils/utils_init.py`` on top of the drop-in ``vilbert.vilbert`` when the
reference checkout is mounted (the build container), which is the drop-in claim itself."""
import os
import sys

import numpy as np
import pytest
import torch

from yvb200 import synth, losses
import vilbert.vilbert as V

REFERENCE_ROOT = "/root/reference"


def _golden(golden_dir):
    path = os.path.join(golden_dir, "variants.npz")
    if not os.path.exists(path):
        pytest.skip("variants.npz not generated")
    return np.load(path)


def _config(**over):
    cfg = dict(synth.MICRO_CONFIG, **over)
    c = V.BertConfig(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in cfg.items()})
    c.args = synth.workload_args("micro")
    return c


def _rel(a, ref):
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.linalg.norm(a - ref) / max(np.linalg.norm(ref), 1e-30))


def _lily(config, device):
    from yvb200.lily_compat import Lily
    torch.manual_seed(0)
    m = Lily(config)
    synth.load_synthetic_weights(m, seed=0)
    return m.to(device).eval()


def _batch(device):
    return [t.to(device) if torch.is_tensor(t) else t for t in synth.make_batch("micro", seed=1)]


def _check_fixed(g, device, tol_out, tol_grad):
    args = synth.workload_args("micro")
    model = _lily(_config(fixed_t_layer=1, fixed_v_layer=0), device)
    b = _batch(device)
    out = model(*synth.model_inputs(b))
    ld = losses.step_losses(b, out, args, training=True)
    tot = losses.total_loss(ld, args)
    tot.backward()
    for k, v in ld.items():
        ref = float(g[f"fixed/loss/{k}"])
        assert abs(float(v) - ref) <= tol_out * max(1.0, abs(ref)), k
    for k, v in out.items():
        assert _rel(v.detach().cpu().numpy(), g[f"fixed/out/{k}"]) < tol_out, k
    dead = {k[len("fixed/nograd/"):] for k in g.files if k.startswith("fixed/nograd/")}
    assert dead == {n for n, p in model.named_parameters() if p.grad is None}
    # the frozen text layer and the embeddings below it really are in the dead set
    assert any(n.startswith("bert.encoder.layer.0.") for n in dead) and "bert.embeddings.word_embeddings.weight" not in dead
    gmax = max(float(np.linalg.norm(g[k])) for k in g.files if k.startswith("fixed/grad/"))
    n = 0
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        ref = g[f"fixed/grad/{name}"]
        if np.linalg.norm(ref) < 1e-6 * gmax:
            continue
        assert _rel(p.grad.detach().cpu().numpy(), ref) < tol_grad, name
        n += 1
    assert n > 50


def _check_attn(g, device, tol):
    model = _lily(_config(), device)
    b = _batch(device)
    inp = synth.model_inputs(b)
    with torch.no_grad():
        seq_t, seq_v, pooled_t, pooled_v, (att_t, att_v, att_c) = model.bert(
            inp[0], inp[1], inp[2], inp[3], inp[4], inp[5], inp[6], output_all_encoded_layers=False,
            output_all_attention_masks=True)
    assert [len(att_t), len(att_v), len(att_c)] == list(g["attn/counts"])
    assert _rel(seq_t.cpu().numpy(), g["attn/seq_t"]) < tol and _rel(seq_v.cpu().numpy(), g["attn/seq_v"]) < tol
    assert _rel(pooled_t.cpu().numpy(), g["attn/pooled_t"]) < tol and _rel(pooled_v.cpu().numpy(), g["attn/pooled_v"]) < tol
    for i, a in enumerate(att_t):
        ref = g[f"attn/t{i}"]
        a = a[..., : ref.shape[-1]]                # (the CUDA path pads the key axis of its score buffer to 8)
        assert tuple(a.shape) == ref.shape and _rel(a.cpu().numpy(), ref) < tol, ("t", i)
    for i, a in enumerate(att_v):
        ref = g[f"attn/v{i}"]
        a = a[..., : ref.shape[-1]]
        assert tuple(a.shape) == ref.shape and _rel(a.cpu().numpy(), ref) < tol, ("v", i)
    for i, (a1, a2) in enumerate(att_c):
        for j, a in enumerate((a1, a2)):
            ref = g[f"attn/c{i}_{j}"]
            a = a[..., : ref.shape[-1]]
            assert tuple(a.shape) == ref.shape and _rel(a.cpu().numpy(), ref) < tol, ("c", i, j)
    # rows of every map are probability distributions
    assert float((att_t[0][..., : g["attn/t0"].shape[-1]].sum(-1) - 1).abs().max()) < 1e-4


def _check_vl(g, device, tol_out, tol_grad):
    torch.manual_seed(0)
    vl = V.VILBertForVLTasks(_config(), num_labels=3, default_gpu=False)
    synth.load_synthetic_weights(vl, seed=0)
    vl = vl.to(device).eval()
    b = _batch(device)
    inp = synth.model_inputs(b)
    outs = vl(inp[0], inp[1], inp[2], inp[3], inp[4], inp[5], inp[6])
    assert len(outs) == 7
    scalar = 0.0
    for i, o in enumerate(outs):
        ref = g[f"vl/out{i}"]
        assert tuple(o.shape) == ref.shape, i
        assert _rel(o.detach().cpu().numpy(), ref) < tol_out, i
        if i == 4:
            m = (inp[5] > 0).unsqueeze(2).float()
            scalar = scalar + ((o * m) ** 2).mean()
        else:
            scalar = scalar + (o ** 2).mean()
    scalar.backward()
    assert abs(float(scalar) - float(g["vl/scalar"])) < tol_out * abs(float(g["vl/scalar"]))
    dead = {k[len("vl/nograd/"):] for k in g.files if k.startswith("vl/nograd/")}
    assert dead == {n for n, p in vl.named_parameters() if p.grad is None}
    gmax = max(float(np.linalg.norm(g[k])) for k in g.files if k.startswith("vl/grad/"))
    n = 0
    for name, p in vl.named_parameters():
        if p.grad is None:
            continue
        ref = g[f"vl/grad/{name}"]
        if np.linalg.norm(ref) < 1e-6 * gmax:
            continue
        assert _rel(p.grad.detach().cpu().numpy(), ref) < tol_grad, name
        n += 1
    assert n > 50


# ------------------------------------------------------------------------------------------------- CPU (host path)
def test_frozen_prefix_host_path(golden_dir):
    _check_fixed(_golden(golden_dir), "cpu", 2e-5, 2e-4)


def test_all_attention_masks_host_path(golden_dir):
    _check_attn(_golden(golden_dir), "cpu", 2e-5)


def test_vl_tasks_host_path(golden_dir):
    _check_vl(_golden(golden_dir), "cpu", 2e-5, 2e-4)


def test_fixed_layer_asserts_like_the_reference():
    """vilbert/vilbert.py:742-743: a frozen prefix longer than the first co-attention position is an AssertionError."""
    model = _lily(_config(fixed_t_layer=2, fixed_v_layer=0), "cpu")
    b = _batch("cpu")
    with pytest.raises(AssertionError):
        model(*synth.model_inputs(b))


def test_unmodified_reference_lily_runs_on_the_dropin(golden_dir):
    """The drop-in claim: the reference's own ``lily.py`` and ``utils/utils_init.py`` (get_model_input,
    get_loss_correct), imported UNMODIFIED from the mounted checkout, on top of this repo's ``vilbert.vilbert``
    reproduce the golden vectors recorded from the all-reference run."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "vilbert")):
        pytest.skip("reference checkout not mounted (GPU box)")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import refload
    from test_oracle_golden import _check_grads, _check_outputs
    refload._install_stubs()
    saved = {k: sys.modules[k] for k in list(sys.modules) if k == "lily" or k == "utils" or k.startswith("utils.")}
    for k in saved:
        del sys.modules[k]
    sys.path.append(REFERENCE_ROOT)              # AFTER the drop-in: ``vilbert.vilbert`` stays this repo's module
    try:
        import importlib
        lily = importlib.import_module("lily")
        ui = importlib.import_module("utils.utils_init")
        assert lily.__file__.startswith(REFERENCE_ROOT) and ui.__file__.startswith(REFERENCE_ROOT)
        assert lily.ViLBertModel is V.BertModel, "lily.py must have picked up the drop-in vilbert.vilbert"
        g = np.load(os.path.join(golden_dir, "micro.npz"))
        args = synth.workload_args("micro")
        config = _config()
        torch.manual_seed(0)
        model = lily.Lily(config)
        synth.load_synthetic_weights(model, seed=0)
        model.eval()
        batch = tuple(synth.make_batch("micro", seed=1))
        outputs = model(*ui.get_model_input(batch))
        total = 0.0
        for task in ("vision", "language", "ranking", "traj"):
            _, _, loss, _ = ui.get_loss_correct(batch, outputs, task, args, None, True)
            assert abs(float(loss) - float(g[f"loss/{task}"])) <= 2e-5 * max(1.0, abs(float(g[f"loss/{task}"]))), task
            total = total + (args.traj_loss_scale * loss if task == "traj" else loss)
        total.backward()
        _check_outputs(g, {k: v.detach() for k, v in outputs.items()}, 2e-5)
        grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
        assert _check_grads(g, grads, 2e-4) > 100
        assert len(model.state_dict()) == len([k for k in g.files if k.startswith("grad/") or k.startswith("nograd/")]) + 1
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == "lily" or k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]
        sys.modules.update(saved)


# ------------------------------------------------------------------------------------------------- GPU (CUDA kernels)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.mark.gpu
def test_frozen_prefix_cuda(golden_dir):
    _need_gpu()
    _check_fixed(_golden(golden_dir), "cuda", 1e-3, 1e-3)


@pytest.mark.gpu
def test_all_attention_masks_cuda(golden_dir):
    _need_gpu()
    _check_attn(_golden(golden_dir), "cuda", 1e-3)


@pytest.mark.gpu
def test_vl_tasks_cuda(golden_dir):
    _need_gpu()
    _check_vl(_golden(golden_dir), "cuda", 1e-3, 1e-3)
