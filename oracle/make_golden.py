"""TEST INFRASTRUCTURE ONLY -- generate golden vectors from the *real* reference.

Run in the build container (``/root/reference`` mounted):

    python oracle/make_golden.py [micro cfg1 cfg2]

For each workload it instantiates the reference's own ``Lily`` (lily.py:23) on the reference's own
``BertConfig`` (vilbert/vilbert.py:129), loads the name-keyed synthetic weights of ``yvb200.synth``,
runs ``model.eval()`` forward through the reference's ``get_model_input`` (utils/utils_init.py:34),
evaluates the four losses with the reference's ``get_loss_correct`` (utils/utils_init.py:108) in the
``training=True`` branch, sums them like ``train_epoch`` (utils/utils_init.py:217-224), calls
``loss.backward()`` and stores:

  micro : complete outputs + complete gradients (narrow model, a few hundred KB)
  cfg1/2: ranking/traj logits, the loss scalars, and for the big tensors (vision / language outputs,
          every parameter gradient) the L2 norm plus values at seeded sample positions.

The files land in ``tests/golden/<workload>.npz`` and are committed; the GPU box never needs
``/root/reference``.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
sys.path.insert(0, HERE)

from yvb200 import synth  # noqa: E402
import refload  # noqa: E402

N_SAMPLE_OUT = 4096
N_SAMPLE_GRAD = 64


def sample_positions(name: str, numel: int, k: int) -> np.ndarray:
    g = torch.Generator().manual_seed(synth._seed_of("sample:" + name, 7))
    if numel <= k:
        return np.arange(numel, dtype=np.int64)
    return torch.randint(0, numel, (k,), generator=g).numpy().astype(np.int64)


def run(workload: str):
    vb, lily, ui = refload.load_reference()
    w = synth.WORKLOADS[workload]
    cfgd = synth.CONFIGS[w["config"]]
    args = synth.workload_args(workload)
    config = vb.BertConfig(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in cfgd.items()})
    config.args = args
    torch.manual_seed(0)
    model = lily.Lily(config)
    synth.load_synthetic_weights(model, seed=0)
    model.eval()
    batch = synth.make_batch(workload, seed=1)
    t0 = time.time()
    outputs = model(*ui.get_model_input(tuple(batch)))
    total = 0.0
    loss_vals = {}
    for task in ("vision", "language", "ranking", "traj"):
        if task not in outputs:                 # objective switched off for this workload (ranking-only fine-tune)
            continue
        _, _, loss, _ = ui.get_loss_correct(tuple(batch), outputs, task, args, None, True)
        loss_vals[task] = float(loss.detach())
        total = total + (args.traj_loss_scale * loss if task == "traj" else loss)
    model.zero_grad()
    total.backward()
    dt = time.time() - t0
    out = {"total_loss": np.float64(float(total.detach())), "ref_seconds": np.float64(dt)}
    for k, v in loss_vals.items():
        out["loss/" + k] = np.float64(v)
    full = workload == "micro"        # (micro_pad is stored sampled: it pins the padding / degenerate-mask paths)
    for k, v in outputs.items():
        a = v.detach().float().numpy()
        if full or a.size <= N_SAMPLE_OUT:
            out["out/" + k] = a
        else:
            pos = sample_positions("out/" + k, a.size, N_SAMPLE_OUT)
            out["outpos/" + k] = pos
            out["outval/" + k] = a.reshape(-1)[pos]
            out["outnorm/" + k] = np.float64(np.linalg.norm(a.astype(np.float64)))
    n_grad = 0
    for name, p in model.named_parameters():
        if p.grad is None:
            out["nograd/" + name] = np.int8(1)
            continue
        gr = p.grad.detach().float().numpy()
        n_grad += 1
        if full:
            out["grad/" + name] = gr
        else:
            pos = sample_positions("grad/" + name, gr.size, N_SAMPLE_GRAD)
            out["gradpos/" + name] = pos
            out["gradval/" + name] = gr.reshape(-1)[pos]
            out["gradnorm/" + name] = np.float64(np.linalg.norm(gr.astype(np.float64)))
    path = os.path.join(ROOT, "tests", "golden", f"{workload}.npz")
    np.savez_compressed(path, **out)
    print(f"{workload}: total_loss={float(total):.6f} losses={loss_vals} params_with_grad={n_grad} "
          f"ref fwd+bwd {dt:.1f}s -> {path} ({os.path.getsize(path)/1e3:.0f} KB)")


def run_adamw():
    """3 steps of the reference's own AdamW (vilbert/optimization.py:107) with the reference's param grouping
    (vilbert/vilbert_init.py:9-18) on a handful of named tensors and seeded gradients."""
    import importlib
    refload.load_reference()
    sys.path.insert(0, refload.REFERENCE_ROOT)
    opt = importlib.import_module("vilbert.optimization")
    sys.path.remove(refload.REFERENCE_ROOT)
    names = {"enc.dense.weight": (48, 40), "enc.dense.bias": (48,), "enc.LayerNorm.weight": (48,),
             "enc.LayerNorm.bias": (48,), "emb.word_embeddings.weight": (37, 24)}
    g = torch.Generator().manual_seed(123)
    params = {n: torch.nn.Parameter(torch.randn(s, generator=g) * 0.05) for n, s in names.items()}
    groups = [{"params": [], "weight_decay": 0.0}, {"params": [], "weight_decay": 0.01}]
    for n, p_ in params.items():
        groups[0 if any(nd in n for nd in ("bias", "LayerNorm.weight", "LayerNorm.bias")) else 1]["params"].append(p_)
    optim = opt.AdamW(groups, lr=4e-5)
    out = {}
    for n, p_ in params.items():
        out["p0/" + n] = p_.detach().numpy().copy()
    lrs = [4e-5, 3e-5, 2.5e-5]
    for step in range(3):
        for grp in optim.param_groups:
            grp["lr"] = lrs[step]
        for n, p_ in params.items():
            p_.grad = torch.randn(p_.shape, generator=g) * (0.01 if step != 1 else 1e-4)
            out[f"g{step}/" + n] = p_.grad.numpy().copy()
        optim.step()
        for n, p_ in params.items():
            out[f"p{step + 1}/" + n] = p_.detach().numpy().copy()
    for n, p_ in params.items():
        out["m3/" + n] = optim.state[p_]["exp_avg"].numpy().copy()
        out["v3/" + n] = optim.state[p_]["exp_avg_sq"].numpy().copy()
    out["lrs"] = np.array(lrs)
    path = os.path.join(ROOT, "tests", "golden", "adamw.npz")
    np.savez_compressed(path, **out)
    print("adamw golden ->", path)


def run_masking():
    """The reference's own randomize_tokens / randomize_regions (utils/dataset/common.py:213-300) on seeded inputs.
    The random draws the functions make internally (torch.rand_like, torch.randint_like, np.random.choice) are
    recorded by replaying the generator state, so that a device implementation fed with the same draws must
    reproduce the outputs bit for bit."""
    import importlib
    import types
    refload.load_reference()
    sys.path.insert(0, refload.REFERENCE_ROOT)
    common = importlib.import_module("utils.dataset.common")
    sys.path.remove(refload.REFERENCE_ROOT)
    vocab = {str(i): i for i in range(30521)}
    vocab["[MASK]"] = 103
    tokenizer = types.SimpleNamespace(vocab=vocab)
    out = {"vocab_size": np.array(len(vocab)), "mask_id": np.array(103)}
    for case, rate in (("plain", 0.0), ("actions", 0.5)):
        torch.manual_seed(11 if rate == 0.0 else 12)
        np.random.seed(5)
        tokens = torch.randint(1000, 30000, (6, 20))
        tokens[:, 0] = 101
        tokens[1, 15:] = 0
        tokens[4, 9:] = 0
        for (r, c, a) in ((0, 3, 2187), (0, 7, 2830), (2, 5, 2157), (3, 11, 2187), (5, 2, 2830), (5, 18, 2157)):
            tokens[r, c] = a
        mask = tokens > 0
        args = types.SimpleNamespace(mask_action_rate=rate)
        st, nst = torch.get_rng_state(), np.random.get_state()
        pr = torch.rand_like(tokens.float())
        rnd = torch.randint_like(tokens, len(vocab))
        forced = torch.zeros_like(tokens, dtype=torch.uint8)
        if rate > 0:
            xs, ys = [], []
            for ac in (2187, 2830, 2157):
                ix = torch.where(tokens == ac)
                xs.append(ix[0]); ys.append(ix[1])
            xs, ys = torch.cat(xs), torch.cat(ys)
            for mi in np.random.choice(range(len(xs)), int(rate * len(xs))):
                forced[xs[mi], ys[mi]] = 1
        torch.set_rng_state(st); np.random.set_state(nst)
        o_tok, o_tgt = common.randomize_tokens(tokens.clone(), mask, tokenizer, args)
        out.update({f"tok/{case}/tokens": tokens.numpy(), f"tok/{case}/mask": mask.numpy(), f"tok/{case}/p": pr.numpy(),
                    f"tok/{case}/random": rnd.numpy(), f"tok/{case}/forced": forced.numpy(),
                    f"tok/{case}/out_tokens": o_tok.numpy(), f"tok/{case}/out_targets": o_tgt.numpy()})
    torch.manual_seed(21)
    feats = torch.randn(5, 36, 16)
    probs = torch.softmax(torch.randn(5, 36, 12), -1)
    rmask = torch.ones(5, 36, dtype=torch.long)
    rmask[1, 30:] = 0
    rmask[3, 18:] = 0
    st = torch.get_rng_state()
    pr = torch.rand_like(rmask.float())
    torch.set_rng_state(st)
    o_f, o_t, o_m = common.randomize_regions(feats.clone(), probs, rmask)
    out.update({"reg/features": feats.numpy(), "reg/probs": probs.numpy(), "reg/mask": rmask.numpy(), "reg/p": pr.numpy(),
                "reg/out_features": o_f.numpy(), "reg/out_targets": o_t.numpy(), "reg/out_targets_mask": o_m.numpy()})
    path = os.path.join(ROOT, "tests", "golden", "masking.npz")
    np.savez_compressed(path, **out)
    print("masking golden ->", path, f"({os.path.getsize(path)/1e3:.0f} KB)")


def run_featstore():
    """The reference's own YTbFeaturesReader.__getitem__ (utils/dataset/features_reader.py:153-182) on synthetic
    records in both key conventions (raw float32 blobs / base64 strings).  The LMDB layer is replaced by a dict: the
    base class lookup is patched to return the unpickled records, everything above it is the reference's code."""
    import base64
    import importlib
    import pickle
    refload.load_reference()
    sys.path.insert(0, refload.REFERENCE_ROOT)
    fr = importlib.import_module("utils.dataset.features_reader")
    sys.path.remove(refload.REFERENCE_ROOT)
    rng = np.random.RandomState(7)
    store = {}
    for i, (key, k, old) in enumerate((("vidA/000012", 3, True), ("vidA/000030", 2, False), ("vidB/000001", 4, False),
                                       ("vidB/000077", 1, True))):
        w, h = 640 + 16 * i, 360 + 8 * i
        feats = rng.randn(k, 2048).astype(np.float32)
        x1 = rng.uniform(0, w / 2, k); y1 = rng.uniform(0, h / 2, k)
        boxes = np.stack([x1, y1, x1 + rng.uniform(8, w / 2, k), y1 + rng.uniform(8, h / 2, k)], 1).astype(np.float32)
        probs = rng.dirichlet(np.ones(1601) * 0.1, k).astype(np.float32)
        if old:
            item = {"image_width": w, "image_height": h, "feature": feats.tobytes(), "bbox": boxes.tobytes(),
                    "cls_prob": probs.tobytes()}
        else:
            item = {"image_w": str(w), "image_h": str(h), "features": base64.b64encode(feats.tobytes()),
                    "boxes": base64.b64encode(boxes.tobytes()), "cls_prob": base64.b64encode(probs.tobytes())}
        store[key] = pickle.dumps(item)
    reader = object.__new__(fr.YTbFeaturesReader)
    reader.keys = {k: 0 for k in store}
    fr.FeaturesReader.__getitem__ = lambda self, keys: [pickle.loads(store[k]) for k in keys]
    out = {"keys": np.array(list(store)), "records": np.array([np.frombuffer(v, dtype=np.uint8) for v in store.values()],
                                                                dtype=object)}
    queries = {"q0": ("vidA/000012", "vidA/000030", "vidB/000001"), "q1": ("vidB/000077",),
               "q2": ("vidB/000001", "vidA/000012", "vidB/000001", "vidB/000077")}
    for name, q in queries.items():
        f, l, p = reader[q]
        out[f"{name}/query"] = np.array(q)
        out[f"{name}/features"], out[f"{name}/locations"], out[f"{name}/probs"] = f, l, p
    # trajectory assembly: the reference's BaseDataset._get_visual_features (utils/dataset/all_dataset.py:294-345) on
    # top of that reader, then the float32 / int64 conversion of __getitem__ (:236-239)
    import inspect
    import types
    inspect.ArgSpec = getattr(inspect, "ArgSpec", inspect.FullArgSpec)      # removed from Python 3.11 (import-time only)
    sys.path.insert(0, refload.REFERENCE_ROOT)
    ds = importlib.import_module("utils.dataset.all_dataset")
    sys.path.remove(refload.REFERENCE_ROOT)
    fake = types.SimpleNamespace(args=types.SimpleNamespace(max_path_length=4, max_num_boxes=4), _features_reader=reader,
                                 get_feature_key=lambda listing, pid: f"{listing}/{pid:06d}")
    trajectories = {"t0": [("vidA", 12), ("vidB", (1, 77)), ("vidA", 30)], "t1": [("vidB", 77)],
                    "t2": [("vidA", 30), ("vidA", 12), ("vidB", 1), ("vidB", 77)]}
    for name, traj in trajectories.items():
        f, b, p, m = ds.BaseDataset._get_visual_features(fake, traj)
        out[f"{name}/steps"] = np.array(["|".join(fake.get_feature_key(l, q) for q in ((pid,) if isinstance(pid, int) else pid))
                                         for l, pid in traj])
        out[f"{name}/features"] = torch.from_numpy(np.array(f)).float().numpy()
        out[f"{name}/boxes"] = torch.from_numpy(np.array(b)).float().numpy()
        out[f"{name}/probs"] = torch.from_numpy(np.array(p)).float().numpy()
        out[f"{name}/masks"] = torch.from_numpy(np.array(m)).long().numpy()
    path = os.path.join(ROOT, "tests", "golden", "featstore.npz")
    np.savez_compressed(path, **out)
    print("featstore golden ->", path, f"({os.path.getsize(path)/1e3:.0f} KB)")


def run_variants():
    """Less-travelled paths of the reference on the narrow ``micro`` model, complete tensors:
      fixed/ : ``Lily`` with ``fixed_t_layer=1`` (frozen text prefix, vilbert/vilbert.py:745-764): outputs, losses, every
               gradient, and which parameters get none;
      attn/  : ``BertModel.forward(..., output_all_attention_masks=True)`` (vilbert/vilbert.py:1242-1337): every
               attention-probability map of the three layer families plus the two pooled outputs;
      vl/    : ``VILBertForVLTasks`` (vilbert/vilbert.py:1457-1520): the 7-tuple and the gradients of
               sum_i mean(out_i^2) (the -10000 mask term of ``vision_logit`` is left out of the scalar)."""
    vb, lily, ui = refload.load_reference()
    wl = "micro"
    cfgd = dict(synth.CONFIGS["micro"])
    args = synth.workload_args(wl)
    batch = synth.make_batch(wl, seed=1)
    out = {}

    def mk_config(**over):
        c = vb.BertConfig(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in dict(cfgd, **over).items()})
        c.args = args
        return c

    # ---- frozen text prefix
    torch.manual_seed(0)
    model = lily.Lily(mk_config(fixed_t_layer=1, fixed_v_layer=0))
    synth.load_synthetic_weights(model, seed=0)
    model.eval()
    outputs = model(*ui.get_model_input(tuple(batch)))
    total = 0.0
    for task in ("vision", "language", "ranking", "traj"):
        _, _, loss, _ = ui.get_loss_correct(tuple(batch), outputs, task, args, None, True)
        out["fixed/loss/" + task] = np.float64(float(loss.detach()))
        total = total + (args.traj_loss_scale * loss if task == "traj" else loss)
    model.zero_grad()
    total.backward()
    out["fixed/total_loss"] = np.float64(float(total.detach()))
    for k, v in outputs.items():
        out["fixed/out/" + k] = v.detach().float().numpy()
    for name, p in model.named_parameters():
        if p.grad is None:
            out["fixed/nograd/" + name] = np.int8(1)
        else:
            out["fixed/grad/" + name] = p.grad.detach().float().numpy()

    # ---- attention maps
    torch.manual_seed(0)
    model = lily.Lily(mk_config())
    synth.load_synthetic_weights(model, seed=0)
    model.eval()
    inp = ui.get_model_input(tuple(batch))
    with torch.no_grad():
        seq_t, seq_v, pooled_t, pooled_v, (att_t, att_v, att_c) = model.bert(
            inp[0], inp[1], inp[2], inp[3], inp[4], inp[5], inp[6], output_all_encoded_layers=False,
            output_all_attention_masks=True)
    out["attn/seq_t"], out["attn/seq_v"] = seq_t.numpy(), seq_v.numpy()
    out["attn/pooled_t"], out["attn/pooled_v"] = pooled_t.numpy(), pooled_v.numpy()
    for i, a in enumerate(att_t):
        out[f"attn/t{i}"] = a.numpy()
    for i, a in enumerate(att_v):
        out[f"attn/v{i}"] = a.numpy()
    for i, (a1, a2) in enumerate(att_c):
        out[f"attn/c{i}_0"], out[f"attn/c{i}_1"] = a1.numpy(), a2.numpy()
    out["attn/counts"] = np.array([len(att_t), len(att_v), len(att_c)])

    # ---- VILBertForVLTasks
    torch.manual_seed(0)
    vl = vb.VILBertForVLTasks(mk_config(), num_labels=3, default_gpu=False)
    synth.load_synthetic_weights(vl, seed=0)
    vl.eval()
    outs = vl(inp[0], inp[1], inp[2], inp[3], inp[4], inp[5], inp[6])
    scalar = 0.0
    for i, o in enumerate(outs):
        out[f"vl/out{i}"] = o.detach().float().numpy()
        if i == 4:      # vision_logit carries the additive -10000 mask: keep the unmasked entries only
            m = (inp[5] > 0).unsqueeze(2).float()
            scalar = scalar + ((o * m) ** 2).mean()
        else:
            scalar = scalar + (o ** 2).mean()
    vl.zero_grad()
    scalar.backward()
    out["vl/scalar"] = np.float64(float(scalar.detach()))
    for name, p in vl.named_parameters():
        if p.grad is None:
            out["vl/nograd/" + name] = np.int8(1)
        else:
            out["vl/grad/" + name] = p.grad.detach().float().numpy()
    path = os.path.join(ROOT, "tests", "golden", "variants.npz")
    np.savez_compressed(path, **out)
    print("variants golden ->", path, f"({os.path.getsize(path)/1e3:.0f} KB)")


if __name__ == "__main__":
    for wl in (sys.argv[1:] or ["micro", "cfg1", "cfg2", "adamw", "masking", "featstore", "variants"]):
        if wl == "variants":
            run_variants()
        elif wl == "adamw":
            run_adamw()
        elif wl == "featstore":
            run_featstore()
        elif wl == "masking":
            run_masking()
        else:
            run(wl)
