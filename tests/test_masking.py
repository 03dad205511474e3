"""Batch masking (SURVEY.md 8f "next" #3): the numpy oracle against golden vectors recorded from the reference's own
randomize_tokens / randomize_regions (CPU), and the CUDA kernels against both (bit-exact: integer / byte work)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import masking_oracle as MO  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "masking.npz"))


@pytest.mark.parametrize("case", ["plain", "actions"])
def test_oracle_tokens_match_reference(case):
    t, tg = MO.randomize_tokens(G[f"tok/{case}/tokens"], G[f"tok/{case}/mask"], G[f"tok/{case}/p"], G[f"tok/{case}/random"],
                                int(G["mask_id"]), G[f"tok/{case}/forced"])
    assert np.array_equal(t, G[f"tok/{case}/out_tokens"])
    assert np.array_equal(tg, G[f"tok/{case}/out_targets"])
    assert (tg != -1).sum() > 0


def test_oracle_regions_match_reference():
    f, t, m = MO.randomize_regions(G["reg/features"], G["reg/probs"], G["reg/mask"], G["reg/p"])
    assert np.array_equal(f, G["reg/out_features"])
    assert np.array_equal(t, G["reg/out_targets"])
    assert np.array_equal(m, G["reg/out_targets_mask"])
    assert m.sum() > 0 and (m[G["reg/mask"] == 0] == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["plain", "actions"])
def test_cuda_tokens_bit_exact(case):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from yvb200 import masking
    dev = "cuda"
    tok = torch.from_numpy(G[f"tok/{case}/tokens"]).to(dev)
    out_t, out_tg = masking.mask_tokens(tok.clone(), torch.from_numpy(G[f"tok/{case}/mask"]).to(dev),
                                        torch.from_numpy(G[f"tok/{case}/p"]).to(dev),
                                        torch.from_numpy(G[f"tok/{case}/random"]).to(dev), int(G["mask_id"]),
                                        torch.from_numpy(G[f"tok/{case}/forced"]).to(dev))
    assert np.array_equal(out_t.cpu().numpy(), G[f"tok/{case}/out_tokens"])
    assert np.array_equal(out_tg.cpu().numpy(), G[f"tok/{case}/out_targets"])


@pytest.mark.gpu
def test_cuda_regions_bit_exact_and_full_size():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from yvb200 import masking
    dev = "cuda"
    f, t, m = masking.mask_regions(torch.from_numpy(G["reg/features"]).to(dev), torch.from_numpy(G["reg/probs"]).to(dev),
                                   torch.from_numpy(G["reg/mask"]).to(dev), torch.from_numpy(G["reg/p"]).to(dev))
    assert np.array_equal(f.cpu().numpy(), G["reg/out_features"])
    assert np.array_equal(t.cpu().numpy(), G["reg/out_targets"])
    assert np.array_equal(m.cpu().numpy(), G["reg/out_targets_mask"])
    # full cfg2 size against the oracle: 8 pairs x 288 regions x 2048 features / 1601 classes, ragged masks
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(8, 288, 2048, generator=g)
    probs = torch.softmax(torch.randn(8, 288, 1601, generator=g), -1)
    mask = torch.ones(8, 288, dtype=torch.long)
    mask[1::2, 252:] = 0
    p = torch.rand(8, 288, generator=g)
    wf, wt, wm = MO.randomize_regions(feats.numpy(), probs.numpy(), mask.numpy(), p.numpy())
    f, t, m = masking.mask_regions(feats.to(dev), probs.to(dev), mask.to(dev), p.to(dev))
    assert np.array_equal(f.cpu().numpy(), wf) and np.array_equal(t.cpu().numpy(), wt) and np.array_equal(m.cpu().numpy(), wm)
    # the drop-in entry points with the reference's signatures draw their own numbers on the device
    import types
    tok = torch.randint(1000, 30000, (8, 80), generator=g)
    tok[:, 70:] = 0
    tokenizer = types.SimpleNamespace(vocab={**{str(i): i for i in range(30521)}, "[MASK]": 103})
    o, tg = masking.randomize_tokens(tok.to(dev), (tok > 0).to(dev), tokenizer, types.SimpleNamespace(mask_action_rate=0.0))
    sel = tg != -1
    assert 0.05 < float(sel.float().sum() / (tok > 0).sum()) < 0.3          # ~15 % of the real tokens are supervised
    assert bool((tg[sel] == tok.to(dev)[sel]).all()) and bool((tg[tok.to(dev) == 0] == -1).all())
    assert abs(float((o[sel] == 103).float().mean()) - 0.8) < 0.15          # 80 % [MASK], 10 % random, 10 % kept
