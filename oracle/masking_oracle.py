"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's batch-masking functions, SURVEY.md 8(f)
"next" #3.  Pinned against outputs of the real reference (tests/golden/masking.npz, written by oracle/make_golden.py).
Only tests/ may import this module; the product path is youtube-vln_b200/yvb200/masking.py on the CUDA kernels.

The random draws the reference makes internally are explicit inputs here (``p``: uniform [0,1) per element,
``random``: replacement token ids, ``forced``: action-word positions picked by np.random.choice), so the functions are
deterministic integer / byte work and parity is bit-exact.
"""
import numpy as np

# thresholds exactly as the reference computes them (Python doubles, then compared against float32 tensors, i.e.
# cast to float32 by the comparison): utils/dataset/common.py:233,254,258,262 and :292,297
T_MASK = np.float32(0.85)
T_RANDOM = np.float32(0.85 + 0.15 * 0.8)
T_KEEP = np.float32(0.85 + 0.15 * 0.9)
T_ZERO = np.float32(0.85 + 0.15 * 0.1)
P_FORCED = np.float32(0.85 * 0.9)


def randomize_tokens(tokens, mask, p, random, mask_id, forced=None):
    """utils/dataset/common.py:213-270.  tokens i64 [R,T], mask bool [R,T], p f32 [R,T], random i64 [R,T]."""
    tokens = tokens.copy()
    targets = np.full_like(tokens, -1)
    pe = p.astype(np.float32) * mask.astype(np.float32)
    if forced is not None and forced.any():                      # :236-252 (mask_action_rate > 0)
        f = forced.astype(bool)
        targets[f] = tokens[f]
        tokens[f] = mask_id
        pe[f] = P_FORCED
    sel = pe >= T_MASK                                           # :254-258
    targets[sel] = tokens[sel]
    tokens[sel] = mask_id
    sel = pe >= T_RANDOM                                         # :260-262
    tokens[sel] = random[sel]
    sel = pe >= T_KEEP                                           # :264-266
    tokens[sel] = targets[sel]
    return tokens, targets


def randomize_regions(features, probs, mask, p):
    """utils/dataset/common.py:272-300.  features f32 [R,N,F], probs f32 [R,N,C], mask i64 [R,N], p f32 [R,N]."""
    features = features.copy()
    targets = np.ones_like(probs) / np.float32(probs.shape[-1])
    targets_mask = np.zeros_like(mask)
    pe = p.astype(np.float32) * mask.astype(np.float32)
    sel = pe >= T_MASK
    targets[sel] = probs[sel]
    targets_mask[sel] = 1
    features[pe >= T_ZERO] = 0
    return features, targets, targets_mask
