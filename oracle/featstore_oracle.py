"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's feature-store read path (SURVEY.md 8f "next" #4):
utils/dataset/features_reader.py:124-150 (record decoding, both key conventions), :84-118 (box / location encoding) and
:153-182 (trajectory assembly with the global mean-feature row).  Pinned bit-exact against outputs of the real reader
(tests/golden/featstore.npz, written by oracle/make_golden.py).  Only tests/ and tools/featstore_bench.py's baseline leg
may import this module; the product path is youtube-vln_b200/yvb200/featstore.py.
"""
import base64
import pickle

import numpy as np


def decode_item(item: dict):
    """features_reader.py:124-150: (features [K,2048], boxes [K,4], cls_prob [K,1601], image_w, image_h)."""
    old = "image_width" in item
    w = int(item["image_width" if old else "image_w"])
    h = int(item["image_height" if old else "image_h"])
    raw = (lambda k_old, k_new: item[k_old] if old else base64.b64decode(item[k_new]))
    features = np.frombuffer(raw("feature", "features"), dtype=np.float32).reshape((-1, 2048))
    boxes = np.frombuffer(raw("bbox", "boxes"), dtype=np.float32).reshape((-1, 4))
    cls_prob = np.frombuffer(raw("cls_prob", "cls_prob"), dtype=np.float32).reshape((-1, 1601))
    return features, boxes, cls_prob, w, h


def encode_boxes(boxes: np.ndarray, w: int, h: int) -> np.ndarray:
    """features_reader.py:84-103: [x1/w, y1/h, x2/w, y2/h, area/(w*h)] in float32."""
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    area /= w * h
    out = np.zeros(shape=(len(boxes), 5), dtype=np.float32)
    out[:, 0] = boxes[:, 0] / w
    out[:, 1] = boxes[:, 1] / h
    out[:, 2] = boxes[:, 2] / w
    out[:, 3] = boxes[:, 3] / h
    out[:, 4] = area
    return out


def assemble(items):
    """features_reader.py:153-182 for a list of raw (already unpickled) items of one trajectory."""
    l_boxes, l_probs, l_features = [], [], []
    for item in items:
        f, b, p, w, h = decode_item(item)
        l_boxes.append(encode_boxes(b, w, h))
        l_probs.append(p)
        l_features.append(f)
    features = np.concatenate(l_features, axis=0)
    boxes = np.concatenate(l_boxes, axis=0)
    probs = np.concatenate(l_probs, axis=0)
    locations = np.ones(shape=(len(boxes), 11), dtype=np.float32)
    locations[:, 0:5] = boxes[:, 0:5]
    if features.size == 0:
        raise RuntimeError("Features could not be correctly read")
    g_feature = features.mean(axis=0, keepdims=True)
    g_location = np.array([[0, 0, 1, 1, 1, 0, 1, 0, 1, 0, 1, ]])
    g_prob = np.ones(shape=(1, 1601)) / 1601
    return (np.concatenate([g_feature, features], axis=0), np.concatenate([g_location, locations], axis=0),
            np.concatenate([g_prob, probs], axis=0))


def read(store: dict, keys):
    """``store``: key -> pickled record bytes (what the LMDB holds)."""
    return assemble([pickle.loads(store[k]) for k in keys])


def visual_features(store: dict, steps, max_path_length: int, max_num_boxes: int):
    """utils/dataset/all_dataset.py:294-345 (+ the float32 / int64 conversion of :236-239): ``steps`` is a list of key
    tuples (one tuple of frame keys per trajectory step).  float64 staging arrays exactly as the reference."""
    path_length = min(len(steps), max_path_length)
    pf, pb, pp, pm = [], [], [], []
    for i, keys in enumerate(steps):
        features, boxes, probs = read(store, keys)
        nb = min(len(boxes), max_num_boxes)
        f = np.zeros((max_num_boxes, 2048)); f[:nb] = features[:nb]
        b = np.zeros((max_num_boxes, 12)); b[:nb, :11] = boxes[:nb, :11]; b[:, 11] = np.ones(max_num_boxes) * i
        p = np.zeros((max_num_boxes, 1601)); p[:nb] = probs[:nb]
        pf.append(f); pb.append(b); pp.append(p); pm.append([1] * nb + [0] * (max_num_boxes - nb))
    for idx in range(path_length, max_path_length):
        b = np.zeros((max_num_boxes, 12)); b[:, 11] = np.ones(max_num_boxes) * idx
        pf.append(np.zeros((max_num_boxes, 2048))); pb.append(b); pp.append(np.zeros((max_num_boxes, 1601)))
        pm.append([0] * max_num_boxes)
    return (np.vstack(pf).astype(np.float32), np.vstack(pb).astype(np.float32), np.vstack(pp).astype(np.float32),
            np.hstack(pm).astype(np.int64))
