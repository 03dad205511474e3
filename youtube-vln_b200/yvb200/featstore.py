"""Contiguous feature shards (SURVEY.md 8f "next" #4) -- a replacement for the reference's per-frame LMDB of pickled
dicts (utils/dataset/features_reader.py; writer scripts/video_process/convert_to_lmdb.py:70-95).

The reference stores one pickled dict per frame whose arrays are raw float32 blobs or base64 strings; every read
unpickles, base64-decodes, re-encodes the boxes and concatenates.  A shard holds the same information decoded ONCE:

    header   : magic ``YVFS0001``, counts, section offsets (little-endian u64)
    index    : int64 [n + 1] row offsets (frame i owns rows [off[i], off[i+1])) and the utf-8 keys
    features : float32 [rows, 2048]      region features, exactly the reference's bytes
    boxes5   : float32 [rows, 5]         [x1/w, y1/h, x2/w, y2/h, area/(w*h)] as features_reader.py:84-103 computes it
    probs    : float32 [rows, 1601]      object-class probabilities

Sections are 4096-byte aligned, so a reader can ``np.memmap`` them (or hand the file to a pinned / GPUDirect copy) and
a trajectory is a handful of contiguous row ranges.  ``ShardReader.__getitem__(keys)`` returns what
``BaseFeaturesReader.__getitem__`` returns for the same keys -- same dtypes, same global mean-feature row -- bit for bit
(tests/test_featstore.py against golden vectors recorded from the reference reader).
"""
import base64
import os
import pickle
import struct
from typing import Dict, Iterable, List, Mapping, Sequence, Tuple, Union

import numpy as np

MAGIC = b"YVFS0001"
F_DIM, P_DIM = 2048, 1601
_ALIGN = 4096
_HEADER = struct.Struct("<8s7Q")       # magic, n, rows, off_index, off_keys, off_features, off_boxes5, off_probs


def decode_item(item: Mapping) -> Tuple[np.ndarray, np.ndarray, np.ndarray, int, int]:
    """One LMDB record (already unpickled) -> (features [K,2048], boxes [K,4], cls_prob [K,1601], image_w, image_h).
    Both conventions of the reference: raw blobs under ``feature``/``bbox`` with ``image_width``/``image_height``,
    or base64 strings under ``features``/``boxes`` with ``image_w``/``image_h`` (features_reader.py:124-150)."""
    old = "image_width" in item
    w = int(item["image_width" if old else "image_w"])
    h = int(item["image_height" if old else "image_h"])

    def blob(k_old, k_new):
        return item[k_old] if old else base64.b64decode(item[k_new])

    features = np.frombuffer(blob("feature", "features"), dtype=np.float32).reshape((-1, F_DIM))
    boxes = np.frombuffer(blob("bbox", "boxes"), dtype=np.float32).reshape((-1, 4))
    cls_prob = np.frombuffer(blob("cls_prob", "cls_prob"), dtype=np.float32).reshape((-1, P_DIM))
    if not (len(features) == len(boxes) == len(cls_prob)):
        raise ValueError("record with inconsistent numbers of regions")
    return features, boxes, cls_prob, w, h


def encode_boxes(boxes: np.ndarray, w: int, h: int) -> np.ndarray:
    """The 5-column region encoding of features_reader.py:84-103 (float32 arithmetic in the same order)."""
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    area /= w * h
    out = np.zeros((len(boxes), 5), dtype=np.float32)
    out[:, 0] = boxes[:, 0] / w
    out[:, 1] = boxes[:, 1] / h
    out[:, 2] = boxes[:, 2] / w
    out[:, 3] = boxes[:, 3] / h
    out[:, 4] = area
    return out


def _pad(fh, align=_ALIGN):
    pos = fh.tell()
    fh.write(b"\0" * ((-pos) % align))
    return fh.tell()


class ShardWriter:
    """Append frames, then ``close()``.  Rows are buffered per section in temporary files next to the target, so a
    shard larger than host memory can be written."""

    def __init__(self, path: Union[str, os.PathLike]):
        self.path = str(path)
        self._keys: List[str] = []
        self._set = set()
        self._offsets: List[int] = [0]
        self._tmp = {name: open(f"{self.path}.{name}.tmp", "wb") for name in ("features", "boxes5", "probs")}
        self._closed = False

    def add(self, key: str, item: Mapping):
        """``item``: an unpickled LMDB record of the reference (either convention)."""
        if key in self._set:
            return                                                # the reference's LMDBWriter ignores duplicates too
        f, b, p, w, h = decode_item(item)
        self._tmp["features"].write(np.ascontiguousarray(f, dtype=np.float32).tobytes())
        self._tmp["boxes5"].write(encode_boxes(b, w, h).tobytes())
        self._tmp["probs"].write(np.ascontiguousarray(p, dtype=np.float32).tobytes())
        self._keys.append(key)
        self._set.add(key)
        self._offsets.append(self._offsets[-1] + len(f))

    def close(self):
        if self._closed:
            return
        self._closed = True
        for fh in self._tmp.values():
            fh.close()
        n, rows = len(self._keys), self._offsets[-1]
        keys_blob = "\n".join(self._keys).encode("utf-8")
        with open(self.path, "wb") as out:
            out.write(b"\0" * _HEADER.size)
            off_index = _pad(out)
            out.write(np.asarray(self._offsets, dtype="<i8").tobytes())
            off_keys = out.tell()
            out.write(struct.pack("<Q", len(keys_blob)) + keys_blob)
            offs = {}
            for name in ("features", "boxes5", "probs"):
                offs[name] = _pad(out)
                with open(f"{self.path}.{name}.tmp", "rb") as src:
                    while True:
                        chunk = src.read(1 << 24)
                        if not chunk:
                            break
                        out.write(chunk)
                os.remove(f"{self.path}.{name}.tmp")
            out.seek(0)
            out.write(_HEADER.pack(MAGIC, n, rows, off_index, off_keys, offs["features"], offs["boxes5"], offs["probs"]))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def convert(records: Iterable[Tuple[Union[str, bytes], Union[bytes, Mapping]]], path) -> int:
    """Write a shard from ``(key, value)`` pairs as an LMDB cursor yields them (value: pickled dict or dict; the
    ``keys`` bookkeeping entry of the reference's LMDB is skipped).  Returns the number of frames written."""
    n = 0
    with ShardWriter(path) as w:
        for key, value in records:
            key = key.decode() if isinstance(key, (bytes, bytearray, memoryview)) else key
            if key == "keys":
                continue
            item = pickle.loads(bytes(value)) if isinstance(value, (bytes, bytearray, memoryview)) else value
            w.add(key, item)
            n += 1
    return n


class ShardReader:
    """Read-only view of one or several shards with the interface of the reference's ``BaseFeaturesReader``:
    ``len(reader)``, ``reader.keys`` and ``reader[(key, ...)] -> (features, locations, probs)``."""

    def __init__(self, path: Union[str, os.PathLike, Sequence[Union[str, os.PathLike]]]):
        paths = [path] if isinstance(path, (str, os.PathLike)) else list(path)
        self._shards = []
        self.keys: Dict[str, Tuple[int, int]] = {}
        for si, p in enumerate(paths):
            with open(p, "rb") as fh:
                magic, n, rows, off_index, off_keys, off_f, off_b, off_p = _HEADER.unpack(fh.read(_HEADER.size))
                if magic != MAGIC:
                    raise RuntimeError(f"{p}: not a yvb200 feature shard")
                fh.seek(off_keys)
                (klen,) = struct.unpack("<Q", fh.read(8))
                keys = fh.read(klen).decode("utf-8").split("\n") if n else []
            offsets = np.memmap(p, dtype="<i8", mode="r", offset=off_index, shape=(n + 1,))
            shard = dict(
                offsets=offsets,
                features=np.memmap(p, dtype=np.float32, mode="r", offset=off_f, shape=(rows, F_DIM)),
                boxes5=np.memmap(p, dtype=np.float32, mode="r", offset=off_b, shape=(rows, 5)),
                probs=np.memmap(p, dtype=np.float32, mode="r", offset=off_p, shape=(rows, P_DIM)))
            self._shards.append(shard)
            for i, k in enumerate(keys):
                self.keys.setdefault(k, (si, i))

    def __len__(self):
        return len(self.keys)

    def rows(self, key: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Zero-copy views (features [K,2048], boxes5 [K,5], probs [K,1601]) of one frame."""
        si, i = self.keys[key]
        sh = self._shards[si]
        lo, hi = int(sh["offsets"][i]), int(sh["offsets"][i + 1])
        return sh["features"][lo:hi], sh["boxes5"][lo:hi], sh["probs"][lo:hi]

    def __getitem__(self, keys: Sequence[str]):
        for key in keys:
            if not isinstance(key, str) or key not in self.keys:
                raise TypeError(f"invalid key: {key}")              # features_reader.py:48-50
        parts = [self.rows(k) for k in keys]
        total = sum(len(f) for f, _, _ in parts)
        if total == 0:
            raise RuntimeError("Features could not be correctly read")
        # one output allocation per array; row 0 is the global entry of features_reader.py:170-180
        features = np.empty((total + 1, F_DIM), dtype=np.float32)
        locations = np.ones((total + 1, 11), dtype=np.float64)
        probs = np.empty((total + 1, P_DIM), dtype=np.float64)
        r = 1
        for f, b, p in parts:
            features[r:r + len(f)] = f
            locations[r:r + len(f), 0:5] = b
            probs[r:r + len(f)] = p
            r += len(f)
        features[0] = features[1:].mean(axis=0, keepdims=True)
        locations[0] = (0, 0, 1, 1, 1, 0, 1, 0, 1, 0, 1)
        probs[0] = np.ones(shape=(1, P_DIM)) / P_DIM
        return features, locations, probs


def assemble_path(reader: "ShardReader", steps: Sequence[Sequence[str]], max_path_length: int, max_num_boxes: int,
                  out=None):
    """One trajectory as the model wants it -- the reference's ``BaseDataset._get_visual_features``
    (utils/dataset/all_dataset.py:294-345) followed by the float32 / int64 conversion of ``__getitem__`` (:236-239).

    ``steps[i]`` is the tuple of frame keys of trajectory step i (several keys = several photos of one step: their
    regions share one global row, features_reader.py:153-182).  Every step is cut / zero-padded to ``max_num_boxes``
    rows, column 11 of the boxes carries the step index, missing steps up to ``max_path_length`` are all padding.
    Returns ``(features f32 [S*B, 2048], boxes f32 [S*B, 12], probs f32 [S*B, 1601], masks i64 [S*B])`` with
    S = max(len(steps), max_path_length) -- bit-identical to the reference, but written once, in float32, straight
    into ``out`` (e.g. slices of a pinned batch buffer) instead of through per-step float64 staging arrays."""
    B = int(max_num_boxes)
    S = max(len(steps), int(max_path_length))
    if out is None:
        out = (np.empty((S * B, F_DIM), np.float32), np.empty((S * B, 12), np.float32),
               np.empty((S * B, P_DIM), np.float32), np.empty((S * B,), np.int64))
    feats, boxes, probs, masks = out
    if feats.shape != (S * B, F_DIM) or boxes.shape != (S * B, 12) or probs.shape != (S * B, P_DIM) or masks.shape != (S * B,):
        raise ValueError("assemble_path: output buffers do not match the trajectory shape")
    feats[:] = 0
    boxes[:] = 0
    probs[:] = 0
    masks[:] = 0
    for i in range(S):
        boxes[i * B:(i + 1) * B, 11] = i
    for i, keys in enumerate(steps):
        for key in keys:
            if not isinstance(key, str) or key not in reader.keys:
                raise TypeError(f"invalid key: {key}")
        parts = [reader.rows(k) for k in keys]
        total = sum(len(f) for f, _, _ in parts)
        if total == 0:
            raise RuntimeError("Features could not be correctly read")
        lo = i * B
        # row 0 of a step: the global entry (mean feature over ALL regions of the step, unit box, uniform probabilities)
        if len(parts) == 1:
            g = parts[0][0].mean(axis=0, keepdims=True)
        else:
            g = np.concatenate([f for f, _, _ in parts], axis=0).mean(axis=0, keepdims=True)
        feats[lo] = g
        boxes[lo, :11] = (0, 0, 1, 1, 1, 0, 1, 0, 1, 0, 1)
        probs[lo] = (np.ones(shape=(1, P_DIM)) / P_DIM)
        n = 1
        for f, b5, p in parts:
            take = min(len(f), B - n)
            if take <= 0:
                break
            feats[lo + n:lo + n + take] = f[:take]
            boxes[lo + n:lo + n + take, 0:5] = b5[:take]
            boxes[lo + n:lo + n + take, 5:11] = 1
            probs[lo + n:lo + n + take] = p[:take]
            n += take
        masks[lo:lo + min(n, B)] = 1
    return feats, boxes, probs, masks
