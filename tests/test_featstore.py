"""Feature-store path (SURVEY.md 8f "next" #4), host only: the numpy oracle and the shard reader against golden vectors
recorded from the reference's own YTbFeaturesReader (both record conventions), bit-exact, plus format round trips."""
import os
import pickle
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
import featstore_oracle as FO  # noqa: E402
from yvb200 import featstore as FS  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "featstore.npz"), allow_pickle=True)
STORE = {str(k): bytes(v.tobytes()) for k, v in zip(G["keys"], G["records"])}
QUERIES = ["q0", "q1", "q2"]


def _same(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.parametrize("q", QUERIES)
def test_oracle_matches_reference_reader(q):
    f, l, p = FO.read(STORE, [str(k) for k in G[f"{q}/query"]])
    assert _same(f, G[f"{q}/features"]) and _same(l, G[f"{q}/locations"]) and _same(p, G[f"{q}/probs"])


@pytest.fixture(scope="module")
def shard(tmp_path_factory):
    path = tmp_path_factory.mktemp("fs") / "frames.yvfs"
    # what an LMDB cursor yields, including the reference's bookkeeping entry
    records = [(b"keys", pickle.dumps([k.encode() for k in STORE]))] + [(k.encode(), v) for k, v in STORE.items()]
    assert FS.convert(records, path) == len(STORE)
    return path


@pytest.mark.parametrize("q", QUERIES)
def test_shard_reader_matches_reference_reader(shard, q):
    r = FS.ShardReader(shard)
    assert len(r) == len(STORE) and set(r.keys) == set(STORE)
    f, l, p = r[tuple(str(k) for k in G[f"{q}/query"])]
    assert _same(f, G[f"{q}/features"]) and _same(l, G[f"{q}/locations"]) and _same(p, G[f"{q}/probs"])


def test_shard_layout_and_errors(shard, tmp_path):
    r = FS.ShardReader(shard)
    with pytest.raises(TypeError):
        r[("nope/0",)]
    with pytest.raises(TypeError):
        r[(3,)]
    # sections are page aligned and hold exactly the reference's float32 bytes
    raw = open(shard, "rb").read()
    magic, n, rows, off_i, off_k, off_f, off_b, off_p = FS._HEADER.unpack(raw[:FS._HEADER.size])
    assert magic == FS.MAGIC and n == len(STORE) and off_f % 4096 == 0 and off_b % 4096 == 0 and off_p % 4096 == 0
    for key in STORE:
        f0, b0, p0, w, h = FO.decode_item(pickle.loads(STORE[key]))
        f, b5, p = r.rows(key)
        assert np.array_equal(f, f0) and np.array_equal(p, p0) and np.array_equal(b5, FO.encode_boxes(b0, w, h))
    assert rows == sum(len(FO.decode_item(pickle.loads(v))[0]) for v in STORE.values())
    # several shards behave like several LMDBs (first occurrence of a key wins); an empty shard is legal
    empty = tmp_path / "empty.yvfs"
    FS.ShardWriter(empty).close()
    r2 = FS.ShardReader([empty, shard, shard])
    assert len(r2) == len(STORE)
    f, _, _ = r2[(str(G["q1/query"][0]),)]
    assert _same(f, G["q1/features"])
    with pytest.raises(RuntimeError):
        FS.ShardReader(__file__)


TRAJ = ["t0", "t1", "t2"]


def _steps(name):
    return [tuple(str(x).split("|")) for x in G[f"{name}/steps"]]


@pytest.mark.parametrize("t", TRAJ)
def test_oracle_trajectory_matches_reference_dataset(t):
    f, b, p, m = FO.visual_features(STORE, _steps(t), 4, 4)
    assert _same(f, G[f"{t}/features"]) and _same(b, G[f"{t}/boxes"]) and _same(p, G[f"{t}/probs"]) and _same(m, G[f"{t}/masks"])


@pytest.mark.parametrize("t", TRAJ)
def test_assemble_path_matches_reference_dataset(shard, t):
    r = FS.ShardReader(shard)
    f, b, p, m = FS.assemble_path(r, _steps(t), 4, 4)
    assert _same(f, G[f"{t}/features"]) and _same(b, G[f"{t}/boxes"]) and _same(p, G[f"{t}/probs"]) and _same(m, G[f"{t}/masks"])
    # into caller-provided (e.g. pinned) buffers that hold garbage
    out = (np.full_like(f, 7), np.full_like(b, 7), np.full_like(p, 7), np.full_like(m, 7))
    FS.assemble_path(r, _steps(t), 4, 4, out=out)
    assert all(_same(x, y) for x, y in zip(out, (f, b, p, m)))
    with pytest.raises(TypeError):
        FS.assemble_path(r, [("missing/1",)], 4, 4)
    with pytest.raises(ValueError):
        FS.assemble_path(r, _steps(t), 5, 4, out=out)
