"""Developer tool (host only): trajectories/s of the feature-store read path -- the reference's per-frame decode
(pickle + base64 + box encoding + concatenation, restated in oracle/featstore_oracle.py) against yvb200.featstore
shards, on synthetic frames of the cfg2 shape (8 frames x 36 regions per trajectory)."""
import base64, os, pickle, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import featstore_oracle as FO
from yvb200 import featstore as FS

rng = np.random.RandomState(0)
frames, regions, n_frames = 8, 36, 400
store = {}
for i in range(n_frames):
    f = rng.randn(regions, 2048).astype(np.float32)
    b = np.abs(rng.randn(regions, 4)).astype(np.float32) * 100
    p = rng.dirichlet(np.ones(1601) * 0.1, regions).astype(np.float32)
    store[f"vid/{i:06d}"] = pickle.dumps({"image_w": "640", "image_h": "360", "features": base64.b64encode(f.tobytes()),
                                          "boxes": base64.b64encode(b.tobytes()), "cls_prob": base64.b64encode(p.tobytes())})
keys = list(store)
trajs = [tuple(keys[j] for j in rng.choice(n_frames, frames, replace=False)) for _ in range(200)]
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "s.yvfs")
    t0 = time.perf_counter(); FS.convert(store.items(), path); t_conv = time.perf_counter() - t0
    reader = FS.ShardReader(path)
    for name, fn in (("reference decode (oracle port)", lambda q: FO.read(store, q)), ("shard reader", lambda q: reader[q])):
        fn(trajs[0])
        t0 = time.perf_counter()
        for q in trajs:
            out = fn(q)
        dt = time.perf_counter() - t0
        mb = sum(a.nbytes for a in out) / 1e6
        print(f"{name:32s}: {len(trajs)/dt:8.1f} trajectories/s  ({dt/len(trajs)*1e3:.2f} ms each, {mb:.1f} MB out)")
    a, b = FO.read(store, trajs[3]), reader[trajs[3]]
    print("bit-exact:", all(np.array_equal(x, y) and x.dtype == y.dtype for x, y in zip(a, b)),
          f"| shard {os.path.getsize(path)/1e6:.0f} MB vs pickled records {sum(len(v) for v in store.values())/1e6:.0f} MB, convert {t_conv:.2f} s")
