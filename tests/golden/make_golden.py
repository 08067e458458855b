"""Generates tests/golden/*.npz by RUNNING THE REFERENCE ITSELF: oracle/_ref/Base.so, compiled from
/root/reference/base/Base.cpp (oracle/Makefile), driven with the ctypes call sequence of
/root/reference/Config.py:160-164,347 and distribute_training.py:467-475.

The reference ships no tests or golden vectors of its own, so these are the
pin for the integer half of the oracle.  Run from the repo root (needs /root/reference):

    python tests/golden/make_golden.py

Each case runs in a fresh process so that libc rand() — the seed source of Random.h:12 — starts
from its default state (seeds 1804289383, 846930886, ...).  The dataset is regenerated from
openkeonspark_b200.datagen with fixed seeds; its arrays are stored in the .npz too so the fixture
is self-contained.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (shape, datagen kwargs, bern, W, [(B, k, kr) ...])
    "tiny_uniform": ("tiny", dict(seed=3, dup_train=20), 0, 8, [(50, 1, 0), (13, 2, 1), (7, 10, 0)]),
    "small_zipf_bern": ("small", dict(seed=1, zipf=True, dup_train=50), 1, 8, [(600, 1, 0), (101, 3, 2), (4831, 1, 0)]),
    "small_w3": ("small", dict(seed=2), 1, 3, [(100, 2, 0), (2, 1, 1)]),
}


def child(name):
    from openkeonspark_b200 import datagen
    from oracle.harness import RefLib
    shape, kw, bern, W, calls = CASES[name]
    d = tempfile.mkdtemp() + "/"
    g = datagen.make_shape(shape, **kw)
    datagen.write_dataset(g, d, ontology=True)
    ref = RefLib().init(d, bern=bern, W=W)
    txt = lambda n: np.frombuffer(open(d + n, "rb").read(), dtype=np.uint8)
    out = {"type_constrain_txt": txt("type_constrain.txt"), "ontology_constrain_txt": txt("ontology_constrain.txt"),
           "train": g.train, "valid": g.valid, "test": g.test, "E": g.E, "R": g.R, "bern": bern, "W": W,
           "seeds": ref.seeds(), "calls": np.array(calls)}
    for i, (B, k, kr) in enumerate(calls):
        for rep in range(2):
            h, t, r, y = ref.sampling(B, k, kr)
            out["s%d_%d_h" % (i, rep)] = h.astype(np.int32)
            out["s%d_%d_t" % (i, rep)] = t.astype(np.int32)
            out["s%d_%d_r" % (i, rep)] = r.astype(np.int32)
    out["seeds_after"] = ref.seeds()
    # ranking records for pseudo-random score vectors (ties included)
    rng = np.random.default_rng(99)
    n_test = ref.L.getTestTotal()
    idx = np.arange(0, n_test, max(1, n_test // 12))
    scores = rng.standard_normal((idx.size, g.E)).astype(np.float32)
    scores[::2] = np.round(scores[::2] * 2) / 2
    rec = np.zeros((idx.size, 2, 8), np.int64)
    for a, i in enumerate(idx):
        rec[a, 0] = ref.rank(0, int(i), scores[a])
        rec[a, 1] = ref.rank(1, int(i), scores[a])
    out["rank_idx"], out["rank_scores"], out["rank_rec"] = idx, scores, rec
    # triple classification: negatives (libc rand stream continues), thresholds, accuracy
    vb = ref.tc_batch(1)
    tb = ref.tc_batch(0)
    out["valid_batch"], out["test_batch"] = np.stack(vb), np.stack(tb)
    sp, sn = rng.random(len(vb[0])).astype(np.float32) * 3, rng.random(len(vb[0])).astype(np.float32) * 3 + 0.5
    th = ref.best_threshold(sp, sn)
    tp, tn = rng.random(len(tb[0])).astype(np.float32) * 3, rng.random(len(tb[0])).astype(np.float32) * 3 + 0.5
    out["tc_vp"], out["tc_vn"], out["tc_tp"], out["tc_tn"], out["tc_thresh"] = sp, sn, tp, tn, th
    out["tc_acc"] = np.float32(ref.tc_eval(th, tp, tn))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
    else:
        for name in CASES:
            subprocess.run([sys.executable, __file__, name], check=True, stdout=subprocess.DEVNULL)
            print("wrote", name)
