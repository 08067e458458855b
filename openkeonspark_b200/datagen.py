"""Synthetic knowledge-graph datasets in the reference's on-disk format.

There are no datasets offline, so benchmarks and parity tests run on seed-fixed synthetic
graphs with the shapes BASELINE.json names.  The file format is the one the reference's
reader consumes (/root/reference/base/Reader.h:27-179, 186-292, 302-365, 376-449):

  entity2id.txt / relation2id.txt   first line = count (only the count is read)
  train2id.txt / valid2id.txt / test2id.txt
                                    first line = count, then one "h t r" line per triple
                                    (head, TAIL, relation order: Reader.h:92-94)
  type_constrain.txt                first line = #relations, then per relation two lines
                                    "r n h1 .. hn" (head types) and "r n t1 .. tn" (tail types),
                                    built the way main_spark.py:209-290 (n_n) builds it
  ontology_constrain.txt (optional) first line = #entities listed, then per entity two lines
                                    "e n sup1 .. supn" and "e n sub1 .. subn"
  batch2id.txt (optional)           first line = size of the incremental batch (Reader.h:61-67)
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

# Shapes named by BASELINE.json / SURVEY.md section 8(d).
SHAPES = {
    "fb15k": dict(E=14951, R=1345, n_train=483142, n_valid=50000, n_test=59071),
    "wn18": dict(E=40943, R=18, n_train=141442, n_valid=5000, n_test=5000),
    "dbpedia": dict(E=4_000_000, R=600, n_train=20_000_000, n_valid=100_000, n_test=100_000),
    # small shapes for CPU-side tests (oracle finishes in well under a second)
    "tiny": dict(E=60, R=7, n_train=400, n_valid=40, n_test=50),
    "small": dict(E=500, R=23, n_train=6000, n_valid=300, n_test=400),
    # many relations / entities relative to the batch: gradient-row segments stay short (no hub pre-reduction)
    "wide": dict(E=4000, R=800, n_train=8000, n_valid=200, n_test=200),
}


@dataclass
class Graph:
    E: int
    R: int
    train: np.ndarray  # [n,3] int64 columns h,t,r (file order)
    valid: np.ndarray
    test: np.ndarray


def _draw(rng, n, E, R, zipf):
    if zipf:
        # hub entities / frequent relations: drives duplicate-row reduction and (h,r,*) run length
        pe = 1.0 / np.arange(1, E + 1) ** 1.0
        pr = 1.0 / np.arange(1, R + 1) ** 1.2
        perm_e = rng.permutation(E)
        perm_r = rng.permutation(R)
        h = perm_e[rng.choice(E, size=n, p=pe / pe.sum())]
        t = perm_e[rng.choice(E, size=n, p=pe / pe.sum())]
        r = perm_r[rng.choice(R, size=n, p=pr / pr.sum())]
    else:
        h = rng.integers(0, E, size=n)
        t = rng.integers(0, E, size=n)
        r = rng.integers(0, R, size=n)
    return np.stack([h, t, r], axis=1).astype(np.int64)


def make_graph(E, R, n_train, n_valid, n_test, seed=0, zipf=False, dup_train=0):
    """Draw n_train+n_valid+n_test distinct (h,t,r) triples and split them.

    dup_train > 0 appends that many repeated train rows (the reference keeps duplicates in
    trainList_no and samples from them, Reader.h:81-100, while the sorted copies are deduped).
    """
    rng = np.random.default_rng(seed)
    need = n_train + n_valid + n_test
    cap = E * E * R
    if need > cap:
        raise ValueError("more distinct triples requested than exist")
    got = np.zeros((0, 3), dtype=np.int64)
    while got.shape[0] < need:
        cand = _draw(rng, int((need - got.shape[0]) * 1.3) + 16, E, R, zipf)
        got = np.concatenate([got, cand], axis=0)
        key = (got[:, 0] * E + got[:, 1]) * R + got[:, 2]
        _, first = np.unique(key, return_index=True)
        got = got[np.sort(first)]
    got = got[:need]
    got = got[rng.permutation(need)]
    train = got[:n_train]
    if dup_train:
        train = np.concatenate([train, train[rng.integers(0, n_train, size=dup_train)]], axis=0)
        train = train[rng.permutation(train.shape[0])]
    return Graph(E, R, train, got[n_train:n_train + n_valid], got[n_train + n_valid:])


def make_shape(name, seed=0, zipf=False, **kw):
    s = dict(SHAPES[name])
    s.update(kw)
    return make_graph(seed=seed, zipf=zipf, **s)


def _write_triples(path, arr):
    with open(path, "w") as f:
        f.write("%d\n" % arr.shape[0])
        if arr.shape[0]:
            np.savetxt(f, arr, fmt="%d", delimiter=" ")


def type_constraints(g: Graph):
    """Per relation: set of heads and set of tails seen in train+valid+test (main_spark.py:209-290)."""
    allt = np.concatenate([g.train, g.valid, g.test], axis=0)
    heads, tails = {}, {}
    order = np.argsort(allt[:, 2], kind="stable")
    s = allt[order]
    bounds = np.flatnonzero(np.diff(s[:, 2])) + 1
    for blk in np.split(s, bounds):
        if blk.shape[0] == 0:
            continue
        r = int(blk[0, 2])
        heads[r] = np.unique(blk[:, 0])
        tails[r] = np.unique(blk[:, 1])
    return heads, tails


def write_dataset(g: Graph, path, ontology=False, seed=0, new_batch=0):
    """Write the reference's text files for graph `g` into directory `path`."""
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "entity2id.txt"), "w") as f:
        f.write("%d\n" % g.E)
    with open(os.path.join(path, "relation2id.txt"), "w") as f:
        f.write("%d\n" % g.R)
    _write_triples(os.path.join(path, "train2id.txt"), g.train)
    _write_triples(os.path.join(path, "valid2id.txt"), g.valid)
    _write_triples(os.path.join(path, "test2id.txt"), g.test)
    heads, tails = type_constraints(g)
    rng = np.random.default_rng(seed + 12345)
    with open(os.path.join(path, "type_constrain.txt"), "w") as f:
        f.write("%d\n" % len(heads))
        for r in heads:
            # the file lists ids unsorted (dict order in the reference); the reader sorts them
            hs = rng.permutation(heads[r])
            ts = rng.permutation(tails[r])
            f.write("%d\t%d" % (r, hs.size) + "".join("\t%d" % x for x in hs) + "\n")
            f.write("%d\t%d" % (r, ts.size) + "".join("\t%d" % x for x in ts) + "\n")
    if ontology:
        n_ont = min(g.E, max(4, g.E // 3))
        ents = rng.choice(g.E, size=n_ont, replace=False)
        with open(os.path.join(path, "ontology_constrain.txt"), "w") as f:
            f.write("%d\n" % n_ont)
            for e in ents:
                sup = rng.choice(g.E, size=int(rng.integers(0, 5)), replace=False)
                sub = rng.choice(g.E, size=int(rng.integers(0, 5)), replace=False)
                f.write("%d\t%d" % (e, sup.size) + "".join("\t%d" % x for x in sup) + "\n")
                f.write("%d\t%d" % (e, sub.size) + "".join("\t%d" % x for x in sub) + "\n")
    if new_batch:
        with open(os.path.join(path, "batch2id.txt"), "w") as f:
            f.write("%d\n" % new_batch)
    return path


def xavier_normal(rng, rows, cols):
    """fp32 N(0, sigma^2) with sigma = sqrt(2/(rows+cols)) — the scale of
    tf.contrib.layers.xavier_initializer(uniform=False) (TransE.py:21-22).  TF's RNG stream is
    not reproducible, so parity tests inject parameters through Config.set_parameters."""
    sigma = np.sqrt(2.0 / (rows + cols))
    return (rng.standard_normal((rows, cols)) * sigma).astype(np.float32)
