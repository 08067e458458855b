"""TEST INFRASTRUCTURE ONLY — Python bindings for the two CPU checkers.

  COracle  : oracle/_build/libkge_oracle.so, our plain-C restatement (kge_oracle.c)
  RefLib   : oracle/_ref/Base.so, the reference's own native library compiled from
             /root/reference/base/Base.cpp by oracle/Makefile (process-global state, exactly the
             ctypes call sequence of /root/reference/Config.py:30-51,160-164,347)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  Nothing under openkeonspark_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libkge_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "Base.so")

_vp = C.c_void_p
_i64 = C.c_int64


def build(quiet=True):
    """Compile kge_oracle.c and, when /root/reference is present, the reference Base.so."""
    r = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if not quiet:
        print(r.stdout)


def _p(a):
    return _vp(a.ctypes.data) if a is not None else _vp(0)


def _i64a(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int64))


# ------------------------------------------------------------------------------ dataset files
def read_triples(path):
    with open(path) as f:
        n = int(f.readline())
        arr = np.loadtxt(f, dtype=np.int64, ndmin=2) if n else np.zeros((0, 3), np.int64)
    assert arr.shape[0] == n
    return arr  # columns h, t, r


def read_count(path):
    with open(path) as f:
        return int(f.readline())


def read_lists(path):
    """type_constrain.txt / ontology_constrain.txt -> (keys, offA, flatA, offB, flatB)."""
    with open(path) as f:
        toks = np.array(f.read().split(), dtype=np.int64)
    n = int(toks[0])
    pos = 1
    keys, offa, offb, fa, fb = [], [0], [0], [], []
    for _ in range(n):
        k, c = int(toks[pos]), int(toks[pos + 1])
        fa.append(toks[pos + 2:pos + 2 + c])
        pos += 2 + c
        k2, c2 = int(toks[pos]), int(toks[pos + 1])
        fb.append(toks[pos + 2:pos + 2 + c2])
        pos += 2 + c2
        keys.append(k)
        offa.append(offa[-1] + c)
        offb.append(offb[-1] + c2)
    cat = lambda l: _i64a(np.concatenate(l)) if l else np.zeros(0, np.int64)
    return _i64a(keys), _i64a(offa), cat(fa), _i64a(offb), cat(fb)


# ------------------------------------------------------------------------------ C restatement
class COracle:
    def __init__(self, path=None):
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        self.L = L
        L.orc_new.restype = _vp
        L.orc_new_tail.restype = _i64
        L.orc_new_head.restype = _i64
        L.orc_new_rel.restype = _i64
        L.orc_n_dedup.restype = _i64
        L.orc_tc_eval.restype = C.c_float
        L.orc_find.restype = C.c_int
        self.o = _vp(L.orc_new())
        self.W = 0
        if path is not None:
            self.load(path)

    def load(self, path, test=True):
        j = lambda n: os.path.join(path, n)
        self.E = read_count(j("entity2id.txt"))
        self.R = read_count(j("relation2id.txt"))
        tr = read_triples(j("train2id.txt"))
        nb = read_count(j("batch2id.txt")) if os.path.exists(j("batch2id.txt")) else 0
        self.n_raw = tr.shape[0]
        h, t, r = (_i64a(tr[:, k]) for k in range(3))
        self.L.orc_load_train(self.o, _i64(self.E), _i64(self.R), _p(h), _p(t), _p(r), _i64(self.n_raw), _i64(nb))
        if test:
            te, va = read_triples(j("test2id.txt")), read_triples(j("valid2id.txt"))
            self.n_test, self.n_valid = te.shape[0], va.shape[0]
            a = [_i64a(te[:, k]) for k in range(3)] + [_i64a(va[:, k]) for k in range(3)]
            self.L.orc_load_test(self.o, _p(a[0]), _p(a[1]), _p(a[2]), _i64(self.n_test),
                                 _p(a[3]), _p(a[4]), _p(a[5]), _i64(self.n_valid))
            k, oa, fa, ob, fb = read_lists(j("type_constrain.txt"))
            self.L.orc_load_types(self.o, _i64(k.size), _p(k), _p(oa), _p(fa), _p(ob), _p(fb))
            if os.path.exists(j("ontology_constrain.txt")):
                k, oa, fa, ob, fb = read_lists(j("ontology_constrain.txt"))
                self.L.orc_load_ontology(self.o, _i64(k.size), _p(k), _p(oa), _p(fa), _p(ob), _p(fb))
        return self

    def load_arrays(self, E, R, train, valid=None, test=None, heads=None, tails=None):
        """Same as load() from [n,3] int64 arrays (columns h, t, r) — full-size synthetic graphs that are never written to
        text files.  heads / tails: {relation: sorted ids} type constraints (datagen.type_constraints)."""
        self.E, self.R, self.n_raw = int(E), int(R), int(train.shape[0])
        h, t, r = (_i64a(train[:, k]) for k in range(3))
        self.L.orc_load_train(self.o, _i64(self.E), _i64(self.R), _p(h), _p(t), _p(r), _i64(self.n_raw), _i64(0))
        if test is not None:
            self.n_test, self.n_valid = int(test.shape[0]), int(valid.shape[0])
            a = [_i64a(test[:, k]) for k in range(3)] + [_i64a(valid[:, k]) for k in range(3)]
            self.L.orc_load_test(self.o, _p(a[0]), _p(a[1]), _p(a[2]), _i64(self.n_test),
                                 _p(a[3]), _p(a[4]), _p(a[5]), _i64(self.n_valid))
            keys = sorted(heads)
            offa, offb = np.zeros(len(keys) + 1, np.int64), np.zeros(len(keys) + 1, np.int64)
            for i, k in enumerate(keys):
                offa[i + 1] = offa[i] + heads[k].size
                offb[i + 1] = offb[i] + tails[k].size
            fa = _i64a(np.concatenate([heads[k] for k in keys])) if keys else np.zeros(0, np.int64)
            fb = _i64a(np.concatenate([tails[k] for k in keys])) if keys else np.zeros(0, np.int64)
            ka = _i64a(keys)                     # a named local: _p() of a temporary would hand C a dangling address
            self.L.orc_load_types(self.o, _i64(len(keys)), _p(ka), _p(offa), _p(fa), _p(offb), _p(fb))
        return self

    def set_streams(self, seeds, bern=0):
        s = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
        self.W = s.size
        self.L.orc_set_streams(self.o, _i64(s.size), _p(s), C.c_int(bern))

    def libc_seeds(self, W):
        s = np.zeros(W, np.uint64)
        self.L.orc_libc_seeds(_p(s), _i64(W))
        return s

    def streams(self):
        s = np.zeros(self.W, np.uint64)
        self.L.orc_get_streams(self.o, _p(s))
        return s

    def sampling(self, B, k=1, kr=0):
        S = B * (1 + k + kr)
        h, t, r = (np.zeros(S, np.int64) for _ in range(3))
        y = np.zeros(S, np.float32)
        self.L.orc_sampling(self.o, _p(h), _p(t), _p(r), _p(y), _i64(B), _i64(k), _i64(kr))
        return h, t, r, y

    def find(self, h, t, r):
        return bool(self.L.orc_find(self.o, _i64(h), _i64(t), _i64(r)))

    def rank(self, side, index, scores):
        s = np.ascontiguousarray(scores, dtype=np.float32)
        out = np.zeros(8, np.int64)
        self.L.orc_rank(self.o, C.c_int(side), _i64(index), _p(s), _p(out))
        return out

    def get_list(self, which):
        n = self.n_valid if which else self.n_test
        h, t, r = (np.zeros(n, np.int64) for _ in range(3))
        self.L.orc_get_list(self.o, C.c_int(which), _p(h), _p(t), _p(r))
        return h, t, r

    def tc_batch(self, which):
        n = self.n_valid if which else self.n_test
        a = [np.zeros(n, np.int64) for _ in range(6)]
        self.L.orc_tc_batch(self.o, C.c_int(which), *[_p(x) for x in a])
        return a

    def best_threshold(self, pos, neg, thresh=None):
        th = np.zeros(self.R, np.float32) if thresh is None else thresh
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1)
        neg = np.ascontiguousarray(neg, np.float32).reshape(-1)
        self.L.orc_best_threshold(self.o, _p(th), _p(pos), _p(neg))
        return th

    def tc_eval(self, thresh, pos, neg):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1)
        neg = np.ascontiguousarray(neg, np.float32).reshape(-1)
        cnt = np.zeros(4, np.int64)
        acc = self.L.orc_tc_eval(self.o, _p(thresh), _p(pos), _p(neg), _p(cnt))
        return float(acc), cnt

    def means(self):
        a, b = np.zeros(self.R, np.float32), np.zeros(self.R, np.float32)
        self.L.orc_means(self.o, _p(a), _p(b))
        return a, b

    def predict(self, model, params, h, t, r):
        """Canonical-order fp32 scores.  model in {"TransE","TransH","TransR","TransD"}."""
        mid = {"TransE": 0, "TransH": 1, "TransR": 2, "TransD": 3}[model]
        f = lambda k: np.ascontiguousarray(params[k], np.float32) if k in params else None
        ent, rel = f("ent_embeddings"), f("rel_embeddings")
        aux_ent = f("ent_transfer")
        aux_rel = f({1: "normal_vectors", 2: "transfer_matrix", 3: "rel_transfer"}.get(mid, "-"))
        h, t, r = _i64a(h), _i64a(t), _i64a(r)
        out = np.zeros(h.size, np.float32)
        self.L.orc_predict(C.c_int(mid), C.c_int(ent.shape[1]), C.c_int(rel.shape[1]), _p(ent), _p(rel),
                           _p(aux_ent), _p(aux_rel), _p(h), _p(t), _p(r), _i64(h.size), _p(out))
        return out


# ------------------------------------------------------------------------------ the reference itself
class RefLib:
    """The reference's Base.so.  State is process-global: one dataset per process."""

    def __init__(self, so=REF_SO):
        if not os.path.exists(so):
            build()
        if not os.path.exists(so):
            raise FileNotFoundError(so + " (needs /root/reference to build)")
        L = C.CDLL(so)
        self.L = L
        L.sampling.argtypes = [_vp, _vp, _vp, _vp, _i64, _i64, _i64]
        for n in ("getTailBatch", "getHeadBatch"):
            getattr(L, n).argtypes = [_i64, _vp, _vp, _vp]
        for n in ("testTail", "testHead"):
            getattr(L, n).argtypes = [_i64, _vp]
            getattr(L, n).restype = C.POINTER(_i64 * 8)
        L.getTestBatch.argtypes = [_vp] * 6
        L.getValidBatch.argtypes = [_vp] * 6
        L.getBestThreshold.argtypes = [_vp] * 3
        L.test_triple_classification.argtypes = [_vp] * 4     # Config.py:45 lists 3; the C function takes 4
        for n in ("getEntityTotal", "getRelationTotal", "getTrainTotal", "getTrainTotal_", "getTestTotal",
                  "getValidTotal", "getTripleTotal", "getBatchTotal", "getWorkThreads"):
            getattr(L, n).restype = _i64
        L.setWorkThreads.argtypes = [_i64]
        L.setBern.argtypes = [_i64]

    def init(self, path, bern=0, W=8, test=True, ontology=True):
        """Config.init() call sequence (Config.py:160-164) + init_link_prediction (:78-80)."""
        if not path.endswith("/"):
            path += "/"
        self.L.setInPath(C.create_string_buffer(path.encode(), len(path) * 2))
        self.L.setBern(bern)
        self.L.setWorkThreads(W)
        self.L.randReset()
        self.L.importTrainFiles()
        self.W = W
        self.E = self.L.getEntityTotal()
        self.R = self.L.getRelationTotal()
        if test:
            self.L.importTestFiles()
            self.L.importTypeFiles()
            if ontology:
                self.L.importOntologyFiles()
        return self

    def seeds(self):
        """Current per-thread LCG states (global `next_random`, Random.h:6)."""
        ptr = C.c_void_p.in_dll(self.L, "next_random").value
        return np.ctypeslib.as_array((C.c_uint64 * self.W).from_address(ptr)).copy()

    def sampling(self, B, k=1, kr=0):
        S = B * (1 + k + kr)
        h, t, r = (np.zeros(S, np.int64) for _ in range(3))
        y = np.zeros(S, np.float32)
        self.L.sampling(_p(h), _p(t), _p(r), _p(y), B, k, kr)
        return h, t, r, y

    def rank(self, side, index, scores):
        s = np.ascontiguousarray(scores, dtype=np.float32)
        fn = self.L.testTail if side else self.L.testHead
        return np.array(list(fn(index, _p(s)).contents), dtype=np.int64)

    def candidates(self, side, index):
        a = [np.zeros(self.E, np.int64) for _ in range(3)]
        (self.L.getTailBatch if side else self.L.getHeadBatch)(index, *[_p(x) for x in a])
        return a

    def tc_batch(self, which):
        n = self.L.getValidTotal() if which else self.L.getTestTotal()
        a = [np.zeros(n, np.int64) for _ in range(6)]
        (self.L.getValidBatch if which else self.L.getTestBatch)(*[_p(x) for x in a])
        return a

    def best_threshold(self, pos, neg, thresh=None):
        th = np.zeros(self.R, np.float32) if thresh is None else thresh
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1)
        neg = np.ascontiguousarray(neg, np.float32).reshape(-1)
        self.L.getBestThreshold(_p(th), _p(pos), _p(neg))
        return th

    def tc_eval(self, thresh, pos, neg):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1)
        neg = np.ascontiguousarray(neg, np.float32).reshape(-1)
        acc = np.zeros(1, np.float32)
        self.L.test_triple_classification(_p(thresh), _p(pos), _p(neg), _p(acc))
        return float(acc[0])
