"""BASELINE.json's full sizes (FB15K-shaped graph: 14,951 entities / 1,345 relations / 483,142 train triples, B = 4,831).

Where the C oracle finishes in seconds the comparison is direct and bit-exact (one sampled batch, a sample of ranking
queries); the rest are size-independent properties of the domain: filtered negatives are never known triples,
determinism, chunked == step-by-step, the loss equals the hinge of the canonical predict scores, SGD leaves untouched rows
alone and is linear in the learning rate, filtered <= raw and constrained <= unconstrained ranks, candidate shards sum
to the whole, and a checksum of the better-than counts against a torch recomputation."""
import ctypes

import numpy as np
import pytest

from conftest import make_params

pytestmark = pytest.mark.gpu
vp = ctypes.c_void_p


@pytest.fixture(scope="module")
def fb15k(tmp_path_factory):
    from openkeonspark_b200 import datagen
    d = str(tmp_path_factory.mktemp("fb15k")) + "/"
    g = datagen.make_shape("fb15k", seed=0)
    datagen.write_dataset(g, d, ontology=False)
    return d, g


def _con(path, model, D, opt="SGD", k=1, lr=0.01, lp=False):
    import openkeonspark_b200 as okb
    con = okb.Config(private_context=True)
    con.set_in_path(path)
    con.set_nbatches(100); con.set_ent_neg_rate(k); con.set_alpha(lr); con.set_opt_method(opt); con.set_dimension(D)
    con.set_test_link_prediction(lp); con.set_test_head(1)
    con.workThreads = 8
    con.init()
    con.set_model_and_session(getattr(okb, model))
    con.set_parameters(make_params(model, con.entTotal, con.relTotal, D, seed=11))
    seeds = np.arange(1, 9, dtype=np.uint64) * np.uint64(2654435761)
    con.ctx.call("okb_set_streams", vp(seeds.ctypes.data), 8)
    return con, seeds


def test_fullsize_sampler_bit_exact_and_filtered(built, fb15k):
    from oracle.harness import COracle
    path, g = fb15k
    con, seeds = _con(path, "TransE", 50, k=2)
    assert con.batch_size == 4831 and con.entTotal == 14951 and con.trainTotal == 483142
    orc = COracle().load(path, test=False)
    orc.set_streams(seeds, 0)
    known = set(map(tuple, g.train.tolist()))              # columns h, t, r
    for it in range(2):
        con.sampling()
        eh, et, er, ey = orc.sampling(4831, 2, 0)
        assert np.array_equal(con.batch_h, eh) and np.array_equal(con.batch_t, et) and np.array_equal(con.batch_r, er)
        assert np.array_equal(con.batch_y, ey)
        B = con.batch_size
        h, t, r = con.batch_h, con.batch_t, con.batch_r
        for b in range(0, B, 7):
            assert (h[b], t[b], r[b]) in known                                   # positives are train rows
            for m in (1, 2):
                nh, nt, nr = h[b + m * B], t[b + m * B], r[b + m * B]
                assert nr == r[b] and (nh != h[b]) != (nt != t[b])                 # exactly one side replaced
                assert (nh, nt, nr) not in known                                   # Corrupt.h:7-69: never a known triple


def test_fullsize_train_properties(built, fb15k):
    import torch
    path, g = fb15k
    # determinism and chunked == step-by-step (Adam, TransH D=100: the bench workload)
    runs = []
    for mode in ("steps", "steps", "chunk"):
        con, _ = _con(path, "TransH", 100, opt="Adam", lr=0.001)
        if mode == "chunk":
            losses = con.train_chunk_device(6).cpu().numpy()
        else:
            con.plan_ahead = 6
            losses = np.array([float(con.next_step_device().item()) for _ in range(6)], np.float32)
        runs.append((losses, con.get_parameters()))
    for a, b in ((0, 1), (0, 2)):
        assert np.array_equal(runs[a][0], runs[b][0])
        for k in runs[a][1]:
            assert np.array_equal(runs[a][1][k], runs[b][1][k]), k
    # the loss is the mean hinge of the canonical predict scores of the same batch (cross-kernel consistency)
    con, _ = _con(path, "TransE", 50, k=1)
    con.sampling()
    h, t, r = con.batch_h.copy(), con.batch_t.copy(), con.batch_r.copy()
    s = con.test_step(h, t, r).reshape(-1).astype(np.float64) * 50          # TransE predict is the mean over d
    B = con.batch_size
    want = np.maximum(s[:B] - s[B:] + 1.0, 0.0).mean()
    x0 = con.get_parameters()
    got = float(con.train_step_device(0).item())
    assert abs(got - want) <= 2e-5 * want, (got, want)
    # SGD: rows outside the batch are untouched; the update is linear in the learning rate
    x1 = con.get_parameters()
    touched = np.zeros(con.entTotal, bool); touched[h] = True; touched[t] = True
    assert np.array_equal(x1["ent_embeddings"][~touched], x0["ent_embeddings"][~touched]) and (~touched).sum() > 1000
    assert np.abs(x1["ent_embeddings"][touched] - x0["ent_embeddings"][touched]).max() > 0
    con2, _ = _con(path, "TransE", 50, k=1, lr=0.02)
    con2.sampling()
    assert np.array_equal(con2.batch_h, h)
    con2.train_step_device(0)
    x2 = con2.get_parameters()
    for k in x0:
        d1, d2 = (x1[k] - x0[k]).astype(np.float64), (x2[k] - x0[k]).astype(np.float64)
        assert np.abs(d2 - 2 * d1).max() <= 1e-6, k
    del torch


def test_fullsize_ranking(built, fb15k):
    import torch
    from oracle.harness import COracle
    path, g = fb15k
    con, _ = _con(path, "TransH", 100, lp=True)
    n = 2048
    rec = con.link_prediction_records(0, n).cpu().numpy()                  # [n, 2, 8]
    raw, flt, rawc, fltc = rec[..., 0], rec[..., 1], rec[..., 2], rec[..., 3]
    assert (flt <= raw).all() and (rawc <= raw).all() and (fltc <= rawc).all() and (fltc <= flt).all()
    assert (raw >= 0).all() and (raw < con.entTotal).all() and raw.mean() > 100        # untrained tables: far from rank 0
    # candidate shards sum to the whole (8 contiguous ranges, the 8-GPU evaluation layout)
    dev = con.trainModel.device
    counts = torch.zeros(n * 8, dtype=torch.int64, device=dev)
    best = torch.full((n * 8,), -1, dtype=torch.int64, device=dev)
    m = con._cmodel()
    E = con.entTotal
    for gq in range(8):
        con.ctx.call("okb_rank", ctypes.byref(m), 0, n, 1, E * gq // 8, E * (gq + 1) // 8, vp(counts.data_ptr()), vp(best.data_ptr()), None)
    out = torch.empty(n * 16, dtype=torch.int64, device=dev)
    con.ctx.call("okb_rank_finalize", 0, n, vp(counts.data_ptr()), vp(best.data_ptr()), vp(out.data_ptr()), None)
    assert np.array_equal(out.view(n, 2, 8).cpu().numpy(), rec)
    # direct comparison with the oracle for a sample of queries (canonical scores -> testHead/testTail)
    orc = COracle(path)
    P = con.get_parameters()
    th, tt, tr = orc.get_list(0)
    ents = np.arange(orc.E)
    checksum_gpu, checksum_ref = 0, 0
    for i in range(0, n, 97):
        for side in (0, 1):
            s = (orc.predict("TransH", P, np.full(orc.E, th[i]), ents, np.full(orc.E, tr[i])) if side else
                 orc.predict("TransH", P, ents, np.full(orc.E, tt[i]), np.full(orc.E, tr[i])))
            exp = orc.rank(side, i, s)
            assert np.array_equal(rec[i, side], exp), (i, side, rec[i, side], exp)
            # checksum of the raw better-than counts against a recomputation from the GPU's own predict kernel
            hh = np.full(E, th[i]) if side else ents
            ttt = ents if side else np.full(E, tt[i])
            sg = con.test_step(hh, ttt, np.full(E, tr[i])).reshape(-1)
            tgt = tt[i] if side else th[i]
            checksum_ref += int((sg < sg[tgt]).sum())
            checksum_gpu += int(rec[i, side, 0])
    assert checksum_gpu == checksum_ref
