"""The rest of the reference's user-facing surface on the hot path, against the CPU checkers:

  predict_head_entity / predict_tail_entity / predict_relation / predict_triple   /root/reference/Config.py:574-663
  plot_roc + get_n_interval / get_TPFP                                            Config.py:519-571, base/Test.h:390-444
  getHeadBatch / getTailBatch                                                     base/Test.h:10-26
  ids outside the tables                                                          TF embedding_lookup raises InvalidArgument

Integer results (top-k ids, TP/FP counts, candidate fills) are bit-exact: top-k is an argsort of scores that are
bit-identical to the canonical-order oracle, the counts are fp32 compares on identical score arrays."""
import ctypes
import os

import numpy as np
import pytest

from conftest import make_params

pytestmark = pytest.mark.gpu
vp = ctypes.c_void_p


def _config(path, model, D, tc=True, Dr=None):
    import openkeonspark_b200 as okb
    con = okb.Config(private_context=True)
    con.set_in_path(path)
    con.set_nbatches(4)
    con.set_dimension(D)
    if Dr is not None:
        con.set_rel_dimension(Dr)
    con.set_test_triple_classification(tc)
    con.init()
    con.set_model_and_session(getattr(okb, model))
    return con


@pytest.mark.parametrize("model,D,Dr", [("TransE", 50, None), ("TransH", 100, None), ("TransD", 33, None), ("TransR", 40, 24)])
def test_predict_topk_matches_oracle(built, small_ds, model, D, Dr, capsys):
    from oracle.harness import COracle
    con = _config(small_ds, model, D, Dr=Dr)
    P = make_params(model, con.entTotal, con.relTotal, D, seed=21, Dr=Dr)
    con.set_parameters(P)
    orc = COracle(small_ds)
    E, R = orc.E, orc.R
    ents, rels = np.arange(E), np.arange(R)
    for (a, r, k) in ((3, 5, 10), (E - 1, 0, 1), (17, R - 1, E)):
        got = con.predict_head_entity(a, r, k)                       # Config.py:574-593: (t, r, k)
        exp = orc.predict(model, P, ents, np.full(E, a), np.full(E, r)).argsort()[:k]
        assert np.array_equal(got, exp), (model, "head", a, r, k)
        got = con.predict_tail_entity(a, r, k)                       # Config.py:595-614: (h, r, k)
        exp = orc.predict(model, P, np.full(E, a), ents, np.full(E, r)).argsort()[:k]
        assert np.array_equal(got, exp), (model, "tail", a, r, k)
    for (h, t, k) in ((3, 9, 5), (0, E - 1, R)):
        got = con.predict_relation(h, t, k)                          # Config.py:616-635; TransR: every row through M_{r[0]} = M_0
        exp = orc.predict(model, P, np.full(R, h), np.full(R, t), rels).argsort()[:k]
        assert np.array_equal(got, exp), (model, "relation", h, t, k)
    assert "[" in capsys.readouterr().out                            # the reference prints the ids


def test_predict_triple_matches_oracle(built, small_ds, capsys):
    from oracle.harness import COracle
    con = _config(small_ds, "TransE", 50)
    P = make_params("TransE", con.entTotal, con.relTotal, 50, seed=22)
    con.set_parameters(P)
    orc = COracle(small_ds)
    # explicit threshold (Config.py:646-652): strict <
    for h, t, r in ((1, 2, 3), (40, 41, 0), (7, 7, 7)):
        s = float(orc.predict("TransE", P, [h], [t], [r])[0])
        for thresh in (s, np.nextafter(np.float32(s), np.float32(10)), 0.0):
            assert con.predict_triple(h, t, r, thresh) == (np.float32(s) < np.float32(thresh))
    # threshold fitted on the valid triples (Config.py:653-659)
    ok = con.predict_triple(5, 6, 2)
    vpos = orc.predict("TransE", P, con.valid_pos_h, con.valid_pos_t, con.valid_pos_r)
    vneg = orc.predict("TransE", P, con.valid_neg_h, con.valid_neg_t, con.valid_neg_r)
    th = orc.best_threshold(vpos, vneg)
    assert np.array_equal(th, con.relThresh)
    assert ok == bool(orc.predict("TransE", P, [5], [6], [2])[0] < th[2])
    out = capsys.readouterr().out
    assert "is correct" in out or "is wrong" in out


def test_roc_helpers_match_reference(built, small_ds):
    """okb_n_interval / okb_tpfp and Config.plot_roc against the reference's own get_n_interval / get_TPFP."""
    from oracle import harness
    if not os.path.exists(harness.REF_SO):
        pytest.skip("reference Base.so not built")
    con = _config(small_ds, "TransH", 50)
    P = make_params("TransH", con.entTotal, con.relTotal, 50, seed=23)
    con.set_parameters(P)
    ref = harness.RefLib().init(small_ds, bern=0, W=2)
    L = ref.L
    L.get_n_interval.restype = ctypes.c_int64
    L.get_n_interval.argtypes = [ctypes.c_int64, vp, vp]
    L.get_TPFP.argtypes = [ctypes.c_int64, vp, vp, vp, vp]
    checked = 0
    import openkeonspark_b200 as okb
    valid_rels = set(int(x) for x in harness.COracle(small_ds).get_list(1)[2])
    for rel in range(con.relTotal):
        if rel not in valid_rels:
            with pytest.raises(okb.OkbError, match="no valid triples"):
                con.plot_roc(rel, fig_name=os.devnull)
            continue
        fpr, tpr, auc = con.plot_roc(rel, fig_name=os.devnull)
        # plot_roc leaves the score inputs behind: recompute them the way it did and hand the SAME arrays to the reference
        pv = con.test_step(con.valid_pos_h, con.valid_pos_t, con.valid_pos_r)
        nv = con.test_step(con.valid_neg_h, con.valid_neg_t, con.valid_neg_r)
        pt = con.test_step(con.test_pos_h, con.test_pos_t, con.test_pos_r)
        nt = con.test_step(con.test_neg_h, con.test_neg_t, con.test_neg_r)
        a = lambda x: vp(x.ctypes.data)
        n_ref = int(L.get_n_interval(rel, a(pv), a(nv)))
        n_got = int(con.lib.okb_n_interval(con.ctx.h, rel, a(pv), a(nv)))
        assert n_got == n_ref, rel
        L.get_TPFP.restype = ctypes.POINTER(ctypes.c_int64 * ((n_ref + 1) * 2))
        exp = np.array(list(L.get_TPFP(rel, a(pv), a(nv), a(pt), a(nt)).contents), dtype=np.int64)
        con.lib.okb_tpfp.restype = ctypes.POINTER(ctypes.c_int64 * ((n_ref + 1) * 2))
        got = np.array(list(con.lib.okb_tpfp(con.ctx.h, rel, a(pv), a(nv), a(pt), a(nt)).contents), dtype=np.int64)
        assert np.array_equal(got, exp), rel
        # Config.py:545-565 on the reference's integers
        TPR, FPR = [], []
        if exp[0] != 0 or exp[n_ref + 1] != 0:
            TPR.append(0); FPR.append(0)
        TPR += [int(x) for x in exp[:n_ref + 1]]
        FPR += [int(x) for x in exp[n_ref + 1:]]
        if TPR[-1] != pt.size or FPR[-1] != nt.size:
            TPR.append(pt.size); FPR.append(nt.size)
        TPR = [x / TPR[-1] for x in TPR]
        FPR = [x / FPR[-1] for x in FPR]
        assert np.allclose(tpr, TPR, rtol=0, atol=0) and np.allclose(fpr, FPR, rtol=0, atol=0), rel
        trap = np.trapezoid if hasattr(np, "trapezoid") else np.trapz
        assert auc == pytest.approx(float(trap(TPR, FPR)), abs=1e-12)
        assert 0.0 <= auc <= 1.0 + 1e-12
        checked += 1
    assert checked >= 5


def test_candidate_fill_matches_reference(built, small_ds):
    """getHeadBatch / getTailBatch of the reference-compatible layer (Test.h:10-26) == the reference's own."""
    from openkeonspark_b200 import _native
    from oracle import harness
    if not os.path.exists(harness.REF_SO):
        pytest.skip("reference Base.so not built")
    lib = _native.load()
    path = small_ds
    lib.setInPath(ctypes.create_string_buffer(path.encode(), len(path) * 2))
    lib.setWorkThreads(ctypes.c_int64(2))
    lib.randReset()
    lib.importTrainFiles(); lib.importTestFiles(); lib.importTypeFiles()
    ref = harness.RefLib().init(small_ds, bern=0, W=2)
    E = ref.E
    lib.getEntityTotal.restype = ctypes.c_int64
    lib.getTestTotal.restype = ctypes.c_int64
    assert lib.getEntityTotal() == E and lib.getTestTotal() == ref.L.getTestTotal()
    for fn_name, side in (("getHeadBatch", 0), ("getTailBatch", 1)):
        fn = getattr(lib, fn_name)
        fn.argtypes = [ctypes.c_int64, vp, vp, vp]
        fn.restype = None
        for idx in (0, 1, ref.L.getTestTotal() // 2, ref.L.getTestTotal() - 1):
            got = [np.full(E, -7, np.int64) for _ in range(3)]
            fn(idx, *[vp(x.ctypes.data) for x in got])
            exp = ref.candidates(side, idx)
            for g, e in zip(got, exp):
                assert np.array_equal(g, e), (fn_name, idx)
    # an index outside the test list prints and returns (Reader.h:36-39 style) instead of reading out of bounds
    got = [np.full(E, -7, np.int64) for _ in range(3)]
    lib.getTailBatch(ctypes.c_int64(10 ** 9), *[vp(x.ctypes.data) for x in got])
    assert all((g == -7).all() for g in got)


@pytest.mark.parametrize("opt", ["SGD", "Adam"])
def test_out_of_range_ids_are_rejected(built, tiny_ds, opt):
    """Config.train_step with an id outside the tables: OKB_ERR_ARG, tables and Adam state untouched."""
    import openkeonspark_b200 as okb
    con = okb.Config(private_context=True)
    con.set_in_path(tiny_ds)
    con.set_nbatches(4); con.set_ent_neg_rate(1); con.set_dimension(16); con.set_opt_method(opt); con.set_alpha(0.01)
    con.workThreads = 2
    con.init()
    con.set_model_and_session(okb.TransE)
    con.set_parameters(make_params("TransE", con.entTotal, con.relTotal, 16, seed=1))
    con.sampling()
    l0 = con.train_step(con.batch_h, con.batch_t, con.batch_r, con.batch_y)
    assert np.isfinite(l0)
    before = con.get_parameters()
    powers = None if con._adam is None else (con._adam["b1p"], con._adam["b2p"])
    step = con._step
    for which, bad in (("batch_h", con.entTotal), ("batch_t", -1), ("batch_r", con.relTotal + 5), ("batch_h", 2 ** 40)):
        con.sampling()
        arr = {n: getattr(con, n).copy() for n in ("batch_h", "batch_t", "batch_r")}
        arr[which][3] = bad
        with pytest.raises(okb.OkbError, match="outside the tables"):
            con.train_step(arr["batch_h"], arr["batch_t"], arr["batch_r"], con.batch_y)
        after = con.get_parameters()
        for k in before:
            assert np.array_equal(before[k], after[k]), (which, bad, k)
        assert con._step == step
        if powers is not None:
            assert (con._adam["b1p"], con._adam["b2p"]) == powers
    # and the context recovers: a clean batch trains again
    con.sampling()
    assert np.isfinite(con.train_step(con.batch_h, con.batch_t, con.batch_r, con.batch_y))
    assert con._step == step + 1
    after = con.get_parameters()
    assert any(not np.array_equal(before[k], after[k]) for k in before)


def test_table_shape_checks(built, tiny_ds):
    import openkeonspark_b200 as okb
    con = okb.Config(private_context=True)
    con.set_in_path(tiny_ds)
    con.set_nbatches(4); con.set_dimension(16)
    con.init()
    con.set_model_and_session(okb.TransE)
    P = make_params("TransE", con.entTotal, con.relTotal, 16, seed=1)
    with pytest.raises(okb.OkbError):
        con.set_parameters({"ent_embeddings": P["ent_embeddings"][:-3]})          # tf.assign would raise on the shape
    con.set_parameters({"ent_embeddings": P["ent_embeddings"][:-3]}, allow_partial_rows=True)   # the incremental-batch case
    con.grow_entities(2)                                                           # tables now disagree with the dataset
    con.sampling_device()
    with pytest.raises(okb.OkbError, match="rows"):
        con.train_step_device(0)


@pytest.mark.parametrize("model,D", [("TransE", 50), ("TransH", 100), ("TransD", 33), ("TransR", 20)])
def test_triple_classification_device_kernels_match_reference(built, small_ds, model, D):
    """csrc/tc.cu (threshold grid search + TP/TN/FP/FN counts on the GPU) and both ways in — Config.test() with device
    scores, and the reference-shaped host-pointer calls — against the reference's own getBestThreshold /
    test_triple_classification (oracle/_ref/Base.so) fed the same score arrays.  Thresholds bit-exact, counts equal."""
    from oracle import harness
    con = _config(small_ds, model, D)
    P = make_params(model, con.entTotal, con.relTotal, D, seed=31)
    con.set_parameters(P)
    con.relThresh[:] = -7.0                            # relations without valid triples must keep their entry (Test.h:310)
    con.test()
    th_gpu, cnt_gpu, acc_gpu = con.relThresh.copy(), con.tc_counts.copy(), float(con.acc[0])
    # the same score arrays, recomputed through the public predict call
    vpos = con.test_step(con.valid_pos_h, con.valid_pos_t, con.valid_pos_r)
    vneg = con.test_step(con.valid_neg_h, con.valid_neg_t, con.valid_neg_r)
    tpos = con.test_step(con.test_pos_h, con.test_pos_t, con.test_pos_r)
    tneg = con.test_step(con.test_neg_h, con.test_neg_t, con.test_neg_r)
    # (a) host-pointer entry points of this library (what the reference's Config.py:503-513 calls)
    th_host = np.full(con.relTotal, -7.0, np.float32)
    con.ctx.call("okb_best_threshold", vp(th_host.ctypes.data), vp(vpos.ctypes.data), vp(vneg.ctypes.data))
    assert np.array_equal(th_host.view(np.uint32), th_gpu.view(np.uint32))
    cnt = np.zeros(4, np.int64); acc = np.zeros(1, np.float32)
    con.ctx.call("okb_tc_eval", vp(th_host.ctypes.data), vp(tpos.ctypes.data), vp(tneg.ctypes.data), vp(cnt.ctypes.data), vp(acc.ctypes.data))
    assert np.array_equal(cnt, cnt_gpu) and float(acc[0]) == acc_gpu
    # (b) the C oracle and (c) the reference library itself
    orc = harness.COracle(small_ds)
    th_o = orc.best_threshold(vpos, vneg, np.full(con.relTotal, -7.0, np.float32))
    assert np.array_equal(th_o.view(np.uint32), th_gpu.view(np.uint32))
    acc_o, cnt_o = orc.tc_eval(th_o, tpos, tneg)
    assert np.array_equal(cnt_o, cnt_gpu) and np.float32(acc_o) == np.float32(acc_gpu)
    if os.path.exists(harness.REF_SO):
        ref = harness.RefLib().init(small_ds, bern=0, W=2)
        th_r = ref.best_threshold(vpos, vneg, np.full(con.relTotal, -7.0, np.float32))
        assert np.array_equal(th_r.view(np.uint32), th_gpu.view(np.uint32))
        assert np.float32(ref.tc_eval(th_r, tpos, tneg)) == np.float32(acc_gpu)
    assert (th_gpu == -7.0).sum() == con.relTotal - len(set(int(x) for x in con.valid_pos_r))
    # early-stop accuracy on the valid ranges: the device count kernel with on_valid = 1 vs a numpy recount
    con._valid_batch_ready = True                      # keep the valid negatives drawn above (a fresh draw would move the thresholds)
    a = con.valid_accuracy()
    vr = np.asarray(con.valid_pos_r)
    ok = (vpos.reshape(-1) <= con.relThresh[vr]).sum() + (vneg.reshape(-1) > con.relThresh[vr]).sum()
    assert np.float32(a) == np.float32(1.0 * ok / (2 * vr.size))


def test_threshold_search_wide_score_range(built, small_ds):
    """Threshold grids with thousands of points per relation and ties between candidates: the FIRST best threshold wins
    (Test.h:333-339), on the device as in the reference."""
    from oracle import harness
    con = _config(small_ds, "TransE", 16)
    con.set_parameters(make_params("TransE", con.entTotal, con.relTotal, 16, seed=3))
    con.ctx.call("okb_tc_batch", 1, *[vp(getattr(con, "valid_" + n + "_addr")) for n in ("pos_h", "pos_t", "pos_r", "neg_h", "neg_t", "neg_r")])
    rng = np.random.default_rng(5)
    n = con.validTotal
    orc = harness.COracle(small_ds)
    for scale, quant in ((40.0, 0.0), (3.0, 0.25), (0.004, 0.0)):
        pos = (rng.random(n) * scale).astype(np.float32)
        neg = (rng.random(n) * scale + 0.3 * scale).astype(np.float32)
        if quant:
            pos, neg = np.round(pos / quant) * np.float32(quant), np.round(neg / quant) * np.float32(quant)
        pos, neg = pos.astype(np.float32), neg.astype(np.float32)
        th = np.zeros(con.relTotal, np.float32)
        con.ctx.call("okb_best_threshold", vp(th.ctypes.data), vp(pos.ctypes.data), vp(neg.ctypes.data))
        exp = orc.best_threshold(pos, neg, np.zeros(con.relTotal, np.float32))
        assert np.array_equal(th.view(np.uint32), exp.view(np.uint32)), (scale, quant)


def _recorrupt(h, t, B, E):
    """Give the first negatives of positives 1..5 a different corrupted entity, keeping the batch's structure (one side
    replaced, the positive's relation — what the kernels, like the sampler, assume of a batch)."""
    for b in range(1, 6):
        s = B + b
        if h[s] != h[b]:
            h[s] = (h[s] + 7) % E
            if h[s] == h[b]:
                h[s] = (h[s] + 1) % E
        else:
            t[s] = (t[s] + 7) % E
            if t[s] == t[b]:
                t[s] = (t[s] + 1) % E


def test_host_batch_fast_path_and_fallback(built, small_ds):
    """Config.sampling() leaves the batch resident and planned; Config.train_step(batch_h, ...) with the UNCHANGED arrays only
    verifies them on the device (no narrowing, no re-plan).  Arrays changed in between — in place or replaced — must give
    exactly what the general path gives: the step is skipped on the device (tables untouched) and redone from the caller's
    arrays."""
    import openkeonspark_b200 as okb

    def fresh():
        con = okb.Config(private_context=True)
        con.set_in_path(small_ds)
        con.set_nbatches(6); con.set_ent_neg_rate(2); con.set_dimension(32); con.set_opt_method("Adam"); con.set_alpha(0.01)
        con.workThreads = 4
        con.init()
        seeds = np.arange(1, 5, dtype=np.uint64) * np.uint64(40503)
        con.ctx.call("okb_set_streams", vp(seeds.ctypes.data), 4)
        con.set_model_and_session(okb.TransH)
        con.set_parameters(make_params("TransH", con.entTotal, con.relTotal, 32, seed=2))
        return con

    a, b, c = fresh(), fresh(), fresh()
    la, lb, lc = [], [], []
    for it in range(4):
        a.sampling()                                          # fast path: hand the pinned block straight back
        la.append(a.train_step(a.batch_h, a.batch_t, a.batch_r, a.batch_y))
        b.sampling()                                          # general path: copies of the arrays (not the pinned block)
        lb.append(b.train_step(b.batch_h.copy(), b.batch_t.copy(), b.batch_r.copy(), b.batch_y))
        c.sampling()                                          # changed IN PLACE after sampling(): another corrupted entity
        _recorrupt(c.batch_h, c.batch_t, c.batch_size, c.entTotal)
        lc.append(c.train_step(c.batch_h, c.batch_t, c.batch_r, c.batch_y))
    assert la == lb
    pa, pb = a.get_parameters(), b.get_parameters()
    for k in pa:
        assert np.array_equal(pa[k], pb[k]), k
    assert a._step == b._step == c._step == 4
    # c: every step fell back (its arrays never matched the resident batch) and still trained, on the MODIFIED batches
    from oracle import models_ref
    d = fresh()
    ref = models_ref.Trainer("TransH", make_params("TransH", d.entTotal, d.relTotal, 32, seed=2), margin=1.0, lr=0.01, opt="Adam")
    for it in range(4):
        d.sampling()
        B = d.batch_size
        h, t, r = d.batch_h.copy(), d.batch_t.copy(), d.batch_r.copy()
        _recorrupt(h, t, B, d.entTotal)
        l = ref.step(h, t, r, B, 2, 0)
        assert abs(lc[it] - l) <= (2e-5 if it == 0 else 1e-3) * abs(l) + 1e-6, (it, lc[it], l)
        d.train_step(h, t, r, d.batch_y)                      # keeps d's sampler in step with c's
    pc, pd = c.get_parameters(), d.get_parameters()
    for k in pc:
        assert np.array_equal(pc[k], pd[k]), k               # fallback == general path, bit for bit
