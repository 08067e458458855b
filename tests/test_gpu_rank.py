"""Scores and link-prediction ranks vs the CPU oracle: BIT-EXACT (canonical fp32 order)."""
import ctypes

import numpy as np
import pytest

from conftest import make_params

pytestmark = pytest.mark.gpu


def _config(path, model, D, test_head=1, Dr=None):
    import openkeonspark_b200 as okb
    con = okb.Config(private_context=True)
    con.set_in_path(path)
    con.set_nbatches(4)
    con.set_dimension(D)
    if Dr is not None:
        con.set_rel_dimension(Dr)
    con.set_test_link_prediction(True)
    con.set_test_triple_classification(True)
    con.set_test_head(test_head)
    con.init()
    con.set_model_and_session(getattr(okb, model))
    return con


@pytest.mark.parametrize("model", ["TransE", "TransH", "TransD", "TransR"])
@pytest.mark.parametrize("D", [50, 100, 33])
def test_predict_bit_exact(built, small_ds, model, D):
    from oracle.harness import COracle
    con = _config(small_ds, model, D)
    P = make_params(model, con.entTotal, con.relTotal, D, seed=11)
    con.set_parameters(P)
    orc = COracle(small_ds)
    rng = np.random.default_rng(0)
    n = 777
    h, t, r = rng.integers(0, orc.E, n), rng.integers(0, orc.E, n), rng.integers(0, orc.R, n)
    got = con.test_step(h, t, r)
    assert got.shape == ((n,) if model == "TransE" else (n, 1))
    exp = orc.predict(model, P, h, t, r)
    assert np.array_equal(got.reshape(-1).view(np.uint32), exp.view(np.uint32))


@pytest.mark.parametrize("model", ["TransE", "TransH", "TransD", "TransR"])
@pytest.mark.parametrize("D,test_head", [(50, 1), (100, 0)])
def test_link_prediction_records_bit_exact(built, small_ds, model, D, test_head):
    """GPU all-entity ranking == oracle testHead/testTail fed the oracle's own canonical scores."""
    from oracle.harness import COracle
    con = _config(small_ds, model, D, test_head)
    P = make_params(model, con.entTotal, con.relTotal, D, seed=5)
    # make near-ties likely: quantise a few entity rows so that distinct candidates share scores
    P["ent_embeddings"][10:40] = P["ent_embeddings"][50:80]
    con.set_parameters(P)
    orc = COracle(small_ds)
    rec = con.link_prediction_records().cpu().numpy()
    th, tt, tr = orc.get_list(0)
    ents = np.arange(orc.E)
    for i in range(0, orc.n_test, 3):
        for side in ((0, 1) if test_head else (1,)):
            if side:
                s = orc.predict(model, P, np.full(orc.E, th[i]), ents, np.full(orc.E, tr[i]))
            else:
                s = orc.predict(model, P, ents, np.full(orc.E, tt[i]), np.full(orc.E, tr[i]))
            exp = orc.rank(side, i, s)
            assert np.array_equal(rec[i, side], exp), (i, side, rec[i, side], exp)


def test_candidate_sharding_sums_to_full(built, small_ds):
    """Candidate-sharded evaluation: per-shard counts add / argmins min-combine to the single-GPU result."""
    import torch
    con = _config(small_ds, "TransH", 50)
    con.set_parameters(make_params("TransH", con.entTotal, con.relTotal, 50, seed=2))
    full = con.link_prediction_records().cpu().numpy()
    E, n = con.entTotal, con.testTotal
    dev = con.trainModel.device
    counts = torch.zeros(n * 8, dtype=torch.int64, device=dev)
    best = torch.full((n * 8,), -1, dtype=torch.int64, device=dev)
    m = con._cmodel()
    cuts = [0, 130, 131, 300, E]
    vp = ctypes.c_void_p
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        con.ctx.call("okb_rank", ctypes.byref(m), 0, n, 1, lo, hi, vp(counts.data_ptr()), vp(best.data_ptr()), None)
    out = torch.empty(n * 16, dtype=torch.int64, device=dev)
    con.ctx.call("okb_rank_finalize", 0, n, vp(counts.data_ptr()), vp(best.data_ptr()), vp(out.data_ptr()), None)
    assert np.array_equal(out.view(n, 2, 8).cpu().numpy(), full)


def test_reference_abi_rank_given_scores(built, small_ds):
    """testHead/testTail symbols with caller-provided score vectors == the oracle (Test.h:31-249)."""
    from openkeonspark_b200 import _native
    from oracle.harness import COracle
    lib = _native.load()
    path = small_ds
    lib.setInPath(ctypes.create_string_buffer(path.encode(), len(path) * 2))
    lib.setWorkThreads(ctypes.c_int64(2))
    lib.randReset()
    lib.importTrainFiles(); lib.importTestFiles(); lib.importTypeFiles(); lib.importOntologyFiles()
    orc = COracle(small_ds)
    rng = np.random.default_rng(4)
    for fn, side in ((lib.testHead, 0), (lib.testTail, 1)):
        fn.argtypes = [ctypes.c_int64, ctypes.c_void_p]
        fn.restype = ctypes.POINTER(ctypes.c_int64 * 8)
        for idx in range(0, orc.n_test, 11):
            s = rng.standard_normal(orc.E).astype(np.float32)
            if idx % 2:
                s = np.round(s * 2) / 2
            got = np.array(list(fn(idx, s.ctypes.data).contents))
            assert np.array_equal(got, orc.rank(side, idx, s)), (idx, side)


def test_triple_classification_matches_oracle(built, small_ds):
    """Config.test() triple classification: thresholds, TP/TN/FP/FN and accuracy (Test.h:304-387)."""
    from oracle.harness import COracle
    con = _config(small_ds, "TransE", 50)
    P = make_params("TransE", con.entTotal, con.relTotal, 50, seed=9)
    con.set_parameters(P)
    con.set_test_link_prediction(False)
    con.test_link_prediction = False
    con.test()
    orc = COracle(small_ds)
    vp = orc.predict("TransE", P, con.valid_pos_h, con.valid_pos_t, con.valid_pos_r)
    vn = orc.predict("TransE", P, con.valid_neg_h, con.valid_neg_t, con.valid_neg_r)
    th = orc.best_threshold(vp, vn)
    assert np.array_equal(th, con.relThresh)
    tp = orc.predict("TransE", P, con.test_pos_h, con.test_pos_t, con.test_pos_r)
    tn = orc.predict("TransE", P, con.test_neg_h, con.test_neg_t, con.test_neg_r)
    acc, cnt = orc.tc_eval(th, tp, tn)
    assert np.array_equal(cnt, con.tc_counts)
    assert np.float32(acc) == con.acc[0]
    # negatives must be type-constrained and unknown (Corrupt.h:118-137)
    for h, t, r in list(zip(con.test_neg_h, con.test_neg_t, con.test_neg_r))[:50]:
        assert not orc.find(h, t, r)


def test_transr_rectangular_matrix(built, small_ds):
    """TransR with ent_size != rel_size (set_ent_dimension / set_rel_dimension, Config.py:286-296)."""
    from oracle.harness import COracle
    con = _config(small_ds, "TransR", 40, 1, Dr=24)
    P = make_params("TransR", con.entTotal, con.relTotal, 40, seed=5, Dr=24)
    con.set_parameters(P)
    orc = COracle(small_ds)
    rng = np.random.default_rng(1)
    h, t = rng.integers(0, orc.E, 300), rng.integers(0, orc.E, 300)
    r = np.full(300, 7)
    r[1:] = rng.integers(0, orc.R, 299)          # mixed relations: the matrix of r[0] is used for every row (TransR.py:83)
    got = con.test_step(h, t, r).reshape(-1)
    assert np.array_equal(got.view(np.uint32), orc.predict("TransR", P, h, t, r).view(np.uint32))
    rec = con.link_prediction_records().cpu().numpy()
    th, tt, tr = orc.get_list(0)
    ents = np.arange(orc.E)
    for i in range(0, orc.n_test, 17):
        s = orc.predict("TransR", P, np.full(orc.E, th[i]), ents, np.full(orc.E, tr[i]))
        assert np.array_equal(rec[i, 1], orc.rank(1, i, s)), i
        s = orc.predict("TransR", P, ents, np.full(orc.E, tt[i]), np.full(orc.E, tr[i]))
        assert np.array_equal(rec[i, 0], orc.rank(0, i, s)), i


@pytest.mark.parametrize("D,Dr", [(100, 100), (40, 24), (64, 112)])
def test_transr_tensor_core_projection(built, small_ds, D, Dr):
    """TransR ranking with candidates projected on tcgen05 (3xTF32): the 8-int records must agree with the
    bit-exact canonical path except where two scores are within fp32 rounding of each other.
    Tolerance: counts may differ by at most 2 per record and at least 97 % of the records are identical."""
    con = _config(small_ds, "TransR", D, 1, Dr=Dr)
    P = make_params("TransR", con.entTotal, con.relTotal, D, seed=5, Dr=Dr)
    con.set_parameters(P)
    exact = con.link_prediction_records().cpu().numpy()
    con.transr_tensor_cores = True
    tc = con.link_prediction_records().cpu().numpy()
    assert np.abs(exact[..., :4] - tc[..., :4]).max() <= 2
    same = (exact == tc).all(axis=-1).mean()
    assert same >= 0.97, same
