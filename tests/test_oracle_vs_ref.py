"""The C restatement against the reference library itself (oracle/_ref/Base.so, compiled from
/root/reference/base/Base.cpp).  Skipped where the reference library is absent."""
import os

import numpy as np
import pytest

from oracle import harness

needs_ref = pytest.mark.skipif(not os.path.exists(harness.REF_SO) and not os.path.exists("/root/reference/base/Base.cpp"),
                               reason="reference Base.so not available")


@needs_ref
@pytest.mark.parametrize("bern", [0, 1])
@pytest.mark.parametrize("B,k,kr,W", [(100, 1, 0, 8), (101, 3, 0, 8), (37, 2, 1, 5), (3, 1, 0, 8), (600, 10, 2, 1)])
def test_sampling_matches_reference(built, small_ds, bern, B, k, kr, W):
    ref = harness.RefLib().init(small_ds, bern=bern, W=W)
    orc = harness.COracle(small_ds)
    orc.set_streams(ref.seeds(), bern)
    for it in range(3):
        a, b = ref.sampling(B, k, kr), orc.sampling(B, k, kr)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), it
    assert np.array_equal(ref.seeds(), orc.streams())


@needs_ref
def test_index_statistics_match_reference(built, small_ds):
    ref = harness.RefLib().init(small_ds, bern=1, W=2, test=False)
    orc = harness.COracle(small_ds)
    assert ref.L.getTrainTotal() == orc.L.orc_n_dedup(orc.o)      # deduplicated (Reader.h:103-123)
    assert ref.L.getTrainTotal_() == orc.n_raw
    ref.L.importTestFiles()
    assert ref.L.getTrainTotal() == orc.n_raw                      # Reader.h:227 overwrites it with the file header
    assert ref.L.getTripleTotal() == orc.n_test + orc.n_raw + orc.n_valid


@needs_ref
def test_rank_matches_reference(built, small_ds):
    ref = harness.RefLib().init(small_ds, 0, 8)
    orc = harness.COracle(small_ds)
    rng = np.random.default_rng(0)
    for idx in range(0, orc.n_test, 5):
        for side in (0, 1):
            s = rng.standard_normal(orc.E).astype(np.float32)
            if idx % 3 == 0:
                s = np.round(s * 2) / 2
            assert np.array_equal(ref.rank(side, idx, s), orc.rank(side, idx, s)), (idx, side)


@needs_ref
def test_rank_without_ontology_file(built, small_uniform_ds, tmp_path):
    import shutil
    d = str(tmp_path) + "/"
    for f in os.listdir(small_uniform_ds):
        if f != "ontology_constrain.txt":
            shutil.copy(os.path.join(small_uniform_ds, f), d + f)
    ref = harness.RefLib().init(d, 0, 8)
    orc = harness.COracle(d)
    s = np.random.default_rng(1).standard_normal(orc.E).astype(np.float32)
    for idx in (0, 5, 17):
        assert np.array_equal(ref.rank(1, idx, s), orc.rank(1, idx, s))


@needs_ref
def test_triple_classification_matches_reference(built, small_ds):
    ref = harness.RefLib().init(small_ds, 0, 8)
    orc = harness.COracle(small_ds)
    lists = orc.get_list(1)
    vb = ref.tc_batch(1)
    assert all(np.array_equal(a, b) for a, b in zip(vb[:3], lists))
    rng = np.random.default_rng(3)
    sp = (rng.random(orc.n_valid) * 4).astype(np.float32)
    sn = (rng.random(orc.n_valid) * 4 + 1).astype(np.float32)
    th_r, th_o = ref.best_threshold(sp, sn), orc.best_threshold(sp, sn)
    assert np.array_equal(th_r, th_o)
    tp = (rng.random(orc.n_test) * 4).astype(np.float32)
    tn = (rng.random(orc.n_test) * 4 + 1).astype(np.float32)
    acc_o, _ = orc.tc_eval(th_o, tp, tn)
    assert np.float32(ref.tc_eval(th_r, tp, tn)) == np.float32(acc_o)


def test_scores_match_fp64_shadow(built, tiny_ds):
    """Floating-point half (parity unpinned by the reference): canonical C scores vs torch fp64."""
    import torch
    from conftest import make_params
    from oracle import models_ref
    orc = harness.COracle(tiny_ds)
    rng = np.random.default_rng(0)
    h, t, r = rng.integers(0, orc.E, 200), rng.integers(0, orc.E, 200), rng.integers(0, orc.R, 200)
    for model in models_ref.NAMES:
        P = make_params(model, orc.E, orc.R, 24, seed=1)
        got = orc.predict(model, P, h, t, r)
        if model == "TransR":
            r0 = np.full_like(r, r[0])
            P64 = {k: torch.tensor(v, dtype=torch.float64) for k, v in P.items()}
            exp = models_ref.predict_fn(model, P64, h, t, r).numpy().reshape(-1)
        else:
            P64 = {k: torch.tensor(v, dtype=torch.float64) for k, v in P.items()}
            exp = models_ref.predict_fn(model, P64, h, t, r).numpy().reshape(-1)
        assert np.allclose(got, exp, rtol=2e-6, atol=2e-6), model


def test_loss_gradients_fp32_vs_fp64(built, tiny_ds):
    import torch
    from conftest import make_params
    from oracle import models_ref
    orc = harness.COracle(tiny_ds)
    orc.set_streams([11, 22, 33, 44], 1)
    h, t, r, _ = orc.sampling(40, 3, 1)
    for model in models_ref.NAMES:
        P = make_params(model, orc.E, orc.R, 16, seed=2)
        for opt in ("SGD", "Adam"):
            a = models_ref.Trainer(model, P, lr=0.01, opt=opt)
            b = models_ref.Trainer(model, P, lr=0.01, opt=opt, dtype=torch.float64)
            for _ in range(2):
                la, lb = a.step(h, t, r, 40, 3, 1), b.step(h, t, r, 40, 3, 1)
            assert abs(la - lb) < 1e-5
            for k, v in b.params().items():
                assert np.allclose(a.params()[k], v, atol=5e-6), (model, opt, k)
