"""The C oracle (oracle/kge_oracle.c) against golden vectors produced by the reference's own
Base.so (tests/golden/make_golden.py).  CPU only; nothing here needs /root/reference."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["tiny_uniform", "small_zipf_bern", "small_w3"]


def materialise(g, d):
    """Write the fixture's dataset back to the reference's text format."""
    from openkeonspark_b200 import datagen
    graph = datagen.Graph(int(g["E"]), int(g["R"]), g["train"], g["valid"], g["test"])
    datagen.write_dataset(graph, d)
    open(os.path.join(d, "type_constrain.txt"), "wb").write(g["type_constrain_txt"].tobytes())
    open(os.path.join(d, "ontology_constrain.txt"), "wb").write(g["ontology_constrain_txt"].tobytes())
    return d


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_vectors(built, tmp_path, name):
    from oracle.harness import COracle
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    d = materialise(g, str(tmp_path) + "/")
    orc = COracle(d)
    orc.set_streams(g["seeds"], int(g["bern"]))
    for i, (B, k, kr) in enumerate(g["calls"]):
        for rep in range(2):
            h, t, r, y = orc.sampling(int(B), int(k), int(kr))
            assert np.array_equal(h, g["s%d_%d_h" % (i, rep)])
            assert np.array_equal(t, g["s%d_%d_t" % (i, rep)])
            assert np.array_equal(r, g["s%d_%d_r" % (i, rep)])
            assert np.array_equal(y[:B], np.ones(B, np.float32)) and np.all(y[B:] == -1)
    assert np.array_equal(orc.streams(), g["seeds_after"])
    for a, i in enumerate(g["rank_idx"]):
        for side in (0, 1):
            assert np.array_equal(orc.rank(side, int(i), g["rank_scores"][a]), g["rank_rec"][a, side])
    th = orc.best_threshold(g["tc_vp"], g["tc_vn"])
    assert np.array_equal(th, g["tc_thresh"])
    acc, _ = orc.tc_eval(th, g["tc_tp"], g["tc_tn"])
    assert np.float32(acc) == g["tc_acc"]
    # the negatives of the golden batches are type-constrained unknown tails (Corrupt.h:118-137)
    vb = g["valid_batch"]
    for h, t, r in list(zip(vb[3], vb[4], vb[5]))[:40]:
        assert not orc.find(int(h), int(t), int(r))


def test_first_seeds_are_glibc_default_rand():
    """Random.h:12 seeds from an un-srand()ed libc rand(): 1804289383, 846930886, ... in a fresh process."""
    g = np.load(os.path.join(HERE, "golden", "tiny_uniform.npz"))
    assert list(g["seeds"][:4]) == [1804289383, 846930886, 1681692777, 1714636915]
