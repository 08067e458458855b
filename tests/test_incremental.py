"""Incremental batches: the file side (main_spark.py:100-195) on CPU, the model side (table growth) and early stopping on the GPU."""
import os
import shutil
import tempfile

import numpy as np
import pytest


def _write(path, name, lines):
    with open(os.path.join(path, name), "w") as f:
        f.write(str(len(lines)) + "\n")
        f.writelines(l + "\n" for l in lines)


def test_feed_batch_files():
    from openkeonspark_b200 import incremental
    d = tempfile.mkdtemp()
    try:
        _write(d, "entity2id.txt", ["e0\t0", "e1\t1", "e2\t2"])
        _write(d, "train2id.txt", ["0 1 0", "1 2 0"])
        _write(d, "test2id.txt", ["0 2 0"])
        _write(d, "valid2id.txt", ["2 0 0"])
        assert not incremental.is_new_batch(d)
        _write(d, "batch2id.txt", ["3 0 0", "4 3 0", "0 4 0"])
        _write(d, "batchEntity2id.txt", ["e3\t3", "e4\t4"])
        _write(d, "batchTest2id.txt", ["3 4 0"])
        _write(d, "batchValid2id.txt", [])
        assert incremental.is_new_batch(d)
        r = incremental.feed_batch(d)
        assert r == {"new_entities": 2, "final_entities": 5, "new_train": 3, "new_test": 1, "new_valid": 0}
        lines = open(os.path.join(d, "train2id.txt")).read().split("\n")
        assert lines[0] == "5" and lines[1:6] == ["0 1 0", "1 2 0", "3 0 0", "4 3 0", "0 4 0"]     # new triples are the LAST rows
        assert open(os.path.join(d, "entity2id.txt")).readline().strip() == "5"
        assert open(os.path.join(d, "test2id.txt")).readline().strip() == "2"
        assert open(os.path.join(d, "valid2id.txt")).readline().strip() == "1"
        # batch2id.txt stays: its first line switches the sampler to the last newBatchTotal rows (Reader.h:61-67)
        assert open(os.path.join(d, "batch2id.txt")).readline().strip() == "3"
        incremental.remove_batch_files(d)
        assert not incremental.is_new_batch(d) and not os.path.exists(os.path.join(d, "batch2id.txt"))
    finally:
        shutil.rmtree(d)


@pytest.mark.gpu
def test_grow_entities_and_restore_into_grown_tables(tmp_path):
    import torch
    import openkeonspark_b200 as okb
    from openkeonspark_b200 import datagen
    g = datagen.make_shape("small", seed=3)
    con = okb.Config(private_context=True)
    con.set_nbatches(10); con.set_dimension(20); con.set_opt_method("Adam")
    con.init_from_arrays(g.E, g.R, g.train, g.valid, g.test)
    con.set_model_and_session(okb.TransD)
    con.train_chunk_device(4)
    before = con.get_parameters()
    m_before = con._adam["m_ent_embeddings"].clone()
    assert con.grow_entities(7) == g.E + 7
    after = con.get_parameters()
    for k in ("ent_embeddings", "ent_transfer"):
        assert after[k].shape == (g.E + 7, 20)
        assert np.array_equal(after[k][:g.E], before[k])                   # old rows untouched
        std = np.sqrt(2.6 / (g.E + 7 + 20))
        assert np.all(np.abs(after[k][g.E:]) <= 2 * std + 1e-7) and after[k][g.E:].std() > 0.3 * std   # truncated normal of the FINAL shape
    assert np.array_equal(after["rel_embeddings"], before["rel_embeddings"])
    assert torch.equal(con._adam["m_ent_embeddings"][:g.E], m_before) and float(con._adam["m_ent_embeddings"][g.E:].abs().sum()) == 0.0
    # restart flow: a checkpoint of the old shape loads into a Config built on the grown dataset
    con.set_export_files(str(tmp_path / "ckpt.pt"))
    con.grow_entities(0)
    small = okb.Config(private_context=True)
    small.set_nbatches(10); small.set_dimension(20); small.set_opt_method("Adam")
    small.init_from_arrays(g.E, g.R, g.train, g.valid, g.test)
    small.set_model_and_session(okb.TransD)
    small.set_export_files(str(tmp_path / "ckpt.pt"))
    small.save_tensorflow()
    big = okb.Config(private_context=True)
    big.set_nbatches(1); big.set_dimension(20); big.set_opt_method("Adam")      # batch = the 2 new rows (Config.py:189-210 with bt > 0)
    train2 = np.concatenate([g.train, np.array([[g.E, 0, 0], [1, g.E + 1, 0]])])
    big.init_from_arrays(g.E + 2, g.R, train2, g.valid, g.test, new_batch=2)
    big.set_model_and_session(okb.TransD)
    fresh = big.get_parameters()["ent_embeddings"][g.E:].copy()
    big.set_import_files(str(tmp_path / "ckpt.pt"))
    big.restore_tensorflow()
    p = big.get_parameters()
    assert np.array_equal(p["ent_embeddings"][:g.E], small.get_parameters()["ent_embeddings"])
    assert np.array_equal(p["ent_embeddings"][g.E:], fresh)
    big.train_chunk_device(3)                                              # trains on the two new rows only (Base.cpp:101-103)
    hb = big.ctx  # noqa: F841


@pytest.mark.gpu
def test_early_stopping_restores_best_model(small_uniform_ds):
    """distribute_training.py:286-356: with a learning rate that makes the loss rise, the loss patience runs out, training
    stops early and the tables of the best check are restored."""
    import openkeonspark_b200 as okb
    from conftest import make_params
    con = okb.Config(private_context=True)
    con.set_in_path(small_uniform_ds)
    con.set_nbatches(4); con.set_train_times(40); con.set_dimension(16); con.set_alpha(0.5); con.set_opt_method("SGD")
    con.set_valid_triple_classification(True)
    con.init()
    con.set_model_and_session(okb.TransE)
    con.set_parameters(make_params("TransE", con.entTotal, con.relTotal, 16, seed=1))
    con.set_early_stopping(patience=2, start_step=1, stopping_step=1)
    con.run()
    es = con.early_stop
    assert es is not None and len(es["checks"]) >= 1
    steps = [c[0] for c in es["checks"]]
    assert steps[0] == 4 and all(b - a == 4 for a, b in zip(steps[:-1], steps[1:]))      # a check per epoch from epoch 1
    assert all(0.0 <= c[1] <= 1.0 for c in es["checks"])
    if es["reason"] is not None:
        assert con._step == es["best_step"] and es["best_step"] in steps                  # best model restored
        assert len(steps) < 40
    # the valid accuracy equals a direct count over the valid ranges with the fitted thresholds
    acc = con.valid_accuracy()
    pos = con.test_step(con.valid_pos_h, con.valid_pos_t, con.valid_pos_r).reshape(-1)
    neg = con.test_step(con.valid_neg_h, con.valid_neg_t, con.valid_neg_r).reshape(-1)
    th = con.relThresh[con.valid_pos_r]
    want = (np.sum(pos <= th) + np.sum(neg > th)) / (2.0 * pos.size)
    assert abs(acc - want) < 1e-6
