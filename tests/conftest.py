import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """The C-ABI library and both CPU checkers exist (built here; prebuilt on the GPU box)."""
    from openkeonspark_b200 import _native, build as b
    from oracle import harness
    if not os.path.exists(_native.LIB_PATH):
        b.build()
    if not os.path.exists(harness.ORACLE_SO):
        harness.build()
    return True


def _dataset(tmp_path_factory, shape, **kw):
    from openkeonspark_b200 import datagen
    d = str(tmp_path_factory.mktemp(shape)) + "/"
    g = datagen.make_shape(shape, **{k: v for k, v in kw.items() if k in ("seed", "zipf", "dup_train")})
    datagen.write_dataset(g, d, ontology=kw.get("ontology", True), new_batch=kw.get("new_batch", 0))
    return d


@pytest.fixture(scope="session")
def tiny_ds(tmp_path_factory):
    return _dataset(tmp_path_factory, "tiny", seed=3, dup_train=20)


@pytest.fixture(scope="session")
def small_ds(tmp_path_factory):
    return _dataset(tmp_path_factory, "small", seed=1, zipf=True, dup_train=50)


@pytest.fixture(scope="session")
def small_uniform_ds(tmp_path_factory):
    return _dataset(tmp_path_factory, "small", seed=2)


def make_params(model, E, R, D, seed=0, Dr=None):
    from openkeonspark_b200 import datagen
    rng = np.random.default_rng(seed)
    Dr = D if Dr is None else Dr
    P = {"ent_embeddings": datagen.xavier_normal(rng, E, D), "rel_embeddings": datagen.xavier_normal(rng, R, Dr)}
    if model == "TransH":
        P["normal_vectors"] = datagen.xavier_normal(rng, R, D)
    if model == "TransR":
        P["transfer_matrix"] = datagen.xavier_normal(rng, R, D * Dr)
    if model == "TransD":
        P["ent_transfer"] = datagen.xavier_normal(rng, E, D)
        P["rel_transfer"] = datagen.xavier_normal(rng, R, Dr)
    return P


@pytest.fixture(scope="session")
def wide_ds(tmp_path_factory):
    return _dataset(tmp_path_factory, "wide", seed=4)
