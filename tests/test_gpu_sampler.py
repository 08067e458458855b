"""GPU sampler vs the CPU oracle: corrupted-triple ids must be BIT-EXACT (integer work)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEEDS = np.array([1804289383, 846930886, 1681692777, 1714636915, 1957747793, 424238335, 719885386, 1649760492], dtype=np.uint64)


def _ctx(path, W, bern, seeds=None):
    from openkeonspark_b200 import _native
    c = _native.Ctx()
    c.call("okb_set_in_path", path.encode())
    c.call("okb_set_bern", bern)
    s = np.ascontiguousarray((SEEDS if seeds is None else seeds)[:W])
    c.call("okb_set_streams", ctypes.c_void_p(s.ctypes.data), W)
    c.call("okb_import_train_files")
    return c


def _gpu_batch(c, B, k, kr, step=0):
    S = B * (1 + k + kr)
    h, t, r = (np.zeros(S, np.int64) for _ in range(3))
    y = np.zeros(S, np.float32)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    c.call("okb_batch_to_host", step, p(h), p(t), p(r), p(y), None)
    return h, t, r, y


@pytest.mark.parametrize("bern", [0, 1])
@pytest.mark.parametrize("B,k,kr,W", [(100, 1, 0, 8), (101, 3, 0, 8), (37, 2, 1, 5), (3, 1, 0, 8), (600, 10, 2, 1), (4831, 1, 0, 8)])
def test_sampler_bit_exact(built, small_ds, bern, B, k, kr, W):
    from oracle.harness import COracle
    orc = COracle(small_ds)
    orc.set_streams(SEEDS[:W], bern)
    c = _ctx(small_ds, W, bern)
    for it in range(3):                      # stream state persists across calls (Random.h:6)
        exp = orc.sampling(B, k, kr)
        c.call("okb_sample", B, k, kr, 1, 0, W, None)
        got = _gpu_batch(c, B, k, kr)
        for a, b, nm in zip(exp, got, "htry"):
            assert np.array_equal(a, b), (nm, it, np.flatnonzero(a != b)[:5])
    st = np.zeros(W, np.uint64)
    c.call("okb_get_streams", ctypes.c_void_p(st.ctypes.data), W)
    assert np.array_equal(st, orc.streams())
    c.close()


def test_multi_step_launch_equals_consecutive_calls(built, small_ds):
    """One launch producing 5 sampling() calls == 5 reference calls (LCG jump-ahead)."""
    from oracle.harness import COracle
    B, k, kr, W = 333, 2, 1, 8
    orc = COracle(small_ds)
    orc.set_streams(SEEDS[:W], 1)
    c = _ctx(small_ds, W, 1)
    c.call("okb_sample", B, k, kr, 5, 0, W, None)
    for step in range(5):
        exp = orc.sampling(B, k, kr)
        got = _gpu_batch(c, B, k, kr, step)
        for a, b in zip(exp, got):
            assert np.array_equal(a, b), step
    st = np.zeros(W, np.uint64)
    c.call("okb_get_streams", ctypes.c_void_p(st.ctypes.data), W)
    assert np.array_equal(st, orc.streams())
    c.close()


def test_incremental_batch_mode(built, tmp_path_factory):
    """batch2id.txt present: positives come from the last newBatchTotal rows (Base.cpp:101-103)."""
    from openkeonspark_b200 import datagen
    from oracle.harness import COracle
    d = str(tmp_path_factory.mktemp("inc")) + "/"
    g = datagen.make_shape("small", seed=5)
    datagen.write_dataset(g, d, new_batch=700)
    orc = COracle(d)
    orc.set_streams(SEEDS[:4], 0)
    c = _ctx(d, 4, 0)
    exp = orc.sampling(250, 2, 0)
    c.call("okb_sample", 250, 2, 0, 1, 0, 4, None)
    got = _gpu_batch(c, 250, 2, 0)
    for a, b in zip(exp, got):
        assert np.array_equal(a, b)
    c.close()


def test_reference_abi_sampling(built, small_ds):
    """The Base.so-compatible symbols: same call sequence as the reference's Config.init/sampling."""
    from openkeonspark_b200 import _native
    from oracle.harness import COracle
    lib = _native.load()
    path = small_ds
    lib.setInPath(ctypes.create_string_buffer(path.encode(), len(path) * 2))
    lib.setBern(ctypes.c_int64(1))
    lib.setWorkThreads(ctypes.c_int64(8))
    lib.randReset()
    lib.importTrainFiles()
    ctx = _native.Ctx(default=True)
    st = np.zeros(8, np.uint64)
    ctx.call("okb_get_streams", ctypes.c_void_p(st.ctypes.data), 8)
    orc = COracle(small_ds)
    orc.set_streams(st, 1)
    lib.sampling.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64] * 3
    B, k = 200, 2
    S = B * (1 + k)
    h, t, r = (np.zeros(S, np.int64) for _ in range(3))
    y = np.zeros(S, np.float32)
    lib.sampling(h.ctypes.data, t.ctypes.data, r.ctypes.data, y.ctypes.data, B, k, 0)
    exp = orc.sampling(B, k, 0)
    for a, b in zip(exp, (h, t, r, y)):
        assert np.array_equal(a, b)
    lib.getEntityTotal.restype = ctypes.c_int64
    assert lib.getEntityTotal() == orc.E


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("B,k,kr,W", [(100, 1, 0, 8), (37, 2, 1, 5), (4831, 1, 0, 8)])
def test_sample_to_host_bit_exact(built, small_ds, pinned, B, k, kr, W):
    """okb_sample_to_host = one reference sampling() call (Base.cpp:153-177) into the caller's int64 arrays: with one
    page-locked block the sample kernel mirrors the batch over PCIe itself, otherwise okb_sample + okb_batch_to_host.
    Ids, the device-resident copy and the stream states must match the CPU oracle bit for bit in both forms."""
    import torch
    from oracle.harness import COracle
    orc = COracle(small_ds)
    orc.set_streams(SEEDS[:W], 1)
    c = _ctx(small_ds, W, 1)
    S = B * (1 + k + kr)
    blk = torch.zeros(3 * S, dtype=torch.int64)
    if pinned:
        blk = blk.pin_memory()
    a = blk.numpy()
    p = lambda off: ctypes.c_void_p(blk.data_ptr() + 8 * off)
    for it in range(3):
        exp = orc.sampling(B, k, kr)
        a[:] = -7
        c.call("okb_sample_to_host", B, k, kr, 0, W, p(0), p(S), p(2 * S), None)
        for j, nm in enumerate("htr"):
            assert np.array_equal(exp[j], a[j * S:(j + 1) * S]), (nm, it)
        dev = _gpu_batch(c, B, k, kr)                 # the batch also stays resident as step 0
        for x, y_, nm in zip(exp, dev, "htry"):
            assert np.array_equal(x, y_), ("device", nm, it)
    st = np.zeros(W, np.uint64)
    c.call("okb_get_streams", ctypes.c_void_p(st.ctypes.data), W)
    assert np.array_equal(st, orc.streams())
    c.close()
