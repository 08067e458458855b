"""Fused train step vs the CPU restatement of the TF graphs (oracle/models_ref.py).

Stated fp32 tolerances (against the fp64 run of the restatement):
  loss .................. |dl| <= 2e-5 * |l| + 1e-6
  SGD tables ............ max|err| <= max(4 x the fp32 CPU restatement's own error, 1e-4 * max|update| + 2e-6)
  Adam slots m, v ....... (after the FIRST step) linear / quadratic in the gradient, so they carry the gradient check:
                          99.9 % of elements: |dm| <= 1e-4 * max|m| + 1e-9 ; |dv| <= 2e-4 * max|v| + 1e-12
  Adam tables ........... Adam normalises every element's step to ~lr regardless of |g|, so an element whose
                          gradient is ~0 flips between -lr and +lr on fp32 rounding noise (the fp32 CPU
                          restatement itself is up to lr away from fp64 there), and the flips feed the next step.
                          Hence after 3 steps: median error within 1e-5, no element further than 2 * lr * steps;
                          the sharp check is the one on the slots after the first step.
The oracle itself is "parity unpinned" by the reference (TF 1.x is not installable)."""
import ctypes

import numpy as np
import pytest

from conftest import make_params

pytestmark = pytest.mark.gpu


def _config(path, model, D, k, kr, opt, nbatches=6, alpha=0.01, margin=1.0, W=4, bern=1, Dr=None):
    import openkeonspark_b200 as okb
    con = okb.Config(private_context=True)
    con.set_in_path(path)
    con.set_nbatches(nbatches)
    con.set_ent_neg_rate(k)
    con.set_rel_neg_rate(kr)
    con.set_margin(margin)
    con.set_alpha(alpha)
    con.set_opt_method(opt)
    con.set_dimension(D)
    if Dr is not None:
        con.set_rel_dimension(Dr)
    con.set_bern(bern)
    con.workThreads = W
    con.init()
    # fixed streams: init() seeds from libc rand(), whose position depends on what ran before in the process
    seeds = np.arange(1, W + 1, dtype=np.uint64) * np.uint64(2654435761)
    con.ctx.call("okb_set_streams", ctypes.c_void_p(seeds.ctypes.data), W)
    con.set_model_and_session(getattr(okb, model))
    return con


def _check_adam_slots(con, ref64):
    """m and v are linear / quadratic in the gradient.  A hinge term within rounding of 0 may be active in fp32
    and inactive in fp64 (whole rows then differ), so the bound is on the 99.9 % quantile, not the maximum."""
    for name in ref64.m:
        m_gpu, v_gpu = con._adam["m_" + name].cpu().numpy(), con._adam["v_" + name].cpu().numpy()
        m_ref, v_ref = ref64.m[name].numpy(), ref64.v[name].numpy()
        assert np.quantile(np.abs(m_gpu - m_ref), 0.999) <= 1e-4 * np.abs(m_ref).max() + 1e-9, name
        assert np.quantile(np.abs(v_gpu - v_ref), 0.999) <= 2e-4 * np.abs(v_ref).max() + 1e-12, name


@pytest.mark.parametrize("model", ["TransE", "TransH", "TransD"])
@pytest.mark.parametrize("opt", ["SGD", "Adam"])
@pytest.mark.parametrize("D,k,kr", [(50, 1, 0), (100, 3, 1), (20, 2, 0), (200, 1, 0), (100, 1, 0), (33, 1, 0), (33, 2, 1)])
def test_train_step_parity(built, small_ds, model, opt, D, k, kr):
    import torch
    from oracle import models_ref
    con = _config(small_ds, model, D, k, kr, opt)
    P = make_params(model, con.entTotal, con.relTotal, D, seed=7)
    con.set_parameters(P)
    ref32 = models_ref.Trainer(model, P, margin=1.0, lr=0.01, opt=opt)
    ref64 = models_ref.Trainer(model, P, margin=1.0, lr=0.01, opt=opt, dtype=torch.float64)
    B = con.batch_size
    for it in range(3):
        con.sampling()                                   # GPU sampler -> host arrays (bit-exact, tested elsewhere)
        h, t, r = con.batch_h.copy(), con.batch_t.copy(), con.batch_r.copy()
        loss = float(con.train_step_device(0).item())
        l32 = ref32.step(h, t, r, B, k, kr)
        l64 = ref64.step(h, t, r, B, k, kr)
        # Adam: from the second step on the fp32/fp64 tables differ by up to lr in near-zero-gradient elements (see
        # the module docstring), which feeds back into the loss; only the first step isolates the forward pass
        tol = 2e-5 if (opt == "SGD" or it == 0) else 1e-3
        assert abs(loss - l64) <= tol * abs(l64) + 1e-6, (it, loss, l32, l64)
        if it == 0 and opt == "Adam":          # after ONE step m = 0.1 g and v = 0.001 g^2: the pure gradient check
            _check_adam_slots(con, ref64)
    got = con.get_parameters()
    exp = ref64.params()
    for name in exp:
        delta = np.abs(exp[name] - P[name]).max()
        err = np.abs(got[name] - exp[name])
        err32 = np.abs(ref32.params()[name] - exp[name]).max()
        if opt == "SGD":
            assert err.max() <= max(4 * err32, 1e-4 * delta + 2e-6), (name, err.max(), err32, delta)
        else:
            assert np.median(err) <= 1e-5, (name, np.median(err))
            assert err.max() <= 2 * 0.01 * 3 + 1e-6, (name, err.max())


def test_train_step_deterministic(built, small_ds):
    """Same seeds + same parameters -> bit-identical tables (sorted, fixed-order gradient reduction)."""
    outs = []
    for rep in range(2):
        con = _config(small_ds, "TransH", 100, 2, 0, "Adam")
        seeds = np.arange(1, 5, dtype=np.uint64) * 1234567
        con.ctx.call("okb_set_streams", ctypes.c_void_p(seeds.ctypes.data), 4)
        con.set_parameters(make_params("TransH", con.entTotal, con.relTotal, 100, seed=3))
        for it in range(4):
            con.sampling_device()
            con.train_step_device(0)
        outs.append(con.get_parameters())
    for name in outs[0]:
        assert np.array_equal(outs[0][name], outs[1][name]), name


def test_train_step_host_batch_api(built, tiny_ds):
    """Config.train_step(batch_h, batch_t, batch_r, batch_y) — the reference's signature (Config.py:464)."""
    from oracle import models_ref
    con = _config(tiny_ds, "TransE", 16, 2, 0, "SGD", nbatches=4)
    P = make_params("TransE", con.entTotal, con.relTotal, 16, seed=1)
    con.set_parameters(P)
    con.sampling()
    loss = con.train_step(con.batch_h, con.batch_t, con.batch_r, con.batch_y)
    ref = models_ref.Trainer("TransE", P, lr=0.01)
    l = ref.step(con.batch_h, con.batch_t, con.batch_r, con.batch_size, 2, 0)
    assert abs(loss - l) < 1e-5
    got = con.get_parameters()
    for name, v in ref.params().items():
        assert np.allclose(got[name], v, atol=1e-6)


def test_run_loop_reduces_loss(built, small_uniform_ds):
    con = _config(small_uniform_ds, "TransE", 32, 1, 0, "SGD", nbatches=10, alpha=0.5)
    con.set_train_times(6)
    losses = con.run()
    assert losses[-1] < losses[0]


def test_chunked_lookahead_equals_step_by_step(built, small_ds):
    """next_step_device() (64 steps sampled + planned per launch) == sampling_device()+train_step_device() per step."""
    outs = []
    for chunked in (False, True):
        con = _config(small_ds, "TransD", 50, 2, 1, "Adam")
        seeds = np.arange(1, 5, dtype=np.uint64) * 999331
        con.ctx.call("okb_set_streams", ctypes.c_void_p(seeds.ctypes.data), 4)
        con.set_parameters(make_params("TransD", con.entTotal, con.relTotal, 50, seed=3))
        con.plan_ahead = 5
        losses = []
        for it in range(12):
            if chunked:
                losses.append(float(con.next_step_device().item()))
            else:
                con.sampling_device()
                losses.append(float(con.train_step_device(0).item()))
        outs.append((losses, con.get_parameters()))
    # and the one-call-per-chunk C loop (Config.train_chunk_device / okb_train_steps)
    con = _config(small_ds, "TransD", 50, 2, 1, "Adam")
    con.ctx.call("okb_set_streams", ctypes.c_void_p(seeds.ctypes.data), 4)
    con.set_parameters(make_params("TransD", con.entTotal, con.relTotal, 50, seed=3))
    losses = []
    for n in (5, 5, 2):
        losses += [float(x) for x in con.train_chunk_device(n).cpu().numpy()]
    outs.append((losses, con.get_parameters()))
    assert outs[0][0] == outs[1][0] == outs[2][0]
    for name in outs[0][1]:
        assert np.array_equal(outs[0][1][name], outs[1][1][name]), name
        assert np.array_equal(outs[0][1][name], outs[2][1][name]), name


@pytest.mark.parametrize("opt", ["SGD", "Adam"])
@pytest.mark.parametrize("D,Dr,k", [(100, 100, 1), (20, 20, 2), (40, 24, 3), (64, 128, 10)])
def test_transr_train_step_parity(built, small_ds, opt, D, Dr, k):
    """TransR: relation-bucketed kernel (projection, dA, dM contractions) vs the TF-graph restatement."""
    import torch
    from oracle import models_ref
    con = _config(small_ds, "TransR", D, k, 0, opt, Dr=Dr)
    P = make_params("TransR", con.entTotal, con.relTotal, D, seed=7, Dr=Dr)
    con.set_parameters(P)
    ref32 = models_ref.Trainer("TransR", P, margin=1.0, lr=0.01, opt=opt)
    ref64 = models_ref.Trainer("TransR", P, margin=1.0, lr=0.01, opt=opt, dtype=torch.float64)
    B = con.batch_size
    for it in range(3):
        con.sampling()
        h, t, r = con.batch_h.copy(), con.batch_t.copy(), con.batch_r.copy()
        loss = float(con.train_step_device(0).item())
        ref32.step(h, t, r, B, k, 0)
        l64 = ref64.step(h, t, r, B, k, 0)
        tol = 2e-5 if (opt == "SGD" or it == 0) else 1e-3
        assert abs(loss - l64) <= tol * abs(l64) + 1e-6, (it, loss, l64)
        if it == 0 and opt == "Adam":
            _check_adam_slots(con, ref64)
    got, exp = con.get_parameters(), ref64.params()
    for name in exp:
        delta = np.abs(exp[name] - P[name]).max()
        err = np.abs(got[name] - exp[name])
        err32 = np.abs(ref32.params()[name] - exp[name]).max()
        if opt == "SGD":
            assert err.max() <= max(4 * err32, 1e-4 * delta + 2e-6), (name, err.max(), err32, delta)
        else:
            assert np.median(err) <= 1e-5, (name, np.median(err))
            assert err.max() <= 2 * 0.01 * 3 + 1e-6, (name, err.max())


@pytest.mark.parametrize("opt", ["SGD", "Adam"])
@pytest.mark.parametrize("D,Dr,k,kr", [(20, 20, 1, 1), (40, 24, 2, 2), (100, 100, 1, 1), (16, 32, 3, 1)])
def test_transr_relation_negatives_parity(built, small_ds, opt, D, Dr, k, kr):
    """TransR with rel_neg_rate > 0 (TransR.py:61-65: every negative is projected by ITS relation's matrix): the general
    kernel (one CTA per positive, a [d rel | d M] gradient row per (positive, relation slot), summed per relation in slot
    order by the relation update) against the TF-graph restatement, same bars as the default TransR path.  small_ds has 23
    relations, so negative relations coincide with other positives' relations and long relation segments occur."""
    import torch
    from oracle import models_ref
    con = _config(small_ds, "TransR", D, k, kr, opt, Dr=Dr)
    P = make_params("TransR", con.entTotal, con.relTotal, D, seed=7, Dr=Dr)
    con.set_parameters(P)
    ref32 = models_ref.Trainer("TransR", P, margin=1.0, lr=0.01, opt=opt)
    ref64 = models_ref.Trainer("TransR", P, margin=1.0, lr=0.01, opt=opt, dtype=torch.float64)
    B = con.batch_size
    for it in range(3):
        con.sampling()
        h, t, r = con.batch_h.copy(), con.batch_t.copy(), con.batch_r.copy()
        loss = float(con.train_step_device(0).item())
        ref32.step(h, t, r, B, k, kr)
        l64 = ref64.step(h, t, r, B, k, kr)
        tol = 2e-5 if (opt == "SGD" or it == 0) else 1e-3
        assert abs(loss - l64) <= tol * abs(l64) + 1e-6, (it, loss, l64)
        if it == 0 and opt == "Adam":
            _check_adam_slots(con, ref64)
    got, exp = con.get_parameters(), ref64.params()
    for name in exp:
        delta = np.abs(exp[name] - P[name]).max()
        err = np.abs(got[name] - exp[name])
        err32 = np.abs(ref32.params()[name] - exp[name]).max()
        if opt == "SGD":
            assert err.max() <= max(4 * err32, 1e-4 * delta + 2e-6), (name, err.max(), err32, delta)
        else:
            assert np.median(err) <= 1e-5, (name, np.median(err))
            assert err.max() <= 2 * 0.01 * 3 + 1e-6, (name, err.max())
    # the chunked entry point takes the same path
    l2 = con.train_chunk_device(2, 0)
    assert np.all(np.isfinite(l2.cpu().numpy()))


def test_transr_relation_negatives_cannot_be_relation_sharded(built, small_ds):
    """A relation negative belongs to two relation shards: refused loudly, not computed wrongly."""
    from openkeonspark_b200 import OkbError
    con = _config(small_ds, "TransR", 20, 1, 1, "SGD")
    con.ctx.call("okb_transr_set_shard", 0, 5)
    con.sampling_device()
    with pytest.raises(OkbError):
        con.train_step_device(0)


@pytest.mark.parametrize("model,opt", [("TransE", "SGD"), ("TransH", "Adam"), ("TransD", "SGD"), ("TransR", "Adam")])
def test_short_segment_path_parity(built, wide_ds, model, opt):
    """Graph with many relations/entities relative to the batch: the update runs WITHOUT the hub pre-reduction
    (small_ds, with 23 Zipf relations, always takes the pre-reduced path)."""
    import torch
    from oracle import models_ref
    con = _config(wide_ds, model, 20, 1, 0, opt, nbatches=80)
    P = make_params(model, con.entTotal, con.relTotal, 20, seed=7)
    con.set_parameters(P)
    ref64 = models_ref.Trainer(model, P, margin=1.0, lr=0.01, opt=opt, dtype=torch.float64)
    con.sampling()
    h, t, r = con.batch_h.copy(), con.batch_t.copy(), con.batch_r.copy()
    loss = float(con.train_step_device(0).item())
    l64 = ref64.step(h, t, r, con.batch_size, 1, 0)
    assert abs(loss - l64) <= 2e-5 * abs(l64) + 1e-6
    got, exp = con.get_parameters(), ref64.params()
    for name in exp:
        if opt == "SGD":
            assert np.abs(got[name] - exp[name]).max() <= 1e-4 * np.abs(exp[name] - P[name]).max() + 2e-6, name
        else:
            _check_adam_slots(con, ref64)


@pytest.mark.parametrize("model,opt,D,k,kr,nbatches,ds", [
    ("TransH", "Adam", 100, 1, 0, 6, "small"), ("TransE", "SGD", 50, 3, 2, 6, "small"), ("TransD", "Adam", 33, 2, 1, 40, "small"),
    ("TransR", "Adam", 20, 2, 0, 6, "small"), ("TransE", "SGD", 16, 2, 0, 4, "tiny"), ("TransH", "SGD", 64, 1, 0, 3, "wide"),
    ("TransE", "Adam", 32, 10, 0, 1, "small")])
def test_single_kernel_plan_equals_multi_kernel_plan(built, small_ds, tiny_ds, wide_ds, model, opt, D, k, kr, nbatches, ds):
    """A one-step plan runs as ONE kernel (plan_small_kernel: keys, stable 2-pass sort, row map) when it fits one CTA;
    OKB_FLAG_PLAN_MULTI routes it through the general multi-kernel segmented sort.  Both are stable sorts of the same
    keys, so losses and tables must be bit-identical (the last case exceeds the one-CTA limit and checks the fallback)."""
    path = {"small": small_ds, "tiny": tiny_ds, "wide": wide_ds}[ds]
    outs = []
    for multi in (0, 1):
        con = _config(path, model, D, k, kr, opt, nbatches=nbatches)
        con.ctx.call("okb_set_flag", 9, multi)
        con.set_parameters(make_params(model, con.entTotal, con.relTotal, D, seed=5))
        losses = []
        for it in range(4):
            con.sampling_device()
            losses.append(float(con.train_step_device(0).item()))
        outs.append((losses, con.get_parameters()))
    assert outs[0][0] == outs[1][0]
    for name in outs[0][1]:
        assert np.array_equal(outs[0][1][name], outs[1][1][name]), name


@pytest.mark.parametrize("E,R,n_train,nbatches,k,kr", [(3000, 200, 9000, 3, 1, 0),      # 12-bit keys: 6-bit digits
                                                        (20000, 100, 12000, 3, 2, 1),    # 15-bit keys: 8-bit digits
                                                        (40000, 18, 20000, 4, 1, 0)])    # 16-bit keys (WN18-sized id space)
def test_single_kernel_plan_digit_widths(built, E, R, n_train, nbatches, k, kr):
    """The one-step plan kernel with the digit widths the file-based fixtures do not reach (they give 5 and 7 bits)."""
    import contextlib, io
    import openkeonspark_b200 as okb
    from openkeonspark_b200 import datagen
    g = datagen.make_graph(E, R, n_train, 200, 200, seed=E)
    outs = []
    for multi in (0, 1):
        con = okb.Config(private_context=True)
        con.set_nbatches(nbatches); con.set_ent_neg_rate(k); con.set_rel_neg_rate(kr); con.set_dimension(16)
        con.set_opt_method("SGD"); con.set_alpha(0.01); con.workThreads = 4
        with contextlib.redirect_stdout(io.StringIO()):
            con.init_from_arrays(g.E, g.R, g.train, g.valid, g.test)
        seeds = np.arange(1, 5, dtype=np.uint64) * np.uint64(2654435761)
        con.ctx.call("okb_set_streams", ctypes.c_void_p(seeds.ctypes.data), 4)
        con.set_model_and_session(okb.TransE)
        con.ctx.call("okb_set_flag", 9, multi)
        con.set_parameters(make_params("TransE", con.entTotal, con.relTotal, 16, seed=5))
        assert con.batch_size * (3 + k + kr) <= 32768
        losses = []
        for it in range(3):
            con.sampling_device()
            losses.append(float(con.train_step_device(0).item()))
        outs.append((losses, con.get_parameters()))
    assert outs[0][0] == outs[1][0]
    for name in outs[0][1]:
        assert np.array_equal(outs[0][1][name], outs[1][1][name]), name


@pytest.mark.parametrize("threads", ["512", "640"])
@pytest.mark.parametrize("model,opt,D,k,kr,ds,single_warp", [
    ("TransH", "Adam", 100, 1, 0, "small", 0),        # the bench workload's kernel pair: k = 1 grad body + dense Adam; hub rows (23 Zipf relations)
    ("TransE", "SGD", 50, 1, 0, "small", 0),          # 64-bit row fragments
    ("TransD", "Adam", 100, 1, 0, "wide", 0),         # two tables per entity, no hub rows
    ("TransE", "SGD", 200, 1, 0, "wide", 0),          # two 128-bit vectors per lane
    ("TransE", "Adam", 64, 1, 0, "small", 0)])
def test_persistent_chunk_kernel_equals_per_phase_kernels(built, small_ds, wide_ds, monkeypatch, model, opt, D, k, kr, ds, single_warp, threads):
    """OKB_FLAG_CHUNK_KERNEL = 1: okb_train_steps runs a chunk as ONE persistent cooperative kernel (csrc/chunk.cu: grad ->
    grid barrier -> update per step) where covered (the k = 1 batch); the default is the per-phase kernels (measured
    faster).  Same bodies, same fp32 order: losses and tables must be bit-identical, for both CTA sizes."""
    monkeypatch.setenv("OKB200_CHUNK_THREADS", threads)
    path = {"small": small_ds, "wide": wide_ds}[ds]
    outs = []
    for chunk_kernel in (0, 1):
        con = _config(path, model, D, k, kr, opt, nbatches=5)
        con.ctx.call("okb_set_flag", 10, chunk_kernel)
        con.ctx.call("okb_set_flag", 8, single_warp)
        con.set_parameters(make_params(model, con.entTotal, con.relTotal, D, seed=5))
        from openkeonspark_b200 import _native
        l0 = _native.load().okb_launch_count()
        losses = []
        for n in (7, 2, 5):
            losses += [float(x) for x in con.train_chunk_device(n, 0).cpu().numpy()]
        outs.append((losses, con.get_parameters(), _native.load().okb_launch_count() - l0))
    assert outs[0][0] == outs[1][0], (outs[0][0], outs[1][0])
    for name in outs[0][1]:
        assert np.array_equal(outs[0][1][name], outs[1][1][name]), name
    assert outs[1][2] < outs[0][2] - 2 * 10, (outs[0][2], outs[1][2])       # 14 steps: >= 28 grad/update launches became 3


@pytest.mark.parametrize("model,D,ds", [("TransH", 100, "small"), ("TransD", 50, "wide"), ("TransE", 33, "small")])
def test_adam_vectors_per_thread_bit_identical(built, small_ds, wide_ds, model, D, ds):
    """OKB_FLAG_ADAM_VPT: 1..4 vectors per thread in the dense Adam pass (the gradient sums and the Adam rule are the same
    functions, only the thread -> vector mapping changes): losses, tables and slots bit-identical."""
    path = {"small": small_ds, "wide": wide_ds}[ds]
    outs = []
    for vpt in (1, 2, 3, 4):
        con = _config(path, model, D, 2, 0, "Adam", nbatches=5)
        con.ctx.call("okb_set_flag", 11, vpt)
        con.set_parameters(make_params(model, con.entTotal, con.relTotal, D, seed=5))
        losses = [float(x) for x in con.train_chunk_device(5, 0).cpu().numpy()]
        slots = {k: v.cpu().numpy() for k, v in con._adam.items() if hasattr(v, "cpu")}
        outs.append((losses, con.get_parameters(), slots))
    for o in outs[1:]:
        assert o[0] == outs[0][0]
        for name in outs[0][1]:
            assert np.array_equal(o[1][name], outs[0][1][name]), name
        for name in outs[0][2]:
            assert np.array_equal(o[2][name], outs[0][2][name]), name


@pytest.mark.parametrize("opt", ["SGD", "Adam"])
@pytest.mark.parametrize("D,Dr,k,ds", [(100, 100, 1, "small"), (40, 24, 3, "small"), (64, 128, 2, "wide"), (20, 20, 1, "wide")])
def test_transr_fused_kernel_equals_two_kernel_form(built, small_ds, wide_ds, opt, D, Dr, k, ds):
    """OKB_FLAG_TRANSR_FUSED (an A/B experiment, default off): persistent kernel, M_r double-buffered by bulk-async copies, relation-side update
    applied in place — against the round-1 form (one CTA per relation + a separate relation-update kernel through the
    per-relation gradient rows).  Same contractions in the same order: losses equal, tables within 1e-7 (the update is the
    same expression compiled in two kernels)."""
    path = {"small": small_ds, "wide": wide_ds}[ds]
    outs = []
    for fused in (0, 1):
        con = _config(path, "TransR", D, k, 0, opt, Dr=Dr, nbatches=5)
        con.ctx.call("okb_set_flag", 12, fused)
        con.set_parameters(make_params("TransR", con.entTotal, con.relTotal, D, seed=9, Dr=Dr))
        losses = [float(x) for x in con.train_chunk_device(4, 0).cpu().numpy()]
        con.sampling_device()
        losses.append(float(con.train_step_device(0).item()))
        slots = {} if con._adam is None else {kk: v.cpu().numpy() for kk, v in con._adam.items() if hasattr(v, "cpu")}
        outs.append((losses, con.get_parameters(), slots))
    assert np.allclose(outs[0][0], outs[1][0], rtol=1e-6, atol=0), (outs[0][0], outs[1][0])
    for name in outs[0][1]:
        assert np.allclose(outs[0][1][name], outs[1][1][name], rtol=0, atol=(1e-7 if opt == "SGD" else 2e-2 * 0 + 1e-6)), name
    for name in outs[0][2]:
        assert np.allclose(outs[0][2][name], outs[1][2][name], rtol=1e-5, atol=1e-9), name
