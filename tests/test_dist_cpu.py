"""Multi-process host logic on CPU (gloo, world size 2): slice geometry, the in-place all-gather layout
of the gradient buffers, and the count / packed-argmin reductions of candidate-sharded evaluation.
The CUDA entry points are replaced by a recording fake; the real kernels are covered by -m gpu tests."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_geometry():
    from openkeonspark_b200 import parallel
    # Base.cpp:85-92: B % W == 0 -> slices of B/W; else B/W+1 with the tail clamped
    chunk, rng = parallel.partition(4831, 8, 8)
    assert chunk == 604 and rng[0] == (0, 604) and rng[7] == (4228, 4831)
    chunk, rng = parallel.partition(4831, 8, 2)
    assert chunk == 2416 and rng == [(0, 2416), (2416, 4831)]
    chunk, rng = parallel.partition(4800, 8, 4)
    assert chunk == 1200 and rng[3] == (3600, 4800)
    chunk, rng = parallel.partition(3, 8, 2)          # later streams own empty slices
    assert rng == [(0, 3), (3, 3)]
    with pytest.raises(ValueError):
        parallel.partition(100, 6, 4)
    los = [parallel.candidate_range(14951, 8, r) for r in range(8)]
    assert los[0][0] == 0 and los[-1][1] == 14951 and all(a[1] == b[0] for a, b in zip(los[:-1], los[1:]))


def test_owner_rows_geometry():
    """Row ownership of the owner-sharded mode mirrors csrc/train.cu: blocks of ceil(rows / world), tail clamped."""
    from openkeonspark_b200 import parallel
    assert parallel.owner_rows(14951, 8)[0] == (0, 1869) and parallel.owner_rows(14951, 8)[7] == (13083, 14951)
    assert parallel.owner_rows(18, 8) == [(0, 3), (3, 6), (6, 9), (9, 12), (12, 15), (15, 18), (18, 18), (18, 18)]
    for rows, world in ((1345, 2), (40943, 4), (5, 16)):
        r = parallel.owner_rows(rows, world)
        assert r[0][0] == 0 and r[-1][1] == rows and all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))


class FakeCtx:
    """Stands in for _native.Ctx: fills this rank's rows exactly where okb_grad would."""

    def __init__(self, con):
        self.con, self.calls = con, []

    def call(self, name, *args):
        self.calls.append(name)
        con = self.con
        if name == "okb_grad_sizes":
            B, k, kr = args[1], args[2], args[3]
            args[4]._obj.value, args[5]._obj.value = B * (2 + k), con.D
            args[6]._obj.value, args[7]._obj.value = B * (1 + kr), 2 * con.D
        elif name == "okb_transr_set_shard":
            con.shard = (args[0], args[1])
        elif name == "okb_grad" and con.trainModel.name == "TransR":
            # relation-sharded: the full positive range is passed, only the positives of this rank's relations are written
            assert (args[3], args[4]) == (0, con.batch_size)
            b = con._world._bufs
            for bb in range(con.batch_size):
                r = bb % con.relTotal
                if con.shard[0] <= r < con.shard[1]:
                    b["gent"][bb * 3:(bb + 1) * 3] = float(bb + 1)
                    b["loss"][bb] = 0.5 * bb
            b["grel"][con.shard[0]:con.shard[1]] = float(con._world.rank + 1)
        elif name == "okb_grad":
            lo, hi = args[3], args[4]
            b = con._world._bufs
            ne, nr = b["ne"], b["nr"]
            for bb in range(lo, hi):
                b["gent"][bb * ne:(bb + 1) * ne] = float(bb + 1)
                b["grel"][bb * nr:(bb + 1) * nr] = -float(bb + 1)
                b["loss"][bb] = 0.5 * bb
        elif name == "okb_update":
            con.updated = {k: v.clone() if torch.is_tensor(v) else v for k, v in con._world._bufs.items()}


class FakeModel:
    device = torch.device("cpu")
    name = "TransE"
    parameter_lists = {}


class FakeCon:
    def __init__(self, B, W, D, k, kr):
        self.batch_size, self.workThreads, self.D = B, W, D
        self.negative_ent, self.negative_rel = k, kr
        self.trainModel = FakeModel()
        self._loss_dev = torch.zeros(1)
        self.ctx = FakeCtx(self)
        self._adam = None

    def _ensure_model(self):
        pass


def _worker(rank, world, port, B, W):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from openkeonspark_b200 import parallel
    sys.modules["openkeonspark_b200.Config"]._stream = lambda: None     # no CUDA stream on the CPU box
    con = FakeCon(B, W, 4, 2, 1)
    dp = parallel.attach(con)
    assert dp.mode == "exact"                        # gloo / CPU: the peer-memory mode is never chosen
    from openkeonspark_b200._native import okb_hyper, okb_model
    dp.train_step(con, okb_model(), okb_hyper(), 0)
    assert con.ctx.calls == ["okb_grad_sizes", "okb_plan", "okb_grad", "okb_update"]
    u = con.updated
    # after the all-gather EVERY rank holds every positive's rows, at the positive's own offset
    for bb in range(B):
        assert torch.all(u["gent"][bb * 4:(bb + 1) * 4] == float(bb + 1))
        assert torch.all(u["grel"][bb * 2:(bb + 1) * 2] == -float(bb + 1))
        assert u["loss"][bb] == 0.5 * bb
    # candidate-sharded evaluation: counts add up, packed argmins min-combine, sentinel survives
    counts = torch.tensor([rank + 1, 0, 5], dtype=torch.int64)
    best = torch.tensor([-1, (3 << 32) | (7 + rank), -1 if rank == 0 else (1 << 32) | 9], dtype=torch.int64)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    best = parallel.allreduce_best(best)
    assert counts.tolist() == [3, 0, 10]
    assert best.tolist() == [-1, (3 << 32) | 7, (1 << 32) | 9]
    dist.destroy_process_group()


def _worker_relation(rank, world, port, B, R):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from openkeonspark_b200 import parallel
    sys.modules["openkeonspark_b200.Config"]._stream = lambda: None
    con = FakeCon(B, 4, 4, 1, 0)
    con.trainModel = FakeModel()
    con.trainModel.name = "TransR"
    con.relTotal = R
    # relation tables whose rows identify their owner: the gather must deliver every owner's rows to every rank
    rel = torch.full((R, 3), -1.0)
    mat = torch.full((R, 6), -1.0)
    con.trainModel.parameter_lists = {"rel_embeddings": rel, "transfer_matrix": mat}
    orig_sizes = FakeCtx.call

    def sizes(self, name, *args):
        if name == "okb_grad_sizes":
            args[4]._obj.value, args[5]._obj.value, args[6]._obj.value, args[7]._obj.value = B * 3, 4, R, 4 + 16
            self.calls.append(name)
        else:
            orig_sizes(self, name, *args)
    FakeCtx.call = sizes
    dp = parallel.attach(con)
    assert dp.mode == "relation" and con.shard == parallel.owner_rows(R, world)[rank]
    with pytest.raises(ValueError):
        parallel.attach(con, mode="exact")           # TransR has one data-parallel mode
    from openkeonspark_b200._native import okb_hyper, okb_model
    dp.train_step(con, okb_model(), okb_hyper(), 0)
    assert con.ctx.calls == ["okb_transr_set_shard", "okb_grad_sizes", "okb_plan", "okb_grad", "okb_update"]
    u = con.updated
    for bb in range(B):                              # every positive's rows arrive exactly once (one writer + zeros)
        assert torch.all(u["gent"][bb * 3:(bb + 1) * 3] == float(bb + 1)) and u["loss"][bb] == 0.5 * bb
    lo, hi = con.shard
    rel[lo:hi] = float(rank + 10)
    mat[lo:hi] = float(rank + 20)
    dp._stale = True
    dp.gather_relations(con)
    for q, (a, b) in enumerate(parallel.owner_rows(R, world)):
        assert torch.all(rel[a:b] == float(q + 10)) and torch.all(mat[a:b] == float(q + 20))
    dist.destroy_process_group()


@pytest.mark.parametrize("B,R", [(37, 7), (16, 2)])
def test_relation_sharded_host_logic_gloo_world2(B, R):
    port = 31500 + (os.getpid() + B) % 2000
    mp.spawn(_worker_relation, args=(2, port, B, R), nprocs=2, join=True)


@pytest.mark.parametrize("B,W", [(37, 4), (40, 8), (3, 8)])
def test_data_parallel_host_logic_gloo_world2(B, W):
    port = 29500 + (os.getpid() + B) % 2000
    mp.spawn(_worker, args=(2, port, B, W), nprocs=2, join=True)


def test_default_form_of_the_owner_sharded_step(monkeypatch):
    """scatter up to 4 ranks, push above; an explicit form or the A/B switches override; unknown names are refused."""
    from openkeonspark_b200 import parallel
    monkeypatch.delenv("OKB200_DP_FORM", raising=False)
    assert [parallel.default_form(w) for w in (1, 2, 4, 5, 8, 16)] == ["scatter", "scatter", "scatter", "push", "push", "push"]
    assert parallel.default_form(8, form="scatter") == "scatter" and parallel.default_form(2, form="gather") == "gather"
    assert parallel.default_form(2, pull=True) == "pull"
    monkeypatch.setenv("OKB200_DP_FORM", "push")
    assert parallel.default_form(2) == "push" and parallel.default_form(2, form="scatter") == "scatter"
    with pytest.raises(ValueError):
        parallel.default_form(2, form="ring")
