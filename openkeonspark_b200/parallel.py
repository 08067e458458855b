"""Synchronous data-parallel training and candidate-sharded evaluation inside one box.

Replaces the reference's asynchronous TensorFlowOnSpark parameter-server path
(distribute_training.py:161-364: between-graph replication, variables on /job:ps, gRPC) with one
process per GPU over torch.distributed (NCCL over NVLink):

  train  The global batch IS the reference batch at workThreads = W (W a multiple of the world size);
         every rank samples the full batch (integer work, bit-identical everywhere) and plans it, but
         computes gradients only for the positives of ITS streams [g*W/G, (g+1)*W/G) (Base.cpp:85-92
         slice geometry).  One all-gather of the gradient rows, then every rank applies the same
         sorted, fixed-order update -> replicas stay bit-identical without a parameter broadcast.
  eval   candidate entities are split into G contiguous ranges; tables are replicated, so every rank
         computes each query's reference score bit-identically, counts better candidates in its range,
         and the integer counts are all-reduced (sum) / the packed argmins all-reduced (min).
         (The reference splits QUERIES across workers instead, distribute_training.py:430-441.)
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

_vp = ctypes.c_void_p
_I64_MAX = 0x7FFFFFFFFFFFFFFF


def partition(batch_size, work_threads, world):
    """Positives owned by each rank: rank g owns the slices of streams [g*W/G, (g+1)*W/G).

    Returns (chunk, [(lo, hi)] per rank); chunk = positives per rank before clamping, so rank g's
    range starts at g*chunk (gradient buffers are laid out in chunks of this size)."""
    if work_threads % world != 0:
        raise ValueError("workThreads (%d) must be a multiple of the world size (%d)" % (work_threads, world))
    per = batch_size // work_threads + (1 if batch_size % work_threads else 0)     # Base.cpp:85-92
    chunk = per * (work_threads // world)
    return chunk, [(min(batch_size, g * chunk), min(batch_size, (g + 1) * chunk)) for g in range(world)]


def candidate_range(n_ent, world, rank):
    lo = (n_ent * rank) // world
    hi = (n_ent * (rank + 1)) // world
    return lo, hi


def allreduce_best(best, group=None):
    """min-combine packed (score_bits << 32 | id) words where all-ones means "no candidate".
    Scores are non-negative so valid words are positive int64; the sentinel (-1) is lifted to
    INT64_MAX for the MIN reduction and restored afterwards."""
    lifted = torch.where(best < 0, torch.full_like(best, _I64_MAX), best)
    dist.all_reduce(lifted, op=dist.ReduceOp.MIN, group=group)
    return torch.where(lifted == _I64_MAX, torch.full_like(lifted, -1), lifted)


class DataParallel:
    def __init__(self, con, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.chunk, self.ranges = partition(con.batch_size, con.workThreads, self.world)
        self._bufs = None
        self._coalesce = dist.get_backend(group) == "nccl"

    def _buffers(self, con, m):
        if self._bufs is not None:
            return self._bufs
        er, ec, rr, rc = (ctypes.c_int64() for _ in range(4))
        con.ctx.call("okb_grad_sizes", ctypes.byref(m), con.batch_size, con.negative_ent, con.negative_rel,
                     ctypes.byref(er), ctypes.byref(ec), ctypes.byref(rr), ctypes.byref(rc))
        ne, nr = er.value // con.batch_size, rr.value // con.batch_size
        dev = con.trainModel.device
        padded = self.chunk * self.world
        self._bufs = dict(
            gent=torch.zeros(padded * ne, ec.value, dtype=torch.float32, device=dev),
            grel=torch.zeros(padded * nr, rc.value, dtype=torch.float32, device=dev),
            loss=torch.zeros(padded, dtype=torch.float32, device=dev), ne=ne, nr=nr)
        return self._bufs

    def train_step(self, con, m, hp, step):
        from .Config import _stream
        b = self._buffers(con, m)
        lo, hi = self.ranges[self.rank]
        s = _stream()
        con.ctx.call("okb_plan", step, s)
        con.ctx.call("okb_grad", ctypes.byref(m), ctypes.byref(hp), step, lo, hi, _vp(b["gent"].data_ptr()),
                     _vp(b["grel"].data_ptr()), _vp(b["loss"].data_ptr()), s)
        c, g = self.chunk, self.rank
        # in-place all-gather: each rank's slot of the full buffer is its own contribution; the three gathers are
        # issued as ONE NCCL group (one launch, one synchronisation over NVLink) when the backend supports it
        parts = [(b["gent"], b["gent"][g * c * b["ne"]:(g + 1) * c * b["ne"]]),
                 (b["grel"], b["grel"][g * c * b["nr"]:(g + 1) * c * b["nr"]]),
                 (b["loss"], b["loss"][g * c:(g + 1) * c])]
        if self._coalesce:
            try:
                with dist._coalescing_manager(group=self.group, device=b["gent"].device, async_ops=False):
                    for full, mine in parts:
                        dist.all_gather_into_tensor(full, mine, group=self.group)
            except Exception:                      # backend without coalesced all-gather (gloo): plain calls
                self._coalesce = False
                for full, mine in parts:
                    dist.all_gather_into_tensor(full, mine, group=self.group)
        else:
            for full, mine in parts:
                dist.all_gather_into_tensor(full, mine, group=self.group)
        con.ctx.call("okb_update", ctypes.byref(m), ctypes.byref(hp), step, _vp(b["gent"].data_ptr()),
                     _vp(b["grel"].data_ptr()), _vp(b["loss"].data_ptr()), _vp(con._loss_dev.data_ptr()), s)

    def link_prediction(self, con, q_lo=0, q_hi=None):
        lo, hi = candidate_range(con.entTotal, self.world, self.rank)

        def reduce_fn(counts, best):
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=self.group)
            return counts, allreduce_best(best, self.group)

        return con.link_prediction_records(q_lo, q_hi, lo, hi, reduce_fn)


def attach(con, group=None):
    """Enable data-parallel mode on a Config whose init() has run (torch.distributed initialised)."""
    con._world = DataParallel(con, group)
    return con._world
