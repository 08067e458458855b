"""Synchronous data-parallel training and candidate-sharded evaluation inside one box.

Replaces the reference's asynchronous TensorFlowOnSpark parameter-server path
(distribute_training.py:161-364: between-graph replication, variables on /job:ps, gRPC) with one
process per GPU over torch.distributed (NCCL over NVLink):

  train  The global batch IS the reference batch at workThreads = W (W a multiple of the world size);
         every rank samples the full batch (integer work, bit-identical everywhere) and plans it, but
         computes gradients only for the positives of ITS streams [g*W/G, (g+1)*W/G) (Base.cpp:85-92
         slice geometry).  One all-gather of the gradient rows, then every rank applies the same
         sorted, fixed-order update -> replicas stay bit-identical without a parameter broadcast.
  train  (owner-sharded, the default on GPUs for TransE/H/D) every rank keeps the full tables in a peer arena the other
         ranks map over NVLink (CUDA IPC) and OWNS the update of a contiguous range of rows; the reduce-scatter and the
         all-gather of a synchronous step are plain peer stores inside the step's own kernels (csrc/train.cu, "data
         parallel, owner-sharded"), no NCCL call per step.  Forms (`form=`):
           scatter (default up to 4 ranks)  every rank samples + plans the GLOBAL batch; the grad kernel stores each
                   gradient row straight into its row owner's arena; the owner runs the single-GPU update over its rows
                   and stores the new rows into every rank's table: two kernels per step, losses / tables / records
                   bit-identical to ONE GPU training the global batch;
           push    (default above 4 ranks)  every rank samples + plans only ITS positives and pushes per-row partial
                   sums into the owner's staging slab; three kernels per step; replicas bit-identical to each other,
                   against the single-GPU order the sums differ by fp32 re-association ((a+b)+(c+d)): a tolerance;
           gather / pull  measured-slower alternatives kept for A/B runs.
         `mode="exact"` keeps the NCCL all-gather path below (CPU / gloo tests).
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
    mode = "exact"

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


class RelationSharded(DataParallel):
    """TransR: ranks split the RELATIONS, not the positives (csrc/transr.cu).

    TransR's train kernel owns one relation per CTA and emits that relation's matrix gradient already reduced, so the
    natural shard is a contiguous range of relations per rank: a rank computes the positives of its relations and is the
    only one that ever updates (or, during training, reads) those relations' rel_embeddings / transfer_matrix rows — the
    40 KB-per-relation operand never crosses NVLink.  Per step the ranks exchange only the entity gradient rows and loss
    terms (each row is written by exactly one rank, so the all-reduce adds zeros: results are bit-identical to one GPU)
    and apply the same entity update.  The relation tables are gathered from their owners when something reads them
    (get_parameters, checkpoints, evaluation) — those calls are collectives in this mode."""
    mode = "relation"

    def __init__(self, con, group=None):
        super().__init__(con, group)
        con._ensure_model()
        if con.trainModel.name != "TransR":
            raise ValueError("relation-sharded mode is TransR's; use mode='owner' or 'exact'")
        if con.negative_rel:
            raise ValueError("TransR data parallelism needs rel_neg_rate == 0 (a relation negative belongs to two shards)")
        R = con.relTotal
        self.rel_ranges = owner_rows(R, self.world)
        lo, hi = self.rel_ranges[self.rank]
        if hi <= lo:
            raise ValueError("more ranks than relations")
        con.ctx.call("okb_transr_set_shard", lo, hi)
        self._stale = False          # the relation rows of the other shards are out of date

    def train_step(self, con, m, hp, step):
        from .Config import _stream
        b = self._buffers(con, m)
        s = _stream()
        con.ctx.call("okb_plan", step, s)
        b["gent"].zero_(); b["loss"].zero_()
        con.ctx.call("okb_grad", ctypes.byref(m), ctypes.byref(hp), step, 0, con.batch_size, _vp(b["gent"].data_ptr()),
                     _vp(b["grel"].data_ptr()), _vp(b["loss"].data_ptr()), s)
        dist.all_reduce(b["gent"], group=self.group)
        dist.all_reduce(b["loss"], group=self.group)
        con.ctx.call("okb_update", ctypes.byref(m), ctypes.byref(hp), step, _vp(b["gent"].data_ptr()),
                     _vp(b["grel"].data_ptr()), _vp(b["loss"].data_ptr()), _vp(con._loss_dev.data_ptr()), s)
        self._stale = True

    def _buffers(self, con, m):
        if self._bufs is None:
            er, ec, rr, rc = (ctypes.c_int64() for _ in range(4))
            con.ctx.call("okb_grad_sizes", ctypes.byref(m), con.batch_size, con.negative_ent, con.negative_rel,
                         ctypes.byref(er), ctypes.byref(ec), ctypes.byref(rr), ctypes.byref(rc))
            dev = con.trainModel.device
            self._bufs = dict(gent=torch.zeros(er.value, ec.value, dtype=torch.float32, device=dev),
                              grel=torch.zeros(rr.value, rc.value, dtype=torch.float32, device=dev),       # one row per RELATION
                              loss=torch.zeros(con.batch_size, dtype=torch.float32, device=dev))
        return self._bufs

    def gather_relations(self, con, adam=False):
        """Collective: every rank receives the other shards' relation rows (and, adam=True, their Adam slots)."""
        if not self._stale and not adam:
            return
        R = con.relTotal
        per = (R + self.world - 1) // self.world
        names = ["rel_embeddings", "transfer_matrix"]
        tensors = [con.trainModel.parameter_lists[n] for n in names]
        if adam and con._adam is not None:
            tensors += [con._adam[p + n] for n in names for p in ("m_", "v_")]
        lo, hi = self.rel_ranges[self.rank]
        for t in tensors:
            pad = torch.zeros(per * self.world, t.shape[1], dtype=t.dtype, device=t.device)
            mine = torch.zeros(per, t.shape[1], dtype=t.dtype, device=t.device)
            mine[:hi - lo].copy_(t[lo:hi])
            dist.all_gather_into_tensor(pad, mine, group=self.group)
            t.copy_(pad[:R])
        self._stale = False

    def link_prediction(self, con, q_lo=0, q_hi=None):
        self.gather_relations(con)
        return super().link_prediction(con, q_lo, q_hi)


def owner_rows(n_rows, world):
    """Row ranges owned by each rank in owner-sharded mode: blocks of ceil(n_rows / world) (csrc/train.cu okb_dp_*)."""
    per = (n_rows + world - 1) // world
    return [(min(n_rows, g * per), min(n_rows, (g + 1) * per)) for g in range(world)]


def default_form(world, pull=False, form=None):
    """Form of the owner-sharded step: an explicit `form` wins, then the A/B switches (`pull`, OKB200_DP_FORM), then the
    measured default by world size — scatter up to 4 ranks, push above (profiles/r02_dp_phase_traces.txt)."""
    import os
    if form is None:
        form = os.environ.get("OKB200_DP_FORM", "scatter" if world <= 4 else "push")
    if pull:
        form = "pull"
    if form not in ("gather", "scatter", "push", "pull"):
        raise ValueError("form must be 'gather', 'scatter', 'push' or 'pull'")
    return form


class _ArenaView:
    """__cuda_array_interface__ over raw device memory, so torch can wrap a slice of the peer arena."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2,
                                         "strides": None}


class OwnerSharded(DataParallel):
    """Owner-sharded synchronous data parallelism over peer memory (see the module docstring)."""
    mode = "owner"

    def __init__(self, con, group=None, pull=None, form=None):
        super().__init__(con, group)
        import os
        pull = os.environ.get("OKB200_DP_PULL") == "1" if pull is None else pull      # measured slower: A/B runs only
        # form: "scatter" = the grad kernel stores gradient rows straight into their row owner's arena (two kernels per
        # step, bit-identical to one GPU); "push" = per-rank partial sums through staging slabs (three kernels per step)
        # default by world size: measured crossover (profiles/r02_dp_phase_traces.txt) — the scatter form's grad kernel slows
        # down as the share of remote rows grows and its owner update pre-reduces the GLOBAL batch's long segments
        # "gather": gradient rows go to EVERY rank and every rank runs the full update — one exchange per step, bit-identical
        # to one GPU too, but measured slower even at 2 ranks (61.6 vs 53.8 us per step: the grad kernel's peer stores run at
        # ~350 GB/s, so doubling them costs more than the second exchange); kept as an option
        form = default_form(self.world, pull=pull, form=form)
        self.form = form
        self.prefetch = os.environ.get("OKB200_DP_PREFETCH", "1") == "1" and form != "pull"
        from ._native import okb_dp
        from .Config import _AUX_ENT, _AUX_REL
        con._ensure_model()
        if con.trainModel.name == "TransR":
            raise ValueError("owner-sharded mode does not cover TransR (its gradients are reduced per relation): "
                             "use mode='relation' (parallel.attach picks it for TransR)")
        if self.world > 16:
            raise ValueError("owner-sharded mode supports up to 16 ranks per box")
        m = con._cmodel()
        lay = okb_dp()
        if pull:   # reserve the "pull" slices: the plan of up to plan_ahead steps, this rank's gradient rows and loss terms
            lay.plan_steps, lay.max_local = int(con.plan_ahead), int(self.chunk)
            lay.neg_ent, lay.neg_rel = int(con.negative_ent), int(con.negative_rel)
            con.ctx.call("okb_set_flag", 7, 1)
        if form in ("scatter", "gather"):
            lay.scatter, lay.global_batch = (1 if form == "scatter" else 2), int(con.batch_size)
            lay.neg_ent, lay.neg_rel = int(con.negative_ent), int(con.negative_rel)
        con.ctx.call("okb_dp_layout", ctypes.byref(m), self.world, ctypes.byref(lay))
        own, handle = _vp(), (ctypes.c_ubyte * 64)()
        con.ctx.call("okb_peer_alloc", lay.arena_bytes, ctypes.byref(own), handle)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=self.group)
        self._opened = []
        for q in range(self.world):
            if q == self.rank:
                lay.arena[q] = own.value
            else:
                peer = _vp()
                con.ctx.call("okb_peer_open", (ctypes.c_ubyte * 64).from_buffer_copy(handles[q]), ctypes.byref(peer))
                lay.arena[q] = peer.value
                self._opened.append(peer.value)
        # re-home the tables into the arena (values preserved); the Adam slots stay ordinary local tensors
        name = con.trainModel.name
        offs = {"ent_embeddings": lay.off_ent, "rel_embeddings": lay.off_rel}
        if _AUX_ENT.get(name):
            offs[_AUX_ENT[name]] = lay.off_ent_aux
        if _AUX_REL.get(name):
            offs[_AUX_REL[name]] = lay.off_rel_aux
        P = con.trainModel.parameter_lists
        dev = con.trainModel.device
        self._views = []
        for k, off in offs.items():
            view = _ArenaView(own.value + off, P[k].shape)
            t = torch.as_tensor(view, device=dev)
            t.copy_(P[k])
            P[k] = t
            self._views.append(view)
        con._model_struct = None
        lay.rank, lay.world = self.rank, self.world
        lay.b_lo, lay.b_hi = self.ranges[self.rank]
        self._lay, self._own = lay, own.value
        per = con.workThreads // self.world
        # scatter form: every rank samples (and plans) the global batch; the other forms only this rank's streams
        self.streams = (0, con.workThreads) if form in ("scatter", "gather") else (self.rank * per, (self.rank + 1) * per)
        con.ctx.call("okb_dp_attach", ctypes.byref(lay))
        torch.cuda.synchronize()
        dist.barrier(group=self.group)                # every arena initialised before anyone pushes into it

    def _sample(self, con, n):
        from .Config import _stream
        con.ctx.call("okb_sample", con.batch_size, con.negative_ent, con.negative_rel, n, self.streams[0], self.streams[1], _stream())
        con._chunk_pos, con._chunk_len = 0, n

    def next_step(self, con):
        """One train step; batches are sampled (this rank's streams only) and planned plan_ahead steps at a time."""
        from .Config import _stream
        if con._chunk_pos >= con._chunk_len:
            n = con.batch_size * (3 + con.negative_ent + con.negative_rel) // (1 if self.form in ("scatter", "gather") else self.world)
            self._sample(con, max(1, min(int(con.plan_ahead), (1 << 24) // max(n, 1))))
        m = con._cmodel()
        (hp,), powers = con._hypers(1)
        con.ctx.call("okb_dp_train_steps", ctypes.byref(m), ctypes.byref(hp), con._chunk_pos, 1, _vp(con._loss_dev.data_ptr()), _stream())
        con._chunk_pos += 1
        con._commit_powers(powers, 1)
        return con._loss_dev

    def train_chunk(self, con, n):
        """n steps in ONE library call (sample + plan + n x [grad, reduce/push, owner update])."""
        from ._native import okb_hyper
        from .Config import _stream
        # chunk pipeline: this chunk is the one the side stream produced during the previous call (else it is produced now),
        # and the next one is started before this chunk's steps are issued, so it is sampled + planned under them
        geo = (con.batch_size, con.negative_ent, con.negative_rel, n, self.streams[0], self.streams[1], _stream())
        con.ctx.call("okb_dp_chunk_begin", *geo)
        if self.prefetch:
            con.ctx.call("okb_dp_chunk_prefetch", *geo)
        con._chunk_pos, con._chunk_len = 0, n
        m = con._cmodel()
        hl, powers = con._hypers(n)
        hps = (okb_hyper * n)(*hl)
        if getattr(con, "_loss_chunk", None) is None or con._loss_chunk.numel() < n:
            con._loss_chunk = torch.zeros(n, dtype=torch.float32, device=con.trainModel.device)
        con.ctx.call("okb_dp_train_steps", ctypes.byref(m), hps, 0, n, _vp(con._loss_chunk.data_ptr()), _stream())
        con._commit_powers(powers, n)
        con._chunk_pos = con._chunk_len = 0
        return con._loss_chunk[:n]

    def train_step(self, con, m, hp, step):
        raise RuntimeError("owner-sharded mode steps through next_step()/train_chunk()")

    def quiesce(self, con):
        """Returns when every rank's row updates of all steps issued so far have landed in THIS rank's tables (a rank's
        last owner-update kernel stores into its peers' arenas).  NOT a collective — `if rank == 0: con.get_parameters()`
        is fine: it waits on flags the peers publish at the end of their own train calls."""
        from .Config import _stream
        con.ctx.call("okb_dp_quiesce", _stream())
        torch.cuda.synchronize()

    def link_prediction(self, con, q_lo=0, q_hi=None):
        self.quiesce(con)
        return super().link_prediction(con, q_lo, q_hi)

    def sync_adam_slots(self, con):
        """Make every rank's Adam m / v complete (each rank only maintains the rows it owns): for checkpoints."""
        if con._adam is None:
            return
        for k, v in con._adam.items():
            if not torch.is_tensor(v):
                continue
            rows = v.shape[0]
            per = (rows + self.world - 1) // self.world
            pad = torch.zeros(per * self.world, *v.shape[1:], dtype=v.dtype, device=v.device)
            lo, hi = owner_rows(rows, self.world)[self.rank]
            dist.all_gather_into_tensor(pad, torch.nn.functional.pad(v[lo:hi], (0, 0, 0, per - (hi - lo))), group=self.group)
            v.copy_(pad[:rows])

    def close(self, con):
        """Collective: every rank must have stopped writing into its peers before the mappings go away."""
        self.quiesce(con)
        dist.barrier(group=self.group)
        con.ctx.call("okb_dp_detach")
        for p in self._opened:
            con.ctx.call("okb_peer_close", _vp(p))
        self._opened = []


def attach(con, group=None, mode="auto", pull=None, form=None):
    """Enable data-parallel mode on a Config whose init() has run (torch.distributed initialised).
    mode: "owner" (peer-memory owner-sharded update), "exact" (all-gather of gradient rows, bit-identical to one GPU),
    "auto" = owner on NCCL/GPU for TransE/H/D when the batch touches a sizeable share of the rows, else exact."""
    con._ensure_model()
    if con.trainModel.name == "TransR":
        if mode not in ("auto", "relation"):
            raise ValueError("TransR trains data-parallel in mode='relation' only (its gradients are reduced per relation)")
        con._world = RelationSharded(con, group)
        return con._world
    if mode == "relation":
        raise ValueError("mode='relation' is TransR's")
    if mode == "auto":
        mode = "exact"
        if dist.get_backend(group) == "nccl" and dist.get_world_size(group) <= 16:      # peer memory needs GPUs
            con._ensure_model()
            dense = con.batch_size * (3 + con.negative_ent + con.negative_rel) * 4 >= con.entTotal + con.relTotal
            _, ranges = partition(con.batch_size, con.workThreads, dist.get_world_size(group))
            if dense and con.trainModel.name != "TransR" and all(hi > lo for lo, hi in ranges):    # every rank must own positives
                mode = "owner"
    con._world = OwnerSharded(con, group, pull, form) if mode == "owner" else DataParallel(con, group)
    return con._world
