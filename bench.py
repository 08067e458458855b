#!/usr/bin/env python
"""bench.py — the reference's headline metric (train triples/s + link-prediction queries/s) on B200.

    python bench.py [--gpus N --steps K --warmup W]              this repo's CUDA path
    python bench.py --impl reference [--steps K --warmup W]      the reference's CPU path, same config

Headline workload (BASELINE.json configs[1]): TransH dim=100, Adam, margin 1, ent_neg_rate 1, FB15K-shaped
synthetic graph (14,951 entities / 1,345 relations / 483,142 train triples), nbatches=100 so B = 4,831
positives per step per GPU (weak scaling: the global batch is N x 4,831 = the reference batch at
workThreads = 8N).  One "step" = sampling() + loss_def + optimizer (distribute_training.py:274-282):
GPU sampler -> plan (keys + radix sort) -> fused grad kernel -> fused segmented-reduce + TF1-Adam update.
The sampler and the plan run INSIDE the timed region whatever --steps is (chunks of min(64, steps) steps on the
launching stream; the run fails if the instrumented pass saw no sampler or plan launch).

The one JSON line also carries `configs`: train triples/s, link-prediction queries/s and roofline fractions for all
five BASELINE.json configs (sampler + plan included; the chunked loop the product's Config.run() executes).
With --gpus N > 1 it ends with `dp_check` (replica tables identical on every rank; candidate-sharded
link-prediction records == the unsharded ones) and a strong-scaling point beside the weak one.
"""
from __future__ import annotations

import argparse
import contextlib
import ctypes
import io
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "train triples/s (+ link-pred queries/s)"
W_PER_GPU = 8
# BASELINE.json `configs`, in order.  Optimizer: the reference's default (main_spark.py:316 SGD) where the config names
# none; learning rates: Config.py:66 default.  bern as in the reference default (0) except the negative-sampling-heavy
# config 3, which uses the bern coin (Base.cpp:116-117).
CONFIGS = [
    dict(id=1, model="TransE", dim=50, opt="SGD", k=1, shape="fb15k", nbatches=100, bern=0,
         name="TransE dim=50 L1 SGD margin=1 k=1, FB15K-shaped 14951/1345/483142, B=4831"),
    dict(id=2, model="TransH", dim=100, opt="Adam", k=1, shape="fb15k", nbatches=100, bern=0,
         name="TransH dim=100 Adam k=1 margin=1, FB15K-shaped 14951/1345/483142, B=4831"),
    dict(id=3, model="TransD", dim=100, opt="SGD", k=10, shape="wn18", nbatches=100, bern=1,
         name="TransD dim=100 SGD k=10 bern, WN18-shaped 40943/18/141442, B=1414"),
    dict(id=4, model="TransR", dim=100, opt="SGD", k=1, shape="fb15k", nbatches=100, bern=0,
         name="TransR ent=rel dim=100 SGD k=1, FB15K-shaped 14951/1345/483142, B=4831"),
    dict(id=5, model="TransE", dim=200, opt="SGD", k=1, shape="dbpedia", nbatches=0, bern=0,
         name="TransE dim=200 SGD k=1, DBpedia-shaped 4000000/600/20000000, B=2000 (auto rule)"),
]
HEAD = CONFIGS[1]
FP32_ALU_TFADD = 148 * 128 * 1.965e9 / 1e12      # lanes x SMs x max SM clock (derived, not measured)


def workload_string():
    return HEAD["name"] + " per GPU (nbatches=100)"


def config_block(world):
    """Identical in both arms (the driver compares them)."""
    return {"workload": workload_string(), "global_batch": 4831 * world, "parallelism": "dp%d" % world,
            "l2": "flushed between timed steps (256 MiB fill, outside the per-step events)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the train kernels, from the committed `ncu --set full`
    capture of this round (profiles/r02_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep); None if absent."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:  # noqa: BLE001
            return None
    return None


_GRAPHS = {}


def graph(shape):
    if shape not in _GRAPHS:
        from openkeonspark_b200 import datagen
        if shape == "dbpedia":
            # 20 M distinct triples: one vectorised draw + dedupe (datagen.make_graph's loop is sized for small graphs);
            # 100,000 valid + 100,000 test triples
            E, R, N, nv = 4_000_000, 600, 20_000_000, 100_000
            rng = np.random.default_rng(0)
            need = N + 2 * nv
            m = need + need // 50
            raw = np.stack([rng.integers(0, E, m), rng.integers(0, E, m), rng.integers(0, R, m)], 1)
            key = (raw[:, 0] * E + raw[:, 1]) * R + raw[:, 2]
            _, first = np.unique(key, return_index=True)
            raw = raw[np.sort(first)][:need]
            _GRAPHS[shape] = datagen.Graph(E, R, raw[:N], raw[N:N + nv], raw[N + nv:])
        else:
            _GRAPHS[shape] = datagen.make_shape(shape, seed=0)
    return _GRAPHS[shape]


def params_for(cfg, g, seed=0):
    from openkeonspark_b200 import datagen
    rng = np.random.default_rng(seed)
    D = cfg["dim"]
    if g.E > 1_000_000:      # 800 M normals: draw in fp32 directly
        import torch
        gen = torch.Generator().manual_seed(seed)
        ent = (torch.randn(g.E, D, generator=gen) * float(np.sqrt(2.0 / (g.E + D)))).numpy()
    else:
        ent = datagen.xavier_normal(rng, g.E, D)
    P = {"ent_embeddings": ent, "rel_embeddings": datagen.xavier_normal(rng, g.R, D)}
    if cfg["model"] == "TransH":
        P["normal_vectors"] = datagen.xavier_normal(rng, g.R, D)
    if cfg["model"] == "TransR":
        P["transfer_matrix"] = datagen.xavier_normal(rng, g.R, D * D)
    if cfg["model"] == "TransD":
        P["ent_transfer"] = datagen.xavier_normal(rng, g.E, D)
        P["rel_transfer"] = datagen.xavier_normal(rng, g.R, D)
    return P


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp"

    def __init__(self, index=0):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self, t_from=None, t_to=None):
        """Median SM clock over the samples taken inside [t_from, t_to] (time.time() values; all samples if None)."""
        import datetime
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:  # noqa: BLE001
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                if t_from is not None and len(f) > 7:
                    ts = datetime.datetime.strptime(f[7], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    if ts < t_from - 0.11 or ts > t_to + 0.11:
                        continue
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU baseline
class native_stdout_to_stderr:
    """The reference's Base.so prints with C printf; stdout must carry exactly ONE JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        try:
            ctypes.CDLL(None).fflush(None)
        except Exception:  # noqa: BLE001
            pass
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


class CpuReference:
    """The reference's serial loop on host cores (distribute_training.py:274-282): sampling() by the reference's own
    Base.so (oracle/_ref, compiled from /root/reference/base/Base.cpp; the C restatement if absent) followed by the
    TF-graph step restated in torch-CPU (TensorFlow 1.x is not installable)."""

    def __init__(self, cfg, g, threads):
        import tempfile

        import torch
        from openkeonspark_b200 import datagen
        from oracle import harness, models_ref
        torch.set_num_threads(threads)
        self.cfg, self.threads = cfg, threads
        self.B = g.train.shape[0] // cfg["nbatches"]
        d = tempfile.mkdtemp() + "/"
        datagen.write_dataset(g, d)
        self.kind, self.ref = "port", None
        if os.path.exists(harness.REF_SO):
            self.ref = harness.RefLib().init(d, bern=cfg["bern"], W=min(threads, 8), test=True, ontology=True)   # a missing ontology file is fine (Reader.h)
            self.sample = lambda: self.ref.sampling(self.B, cfg["k"], 0)
            self.native = "reference Base.so sampling() at workThreads=%d" % min(threads, 8)
            self.kind = "reference"
        else:
            orc = harness.COracle().load(d, test=False)
            orc.set_streams(np.arange(1, 9, dtype=np.uint64), cfg["bern"])
            self.sample = lambda: orc.sampling(self.B, cfg["k"], 0)
            self.native = "C restatement of sampling() (1 thread)"
        self.tr = models_ref.Trainer(cfg["model"], params_for(cfg, g), margin=1.0, lr=0.001, opt=cfg["opt"])

    def step(self):
        h, t, r, _ = self.sample()
        self.tr.step(h, t, r, self.B, self.cfg["k"], 0)

    def link_prediction(self, seconds=3.0):
        """getHead/TailBatch -> predict -> testHead/Tail, one query at a time, single-threaded native ranking (the reference
        parallelises evaluation only across Spark workers, distribute_training.py:430-441)."""
        if self.ref is None:
            return None
        try:
            nq, t0 = 0, time.perf_counter()
            n_test = self.ref.L.getTestTotal()
            while nq < 16 or time.perf_counter() - t0 < seconds:
                i = (nq // 2) * 97 % n_test
                side = nq % 2
                ch, ct, cr = self.ref.candidates(side, i)
                self.ref.rank(side, i, self.tr.predict(ch, ct, cr))
                nq += 1
            return {"queries_per_s": nq / (time.perf_counter() - t0), "queries": nq,
                    "what": "reference Base.so getHead/TailBatch + testHead/testTail (1 thread) around the torch-CPU predict_def (%d threads)" % self.threads}
        except Exception as e:  # noqa: BLE001
            return {"error": str(e)}

    def describe(self, n):
        return "%d steps of B=%d: %s + torch-CPU fp32 restatement of %s loss_def/%s (%d threads)" % (
            n, self.B, self.native, self.cfg["model"], self.cfg["opt"], self.threads)


def run_reference(args):
    """The reference arm: K timed "steps", each the mean of R back-to-back CPU train steps with R sized so that the timed
    region holds >= 10 s of CPU work (a 20-step sample of an 8 ms step is 0.16 s: too noisy a denominator)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    g = graph(HEAD["shape"])
    cores = os.cpu_count() or 1
    with native_stdout_to_stderr():
        cpu = CpuReference(HEAD, g, cores)
        t_step = 1e9
        for _ in range(max(1, args.warmup)):
            t0 = time.perf_counter()
            cpu.step()
            t_step = min(t_step, time.perf_counter() - t0)     # the first step pays one-off allocations
        reps = max(1, min(500, int(math.ceil(10.5 / max(args.steps * t_step, 1e-6)))))
        ts = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            for _ in range(reps):
                cpu.step()
            ts.append((time.perf_counter() - t0) / reps)
        lp = cpu.link_prediction()
    sec = sum(ts) / len(ts)
    val = cpu.B / sec
    sample = cpu.describe(args.steps * reps) + "; each of the %d reported steps is the mean of %d consecutive CPU steps (%.1f s timed)" % (
        args.steps, reps, sum(ts) * reps)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "triples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_block(world),
            "cpu_baseline": {"value": val, "unit": "triples/s", "cores": cores, "kind": cpu.kind, "sample": sample, "link_prediction": lp},
            "e2e": {"value": val, "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm helpers
def make_con(cfg, g, world, rank, global_batch=None, work_threads=None, lp=True):
    """A Config for `cfg` on graph g.  global_batch None: weak scaling (per-GPU batch = the config's B)."""
    import openkeonspark_b200 as okb
    from openkeonspark_b200 import parallel
    con = okb.Config(private_context=True)
    con.set_nbatches(cfg["nbatches"])
    con.set_ent_neg_rate(cfg["k"])
    con.set_margin(1.0)
    con.set_alpha(0.001)
    con.set_opt_method(cfg["opt"])
    con.set_dimension(cfg["dim"])
    con.set_bern(cfg["bern"])
    con.workThreads = work_threads if work_threads is not None else W_PER_GPU * world
    con.test_head = 1
    with contextlib.redirect_stdout(io.StringIO()):
        con.init_from_arrays(g.E, g.R, g.train, g.valid if lp else None, g.test if lp else None)
    B_local = con.batch_size
    if global_batch is not None:
        con.batch_size = int(global_batch)
        con._alloc_batch()
    elif world > 1:
        con.batch_size = B_local * world
        con._alloc_batch()
    con.set_model_and_session(getattr(okb, cfg["model"]))
    con.set_parameters(params_for(cfg, g))
    seeds = np.arange(1, con.workThreads + 1, dtype=np.uint64) * np.uint64(2654435761)
    con.ctx.call("okb_set_streams", ctypes.c_void_p(seeds.ctypes.data), con.workThreads)
    if world > 1:
        parallel.attach(con)
    return con, B_local


def train_bytes_per_positive(cfg, g, B):
    """Algorithmic bytes per positive triple: every row a positive group touches is read once and written once
    = 2 * 4 * D * [(2 + k) c_e + c_r]; TransR adds its matrices, 2 * 4 * D * D per distinct relation of the batch."""
    D, k = cfg["dim"], cfg["k"]
    ce = 2 if cfg["model"] == "TransD" else 1
    cr = {"TransE": 1, "TransH": 2, "TransD": 2, "TransR": 1}[cfg["model"]]
    b = 2 * 4 * D * ((2 + k) * ce + cr)
    if cfg["model"] == "TransR":
        rb = g.R * (1.0 - (1.0 - 1.0 / g.R) ** B)          # expected distinct relations among B uniform draws
        b += 2 * 4 * D * D * rb / B
    return b


def dense_adam_bytes(cfg, g):
    """TF1 Adam moves every row every step: 6 * 4 B per table element (var, m, v read + write)."""
    if cfg["opt"] != "Adam":
        return 0
    D = cfg["dim"]
    rows_e = g.E * (2 if cfg["model"] == "TransD" else 1)
    rows_r = g.R * {"TransE": 1, "TransH": 2, "TransD": 2, "TransR": 1 + D}[cfg["model"]]
    return 24 * (rows_e + rows_r) * D


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="okb200")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lp-queries", type=int, default=1 << 30, help="test triples ranked for the link-prediction figure (default: the whole test set)")
    ap.add_argument("--plan-ahead", type=int, default=64)
    ap.add_argument("--configs", default="1,2,3,4,5", help="BASELINE configs measured for the `configs` array (the headline is always config 2)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling point at N > 1")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    from openkeonspark_b200 import _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    clocks = ClockSampler(local) if rank == 0 else None      # nvidia-smi needs ~0.2 s to produce its first sample: start it early
    lib = _native.load()
    peak, peak_src = peaks()
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    align = torch.zeros(1, device=dev)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    g = graph(HEAD["shape"])
    con, B_local = make_con(HEAD, g, world, rank)
    chunk_steps = max(1, min(args.plan_ahead, args.steps))     # sampler + plan run once per chunk, on the launching stream
    con.plan_ahead = chunk_steps

    def begin_chunk_if_needed():
        """Sample + plan the next chunk ON THE TIMED STREAM (no side-stream look-ahead here: every launch of the step's
        work must sit between the events that time it)."""
        if con._world is not None and con._world.mode == "owner":
            return                                    # next_step() samples + plans its rank's positives itself
        if con._chunk_pos >= con._chunk_len:
            con.ctx.call("okb_chunk_begin", con.batch_size, con.negative_ent, con.negative_rel, chunk_steps,
                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            con._chunk_pos, con._chunk_len = 0, chunk_steps

    def one_step():
        if con._world is not None:
            con.next_step_device()
            return
        begin_chunk_if_needed()
        con.train_step_device(con._chunk_pos)
        con._chunk_pos += 1

    t_clk0 = time.time()                                     # clock samples are kept from warm-up to the end of the timed regions
    for _ in range(args.warmup):
        one_step()
    con._chunk_pos = con._chunk_len = 0                      # the timed region starts at a chunk boundary
    barrier()

    def timed_region(kernel_events):
        """K steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps."""
        con.ctx.call("okb_prof_enable", 1 if kernel_events else 0)
        l0 = lib.okb_launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        con._chunk_pos = con._chunk_len = 0
        barrier()
        w0 = time.perf_counter()
        for a, b in ev:
            if flush is not None:
                flush.fill_(1)
                if world > 1:
                    # the 256 MiB fills finish at different times on different ranks; a synchronous step would then time
                    # that skew as flag-wait inside the event pair.  A tiny all-reduce (stream-ordered, outside the
                    # events) lines the ranks up again, as they are in the real loop, which has no flush.
                    dist.all_reduce(align)
            a.record()
            one_step()
            b.record()
        barrier()
        wall = time.perf_counter() - w0
        con.ctx.call("okb_prof_enable", 0)
        ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
        return ms, wall, lib.okb_launch_count() - l0

    # ---------------- region A (headline): no per-kernel events inside the steps
    ms_total, t_wall, launches = timed_region(False)
    ms_per_step = ms_total / args.steps
    value = con.batch_size * args.steps / (ms_total * 1e-3)      # global positives per second

    # ---------------- region B (roofline): the same K steps again with a CUDA-event pair around each kernel
    # (the extra event records cost ~10 us per step, which is why the headline is taken without them)
    ms_instr, _, _ = timed_region(True)
    prof = {}
    for name, kid in (("sample", 0), ("plan", 1), ("grad", 2), ("update", 3), ("dp_push", 6), ("dp_owner", 7)):
        ms, cnt = ctypes.c_double(), ctypes.c_int64()
        con.ctx.call("okb_prof_read", kid, ctypes.byref(ms), ctypes.byref(cnt))
        prof[name] = (ms.value, cnt.value)
    if prof["sample"][1] <= 0 or prof["plan"][1] <= 0:
        raise SystemExit("bench.py: the timed region executed no sampler / plan launch (%r) — the headline would skip work" % (prof,))

    # ---------------- region C: the real training loop — one library call per chunk of steps, no L2 flush.  Single GPU:
    # the next chunk is sampled + planned on the library's side stream while this one trains; the region ends with the
    # main stream waiting for the last chunk produced inside it, so n chunks are consumed AND n chunks are produced.
    con.plan_ahead = args.plan_ahead
    n_chunks = max(2, -(-args.steps // con.plan_ahead))
    con.train_chunk_device()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    a.record()
    for _ in range(n_chunks):
        con.train_chunk_device()
    if con._world is None:
        con.ctx.call("okb_chunk_begin", con.batch_size, con.negative_ent, con.negative_rel, con.plan_ahead,
                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    b.record()
    barrier()
    ms_c = max_over_ranks(a.elapsed_time(b))
    chunk = {"value": con.batch_size * n_chunks * con.plan_ahead / (ms_c * 1e-3), "unit": "triples/s",
             "ms_per_step": ms_c / (n_chunks * con.plan_ahead), "steps": n_chunks * con.plan_ahead,
             "wall_ms_per_step": (time.perf_counter() - w0) * 1e3 / (n_chunks * con.plan_ahead),
             "what": "Config.train_chunk_device(): %d steps per library call, tables L2-resident (no flush); sampler + plan of "
                     "every chunk inside the region" % con.plan_ahead}
    clk = clocks.stop(t_clk0, time.time()) if clocks else None

    # ---------------- e2e: the reference-shaped loop through the public API with HOST buffers
    # con.sampling() fills the numpy batch_h/t/r/y (D2H); con.train_step(...) feeds them back (H2D) and
    # returns the loss as a Python float (D2H) — distribute_training.py:274-282.  Always >= 500 steps.
    e2e_steps = max(500, min(args.steps, 2000))
    for _ in range(5):
        con.sampling(); con.train_step(con.batch_h, con.batch_t, con.batch_r, con.batch_y)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        con.sampling()
        con.train_step(con.batch_h, con.batch_t, con.batch_r, con.batch_y)
    barrier()
    e2e_sec = max_over_ranks(time.perf_counter() - t0)
    S = con.batch_seq_size
    e2e = {"value": con.batch_size * e2e_steps / e2e_sec, "unit": "triples/s", "h2d_bytes_per_step": 3 * 8 * S,
           "d2h_bytes_per_step": 3 * 8 * S + 4, "steps": e2e_steps, "ms_per_step": e2e_sec * 1e3 / e2e_steps}

    # ---------------- roofline of the dominant kernel
    # Algorithmic bytes per launch: Adam update = 6*4 B per table element (var, m, v read+write)
    # over ALL rows (TF1 dense-decay semantics) + one read of every gradient row;
    # grad kernel = gather of (2+k) entity rows + 2 relation-side rows per positive + the gradient rows it writes.
    D, k = HEAD["dim"], HEAD["k"]
    n_tab = (g.E + 2 * g.R) * D
    ne_rows, nr_rows = con.batch_size * (2 + k), con.batch_size
    bytes_update = 24 * n_tab + 4 * (ne_rows * D + nr_rows * 2 * D)
    bytes_grad = (con.batch_size // world) * 4 * D * ((2 + k) + 2) * 2
    kern = {}
    if world == 1:
        for name, nbytes in (("update", bytes_update), ("grad", bytes_grad)):
            ms, cnt = prof.get(name, (0, 0))
            if cnt:
                kern[name] = {"ms": ms / cnt, "gbs": nbytes / (ms / cnt * 1e-3) / 1e9, "bytes": nbytes}
    else:
        ms, cnt = prof.get("grad", (0, 0))
        if cnt:
            kern["grad"] = {"ms": ms / cnt, "gbs": bytes_grad / (ms / cnt * 1e-3) / 1e9, "bytes": bytes_grad}
    dom = max(kern, key=lambda n: kern[n]["ms"]) if kern else None
    traffic = ncu_traffic()
    roofline = None
    per_launch = {n: (prof[n][0] / prof[n][1] if prof[n][1] else None) for n in prof}
    if dom:
        kname = {"update": "adam_tile_kernel", "grad": "grad_k1_kernel"}[dom]
        tr = None
        if traffic and world == 1:
            for kk, v in traffic.get("kernels", {}).items():
                if kname in kk:
                    tr = v.get("dram_bytes_per_launch")
        roofline = {"kernel": kname, "bound": "hbm", "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": kern[dom]["gbs"] / peak, "traffic": tr, "traffic_source": (traffic or {}).get("source"),
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": kern[dom]["bytes"], "avg_launch_ms": kern[dom]["ms"],
                    "measured": "CUDA-event pair around every launch of the kernel in a second pass of the same %d steps (L2 flushed between steps)" % args.steps,
                    "instrumented_ms_per_step": ms_instr / args.steps,
                    "per_launch_ms": per_launch, "launches_in_pass": {n: prof[n][1] for n in prof},
                    "step": {"algorithmic_bytes": 24 * n_tab + (con.batch_size // world) * 2 * 4 * D * ((2 + k) + 2),
                             "GB/s": (24 * n_tab + (con.batch_size // world) * 2 * 4 * D * ((2 + k) + 2)) / (ms_per_step * 1e-3) / 1e9,
                             "frac": (24 * n_tab + (con.batch_size // world) * 2 * 4 * D * ((2 + k) + 2)) / (ms_per_step * 1e-3) / 1e9 / peak,
                             "frac_chunked_loop": (24 * n_tab + (con.batch_size // world) * 2 * 4 * D * ((2 + k) + 2)) / (chunk["ms_per_step"] * 1e-3) / 1e9 / peak,
                             "what": "SURVEY 8(d) bytes of one whole step / ms_per_step (flushed) and / the chunked loop's ms_per_step"},
                    "other_kernels": {n: {"GB/s": kern[n]["gbs"], "frac": kern[n]["gbs"] / peak, "bytes": kern[n]["bytes"]} for n in kern if n != dom}}

    # ---------------- secondary metric of the headline config: filtered link-prediction queries/s (both sides)
    def measure_lp(c, cfg, gg, nq):
        rec_fn = (lambda: c._world.link_prediction(c, 0, nq)) if c._world is not None else (lambda: c.link_prediction_records(0, nq))
        rec_fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rec_fn()
        b.record()
        barrier()
        ms = max_over_ranks(a.elapsed_time(b))
        qps = 2 * nq / (ms * 1e-3)
        Dd = cfg["dim"]
        ce = 2 if cfg["model"] == "TransD" else 1
        return {"queries_per_s": qps, "queries": 2 * nq, "ms": ms,
                "roofline": {"bound": "fp32_alu", "achieved": qps * 2.0 * gg.E * Dd / 1e12 / world, "peak": FP32_ALU_TFADD, "unit": "TFADD/s per GPU",
                             "frac": qps * 2.0 * gg.E * Dd / 1e12 / FP32_ALU_TFADD / world,
                             "hbm_canonical_frac": qps * 4.0 * Dd * gg.E * ce / 1e9 / peak / world,
                             "what": "2 FADD per (query, candidate, dim) against 148 SM x 128 lanes x 1.965 GHz (derived peak); "
                                     "hbm_canonical_frac = queries/s x 4*D*E*c_e bytes (SURVEY 8d: each query re-streams the table) / HBM peak — "
                                     "above 1 because one table pass serves a whole batch of queries"},
                "what": "filtered+raw+type-constrained ranks, head and tail side, all %d candidates%s" % (
                    gg.E, " sharded over %d GPUs" % world if world > 1 else "")}

    lp = None
    try:
        lp = measure_lp(con, HEAD, g, min(args.lp_queries, con.testTotal))
        if world == 1:      # the same evaluation in the regime of a TRAINED model: few candidates beat the target
            nq_t = min(args.lp_queries, con.testTotal)
            g_t = planted_test_set(dev, HEAD, g, nq_t)
            con_t, _ = make_con(HEAD, g_t, 1, 0)
            r_t = measure_lp(con_t, HEAD, g_t, nq_t)
            rec = con_t.link_prediction_records(0, min(nq_t, 4096)).cpu().numpy()
            lp["trained_like"] = {"queries_per_s": r_t["queries_per_s"], "ms": r_t["ms"], "queries": r_t["queries"],
                                  "roofline_frac_fp32_alu": r_t["roofline"]["frac"],
                                  "mean_raw_rank_tail": float(rec[:, 1, 0].mean()), "mean_raw_rank_head": float(rec[:, 0, 0].mean()),
                                  "what": "same tables, test triples re-planted so that the true entity is the best of 64 random candidates "
                                          "(raw mean rank ~ E/65, the regime of a trained model): the better-than counting epilogue runs for ~1.5 % "
                                          "of the candidates instead of ~50 % with random tables"}
            del con_t
            torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001  (the secondary metric must not kill the headline line)
        lp = lp or {}
        lp["error"] = repr(e)

    # ---------------- data-parallel correctness in the driver's own record (N > 1)
    dp_check = None
    if world > 1:
        try:
            dp_check = run_dp_check(con, dist, dev, world, rank)
        except Exception as e:  # noqa: BLE001
            dp_check = "FAILED: %r" % (e,)

    # ---------------- strong scaling beside the weak point: SURVEY cfg2's own case, global B = 4,831 split over N ranks
    strong = None
    if world > 1 and not args.no_strong:
        try:
            if con._world is not None and hasattr(con._world, "close"):
                con._world.close(con)
            con_s, _ = make_con(HEAD, g, world, rank, global_batch=B_local, work_threads=8 if 8 % world == 0 else world, lp=False)
            con_s.plan_ahead = args.plan_ahead
            con_s.train_chunk_device()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n_chunks):
                con_s.train_chunk_device()
            b.record()
            barrier()
            ms_s = max_over_ranks(a.elapsed_time(b))
            strong = {"global_batch": B_local, "value": B_local * n_chunks * con_s.plan_ahead / (ms_s * 1e-3), "unit": "triples/s",
                      "ms_per_step": ms_s / (n_chunks * con_s.plan_ahead), "mode": con_s._world.mode,
                      "what": "the reference batch (B=4831, workThreads=8) split over %d ranks: %d positives per rank, chunked loop" % (world, -(-B_local // world))}
            if hasattr(con_s._world, "close"):
                con_s._world.close(con_s)
            del con_s
        except Exception as e:  # noqa: BLE001
            strong = {"error": repr(e)}

    # ---------------- all five BASELINE configs (train: chunked loop incl. sampler + plan; LP: a slice of the test set)
    configs_out = []
    want = [int(x) for x in args.configs.split(",") if x.strip()]
    for cfg in CONFIGS:
        if cfg["id"] not in want:
            continue
        entry = {"id": cfg["id"], "workload": cfg["name"], "n_gpus": world}
        try:
            if cfg["id"] == HEAD["id"]:
                cc, gg, Bl = con, g, B_local
                if world > 1 and strong is not None:     # the headline context was closed for the strong-scaling run
                    cc = None
            else:
                gg = graph(cfg["shape"])
                if world > 1 and cfg["id"] == 5:
                    cc, Bl = make_con(cfg, gg, 1, 0)       # tables replicated, no data-parallel train (B=2000 does not shard usefully)
                else:
                    cc, Bl = make_con(cfg, gg, world, rank)
            if cc is None:
                entry["train"] = {"triples_per_s": chunk["value"], "ms_per_step": chunk["ms_per_step"], "from": "training_loop_chunked"}
                entry["link_prediction"] = lp
                configs_out.append(entry)
                continue
            # ---- train
            if world > 1 and cfg["id"] == 5:
                entry["train"] = {"skipped": "measured at N=1: the auto batch rule gives B=2000, which does not shard usefully"}
            else:
                cc.plan_ahead = args.plan_ahead
                nch = 3 if cfg["id"] != HEAD["id"] else n_chunks
                cc.train_chunk_device()
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                if flush is not None:
                    flush.fill_(1)
                a.record()
                for _ in range(nch):
                    cc.train_chunk_device()
                if cc._world is None:
                    cc.ctx.call("okb_chunk_begin", cc.batch_size, cc.negative_ent, cc.negative_rel, cc.plan_ahead,
                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
                b.record()
                barrier()
                ms = max_over_ranks(a.elapsed_time(b)) / (nch * cc.plan_ahead)
                tps = cc.batch_size / (ms * 1e-3)
                bpp = train_bytes_per_positive(cfg, gg, Bl)
                step_bytes = bpp * cc.batch_size + dense_adam_bytes(cfg, gg)
                entry["train"] = {"triples_per_s": tps, "ms_per_step": ms, "batch": cc.batch_size, "steps": nch * cc.plan_ahead,
                                  "mode": "single GPU" if cc._world is None else cc._world.mode,
                                  "roofline": {"bound": "hbm", "achieved": step_bytes / (ms * 1e-3) / 1e9 / max(world, 1), "peak": peak, "unit": "GB/s",
                                               "frac": step_bytes / (ms * 1e-3) / 1e9 / max(world, 1) / peak,
                                               "algorithmic_bytes_per_step": step_bytes,
                                               "what": "SURVEY 8(d): 2*4*D*[(2+k)c_e+c_r] bytes per positive (+ TransR matrices, + 24 B per table element "
                                                       "for TF1 dense Adam) / step time of the chunked loop (sampler + plan included), per GPU"}}
            # ---- link prediction
            nq = min(args.lp_queries, cc.testTotal, 512 if cfg["id"] == 5 else 1 << 30)
            if world > 1 and cfg["id"] == 5:
                from openkeonspark_b200 import parallel
                cc._world = parallel.DataParallel(cc)     # candidate-sharded evaluation only: 500 k candidates per rank at N=8
            entry["link_prediction"] = measure_lp(cc, cfg, gg, nq)
            if world > 1 and cfg["id"] == 5:
                cc._world = None
            if cc is not con:
                if cc._world is not None and hasattr(cc._world, "close"):
                    cc._world.close(cc)
                del cc
                torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001  (one config must not kill the line)
            entry["error"] = repr(e)
        if cfg["shape"] == "dbpedia":
            _GRAPHS.pop("dbpedia", None)
        configs_out.append(entry)

    # ---------------- the reference's CPU path beside it (rank 0, N = 1): ~10 s of CPU work
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        with native_stdout_to_stderr():
            cr = CpuReference(HEAD, g, cores)
            for _ in range(2):
                cr.step()
            ts, t_start = [], time.perf_counter()
            while len(ts) < 12 or time.perf_counter() - t_start < 10.0:
                t0 = time.perf_counter()
                cr.step()
                ts.append(time.perf_counter() - t0)
            cpu_lp = cr.link_prediction()
        sec = sum(ts) / len(ts)
        cpu = {"value": cr.B / sec, "unit": "triples/s", "cores": cores, "kind": cr.kind, "sample": cr.describe(len(ts)),
               "link_prediction": cpu_lp}

    if rank == 0:
        cfgb = config_block(world)
        if flush is None:
            cfgb["l2"] = "not flushed"
        line = {"metric": METRIC, "value": value, "unit": "triples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": cfgb,
                "measurement": {"timing": "per-step CUDA events summed; max over ranks",
                                "chunk": "sampler + plan of %d steps at a time, on the timed stream, inside the step events" % chunk_steps,
                                "parallelism": "dp%d%s" % (world, "" if world == 1 else " (%s)" % (con._world.mode if con._world is not None else "owner")),
                                "workThreads": con.workThreads, "batch_per_gpu": B_local},
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "link_prediction": lp, "training_loop_chunked": chunk, "configs": configs_out,
                "wall_s_timed_region_incl_flush": t_wall}
        if world > 1:
            line["dp_check"] = dp_check
            line["strong_scaling"] = strong
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def planted_test_set(dev, cfg, g, n, n_cand=64):
    """A test set on which the CURRENT tables rank the true entity well — the regime of a trained model (raw mean rank
    ~ E / 65, like published TransE numbers on FB15K) instead of random tables, where half of all candidates beat the target
    and the counting epilogue of the ranking kernel is over-weighted.  For each of the first n test triples the tail is
    replaced by the best of n_cand random candidates under the model's own score, then the head likewise."""
    import torch
    from openkeonspark_b200 import datagen
    P = {k: torch.as_tensor(v, device=dev) for k, v in params_for(cfg, g).items()}      # the tables make_con() starts from
    ent, rel = P["ent_embeddings"], P["rel_embeddings"]
    l2n = lambda x: x * torch.rsqrt(torch.clamp((x * x).sum(-1, keepdim=True), min=1e-12))

    def score(h, t, r):                                     # index tensors of equal shape -> scores of that shape
        he, te, re = ent[h], ent[t], rel[r]
        if cfg["model"] == "TransH":
            nv = l2n(P["normal_vectors"][r])
            he = he - (he * nv).sum(-1, keepdim=True) * nv
            te = te - (te * nv).sum(-1, keepdim=True) * nv
        return (l2n(he) + l2n(re) - l2n(te)).abs().sum(-1)

    test = torch.as_tensor(np.ascontiguousarray(g.test[:n]), device=dev)
    gen = torch.Generator(device=dev).manual_seed(7)
    out = []
    for lo in range(0, n, 4096):
        blk = test[lo:lo + 4096]
        h, t, r = blk[:, 0:1], blk[:, 1:2], blk[:, 2:3]
        cand = torch.randint(0, g.E, (blk.shape[0], n_cand), device=dev, generator=gen)
        cand[:, 0] = t[:, 0]
        t2 = cand.gather(1, score(h.expand_as(cand), cand, r.expand_as(cand)).argmin(1, keepdim=True))
        cand = torch.randint(0, g.E, (blk.shape[0], n_cand), device=dev, generator=gen)
        cand[:, 0] = h[:, 0]
        h2 = cand.gather(1, score(cand, t2.expand_as(cand), r.expand_as(cand)).argmin(1, keepdim=True))
        out.append(torch.cat([h2, t2, r], 1))
    new_test = torch.cat(out, 0).cpu().numpy().astype(np.int64)
    return datagen.Graph(g.E, g.R, g.train, g.valid, new_test)


def run_dp_check(con, dist, dev, world, rank):
    """(1) replica tables identical on every rank after everything trained so far (all-reduced min/max of a 64-bit checksum);
    (2) 64 candidate-sharded link-prediction records == the same records ranked unsharded on every rank."""
    import torch
    P = con.get_parameters()                              # owner mode: settles peers' last row updates first
    sums = []
    for kname in sorted(P):
        a = np.ascontiguousarray(P[kname]).view(np.uint32).astype(np.uint64)
        w = (np.arange(a.size, dtype=np.uint64) % np.uint64(1000003)) + np.uint64(1)
        sums.append(int((a.reshape(-1) * w).sum(dtype=np.uint64) & np.uint64(0x7FFFFFFFFFFFFFFF)))
    t = torch.tensor(sums, dtype=torch.int64, device=dev)
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if not torch.equal(lo, hi):
        return "FAILED: replica tables differ across ranks"
    nq = min(64, con.testTotal)
    sharded = con._world.link_prediction(con, 0, nq)
    full = con.link_prediction_records(0, nq)
    ok = torch.tensor([1 if torch.equal(sharded, full) else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) != 1:
        return "FAILED: candidate-sharded link-prediction records differ from the unsharded ones"
    return "ok"


if __name__ == "__main__":
    main()
