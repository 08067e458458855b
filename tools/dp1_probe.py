"""Owner-sharded data-parallel step on ONE rank (world size 1): every "peer" store is local, there is no flag skew, so the
difference to the plain single-GPU step is the cost of the kernel structure alone (and the process can run under ncu).

    python tools/dp1_probe.py [form ...]          forms: none (plain single GPU), push, scatter

Prints, per form: us per step of the chunked loop (tables L2-resident), us per step with the L2 flushed between steps, and
the per-kernel event times of the library's profiler."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import bench


def run(form, g, steps=64, mult=1):
    from openkeonspark_b200 import parallel
    con, B_local = bench.make_con(bench.HEAD, g, 1, 0, lp=False, global_batch=None if mult == 1 else 4831 * mult,
                                  work_threads=8 * mult)
    if form != "none":
        parallel.attach(con, mode="owner", form=form)
    con.plan_ahead = steps
    for _ in range(3):
        con.train_chunk_device()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(4):
        con.train_chunk_device()
    b.record()
    torch.cuda.synchronize()
    chunked = a.elapsed_time(b) / (4 * steps) * 1e3
    # flushed, per-step events
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    con.plan_ahead = 20
    con._chunk_pos = con._chunk_len = 0
    tot = 0.0
    con.ctx.call("okb_prof_enable", 1)
    for i in range(20):
        flush.fill_(1)
        a.record()
        if form == "none":
            if con._chunk_pos >= con._chunk_len:
                con.ctx.call("okb_chunk_begin", con.batch_size, con.negative_ent, con.negative_rel, 20,
                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
                con._chunk_pos, con._chunk_len = 0, 20
            con.train_step_device(con._chunk_pos)
            con._chunk_pos += 1
        else:
            con.next_step_device()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    prof = {}
    for name, kid in (("sample", 0), ("plan", 1), ("grad", 2), ("update", 3), ("dp_push", 6), ("dp_owner", 7)):
        ms, cnt = ctypes.c_double(), ctypes.c_int64()
        con.ctx.call("okb_prof_read", kid, ctypes.byref(ms), ctypes.byref(cnt))
        if cnt.value:
            prof[name] = round(ms.value / cnt.value * 1e3, 2)
    con.ctx.call("okb_prof_enable", 0)
    print("form=%-8s batch=%d  chunked %.2f us/step   flushed+instrumented %.2f us/step   per launch (us): %s"
          % (form, con.batch_size, chunked, tot / 20 * 1e3, prof), flush=True)
    if con._world is not None:
        con._world.close(con)


def main():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    g = bench.graph(bench.HEAD["shape"])
    forms = [a for a in sys.argv[1:] if not a.startswith("x")] or ["none", "push", "scatter"]
    mult = max([int(a[1:]) for a in sys.argv[1:] if a.startswith("x")] or [1])
    for f in forms:
        run(f, g, mult=mult)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
