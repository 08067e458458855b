"""Phase timeline of the owner-sharded (scatter form) data-parallel step from the kernels' own %globaltimer stamps
(OKB_FLAG_DP_TRACE).  Run under torchrun on N GPUs of one box (N = 1 works too):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_trace.py

Prints, per rank, the median over the traced steps of each phase (us):
  x_wait   grad kernel: grid dependency resolved -> first block past the "tables complete" flags
  x_skew   first -> last block past those flags
  grad     first block past the flags -> last block out
  gap1     last grad block out -> owner-update kernel's grid dependency resolved (drain of the peer stores + launch)
  s_wait   -> first block past the "gradient rows landed" flags (handshake + waiting for the slowest rank's grad kernel)
  hub      -> last hub block done          tiles    -> last tile out (from s_wait's end)
  gap2     last tile out -> next step's grad kernel dependency resolved
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import bench


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29578")
    if "RANK" in os.environ:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    world, rank = dist.get_world_size(), dist.get_rank()
    from openkeonspark_b200 import parallel
    g = bench.graph(bench.HEAD["shape"])
    con, _ = bench.make_con(bench.HEAD, g, 1, 0, lp=False, global_batch=4831 * world, work_threads=8 * world)
    form = os.environ.get("OKB200_DP_FORM", "scatter")
    flushed = os.environ.get("OKB200_TRACE_FLUSH") == "1"
    parallel.attach(con, mode="owner", form=form)
    hs = int(os.environ.get("OKB200_DP_HANDSHAKE", "-1"))
    if hs >= 0:
        con.ctx.call("okb_set_flag", 14, hs)
    steps = 48
    con.plan_ahead = steps
    for _ in range(3):
        con.train_chunk_device()
    torch.cuda.synchronize()
    dist.barrier()
    con.ctx.call("okb_set_flag", 13, 1)
    if flushed:                                        # bench.py's headline regime: L2 flushed and ranks re-aligned before every step
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        align = torch.zeros(1, device="cuda")
        con._chunk_pos = con._chunk_len = 0
        for _ in range(steps):
            flush.fill_(1)
            dist.all_reduce(align)
            con.next_step_device()
    else:
        con.train_chunk_device()
    buf = np.zeros(64 * 16, np.uint64)
    con.ctx.call("okb_debug_dp_trace", ctypes.c_void_p(buf.ctypes.data))
    con.ctx.call("okb_set_flag", 13, 0)
    t = buf.reshape(64, 16)
    for s in (3, 4, 8, 9):
        t[:, s] = ~t[:, s]
    none = np.uint64(0xFFFFFFFFFFFFFFFF)
    rows = [r for r in range(64) if t[r, 1] != none and t[r, 9] != 0 and t[r, 7] != none]
    rows.sort(key=lambda r: int(t[r, 1]))
    ph = {k: [] for k in ("x_wait", "x_skew", "grad", "gap1", "push", "gapB", "s_wait", "hub", "tiles", "gap2", "step")}
    f = lambda a, b: (int(a) - int(b)) / 1e3
    for s in (12,):
        t[:, s] = ~t[:, s]
    for i, r in enumerate(rows[1:-1], 1):
        ph["x_wait"].append(f(t[r, 2], t[r, 1])); ph["x_skew"].append(f(t[r, 3], t[r, 2])); ph["grad"].append(f(t[r, 4], t[r, 2]))
        if form == "push":      # gap1: last grad block out -> push kernel's dependency resolved; gapB: last push block out -> owner's
            ph["gap1"].append(f(t[r, 11], t[r, 4])); ph["push"].append(f(t[r, 12], t[r, 11])); ph["gapB"].append(f(t[r, 6], t[r, 12]))
        else:
            ph["gap1"].append(f(t[r, 6], t[r, 4])); ph["push"].append(0.0); ph["gapB"].append(0.0)
        ph["s_wait"].append(f(t[r, 7], t[r, 6]))
        ph["hub"].append(f(t[r, 8], t[r, 7]) if t[r, 8] != 0 else 0.0); ph["tiles"].append(f(t[r, 9], t[r, 7]))
        nxt = rows[i + 1]
        ph["gap2"].append(f(t[nxt, 1], t[r, 9])); ph["step"].append(f(t[nxt, 1], t[r, 1]))
    for q in range(world):
        if q == rank:
            print("%s%s hs=%s rank %d (%d steps): " % (form, " flushed" if flushed else "", os.environ.get("OKB200_DP_HANDSHAKE", "default"), rank, len(ph["step"])) + "  ".join("%s %.2f" % (k, float(np.median(v))) for k, v in ph.items()), flush=True)
        dist.barrier()
    con._world.close(con)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
