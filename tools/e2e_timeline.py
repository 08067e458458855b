"""Host-side timeline of the host-buffer loop: mean wall time of sampling() and train_step() inside the steady-state loop,
next to the per-kernel CUDA-event times of the same loop (library profiler; events serialise the PDL overlap)."""
import contextlib, ctypes, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
g = bench.graph(bench.HEAD["shape"])
con, _ = bench.make_con(bench.HEAD, g, 1, 0, lp=False)
for _ in range(20):
    con.sampling(); con.train_step(con.batch_h, con.batch_t, con.batch_r, con.batch_y)
N = 500
ts = np.zeros((N, 3))
torch.cuda.synchronize()
for i in range(N):
    ts[i, 0] = time.perf_counter()
    con.sampling()
    ts[i, 1] = time.perf_counter()
    con.train_step(con.batch_h, con.batch_t, con.batch_r, con.batch_y)
    ts[i, 2] = time.perf_counter()
torch.cuda.synchronize()
d = np.diff(ts, axis=1) * 1e6
gap = (ts[1:, 0] - ts[:-1, 2]) * 1e6
print("sampling() %.1f us   train_step() %.1f us   loop overhead %.1f us   step %.1f us" % (d[:, 0].mean(), d[:, 1].mean(), gap.mean(), (ts[-1, 2] - ts[0, 0]) / N * 1e6))
con.ctx.call("okb_prof_enable", 1)
for i in range(100):
    con.sampling(); con.train_step(con.batch_h, con.batch_t, con.batch_r, con.batch_y)
for name, kid in (("sample", 0), ("plan", 1), ("grad", 2), ("update", 3)):
    ms, cnt = ctypes.c_double(), ctypes.c_int64()
    con.ctx.call("okb_prof_read", kid, ctypes.byref(ms), ctypes.byref(cnt))
    if cnt.value:
        print("  %-7s %.1f us x %d" % (name, ms.value / cnt.value * 1e3, cnt.value))
