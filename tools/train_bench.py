"""Per-config train-step timing (not the driver's bench): python tools/train_bench.py MODEL DIM SHAPE OPT K [NBATCHES] [BERN]
Prints microseconds per step for the chunked loop (Config.train_chunk_device) and a per-kernel breakdown."""
import contextlib
import ctypes
import io
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import openkeonspark_b200 as okb  # noqa: E402
from openkeonspark_b200 import datagen  # noqa: E402

model, D, shape, opt, k = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4], int(sys.argv[5])
nb = int(sys.argv[6]) if len(sys.argv) > 6 else 100
bern = int(sys.argv[7]) if len(sys.argv) > 7 else 0
t0 = time.time()
g = datagen.make_shape(shape, seed=0)
t1 = time.time()
con = okb.Config(private_context=True)
con.set_nbatches(nb); con.set_dimension(D); con.set_opt_method(opt); con.set_ent_neg_rate(k); con.set_bern(bern)
con.workThreads = 8
with contextlib.redirect_stdout(io.StringIO()):
    con.init_from_arrays(g.E, g.R, g.train, g.valid, g.test)
t2 = time.time()
con.set_model_and_session(getattr(okb, model))
n = con.plan_ahead
for _ in range(2):
    con.train_chunk_device(n)
torch.cuda.synchronize()
reps = 4
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    losses = con.train_chunk_device(n)
b.record(); torch.cuda.synchronize()
steps = reps * losses.numel()
us = a.elapsed_time(b) * 1e3 / steps
con.ctx.call("okb_prof_enable", 1)
losses = con.train_chunk_device(n)
torch.cuda.synchronize()
con.ctx.call("okb_prof_enable", 0)
br = {}
for nm, kid in (("sample", 0), ("plan", 1), ("grad", 2), ("update", 3)):
    t, c = ctypes.c_double(), ctypes.c_int64()
    con.ctx.call("okb_prof_read", kid, ctypes.byref(t), ctypes.byref(c))
    br[nm] = "%.1f us x%d" % (t.value * 1e3 / max(c.value, 1), c.value)
print("%s D=%d %s %s k=%d B=%d chunk=%d: %.1f us/step, %.3g triples/s | gen %.1fs load %.1fs | %s | loss %.4f" %
      (model, D, shape, opt, k, con.batch_size, losses.numel(), us, con.batch_size / us * 1e6, t1 - t0, t2 - t1, br, float(losses[-1])))
