"""Link-prediction timing (not the driver bench): python tools/lp_bench.py MODEL DIM N_TEST_TRIPLES HEADS [tc]\nPrints rank-kernel / prep milliseconds (CUDA events) and queries/s on the FB15K-shaped graph."""
import sys, time, ctypes, numpy as np, torch, contextlib, io
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import openkeonspark_b200 as okb
from openkeonspark_b200 import datagen
from conftest import make_params
model, D, nq, heads = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
tc = len(sys.argv) > 5 and sys.argv[5] == 'tc'
g = datagen.make_shape("fb15k", seed=0)
con = okb.Config(private_context=True)
con.set_nbatches(100); con.set_dimension(D); con.workThreads = 8; con.test_head = heads
with contextlib.redirect_stdout(io.StringIO()):
    con.init_from_arrays(g.E, g.R, g.train, g.valid, g.test)
con.set_model_and_session(getattr(okb, model))
con.set_parameters(make_params(model, g.E, g.R, D, seed=0))
con.transr_tensor_cores = tc
con.link_prediction_records(0, nq)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
con.ctx.call("okb_prof_enable", 1)
a.record(); con.link_prediction_records(0, nq); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
for nm, kid in (("rank", 4), ("prep", 5)):
    t, n = ctypes.c_double(), ctypes.c_int64()
    con.ctx.call("okb_prof_read", kid, ctypes.byref(t), ctypes.byref(n))
    print("   ", nm, "ms %.3f" % t.value, "launch groups", n.value)
print("tc" if tc else "", model, D, "queries", nq * (2 if heads else 1), "ms %.3f" % ms, "q/s %.3g" % (nq * (2 if heads else 1) / ms * 1e3))
