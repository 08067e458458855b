"""Where the host-buffer (e2e) step goes: wall time of con.sampling() and con.train_step() and of their pieces."""
import contextlib, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import openkeonspark_b200 as okb
from openkeonspark_b200 import datagen
g = datagen.make_shape("fb15k", seed=0)
con = okb.Config(private_context=True)
con.set_nbatches(100); con.set_dimension(100); con.set_opt_method("Adam"); con.workThreads = 8
with contextlib.redirect_stdout(io.StringIO()):
    con.init_from_arrays(g.E, g.R, g.train, g.valid, g.test)
con.set_model_and_session(okb.TransH)
for _ in range(5):
    con.sampling(); con.train_step(con.batch_h, con.batch_t, con.batch_r, con.batch_y)
N = 200
def timeit(f):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(N): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / N * 1e6
print("sampling()            %.1f us" % timeit(con.sampling))
print("  sampling_device()   %.1f us (async)" % timeit(con.sampling_device))
from openkeonspark_b200.Config import _vp, _stream
def _y():
    con.batch_y[:con.batch_size] = 1.0; con.batch_y[con.batch_size:] = -1.0
print("  label fill          %.1f us" % timeit(_y))
print("  okb_sample_to_host  %.1f us (sync)" % timeit(lambda: con.ctx.call("okb_sample_to_host", con.batch_size, con.negative_ent, con.negative_rel, 0, con.workThreads, _vp(con.batch_h_addr), _vp(con.batch_t_addr), _vp(con.batch_r_addr), _stream())))
print("  okb_sample + okb_batch_to_host %.1f us (sync)" % timeit(lambda: (con.sampling_device(), con.ctx.call("okb_batch_to_host", 0, _vp(con.batch_h_addr), _vp(con.batch_t_addr), _vp(con.batch_r_addr), None, _stream()))))
print("train_step()          %.1f us" % timeit(lambda: con.train_step(con.batch_h, con.batch_t, con.batch_r, con.batch_y)))
print("  _hyper()            %.1f us" % timeit(lambda: con._hypers(1)))
print("  train_step_device   %.1f us (async, incl. 1-step plan)" % timeit(lambda: con.train_step_device(0)))
print("  loss .item()        %.1f us" % timeit(lambda: con._loss_dev.item()))
both = timeit(lambda: (con.sampling(), con.train_step(con.batch_h, con.batch_t, con.batch_r, con.batch_y)))
print("sampling+train_step   %.1f us -> %.3g triples/s" % (both, con.batch_size / both * 1e6))
