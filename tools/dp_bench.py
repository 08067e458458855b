"""Owner-sharded data-parallel step breakdown (torchrun, N ranks; N=1 isolates the kernels from NVLink):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 tools/dp_bench.py [MODEL DIM OPT]
Prints per-kernel CUDA-event times (they include the time a kernel spends waiting for its peers)."""
import contextlib, ctypes, io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import openkeonspark_b200 as okb
from openkeonspark_b200 import datagen, parallel

model, D, opt = (sys.argv[1:4] + ["TransH", "100", "Adam"][len(sys.argv) - 1:])[:3] if len(sys.argv) > 1 else ("TransH", "100", "Adam")
D = int(D)
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world, rank = dist.get_world_size(), dist.get_rank()
g = datagen.make_shape("fb15k", seed=0)
con = okb.Config(private_context=True)
con.set_nbatches(100); con.set_dimension(D); con.set_opt_method(opt); con.workThreads = 8 * world
with contextlib.redirect_stdout(io.StringIO()):
    con.init_from_arrays(g.E, g.R, g.train, g.valid, g.test)
con.batch_size *= world
con._alloc_batch()
con.set_model_and_session(getattr(okb, model))
parallel.attach(con, mode="owner")
for _ in range(3):
    con.train_chunk_device(64)
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(4):
    losses = con.train_chunk_device(64)
b.record(); torch.cuda.synchronize()
us = a.elapsed_time(b) * 1e3 / 256
con.ctx.call("okb_prof_enable", 1)
con.train_chunk_device(64)
torch.cuda.synchronize()
con.ctx.call("okb_prof_enable", 0)
br = {}
for nm, kid in (("sample", 0), ("plan", 1), ("grad", 2), ("push", 6), ("owner", 7)):
    t, c = ctypes.c_double(), ctypes.c_int64()
    con.ctx.call("okb_prof_read", kid, ctypes.byref(t), ctypes.byref(c))
    br[nm] = "%.1f us x%d" % (t.value * 1e3 / max(c.value, 1), c.value)
print("rank %d/%d %s D=%d %s global B=%d: %.1f us/step, %.3g triples/s | %s | loss %.4f" %
      (rank, world, model, D, opt, con.batch_size, us, con.batch_size / us * 1e6, br, float(losses[-1])), flush=True)
dist.barrier()
con._world.close(con)
dist.destroy_process_group()
