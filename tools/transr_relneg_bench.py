import sys; sys.path.insert(0,'.')
import torch, bench
cfg=dict(bench.CONFIGS[3]); g=bench.graph(cfg["shape"])
import openkeonspark_b200 as okb, contextlib, io, numpy as np, ctypes
con=okb.Config(private_context=True)
con.set_nbatches(cfg["nbatches"]); con.set_ent_neg_rate(1); con.set_rel_neg_rate(1); con.set_margin(1.0); con.set_alpha(0.001)
con.set_opt_method("SGD"); con.set_dimension(100); con.set_bern(0); con.workThreads=8
with contextlib.redirect_stdout(io.StringIO()):
    con.init_from_arrays(g.E,g.R,g.train,None,None)
con.set_model_and_session(okb.TransR)
con.plan_ahead=8
for _ in range(2): con.train_chunk_device()
torch.cuda.synchronize()
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(4): con.train_chunk_device()
b.record(); torch.cuda.synchronize()
us=a.elapsed_time(b)/32*1e3
print("TransR D=100 k=1 kr=1 B=%d: %.1f us/step = %.2f M triples/s"%(con.batch_size,us,con.batch_size/us))
