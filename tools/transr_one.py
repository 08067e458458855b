"""A few TransR train steps at the cfg-4 shape (for ncu): python tools/transr_one.py ROWS"""
import sys; sys.path.insert(0, '.')
import torch, bench
cfg = bench.CONFIGS[3]; g = bench.graph(cfg["shape"])
con, _ = bench.make_con(cfg, g, 1, 0, lp=False)
con.ctx.call("okb_set_flag", 15, int(sys.argv[1]) if len(sys.argv) > 1 else 0)
con.plan_ahead = 4
for _ in range(2):
    con.train_chunk_device()
torch.cuda.synchronize()
