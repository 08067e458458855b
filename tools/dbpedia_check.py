"""DBpedia-shaped scale check (BASELINE.json configs[4]): TransE D=200, ~4 M entities / 600 relations / 20 M triples.
Loads from arrays, trains a few chunks (auto batch rule -> B=2000, and a large batch), ranks a slice of the test set."""
import contextlib, ctypes, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import openkeonspark_b200 as okb
from openkeonspark_b200 import datagen

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
E, R, N = int(4_000_000 * scale), 600, int(20_000_000 * scale)
t0 = time.time()
rng = np.random.default_rng(0)
need = N + 200_000
raw = np.stack([rng.integers(0, E, need + need // 50), rng.integers(0, E, need + need // 50), rng.integers(0, R, need + need // 50)], 1)
key = (raw[:, 0] * E + raw[:, 1]) * R + raw[:, 2]
_, first = np.unique(key, return_index=True)
raw = raw[np.sort(first)][:need]
train, valid, test = raw[:N], raw[N:N + 100_000], raw[N + 100_000:]
print("graph: E=%d R=%d train=%d (%.1fs)" % (E, R, train.shape[0], time.time() - t0), flush=True)
for nb, label in ((0, "auto batch rule"), (40, "large batch")):
    con = okb.Config(private_context=True)
    con.set_nbatches(nb); con.set_dimension(200); con.set_opt_method("SGD"); con.workThreads = 8
    con.test_head = 1
    t1 = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        con.init_from_arrays(E, R, train, valid, test)
    t2 = time.time()
    con.set_model_and_session(okb.TransE)
    n = min(con.plan_ahead, 16)
    con.train_chunk_device(n); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); losses = con.train_chunk_device(n); b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / losses.numel()
    print("%s: B=%d nbatches=%d load %.1fs | %.1f us/step %.3g triples/s loss %.4f | mem %.1f GB" %
          (label, con.batch_size, con.nbatches, t2 - t1, us, con.batch_size / us * 1e6, float(losses[-1]), torch.cuda.memory_allocated() / 2**30), flush=True)
    if nb == 0:
        nq = 256
        con.link_prediction_records(0, 8); torch.cuda.synchronize()
        a.record(); rec = con.link_prediction_records(0, nq); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print("link prediction: %d queries (both sides) over %d candidates: %.1f ms -> %.1f q/s; mean filtered tail rank %.1f" %
              (2 * nq, E, ms, 2 * nq / ms * 1e3, float(rec[:, 1, 1].float().mean())), flush=True)
    del con
    torch.cuda.empty_cache()
