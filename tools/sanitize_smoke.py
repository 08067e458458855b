"""Small invocations of every hot kernel (all four models, SGD/Adam, k = 1 / several negatives, relation negatives, odd D,
host-buffer step, chunked steps, evaluation) — sized for `compute-sanitizer --tool memcheck|racecheck python
tools/sanitize_smoke.py` where the pool allows it (it does not on the round-1 boxes), otherwise a quick all-kernels run."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import openkeonspark_b200 as okb
from openkeonspark_b200 import datagen
from conftest import make_params
d = tempfile.mkdtemp() + "/"
datagen.write_dataset(datagen.make_shape("tiny", seed=3, dup_train=20), d, ontology=True)
for model, opt, k, kr, D in (("TransE", "SGD", 1, 0, 50), ("TransH", "Adam", 1, 0, 100), ("TransD", "Adam", 4, 1, 100),
                             ("TransH", "SGD", 3, 0, 33), ("TransR", "Adam", 2, 0, 20)):
    con = okb.Config(private_context=True)
    con.set_in_path(d); con.set_nbatches(4); con.set_ent_neg_rate(k); con.set_rel_neg_rate(kr); con.set_opt_method(opt)
    con.set_dimension(D); con.set_test_link_prediction(True); con.set_test_triple_classification(True); con.set_test_head(1)
    con.init(); con.set_model_and_session(getattr(okb, model))
    con.set_parameters(make_params(model, con.entTotal, con.relTotal, D, seed=1))
    con.sampling(); con.train_step(con.batch_h, con.batch_t, con.batch_r, con.batch_y)
    con.plan_ahead = 3
    for _ in range(4):
        con.next_step_device()
    con.train_chunk_device(3)
    con.test()
    torch.cuda.synchronize()
    print("ok", model, opt, k, kr, D, flush=True)
