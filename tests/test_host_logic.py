"""Host-side logic that needs no GPU: metric accumulation, batch geometry, dataset writer."""
import numpy as np


def reference_style_accumulate(records, test_head):
    """Literal per-triple restatement of distribute_training.py:477-590 for cross-checking."""
    from openkeonspark_b200 import metrics
    d = metrics.empty_metrics(test_head)
    for rec in records:
        for side, si in (("r", 1), ("l", 0)):
            if side == "l" and not test_head:
                continue
            s, f, sc, fc, m, fm, cm, fcm = [int(x) for x in rec[si]]
            if f < 10: d[side + "_filter_tot"] += 1
            if s < 10: d[side + "_tot"] += 1
            if f < 3: d[side + "3_filter_tot"] += 1
            if s < 3: d[side + "3_tot"] += 1
            if fc < 10: d[side + "_filter_tot_constrain"] += 1
            if sc < 10: d[side + "_tot_constrain"] += 1
            if fc < 3: d[side + "3_filter_tot_constrain"] += 1
            if sc < 3: d[side + "3_tot_constrain"] += 1
            for cnt, cls, pre, suf in ((f, fm, "_filter", ""), (s, m, "", ""), (fc, fcm, "_filter", "_constrain"), (sc, cm, "", "_constrain")):
                if cnt < 1: d[side + "1" + pre + "_tot" + suf] += 1
                elif cls == 1: d[side + pre + "_gen_err" + suf] += 1
                elif cls == 2: d[side + pre + "_spec_err" + suf] += 1
                else: d[side + pre + "_mis_err" + suf] += 1
            d[side + "_filter_rank"] += 1 + f
            d[side + "_rank"] += 1 + s
            d[side + "_filter_reci_rank"] += 1.0 / (1 + f)
            d[side + "_reci_rank"] += 1.0 / (1 + s)
            d[side + "_filter_rank_constrain"] += 1 + fc
            d[side + "_rank_constrain"] += 1 + sc
            d[side + "_filter_reci_rank_constrain"] += 1.0 / (1 + fc)
            d[side + "_reci_rank_constrain"] += 1.0 / (1 + sc)
    return d


def test_metric_keys_are_the_reference_contract():
    from openkeonspark_b200 import metrics
    d = metrics.empty_metrics(True)
    assert len(d) == 64 and len(metrics.empty_metrics(False)) == 32
    for k in ("r_tot", "r_filter_tot", "r_tot_constrain", "r_filter_tot_constrain", "r1_tot", "r3_filter_tot_constrain",
              "r_rank", "r_filter_reci_rank_constrain", "r_mis_err", "r_filter_spec_err", "l_gen_err_constrain", "l1_filter_tot"):
        assert k in d, k


def test_metric_accumulation_matches_per_triple_loop():
    from openkeonspark_b200 import metrics
    rng = np.random.default_rng(0)
    rec = np.zeros((300, 2, 8), np.int64)
    rec[:, :, :4] = rng.integers(0, 30, (300, 2, 4))
    rec[:, :, 4:] = rng.integers(0, 4, (300, 2, 4))
    for th in (True, False):
        a = metrics.accumulate(rec, th)
        b = reference_style_accumulate(rec, th)
        assert a.keys() == b.keys()
        for k in a:
            assert abs(a[k] - b[k]) < 1e-9, k
        fin = metrics.finalize(a, 300)
        assert abs(fin["r_tot"] - a["r_tot"] / 300) < 1e-12
        assert "LINK PREDICTION RESULTS" in metrics.format_table(fin, th)


def test_dataset_writer_roundtrip(tmp_path):
    from openkeonspark_b200 import datagen
    from oracle import harness
    g = datagen.make_shape("tiny", seed=0, dup_train=5)
    d = datagen.write_dataset(g, str(tmp_path) + "/", ontology=True)
    assert harness.read_count(d + "entity2id.txt") == g.E
    tr = harness.read_triples(d + "train2id.txt")
    assert np.array_equal(tr, g.train) and tr.shape[0] == 405
    keys, oa, fa, ob, fb = harness.read_lists(d + "type_constrain.txt")
    assert keys.size == len(set(np.concatenate([g.train, g.valid, g.test])[:, 2]))
    # every (h, r) of the graph has h in r's head list
    heads, tails = datagen.type_constraints(g)
    for h, t, r in g.test:
        assert h in heads[r] and t in tails[r]


def test_shapes_named_by_baseline():
    from openkeonspark_b200 import datagen
    assert datagen.SHAPES["fb15k"] == dict(E=14951, R=1345, n_train=483142, n_valid=50000, n_test=59071)
    assert datagen.SHAPES["wn18"]["E"] == 40943 and datagen.SHAPES["wn18"]["R"] == 18
