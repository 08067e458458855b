"""Dataset preparation (openkeonspark_b200.prep = the reference's split/generate.py) and its hand-over to the
incremental-batch file merge: CPU only."""
import os
import shutil

import numpy as np

TYPE = "<http://www.w3.org/1999/02/22-rdf-syntax-ns#type>"


def _nt(path, triples):
    with open(path, "w") as f:
        f.writelines("%s %s %s .\n" % t for t in triples)


def test_generate_batches_and_feed(tmp_path):
    from openkeonspark_b200 import incremental, prep
    from oracle import harness
    rng = np.random.default_rng(0)
    ents = ["<http://x/e%d>" % i for i in range(60)]
    classes = ["<http://x/o/Thing>", "<http://x/o/Agent>", "<http://x/o/Person>", "<http://x/o/Place>"]
    tgt = [(e, TYPE, classes[int(rng.integers(0, 4))]) for e in ents]
    rest = [(ents[int(rng.integers(0, 60))], "<http://x/p%d>" % int(rng.integers(0, 5)), ents[int(rng.integers(0, 60))]) for _ in range(240)]
    rest += [(ents[3], "<http://x/name>", '"a literal"')] * 4                       # data properties are skipped (:139-143)
    _nt(tmp_path / "t.nt", tgt)
    _nt(tmp_path / "r.nt", rest)
    (tmp_path / "h.txt").write_text("Thing\n\tAgent\n\t\tPerson\n\tPlace\n\tGhost\n")   # Ghost never occurs: dropped with a warning
    out = str(tmp_path / "out")
    s = prep.generate(str(tmp_path / "t.nt"), str(tmp_path / "r.nt"), out, n_batches=3, class_hierarchy=str(tmp_path / "h.txt"), seed=1)
    n = len(tgt) + len(set()) + 240
    assert s["triples"] == n and s["skipped_literals"] == 4 and s["missing_classes"] == ["Ghost"]
    assert s["relations"] == 6 and s["entities"] == len({x for t in tgt + rest[:240] for x in (t[0], t[2])})
    # the batches partition the triples; test/valid hold target-relation triples only; every id is in range
    tot = 0
    rel = dict(l.strip().split("\t") for l in open(os.path.join(out, "3", "0", "relation2id.txt")).readlines()[1:])
    type_id = int(rel[TYPE])
    seen_entities = 0
    for b, info in enumerate(s["batches"]):
        d = info["dir"]
        names = ("train2id.txt", "test2id.txt", "valid2id.txt", "entity2id.txt") if b == 0 else \
                ("batch2id.txt", "batchTest2id.txt", "batchValid2id.txt", "batchEntity2id.txt")
        tr, te, va = (harness.read_triples(os.path.join(d, f)) for f in names[:3])
        assert (tr.shape[0], te.shape[0], va.shape[0]) == (info["train"], info["test"], info["valid"])
        tot += tr.shape[0] + te.shape[0] + va.shape[0]
        assert all(x.size == 0 or (x[:, 2] == type_id).all() for x in (te, va))
        seen_entities += int(open(os.path.join(d, names[3])).readline())
        for x in (tr, te, va):
            assert x.size == 0 or (x[:, :2].max() < seen_entities and x[:, 2].max() < 6)      # only entities introduced so far
        assert os.path.isdir(os.path.join(d, "model"))
        k, sup_l, sup_ids, sub_l, sub_ids = harness.read_lists(os.path.join(d, "ontology_constrain.txt"))
        assert k.size <= 4 and (k < seen_entities).all()
        assert sum(info["structure"][f][c] for f in ("train", "test", "valid") for c in ("1-1", "1-N", "N-1", "N-N")) == \
            info["train"] + info["test"] + info["valid"]
    assert tot == n and seen_entities == s["entities"]
    # Person's super-classes are Agent and Thing, Thing's sub-classes are everything below it (last batch: all classes known)
    ent = {}
    for b in range(3):
        f = os.path.join(out, "3", str(b), "entity2id.txt" if b == 0 else "batchEntity2id.txt")
        ent.update(dict((l.split("\t")[0], int(l.split("\t")[1])) for l in open(f).readlines()[1:]))
    lines = [l.split() for l in open(os.path.join(out, "3", "2", "ontology_constrain.txt")).readlines()[1:]]
    rows = {}
    for sup, sub in zip(lines[0::2], lines[1::2]):
        rows[int(sup[0])] = (set(map(int, sup[2:])), set(map(int, sub[2:])))
    used = {c for _, _, c in tgt}
    if classes[2] in used:
        want = {ent[c] for c in (classes[0], classes[1]) if c in used}
        assert rows[ent[classes[2]]][0] == want
    assert rows[ent[classes[0]]][1] == {ent[c] for c in classes[1:] if c in used}
    # batch 1 handed to the incremental merge: counts add up, new triples are the LAST rows of train2id.txt
    d0 = os.path.join(out, "3", "0")
    for f in ("batch2id.txt", "batchEntity2id.txt", "batchTest2id.txt", "batchValid2id.txt"):
        shutil.copy(os.path.join(out, "3", "1", f), d0)
    assert incremental.is_new_batch(d0)
    r = incremental.feed_batch(d0)
    b0, b1 = s["batches"][0], s["batches"][1]
    assert r == {"new_entities": b1["new_entities"], "final_entities": b0["new_entities"] + b1["new_entities"],
                 "new_train": b1["train"], "new_test": b1["test"], "new_valid": b1["valid"]}
    tr = harness.read_triples(os.path.join(d0, "train2id.txt"))
    assert tr.shape[0] == b0["train"] + b1["train"]
    assert np.array_equal(tr[b0["train"]:], harness.read_triples(os.path.join(out, "3", "1", "batch2id.txt")))
