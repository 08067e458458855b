"""Dataset preparation — the reference's split/generate.py as a function (file side only, no GPU).

From two N-Triples files (triples of the target relation/s, all other triples) and an optional tab-indented class
hierarchy it assigns numerical ids, cuts the sorted triple list into `n_batches` contiguous batches, splits each batch
into train / test / valid (test and valid hold target-relation triples only), and writes, under `out_dir/<n_batches>/`:

    0/entity2id.txt relation2id.txt train2id.txt test2id.txt valid2id.txt ontology_constrain.txt model/
    b/batchEntity2id.txt batch2id.txt batchTest2id.txt batchValid2id.txt ontology_constrain.txt model/      (b >= 1)

i.e. exactly the files `incremental.feed_batch` merges and `Config.init()` reads.  Behaviour follows
/root/reference/split/generate.py: ids by first appearance in the lexicographically sorted triple lines (head, tail,
relation; :141-166), batch b = triples [b*floor(n/N), (b+1)*floor(n/N)) with the last batch taking the remainder
(:186-206), an entity belongs to the first batch it appears in (:214-225), the ontology file of batch b lists the
classes whose id is <= the largest entity id seen so far, all ancestors as super-classes and all descendants as
sub-classes (:96-127, :243-276), test/valid sizes are int(#target triples * pct / 100) of the shuffled batch (:281-307).
The reference shuffles with an unseeded `random.shuffle`; here a seed can be given.
"""
from __future__ import annotations

import math
import os
import random

RDF_TYPE = "<http://www.w3.org/1999/02/22-rdf-syntax-ns#type>"


def _read_nt(path):
    with open(path, "r") as f:
        return [l.replace(" .\n", "").rstrip("\n") for l in f if l.strip()]


def _spo(line):
    a = line.split(" ")
    return a[0].strip(), a[1].strip(), a[2].strip()


def read_class_hierarchy(path, short_to_iri):
    """Tab-indented tree (generate.py:96-127): {class IRI: {"sup": ancestors, "sub": descendants}}; classes that do not
    occur as a type in the dataset are dropped with a warning."""
    classes, levels, missing = {}, {}, []
    with open(path, "r") as f:
        lines = [l.rstrip("\n") for l in f if l.strip()]
    for line in lines:
        level, name = line.count("\t"), line.strip()
        if name not in short_to_iri:
            missing.append(name)
            continue
        cur = short_to_iri[name]
        levels[level] = cur
        classes[cur] = {"sup": set(), "sub": set()}
        for j in range(level):
            if j in levels:
                classes[cur]["sup"].add(levels[j])
                classes[levels[j]]["sub"].add(cur)
    return classes, missing


def relation_structure(train, valid, test):
    """1-1 / 1-N / N-1 / N-N counts per file (generate.py:377-498): tails per (h, r) and heads per (r, t) over all three."""
    lef, rig = {}, {}
    for h, t, r in list(train) + list(valid) + list(test):
        lef.setdefault((h, r), []).append(t)
        rig.setdefault((r, t), []).append(h)
    rellef, totlef, relrig, totrig = {}, {}, {}, {}
    for (h, r), v in lef.items():
        rellef[r] = rellef.get(r, 0) + len(v)
        totlef[r] = totlef.get(r, 0) + 1.0
    for (r, t), v in rig.items():
        relrig[r] = relrig.get(r, 0) + len(v)
        totrig[r] = totrig.get(r, 0) + 1.0
    out = {}
    for name, rows in (("train", train), ("test", test), ("valid", valid)):
        c = {"1-1": 0, "1-N": 0, "N-1": 0, "N-N": 0}
        for h, t, r in rows:
            rign, lefn = rellef[r] / totlef[r], relrig[r] / totrig[r]
            c[("N" if lefn > 1.5 else "1") + "-" + ("N" if rign > 1.5 else "1")] += 1
        out[name] = c
    return out


def _write_triples(path, rows):
    with open(path, "w") as f:
        f.write(str(len(rows)) + "\n")
        f.writelines("%d %d %d\n" % r for r in rows)


def generate(path_target, path_rest, out_dir, n_batches=10, target_relations=(RDF_TYPE,), class_hierarchy=None,
             skip_data_property=True, test_pct=10, valid_pct=10, seed=None):
    """Write the batch folders and return a summary dict (counts, per-batch sizes and relation-structure statistics)."""
    rng = random.Random(seed)
    lines_t = _read_nt(path_target)
    type_tails = {_spo(l)[2] for l in lines_t}
    short_to_iri = {t.split("/")[-1].replace(">", ""): t for t in type_tails}
    classes, missing = ({}, [])
    if class_hierarchy is not None:
        classes, missing = read_class_hierarchy(class_hierarchy, short_to_iri)
    triples, skipped = list(lines_t), 0
    for l in _read_nt(path_rest):
        if skip_data_property and not _spo(l)[2].startswith("<"):       # literal tail = data property (:139-143)
            skipped += 1
            continue
        triples.append(l)
    triples.sort()
    ent, rel = {}, {}
    for l in triples:
        h, r, t = _spo(l)
        for e in (h, t):
            if e not in ent:
                ent[e] = len(ent)
        if r not in rel:
            rel[r] = len(rel)
    ids = [(ent[h], ent[t], rel[r]) for h, r, t in map(_spo, triples)]           # file order is h t r
    is_target = [_spo(l)[1] in target_relations for l in triples]

    n, bs = len(triples), int(math.floor(len(triples) / n_batches))
    root = os.path.join(out_dir, str(n_batches))
    seen, max_id, lef, summary = set(), -1, 0, {"entities": len(ent), "relations": len(rel), "triples": n,
                                                   "skipped_literals": skipped, "missing_classes": missing, "batches": []}
    for b in range(n_batches):
        d = os.path.join(root, str(b))
        os.makedirs(os.path.join(d, "model"), exist_ok=True)
        rig = n if b + 1 == n_batches else lef + bs
        idx = list(range(lef, rig))
        new_entities = []
        for i in idx:
            h, _, t = _spo(triples[i])
            for e in (h, t):
                if e not in seen:
                    seen.add(e)
                    new_entities.append(e)
        if new_entities:
            max_id = max(max_id, max(ent[e] for e in new_entities))
        with open(os.path.join(d, "entity2id.txt" if b == 0 else "batchEntity2id.txt"), "w") as f:
            f.write(str(len(new_entities)) + "\n")
            f.writelines("%s\t%d\n" % (e, ent[e]) for e in new_entities)
        if class_hierarchy is not None:
            known = [c for c in classes if ent[c] <= max_id]
            with open(os.path.join(d, "ontology_constrain.txt"), "w") as f:
                f.write(str(len(known)) + "\n")
                for c in known:
                    for key in ("sup", "sub"):
                        lst = sorted(ent[v] for v in classes[c][key] if ent[v] <= max_id)
                        f.write("\t".join([str(ent[c]), str(len(lst))] + [str(v) for v in lst]) + "\n")
        rng.shuffle(idx)
        tgt = [i for i in idx if is_target[i]]
        rest = [i for i in idx if not is_target[i]]
        n_test, n_valid = int(len(tgt) * test_pct / 100), int(len(tgt) * valid_pct / 100)
        test = [ids[i] for i in tgt[:n_test]]
        valid = [ids[i] for i in tgt[n_test:n_test + n_valid]]
        train_idx = tgt[n_test + n_valid:] + rest
        rng.shuffle(train_idx)
        train = [ids[i] for i in train_idx]
        pre = "" if b == 0 else "batch"
        _write_triples(os.path.join(d, "train2id.txt" if b == 0 else "batch2id.txt"), train)
        _write_triples(os.path.join(d, "test2id.txt" if b == 0 else pre + "Test2id.txt"), test)
        _write_triples(os.path.join(d, "valid2id.txt" if b == 0 else pre + "Valid2id.txt"), valid)
        summary["batches"].append({"dir": d, "new_entities": len(new_entities), "train": len(train), "test": len(test),
                                   "valid": len(valid), "structure": relation_structure(train, valid, test)})
        lef = rig
    with open(os.path.join(root, "0", "relation2id.txt"), "w") as f:
        f.write(str(len(rel)) + "\n")
        f.writelines("%s\t%d\n" % kv for kv in rel.items())
    return summary
