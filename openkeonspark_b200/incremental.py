"""Incremental batches — the file side of main_spark.py:100-195 (feed_batch and friends).

A new batch arrives as four files next to the dataset: batch2id.txt (new train triples), batchEntity2id.txt (new
entities), batchTest2id.txt, batchValid2id.txt.  feed_batch() appends them to train2id.txt / entity2id.txt /
test2id.txt / valid2id.txt and fixes the counts in the first lines; batch2id.txt stays in place during training because
its first line is what switches the sampler to "positives from the last newBatchTotal rows" (Reader.h:61-67,
Base.cpp:101-103); remove_batch_files() deletes the four files afterwards.  The model side (new rows in the entity
tables) is Config.grow_entities()."""
from __future__ import annotations

import os

TRIPLES, ENTITIES, TEST, VALID = "train2id.txt", "entity2id.txt", "test2id.txt", "valid2id.txt"
NEW_TRIPLES, NEW_ENTITIES, NEW_TEST, NEW_VALID = "batch2id.txt", "batchEntity2id.txt", "batchTest2id.txt", "batchValid2id.txt"


def is_new_batch(path):
    """main_spark.py:163-170"""
    return all(os.path.isfile(os.path.join(path, f)) for f in (NEW_TRIPLES, NEW_ENTITIES, NEW_TEST, NEW_VALID))


def _append(path, target, batch):
    """main_spark.py:100-133 (update_triples); returns the number of appended lines"""
    with open(os.path.join(path, batch)) as f:
        n = int(f.readline().strip())
        new = [f.readline() for _ in range(n)]
    if n > 0:
        with open(os.path.join(path, target)) as f:
            lines = f.readlines()
        lines[0] = str(int(lines[0]) + n) + "\n"
        if not lines[-1].endswith("\n"):
            lines[-1] += "\n"
        with open(os.path.join(path, target), "w") as f:
            f.writelines(lines + new)
    return n


def feed_batch(path):
    """Merge the four batch files into the dataset (main_spark.py:136-160).  Returns a dict with the number of new
    entities, train / test / valid triples and the final entity count (what update_entities_and_model returns)."""
    n_ent = _append(path, ENTITIES, NEW_ENTITIES)
    with open(os.path.join(path, ENTITIES)) as f:
        final = int(f.readline().strip())
    return {"new_entities": n_ent, "final_entities": final, "new_train": _append(path, TRIPLES, NEW_TRIPLES),
            "new_test": _append(path, TEST, NEW_TEST), "new_valid": _append(path, VALID, NEW_VALID)}


def remove_batch_files(path):
    """main_spark.py:173-180"""
    for f in (NEW_TRIPLES, NEW_ENTITIES, NEW_TEST, NEW_VALID):
        os.remove(os.path.join(path, f))
