"""Link-prediction metric accumulation — the output contract of the reference's evaluation loop.

The reference turns each query's 8-int record (base/Test.h:99-134) into ~30 dictionary updates
(distribute_training.py:375-420 init, :477-527 tail side, :541-590 head side) and the driver divides
every accumulator by testTotal and prints two tables (main_spark.py:447-469).  Same keys, same
arithmetic, vectorised over all queries.
"""
from __future__ import annotations

import numpy as np

_SIDES = {"r": 1, "l": 0}      # r = tail replaced (testTail), l = head replaced (testHead)
_VARIANTS = [("", 0), ("_filter", 1), ("_constrain", 2), ("_filter_constrain", 3)]


def _key(side, stem, variant):
    # e.g. r_tot, r_filter_tot, r_tot_constrain, r_filter_tot_constrain   (distribute_training.py:376-386)
    filt = "_filter" if "filter" in variant else ""
    cons = "_constrain" if "constrain" in variant else ""
    return "%s%s%s%s" % (side, filt, stem, cons)


def empty_metrics(test_head):
    d = {}
    for side in (("r", "l") if test_head else ("r",)):
        for variant, _ in _VARIANTS:
            for stem in ("_tot", "_rank", "_reci_rank", "_mis_err", "_spec_err", "_gen_err"):
                d[_key(side, stem, variant)] = 0.0
            for n in ("1", "3"):
                d[_key(side + n, "_tot", variant)] = 0.0
    return d


def accumulate(records, test_head, d=None):
    """records: int64 [n_queries, 2, 8] (side 0 = head, 1 = tail).  Returns the SUMS (not yet divided)."""
    if d is None:
        d = empty_metrics(test_head)
    rec = np.asarray(records, dtype=np.int64)
    for side, si in _SIDES.items():
        if side == "l" and not test_head:
            continue
        for variant, vi in _VARIANTS:
            s = rec[:, si, vi]
            cls = rec[:, si, 4 + vi]
            d[_key(side, "_tot", variant)] += float((s < 10).sum())              # hits@10
            d[_key(side + "3", "_tot", variant)] += float((s < 3).sum())         # hits@3
            hit1 = s < 1
            d[_key(side + "1", "_tot", variant)] += float(hit1.sum())            # hits@1
            miss = ~hit1                                                          # ontology class of the top-1 error
            d[_key(side, "_gen_err", variant)] += float((miss & (cls == 1)).sum())
            d[_key(side, "_spec_err", variant)] += float((miss & (cls == 2)).sum())
            d[_key(side, "_mis_err", variant)] += float((miss & (cls != 1) & (cls != 2)).sum())
            d[_key(side, "_rank", variant)] += float((1 + s).sum())              # MR
            d[_key(side, "_reci_rank", variant)] += float(np.divide(1.0, 1 + s).sum())   # MRR
    return d


def finalize(d, test_total):
    """main_spark.py:447-448: every accumulator divided by testTotal."""
    return {k: float(np.divide(v, test_total)) for k, v in d.items()}


def format_table(d, test_head):
    """The two tables of main_spark.py:451-469."""
    hdr = "{:<20}" * 9
    row = "{:<20}" + "{:<20.5f}" * 8
    cols = ("_reci_rank", "_rank", "_tot", "3_tot", "1_tot", "_gen_err", "_spec_err", "_mis_err")

    def vals(side, variant):
        out = []
        for c in cols:
            if c in ("3_tot", "1_tot"):
                out.append(d[_key(side + c[0], "_tot", variant)])
            else:
                out.append(d[_key(side, c, variant)])
        return out

    lines = ["", " ========== LINK PREDICTION RESULTS =========="]
    for title, base in (("No type constraint results:", ""), ("Type constraint results:", "_constrain")):
        lines.append(title)
        lines.append(hdr.format("metric", "MRR", "MR", "hit@10", "hit@3", "hit@1", "hit@1GenError", "hit@1SpecError", "hit@1MisError"))
        for label, variant in (("raw", base), ("filter", "_filter" + base)):
            r = vals("r", variant)
            if test_head:
                l = vals("l", variant)
                lines.append(row.format("l(%s):" % label, *l))
            lines.append(row.format("r(%s):" % label, *r))
            if test_head:
                lines.append(row.format("mean(%s):" % label, *[np.divide(a + b, 2) for a, b in zip(l, r)]))
                lines.append("")
    return "\n".join(lines)
