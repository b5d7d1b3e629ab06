"""shared test helpers: golden input loading, stage application"""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

with open(os.path.join(GOLD, "expected.json")) as _f:
    EXPECTED = json.load(_f)
GOLDEN_KEYS = sorted(k for k in EXPECTED if not k.startswith("_"))
STAGE_ORDER = ("build", "rlcc5", "rlcc5_filter3_1")


def load_input(name):
    z = np.load(os.path.join(GOLD, "inputs", name + ".npz"))
    ids = z["ids"].astype(np.int32)
    off = z["off"].astype(np.int64)
    ps = z["pos_start"].astype(np.int32) if "pos_start" in z.files else None
    pe = z["pos_end"].astype(np.int32) if "pos_end" in z.files else None
    return [str(x) for x in z["vocab"]], ids, off, ps, pe


def apply_stage(graph, stage):
    """graph: anything with remove_low_coverage_components / filter_graph"""
    if stage == "rlcc5":
        graph.remove_low_coverage_components(5)
    elif stage == "rlcc5_filter3_1":
        graph.filter_graph(3, 1)


def read_dict(vocab, ids, off, ps=None, pe=None):
    toks = [("+" if x > 0 else "-") + vocab[abs(int(x)) - 1] for x in ids.tolist()]
    reads = {"read%09d" % i: toks[off[i]:off[i + 1]] for i in range(len(off) - 1)}
    pos = None
    if ps is not None:
        pos = {r: [[int(ps[j]), int(pe[j])] for j in range(off[i], off[i + 1])] for i, r in enumerate(reads)}
    return reads, pos
