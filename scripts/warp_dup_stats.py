"""How much would warp-aggregated atomics save in k_insert_windows?  (VERDICT r1, row g)

In the insert kernel the 32 lanes of a warp instruction handle the windows that START at 32 consecutive call
positions (chunk base + 32 * iteration + lane).  A warp-aggregated atomic (`__match_any_sync` on the slot, one
atomic per group) saves exactly the lanes whose window hits a node -- or whose window pair hits an undirected edge --
that another lane of the same instruction also hits.  This script counts those on the oracle's build of the bench
workloads: duplicates per warp instruction = (valid lanes) - (distinct targets), summed over all instructions.

    python scripts/warp_dup_stats.py > profiles/r2_warp_duplicates.md
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amira_b200 import synth  # noqa: E402
from oracle import c_oracle  # noqa: E402


def stats(name, n_reads, k):
    ids, off = synth.generate(synth.CONFIGS[name], 0, n_reads)
    a = c_oracle.COracleGraph(ids, off, k).arrays()
    nwin = np.diff(a["win_off"])
    read_of = np.repeat(np.arange(len(nwin)), nwin)
    pos = off[:-1][read_of] + (np.arange(len(read_of)) - a["win_off"][:-1][read_of])     # call position of every window
    grp = pos // 32
    node = a["win_node"].astype(np.int64)
    W = len(node)
    n_instr = len(np.unique(grp))
    distinct = len(np.unique(grp * (node.max() + 1) + node))
    # adjacent pairs of one read: the pair is handled by the lane of its first window
    same = read_of[1:] == read_of[:-1]
    a_, b_ = node[:-1][same], node[1:][same]
    rel = (a["win_dir"][:-1][same].astype(np.int64) * a["win_dir"][1:][same] > 0).astype(np.int64)
    ekey = (np.minimum(a_, b_) * (node.max() + 1) + np.maximum(a_, b_)) * 2 + rel
    egrp = grp[:-1][same]
    P = len(ekey)
    order = np.lexsort((ekey, egrp))
    eg, ek = egrp[order], ekey[order]
    edistinct = 1 + int(np.count_nonzero((eg[1:] != eg[:-1]) | (ek[1:] != ek[:-1]))) if P else 0
    return {"workload": "%s, %d reads, k=%d" % (name, n_reads, k), "windows": W, "warp_instr": n_instr,
            "node_dups": W - distinct, "node_dup_frac": (W - distinct) / W, "pairs": P, "edge_dups": P - edistinct,
            "edge_dup_frac": (P - edistinct) / max(P, 1)}


if __name__ == "__main__":
    rows = [stats("c5", 300_000, 5), stats("c2", 50_000, 3), stats("c3", 200_000, 3), stats("c4", 300_000, 3)]
    print("# Duplicate atomic targets inside one warp instruction of k_insert_windows (r2)\n")
    print("`python scripts/warp_dup_stats.py` on the C oracle's build (exact: lanes of an instruction = windows starting at 32 "
          "consecutive call positions).\nA warp-aggregated atomic saves exactly the duplicate lanes.\n")
    print("| workload | windows | warp instructions | duplicate node targets | share | adjacent pairs | duplicate edge targets | share |")
    print("|---|---|---|---|---|---|---|---|")
    for r in rows:
        print("| %s | %d | %d | %d | %.3f%% | %d | %d | %.3f%% |" % (r["workload"], r["windows"], r["warp_instr"], r["node_dups"],
              100 * r["node_dup_frac"], r["pairs"], r["edge_dups"], 100 * r["edge_dup_frac"]))
