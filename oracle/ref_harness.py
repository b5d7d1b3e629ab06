"""Import the UNMODIFIED upstream Amira package -- TEST INFRASTRUCTURE ONLY.

From the read-only checkout (``/root/reference``, build container) or, where that does not exist (the GPU box),
from ``baseline/_ref/`` -- the git-ignored ``pip install --target`` copy made by ``baseline/install_ref.py`` that
travels with the repository snapshot.  Used by ``oracle/make_golden.py`` to generate ``tests/golden/`` and by
``bench.py`` to time upstream's own ``GeneMerGraph(readDict, k)`` on the host cores beside the GPU numbers.

``amira/construct_graph.py:9,11,19`` imports ``sourmash``, ``suffix_tree`` and (through
``graph_utils``) ``matplotlib``/``pysam`` at module import; none is called on the graph-build
path, so empty stand-in modules are registered before the import.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    cands = [os.environ.get("AMIRA_REFERENCE_ROOT"), "/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "amira", "construct_graph.py")):
            return c
    return cands[1]


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "amira", "construct_graph.py"))


def load():
    """returns the upstream ``amira.construct_graph`` module"""
    if not available():
        raise RuntimeError("upstream checkout not present at " + REFERENCE_ROOT)
    for name, attrs in (("sourmash", ()), ("suffix_tree", ("Tree",)), ("pysam", ()),
                        ("matplotlib", ()), ("matplotlib.pyplot", ())):
        if name not in sys.modules:
            try:
                __import__(name)
                continue
            except Exception:
                pass
            mod = types.ModuleType(name)
            for a in attrs:
                setattr(mod, a, type(a, (), {}))
            sys.modules[name] = mod
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import amira.construct_graph as cg  # noqa: E402
    return cg


def reference_arrays(graph, read_ids, vocabulary) -> dict:
    """Flatten an upstream ``GeneMerGraph`` object into graph arrays (same layout as the C-ABI exports)."""
    rank = {n: i + 1 for i, n in enumerate(vocabulary)}
    k = graph.get_kmerSize()
    nodes = list(graph.get_nodes().values())
    nidx = {h: i for i, h in enumerate(graph.get_nodes())}
    edges = list(graph.get_edges().values())
    eidx = {h: i for i, h in enumerate(graph.get_edges())}
    ridx = {r: i for i, r in enumerate(read_ids)}

    def csr(lists, conv):
        off = np.zeros(len(lists) + 1, np.int64)
        flat = []
        for i, l in enumerate(lists):
            flat.extend(conv(x) for x in l)
            off[i + 1] = len(flat)
        return off, np.asarray(flat, np.int32).reshape(-1)

    out = {"k": np.int32(k)}
    out["node_key"] = np.asarray(
        [[rank[g.get_name()] * g.get_strand() for g in n.get_canonical_geneMer()] for n in nodes],
        np.int32).reshape(len(nodes), max(k, 0))
    out["node_cov"] = np.asarray([n.get_node_coverage() for n in nodes], np.uint32)
    out["node_dir"] = np.asarray([n.get_geneMer().get_geneMerDirection() for n in nodes], np.int8)
    out["node_comp"] = np.asarray([n.get_component() for n in nodes], np.uint32)
    out["node_reads_off"], out["node_reads"] = csr([n.get_list_of_reads() for n in nodes], ridx.__getitem__)
    out["fw_off"], out["fw_edges"] = csr([n.get_forward_edge_hashes() for n in nodes], eidx.__getitem__)
    out["bw_off"], out["bw_edges"] = csr([n.get_backward_edge_hashes() for n in nodes], eidx.__getitem__)
    out["edge_src"] = np.asarray([nidx[e.get_sourceNode().__hash__()] for e in edges], np.int32)
    out["edge_tgt"] = np.asarray([nidx[e.get_targetNode().__hash__()] for e in edges], np.int32)
    out["edge_sd"] = np.asarray([e.get_sourceNodeDirection() for e in edges], np.int8)
    out["edge_td"] = np.asarray([e.get_targetNodeDirection() for e in edges], np.int8)
    out["edge_cov"] = np.asarray([e.get_edge_coverage() for e in edges], np.uint32)
    win_off = np.zeros(len(read_ids) + 1, np.int64)
    wn, wd, ws, we = [], [], [], []
    rn, rd, rp = graph.get_readNodes(), graph.get_readNodeDirections(), graph.get_readNodePositions()
    for i, rid in enumerate(read_ids):
        for h, d, p in zip(rn.get(rid, []), rd.get(rid, []), rp.get(rid, [])):
            wn.append(-1 if h is None else nidx[h])
            wd.append(0 if d is None else d)
            ws.append(-1 if p is None else p[0])
            we.append(-1 if p is None else p[1])
        win_off[i + 1] = len(wn)
    out["win_off"] = win_off
    out["win_node"] = np.asarray(wn, np.int32).reshape(-1)
    out["win_dir"] = np.asarray(wd, np.int8).reshape(-1)
    out["win_start"] = np.asarray(ws, np.int32).reshape(-1)
    out["win_end"] = np.asarray(we, np.int32).reshape(-1)
    short = graph.get_short_read_annotations()
    out["is_short"] = np.asarray([r in short for r in read_ids], np.uint8)
    tc = graph.get_reads_to_correct()
    out["to_correct"] = np.asarray([r in tc for r in read_ids], np.uint8)
    return out
