"""Digests of a (possibly sharded) build for the full-size parity checks of bench.py -- TEST INFRASTRUCTURE ONLY.

The C oracle needs 4 s .. 2 min for the bench workloads (1.25M .. 10M reads), too long to repeat inside every
bench run at 2, 4 and 8 GPUs, so its results are committed as SHA-256 digests (tests/golden/bench_digests.json,
written by scripts/make_bench_digests.py) in a form that does not depend on how an implementation distributes
the graph:

  global   the node / edge / adjacency tables of the whole read set, in upstream's insertion orders
  rank r   what belongs to the r-th contiguous read shard: its per-read lists (window offsets rebased to the
           shard, GLOBAL node indices) and its share of every node's read list as a CSR over the global nodes
           (global read indices)
"""
from __future__ import annotations

import hashlib

import numpy as np

GLOBAL_FIELDS = ("node_key", "node_cov", "node_dir", "node_comp", "fw_off", "fw_edges", "bw_off", "bw_edges",
                 "edge_src", "edge_tgt", "edge_sd", "edge_td", "edge_cov")
RANK_FIELDS = ("win_off", "win_node", "win_dir", "is_short", "to_correct", "node_reads_off", "node_reads")
_DT = {"node_key": np.int32, "node_cov": np.uint32, "node_dir": np.int8, "node_comp": np.uint32, "fw_off": np.int64,
       "fw_edges": np.int32, "bw_off": np.int64, "bw_edges": np.int32, "edge_src": np.int32, "edge_tgt": np.int32,
       "edge_sd": np.int8, "edge_td": np.int8, "edge_cov": np.uint32, "win_off": np.int64, "win_node": np.int32,
       "win_dir": np.int8, "is_short": np.uint8, "to_correct": np.uint8, "node_reads_off": np.int64,
       "node_reads": np.int32}


def _sha(a, f):
    return hashlib.sha256(np.ascontiguousarray(np.asarray(a).astype(_DT[f], copy=False)).tobytes()).hexdigest()


def digest_fields(arrays: dict, fields) -> dict:
    return {f: _sha(arrays[f], f) for f in fields}


def global_digest(arrays: dict) -> dict:
    d = digest_fields(arrays, GLOBAL_FIELDS)
    d["nodes"], d["edges"] = int(len(arrays["node_cov"])), int(len(arrays["edge_cov"]))
    return d


def rank_digest(piece: dict) -> dict:
    d = digest_fields(piece, RANK_FIELDS)
    d["windows"], d["incidences"] = int(len(piece["win_node"])), int(len(piece["node_reads"]))
    return d


def slice_rank(a: dict, read_lo: int, read_hi: int) -> dict:
    """the part of a whole-read-set build (oracle arrays) that belongs to reads [read_lo, read_hi)"""
    w0, w1 = int(a["win_off"][read_lo]), int(a["win_off"][read_hi])
    out = {"win_off": a["win_off"][read_lo:read_hi + 1] - a["win_off"][read_lo],
           "win_node": a["win_node"][w0:w1], "win_dir": a["win_dir"][w0:w1],
           "is_short": a["is_short"][read_lo:read_hi], "to_correct": a["to_correct"][read_lo:read_hi]}
    n = len(a["node_cov"])
    reads = a["node_reads"]
    mine = (reads >= read_lo) & (reads < read_hi)
    # reads are ascending inside every node's list: the shard's entries of a node are contiguous
    csum = np.zeros(len(reads) + 1, np.int64)
    np.cumsum(mine, out=csum[1:])
    roff = csum[a["node_reads_off"]]
    assert len(roff) == n + 1
    out["node_reads_off"] = roff
    out["node_reads"] = reads[mine]
    return out


def diff_digests(got: dict, want: dict) -> list:
    return sorted(k for k in want if got.get(k) != want[k])
