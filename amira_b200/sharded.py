"""Host side of the multi-GPU build: one process per GPU, contiguous read shards in rank order.

The device work and the NCCL exchange live behind the C ABI (``amira_gmg_comm_init`` + a collective
``amira_gmg_build``, see include/amira_gmg.h and csrc/sharded.cuh).  This module is the plumbing around
it: which reads a rank owns, how the ncclUniqueId travels (``torch.distributed``, any backend), and
how rank-local exports are stitched back into the arrays a single-GPU build would have exported.
Upstream has no counterpart (its joblib fan-out, amira/graph_utils.py:105-124, is disabled)."""
from __future__ import annotations

import numpy as np


def shard_range(n_reads: int, rank: int, world: int):
    """reads [lo, hi) of rank: contiguous, in rank order, sizes differing by at most one"""
    base, extra = divmod(int(n_reads), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_csr(ids: np.ndarray, off: np.ndarray, rank: int, world: int, pos_start=None, pos_end=None):
    """this rank's slice of a global CSR (offsets rebased to zero)"""
    lo, hi = shard_range(len(off) - 1, rank, world)
    a, b = int(off[lo]), int(off[hi])
    o = (off[lo:hi + 1] - off[lo]).astype(np.int64)
    ps = None if pos_start is None else np.ascontiguousarray(pos_start[a:b])
    pe = None if pos_end is None else np.ascontiguousarray(pos_end[a:b])
    return np.ascontiguousarray(ids[a:b]), o, ps, pe


def init_comm(dg, group=None, device=None):
    """create the library's NCCL communicator: rank 0 makes the id, torch.distributed broadcasts it"""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.from_numpy(np.ascontiguousarray(dg.nccl_unique_id()).copy())
    if dist.get_backend(group) == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        uid = uid.to(dev)
    dist.broadcast(uid, dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    dg.comm_init(uid.cpu().numpy(), rank, world)
    return rank, world


REPLICATED = ("node_key", "node_cov", "node_dir", "node_comp", "fw_off", "fw_edges", "bw_off", "bw_edges",
              "edge_src", "edge_tgt", "edge_sd", "edge_td", "edge_cov")
PER_WINDOW = ("win_node", "win_dir", "win_start", "win_end")
PER_READ = ("is_short", "to_correct")


def assemble_arrays(pieces: list) -> dict:
    """rank-ordered exports of a sharded build -> the arrays of the equivalent single-GPU build

    Node / edge tables are replicated (checked); per-read and per-window arrays concatenate in rank
    order; a node's read list is the concatenation of its per-rank lists (shards are contiguous, so
    the result is ascending, as upstream's first-touch order is)."""
    first = pieces[0]
    out = {"k": first["k"]}
    for f in REPLICATED:
        for p in pieces[1:]:
            if not np.array_equal(first[f], p[f]):
                raise ValueError("replicated field %s differs between ranks" % f)
        out[f] = first[f]
    for f in PER_WINDOW + PER_READ:
        if f in first:
            out[f] = np.concatenate([p[f] for p in pieces])
    shift, offs = 0, [np.zeros(1, np.int64)]
    for p in pieces:
        offs.append(p["win_off"][1:] + shift)
        shift += int(p["win_off"][-1])
    out["win_off"] = np.concatenate(offs)
    n = len(first["node_cov"])
    counts = np.zeros(n, np.int64)
    for p in pieces:
        counts += np.diff(p["node_reads_off"])
    roff = np.zeros(n + 1, np.int64)
    np.cumsum(counts, out=roff[1:])
    reads = np.empty(int(roff[-1]), np.int32)
    cursor = roff[:-1].copy()
    for p in pieces:
        c = np.diff(p["node_reads_off"])
        dst = np.repeat(cursor, c) + (np.arange(int(c.sum()), dtype=np.int64) - np.repeat(p["node_reads_off"][:-1], c))
        reads[dst] = p["node_reads"]
        cursor += c
    out["node_reads_off"], out["node_reads"] = roff, reads
    return out
