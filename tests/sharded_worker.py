"""Worker of the multi-process tests (launched with torch.distributed.run, one process per rank).

    --backend nccl : the real thing, one GPU per rank: collective amira_gmg_build over the rank's shard
    --backend gloo : CPU box: the device is replaced by a test double that emulates the collective build
                     with the C oracle, so that the host-side logic (shard ranges, id broadcast, assembly
                     of the rank-local exports) runs at world_size > 1

Rank 0 gathers every rank's exports, stitches them with amira_b200.sharded.assemble_arrays and compares the
result, field by field and bit for bit, with the C oracle's build of the whole read set; then the same
after remove_low_coverage_components(5) and filter_graph(3, 1)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class GlooEmulatedDevice:
    """stands in for DeviceGraph on a CPU box: all-gathers the shards over gloo, builds the global graph with
    the oracle and keeps this rank's slice of the per-read outputs (what the library leaves on each GPU)"""

    def __init__(self):
        self.uid = None

    def nccl_unique_id(self):
        return np.arange(128, dtype=np.uint8)

    def comm_init(self, uid, rank, world):
        assert np.array_equal(np.asarray(uid, np.uint8), np.arange(128, dtype=np.uint8)), "id did not travel intact"
        self.rank, self.world = rank, world

    def build(self, ids, off, k, ps=None, pe=None):
        import torch.distributed as dist
        from oracle import c_oracle
        shards = [None] * self.world
        dist.all_gather_object(shards, (ids, off, ps, pe))
        all_ids = np.concatenate([s[0] for s in shards])
        offs, shift = [np.zeros(1, np.int64)], 0
        for s in shards:
            offs.append(s[1][1:] + shift)
            shift += int(s[1][-1])
        self.read_lo = sum(len(s[1]) - 1 for s in shards[:self.rank])
        self.read_hi = self.read_lo + len(off) - 1
        gps = None if ps is None else np.concatenate([s[2] for s in shards])
        gpe = None if pe is None else np.concatenate([s[3] for s in shards])
        self.g = c_oracle.COracleGraph(all_ids, np.concatenate(offs), k, gps, gpe)

    def remove_low_coverage_components(self, c):
        self.g.remove_low_coverage_components(c)

    def filter_graph(self, a, b):
        self.g.filter_graph(a, b)

    def arrays(self):
        a = self.g.arrays()
        lo, hi = self.read_lo, self.read_hi
        w0, w1 = int(a["win_off"][lo]), int(a["win_off"][hi])
        out = dict(a)
        out["win_off"] = a["win_off"][lo:hi + 1] - a["win_off"][lo]
        for f in ("win_node", "win_dir", "win_start", "win_end"):
            out[f] = a[f][w0:w1]
        for f in ("is_short", "to_correct"):
            out[f] = a[f][lo:hi]
        node_of = np.repeat(np.arange(len(a["node_cov"])), np.diff(a["node_reads_off"]))
        mine = (a["node_reads"] >= lo) & (a["node_reads"] < hi)
        out["node_reads"] = a["node_reads"][mine]
        roff = np.zeros(len(a["node_cov"]) + 1, np.int64)
        np.cumsum(np.bincount(node_of[mine], minlength=len(a["node_cov"])), out=roff[1:])
        out["node_reads_off"] = roff
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--config", default="c3")
    ap.add_argument("--reads", type=int, default=30000)
    ap.add_argument("--k", type=int, default=3)
    ap.add_argument("--positions", action="store_true")
    ap.add_argument("--uneven", action="store_true", help="give rank 0 an empty shard and the last rank the remainder")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from amira_b200 import sharded, synth
    from oracle import c_oracle
    from oracle import gmg_oracle as O

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.backend == "nccl":
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from amira_b200.device_graph import DeviceGraph
        dg = DeviceGraph(local_rank)
    else:
        dist.init_process_group("gloo")
        dg = GlooEmulatedDevice()
    gloo = dist.new_group(backend="gloo")      # object gathers of the exports go over gloo in both modes

    ids, off = synth.generate(synth.CONFIGS[args.config], 0, args.reads)
    ps = pe = None
    if args.positions:
        ps, pe = synth.positions_for(off, 3)
    if args.uneven:
        R = len(off) - 1
        cuts = [0, 0] + [R * i // (world - 1) for i in range(1, world)] if world > 1 else [0, R]
        lo, hi = cuts[rank], cuts[rank + 1]
        a, b = int(off[lo]), int(off[hi])
        my = (np.ascontiguousarray(ids[a:b]), (off[lo:hi + 1] - off[lo]).astype(np.int64),
              None if ps is None else np.ascontiguousarray(ps[a:b]), None if pe is None else np.ascontiguousarray(pe[a:b]))
    else:
        my = sharded.shard_csr(ids, off, rank, world, ps, pe)
    r, w = sharded.init_comm(dg, None)
    assert (r, w) == (rank, world)

    ref = c_oracle.COracleGraph(ids, off, args.k, ps, pe) if rank == 0 else None
    failures = []

    def check(stage):
        pieces = [None] * world
        dist.all_gather_object(pieces, dg.arrays(), group=gloo)
        if rank == 0:
            got = sharded.assemble_arrays(pieces)
            d = O.diff_arrays(got, ref.arrays())
            if d:
                failures.append((stage, d))

    for rep in range(2):                       # the second build reuses the handle's tables and buffers
        dg.build(my[0], my[1], args.k, my[2], my[3])
        check("build%d" % rep)
    if args.k > 1:
        dg.remove_low_coverage_components(5)
        if rank == 0:
            ref.remove_low_coverage_components(5)
        check("rlcc5")
        dg.filter_graph(3, 1)
        if rank == 0:
            ref.filter_graph(3, 1)
        check("filter3_1")
    ok = torch.tensor([0 if failures else 1])
    dist.broadcast(ok, 0, group=gloo)
    if rank == 0:
        print("SHARDED_PARITY", "OK" if not failures else "FAIL %r" % (failures,), flush=True)
    dist.barrier(group=gloo)
    dist.destroy_process_group()
    return 0 if int(ok.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
