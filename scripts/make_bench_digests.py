"""Write tests/golden/bench_digests.json: the C oracle's digests of the bench workloads (C5 shards of 1.25M reads
per GPU at 1, 2, 4 and 8 GPUs, k=5; C3 500k reads at k=3/5/7 through remove_low_coverage_components(5) +
filter_graph(3,1); C2; C4 2M reads k=3).  Run in the build container:  python scripts/make_bench_digests.py [names...]"""
import json
import os
import sys
import time
from dataclasses import replace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from amira_b200 import sharded, synth  # noqa: E402
from oracle import c_oracle, digests  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "bench_digests.json")
READS_PER_GPU = 1_250_000


def c5(n_gpus):
    cfg = replace(synth.CONFIGS["c5"], n_reads=READS_PER_GPU * n_gpus)
    ids, off = synth.generate(cfg, 0, cfg.n_reads)
    t = time.time()
    a = c_oracle.COracleGraph(ids, off, cfg.k).arrays()
    out = {"oracle_seconds": round(time.time() - t, 1), "k": cfg.k, "reads": cfg.n_reads, "global": digests.global_digest(a),
           "ranks": []}
    for r in range(n_gpus):
        lo, hi = r * READS_PER_GPU, (r + 1) * READS_PER_GPU
        out["ranks"].append(digests.rank_digest(digests.slice_rank(a, lo, hi)))
    return out


def staged(name, n_reads, k):
    cfg = synth.CONFIGS[name]
    ids, off = synth.generate(cfg, 0, n_reads)
    t = time.time()
    g = c_oracle.COracleGraph(ids, off, k)
    out = {"k": k, "reads": n_reads, "stages": {}}
    R = len(off) - 1

    def snap(stage):
        a = g.arrays()
        out["stages"][stage] = {"global": digests.global_digest(a), "rank": digests.rank_digest(digests.slice_rank(a, 0, R))}
    snap("build")
    g.remove_low_coverage_components(5)
    snap("rlcc5")
    g.filter_graph(3, 1)
    snap("rlcc5_filter3_1")
    out["oracle_seconds"] = round(time.time() - t, 1)
    return out


JOBS = {
    "c5_n1": lambda: c5(1), "c5_n2": lambda: c5(2), "c5_n4": lambda: c5(4), "c5_n8": lambda: c5(8),
    "c3_k3": lambda: staged("c3", 500_000, 3), "c3_k5": lambda: staged("c3", 500_000, 5),
    "c3_k7": lambda: staged("c3", 500_000, 7), "c2_k3": lambda: staged("c2", 50_000, 3),
    "c4_k3": lambda: staged("c4", 2_000_000, 3),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(JOBS)
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for n in names:
        t = time.time()
        res[n] = JOBS[n]()
        print(n, "done in %.0f s" % (time.time() - t), flush=True)
        with open(OUT, "w") as f:
            json.dump(res, f, indent=1, sort_keys=True)
