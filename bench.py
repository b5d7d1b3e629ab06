#!/usr/bin/env python
"""bench.py -- gene-mers/s of the GeneMerGraph build on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement of upstream's path

Workload (config.workload): the C5 metagenome-scale synthetic gene-call set of BASELINE.json
(configs[4]: 10M reads x 30 gene calls, 60k-gene vocabulary, 50 genomes, 1% bad calls, k=5), weak-scaled:
every GPU builds over its contiguous shard of 1.25M reads, so N=8 is the full 10M-read set and N=1 is
one shard (37.4M gene calls, 32.4M gene-mers; 150 MB of gene ids, larger than the 126 MB L2).

A step = one complete build of the graph of the resident reads: window enumeration,
canonicalisation, node/edge hash tables, first-seen ordering, per-read node lists, node->read
incidence, adjacency, connected components -- everything GeneMerGraph.__init__ computes
(upstream amira/construct_graph.py:31-102).

  value   gene-mers/s with the CSR input already resident in HBM (CUDA events on the handle's stream)
  e2e     the same through the C ABI with HOST buffers: pinned host CSR -> amira_gmg_build (H2D inside)
          -> every graph array exported back to pinned host memory (D2H inside)
  roofline / cpu_baseline / clocks: see DESIGN.md "Measurement"

One JSON line on stdout (rank 0).  Everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from dataclasses import replace

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READS_PER_GPU = 1_250_000
METRIC = "gene-mers/sec GeneMerGraph build"
UNIT = "gene-mers/s"
BYTES_PER_GENE_MER = 13          # 4 id in + 4 node idx + 1 direction + 4 node->read incidence out
ATOMICS_PER_GENE_MER = 3         # 1 node update + 2 directed-edge updates


_REAL_STDOUT = None


def emit(obj):
    """the one JSON line, on the process's original stdout"""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload(n_gpus: int, rank: int, reads_per_gpu: int):
    from amira_b200 import synth
    cfg = replace(synth.CONFIGS["c5"], n_reads=reads_per_gpu * n_gpus)
    ids, off = synth.generate(cfg, rank * reads_per_gpu, reads_per_gpu)
    return cfg, ids, off


def config_dict(cfg, n_gpus, reads_per_gpu, extra=None):
    d = {
        "workload": "C5 metagenome-scale synthetic gene calls (BASELINE.json configs[4]), weak-scaled: "
                    "%d reads x 30 calls per GPU, k=%d, 60k-gene vocabulary, 50 genomes, 1%% false/missing/"
                    "strand-flipped calls; 8 GPUs = the 10M-read set" % (reads_per_gpu, cfg.k),
        "reads_per_gpu": reads_per_gpu, "reads_total": reads_per_gpu * n_gpus, "k": cfg.k, "vocab": cfg.vocab,
        "sharding": "contiguous reads per rank; canonical gene-mers owned by hash range, NCCL all-to-all" if n_gpus > 1
                    else "single GPU",
        "l2": "inputs larger than L2 (150 MB of gene ids + ~1 GB of per-build arrays per GPU); no explicit flush",
    }
    if extra:
        d.update(extra)
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.rows, self.proc, self.dev = [], None, device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-i", str(self.dev), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception as e:          # no nvidia-smi: report it instead of failing the bench
            log("clock sampler unavailable:", e)
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def pinned(arr):
    import torch
    t = torch.empty(arr.shape, dtype=torch.from_numpy(arr[:0]).dtype, pin_memory=True)
    t.numpy()[...] = arr
    return t


def time_c_oracle(ids, off, k, n_reads):
    """CPU restatement (oracle/gmg_oracle.c, 1 thread) on the first n_reads reads -> (gene-mers/s, seconds, W)"""
    from amira_b200 import synth
    from oracle import c_oracle
    o = off[: n_reads + 1]
    i = ids[: int(o[-1])]
    W = synth.count_windows(o, k)
    t = time.perf_counter()
    g = c_oracle.COracleGraph(i, o, k)
    dt = time.perf_counter() - t
    del g
    return W / dt, dt, W


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: upstream's algorithm on the host CPU.

    Upstream is pure Python and /root/reference does not travel to the GPU box, so this arm times the
    plain-C restatement of its build (oracle/gmg_oracle.c, pinned to upstream's golden vectors); it
    is single-threaded because upstream's build is (every call site passes cores=1 and the result
    depends on the sequential dict insertion order)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import c_oracle
    c_oracle.build()
    cfg, ids, off = workload(args.gpus, 0, args.reads_per_gpu)
    k = cfg.k
    rate, _, _ = time_c_oracle(ids, off, k, min(20_000, args.reads_per_gpu))
    # bounded sample: the whole run (warmup + steps) stays within ~2 minutes of CPU time
    budget_s = 120.0 / max(1, args.steps + args.warmup)
    per_read = (cfg.fixed_len - k + 1) if cfg.fixed_len else cfg.mean_len
    n = int(min(args.reads_per_gpu, max(20_000, budget_s * rate / per_read)))
    for _ in range(args.warmup):
        time_c_oracle(ids, off, k, n)
    t0 = time.perf_counter()
    W = 0
    for _ in range(args.steps):
        _, _, w = time_c_oracle(ids, off, k, n)
        W += w
    dt = time.perf_counter() - t0
    value = W / dt
    sample = "first %d reads of rank 0's shard (%d gene-mers per step), plain-C restatement of upstream's build" % (
        n, W // max(1, args.steps))
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": config_dict(cfg, args.gpus, args.reads_per_gpu, {"sample_reads": n}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)
    return 0


# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from amira_b200 import _lib, synth
    from amira_b200.device_graph import DeviceGraph

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        log("WORLD_SIZE %d != --gpus %d; using WORLD_SIZE" % (world, args.gpus))
    n_gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg, ids, off = workload(n_gpus, rank, args.reads_per_gpu)
    k = cfg.k
    R, G = len(off) - 1, len(ids)
    W = synth.count_windows(off, k)
    stream = torch.cuda.Stream()
    dg = DeviceGraph(local_rank, stream=stream.cuda_stream, profiling=True)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.from_numpy(dg.nccl_unique_id().copy())
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        dg.comm_init(uid.cpu().numpy(), rank, world)

    h_ids, h_off = pinned(ids), pinned(off)
    with torch.cuda.stream(stream):
        d_ids = h_ids.to("cuda", non_blocking=True)
        d_off = h_off.to("cuda", non_blocking=True)
    stream.synchronize()

    # ---- device-resident arm ------------------------------------------------------------------
    def step_device():
        dg.build(d_ids, d_off, k, on_device=True, wait=False)     # enqueue only: no host synchronisation per build

    if args.table_load > 0:              # developer experiment: hash-table load factor
        step_device()
        s0 = dg.sizes()
        dg.reserve(int(s0["nodes"] / (2 * args.table_load)), int(s0["edges"] / 2 / (2 * args.table_load)))

    for _ in range(args.warmup):
        step_device()
    dg.sync()
    sizes = dg.sizes()
    launches0 = dg.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        smi_id = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        smi_id = str(local_rank)
    sampler = ClockSampler(smi_id)
    if rank == 0:
        sampler.start()
    barrier()
    phase_acc = {}
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
        if args.phases:                      # per-phase events are always recorded; reading them needs a sync
            dg.sync()
            for name, ms in dg.phase_ms().items():
                phase_acc[name] = phase_acc.get(name, 0.0) + ms
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = dg.kernel_launches() - launches0
    if not args.phases:
        # phase times of the last step only (events of one build); average over a few extra untimed steps below
        pass
    t = torch.tensor([ms_total], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    w_all = torch.tensor([W], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(w_all)
    W_all = int(w_all.item())
    value = W_all * args.steps / (ms_total * 1e-3)

    # dominant-kernel duration: CUDA events around k_insert_windows on the handle's stream, averaged over steps
    kern_ms = []
    for _ in range(min(args.steps, 10)):
        step_device()
        dg.sync()
        ph = dg.phase_ms()
        kern_ms.append(ph["insert_kernel"])
        for name, ms in ph.items():
            phase_acc[name] = phase_acc.get(name, 0.0) + ms
    n_ph = len(kern_ms) + (args.steps if args.phases else 0)
    phases = {n: round(v / n_ph, 4) for n, v in phase_acc.items() if v > 0}
    kernel_ms = sum(kern_ms) / len(kern_ms)

    # ---- end-to-end arm: host CSR in, host graph arrays out, through the C ABI ------------------
    # multi-GPU: the node / edge tables are identical on every rank, so rank 0 collects them; every rank
    # exports what it owns (its per-read lists and its share of the node -> read incidence)
    replicated = rank == 0
    dg.build(h_ids.numpy(), h_off.numpy(), k)
    out = dg.arrays(replicated=replicated)
    out_pinned = {n: pinned(a).numpy() for n, a in out.items() if isinstance(a, np.ndarray) and a.ndim >= 1 and
                  n not in ("win_start", "win_end")}
    d2h_bytes = int(sum(a.nbytes for a in out_pinned.values()))
    h2d_bytes = int(ids.nbytes + off.nbytes)
    if world > 1:                                 # bytes per step of the whole job
        tb = torch.tensor([h2d_bytes, d2h_bytes], device="cuda", dtype=torch.int64)
        dist.all_reduce(tb)
        h2d_bytes, d2h_bytes = int(tb[0].item()), int(tb[1].item())

    def step_e2e():
        dg.build(h_ids.numpy(), h_off.numpy(), k)
        dg.arrays(out=out_pinned, replicated=replicated)

    n_e2e = max(3, min(args.steps, 10))
    for _ in range(2):
        step_e2e()
    barrier()
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        step_e2e()
    e1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e0.elapsed_time(e1), wall_ms)      # exports end with a host sync: wall covers the D2H tail
    t = torch.tensor([e2e_ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = W_all * n_e2e / (e2e_ms * 1e-3)
    clocks = sampler.stop() if rank == 0 else None   # sampled from the start of the timed region to here

    # ---- roofline of the dominant kernel and of the whole build --------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            hbm_peak = float(json.load(f)["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        hbm_peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    bytes_alg = BYTES_PER_GENE_MER * W + 4 * R * (k - 1) + 8 * R
    achieved = bytes_alg / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "insert_kernel_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": round(achieved, 2), "peak": hbm_peak, "unit": "GB/s",
                "frac": round(achieved / hbm_peak, 4), "traffic": traffic, "kernel": "k_insert_windows",
                "kernel_ms": round(kernel_ms, 4), "bytes_per_gene_mer": BYTES_PER_GENE_MER,
                "algorithmic_bytes_per_launch": int(bytes_alg), "peak_source": peak_src,
                "whole_build_frac": round(bytes_alg / (ms_total / args.steps * 1e-3) / 1e9 / hbm_peak, 4)}
    if rank == 0 and not args.no_atomic_peak:
        # the second roofline of SURVEY.md 8(d): algorithmic atomics / measured random-address atomic rate
        red, cas, ld = dg.atomic_peak(64 << 20, 1 << 26)
        t_atomic_ms = ATOMICS_PER_GENE_MER * W / red * 1e3
        roofline["atomic"] = {"red_add_per_s": red, "cas_per_s": cas, "sector_load_per_s": ld,
                              "table": "64 MB (L2 resident)",
                              "atomics_per_gene_mer": ATOMICS_PER_GENE_MER,
                              "kernel_frac": round(t_atomic_ms / kernel_ms, 4),
                              "whole_build_frac": round(t_atomic_ms / (ms_total / args.steps), 4)}

    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": config_dict(cfg, n_gpus, args.reads_per_gpu),
        "graph": {"gene_mers": W_all, "nodes": sizes["nodes"], "edges": sizes["edges"]},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": e2e_ms / n_e2e, "steps": n_e2e,
                "what": "pinned host CSR -> amira_gmg_build -> amira_gmg_export_{nodes,edges,reads} into pinned host arrays"
                        + (" (every rank: its per-read lists and incidence share; rank 0: also the replicated node / edge "
                           "tables; bytes are the job's total)" if n_gpus > 1 else "")},
        "gpu_launches": int(launches), "phases_ms": phases, "roofline": roofline, "clocks": clocks,
        "library": _lib.load().amira_version().decode(),
    }

    # ---- CPU baseline beside it (rank 0, N=1 only) ----------------------------------------------
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle
        c_oracle.build()
        n = min(R, args.cpu_sample_reads)
        rate, dt, w = time_c_oracle(ids, off, k, n)
        result["cpu_baseline"] = {
            "value": rate, "unit": UNIT, "cores": 1, "kind": "port", "host_cores": os.cpu_count(), "seconds": round(dt, 2),
            "sample": "first %d of the %d reads of the same workload (%d gene-mers), plain-C restatement of upstream's "
                      "single-process build (oracle/gmg_oracle.c); upstream's own Python path measured 9-17k gene-mers/s "
                      "on one core (BASELINE.md)" % (n, R, w)}
    # ---- ms per graph on the single-isolate config (BASELINE.json configs[1]) -------------------
    if rank == 0 and n_gpus == 1 and not args.no_c2:
        c2 = synth.CONFIGS["c2"]
        i2, o2 = synth.generate(c2, 0, c2.n_reads)
        with torch.cuda.stream(stream):
            di, do = torch.from_numpy(i2).cuda(), torch.from_numpy(o2).cuda()
        stream.synchronize()
        for _ in range(5):
            dg.build(di, do, c2.k, on_device=True, wait=False)
        dg.sync()
        e0.record(stream)
        for _ in range(20):
            dg.build(di, do, c2.k, on_device=True, wait=False)
        e1.record(stream)
        torch.cuda.synchronize()
        w2 = synth.count_windows(o2, c2.k)
        ms2 = e0.elapsed_time(e1) / 20
        result["c2_isolate"] = {"workload": "BASELINE.json configs[1]: 50k reads x ~25 calls, 6k vocab, k=3",
                                "gene_mers": w2, "ms_per_graph": ms2, "gene_mers_per_s": w2 / (ms2 * 1e-3)}
    if rank == 0:
        emit(result)
    dg.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    # NCCL and friends print banners on stdout; the contract is ONE JSON line there, so everything else
    # (including native code writing to fd 1) is sent to stderr and the line goes to the saved descriptor
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads-per-gpu", type=int, default=READS_PER_GPU)
    ap.add_argument("--cpu-sample-reads", type=int, default=READS_PER_GPU)
    ap.add_argument("--phases", action="store_true", help="sync after every step to accumulate per-phase times")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-atomic-peak", action="store_true")
    ap.add_argument("--no-c2", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end arm (profiling runs)")
    ap.add_argument("--table-load", type=float, default=0.0, help="experiment: hash tables sized to this load factor")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
