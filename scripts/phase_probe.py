"""developer probe: per-phase device times of the build on the synthetic configs (not the bench)"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from amira_b200 import synth
from amira_b200.device_graph import DeviceGraph

def run(name, n_reads, k=None, reps=5):
    cfg = synth.CONFIGS[name]
    k = k or cfg.k
    ids, off = synth.generate(cfg, 0, n_reads)
    W = synth.count_windows(off, k)
    d_ids = torch.from_numpy(ids).cuda(); d_off = torch.from_numpy(off).cuda()
    g = DeviceGraph(0, profiling=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = None
    for r in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        g.build(d_ids, d_off, k, on_device=True); g.sync()
        dt = (time.perf_counter() - t) * 1e3
        ph = g.phase_ms()
        if best is None or dt < best[0]:
            best = (dt, ph)
    s = g.sizes()
    dt, ph = best
    print(f"{name} reads={n_reads} k={k} G={len(ids)} W={W} nodes={s['nodes']} edges={s['edges']} wall={dt:.3f} ms "
          f"-> {W/dt/1e6:.2f} G gene-mers/s; launches/build={g.kernel_launches()//reps}")
    print("   " + " ".join(f"{k_}={v:.3f}" for k_, v in ph.items() if v > 0))
    t = time.perf_counter(); g.remove_low_coverage_components(5); g.filter_graph(3, 1); g.sync()
    print(f"   rlcc+filter wall {(time.perf_counter()-t)*1e3:.3f} ms -> {g.sizes()}")
    g.close()

if __name__ == "__main__":
    g = DeviceGraph(0)
    for mb in (1, 16, 64, 128, 192, 512):
        print("random-access peak, table %d MB: red/s %.3g cas/s %.3g 32B-load/s %.3g" % ((mb,) + g.atomic_peak(mb << 20, 1 << 26)))
    g.close()
    run("c2", 50000)
    run("c3", 500000)
    run("c3", 500000, 7)
    run("c4", 2000000)
    run("c5", 1250000)
    run("c5", 10000000, reps=3)
