import sys, torch
sys.path.insert(0, ".")
from amira_b200 import synth
from amira_b200.device_graph import DeviceGraph
for name, n, k in (("c3", 500000, 3), ("c3", 500000, 7), ("c2", 50000, 3), ("c4", 2000000, 3)):
    cfg = synth.CONFIGS[name]
    ids, off = synth.generate(cfg, 0, n)
    d_ids, d_off = torch.from_numpy(ids).cuda(), torch.from_numpy(off).cuda()
    g = DeviceGraph(0, profiling=True)
    for _ in range(3):
        g.build(d_ids, d_off, k, on_device=True); g.sync()
    print(name, k, {k_: round(v, 3) for k_, v in g.phase_ms().items() if v > 0}, g.sizes()["nodes"], g.sizes()["windows"])
    g.close()
