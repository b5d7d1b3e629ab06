"""ncu target: a few builds of one synthetic workload (first build sizes the tables cold, the rest reuse
the previous build's counts).  usage: python scripts/ncu_target.py c5 1250000 [k] [reps] [--no-filter]"""
import sys
import torch
sys.path.insert(0, ".")
from amira_b200 import synth
from amira_b200.device_graph import DeviceGraph

args = [a for a in sys.argv[1:] if not a.startswith("--")]
name, n = args[0], int(args[1])
cfg = synth.CONFIGS[name]
k = int(args[2]) if len(args) > 2 else cfg.k
reps = int(args[3]) if len(args) > 3 else 3
ids, off = synth.generate(cfg, 0, n)
d_ids, d_off = torch.from_numpy(ids).cuda(), torch.from_numpy(off).cuda()
g = DeviceGraph(0, profiling=True)
for _ in range(reps):
    g.build(d_ids, d_off, k, on_device=True)
    g.sync()
    print({k_: round(v, 3) for k_, v in g.phase_ms().items() if v > 0}, g.sizes()["nodes"])
if "--no-filter" not in sys.argv:
    g.remove_low_coverage_components(5)
    g.filter_graph(3, 1)
    g.sync()
    print(g.phase_ms()["filter"], g.sizes())
