"""developer probe: wall time per device-resident build (profiling off, CUDA-graph replay as in the bench)
usage: python scripts/wall_probe.py c5 1250000 [k] [reps]"""
import sys, time
import torch
sys.path.insert(0, ".")
from amira_b200 import synth
from amira_b200.device_graph import DeviceGraph
name, n = sys.argv[1], int(sys.argv[2])
cfg = synth.CONFIGS[name]
k = int(sys.argv[3]) if len(sys.argv) > 3 else cfg.k
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
ids, off = synth.generate(cfg, 0, n)
d_ids, d_off = torch.from_numpy(ids).cuda(), torch.from_numpy(off).cuda()
g = DeviceGraph(0)
for _ in range(4):
    g.build(d_ids, d_off, k, on_device=True)
g.sync()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(reps):
    g.build(d_ids, d_off, k, on_device=True)
g.sync()
torch.cuda.synchronize()
print("%s %d k=%d: %.3f ms per build" % (name, n, k, (time.perf_counter() - t) * 1e3 / reps))
