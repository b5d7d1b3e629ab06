"""per-kernel device times of the last complete build in an ncu launch list: python scripts/launch_list.py gpurun_out/launches_x.csv"""
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
iK, iV = hdr.index('Kernel Name'), hdr.index('Metric Value')
L = [(r[iK], float(r[iV].replace(',', ''))) for r in rows[1:]]
starts = [i for i, (k, v) in enumerate(L) if 'k_read_windows' in k]
seg = L[starts[-1]:]
tot = sum(v for k, v in seg) / 1000
for k, v in seg:
    name = re.sub(r'\(.*', '', k).replace('amira::', '').replace('void ', '')
    print(f"{name[:58]:58s} {v/1000:9.1f} us {100*v/1000/tot:5.1f}%")
print('total %.1f us, %d launches' % (tot, len(seg)))
