"""print the hottest lines of an ncu source-page csv (ncu -i X.ncu-rep --page source [--print-source cuda] --csv)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body.sort(key=lambda r: -int(r[ci["# Samples"]] or 0))
print("total samples", tot)
for r in body[:n]:
    s = int(r[ci["# Samples"]] or 0)
    top = sorted(((int(r[ci[h]] or 0), h) for h in stalls), reverse=True)[:3]
    src = r[ci["Source"]].strip()[:110]
    extra = ""
    if "Instructions Executed" in ci:
        extra = " inst=%s" % r[ci["Instructions Executed"]]
    print("%5.1f%% %-110s %s%s" % (100.0 * s / max(tot, 1), src, " ".join("%s=%d" % (h[6:], v) for v, h in top if v), extra))
