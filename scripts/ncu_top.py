"""hottest source lines of an ncu capture:  python scripts/ncu_top.py <report.ncu-rep> [n]
(uses `ncu -i X --page source --csv --print-source cuda,sass`; needs -lineinfo at compile time)"""
import csv, io, subprocess, sys
rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname, hdr, out = "", None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].rsplit("/", 1)[-1]
    elif len(r) > 5 and r[0] == "Line No":
        hdr = r
        ci = {}
        for i, h in enumerate(hdr):
            ci.setdefault(h, i)
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        out.append((fname, r))
tot = sum(int(r[ci["# Samples"]] or 0) for _, r in out)
tin = sum(int(r[ci["Instructions Executed"]] or 0) for _, r in out)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
out.sort(key=lambda fr: -int(fr[1][ci["# Samples"]] or 0))
print("total samples", tot, "warp instructions", tin)
for f, r in out[:n]:
    s = int(r[ci["# Samples"]] or 0)
    top = sorted(((int(r[ci[h]] or 0), h) for h in stalls), reverse=True)[:3]
    print("%5.1f%% inst %4.1f%% %-18s:%-4s %-90s %s" % (
        100.0 * s / max(tot, 1), 100.0 * int(r[ci["Instructions Executed"]] or 0) / max(tin, 1), f, r[0],
        r[1].strip()[:90], " ".join("%s=%d" % (h[6:], v) for v, h in top if v)))
