"""instruction / stall-sample distribution over the SASS of one `ncu --set full` capture:
python scripts/sass_regions.py gpurun_out/prof_x.ncu-rep [step]"""
import csv, re, subprocess, sys
rep = sys.argv[1]
step = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iN, iE = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
R = rows[2:]
acc = samp = 0; start = 0; ops = []
for i, r in enumerate(R):
    e = int(r[iE]) / 1e6; s = int(r[iN]); acc += e; samp += s
    m = re.search(r'\b(LDG[\.\w]*|STG[\.\w]*|ATOMG[\.\w]*|REDG[\.\w]*|ATOMS[\.\w]*|BAR[\.\w]*|LDS[\.\w]*|STS[\.\w]*|SHFL[\.\w]*)', r[iS])
    if m and e > 0.005: ops.append("%d:%s(%.2fM,%d)" % (i, m.group(1), e, s))
    if (i + 1) % step == 0 or i == len(R) - 1:
        if acc > 0.05: print(f'{start:5d}-{i:5d} inst={acc:8.1f}M samples={samp:6d}  ' + ' '.join(ops)[:400])
        acc = samp = 0; start = i + 1; ops = []
