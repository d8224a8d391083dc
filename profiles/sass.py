"""SASS of one kernel with executed-instruction counts: python profiles/sass.py <rep> <kernel_regex> [min_count]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
minc = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{kern}", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ci = {n: i for i, n in enumerate(hdr)}
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    n = int(r[ci["Instructions Executed"]] or 0)
    tot += n
    if n >= minc:
        print(f"{n:>10d} {int(r[ci['# Samples']] or 0):>5d} {r[ci['Source']].strip()}")
print("total", tot)
