"""Warp instructions executed per CUDA source line of one kernel in an .ncu-rep: python profiles/inst.py <rep> <kernel_regex> [top]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
data, tot, ci, hdr, fname = [], 0, None, None, ""
for r in rows:
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
    elif "# Samples" in r:
        hdr = r
        ci = {n: i for i, n in enumerate(hdr)}
    elif ci and len(r) >= len(hdr) and r[0] not in ("", "Line No"):
        try:
            n = int(r[ci["Instructions Executed"]])
        except ValueError:
            continue
        tot += n
        data.append((n, int(r[ci["# Samples"]] or 0), fname, r[0], r[1].strip()[:100]))
data.sort(key=lambda x: -x[0])
print("total warp instructions", tot)
for n, s, f, ln, src in data[:top]:
    print(f"{100*n/tot:5.1f}% {n:>10d} samples={s:>5d} {f}:{ln:>4s} {src}")
