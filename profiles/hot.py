"""Aggregate an ncu source page by CUDA source line: python profiles/hot.py <rep> <kernel_regex> [top]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
data, tot, fname, ci, hdr = [], 0, "", None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
    elif "# Samples" in r:
        hdr = r
        ci = {n: i for i, n in enumerate(hdr)}
        ci["Source"] = 1
    elif ci and len(r) >= len(hdr) and r[0] not in ("", "Line No"):
        try:
            s = int(r[ci["# Samples"]])
        except ValueError:
            continue
        tot += s
        data.append((s, fname, r))
data.sort(key=lambda x: -x[0])
print(rows[0][:2], "total samples", tot)
cols = [c for c in ("stall_long_sb", "stall_short_sb", "stall_mio", "stall_lg", "stall_math", "stall_barrier", "stall_wait", "stall_not_selected", "stall_branch_resolving", "stall_no_inst", "stall_dispatch") if c in ci]
for s, f, r in data[:top]:
    stalls = sorted(((int(r[ci[c]] or 0), c) for c in cols), reverse=True)[:3]
    print(f"{100*s/max(tot,1):5.1f}% {f}:{r[0]:>4s} {r[1].strip()[:95]:95s} inst={r[ci['Instructions Executed']]:>9s} " + " ".join(f"{c[6:]}={v}" for v, c in stalls if v))
