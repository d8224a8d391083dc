"""Per-kernel summary of an .ncu-rep: python profiles/rawsum.py <rep> [out.csv]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__grid_size',
        'lts__t_sector_hit_rate.pct']
want = [w for w in want if w in idx]
print(' | '.join(f"{w.split('.')[0][-24:]}[{units[idx[w]]}]" for w in want))
for r in rows[2:]:
    print(' | '.join(r[idx[w]][:30] for w in want))
if len(sys.argv) > 2:
    with open(sys.argv[2], 'w') as f:
        w = csv.writer(f)
        w.writerow(want)
        w.writerow([units[idx[k]] for k in want])
        for r in rows[2:]:
            w.writerow([r[idx[k]] for k in want])
