import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from virolution_b200.simulation import GpuSimulation
from virolution_b200.workloads import workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
wl = workload("k2_hiv", n)
sim = GpuSimulation(wl["wildtype"], wl["parameters"], wl["specs"], seed=2026)
sim.set_population_wildtype(n)
for w in range(int(sys.argv[2]) if len(sys.argv) > 2 else 8):
    c0 = sim.get_counters()
    sim.timer_start(); sim.run(20); ms = sim.timer_stop()
    c1 = sim.get_counters()
    print(f"gens {20*w:4d}-{20*w+20:4d}: {ms/20:.4f} ms/gen  records {c1['haplotype_records']} (+{(c1['haplotype_records']-c0['haplotype_records'])/20:.0f}/gen) arena {c1['arena_words']} (+{(c1['arena_words']-c0['arena_words'])/20:.0f}/gen) gc {c1['gc_runs']}")
sim.profile(True); sim.run(10); sim.synchronize(); rep = sim.profile_report(); sim.profile(False)
print({k: round(v[1]/v[0], 4) for k, v in rep.items() if v[1]/v[0] > 0.01}, "sum", round(sum(v[1] for v in rep.values())/10, 4))
sim.close()
