import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from virolution_b200.simulation import GpuSimulation
from virolution_b200.workloads import workload
n = 10_000_000
wl = workload("k2_hiv", n)
sim = GpuSimulation(wl["wildtype"], wl["parameters"], wl["specs"], seed=2026)
sim.set_population_wildtype(n)
sim.run(int(sys.argv[1])); sim.synchronize()
for rep in range(3):
    sim.timer_start(); sim.run(10); ms = sim.timer_stop()
    print("unprofiled", round(ms / 10, 4), sim.get_counters())
    sim.profile(True); sim.timer_start(); sim.run(10); ms = sim.timer_stop(); r = sim.profile_report(); sim.profile(False)
    print("profiled  ", round(ms / 10, 4), "kernel total/gen", round(sum(v[1] for v in r.values()) / 10, 4), {k: (v[0], round(v[1] / 10, 4)) for k, v in r.items() if v[1] / 10 > 0.02})
sim.close()
