"""Profiling driver: K2 (HIV-length, N=1e7) for a few device-resident generations.  Used under ncu:
    python profiles/run_k2.py [generations] [n] [flags]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from virolution_b200.simulation import GpuSimulation  # noqa: E402
from virolution_b200.workloads import workload  # noqa: E402

gens = int(sys.argv[1]) if len(sys.argv) > 1 else 12
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
wl = workload("k2_hiv", n)
sim = GpuSimulation(wl["wildtype"], wl["parameters"], wl["specs"], seed=2026, flags=flags)
sim.set_population_wildtype(n)
sim.run(gens)
sim.synchronize()
st = sim.debug_stamps().astype("int64")
print("plan kernel phases (us): prefix+size %.1f, tile poisson %.1f, fill draws %.1f, scan %.1f" % tuple((st[k + 1] - st[k]) / 1e3 for k in range(4)))
print("population", sim.population_len(), sim.get_counters())
sim.close()
