"""Multi-part transmission on ONE GPU: what an origin of an 8-compartment run does per generation (advance to the offspring
map, draw 8 parts in one plan/resample pass, install them) with per-kernel CUDA-event times.
    python profiles/parts8.py [n] [parts]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from virolution_b200.simulation import GpuSimulation  # noqa: E402
from virolution_b200.workloads import workload  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 8
wl = workload("k2_hiv", n)
sim = GpuSimulation(wl["wildtype"], wl["parameters"], wl["specs"], seed=2026)
sim.set_population_wildtype(n)
sim.run(20)
factors = [0.9] + [0.1 / (parts - 1)] * (parts - 1) if parts > 1 else [1.0]
for rep in range(2):
    if rep == 1:
        sim.profile(True)
    for _ in range(5):
        sim.advance_to_offspring()
        sim.set_population(sim.subsample_populations(factors))
    sim.synchronize()
rep = sim.profile_report()
for k, (cnt, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    if ms / 5 > 0.005:
        print(f"{k[:48]:48s} {cnt / 5:5.1f} x {ms / cnt * 1000:7.1f} us = {ms / 5 * 1000:7.1f} us/gen")
print("population", sim.population_len())
sim.close()
