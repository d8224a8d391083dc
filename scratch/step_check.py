import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from scipy import stats
import helpers
from virolution_b200.simulation import step
T = np.array([[0.7, 0.3, 0.0], [0.2, 0.6, 0.1], [0.0, 0.25, 0.9]])
c = helpers.case("singlehost"); c["parameters"].dilution = 0.01
R, G, C = 300, 3, 3
ro, rg = [], []
for r in range(R):
    cc = dict(c, seed=300 + r)
    w = helpers.make_oracle(cc, n_compartments=C)
    row = []
    for _ in range(G):
        w.step(T, 0); row += [w.population_len(k) for k in range(C)]
    ro.append(row)
    sims = [helpers.make_gpu(cc, seed=int(sys.argv[1]) + r, compartment_id=k) for k in range(C)]
    row = []
    for _ in range(G):
        step(sims, T, 0); row += [s.population_len() for s in sims]
    rg.append(row)
    for s in sims: s.close()
ro, rg = np.array(ro, float), np.array(rg, float)
for k in range(ro.shape[1]):
    print(k, "oracle %.1f +- %.1f  gpu %.1f +- %.1f  KS p=%.3g" % (ro[:, k].mean(), ro[:, k].std(), rg[:, k].mean(), rg[:, k].std(), stats.ks_2samp(ro[:, k], rg[:, k]).pvalue))
