"""vl_step: one generation over several compartments of one process under both transmission rules of the reference —
the parallel build's (src/runner.rs:401-416, sizes driven by the origin) and the serial build's (:550-631, sizes driven
by the target, stochastic rounding of migrant counts) — against the oracle's restatement of the same loops, in
distribution over seeded replicates (population size of every compartment after every generation, distinct haplotypes
and mean mutations at the end)."""
import numpy as np
import pytest
from scipy import stats

import helpers
from virolution_b200.simulation import ReferencePanic, step

pytestmark = pytest.mark.gpu

T_SYM = np.array([[0.8, 0.1, 0.1], [0.1, 0.8, 0.1], [0.1, 0.1, 0.8]])
T_ASYM = np.array([[0.7, 0.3, 0.0], [0.2, 0.6, 0.1], [0.0, 0.25, 0.9]])


def _case():
    c = helpers.case("singlehost")
    c["parameters"].dilution = 0.01  # target = sum(offspring) * dilution stays below max_population: sizes are random
    return c


@pytest.mark.parametrize("rule", [0, 1])
@pytest.mark.parametrize("T", [T_SYM, T_ASYM], ids=["symmetric", "asymmetric"])
def test_step_sizes_agree_in_distribution_with_oracle(rule, T):
    R, G, C = 60, 3, 3
    c = _case()
    rows_o, rows_g = [], []
    for r in range(R):
        cc = dict(c, seed=300 + r)
        w = helpers.make_oracle(cc, n_compartments=C)
        row = []
        for _ in range(G):
            w.step(T, rule)
            row += [w.population_len(k) for k in range(C)]
        off, pos, base = w.export_mutations(0)
        sets = helpers.mutation_sets(off, pos, base)
        row += [len(set(sets)), float(np.mean([len(s) for s in sets])) if sets else 0.0]
        rows_o.append(row)
        sims = [helpers.make_gpu(cc, seed=50000 + r, compartment_id=k) for k in range(C)]
        row = []
        for _ in range(G):
            step(sims, T, rule)
            row += [s.population_len() for s in sims]
        off, pos, base = sims[0].export_mutations()
        sets = helpers.mutation_sets(off, pos, base)
        row += [len(set(sets)), float(np.mean([len(s) for s in sets])) if sets else 0.0]
        rows_g.append(row)
        for s in sims:
            s.close()
    ro, rg = np.array(rows_o, dtype=float), np.array(rows_g, dtype=float)
    alpha = 0.01 / ro.shape[1]
    for k in range(ro.shape[1]):
        if np.all(ro[:, k] == ro[0, k]) and np.all(rg[:, k] == ro[0, k]):
            continue
        p = stats.ks_2samp(ro[:, k], rg[:, k]).pvalue
        assert p > alpha, f"column {k}: KS p={p:.2e} (oracle mean {ro[:, k].mean():.5g}, gpu mean {rg[:, k].mean():.5g})"


def test_parallel_rule_sizes_are_the_floor_of_factor_times_origin_target():
    # src/simulation.rs:516-527: size = (factor * target_size) as usize per (target, origin) pair
    c = helpers.case("epistasis_dense")
    c["parameters"].dilution = 1.0   # sum(offspring) * dilution exceeds max_population: target_size = max_population every generation
    n = c["parameters"].max_population
    sims = [helpers.make_gpu(c, compartment_id=k) for k in range(3)]
    for _ in range(3):
        step(sims, T_ASYM, 0)
        for t, s in enumerate(sims):
            assert s.population_len() == sum(int(T_ASYM[t, o] * n) for o in range(3))
    for s in sims:
        s.close()


def test_serial_rule_panics_like_the_reference_when_migrants_exceed_the_target():
    # usize underflow in amount[t][t] (src/runner.rs:583-590)
    c = helpers.case("epistasis_dense")
    sims = [helpers.make_gpu(c, compartment_id=k) for k in range(2)]
    with pytest.raises(ReferencePanic):
        step(sims, np.array([[0.1, 1.5], [1.5, 0.1]]), 1)
    for s in sims:
        s.close()
