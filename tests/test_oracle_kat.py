"""Known-answer tests the reference holds for this path, restated against the CPU oracle.

Each test names the reference test it restates.  These pin the oracle; the same vectors go through the CUDA path in
tests/test_gpu_kat.py.
"""
import numpy as np
import pytest

from helpers import UNIFORM_SUBST
from oracle import oracle as orc
from virolution_b200.abi import FitnessProvider, HostSpec, Parameters, UtilityFunction
from virolution_b200.config import exponential_table

A, T, C, G = 0, 1, 2, 3


def world(wt, table=None, H=2, epi=None, utility=None, n=0):
    par = Parameters(1e-3, 0.0, H, 100.0, 100, 1.0, UNIFORM_SUBST)
    prov = None
    if table is not None:
        prov = FitnessProvider("SimpleEpistatic" if epi is not None else "NonEpistatic", table=np.asarray(table, dtype=np.float64),
                               epistasis=epi, utility=utility or UtilityFunction("Linear"))
    w = orc.OracleWorld(np.asarray(wt, dtype=np.uint8), par, [HostSpec(0, H, prov, None)])
    if n:
        w.set_population_wildtype(n)
    return w


def test_table_get_value_indexing():
    # src/core/fitness/table.rs:127-142 get_value_indexing: table[pos*4 + idx], A,T,C,G = 0..3
    w = world([A, A], table=[[1., 2., 3., 4.], [5., 6., 7., 8.]])
    expected = {(0, A): 1., (0, T): 2., (0, C): 3., (0, G): 4., (1, A): 5., (1, T): 6., (1, C): 7., (1, G): 8.}
    for (pos, base), v in expected.items():
        assert w.table_get_value(pos, base) == v


def test_host_map_counting_sort():
    # src/core/hosts.rs:294-331 test_host_map: slices come out in DESCENDING infectant order
    w = world([A] * 4, H=2, n=10)
    w.set_infections([0, 1, 0, -1, 1, -1, 0, 1, 0, 1])
    offsets, hosts = w.get_host_map()
    assert hosts[offsets[0]:offsets[1]].tolist() == [8, 6, 2, 0]
    assert hosts[offsets[1]:offsets[2]].tolist() == [9, 7, 4, 1]
    assert w.get_infections().tolist() == [0, 1, 0, -1, 1, -1, 0, 1, 0, 1]  # find_host


def test_initiate_wildtype_and_create_descendant():
    # src/core/haplotype.rs:947-967 initiate_wildtype, create_descendant
    w = world([A, T, C, G])
    wt = w.wildtype()
    assert [w.hap_get_base(wt, i) for i in range(4)] == [A, T, C, G]
    ht = w.create_descendant(wt, [0], [G])
    assert [w.hap_get_base(ht, i) for i in range(4)] == [G, T, C, G]


def test_create_wide_genealogy():
    # src/core/haplotype.rs:969-991
    w = world([A, T, C, G])
    hts = [w.create_descendant(0, [0], [(i % 3) + 1]) for i in range(100)]
    for i, h in enumerate(hts):
        assert w.hap_get_base(h, 0) == (i % 3) + 1


def test_single_recombination():
    # src/core/haplotype.rs:993-1020: left chain all C, right chain all G, window [10, 90)
    w = world([T] * 100)
    left = 0
    for i in range(100):
        left = w.create_descendant(left, [i], [C])
    right = w.create_descendant(0, [0], [G])
    for i in range(1, 100):
        right = w.create_descendant(right, [i], [G])
    rec = w.create_recombinant(left, right, 10, 90)
    assert w.hap_get_sequence(rec).tolist() == [G] * 10 + [C] * 80 + [G] * 10


def test_recombinant_get_sequence():
    # src/core/haplotype.rs:1035-1048 get_sequence: wildtype x mutant, window [25, 75) -> mutation at 0 comes from the right
    w = world([A] * 100)
    hap = w.create_descendant(0, [0], [T])
    rec = w.create_recombinant(0, hap, 25, 75)
    assert w.hap_get_sequence(rec).tolist() == [T] + [A] * 99


def test_library_doc_example():
    # src/lib.rs:122-141
    w = world([A] * 4)
    m1 = w.create_descendant(0, [2], [T])
    m2 = w.create_descendant(0, [1, 2], [G, C])
    rec = w.create_recombinant(m1, m2, 0, 2)
    assert w.hap_get_sequence(0).tolist() == [A, A, A, A]
    assert w.hap_get_sequence(m1).tolist() == [A, A, T, A]
    assert w.hap_get_sequence(m2).tolist() == [A, G, C, A]
    assert w.hap_get_sequence(rec).tolist() == [A, A, C, A]


def test_merge_nodes_net_mutation_set():
    # src/core/haplotype.rs:1073-1117 merge_nodes: wt(A) -> T -> C -> G at site 0 leaves {0: G}
    w = world([A, T, C, G])
    h1 = w.create_descendant(0, [0], [T])
    h2 = w.create_descendant(h1, [0], [C])
    h3 = w.create_descendant(h2, [0], [G])
    assert w.hap_get_mutations(h3) == {0: G}


def test_back_mutation_removed_from_mutation_set():
    # src/core/haplotype.rs:697-698: a change back to the wildtype base removes the key
    w = world([A, T, C, G])
    h1 = w.create_descendant(0, [1], [G])
    h2 = w.create_descendant(h1, [1], [T])
    assert w.hap_get_mutations(h2) == {}
    assert w.hap_get_base(h2, 1) == T


def test_same_base_assertion():
    # src/core/haplotype.rs:250 assert_ne!(from, to)
    w = world([A, T, C, G])
    with pytest.raises(orc.ReferencePanic):
        w.create_descendant(0, [0], [A])


def test_exponential_table_wildtype_entries():
    # src/init/fitness.rs:241-270: wildtype entries are exactly 1, the others are not (neutral weight 0)
    wt = np.zeros(1000, dtype=np.uint8)
    tab = exponential_table(wt, weights=(0.29, 0.51, 0.2, 0.0), rng=np.random.default_rng(3))
    assert tab.shape == (1000, 4)
    assert np.all(tab[:, 0] == 1.0)
    assert np.all(tab[:, 1:] != 1.0)
    assert np.all(tab >= 0.0)


def test_neutral_table_values():
    # src/core/fitness/table.rs:111-125 get_value: neutral model is 1 everywhere
    par = Parameters(1e-3, 0.0, 2, 100.0, 100, 1.0, UNIFORM_SUBST)
    w = orc.OracleWorld(np.zeros(100, dtype=np.uint8), par, [HostSpec(0, 2, FitnessProvider("Neutral"), None)])
    h = w.create_descendant(0, [3, 7], [G, C])
    assert w.hap_raw_fitness(h) == 1.0 and w.hap_fitness(h) == 1.0


def test_update_fitness_fold_order():
    # src/core/fitness/table.rs:65-68: acc = acc * T[pos,to] / T[pos,from], left to right
    tab = np.array([[1.0, 0.3, 0.7, 1.1], [0.9, 1.0, 1.3, 0.2], [1.7, 0.6, 1.0, 0.4]])
    w = world([A, T, C], table=tab)
    h = w.create_descendant(0, [0, 1, 2], [T, C, G])
    acc = 1.0
    for pos, frm, to in ((0, A, T), (1, T, C), (2, C, G)):
        acc = acc * tab[pos, to] / tab[pos, frm]
    assert w.hap_raw_fitness(h) == 1.0 * acc
    h2 = w.create_descendant(h, [0], [G])
    assert w.hap_raw_fitness(h2) == (1.0 * acc) * (1.0 * tab[0, G] / tab[0, T])


def test_utility_functions():
    # src/core/fitness/utility.rs:48-60
    from virolution_b200.abi import vl_utility
    import ctypes as C_
    lib = orc.lib()
    for f in (0.0, 0.3, 1.0, 1.7, 25.0):
        assert lib.orc_utility_apply(C_.byref(UtilityFunction("Linear").to_c()), f) == f
        upper = 1.5
        factor = 1. / (upper - 1.)
        assert lib.orc_utility_apply(C_.byref(UtilityFunction("Algebraic", upper=upper).to_c()), f) == upper * factor * f / (1. + factor * f)
        p, n = 0.25, 3.0
        k = (1. - p) / p
        got = lib.orc_utility_apply(C_.byref(UtilityFunction("Hill", p=p, n=n).to_c()), f)
        assert got == pytest.approx(f ** n / (k + f ** n), rel=1e-15)


def test_epistasis_update_and_compute_quirks():
    # src/core/fitness/epistasis.rs:105-193 on the fixture's pair (100,T)-(102,C)=0.5:
    #  update: multiply when the partner IS present in the mutation set; compute: multiply when the partner is ABSENT
    L = 128
    wt = np.zeros(L, dtype=np.uint8)
    tab = np.ones((L, 4))
    epi = np.array([[100, T, 102, C, 0.5], [20, G, 50, T, 1.1]], dtype=np.float64)
    w = world(wt, table=tab, epi=epi)
    a = w.create_descendant(0, [100], [T])          # partner absent -> update factor 1
    assert w.hap_raw_fitness(a) == 1.0
    b = w.create_descendant(a, [102], [C])          # partner (100,T) present, position below -> 0.5
    assert w.hap_raw_fitness(b) == 0.5
    c = w.create_descendant(0, [100, 102], [T, C])  # both in one node: entries at positions <= the change that belong to the
    assert w.hap_raw_fitness(c) == 0.5              # node itself are skipped, so the pair is counted once
    rec = w.create_recombinant(a, b, 0, 10)         # recombinant -> compute_fitness: (100,T) has partner (102,C) present
    assert w.hap_get_mutations(rec) == {100: T, 102: C}
    assert w.hap_raw_fitness(rec) == 1.0            # ... so the "absent partner" rule does not fire
    rec2 = w.create_recombinant(a, 0, 0, 128)       # only (100,T): partner absent -> multiplied by 0.5
    assert w.hap_get_mutations(rec2) == {100: T}
    assert w.hap_raw_fitness(rec2) == 0.5


def test_samplers_match_exact_distributions():
    # restated rand_distr 0.5.1 Poisson / Binomial and WeightedAliasIndex against scipy's exact pmfs
    from scipy import stats
    n = 200000
    for lam in (0.3, 6.0, 11.9, 12.0, 100.0):
        x = orc.test_poisson(5, lam, n)
        hi = int(stats.poisson.ppf(1 - 1e-6, lam)) + 1
        obs = np.bincount(np.minimum(x.astype(np.int64), hi), minlength=hi + 1)
        exp = stats.poisson.pmf(np.arange(hi + 1), lam) * n
        exp[hi] = stats.poisson.sf(hi - 1, lam) * n
        keep = exp > 5
        chi = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
        assert stats.chi2.sf(chi, keep.sum() - 1) > 1e-4, lam
    for nn, p in ((5386, 1e-6), (9700, 2.16e-5), (30000, 1e-5), (200, 2e-3), (5000, 0.01)):
        x = orc.test_binomial(7, nn, p, n)
        hi = int(stats.binom.ppf(1 - 1e-6, nn, p)) + 1
        obs = np.bincount(np.minimum(x.astype(np.int64), hi), minlength=hi + 1)
        exp = stats.binom.pmf(np.arange(hi + 1), nn, p) * n
        exp[hi] = stats.binom.sf(hi - 1, nn, p) * n
        keep = exp > 5
        chi = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
        assert stats.chi2.sf(chi, max(1, keep.sum() - 1)) > 1e-4, (nn, p)
    wts = np.array([0, 5, 1, 0, 14, 3, 0, 7], dtype=np.uint64)
    x = orc.test_alias(9, wts, n)
    obs = np.bincount(x.astype(np.int64), minlength=wts.size)
    assert np.all(obs[wts == 0] == 0)
    exp = wts / wts.sum() * n
    chi = ((obs[wts > 0] - exp[wts > 0]) ** 2 / exp[wts > 0]).sum()
    assert stats.chi2.sf(chi, (wts > 0).sum() - 1) > 1e-4
    with pytest.raises(orc.ReferencePanic):
        orc.test_alias(1, np.zeros(4, dtype=np.uint64), 10)
