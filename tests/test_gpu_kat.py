"""The reference's known-answer tests for this path, driven through the C ABI on the device
(same vectors as tests/test_oracle_kat.py)."""
import numpy as np
import pytest

import helpers
from helpers import UNIFORM_SUBST
from virolution_b200.abi import FitnessProvider, HostSpec, Parameters, UtilityFunction
from virolution_b200.simulation import GpuSimulation, ReferencePanic, VirolutionError

pytestmark = pytest.mark.gpu
A, T, C, G = 0, 1, 2, 3


def sim(wt, n, H=2, table=None, epi=None, rho=0.0):
    par = Parameters(1e-3, rho, H, 100.0, 1000, 1.0, UNIFORM_SUBST)
    prov = None
    if table is not None:
        prov = FitnessProvider("SimpleEpistatic" if epi is not None else "NonEpistatic", table=np.asarray(table, dtype=np.float64),
                               epistasis=epi, utility=UtilityFunction("Linear"))
    g = GpuSimulation(np.asarray(wt, dtype=np.uint8), par, [HostSpec(0, H, prov, None)])
    g.set_population_wildtype(n)
    return g


def mutate(g, changes):
    """changes: {infectant: [(site, base), ...]} applied as one create_descendant per infectant."""
    n = g.population_len()
    offsets = np.zeros(n + 1, dtype=np.uint64)
    sites, bases = [], []
    for i in range(n):
        for s, b in changes.get(i, []):
            sites.append(s)
            bases.append(b)
        offsets[i + 1] = len(sites)
    g.replay_infections(np.zeros(n, dtype=np.int64))
    g.infect()
    g.replay_mutations(offsets, np.array(sites, dtype=np.uint32), np.array(bases, dtype=np.uint8))
    g.mutate_infectants()


def sequence(g, i, L):
    return [g.get_base(i, p) for p in range(L)]


def mutations(g):
    off, pos, base = g.export_mutations()
    return [dict(m) for m in helpers.mutation_sets(off, pos, base)]


def test_host_map_counting_sort():
    # src/core/hosts.rs:294-331
    g = sim([A] * 4, 10)
    inf = np.array([0, 1, 0, -1, 1, -1, 0, 1, 0, 1], dtype=np.int64)
    g.replay_infections(inf)
    g.infect()
    offsets, hosts = g.get_host_map()
    assert hosts[offsets[0]:offsets[1]].tolist() == [8, 6, 2, 0]
    assert hosts[offsets[1]:offsets[2]].tolist() == [9, 7, 4, 1]
    assert g.get_infections().tolist() == inf.tolist()


def test_wildtype_and_descendant_bases():
    # src/core/haplotype.rs:947-967
    g = sim([A, T, C, G], 2)
    assert sequence(g, 0, 4) == [A, T, C, G]
    mutate(g, {1: [(0, G)]})
    assert sequence(g, 0, 4) == [A, T, C, G]
    assert sequence(g, 1, 4) == [G, T, C, G]


def test_wide_genealogy():
    # src/core/haplotype.rs:969-991
    g = sim([A, T, C, G], 100, H=1)
    mutate(g, {i: [(0, (i % 3) + 1)] for i in range(100)})
    for i in range(100):
        assert g.get_base(i, 0) == (i % 3) + 1


def test_single_recombination():
    # src/core/haplotype.rs:993-1020: left chain all C, right chain all G, window [10, 90)
    g = sim([T] * 100, 2, H=1)
    for i in range(100):
        mutate(g, {1: [(i, C)], 0: [(i, G)]})
    g.replay_infections(np.zeros(2, dtype=np.int64))
    g.infect()
    _, hosts = g.get_host_map()
    assert hosts.tolist() == [1, 0]  # slice position 0 = infectant 1 (C chain) = left, position 1 = infectant 0 (G chain) = right
    g.replay_recombinations([0], [0], [1], [10], [90], [0])  # the recombinant replaces slice position 0
    g.replay_mutations(np.zeros(3, dtype=np.uint64), np.zeros(0, dtype=np.uint32), np.zeros(0, dtype=np.uint8))
    g.mutate_infectants()
    assert sequence(g, 1, 100) == [G] * 10 + [C] * 80 + [G] * 10
    assert sequence(g, 0, 100) == [G] * 100


def test_library_doc_example():
    # src/lib.rs:122-141: recombinant(mutant1, mutant2, 0, 2) == A A C A
    g = sim([A] * 4, 2, H=1)
    mutate(g, {1: [(2, T)], 0: [(1, G), (2, C)]})  # slice order is [1, 0]: g[0] = mutant1, g[1] = mutant2
    assert sequence(g, 1, 4) == [A, A, T, A]
    assert sequence(g, 0, 4) == [A, G, C, A]
    g.replay_infections(np.zeros(2, dtype=np.int64))
    g.infect()
    g.replay_recombinations([0], [0], [1], [0], [2], [0])
    g.replay_mutations(np.zeros(3, dtype=np.uint64), np.zeros(0, dtype=np.uint32), np.zeros(0, dtype=np.uint8))
    g.mutate_infectants()
    assert sequence(g, 1, 4) == [A, A, C, A]


def test_merge_equivalence_and_back_mutation():
    # src/core/haplotype.rs:1073-1117 and :697-698
    g = sim([A, T, C, G], 2, H=1)
    for b in (T, C, G):
        mutate(g, {0: [(0, b)]})
    mutate(g, {1: [(1, G)]})
    mutate(g, {1: [(1, T)]})
    m = mutations(g)
    assert m[0] == {0: G}
    assert m[1] == {}
    assert g.get_base(1, 1) == T


def test_duplicate_site_in_one_event():
    # sites are drawn with replacement (src/simulation.rs:94-103): get_base sees the first change, the mutation set the last
    g = sim([A, T, C, G], 1, H=1)
    mutate(g, {0: [(2, G), (2, T)]})
    assert g.get_base(0, 2) == G
    assert mutations(g)[0] == {2: T}


def test_same_base_assertion():
    # src/core/haplotype.rs:250
    g = sim([A, T, C, G], 1, H=1)
    mutate(g, {0: [(0, A)]})
    with pytest.raises(ReferencePanic):
        g.synchronize()


def test_update_fitness_fold_order():
    # src/core/fitness/table.rs:65-68
    tab = np.array([[1.0, 0.3, 0.7, 1.1], [0.9, 1.0, 1.3, 0.2], [1.7, 0.6, 1.0, 0.4]])
    g = sim([A, T, C], 1, H=1, table=tab)
    mutate(g, {0: [(0, T), (1, C), (2, G)]})
    acc = 1.0
    for pos, frm, to in ((0, A, T), (1, T, C), (2, C, G)):
        acc = acc * tab[pos, to] / tab[pos, frm]
    assert g.get_raw_fitness()[0] == 1.0 * acc
    mutate(g, {0: [(0, G)]})
    assert g.get_raw_fitness()[0] == (1.0 * acc) * (1.0 * tab[0, G] / tab[0, T])


def test_epistasis_quirks():
    # src/core/fitness/epistasis.rs:105-193 — same expectations as tests/test_oracle_kat.py
    L = 128
    epi = np.array([[100, T, 102, C, 0.5], [20, G, 50, T, 1.1]], dtype=np.float64)
    g = sim(np.zeros(L, dtype=np.uint8), 3, H=1, table=np.ones((L, 4)), epi=epi)
    mutate(g, {0: [(100, T)], 1: [(100, T)], 2: [(100, T), (102, C)]})
    assert g.get_raw_fitness().tolist() == [1.0, 1.0, 0.5]
    mutate(g, {1: [(102, C)]})
    assert g.get_raw_fitness().tolist() == [1.0, 0.5, 0.5]


def test_replay_validation_errors():
    g = sim([A, T, C, G], 4)
    with pytest.raises(VirolutionError):
        g.replay_infections(np.zeros(3, dtype=np.int64))
    g.replay_infections(np.array([0, 1, 7, 0], dtype=np.int64))  # host 7 does not exist
    g.infect()
    with pytest.raises(VirolutionError):
        g.synchronize()
