"""Cross-process migrant images on the device (vl_migrants_export_device / vl_migrants_import_device) and the
batched per-target subsample (vl_subsample_populations): what the NCCL exchange of virolution_b200/distributed.py
moves.  Both handles live on cuda:0 here; the image ranges are handed over as raw device pointers."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _grown(case_name, compartment_id, generations=4):
    c = helpers.case(case_name)
    g = helpers.make_gpu(c, compartment_id=compartment_id)
    g.run(generations)
    g.synchronize()
    return c, g


@pytest.mark.parametrize("case_name", ["singlehost", "epistasis_dense", "multihost"])
def test_export_import_roundtrip(case_name):
    c, a = _grown(case_name, 0)
    _, b = _grown(case_name, 1, generations=2)
    a.increment_generation()
    a.infect()
    a.mutate_infectants()
    a.replicate_infectants(to_host=False)
    ts = a.target_size()
    factors = [0.5, 0.0, 0.3]
    parts = a.subsample_populations(factors)
    assert [len(p) for p in parts] == [int(f * ts) for f in factors]
    img, m, w = a.export_migrants(parts)
    P = int(img.n_providers)
    assert m.tolist() == [len(p) for p in parts] and int(img.n_migrants) == int(m.sum()) and int(img.n_words) == int(w.sum())
    got, mo, wo = [], 0, 0
    for k in range(len(parts)):
        got.append(b.import_migrants(img.len_device + 4 * mo if m[k] else 0, (img.raw_device or 0) + 8 * P * mo if m[k] else 0,
                                     (img.entries_device or 0) + 4 * wo if w[k] else 0, int(m[k]), int(w[k])))
        mo += int(m[k])
        wo += int(w[k])
    b.set_population(got)
    a.set_population(parts)
    assert a.population_len() == b.population_len() == int(m.sum())
    oa, pa, ba = a.export_mutations()
    ob, pb, bb = b.export_mutations()
    assert np.array_equal(oa, ob) and np.array_equal(pa, pb) and np.array_equal(ba, bb)
    for sp in range(len(c["specs"])):
        for which in (0, 1):
            helpers.assert_f64_equal(a.get_raw_fitness(sp, which), b.get_raw_fitness(sp, which), f"raw fitness {sp}/{which}")
    # the wildtype is never shipped: it keeps id 0 on the target
    ids_a, ids_b = a.get_haplotype_ids(), b.get_haplotype_ids()
    assert np.array_equal(ids_a == 0, ids_b == 0)
    b.run(3)
    b.synchronize()
    assert b.population_len() > 0
    a.close()
    b.close()


def test_import_rejects_inconsistent_word_count():
    from virolution_b200.simulation import VirolutionError
    _, a = _grown("singlehost", 0)
    _, b = _grown("singlehost", 1, generations=1)
    a.increment_generation(); a.infect(); a.mutate_infectants(); a.replicate_infectants(to_host=False)
    parts = a.subsample_populations([1.0])
    img, m, w = a.export_migrants(parts)
    assert int(w[0]) > 0
    with pytest.raises(VirolutionError):
        b.import_migrants(img.len_device, img.raw_device, img.entries_device, int(m[0]), int(w[0]) - 1)
    a.close()
    b.close()


def test_empty_export():
    _, a = _grown("neutral", 0, generations=1)
    img, m, w = a.export_migrants([])
    assert int(img.n_migrants) == 0 and int(img.n_words) == 0 and m.size == 0
    a.close()


def test_many_parts_in_one_pass_follow_the_sampling_law():
    """vl_subsample_populations draws up to 16 parts per plan/resample pass (more parts = more passes): every part has
    exactly floor(factor * target_size) members, each an iid draw proportional to offspring
    (Population::sample, src/core/population.rs:385-403), and the parts are independent draws."""
    from test_gpu_stochastic import big_singlehost
    n = 6000
    c = big_singlehost(n=n, mu=1e-4, seed=77)
    L = len(c["wildtype"])
    g = helpers.make_gpu(c)
    rng = np.random.default_rng(5)
    # give every infectant its own haplotype: one unique (site, base) change each
    g.increment_generation()
    g.replay_infections(rng.integers(0, n, size=n))
    g.infect()
    sites = (np.arange(n) // 3).astype(np.uint32)
    assert sites.max() < L
    bases = ((c["wildtype"][sites] + 1 + np.arange(n) % 3) % 4).astype(np.uint8)
    g.replay_mutations(np.arange(n + 1, dtype=np.uint64), sites, bases)
    g.mutate_infectants()
    ids = g.get_haplotype_ids().astype(np.int64)
    assert np.unique(ids).size == n
    parent_of = np.full(ids.max() + 1, -1, dtype=np.int64)
    parent_of[ids] = np.arange(n)
    off = rng.poisson(1.3, size=n).astype(np.uint64)
    off[rng.integers(0, n, 20)] = 30
    off[1000:1500] = 0
    ts = g.target_size(off)  # uploads the map the parts are drawn from
    assert ts == float(min(int(off.sum()), n))
    factors = [0.3, 0.2, 0.1] + [0.02] * 17  # 20 parts: two passes
    parts = g.subsample_populations(factors)
    sizes = [len(p) for p in parts]
    assert sizes == [int(f * ts) for f in factors]
    g.set_population(parts)
    drawn = parent_of[g.get_haplotype_ids().astype(np.int64)]
    assert drawn.min() >= 0 and np.all(off[drawn] > 0)
    bounds = np.cumsum([0] + sizes)
    w = off / off.sum()
    for k in range(len(parts)):
        d = drawn[bounds[k]:bounds[k + 1]]
        counts = np.bincount(d, minlength=n)
        for cls in np.unique(off[off > 0]):
            sel = off == cls
            e = w[sel].sum() * d.size
            if e > 20:
                assert abs(counts[sel].sum() - e) / np.sqrt(e) < 5.5, (k, cls)
    # all parts together are one big iid sample as well
    counts = np.bincount(drawn, minlength=n)
    heavy = np.nonzero(off == 30)[0]
    e = w[heavy] * drawn.size
    assert abs(counts[heavy].sum() - e.sum()) / np.sqrt(e.sum()) < 5.5
    assert not np.array_equal(drawn[bounds[3]:bounds[4]], drawn[bounds[4]:bounds[5]]), "equal-sized parts must be different draws"
    g.close()


def test_peer_exchange_equals_in_process_transmission():
    """vl_exchange_step (migrants written by the kernels into the peer's receive buffer, flags waited on by the device;
    both "ranks" live in this process here and hand each other raw pointers instead of IPC handles) against vl_step's
    parallel-build rule (src/runner.rs:401-416) on the same seeds: the same populations generation after generation."""
    from test_gpu_stochastic import big_singlehost, population_signature
    from virolution_b200.simulation import PeerExchangeHandle, step
    c = big_singlehost(n=120000, mu=1e-4, seed=55)
    T = np.array([[0.8, 0.3], [0.2, 0.7]])
    a = [helpers.make_gpu(c, compartment_id=k) for k in range(2)]  # reference arm: vl_step
    b = [helpers.make_gpu(c, compartment_id=k) for k in range(2)]  # peer exchange
    ex = [PeerExchangeHandle(b[k], k, 2, T, words_per_migrant=32) for k in range(2)]
    bufs = [e.recv_buffer() for e in ex]
    for k in range(2):
        ex[k].connect(None, [bufs[t] if t != k else 0 for t in range(2)])
    for g in range(4):
        step(a, T, rule=0)
        for e in ex:      # ranks of one process push all before they pull any
            e.push()
        for s in b:
            s.synchronize()
        for e in ex:
            e.pull()
        for s in b:
            s.synchronize()
        for k in range(2):
            assert a[k].population_len() == b[k].population_len(), (g, k)
            assert population_signature(a[k]) == population_signature(b[k]), f"generation {g}, compartment {k}"
            helpers.assert_f64_equal(a[k].get_raw_fitness(), b[k].get_raw_fitness(), "raw fitness")
    assert a[0].population_len() == int(0.8 * 120000) + int(0.3 * 120000)
    for e in ex:
        e.close()
    for s in a + b:
        s.close()


def _sharded_pair(c, seed, shards=2):
    """`shards` handles that together hold ONE compartment: local host ranges, global max_population."""
    from virolution_b200.abi import Parameters
    from virolution_b200.simulation import GpuSimulation, PeerExchangeHandle
    p = c["parameters"]
    H, M = int(p.host_population_size), int(p.max_population)
    local = Parameters(p.mutation_rate, p.recombination_rate, H // shards, p.basic_reproductive_number, M, p.dilution,
                       p.substitution_matrix, p.n_hits)
    specs = []
    for s in c["specs"]:
        specs.append(type(s)(0, H // shards, s.reproductive, s.infective))
    sims = [GpuSimulation(c["wildtype"], local, specs, seed=seed, compartment_id=k, capacity_infectants=int(1.4 * M / shards) + 65536)
            for k in range(shards)]
    for s in sims:
        s.set_population_wildtype(c["n0"] // shards)
    ex = [PeerExchangeHandle(sims[k], k, shards, None, words_per_migrant=64, sharded=True, shard_seed=seed) for k in range(shards)]
    bufs = [e.recv_buffer() for e in ex]
    for k in range(shards):
        ex[k].connect(None, [bufs[t] if t != k else 0 for t in range(shards)])
    return sims, ex


def _sharded_step(sims, ex):
    for phase in ("begin", "push", "pull"):  # ranks of one process finish a phase before any starts the next
        for e in ex:
            getattr(e, phase)()
        for s in sims:
            s.synchronize()


def test_sharded_compartment_agrees_with_the_unsharded_one():
    """One compartment split over two ranks (vl_exchange_create sharded: device-side multinomial over the (target,
    origin) cells, cells travel like migrants) against the same compartment on one handle: the total population is
    exactly N' every generation, and the trajectory statistics agree in distribution over seeded replicates."""
    from scipy import stats
    from test_gpu_stochastic import big_singlehost
    R, G = 30, 5
    rows_a, rows_b = [], []
    for r in range(R):
        c = big_singlehost(n=40000, mu=1e-4, r0=6.0, dilution=1.0, seed=700 + r)
        a = helpers.make_gpu(c)
        row = []
        for _ in range(G):
            a.run(1)
            row += [a.population_len(), a.mean_fitness()]
        off, pos, base = a.export_mutations()
        row += [pos.size / a.population_len(), len(set(helpers.mutation_sets(off, pos, base)))]
        rows_a.append(row)
        a.close()
        sims, ex = _sharded_pair(c, seed=1700 + r)
        row = []
        for _ in range(G):
            _sharded_step(sims, ex)
            sizes = [s.population_len() for s in sims]
            assert sum(sizes) == 40000, sizes                     # N' = min(W * dilution, max_population) exactly
            assert abs(sizes[0] - 20000) < 6 * 100, sizes           # each new infectant picks its rank uniformly: sd = 100
            mf = sum(s.mean_fitness() * n for s, n in zip(sims, sizes)) / sum(sizes)
            row += [sum(sizes), mf]
        tot, distinct = 0, set()
        for k, s in enumerate(sims):
            off, pos, base = s.export_mutations()
            tot += pos.size
            distinct |= set(helpers.mutation_sets(off, pos, base))
        row += [tot / 40000, len(distinct)]
        rows_b.append(row)
        for e in ex:
            e.close()
        for s in sims:
            s.close()
    ra, rb = np.array(rows_a, dtype=float), np.array(rows_b, dtype=float)
    cols = [k for k in range(ra.shape[1]) if not (np.all(ra[:, k] == ra[0, k]) and np.all(rb[:, k] == ra[0, k]))]
    for k in cols:
        p = stats.ks_2samp(ra[:, k], rb[:, k]).pvalue
        assert p > 0.01 / len(cols), f"column {k}: KS p={p:.2e} (one handle mean {ra[:, k].mean():.6g}, sharded mean {rb[:, k].mean():.6g})"


def _exchange_pair(wl, seed, T, monkeypatch=None, experiment=None):
    from virolution_b200.simulation import GpuSimulation, PeerExchangeHandle
    import os
    old = os.environ.get("VL_EXPERIMENT")
    if experiment is not None:
        os.environ["VL_EXPERIMENT"] = str(experiment)  # read when the handle is created
    try:
        sims = [GpuSimulation(wl["wildtype"], wl["parameters"], wl["specs"], seed=seed, compartment_id=k) for k in range(2)]
    finally:
        if experiment is not None:
            if old is None:
                del os.environ["VL_EXPERIMENT"]
            else:
                os.environ["VL_EXPERIMENT"] = old
    for s in sims:
        s.set_population_wildtype(wl["n0"])
    ex = [PeerExchangeHandle(sims[k], k, 2, T, words_per_migrant=32) for k in range(2)]
    bufs = [e.recv_buffer() for e in ex]
    for k in range(2):
        ex[k].connect(None, [bufs[t] if t != k else 0 for t in range(2)])
    return sims, ex


def _population_state(g):
    off, pos, base = g.export_mutations()
    return off, pos, base, g.get_raw_fitness().view(np.uint64)


@pytest.mark.parametrize("name", ["k2_hiv", "k1_singlehost"])
def test_exchange_run_equals_steps_and_general_kernels(name):
    """The loop bench.py times at --gpus N (vl_exchange_run: own part drawn in place, one-pass export and import, the next
    generation's infection prepared by the draws and the import) against the same generations as single vl_exchange_step
    calls and against the general exchange kernels (several passes, own part copied): the same populations, infectant by
    infectant, on both ranks."""
    from virolution_b200.distributed import migration_matrix
    from virolution_b200.workloads import workload
    wl = workload(name, 150_000)
    if name == "k1_singlehost":
        wl["parameters"].mutation_rate = 2e-5
    T = np.asarray(migration_matrix(2, 0.1), dtype=np.float64)
    arms = {"run": _exchange_pair(wl, 31, T), "steps": _exchange_pair(wl, 31, T), "general": _exchange_pair(wl, 31, T, experiment=256),
            "overlap": _exchange_pair(wl, 31, T, experiment=1024)}  # (own part mutated on the main stream while the migrants travel on a side stream)
    G = 5
    for arm in ("run", "overlap"):
        for e in arms[arm][1]:          # both ranks enqueue all generations; they meet on the device
            e.run(G)
    for arm in ("steps", "general"):
        sims, ex = arms[arm]
        for _ in range(G):
            for e in ex:
                e.push()
            for e in ex:
                e.pull()
    for sims, _ in arms.values():
        for s in sims:
            s.synchronize()
    for k in range(2):
        ref = _population_state(arms["steps"][0][k])
        assert arms["steps"][0][k].population_len() > 100_000
        assert (np.diff(ref[0].astype(np.int64)) > 0).sum() > 1000, "the run must have produced mutants"
        for arm in ("run", "general", "overlap"):
            got = _population_state(arms[arm][0][k])
            for x, y, what in zip(got, ref, ("mutation offsets", "positions", "bases", "raw fitness bits")):
                assert np.array_equal(x, y), f"{name}, rank {k}, {arm} vs steps: {what} differ"
    # and the run continues from the same state: two more generations, compaction forced in between
    for arm in ("run", "steps"):
        sims, ex = arms[arm]
        for s in sims:
            s.collect_garbage()
    for e in arms["run"][1]:
        e.run(2)
    sims, ex = arms["steps"]
    for _ in range(2):
        for e in ex:
            e.push()
        for e in ex:
            e.pull()
    for k in range(2):
        a, b = _population_state(arms["run"][0][k]), _population_state(arms["steps"][0][k])
        for x, y in zip(a, b):
            assert np.array_equal(x, y), f"{name}, rank {k}: after compaction"
    for sims, ex in arms.values():
        for e in ex:
            e.close()
        for s in sims:
            s.close()


def test_exchange_follows_a_transfer_matrix_that_changes():
    """A `transmission` event of the schedule replaces the matrix mid-run (src/config/schedule.rs:196-266): the peer exchange,
    created with the entry-wise maximum of the schedule's matrices, against vl_step called with the matrix of each
    generation.  A matrix with an entry above the creation matrix is refused."""
    from test_gpu_stochastic import big_singlehost, population_signature
    from virolution_b200.simulation import PeerExchangeHandle, VirolutionError, step
    c = big_singlehost(n=100000, mu=1e-4, seed=91)
    T1 = np.array([[0.9, 0.1], [0.1, 0.9]])
    T2 = np.array([[0.6, 0.35], [0.4, 0.65]])
    Tcap = np.maximum(T1, T2)
    a = [helpers.make_gpu(c, compartment_id=k) for k in range(2)]
    b = [helpers.make_gpu(c, compartment_id=k) for k in range(2)]
    ex = [PeerExchangeHandle(b[k], k, 2, Tcap, words_per_migrant=32) for k in range(2)]
    bufs = [e.recv_buffer() for e in ex]
    for k in range(2):
        ex[k].connect(None, [bufs[t] if t != k else 0 for t in range(2)])
    for T in (T1, T1, T2, T2, T1):
        step(a, T, rule=0)
        for e in ex:
            e.set_transfer(T)
        for e in ex:
            e.push()
        for e in ex:
            e.pull()
        for k in range(2):
            assert a[k].population_len() == b[k].population_len() == int(T[k, 0] * 100000) + int(T[k, 1] * 100000)
            assert population_signature(a[k]) == population_signature(b[k])
    with pytest.raises(VirolutionError):
        ex[0].set_transfer(np.array([[0.95, 0.1], [0.1, 0.9]]))
    for e in ex:
        e.close()
    for s in a + b:
        s.close()


@pytest.mark.parametrize("which", ["multihost", "k3_multihost"])
def test_compartment_group_equals_vl_step(which):
    """CompartmentGroup (the compartments of one process stepped through the peer exchange: what the runner mirror and the K3
    block of the bench use) against vl_step's parallel-build rule on three compartments with two host specs each — infectant
    by infectant — with a transfer matrix that changes, then `run` (chunks of generations enqueued per compartment)."""
    from test_gpu_stochastic import population_signature
    from virolution_b200.simulation import CompartmentGroup, step
    from virolution_b200.workloads import workload
    if which == "multihost":
        c = helpers.case("multihost")     # infective fitness, n_hits 3: the general kernels
        mk = lambda k, seed: helpers.make_gpu(c, seed=seed, compartment_id=k)
    else:
        wl = workload("k3_multihost", 30000)
        wl["parameters"].mutation_rate = 2e-5

        def mk(k, seed):
            from virolution_b200.simulation import GpuSimulation
            g = GpuSimulation(wl["wildtype"], wl["parameters"], wl["specs"], seed=seed, compartment_id=k)
            g.set_population_wildtype(wl["n0"])
            return g
    T1 = np.array([[0.8, 0.1, 0.0], [0.2, 0.7, 0.3], [0.0, 0.2, 0.7]])
    T2 = np.array([[0.6, 0.0, 0.2], [0.2, 0.9, 0.0], [0.2, 0.1, 0.8]])
    a = [mk(k, 21) for k in range(3)]
    b = [mk(k, 21) for k in range(3)]
    group = CompartmentGroup(b, np.maximum(T1, T2), words_per_migrant=48)
    for T in (T1, T2, T2, T1):
        step(a, T, 0)
        group.step(T)
        for k in range(3):
            assert a[k].population_len() == b[k].population_len()
            assert population_signature(a[k]) == population_signature(b[k]), f"compartment {k}"
    for _ in range(11):
        step(a, T1, 0)
    group.run(11)             # more than one chunk
    group.synchronize()
    for k in range(3):
        assert population_signature(a[k]) == population_signature(b[k]), f"compartment {k} after run"
        helpers.assert_f64_equal(a[k].get_raw_fitness(), b[k].get_raw_fitness(), "raw fitness")
    group.close()
    for s in a + b:
        s.close()


def test_fast_exchange_three_compartments_with_closed_links_and_small_blocks():
    """The one-pass export / import with three compartments of one process, a transfer matrix with zero entries (targets that
    receive nothing from an origin) and an empty compartment, against vl_step; then blocks that are too small for the
    migrants' records: a capacity error, not a wrong population."""
    from test_gpu_stochastic import population_signature
    from virolution_b200.simulation import CompartmentGroup, GpuSimulation, VirolutionError, step
    from virolution_b200.workloads import workload
    wl = workload("k2_hiv", 60_000)
    wl["parameters"].mutation_rate = 1e-4

    def mk(k, n0):
        g = GpuSimulation(wl["wildtype"], wl["parameters"], wl["specs"], seed=8, compartment_id=k)
        g.set_population_wildtype(n0)
        return g
    T = np.array([[0.7, 0.0, 0.2], [0.3, 0.9, 0.0], [0.0, 0.1, 0.8]])
    a = [mk(k, n0) for k, n0 in enumerate((60_000, 60_000, 0))]      # compartment 2 starts empty and is colonised by 1
    b = [mk(k, n0) for k, n0 in enumerate((60_000, 60_000, 0))]
    group = CompartmentGroup(b, T)
    for g in range(6):
        step(a, T, 0)
        group.step()
        for k in range(3):
            assert a[k].population_len() == b[k].population_len(), (g, k)
            assert population_signature(a[k]) == population_signature(b[k]), (g, k)
    assert b[2].population_len() > 0
    group.close()
    for s in a + b:
        s.close()
    # blocks of one word per migrant: after a few generations the migrants' records do not fit
    c = [mk(k, 60_000) for k in range(2)]
    small = CompartmentGroup(c, np.array([[0.5, 0.5], [0.5, 0.5]]), words_per_migrant=1)
    with pytest.raises(VirolutionError, match="capacity|arena|words"):
        for _ in range(12):
            small.step()
        small.synchronize()
    small.close()
    for s in c:
        s.close()
