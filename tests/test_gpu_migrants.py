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
