"""Population statistics and the final.<i>.csv rows on the device state: frequencies and mean of the
mapped fitness (src/stats/population.rs:15-75), Haplotype::get_string rows (src/core/haplotype.rs:437-453,
src/runner.rs:137-152), checked against the same quantities computed on the host from the exported mutation sets."""
import csv

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["singlehost", "epistasis_dense", "multihost"])
def test_frequencies_mean_and_final_csv(name, tmp_path):
    c = helpers.case(name)
    g = helpers.make_gpu(c)
    g.run(c["generations"])
    g.synchronize()
    n, L = g.population_len(), len(c["wildtype"])
    off, pos, base = g.export_mutations()
    # frequencies: counts of mutation-set entries, the wildtype column is n minus the row sum (the reference's rule)
    cnt = np.zeros((L, 4), dtype=np.int64)
    np.add.at(cnt, (pos.astype(np.int64), base.astype(np.int64)), 1)
    row = cnt.sum(axis=1)
    cnt[np.arange(L), c["wildtype"]] = n - row
    expect = cnt / float(n)
    got = g.frequencies()
    assert np.array_equal(got.view(np.uint64), expect.view(np.uint64)), "frequencies are counts / n: must be bit-identical"
    assert np.allclose(got.sum(axis=1), 1.0)
    # mean mapped fitness per provider
    for sp, spec in enumerate(c["specs"]):
        for which, prov in enumerate((spec.reproductive, spec.infective)):
            m = g.mean_fitness(sp, which)
            if prov is None:
                assert m == 1.0
                continue
            raw = g.get_raw_fitness(sp, which)
            ref = float(np.mean(prov.utility.apply(raw))) if prov.kind != "Neutral" else 1.0
            assert abs(m - ref) <= 1e-12 * abs(ref), (sp, which, m, ref)
    # final.<i>.csv
    path = tmp_path / "final.0.csv"
    g.write_final_csv(str(path), c["wildtype"])
    rows = list(csv.reader(open(path)))
    assert rows[0] == ["haplotype", "count"]
    assert sum(int(r[1]) for r in rows[1:]) == n
    sets = helpers.mutation_sets(off, pos, base)
    letters = "ATCG"
    seen = {}
    for s in sets:
        text = ";".join(f"{p}:{letters[c['wildtype'][p]]}->{letters[b]}" for p, b in s) or "wt"
        seen[text] = seen.get(text, 0) + 1
    agg = {}
    for text, k in rows[1:]:
        agg[text] = agg.get(text, 0) + int(k)
    assert agg == seen, "rows aggregated by haplotype string must equal the population's mutation sets (SURVEY.md D7)"
    # distance to a copy of itself is 0, to another replicate it is the L1 distance of the frequency vectors
    h = helpers.make_gpu(c, seed=c["seed"] + 1)
    h.run(c["generations"])
    assert g.distance(g) == 0.0
    d = g.distance(h)
    assert d == float(np.cumsum(np.abs(got.ravel() - h.frequencies().ravel()))[-1]) and d >= 0.0
    g.close()
    h.close()
