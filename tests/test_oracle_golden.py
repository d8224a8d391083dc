"""The committed oracle replay vectors (tests/golden/oracle_replay_*.npz) are reproduced bit for bit by the current
oracle build — a regression pin: the GPU parity tests replay these files, so the checker must not drift silently."""
import os

import numpy as np
import pytest

import helpers


@pytest.mark.parametrize("name", helpers.GOLDEN_CASES)
def test_oracle_reproduces_golden(name):
    gold = dict(np.load(os.path.join(helpers.GOLDEN, f"oracle_replay_{name}.npz")))
    now = helpers.record_oracle_run(name)
    assert set(gold) == set(now)
    for k in sorted(gold):
        a, b = gold[k], now[k]
        assert a.shape == b.shape, k
        if a.dtype.kind == "f":
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), k
        else:
            assert np.array_equal(a, b), k


def test_oracle_generation_invariants():
    # analytic invariants that do not depend on the restated samplers' streams
    c = helpers.case("singlehost")
    w = helpers.make_oracle(c)
    for _ in range(3):
        log = helpers.oracle_generation(w)
        n = int(log["n"][0])
        assert np.all(log["infections"] >= 0)  # SingleHost, infectivity 1, n_hits 1: everybody infects
        total = int(log["offspring"].sum())
        expect = min(total * c["parameters"].dilution, float(c["parameters"].max_population))
        assert log["target_size"][0] == expect
        assert int(log["n_after"][0]) == int(expect)
        assert log["hostmap_offsets"][-1] == n
