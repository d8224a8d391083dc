"""Host-side configuration resolution against the reference's own tests (src/config/schedule.rs:272-369,
src/config/parameters.rs:112-145)."""
import numpy as np
import pytest

from virolution_b200.abi import Parameters
from virolution_b200.config import (NAMED_TRANSFERS, HostFitness, HostModel, FitnessModel, Schedule, ScheduleRecord,
                                    decode_sequence, encode_sequence, load_settings)


def sched(rows):
    return Schedule([ScheduleRecord(*r) for r in rows])


def test_transfer_matrix_named():
    # schedule.rs:272-311
    plan = sched([("1", "transmission", "migration_fwd"), ("2", "transmission", "migration_rev"), ("3", "transmission", "root_ab"),
                  ("4", "transmission", "root_bc"), ("5", "transmission", "root_cd")])
    for g, name in enumerate(["default", "migration_fwd", "migration_rev", "root_ab", "root_bc", "root_cd"]):
        assert np.array_equal(plan.get_transfer_matrix(g), np.asarray(NAMED_TRANSFERS[name]))


def test_custom_transfer_matrix():
    # schedule.rs:313-324
    plan = sched([("1", "transmission", "[[1, 2], [3, 4]]")])
    assert np.array_equal(plan.get_transfer_matrix(1), np.array([[1.0, 2.0], [3.0, 4.0]]))


def test_sample_size():
    # schedule.rs:326-340
    plan = sched([("1", "sample", "100"), ("2", "sample", "200"), ("3", "sample", "300")])
    assert [plan.get_sample_size(g) for g in range(4)] == [0, 100, 200, 300]


def test_expressions():
    # schedule.rs:342-368
    plan = sched([("{} % 200", "sample", "100"), ("{} % 10", "transmission", "migration_fwd"),
                  ("(5 + {}) % 10", "transmission", "migration_rev")])
    for i in range(0, 1001):
        assert plan.get_sample_size(i) == (100 if i % 200 == 0 else 0)
        name = "migration_fwd" if i % 10 == 0 else ("migration_rev" if (5 + i) % 10 == 0 else "default")
        assert np.array_equal(plan.get_transfer_matrix(i), np.asarray(NAMED_TRANSFERS[name]))


def test_nucleotide_index_order():
    # src/encoding.rs:31-73: A,T,C,G = 0..3; decode accepts upper/lower case and digits
    assert decode_sequence("ATCGatcg0123").tolist() == [0, 1, 2, 3] * 3
    assert encode_sequence(np.array([0, 1, 2, 3])) == "ATCG"
    with pytest.raises(ValueError):
        decode_sequence("ATXG")


def test_multihost_ranges():
    # src/config/parameters.rs:126-142: n = round(fraction*H); lower = i*n; upper = min(lower+n, H)
    par = Parameters(1e-6, 0.0, 1001, 100.0, 1000, 0.02)
    hm = HostModel("MultiHost", fractions=[(0.5, HostFitness(FitnessModel("Neutral"), None)), (0.5, HostFitness(FitnessModel("Neutral"), None))])
    specs = hm.make_definitions(par, np.zeros(10, dtype=np.uint8))
    assert [(s.lo, s.hi) for s in specs] == [(0, 501), (501, 1001)]  # round(500.5) = 501 (half away from zero)
    hm3 = HostModel("MultiHost", fractions=[(0.25, HostFitness()), (0.25, HostFitness()), (0.5, HostFitness())])
    par = Parameters(1e-6, 0.0, 100, 100.0, 1000, 0.02)
    assert [(s.lo, s.hi) for s in hm3.make_definitions(par, np.zeros(10, dtype=np.uint8))] == [(0, 25), (25, 50), (100, 100)]


def test_load_settings_yaml(tmp_path):
    text = """
parameters:
  - mutation_rate: 1e-6
    recombination_rate: 0
    host_population_size: 10000
    basic_reproductive_number: 100.0
    max_population: 10000
    dilution: 0.02
    substitution_matrix:
      - [0.0, 1.0, 1.0, 1.0]
      - [1.0, 0.0, 1.0, 1.0]
      - [1.0, 1.0, 0.0, 1.0]
      - [1.0, 1.0, 1.0, 0.0]
    host_model: !SingleHost
      reproductive:
        distribution: !Exponential
          weights:
            beneficial: 0.29
            deleterious: 0.51
            lethal: 0.2
            neutral: 0.0
          lambda_beneficial: 0.03
          lambda_deleterious: 0.21
        utility: !Algebraic
          upper: 1.5
schedule:
  - generation: "{} % 1"
    event: transmission
    value: "[[0.9, 0.1], [0.1, 0.9]]"
  - generation: "{} % 200"
    event: sample
    value: 1000
"""
    p = tmp_path / "s.yaml"
    p.write_text(text)
    params, host_models, schedule = load_settings(str(p))
    assert params[0].host_population_size == 10000 and params[0].dilution == 0.02 and params[0].n_hits == 1
    wt = np.random.default_rng(0).integers(0, 4, 64).astype(np.uint8)
    specs = host_models[0].make_definitions(params[0], wt, rng=np.random.default_rng(1))
    assert len(specs) == 1 and (specs[0].lo, specs[0].hi) == (0, 10000)
    prov = specs[0].reproductive
    assert prov.kind == "NonEpistatic" and prov.utility.kind == "Algebraic" and prov.utility.upper == 1.5
    assert np.all(prov.table[np.arange(64), wt] == 1.0)
    assert specs[0].infective is None
    assert np.array_equal(schedule.get_transfer_matrix(7), np.array([[0.9, 0.1], [0.1, 0.9]]))
    assert schedule.get_sample_size(400) == 1000 and schedule.get_sample_size(401) == 0


def test_load_parameters_single_block(tmp_path):
    """Parameters::read_from_file (src/config/parameters.rs:182-186): what a `settings` event points to."""
    from virolution_b200.config import load_parameters
    p = tmp_path / "p.yaml"
    p.write_text("""mutation_rate: 2e-6
recombination_rate: 1e-4
host_population_size: 777
basic_reproductive_number: 50.0
max_population: 555
dilution: 0.5
n_hits: 3
substitution_matrix:
  - [0.0, 1.0, 2.0, 1.0]
  - [1.0, 0.0, 1.0, 1.0]
  - [1.0, 1.0, 0.0, 1.0]
  - [1.0, 1.0, 1.0, 0.0]
host_model: !SingleHost
  reproductive:
    distribution: !Neutral
""")
    par = load_parameters(str(p))
    assert (par.mutation_rate, par.recombination_rate, par.host_population_size, par.basic_reproductive_number,
            par.max_population, par.dilution, par.n_hits) == (2e-6, 1e-4, 777, 50.0, 555, 0.5, 3)
    assert par.substitution_matrix[0][2] == 2.0


def test_generation_expressions_are_parsed_not_executed():
    """`eval_int_with_context` semantics of the schedule's generation column (src/config/schedule.rs:250-266): integer
    arithmetic, `/` truncating, `%` with the dividend's sign, `^` = power; anything else is "Invalid generation"."""
    from virolution_b200.config import _eval_generation_expr as ev
    assert ev("x % 10", 20) == 0 and ev("{} % 10", 23) == 3 and ev("(generation - 5) % 7", 12) == 0
    assert ev("x / 2 - 3", 7) == 0          # 7 / 2 == 3
    assert ev("-7 / 2", 0) == -3 and ev("-7 % 3", 0) == -1
    assert ev("2^3 - x", 8) == 0
    for bad in ("9**9**9**9", "__import__('os').system('true')", "x.real", "1/0", "x % 0", "x if x else 1", "1.5 * x", "[x]", "x == 1", ""):
        with pytest.raises(ValueError):
            ev(bad, 3)
    s = sched([("x % 4", "sample", "10"), ("(x - 1) % 4", "sample", "20")])
    assert [s.get_sample_size(g) for g in range(6)] == [10, 20, 0, 0, 10, 20]
