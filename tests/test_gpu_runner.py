"""The host-side mirror of `Runner` (virolution_b200/runner.py) on the reference's integration-test shape
(src/main.rs:63-154: run N compartments for a few generations, then look for `<provider>_table.npy`, the log and the
per-compartment outputs) — with the reference's own fixtures (phiX174, tests/epistasis/*.npy, re-encoded in
tests/golden/reference_fixtures.npz) and its YAML dialect."""
import csv
import os

import numpy as np
import pytest

import helpers
from virolution_b200.config import encode_sequence
from virolution_b200.runner import GpuRunner, block_id

pytestmark = pytest.mark.gpu

YAML = """
parameters:
  - mutation_rate: 1e-4
    recombination_rate: 0
    host_population_size: 3000
    basic_reproductive_number: 100.0
    max_population: 3000
    dilution: 0.02
    substitution_matrix:
      - [0.0, 1.0, 1.0, 1.0]
      - [1.0, 0.0, 1.0, 1.0]
      - [1.0, 1.0, 0.0, 1.0]
      - [1.0, 1.0, 1.0, 0.0]
    host_model: !SingleHost
      reproductive:
        distribution: %s
        utility: !Algebraic
          upper: 1.5
schedule:
  - generation: "{} %% 1"
    event: transmission
    value: "[[0.9, 0.1], [0.1, 0.9]]"
  - generation: "{} %% 3"
    event: sample
    value: 100
"""
EXPONENTIAL = """!Exponential
          weights:
            beneficial: 0.29
            deleterious: 0.51
            lethal: 0.2
            neutral: 0.0
          lambda_beneficial: 0.03
          lambda_deleterious: 0.21"""
EPISTATIC = """!Epistatic
          path: fitness_table.npy
          epi_path: epistasis_table.npy"""


def _setup(tmp_path, distribution):
    fx = helpers.fixtures()
    (tmp_path / "ref.fasta").write_text(">phiX174\n" + encode_sequence(fx["phix174_wildtype"]) + "\n")
    (tmp_path / "settings.yaml").write_text(YAML % distribution)
    np.save(tmp_path / "fitness_table.npy", fx["epistasis_fitness"])
    pairs = fx["epistasis_pairs"]
    rec = np.zeros(len(pairs), dtype=[("pos1", "<u8"), ("base1", "u1"), ("pos2", "<u8"), ("base2", "u1"), ("value", "<f8")])
    for k, f in enumerate(("pos1", "base1", "pos2", "base2", "value")):
        rec[f] = pairs[:, k]
    np.save(tmp_path / "epistasis_table.npy", rec)
    return fx


@pytest.mark.parametrize("distribution", [EXPONENTIAL, EPISTATIC], ids=["singlehost", "epistasis"])
def test_runner_outputs(tmp_path, distribution):
    fx = _setup(tmp_path, distribution)
    out = tmp_path / "out"
    r = GpuRunner(str(tmp_path / "settings.yaml"), str(tmp_path / "ref.fasta"), generations=6, n_compartments=2, outdir=str(out), seed=5)
    r.start()
    # what the reference's integration tests assert: table file and log exist (src/main.rs:84-87)
    assert (out / "reproductive_fitness_0_table.npy").exists() and (out / "virolution.log").exists()
    table = np.load(out / "reproductive_fitness_0_table.npy")
    assert table.shape == (5386, 4) and np.all(table[np.arange(5386), fx["phix174_wildtype"]] == 1.0)
    if distribution is EPISTATIC:
        assert np.array_equal(table, fx["epistasis_fitness"].reshape(-1, 4))
        assert len(np.load(out / "reproductive_fitness_0_epistasis_table.npy")) == len(fx["epistasis_pairs"])
    log = (out / "virolution.log").read_text()
    assert "generation=0" in log and "generation=6" in log and "population_sizes=[3000, 3000]" in log
    for c, sim in enumerate(r.sims):
        rows = list(csv.reader(open(out / f"final.{c}.csv")))
        assert rows[0] == ["haplotype", "count"]
        assert sum(int(x[1]) for x in rows[1:]) == sim.population_len()
        assert any(x[0] == "wt" for x in rows[1:])
        assert any("->" in x[0] for x in rows[1:]), "six generations at mu*L = 0.54 must leave mutants"
        for g in (0, 3, 6):  # sample events: generation % 3 == 0, the generation attribute is the simulation's own counter
            srows = list(csv.reader(open(out / f"sample_{g}_{c}.csv")))
            assert srows[0] == ["block_id", "compartment_id", "generation", "haplotype", "count"]
            assert sum(int(x[4]) for x in srows[1:]) == 100 and all(x[1] == str(c) and x[2] == str(g) for x in srows[1:])
    sizes = [eval(l.split("=")[1]) for l in log.splitlines() if "population_sizes" in l]
    assert len(sizes) == 7 and r.infectant_generations == sum(sum(x) for x in sizes[:-1])
    r.close()


def test_block_id_is_a_stable_letter_code():
    assert block_id(0) == "a" and block_id(25) == "z" and block_id(26) == "ba" and block_id(27) == "bb"


def test_transfer_tables_must_cover_the_compartments(tmp_path):
    _setup(tmp_path, EXPONENTIAL)
    with pytest.raises(ValueError):  # Runner::new, src/runner.rs:78-85: the inline 2x2 matrix cannot serve 3 compartments
        GpuRunner(str(tmp_path / "settings.yaml"), str(tmp_path / "ref.fasta"), generations=1, n_compartments=3, outdir=str(tmp_path / "o"))
