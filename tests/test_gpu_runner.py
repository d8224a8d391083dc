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
    r = GpuRunner(str(tmp_path / "settings.yaml"), str(tmp_path / "ref.fasta"), generations=6, n_compartments=2, outdir=str(out), seed=5,
                  sampling_format="csv", name="trial")
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
        for g in (0, 6, 12):  # sample events at loop generations 0, 3, 6; labels count every compartment's increment
            srows = list(csv.reader(open(out / "samples" / f"sample_{g}_{c}.csv")))
            assert srows[0] == ["block_id", "compartment_id", "generation", "haplotype", "count"]
            assert sum(int(x[4]) for x in srows[1:]) == 100 and all(x[1] == str(c) and x[2] == str(g) for x in srows[1:])
    sizes = [eval(l.split("=")[1]) for l in log.splitlines() if "population_sizes" in l]
    assert len(sizes) == 7 and r.infectant_generations == sum(sum(x) for x in sizes[:-1])
    # src/barcode.rs: header + one line per (sample event, compartment)
    barcodes = (out / "samples" / "barcodes.csv").read_text().splitlines()
    assert barcodes[0] == "barcode,experiment,time,replicate,compartment"
    assert barcodes[1:] == [f"sample_{g}_{c},trial,{g},0,{c}" for g in (0, 6, 12) for c in (0, 1)]
    r.close()


def test_runner_fasta_samples(tmp_path):
    """FastaSampleWriter (src/readwrite/sample_writer.rs:161-208, test_fasta_write :317-345): two lines per sampled
    infectant, head `sequence_id=K;block_id=B;compartment_id=C;generation=G`, the full sequence on one line; every
    sequence differs from the wildtype exactly where the haplotype's mutation list says."""
    fx = _setup(tmp_path, EXPONENTIAL)
    out = tmp_path / "out"
    r = GpuRunner(str(tmp_path / "settings.yaml"), str(tmp_path / "ref.fasta"), generations=3, n_compartments=1, outdir=str(out), seed=9)
    r.start()
    wt = encode_sequence(fx["phix174_wildtype"])
    lines = (out / "samples" / "sample_0_0.fasta").read_text().splitlines()
    assert len(lines) == 200 and lines[0].startswith(">sequence_id=0;block_id=") and lines[0].endswith(";compartment_id=0;generation=0")
    assert all(l == wt for l in lines[1::2])  # generation 0: all wildtype
    lines = (out / "samples" / "sample_3_0.fasta").read_text().splitlines()
    assert len(lines) == 200 and [l.split(";")[0] for l in lines[0::2]] == [f">sequence_id={k}" for k in range(100)]
    n_diff = [sum(a != b for a, b in zip(l, wt)) for l in lines[1::2]]
    assert all(len(l) == len(wt) for l in lines[1::2]) and max(n_diff) >= 1 and max(n_diff) <= 12
    # the same block id <=> the same sequence
    by_block = {}
    for head, seq in zip(lines[0::2], lines[1::2]):
        assert by_block.setdefault(head.split(";")[1], seq) == seq
    r.close()


def test_block_id_is_a_stable_letter_code():
    assert block_id(0) == "a" and block_id(25) == "z" and block_id(26) == "ba" and block_id(27) == "bb"


def test_transfer_tables_must_cover_the_compartments(tmp_path):
    _setup(tmp_path, EXPONENTIAL)
    with pytest.raises(ValueError):  # Runner::new, src/runner.rs:78-85: the inline 2x2 matrix cannot serve 3 compartments
        GpuRunner(str(tmp_path / "settings.yaml"), str(tmp_path / "ref.fasta"), generations=1, n_compartments=3, outdir=str(tmp_path / "o"))


SWITCHED = """mutation_rate: 1e-4
recombination_rate: 0
host_population_size: 1500
basic_reproductive_number: 100.0
max_population: 1200
dilution: 0.02
substitution_matrix:
  - [0.0, 1.0, 1.0, 1.0]
  - [1.0, 0.0, 1.0, 1.0]
  - [1.0, 1.0, 0.0, 1.0]
  - [1.0, 1.0, 1.0, 0.0]
host_model: !SingleHost
  reproductive:
    distribution: !Exponential
      weights: {beneficial: 0.29, deleterious: 0.51, lethal: 0.2, neutral: 0.0}
      lambda_beneficial: 0.03
      lambda_deleterious: 0.21
    utility: !Algebraic
      upper: 1.5
"""


def test_runner_settings_and_stats_events(tmp_path):
    """`settings` events swap the parameter block mid-run (src/runner.rs:373-379) and `stats` events log the device
    reductions (src/runner.rs:634-673): after generation 3 the population is capped by the new max_population; the
    logged frequencies are a distribution per site, the distance matrix is symmetric with a zero diagonal."""
    _setup(tmp_path, EXPONENTIAL)
    (tmp_path / "switched.yaml").write_text(SWITCHED)
    text = (tmp_path / "settings.yaml").read_text()
    text += """  - generation: 3
    event: settings
    value: switched.yaml
  - generation: 7
    event: settings
    value: does_not_exist.yaml
  - generation: "{} % 2"
    event: stats
    value: diversity,average-fitness,distance
"""
    (tmp_path / "settings.yaml").write_text(text)
    out = tmp_path / "out"
    r = GpuRunner(str(tmp_path / "settings.yaml"), str(tmp_path / "ref.fasta"), generations=8, n_compartments=2, outdir=str(out), seed=3)
    seen = []
    collect = r.collect_stats
    r.collect_stats = lambda g: seen.append((g, collect(g))) or seen[-1][1]
    r.start()
    log = (out / "virolution.log").read_text()
    sizes = [eval(line.split("=", 1)[1]) for line in log.splitlines() if line.strip().startswith("population_sizes=")]
    assert len(sizes) == 9
    assert all(s == 3000 for s in sizes[1][:1]) and sizes[3] == [3000, 3000]          # old cap up to the switch
    assert all(max(s) <= 1200 for s in sizes[4:]) and all(min(s) > 0 for s in sizes)  # new cap from the next generation on
    assert "Adjust settings to:" in log and "Error reading settings file `does_not_exist.yaml`" in log
    assert "mean(reproductive_fitness_0)=" in log and "distance=[" in log and "(1)frequencies=" in log
    stats = {g: s for g, s in seen if s}
    assert sorted(stats) == [0, 2, 4, 6]
    for g, s in stats.items():
        for f in s["frequencies"]:
            assert f.shape == (5386, 4) and np.allclose(f.sum(axis=1), 1.0, atol=1e-12)
        d = np.asarray(s["distance"]).reshape(2, 2)
        assert d[0, 0] == 0.0 and d[1, 1] == 0.0 and d[0, 1] == d[1, 0] and (g == 0 or d[0, 1] > 0.0)
        for idx, spec, name, mean in s["mean"]:
            assert name == "reproductive_fitness" and 0.0 < mean <= 1.5
    assert stats[0]["mean"][0][3] == 1.0  # wildtype population: utility(1.0) = 1.5 * 2 * 1 / (1 + 2 * 1) = 1.0
    r.close()
