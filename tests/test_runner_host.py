"""Host-side logic of the runner mirror (virolution_b200/runner.py) that needs no device: the strings and file formats the
reference's writers produce, and the schedule events the loop reacts to."""
import numpy as np
import pytest

from virolution_b200.config import Schedule, ScheduleRecord, load_parameters
from virolution_b200.runner import GpuRunner, block_id, haplotype_string


def test_haplotype_string_like_get_string():
    """Haplotype::get_string (src/core/haplotype.rs:437-453): `pos:FROM->TO` joined by `;`, FROM = the wildtype base,
    `wt` for no mutation."""
    wt = np.array([0, 1, 2, 3, 0, 0], dtype=np.uint8)  # A T C G A A
    assert haplotype_string(wt, [], []) == "wt"
    assert haplotype_string(wt, [1], [3]) == "1:T->G"
    assert haplotype_string(wt, [0, 3, 5], [2, 0, 1]) == "0:A->C;3:G->A;5:A->T"


def test_block_id_letter_code():
    assert [block_id(k) for k in (0, 1, 25, 26, 27, 26 * 26)] == ["a", "b", "z", "ba", "bb", "baa"]
    assert len({block_id(k) for k in range(5000)}) == 5000  # injective


def test_schedule_events_of_the_loop():
    """The events Runner::run looks up every generation (src/runner.rs:358,374,403,634): sample, settings, transmission,
    stats — first matching record wins (src/config/schedule.rs:224-231)."""
    s = Schedule([ScheduleRecord("{} % 10", "sample", "100"), ScheduleRecord("5", "settings", "other.yaml"),
                  ScheduleRecord("{} % 2", "stats", "diversity,distance"), ScheduleRecord("{} % 1", "transmission", "migration_fwd"),
                  ScheduleRecord("{} % 20", "sample", "7")])
    assert s.get_sample_size(0) == 100 and s.get_sample_size(20) == 100 and s.get_sample_size(5) == 0
    assert s.get_event_value("settings", 5) == "other.yaml" and s.get_event_value("settings", 6) is None
    assert s.get_event_value("stats", 4) == "diversity,distance" and s.get_event_value("stats", 3) is None
    assert np.array_equal(s.get_transfer_matrix(3), s.transfers["migration_fwd"])


MINIMAL = """
parameters:
  - mutation_rate: 1e-6
    recombination_rate: 0
    host_population_size: 100
    basic_reproductive_number: 10.0
    max_population: 100
    dilution: 0.1
    substitution_matrix:
      - [0.0, 1.0, 1.0, 1.0]
      - [1.0, 0.0, 1.0, 1.0]
      - [1.0, 1.0, 0.0, 1.0]
      - [1.0, 1.0, 1.0, 0.0]
    host_model: !SingleHost
      reproductive:
        distribution: !Neutral
schedule:
  - generation: "{} % 1"
    event: transmission
    value: migration_fwd
"""


def test_unknown_sampling_format_is_rejected(tmp_path):
    # src/runner.rs:329-332: "Unknown sampling format." — after the settings are read, before anything touches the device
    (tmp_path / "s.yaml").write_text(MINIMAL)
    with pytest.raises(ValueError, match="Unknown sampling format"):
        GpuRunner(str(tmp_path / "s.yaml"), np.zeros(10, dtype=np.uint8), sampling_format="bam", outdir=str(tmp_path / "o"))


def test_runner_fails_loudly_without_a_device(tmp_path):
    """There is no CPU path behind the runner: on a box without CUDA the first handle raises (the product never falls back
    to the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by tests/test_gpu_runner.py")
    (tmp_path / "s.yaml").write_text(MINIMAL)
    with pytest.raises(Exception) as err:
        GpuRunner(str(tmp_path / "s.yaml"), np.zeros(10, dtype=np.uint8), outdir=str(tmp_path / "o"))
    assert "CUDA" in str(err.value) or "device" in str(err.value).lower()


def test_settings_event_file_errors_are_reported_not_raised(tmp_path):
    """Schedule::get_settings maps an unreadable parameter file to None after logging (src/config/schedule.rs:210-222);
    load_parameters is what raises underneath."""
    with pytest.raises(Exception):
        load_parameters(str(tmp_path / "does_not_exist.yaml"))
