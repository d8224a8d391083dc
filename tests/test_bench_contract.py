"""The JSON line contract of bench.py that can be checked without a GPU: the reference arm (the oracle on the host cores)
prints one line with the keys the driver reads, its own `e2e` and `cpu_baseline`, and no device traffic."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--population", "20000"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "infectant-generations/sec" and d["unit"] == "infectant-generations/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and abs(d["value"] - 20000 * 1e3 / d["ms_per_step"]) / d["value"] < 0.05
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and "generation" in cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] == 0  # nothing of the CUDA library runs on this arm
