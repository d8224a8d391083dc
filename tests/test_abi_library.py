"""The C-ABI library loads and exports every symbol include/virolution_gpu.h declares; without a device it fails loudly."""
import ctypes
import os
import re

import numpy as np
import pytest

import helpers
from virolution_b200 import build as vbuild
from virolution_b200 import simulation as S

ROOT = helpers.ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "virolution_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vl_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    lib_path = vbuild.build()
    lib = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 40
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/virolution_gpu.h but not exported"
    assert set(names) == set(S.SIGNATURES), "python binding table and header disagree"
    assert lib.vl_abi_version() == 1


def test_struct_layouts_match_header():
    from virolution_b200 import abi
    assert ctypes.sizeof(abi.vl_utility) == 32
    assert ctypes.sizeof(abi.vl_epi_entry) == 32
    assert ctypes.sizeof(abi.vl_fitness) == 8 + 8 + 8 + 8 + 32
    assert ctypes.sizeof(abi.vl_host_spec) == 16 + 2 * ctypes.sizeof(abi.vl_fitness)
    assert ctypes.sizeof(abi.vl_parameters) == 8 * 6 + 16 * 8 + 8


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device failure mode")
def test_no_device_fails_loudly():
    c = helpers.case("neutral")
    with pytest.raises(S.VirolutionError) as e:
        helpers.make_gpu(c)
    assert e.value.code == -3 and "no CPU path" in str(e.value)
    with pytest.raises(S.VirolutionError):
        S.sampler_poisson(1, 2.0, 10)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        S.load_library(str(tmp_path / "libvirolution_gpu.so"))


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "virolution_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), f"{f} mentions the oracle"
