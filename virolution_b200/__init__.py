"""virolution_b200 — B200-native (sm_100a) population-update path of virolution behind a C ABI.

Only the hot path lives here: the CUDA kernels + C ABI (`csrc/`, `include/virolution_gpu.h`) and the
host-side mirror of the reference's `Simulation` interface (`simulation.py`, `runner.py`).
"""
from .abi import (FitnessProvider, HostSpec, Parameters, UtilityFunction)  # noqa: F401

__all__ = ["FitnessProvider", "HostSpec", "Parameters", "UtilityFunction"]
