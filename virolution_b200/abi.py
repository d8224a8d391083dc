"""ctypes mirror of include/virolution_gpu.h (struct layouts + config marshalling).

Host-side plumbing only: it describes the reference's resolved configuration
(`Parameters` src/config/parameters.rs:14-47, `HostSpec` src/core/hosts.rs:75-78,
`FitnessProvider` src/providers/fitness.rs:18-22) in the C ABI's plain structs.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

VL_OK = 0
VL_ERR_INITIALIZATION = -1
VL_ERR_IMPLEMENTATION = -2
VL_ERR_CUDA = -3
VL_ERR_CAPACITY = -4
VL_ERR_REFERENCE_PANIC = -5
VL_ERR_ARGUMENT = -6

VL_FLAG_TREE_PLAN = 2
VL_FLAG_NO_BUCKETS = 4
VL_FLAG_GENERAL_REPLICATE = 8
VL_FLAG_NO_TILES = 16

UTILITY_LINEAR, UTILITY_ALGEBRAIC, UTILITY_HILL = 0, 1, 2
FITNESS_NONE, FITNESS_NEUTRAL, FITNESS_NONEPISTATIC, FITNESS_EPISTATIC = 0, 1, 2, 3


class vl_utility(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("upper", C.c_double), ("p", C.c_double), ("n", C.c_double)]


class vl_epi_entry(C.Structure):
    _fields_ = [("pos1", C.c_uint64), ("pos2", C.c_uint64), ("value", C.c_double), ("base1", C.c_uint8),
                ("base2", C.c_uint8), ("_pad", C.c_uint8 * 6)]


class vl_fitness(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("table", C.POINTER(C.c_double)),
                ("epi", C.POINTER(vl_epi_entry)), ("n_epi", C.c_uint64), ("utility", vl_utility)]


class vl_host_spec(C.Structure):
    _fields_ = [("lo", C.c_uint64), ("hi", C.c_uint64), ("reproductive", vl_fitness), ("infective", vl_fitness)]


class vl_parameters(C.Structure):
    _fields_ = [("mutation_rate", C.c_double), ("recombination_rate", C.c_double),
                ("host_population_size", C.c_uint64), ("basic_reproductive_number", C.c_double),
                ("max_population", C.c_uint64), ("dilution", C.c_double), ("substitution_matrix", C.c_double * 16),
                ("n_hits", C.c_uint64)]


class vl_config(C.Structure):
    _fields_ = [("n_sites", C.c_uint64), ("wildtype", C.POINTER(C.c_uint8)), ("parameters", vl_parameters),
                ("n_specs", C.c_uint32), ("compartment_id", C.c_uint32), ("specs", C.POINTER(vl_host_spec)),
                ("seed", C.c_uint64), ("device", C.c_int32), ("flags", C.c_int32),
                ("capacity_infectants", C.c_uint64), ("capacity_haplotypes", C.c_uint64),
                ("capacity_arena_words", C.c_uint64)]


class vl_migrant_image(C.Structure):
    _fields_ = [("len_device", C.c_void_p), ("raw_device", C.c_void_p), ("entries_device", C.c_void_p),
                ("n_migrants", C.c_uint64), ("n_words", C.c_uint64), ("n_providers", C.c_uint64),
                ("cap_migrants", C.c_uint64), ("cap_words", C.c_uint64)]


# ----------------------------------------------------------------------------- python-side description

@dataclass
class UtilityFunction:
    """UtilityFunction (src/core/fitness/utility.rs:5-23)."""
    kind: str = "Linear"  # Linear | Algebraic | Hill
    upper: float = 0.0
    p: float = 0.0
    n: float = 0.0

    def to_c(self) -> vl_utility:
        k = {"Linear": UTILITY_LINEAR, "Algebraic": UTILITY_ALGEBRAIC, "Hill": UTILITY_HILL}[self.kind]
        return vl_utility(k, 0, self.upper, self.p, self.n)

    def apply(self, f):
        """Host-side restatement used only for reporting (utility.rs:48-60)."""
        if self.kind == "Algebraic":
            factor = 1.0 / (self.upper - 1.0)
            return self.upper * factor * f / (1.0 + factor * f)
        if self.kind == "Hill":
            k = (1.0 - self.p) / self.p
            fp = f ** self.n
            return fp / (k + fp)
        return f


@dataclass
class FitnessProvider:
    """FitnessProvider (src/providers/fitness.rs:18-22): kind Neutral | NonEpistatic | SimpleEpistatic."""
    kind: str = "Neutral"
    table: Optional[np.ndarray] = None  # (L,4) f64
    epistasis: Optional[np.ndarray] = None  # structured (pos1,base1,pos2,base2,value) or (n,5) array
    utility: UtilityFunction = field(default_factory=UtilityFunction)


@dataclass
class HostSpec:
    """HostSpec (src/core/hosts.rs:75-78) with the BasicHost's providers (src/simulation.rs:22-32)."""
    lo: int
    hi: int
    reproductive: Optional[FitnessProvider] = None
    infective: Optional[FitnessProvider] = None


@dataclass
class Parameters:
    """Parameters (src/config/parameters.rs:14-47) minus host_model (resolved into HostSpecs)."""
    mutation_rate: float
    recombination_rate: float
    host_population_size: int
    basic_reproductive_number: float
    max_population: int
    dilution: float
    substitution_matrix: Sequence[Sequence[float]] = ((0, 1, 1, 1), (1, 0, 1, 1), (1, 1, 0, 1), (1, 1, 1, 0))
    n_hits: int = 1

    def to_c(self) -> vl_parameters:
        sm = (C.c_double * 16)(*[float(x) for row in self.substitution_matrix for x in row])
        return vl_parameters(self.mutation_rate, self.recombination_rate, int(self.host_population_size),
                             self.basic_reproductive_number, int(self.max_population), self.dilution, sm,
                             int(self.n_hits))


def _epi_to_c(epi) -> tuple:
    if epi is None:
        return None, 0
    epi = np.asarray(epi)
    if epi.dtype.names:
        rows = [(int(e["pos1"]), int(e["base1"]), int(e["pos2"]), int(e["base2"]), float(e["value"])) for e in epi]
    else:
        rows = [(int(e[0]), int(e[1]), int(e[2]), int(e[3]), float(e[4])) for e in epi]
    arr = (vl_epi_entry * max(1, len(rows)))()
    for k, (p1, b1, p2, b2, v) in enumerate(rows):
        arr[k].pos1, arr[k].base1, arr[k].pos2, arr[k].base2, arr[k].value = p1, b1, p2, b2, v
    return arr, len(rows)


class MarshalledConfig:
    """Owns every buffer a vl_config points to (keeps them alive for the duration of the call)."""

    def __init__(self, wildtype: np.ndarray, parameters: Parameters, specs: Sequence[HostSpec], seed: int = 0,
                 device: int = 0, compartment_id: int = 0, flags: int = 0, capacity_infectants: int = 0,
                 capacity_haplotypes: int = 0, capacity_arena_words: int = 0):
        self._keep = []
        wt = np.ascontiguousarray(wildtype, dtype=np.uint8)
        if wt.ndim != 1 or (wt.size and wt.max() > 3):
            raise ValueError("wildtype must be a 1-d array of base indices 0..3 (A,T,C,G)")
        self._keep.append(wt)
        L = wt.size
        cspecs = (vl_host_spec * max(1, len(specs)))()
        for i, sp in enumerate(specs):
            cspecs[i].lo, cspecs[i].hi = int(sp.lo), int(sp.hi)
            for name in ("reproductive", "infective"):
                prov: Optional[FitnessProvider] = getattr(sp, name)
                cf = getattr(cspecs[i], name)
                if prov is None:
                    cf.kind = FITNESS_NONE
                    continue
                cf.utility = prov.utility.to_c()
                if prov.kind == "Neutral":
                    cf.kind = FITNESS_NEUTRAL
                    continue
                tab = np.ascontiguousarray(prov.table, dtype=np.float64)
                if tab.size != L * 4:
                    # FitnessTable::from_model size check (src/core/fitness/table.rs:38-44)
                    raise ValueError(f"Fitness table has wrong size: {tab.size} instead of {L * 4}")
                self._keep.append(tab)
                cf.table = tab.ctypes.data_as(C.POINTER(C.c_double))
                if prov.kind == "SimpleEpistatic":
                    cf.kind = FITNESS_EPISTATIC
                    arr, n = _epi_to_c(prov.epistasis)
                    self._keep.append(arr)
                    cf.epi = C.cast(arr, C.POINTER(vl_epi_entry))
                    cf.n_epi = n
                else:
                    cf.kind = FITNESS_NONEPISTATIC
        self._keep.append(cspecs)
        self.cfg = vl_config(L, wt.ctypes.data_as(C.POINTER(C.c_uint8)), parameters.to_c(), len(specs),
                             compartment_id, C.cast(cspecs, C.POINTER(vl_host_spec)), seed, device, flags,
                             capacity_infectants, capacity_haplotypes, capacity_arena_words)

    def ref(self):
        return C.byref(self.cfg)
