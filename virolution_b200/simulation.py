"""Host-side mirror of the reference's `trait Simulation<S>` (src/simulation.rs:177-219) over the C ABI.

`GpuSimulation` has the trait's methods with the trait's names and argument meaning; every one of them is a
thin ctypes call into libvirolution_gpu.so (include/virolution_gpu.h).  There is no CPU path here: if the
library is missing, or no device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from . import abi
from .abi import HostSpec, MarshalledConfig, Parameters, vl_config, vl_migrant_image, vl_parameters

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvirolution_gpu.so")
_LIB = None

u64p, u32p, u8p, i64p, f64p = (C.POINTER(t) for t in (C.c_uint64, C.c_uint32, C.c_uint8, C.c_int64, C.c_double))
_vp = C.c_void_p

# every symbol include/virolution_gpu.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "vl_create": (C.c_int, [C.POINTER(vl_config), C.POINTER(_vp)]),
    "vl_destroy": (C.c_int, [_vp]),
    "vl_last_error": (C.c_char_p, [_vp]),
    "vl_abi_version": (C.c_int, []),
    "vl_increment_generation": (C.c_int, [_vp]),
    "vl_get_generation": (C.c_int, [_vp, u64p]),
    "vl_set_generation": (C.c_int, [_vp, C.c_uint64]),
    "vl_set_parameters": (C.c_int, [_vp, C.POINTER(vl_parameters)]),
    "vl_population_len": (C.c_int, [_vp, u64p]),
    "vl_set_population_wildtype": (C.c_int, [_vp, C.c_uint64]),
    "vl_infect": (C.c_int, [_vp]),
    "vl_mutate_infectants": (C.c_int, [_vp]),
    "vl_replicate_infectants": (C.c_int, [_vp, u64p, C.c_uint64]),
    "vl_target_size": (C.c_int, [_vp, u64p, C.c_uint64, f64p]),
    "vl_sample_indices": (C.c_int, [_vp, u64p, C.c_uint64, C.c_uint64, u64p]),
    "vl_subsample_population": (C.c_int, [_vp, u64p, C.c_uint64, C.c_double, C.POINTER(_vp)]),
    "vl_select": (C.c_int, [_vp, u64p, C.c_uint64, C.POINTER(_vp)]),
    "vl_set_population": (C.c_int, [_vp, C.POINTER(_vp), C.c_uint32]),
    "vl_pop_len": (C.c_int, [_vp, u64p]),
    "vl_pop_free": (C.c_int, [_vp]),
    "vl_next_generation": (C.c_int, [_vp]),
    "vl_run": (C.c_int, [_vp, C.c_uint64]),
    "vl_advance_to_offspring": (C.c_int, [_vp]),
    "vl_step": (C.c_int, [C.POINTER(_vp), C.c_uint32, f64p, C.c_int]),
    "vl_synchronize": (C.c_int, [_vp]),
    "vl_infectant_generations": (C.c_int, [_vp, u64p, C.c_int]),
    "vl_replay_infections": (C.c_int, [_vp, i64p, C.c_uint64]),
    "vl_replay_mutations": (C.c_int, [_vp, u64p, C.c_uint64, u32p, u8p]),
    "vl_replay_recombinations": (C.c_int, [_vp, C.c_uint64, u64p, u32p, u32p, u32p, u32p, u8p]),
    "vl_replay_offspring": (C.c_int, [_vp, u64p, C.c_uint64]),
    "vl_get_infections": (C.c_int, [_vp, i64p, C.c_uint64]),
    "vl_get_host_map": (C.c_int, [_vp, u64p, u64p, u64p]),
    "vl_get_replication_terms": (C.c_int, [_vp, f64p, f64p, f64p, C.c_uint64]),
    "vl_get_offspring_prefix": (C.c_int, [_vp, u64p, C.c_uint64]),
    "vl_get_raw_fitness": (C.c_int, [_vp, C.c_uint32, C.c_int, f64p, C.c_uint64]),
    "vl_export_mutations": (C.c_int, [_vp, u64p, C.c_uint64, u32p, u8p, u64p]),
    "vl_get_base": (C.c_int, [_vp, C.c_uint64, C.c_uint64, u8p]),
    "vl_stats_frequencies": (C.c_int, [_vp, f64p, C.c_uint64]),
    "vl_stats_mean_fitness": (C.c_int, [_vp, C.c_uint32, C.c_int, f64p]),
    "vl_get_haplotype_ids": (C.c_int, [_vp, u32p, C.c_uint64]),
    "vl_get_counters": (C.c_int, [_vp, u64p]),
    "vl_get_capacities": (C.c_int, [_vp, u64p]),
    "vl_collect_garbage": (C.c_int, [_vp]),
    "vl_pop_pack_device": (C.c_int, [_vp, _vp, C.POINTER(_vp), u64p]),
    "vl_pop_unpack_device": (C.c_int, [_vp, _vp, C.c_uint64, C.POINTER(_vp)]),
    "vl_migrants_export_device": (C.c_int, [_vp, C.POINTER(_vp), C.c_uint32, C.POINTER(vl_migrant_image), u64p, u64p]),
    "vl_migrants_import_device": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint64, C.c_uint64, C.POINTER(_vp)]),
    "vl_subsample_populations": (C.c_int, [_vp, f64p, C.c_uint32, C.POINTER(_vp)]),
    "vl_exchange_create": (C.c_int, [_vp, C.c_uint32, C.c_uint32, f64p, C.c_uint64, C.c_int, C.c_uint64, C.POINTER(_vp), _vp]),
    "vl_exchange_begin": (C.c_int, [_vp]),
    "vl_exchange_recv_buffer": (C.c_int, [_vp, C.POINTER(_vp), u64p]),
    "vl_exchange_connect": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "vl_exchange_step": (C.c_int, [_vp]),
    "vl_exchange_run": (C.c_int, [_vp, C.c_uint64]),
    "vl_exchange_set_transfer": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "vl_exchange_push": (C.c_int, [_vp]),
    "vl_exchange_pull": (C.c_int, [_vp]),
    "vl_exchange_destroy": (C.c_int, [_vp]),
    "vl_timer_start": (C.c_int, [_vp]),
    "vl_timer_stop": (C.c_int, [_vp, f64p]),
    "vl_profile": (C.c_int, [_vp, C.c_int]),
    "vl_profile_report": (C.c_int, [_vp, C.c_char_p, C.c_uint64]),
    "vl_debug_stamps": (C.c_int, [_vp, u64p]),
    "vl_debug_last_generation": (C.c_int, [_vp, u64p, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_uint64]),
    "vl_test_poisson": (C.c_int, [C.c_uint64, C.c_double, C.c_uint64, u64p]),
    "vl_test_poisson_fast": (C.c_int, [C.c_uint64, C.c_double, C.c_uint64, u64p]),
    "vl_test_binomial": (C.c_int, [C.c_uint64, C.c_uint64, C.c_double, C.c_uint64, u64p]),
    "vl_test_below": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, u64p]),
}


def load_library(path: Optional[str] = None):
    """dlopen the C-ABI library and bind every declared symbol.  Raises if the library is absent: the product
    path has no fallback."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(f"{p} is missing: build it with `python -m virolution_b200.build` (nvcc, sm_100a). "
                           "virolution_b200 has no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = header/library mismatch
        fn.restype, fn.argtypes = res, args
    if path is None:
        _LIB = lib
    return lib


class VirolutionError(RuntimeError):
    """Non-zero vl_status.  `code` is the status, the message comes from vl_last_error."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code


class ReferencePanic(VirolutionError):
    """The reference panics at this point (VL_ERR_REFERENCE_PANIC)."""


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


class GpuPopulation:
    """Detached `Population<Store<S>>` (src/core/population.rs:156-163) living in device memory of its origin."""

    def __init__(self, ptr, origin: "GpuSimulation"):
        self.ptr = ptr
        self.origin = origin

    def __len__(self) -> int:
        n = C.c_uint64()
        load_library().vl_pop_len(self.ptr, C.byref(n))
        return int(n.value)

    def free(self):
        if self.ptr:
            load_library().vl_pop_free(self.ptr)
            self.ptr = None

    def pack_device(self):
        """(device pointer, bytes) of the self-contained image that travels between GPUs."""
        img, nbytes = _vp(), C.c_uint64()
        self.origin._check(load_library().vl_pop_pack_device(self.origin.ptr, self.ptr, C.byref(img), C.byref(nbytes)))
        return int(img.value), int(nbytes.value)

    def __del__(self):
        try:
            if self.ptr and self.origin.ptr:
                self.free()
        except Exception:
            pass


class GpuSimulation:
    """`impl Simulation<Nt>` on one B200 (one compartment).  Constructor = BasicSimulation::new
    (src/simulation.rs:235-256) with the resolved configuration."""

    def __init__(self, wildtype: np.ndarray, parameters: Parameters, host_specs: Sequence[HostSpec], seed: int = 0,
                 device: int = 0, compartment_id: int = 0, flags: int = 0, capacity_infectants: int = 0,
                 capacity_haplotypes: int = 0, capacity_arena_words: int = 0):
        self.lib = load_library()
        self._cfg = MarshalledConfig(wildtype, parameters, host_specs, seed=seed, device=device,
                                     compartment_id=compartment_id, flags=flags,
                                     capacity_infectants=capacity_infectants, capacity_haplotypes=capacity_haplotypes,
                                     capacity_arena_words=capacity_arena_words)
        self.parameters = parameters
        self.host_specs = list(host_specs)
        self.n_sites = int(len(wildtype))
        self.device = device
        ptr = _vp()
        rc = self.lib.vl_create(self._cfg.ref(), C.byref(ptr))
        if rc != 0:
            raise VirolutionError(rc, (self.lib.vl_last_error(None) or b"").decode())
        self.ptr = ptr

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int):
        if rc != 0:
            msg = (self.lib.vl_last_error(self.ptr) or b"").decode()
            raise (ReferencePanic if rc == abi.VL_ERR_REFERENCE_PANIC else VirolutionError)(rc, msg)

    def close(self):
        if getattr(self, "ptr", None):
            self.lib.vl_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _offspring_args(self, offspring):
        if offspring is None:
            return None, 0, None
        off = np.ascontiguousarray(offspring, dtype=np.uint64)
        return _p(off, C.c_uint64), off.size, off

    # ------------------------------------------------------------------ Simulation trait
    def increment_generation(self):
        self._check(self.lib.vl_increment_generation(self.ptr))

    def get_generation(self) -> int:
        g = C.c_uint64()
        self._check(self.lib.vl_get_generation(self.ptr, C.byref(g)))
        return int(g.value)

    def set_generation(self, generation: int):
        self._check(self.lib.vl_set_generation(self.ptr, generation))

    def set_parameters(self, parameters: Parameters):
        self._check(self.lib.vl_set_parameters(self.ptr, C.byref(parameters.to_c())))
        self.parameters = parameters

    def population_len(self) -> int:
        n = C.c_uint64()
        self._check(self.lib.vl_population_len(self.ptr, C.byref(n)))
        return int(n.value)

    def set_population_wildtype(self, n: int):
        self._check(self.lib.vl_set_population_wildtype(self.ptr, n))

    def get_host_specs(self) -> List[HostSpec]:
        return self.host_specs

    def infect(self):
        self._check(self.lib.vl_infect(self.ptr))

    def mutate_infectants(self):
        self._check(self.lib.vl_mutate_infectants(self.ptr))

    def replicate_infectants(self, out: Optional[np.ndarray] = None, to_host: bool = True):
        """Returns the offspring map (`Vec<usize>`) unless to_host=False (kept on the device)."""
        if not to_host:
            self._check(self.lib.vl_replicate_infectants(self.ptr, None, 0))
            return None
        if out is None:
            out = np.zeros(self.population_len(), dtype=np.uint64)
        self._check(self.lib.vl_replicate_infectants(self.ptr, _p(out, C.c_uint64), out.size))
        return out

    def target_size(self, offspring=None) -> float:
        ptr, n, _keep = self._offspring_args(offspring)
        out = C.c_double()
        self._check(self.lib.vl_target_size(self.ptr, ptr, n, C.byref(out)))
        return float(out.value)

    def sample_indices(self, offspring, amount: int) -> np.ndarray:
        ptr, n, _keep = self._offspring_args(offspring)
        if offspring is not None and n == 0:
            return np.zeros(0, dtype=np.uint64)
        out = np.zeros(max(1, amount), dtype=np.uint64)
        self._check(self.lib.vl_sample_indices(self.ptr, ptr, n, amount, _p(out, C.c_uint64)))
        return out[:amount]

    def subsample_population(self, offspring, factor: float) -> GpuPopulation:
        ptr, n, _keep = self._offspring_args(offspring)
        pop = _vp()
        self._check(self.lib.vl_subsample_population(self.ptr, ptr, n, factor, C.byref(pop)))
        return GpuPopulation(pop, self)

    def subsample_populations(self, factors) -> List[GpuPopulation]:
        """[subsample_population(&offspring, f) for f in factors] on the device-resident offspring map, one sync."""
        f = np.ascontiguousarray(factors, dtype=np.float64)
        arr = (_vp * max(1, f.size))()
        self._check(self.lib.vl_subsample_populations(self.ptr, _p(f, C.c_double), f.size, arr))
        return [GpuPopulation(_vp(arr[k]), self) for k in range(f.size)]

    def select(self, indices) -> GpuPopulation:
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        pop = _vp()
        self._check(self.lib.vl_select(self.ptr, _p(idx, C.c_uint64), idx.size, C.byref(pop)))
        return GpuPopulation(pop, self)

    def set_population(self, parts):
        """set_population(Population::from_iter(parts)); consumes the parts."""
        if isinstance(parts, GpuPopulation):
            parts = [parts]
        arr = (_vp * max(1, len(parts)))(*[p.ptr for p in parts])
        rc = self.lib.vl_set_population(self.ptr, arr, len(parts))
        for p in parts:
            p.ptr = None
        self._check(rc)

    def next_generation(self):
        self._check(self.lib.vl_next_generation(self.ptr))

    # ------------------------------------------------------------------ fused loops
    def advance_to_offspring(self):
        """increment_generation + infect + mutate_infectants + replicate_infectants, offspring map kept on the device."""
        self._check(self.lib.vl_advance_to_offspring(self.ptr))

    def run(self, n_generations: int):
        self._check(self.lib.vl_run(self.ptr, n_generations))

    def synchronize(self):
        self._check(self.lib.vl_synchronize(self.ptr))

    def timer_start(self):
        self._check(self.lib.vl_timer_start(self.ptr))

    def timer_stop(self) -> float:
        """Device milliseconds since timer_start (CUDA events on this handle's stream); synchronises."""
        ms = C.c_double()
        self._check(self.lib.vl_timer_stop(self.ptr, C.byref(ms)))
        return float(ms.value)

    def profile(self, enable: bool):
        self._check(self.lib.vl_profile(self.ptr, int(enable)))

    def profile_report(self) -> dict:
        """{kernel: (launches, total_ms)} of the launches since profile(True)."""
        buf = C.create_string_buffer(1 << 16)
        self._check(self.lib.vl_profile_report(self.ptr, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.rsplit(" ", 2)
            out[name] = (int(cnt), float(ms))
        return out

    def debug_stamps(self) -> np.ndarray:
        out = np.zeros(8, dtype=np.uint64)
        self._check(self.lib.vl_debug_stamps(self.ptr, _p(out, C.c_uint64)))
        return out

    def infectant_generations(self, reset: bool = False) -> int:
        out = C.c_uint64()
        self._check(self.lib.vl_infectant_generations(self.ptr, C.byref(out), int(reset)))
        return int(out.value)

    # ------------------------------------------------------------------ replay hooks
    def replay_infections(self, infections):
        inf = np.ascontiguousarray(infections, dtype=np.int64)
        self._check(self.lib.vl_replay_infections(self.ptr, _p(inf, C.c_int64), inf.size))

    def replay_mutations(self, offsets, sites, bases):
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        s = np.ascontiguousarray(sites, dtype=np.uint32)
        b = np.ascontiguousarray(bases, dtype=np.uint8)
        self._check(self.lib.vl_replay_mutations(self.ptr, _p(off, C.c_uint64), off.size - 1, _p(s, C.c_uint32), _p(b, C.c_uint8)))

    def replay_recombinations(self, host, i, j, s0, s1, which):
        host = np.ascontiguousarray(host, dtype=np.uint64)
        i, j, s0, s1 = (np.ascontiguousarray(a, dtype=np.uint32) for a in (i, j, s0, s1))
        which = np.ascontiguousarray(which, dtype=np.uint8)
        self._check(self.lib.vl_replay_recombinations(self.ptr, host.size, _p(host, C.c_uint64), _p(i, C.c_uint32),
                                                      _p(j, C.c_uint32), _p(s0, C.c_uint32), _p(s1, C.c_uint32),
                                                      _p(which, C.c_uint8)))

    def replay_offspring(self, offspring):
        off = np.ascontiguousarray(offspring, dtype=np.uint64)
        self._check(self.lib.vl_replay_offspring(self.ptr, _p(off, C.c_uint64), off.size))

    # ------------------------------------------------------------------ inspection / export
    def get_infections(self) -> np.ndarray:
        out = np.zeros(self.population_len(), dtype=np.int64)
        self._check(self.lib.vl_get_infections(self.ptr, _p(out, C.c_int64), out.size))
        return out

    def get_host_map(self):
        H = int(self.parameters.host_population_size)
        offsets = np.zeros(H + 1, dtype=np.uint64)
        hosts = np.zeros(max(1, self.population_len()), dtype=np.uint64)
        n = C.c_uint64()
        self._check(self.lib.vl_get_host_map(self.ptr, _p(offsets, C.c_uint64), _p(hosts, C.c_uint64), C.byref(n)))
        return offsets, hosts[:int(n.value)]

    def get_replication_terms(self):
        n = self.population_len()
        f, s, l = (np.zeros(n, dtype=np.float64) for _ in range(3))
        self._check(self.lib.vl_get_replication_terms(self.ptr, _p(f, C.c_double), _p(s, C.c_double), _p(l, C.c_double), n))
        return f, s, l

    def get_offspring_prefix(self) -> np.ndarray:
        out = np.zeros(self.population_len(), dtype=np.uint64)
        self._check(self.lib.vl_get_offspring_prefix(self.ptr, _p(out, C.c_uint64), out.size))
        return out

    def get_raw_fitness(self, spec: int = 0, which: int = 0) -> np.ndarray:
        out = np.zeros(self.population_len(), dtype=np.float64)
        self._check(self.lib.vl_get_raw_fitness(self.ptr, spec, which, _p(out, C.c_double), out.size))
        return out

    def export_mutations(self):
        """(offsets[n+1], positions, bases): Haplotype::get_mutations of every infectant, sorted by position."""
        n = self.population_len()
        total = C.c_uint64()
        offsets = np.zeros(n + 1, dtype=np.uint64)
        self._check(self.lib.vl_export_mutations(self.ptr, _p(offsets, C.c_uint64), n, None, None, C.byref(total)))
        t = int(total.value)
        pos = np.zeros(max(1, t), dtype=np.uint32)
        base = np.zeros(max(1, t), dtype=np.uint8)
        if t:
            self._check(self.lib.vl_export_mutations(self.ptr, _p(offsets, C.c_uint64), n, _p(pos, C.c_uint32),
                                                     _p(base, C.c_uint8), C.byref(total)))
        return offsets, pos[:t], base[:t]

    # ------------------------------------------------------------------ statistics and final.<i>.csv
    def frequencies(self) -> np.ndarray:
        """PopulationFrequencies::frequencies (src/stats/population.rs:15-40): (n_sites, 4) relative frequencies."""
        out = np.zeros((self.n_sites, 4), dtype=np.float64)
        self._check(self.lib.vl_stats_frequencies(self.ptr, _p(out, C.c_double), self.n_sites))
        return out

    def distance(self, other: "GpuSimulation") -> float:
        """PopulationDistance::distance (:66-75): sum of absolute per-site frequency differences, summed in index order."""
        d = np.abs(self.frequencies().ravel() - other.frequencies().ravel())
        return float(np.cumsum(d)[-1]) if d.size else 0.0

    def mean_fitness(self, spec: int = 0, which: int = 0) -> float:
        """PopulationStatistics::mean of the mapped fitness (:50-56)."""
        out = C.c_double()
        self._check(self.lib.vl_stats_mean_fitness(self.ptr, spec, which, C.byref(out)))
        return float(out.value)

    def haplotype_rows(self, wildtype: np.ndarray):
        """Rows of final.<i>.csv (src/runner.rs:137-152): one (Haplotype::get_string, count) per haplotype *record*
        (the reference counts by reference identity), in first-occurrence order."""
        letters = "ATCG"
        ids = self.get_haplotype_ids()
        off, pos, base = self.export_mutations()
        uniq, first, counts = np.unique(ids, return_index=True, return_counts=True)
        rows = []
        for k in np.argsort(first):
            i = int(first[k])
            a, b = int(off[i]), int(off[i + 1])
            text = ";".join(f"{int(p)}:{letters[int(wildtype[int(p)])]}->{letters[int(t)]}" for p, t in zip(pos[a:b], base[a:b]))
            rows.append((text or "wt", int(counts[k])))
        return rows

    def write_final_csv(self, path: str, wildtype: np.ndarray):
        import csv
        with open(path, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["haplotype", "count"])
            w.writerows(self.haplotype_rows(wildtype))

    def get_base(self, infectant: int, position: int) -> int:
        out = C.c_uint8()
        self._check(self.lib.vl_get_base(self.ptr, infectant, position, C.byref(out)))
        return int(out.value)

    def get_haplotype_ids(self) -> np.ndarray:
        out = np.zeros(self.population_len(), dtype=np.uint32)
        self._check(self.lib.vl_get_haplotype_ids(self.ptr, _p(out, C.c_uint32), out.size))
        return out

    def debug_last_generation(self, n: int):
        """(offspring, hosts, host_sums, fitness) of the last device-resident generation, in the order of the population
        it started with (diagnostic; tests)."""
        off = np.zeros(n, dtype=np.uint64)
        hosts = np.zeros(n, dtype=np.int64)
        sums = np.zeros(n, dtype=np.float64)
        fit = np.zeros(n, dtype=np.float64)
        self._check(self.lib.vl_debug_last_generation(self.ptr, _p(off, C.c_uint64), _p(hosts, C.c_int64), _p(sums, C.c_double),
                                                      _p(fit, C.c_double), n))
        return off, hosts, sums, fit

    def get_counters(self) -> dict:
        out = np.zeros(4, dtype=np.uint64)
        self._check(self.lib.vl_get_counters(self.ptr, _p(out, C.c_uint64)))
        return {"haplotype_records": int(out[0]), "arena_words": int(out[1]), "gc_runs": int(out[2]),
                "kernels_launched": int(out[3])}

    def get_capacities(self):
        out = np.zeros(4, dtype=np.uint64)
        self._check(self.lib.vl_get_capacities(self.ptr, _p(out, C.c_uint64)))
        return {"infectants": int(out[0]), "haplotype_records": int(out[1]), "arena_words": int(out[2]), "gc_check_period": int(out[3])}

    def collect_garbage(self):
        self._check(self.lib.vl_collect_garbage(self.ptr))

    def export_migrants(self, parts: Sequence[GpuPopulation]):
        """Flat image of the parts (one contiguous range per part): -> (vl_migrant_image, migrants_per_part, words_per_part).
        The image's device pointers stay valid until the next export on this handle."""
        k = len(parts)
        arr = (_vp * max(1, k))(*[p.ptr for p in parts])
        img = vl_migrant_image()
        m, w = np.zeros(max(1, k), dtype=np.uint64), np.zeros(max(1, k), dtype=np.uint64)
        self._check(self.lib.vl_migrants_export_device(self.ptr, arr, k, C.byref(img), _p(m, C.c_uint64), _p(w, C.c_uint64)))
        return img, m[:k], w[:k]

    def import_migrants(self, len_ptr: int, raw_ptr: int, entries_ptr: int, n_migrants: int, n_words: int) -> GpuPopulation:
        """Detached population from a received image range (device pointers on this handle's GPU, already complete)."""
        pop = _vp()
        self._check(self.lib.vl_migrants_import_device(self.ptr, _vp(len_ptr), _vp(raw_ptr), _vp(entries_ptr), n_migrants, n_words,
                                                       C.byref(pop)))
        return GpuPopulation(pop, self)

    def unpack_device(self, image_ptr: int, nbytes: int) -> GpuPopulation:
        pop = _vp()
        self._check(self.lib.vl_pop_unpack_device(self.ptr, _vp(image_ptr), nbytes, C.byref(pop)))
        return GpuPopulation(pop, self)


class PeerExchangeHandle:
    """vl_exchange of one rank: receive buffer + peer mappings + the per-generation enqueue (include/virolution_gpu.h)."""
    IPC_BYTES = 64

    def __init__(self, sim: GpuSimulation, rank: int, world: int, transfer=None, words_per_migrant: int = 0,
                 sharded: bool = False, shard_seed: int = 0):
        self.sim, self.rank, self.world = sim, rank, world
        T = None
        if transfer is not None:
            T = np.ascontiguousarray(transfer, dtype=np.float64)
            if T.shape != (world, world):
                raise ValueError(f"transfer matrix is {T.shape}, expected {(world, world)}")
        elif not sharded:
            raise ValueError("a transfer matrix is required unless the compartment is sharded")
        self.ipc_handle = np.zeros(self.IPC_BYTES, dtype=np.uint8)
        ptr = _vp()
        sim._check(sim.lib.vl_exchange_create(sim.ptr, rank, world, None if T is None else _p(T, C.c_double), words_per_migrant,
                                              int(sharded), shard_seed, C.byref(ptr), self.ipc_handle.ctypes.data_as(_vp)))
        self.ptr = ptr

    def recv_buffer(self) -> int:
        out, nbytes = _vp(), C.c_uint64()
        self.sim._check(self.sim.lib.vl_exchange_recv_buffer(self.ptr, C.byref(out), C.byref(nbytes)))
        return int(out.value)

    def connect(self, ipc_handles: Optional[np.ndarray] = None, local_ptrs: Optional[Sequence[int]] = None):
        """ipc_handles: (world, 64) uint8 gathered from all ranks; local_ptrs: receive buffers of ranks in this process."""
        h = None if ipc_handles is None else np.ascontiguousarray(ipc_handles, dtype=np.uint8)
        lp = None
        if local_ptrs is not None:
            lp = (_vp * self.world)(*[_vp(p) if p else _vp() for p in local_ptrs])
        self.sim._check(self.sim.lib.vl_exchange_connect(self.ptr, None if h is None else h.ctypes.data_as(_vp), lp))

    def step(self):
        self.sim._check(self.sim.lib.vl_exchange_step(self.ptr))

    def run(self, n_generations: int):
        self.sim._check(self.sim.lib.vl_exchange_run(self.ptr, int(n_generations)))

    def set_transfer(self, T):
        """The matrix of the generations enqueued from now on; entries must not exceed those the handle was created with."""
        T = np.ascontiguousarray(T, dtype=np.float64)
        if T.shape != (self.world, self.world):
            raise ValueError(f"transfer matrix must be {self.world} x {self.world}")
        self.sim._check(self.sim.lib.vl_exchange_set_transfer(self.ptr, _p(T, C.c_double)))

    def begin(self):
        self.sim._check(self.sim.lib.vl_exchange_begin(self.ptr))

    def push(self):
        self.sim._check(self.sim.lib.vl_exchange_push(self.ptr))

    def pull(self):
        self.sim._check(self.sim.lib.vl_exchange_pull(self.ptr))

    def close(self):
        if getattr(self, "ptr", None) and self.sim.ptr:
            self.sim.lib.vl_exchange_destroy(self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def step(sims: Sequence[GpuSimulation], transfer: Optional[np.ndarray] = None, rule: int = 0):
    """One generation over the compartments of this process as Runner::run drives it
    (rule 0: parallel build src/runner.rs:382-422; rule 1: serial build :536-631)."""
    lib = load_library()
    C_ = len(sims)
    T = np.ascontiguousarray(np.eye(C_) if transfer is None else np.asarray(transfer)[:C_, :C_], dtype=np.float64)
    arr = (_vp * C_)(*[s.ptr for s in sims])
    rc = lib.vl_step(arr, C_, _p(T, C.c_double), rule)
    if rc != 0:
        msg = (lib.vl_last_error(sims[0].ptr) or b"").decode()
        raise (ReferencePanic if rc == abi.VL_ERR_REFERENCE_PANIC else VirolutionError)(rc, msg)


class CompartmentGroup:
    """The compartments of ONE process (any devices with peer access; normally one GPU) stepped through the peer exchange
    instead of `vl_step`: every compartment writes its migrants into the others' receive buffers inside the kernels, sizes
    stay on the device, and the host only enqueues — no detached populations, no synchronisation per generation.
    Same transmission block (parallel build, src/runner.rs:382-422) and the same populations as `step(sims, T, 0)`
    (tests/test_gpu_migrants.py).  `transfer_cap` is the entry-wise maximum of every matrix the schedule will ask for."""

    CHUNK = 8  # generations enqueued per compartment before the next one gets its turn (bounded by the launch queue)

    def __init__(self, sims: Sequence[GpuSimulation], transfer_cap, words_per_migrant: int = 0):
        self.sims = list(sims)
        C_ = len(self.sims)
        if C_ < 2:
            raise ValueError("a group has at least two compartments")
        self.cap = np.ascontiguousarray(np.asarray(transfer_cap, dtype=np.float64)[:C_, :C_])
        self.T = self.cap.copy()
        self.ex = [PeerExchangeHandle(self.sims[k], k, C_, self.cap, words_per_migrant) for k in range(C_)]
        bufs = [e.recv_buffer() for e in self.ex]
        for k, e in enumerate(self.ex):
            e.connect(None, [bufs[t] if t != k else 0 for t in range(C_)])

    def set_transfer(self, T):
        T = np.ascontiguousarray(np.asarray(T, dtype=np.float64)[:len(self.sims), :len(self.sims)])
        if not np.array_equal(T, self.T):
            for e in self.ex:
                e.set_transfer(T)
            self.T = T

    def step(self, transfer=None):
        """One generation of every compartment.  Compartments of one process push all before they pull any."""
        if transfer is not None:
            self.set_transfer(transfer)
        for e in self.ex:
            e.push()
        for e in self.ex:
            e.pull()

    def run(self, n_generations: int):
        """n generations; nothing is read back and nothing is waited for on the host."""
        left = int(n_generations)
        while left > 0:
            k = min(left, self.CHUNK)
            for e in self.ex:
                e.run(k)
            left -= k

    def synchronize(self):
        for s in self.sims:
            s.synchronize()

    def close(self):
        for s in self.sims:
            if s.ptr:
                try:
                    s.synchronize()
                except VirolutionError:
                    pass  # (a failed handle stays failed: it is being torn down)
        for e in self.ex:
            e.close()
        self.ex = []


def sampler_poisson(seed: int, lam: float, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint64)
    rc = load_library().vl_test_poisson(seed, lam, n, _p(out, C.c_uint64))
    if rc:
        raise VirolutionError(rc, "vl_test_poisson failed")
    return out


def sampler_poisson_fast(seed: int, lam: float, n: int):
    """(f32-guarded inversion, f64 inversion) of the same uniforms; lam <= 0 spreads the mean over (0, 12) per draw."""
    out = np.zeros(2 * n, dtype=np.uint64)
    rc = load_library().vl_test_poisson_fast(seed, lam, n, _p(out, C.c_uint64))
    if rc:
        raise VirolutionError(rc, "vl_test_poisson_fast failed")
    return out[:n], out[n:]


def sampler_binomial(seed: int, trials: int, p: float, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint64)
    rc = load_library().vl_test_binomial(seed, trials, p, n, _p(out, C.c_uint64))
    if rc:
        raise VirolutionError(rc, "vl_test_binomial failed")
    return out


def sampler_below(seed: int, bound: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint64)
    rc = load_library().vl_test_below(seed, bound, n, _p(out, C.c_uint64))
    if rc:
        raise VirolutionError(rc, "vl_test_below failed")
    return out
