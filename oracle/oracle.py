"""ctypes wrapper of the CPU oracle (oracle/vl_oracle.c).  TEST INFRASTRUCTURE ONLY — see the C header.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

from virolution_b200.abi import HostSpec, MarshalledConfig, Parameters, vl_config, vl_parameters, vl_utility

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libvl_oracle.so")
    src = os.path.join(_HERE, "vl_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp = C.c_void_p
        u64p, u32p, u8p, i64p, f64p = (C.POINTER(t) for t in (C.c_uint64, C.c_uint32, C.c_uint8, C.c_int64, C.c_double))
        sig = {
            "orc_world_create": (vp, [C.POINTER(vl_config), C.c_uint32, C.c_int]),
            "orc_world_destroy": (None, [vp]),
            "orc_last_error": (C.c_char_p, [vp]),
            "orc_set_population_wildtype": (C.c_int, [vp, C.c_uint32, C.c_uint64]),
            "orc_population_len": (C.c_uint64, [vp, C.c_uint32]),
            "orc_set_parameters": (None, [vp, C.c_uint32, C.POINTER(vl_parameters)]),
            "orc_increment_generation": (None, [vp, C.c_uint32]),
            "orc_get_generation": (C.c_uint64, [vp, C.c_uint32]),
            "orc_infectant_generations": (C.c_uint64, [vp, C.c_uint32]),
            "orc_infect": (C.c_int, [vp, C.c_uint32]),
            "orc_set_infections": (C.c_int, [vp, C.c_uint32, i64p, C.c_uint64]),
            "orc_mutate": (C.c_int, [vp, C.c_uint32]),
            "orc_replicate": (C.c_int, [vp, C.c_uint32, u64p]),
            "orc_target_size": (C.c_double, [vp, C.c_uint32, u64p, C.c_uint64]),
            "orc_sample_indices": (C.c_int, [vp, C.c_uint32, u64p, C.c_uint64, C.c_uint64, u64p]),
            "orc_subsample_size": (C.c_uint64, [vp, C.c_uint32, u64p, C.c_uint64, C.c_double]),
            "orc_select": (vp, [vp, C.c_uint32, u64p, C.c_uint64]),
            "orc_subsample": (vp, [vp, C.c_uint32, u64p, C.c_uint64, C.c_double, u64p]),
            "orc_pop_len": (C.c_uint64, [vp]),
            "orc_pop_free": (None, [vp]),
            "orc_set_population": (C.c_int, [vp, C.c_uint32, C.POINTER(vp), C.c_uint32]),
            "orc_next_generation": (C.c_int, [vp, C.c_uint32]),
            "orc_step": (C.c_int, [vp, f64p, C.c_int]),
            "orc_last_indices": (C.c_uint64, [vp, C.c_uint32, C.c_uint32, u64p]),
            "orc_get_infections": (None, [vp, C.c_uint32, i64p]),
            "orc_get_host_map": (C.c_uint64, [vp, C.c_uint32, u64p, u64p]),
            "orc_get_mutation_log": (C.c_uint64, [vp, C.c_uint32, u64p, u32p, u8p]),
            "orc_get_recombination_log": (C.c_uint64, [vp, C.c_uint32, u64p, u32p, u32p, u32p, u32p, u8p]),
            "orc_get_offspring": (None, [vp, C.c_uint32, u64p]),
            "orc_get_replication_terms": (None, [vp, C.c_uint32, f64p, f64p, f64p]),
            "orc_get_raw_fitness": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_int, f64p]),
            "orc_export_mutations": (C.c_uint64, [vp, C.c_uint32, u64p, u32p, u8p]),
            "orc_get_base": (C.c_uint8, [vp, C.c_uint32, C.c_uint64, C.c_uint32]),
            "orc_get_population_ids": (None, [vp, C.c_uint32, i64p]),
            "orc_hap_wildtype": (C.c_int64, [vp]),
            "orc_hap_create_descendant": (C.c_int64, [vp, C.c_int64, u32p, u8p, C.c_uint32]),
            "orc_hap_create_recombinant": (C.c_int64, [vp, C.c_int64, C.c_int64, C.c_uint32, C.c_uint32]),
            "orc_hap_get_base": (C.c_uint8, [vp, C.c_int64, C.c_uint32]),
            "orc_hap_get_sequence": (None, [vp, C.c_int64, u8p]),
            "orc_hap_get_mutations": (C.c_uint32, [vp, C.c_int64, u32p, u8p, C.c_uint32]),
            "orc_hap_raw_fitness": (C.c_double, [vp, C.c_int64, C.c_uint32, C.c_int]),
            "orc_hap_fitness": (C.c_double, [vp, C.c_int64, C.c_uint32, C.c_int]),
            "orc_table_get_value": (C.c_double, [vp, C.c_uint32, C.c_int, C.c_uint32, C.c_uint8]),
            "orc_utility_apply": (C.c_double, [C.POINTER(vl_utility), C.c_double]),
            "orc_set_infectant": (None, [vp, C.c_uint32, C.c_uint64, C.c_int64]),
            "orc_n_nodes": (C.c_uint64, [vp]),
            "orc_test_poisson": (None, [C.c_uint64, C.c_double, C.c_uint64, f64p]),
            "orc_test_binomial": (None, [C.c_uint64, C.c_uint64, C.c_double, C.c_uint64, u64p]),
            "orc_test_alias": (C.c_int, [C.c_uint64, u64p, C.c_uint64, C.c_uint64, u64p]),
            "orc_num_threads": (C.c_int, [vp]),
            "orc_set_mode": (None, [vp, C.c_int]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class ReferencePanic(RuntimeError):
    """The reference panics at this point (e.g. WeightedAliasIndex::new on all-zero weights)."""


class OraclePop:
    def __init__(self, ptr):
        self.ptr = ptr

    def __len__(self):
        return int(lib().orc_pop_len(self.ptr))


class OracleWorld:
    """C compartments sharing one haplotype tree (the reference's Vec<Box<dyn Simulation>>)."""

    def __init__(self, wildtype, parameters: Parameters, specs: Sequence[HostSpec], n_compartments: int = 1,
                 seed: int = 0, n_threads: int = 1):
        self._cfg = MarshalledConfig(wildtype, parameters, specs, seed=seed)
        self.L = int(len(wildtype))
        self.C = n_compartments
        self.H = int(parameters.host_population_size)
        self.ptr = lib().orc_world_create(self._cfg.ref(), n_compartments, n_threads)

    def __del__(self):
        try:
            if self.ptr:
                lib().orc_world_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass

    # --- Simulation trait ---------------------------------------------------
    def set_population_wildtype(self, n, c=0):
        lib().orc_set_population_wildtype(self.ptr, c, n)

    def population_len(self, c=0):
        return int(lib().orc_population_len(self.ptr, c))

    def set_parameters(self, p: Parameters, c=0):
        self.H = int(p.host_population_size)
        lib().orc_set_parameters(self.ptr, c, C.byref(p.to_c()))

    def increment_generation(self, c=0):
        lib().orc_increment_generation(self.ptr, c)

    def get_generation(self, c=0):
        return int(lib().orc_get_generation(self.ptr, c))

    def infect(self, c=0):
        lib().orc_infect(self.ptr, c)

    def set_infections(self, inf, c=0):
        inf = np.ascontiguousarray(inf, dtype=np.int64)
        lib().orc_set_infections(self.ptr, c, _p(inf, C.c_int64), inf.size)

    def mutate_infectants(self, c=0):
        lib().orc_mutate(self.ptr, c)

    def replicate_infectants(self, c=0):
        out = np.zeros(self.population_len(c), dtype=np.uint64)
        lib().orc_replicate(self.ptr, c, _p(out, C.c_uint64))
        return out

    def target_size(self, offspring, c=0):
        off = np.ascontiguousarray(offspring, dtype=np.uint64)
        return float(lib().orc_target_size(self.ptr, c, _p(off, C.c_uint64), off.size))

    def sample_indices(self, offspring, amount, c=0):
        off = np.ascontiguousarray(offspring, dtype=np.uint64)
        out = np.zeros(max(1, amount), dtype=np.uint64)
        if lib().orc_sample_indices(self.ptr, c, _p(off, C.c_uint64), off.size, amount, _p(out, C.c_uint64)) != 0:
            raise ReferencePanic(lib().orc_last_error(self.ptr).decode())
        return out[:amount] if off.size else out[:0]

    def subsample_size(self, offspring, factor, c=0):
        off = np.ascontiguousarray(offspring, dtype=np.uint64)
        return int(lib().orc_subsample_size(self.ptr, c, _p(off, C.c_uint64), off.size, factor))

    def subsample_population(self, offspring, factor, c=0):
        """Returns (pop, indices)."""
        off = np.ascontiguousarray(offspring, dtype=np.uint64)
        size = self.subsample_size(off, factor, c)
        idx = np.zeros(max(1, size), dtype=np.uint64)
        ptr = lib().orc_subsample(self.ptr, c, _p(off, C.c_uint64), off.size, factor, _p(idx, C.c_uint64))
        if not ptr:
            raise ReferencePanic(lib().orc_last_error(self.ptr).decode())
        return OraclePop(ptr), idx[:size]

    def select(self, indices, c=0):
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        return OraclePop(lib().orc_select(self.ptr, c, _p(idx, C.c_uint64), idx.size))

    def set_population(self, parts: Sequence[OraclePop], c=0):
        arr = (C.c_void_p * len(parts))(*[p.ptr for p in parts])
        lib().orc_set_population(self.ptr, c, arr, len(parts))
        for p in parts:
            p.ptr = None

    def next_generation(self, c=0):
        if lib().orc_next_generation(self.ptr, c) != 0:
            raise ReferencePanic(lib().orc_last_error(self.ptr).decode())

    def step(self, transfer=None, rule=0):
        T = np.ascontiguousarray(np.eye(self.C) if transfer is None else transfer, dtype=np.float64)
        assert T.shape == (self.C, self.C)
        if lib().orc_step(self.ptr, _p(T, C.c_double), rule) != 0:
            raise ReferencePanic(lib().orc_last_error(self.ptr).decode())

    def last_indices(self, t, o):
        n = int(lib().orc_last_indices(self.ptr, t, o, None))
        out = np.zeros(max(1, n), dtype=np.uint64)
        lib().orc_last_indices(self.ptr, t, o, _p(out, C.c_uint64))
        return out[:n]

    def infectant_generations(self, c=0):
        return int(lib().orc_infectant_generations(self.ptr, c))

    # --- draw log -----------------------------------------------------------
    def get_infections(self, c=0):
        out = np.zeros(self.population_len(c), dtype=np.int64)
        lib().orc_get_infections(self.ptr, c, _p(out, C.c_int64))
        return out

    def get_host_map(self, c=0):
        offsets = np.zeros(self.H + 1, dtype=np.uint64)
        hosts = np.zeros(max(1, self.population_len(c)), dtype=np.uint64)
        n = int(lib().orc_get_host_map(self.ptr, c, _p(offsets, C.c_uint64), _p(hosts, C.c_uint64)))
        return offsets, hosts[:n]

    def get_mutation_log(self, c=0):
        n = self.population_len(c)
        total = int(lib().orc_get_mutation_log(self.ptr, c, None, None, None))
        offsets = np.zeros(n + 1, dtype=np.uint64)
        sites = np.zeros(max(1, total), dtype=np.uint32)
        bases = np.zeros(max(1, total), dtype=np.uint8)
        lib().orc_get_mutation_log(self.ptr, c, _p(offsets, C.c_uint64), _p(sites, C.c_uint32), _p(bases, C.c_uint8))
        return offsets, sites[:total], bases[:total]

    def get_recombination_log(self, c=0):
        n = int(lib().orc_get_recombination_log(self.ptr, c, None, None, None, None, None, None))
        m = max(1, n)
        host = np.zeros(m, dtype=np.uint64)
        i, j, s0, s1 = (np.zeros(m, dtype=np.uint32) for _ in range(4))
        which = np.zeros(m, dtype=np.uint8)
        lib().orc_get_recombination_log(self.ptr, c, _p(host, C.c_uint64), _p(i, C.c_uint32), _p(j, C.c_uint32),
                                        _p(s0, C.c_uint32), _p(s1, C.c_uint32), _p(which, C.c_uint8))
        return host[:n], i[:n], j[:n], s0[:n], s1[:n], which[:n]

    def get_offspring(self, c=0):
        out = np.zeros(self.population_len(c), dtype=np.uint64)
        lib().orc_get_offspring(self.ptr, c, _p(out, C.c_uint64))
        return out

    def get_replication_terms(self, c=0):
        n = self.population_len(c)
        f, s, l = (np.zeros(n, dtype=np.float64) for _ in range(3))
        lib().orc_get_replication_terms(self.ptr, c, _p(f, C.c_double), _p(s, C.c_double), _p(l, C.c_double))
        return f, s, l

    # --- export -------------------------------------------------------------
    def get_raw_fitness(self, spec=0, which=0, c=0):
        out = np.zeros(self.population_len(c), dtype=np.float64)
        lib().orc_get_raw_fitness(self.ptr, c, spec, which, _p(out, C.c_double))
        return out

    def export_mutations(self, c=0):
        n = self.population_len(c)
        total = int(lib().orc_export_mutations(self.ptr, c, None, None, None))
        offsets = np.zeros(n + 1, dtype=np.uint64)
        pos = np.zeros(max(1, total), dtype=np.uint32)
        base = np.zeros(max(1, total), dtype=np.uint8)
        lib().orc_export_mutations(self.ptr, c, _p(offsets, C.c_uint64), _p(pos, C.c_uint32), _p(base, C.c_uint8))
        return offsets, pos[:total], base[:total]

    def get_base(self, infectant, pos, c=0):
        return int(lib().orc_get_base(self.ptr, c, infectant, pos))

    def get_population_ids(self, c=0):
        out = np.zeros(self.population_len(c), dtype=np.int64)
        lib().orc_get_population_ids(self.ptr, c, _p(out, C.c_int64))
        return out

    # --- haplotype API (reference KATs) ---------------------------------------
    def wildtype(self):
        return 0

    def create_descendant(self, parent, positions, bases):
        pos = np.ascontiguousarray(positions, dtype=np.uint32)
        b = np.ascontiguousarray(bases, dtype=np.uint8)
        r = int(lib().orc_hap_create_descendant(self.ptr, parent, _p(pos, C.c_uint32), _p(b, C.c_uint8), pos.size))
        if r < 0:
            raise ReferencePanic("assert_ne!(from, to) in create_descendant")
        return r

    def create_recombinant(self, left, right, lp, rp):
        return int(lib().orc_hap_create_recombinant(self.ptr, left, right, lp, rp))

    def hap_get_base(self, hid, pos):
        return int(lib().orc_hap_get_base(self.ptr, hid, pos))

    def hap_get_sequence(self, hid):
        out = np.zeros(self.L, dtype=np.uint8)
        lib().orc_hap_get_sequence(self.ptr, hid, _p(out, C.c_uint8))
        return out

    def hap_get_mutations(self, hid):
        cap = 1 << 16
        pos = np.zeros(cap, dtype=np.uint32)
        base = np.zeros(cap, dtype=np.uint8)
        n = int(lib().orc_hap_get_mutations(self.ptr, hid, _p(pos, C.c_uint32), _p(base, C.c_uint8), cap))
        return {int(p): int(b) for p, b in zip(pos[:n], base[:n])}

    def hap_raw_fitness(self, hid, spec=0, which=0):
        return float(lib().orc_hap_raw_fitness(self.ptr, hid, spec, which))

    def hap_fitness(self, hid, spec=0, which=0):
        return float(lib().orc_hap_fitness(self.ptr, hid, spec, which))

    def table_get_value(self, pos, base, spec=0, which=0):
        return float(lib().orc_table_get_value(self.ptr, spec, which, pos, base))

    def set_infectant(self, i, hid, c=0):
        lib().orc_set_infectant(self.ptr, c, i, hid)

    def n_nodes(self):
        return int(lib().orc_n_nodes(self.ptr))

    def num_threads(self):
        return int(lib().orc_num_threads(self.ptr))

    def set_mode(self, no_logs=False, fully_parallel=False):
        """Timing runs: no draw logs (the reference keeps none); fully_parallel: infect, replicate and the sample draws
        run over all threads too (the reference's rayon build runs them serially)."""
        lib().orc_set_mode(self.ptr, (1 if no_logs else 0) | (2 if fully_parallel else 0))


def test_poisson(seed, lam, n):
    out = np.zeros(n, dtype=np.float64)
    lib().orc_test_poisson(seed, lam, n, _p(out, C.c_double))
    return out


def test_binomial(seed, nn, p, n):
    out = np.zeros(n, dtype=np.uint64)
    lib().orc_test_binomial(seed, nn, p, n, _p(out, C.c_uint64))
    return out


def test_alias(seed, weights, n):
    w = np.ascontiguousarray(weights, dtype=np.uint64)
    out = np.zeros(n, dtype=np.uint64)
    if lib().orc_test_alias(seed, _p(w, C.c_uint64), w.size, n, _p(out, C.c_uint64)) != 0:
        raise ReferencePanic("all weights zero")
    return out
