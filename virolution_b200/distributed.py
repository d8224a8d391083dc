"""Transmission between compartments that live in different processes — one process (rank) per GPU.

Reference: the transmission block of `Runner::run`, parallel build (src/runner.rs:401-422):

    new[target] = concat over origin of  sim[origin].subsample_population(&offspring[origin], T[target][origin])

with `T = schedule.get_transfer_matrix(generation)` (row = target, col = origin, src/config/schedule.rs:96-98).
Here compartment c is owned by rank c.  Everything before the transmission block (increment_generation, infect,
mutate_infectants, replicate_infectants) is local to a rank; the block itself becomes

  1. the origin draws its `world` parts on its own GPU (column `rank` of T),
  2. the off-diagonal parts are flattened into one migrant image whose per-target ranges are contiguous
     (vl_migrants_export_device), sizes travel in one all-to-all of 2 x world integers,
  3. three all-to-all-v collectives (entry counts, raw fitness, entries) move the ranges — NCCL over NVLink when
     the process group is NCCL; the same code runs over gloo with CPU tensors in the tests,
  4. every target imports what it received and assembles its population in origin order.

There is no other collective on the data path: compartments are independent between transmission blocks.
`torch.distributed` is the plumbing; all population work happens in the CUDA library behind the C ABI.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def migration_matrix(n: int, migration: float) -> np.ndarray:
    """`1 - migration` on the diagonal, `migration / (n - 1)` elsewhere; identity for n == 1."""
    if n == 1:
        return np.eye(1)
    T = np.full((n, n), migration / (n - 1), dtype=np.float64)
    np.fill_diagonal(T, 1.0 - migration)
    return T


class _DevicePointer:
    """Lets torch view library-owned device memory without a copy (`__cuda_array_interface__`)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


_VIEWS = {}  # (device pointer, typestr) -> longest torch view made of it; the library's staging buffers only ever grow


def _view(ptr: int, n: int, cap: int, typestr: str, dtype: torch.dtype, device) -> torch.Tensor:
    """torch view of n elements of a library buffer that holds `cap`; the view of the whole buffer is made once."""
    if n == 0 or not ptr:
        return torch.empty(0, dtype=dtype, device=device)
    key = (ptr, typestr, str(device))
    v = _VIEWS.get(key)
    if v is None or v.numel() < n:
        if len(_VIEWS) > 64:
            _VIEWS.clear()
        v = torch.as_tensor(_DevicePointer(ptr, max(n, cap), typestr), device=device)
        _VIEWS[key] = v
    return v[:n]


class GpuCompartment:
    """Adapter between a `GpuSimulation` (C ABI handle) and the exchange below: parts are `GpuPopulation`s, images are
    torch views of the library's device buffers."""

    def __init__(self, sim):
        self.sim = sim
        self.device = torch.device("cuda", sim.device)
        self.n_providers = None

    def begin_generation(self):
        self.sim.advance_to_offspring()

    def subsample_parts(self, factors: Sequence[float]):
        return self.sim.subsample_populations(factors)

    def export_parts(self, parts):
        img, m, w = self.sim.export_migrants(parts)
        P = int(img.n_providers)
        self.n_providers = P
        M, W = int(img.n_migrants), int(img.n_words)
        cm, cw = int(img.cap_migrants), int(img.cap_words)
        return (_view(img.len_device, M, cm, "<i4", torch.int32, self.device),
                _view(img.raw_device, M * P, cm * P, "<f8", torch.float64, self.device),
                _view(img.entries_device, W, cw, "<i4", torch.int32, self.device), m.astype(np.int64), w.astype(np.int64), P)

    def import_part(self, lens: torch.Tensor, raw: torch.Tensor, entries: torch.Tensor):
        return self.sim.import_migrants(lens.data_ptr() if lens.numel() else 0, raw.data_ptr() if raw.numel() else 0,
                                        entries.data_ptr() if entries.numel() else 0, lens.numel(), entries.numel())

    def set_population(self, parts):
        self.sim.set_population(parts)

    def synchronize_collectives(self):
        torch.cuda.current_stream(self.device).synchronize()


class CompartmentExchange:
    """One compartment per rank; `step()` is one iteration of the loop body of Runner::run (src/runner.rs:382-422)."""

    def __init__(self, compartment, rank: int, world: int, transfer: Optional[np.ndarray] = None,
                 migration: float = 0.1, group=None):
        self.c, self.rank, self.world, self.group = compartment, rank, world, group
        self.T = np.asarray(migration_matrix(world, migration) if transfer is None else transfer, dtype=np.float64)
        if self.T.shape != (world, world):
            # check_transfer_table_sizes (src/config/schedule.rs:235-237)
            raise ValueError(f"transfer matrix is {self.T.shape}, expected {(world, world)}")
        self.bytes_sent = 0
        self.migrants_sent = 0
        self.trace = None  # set to {} to accumulate host seconds per phase of transmit()

    def _all_to_all(self, send: torch.Tensor, in_split: List[int], out_split: List[int]) -> torch.Tensor:
        recv = torch.empty(int(sum(out_split)), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(recv, send.contiguous(), out_split, in_split, group=self.group)
        return recv

    def transmit(self):
        """The transmission block: parts of this origin -> every target; returns nothing, installs the new population."""
        c, r, W = self.c, self.rank, self.world
        import time
        t_last = [time.perf_counter()]

        def mark(name):
            if self.trace is not None:
                now = time.perf_counter()
                self.trace[name] = self.trace.get(name, 0.0) + now - t_last[0]
                t_last[0] = now
        parts = c.subsample_parts([float(self.T[t, r]) for t in range(W)])
        mark("subsample")
        if W == 1:
            c.set_population(parts)
            return
        others = [t for t in range(W) if t != r]
        lens, raw, entries, m_per, w_per, P = c.export_parts([parts[t] for t in others])
        dev = lens.device
        mark("export")
        counts = torch.zeros(W, 2, dtype=torch.int64)
        for k, t in enumerate(others):
            counts[t, 0], counts[t, 1] = int(m_per[k]), int(w_per[k])
        send_counts = counts.to(dev)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        recv_counts = recv_counts.cpu()
        mark("counts")
        m_in, w_in = counts[:, 0].tolist(), counts[:, 1].tolist()
        m_out, w_out = recv_counts[:, 0].tolist(), recv_counts[:, 1].tolist()
        r_len = self._all_to_all(lens, m_in, m_out)
        r_raw = self._all_to_all(raw, [m * P for m in m_in], [m * P for m in m_out])
        r_ent = self._all_to_all(entries, w_in, w_out)
        c.synchronize_collectives()
        mark("all_to_all")
        self.migrants_sent += int(sum(m_in))
        self.bytes_sent += 4 * int(sum(m_in)) + 8 * P * int(sum(m_in)) + 4 * int(sum(w_in)) + 16 * W
        for t in others:
            parts[t].free()  # shipped
        new_parts, mo, wo = [], 0, 0
        for o in range(W):
            if o == r:
                new_parts.append(parts[r])
            else:
                new_parts.append(c.import_part(r_len[mo:mo + m_out[o]], r_raw[mo * P:(mo + m_out[o]) * P], r_ent[wo:wo + w_out[o]]))
            mo += m_out[o]
            wo += w_out[o]
        mark("import")
        c.set_population(new_parts)  # Population::from_iter: concatenation in origin order (src/core/population.rs:190-205)

        mark("set_population")

    def step(self):
        import time
        t0 = time.perf_counter()
        self.c.begin_generation()
        if self.trace is not None:
            self.trace["begin_generation(enqueue)"] = self.trace.get("begin_generation(enqueue)", 0.0) + time.perf_counter() - t0
        self.transmit()


def sharded_simulation(wildtype, parameters, host_specs, rank: int, world: int, seed: int = 0, device: int = 0, n0: int = 0):
    """The handle of rank `rank`'s shard of ONE compartment (include/virolution_gpu.h, vl_exchange_create sharded): the local
    range of hosts (every spec's range scaled by 1/world), the global max_population, room for 1.4 x its share."""
    from .abi import HostSpec, Parameters
    from .simulation import GpuSimulation
    H, M = int(parameters.host_population_size), int(parameters.max_population)
    if H % world:
        raise ValueError("host_population_size must be a multiple of the number of shards")
    local = Parameters(parameters.mutation_rate, parameters.recombination_rate, H // world, parameters.basic_reproductive_number, M,
                       parameters.dilution, parameters.substitution_matrix, parameters.n_hits)
    specs = [HostSpec(s.lo // world, s.hi // world, s.reproductive, s.infective) for s in host_specs]
    sim = GpuSimulation(wildtype, local, specs, seed=seed, device=device, compartment_id=rank,
                        capacity_infectants=int(1.4 * M / world) + 65536)
    sim.set_population_wildtype(n0 // world)
    return sim


class PeerExchange:
    """The same loop body with the transport fused into the kernels: every origin writes its migrants straight into the
    targets' device memory over NVLink (CUDA IPC mappings), sizes stay on the device, the host only enqueues
    (vl_exchange_* in include/virolution_gpu.h, kernels in csrc/vl_exchange.cuh).  `torch.distributed` is used once, to
    hand the 64-byte IPC handles around."""

    def __init__(self, sim, rank: int, world: int, transfer: Optional[np.ndarray] = None, migration: float = 0.1,
                 words_per_migrant: int = 0, group=None, sharded: bool = False, shard_seed: int = 0):
        from .simulation import PeerExchangeHandle
        self.sim, self.rank, self.world = sim, rank, world
        if sharded:  # one compartment over all ranks (sim from sharded_simulation): the split of the draws is decided on the device
            self.T = None
            self.handle = PeerExchangeHandle(sim, rank, world, None, words_per_migrant, sharded=True, shard_seed=shard_seed)
        else:
            self.T = np.asarray(migration_matrix(world, migration) if transfer is None else transfer, dtype=np.float64)
            self.handle = PeerExchangeHandle(sim, rank, world, self.T, words_per_migrant)
        dev = torch.device("cuda", sim.device)
        mine = torch.from_numpy(self.handle.ipc_handle.copy()).to(dev)
        allh = torch.empty(world * mine.numel(), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, mine, group=group)
        self.handle.connect(allh.cpu().numpy().reshape(world, -1))
        dist.barrier(group=group)  # every rank has mapped every buffer before anybody writes
        self.trace = None

    def step(self):
        self.handle.step()

    def run(self, n_generations: int):
        """n generations enqueued back to back (vl_exchange_run): the ranks meet on the device only."""
        self.handle.run(n_generations)

    def set_transfer(self, T):
        """A `transmission` event of the schedule: the matrix of the generations that follow.  The exchange must have been
        created with the entry-wise maximum of all matrices of the schedule (the buffers are sized once)."""
        self.T = np.asarray(T, dtype=np.float64)
        self.handle.set_transfer(self.T)

    def close(self):
        self.sim.synchronize()
        dist.barrier()
        self.handle.close()
