"""Host-side logic of the cross-process transmission block (virolution_b200/distributed.py) over gloo, world_size 2 and 3.

The exchange is written against a small compartment protocol; on a GPU box the compartment is the CUDA library
(GpuCompartment), here it is a CPU stand-in with the same image format (entry counts with the wildtype marker,
raw fitness per provider, flattened entries), so the test pins: who receives what, in which order the parts are
concatenated (origin order, src/runner.rs:404-416 + src/core/population.rs:190-205), and that every migrant's
haplotype payload arrives intact.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT  # noqa: F401  (puts the repo on sys.path)
from virolution_b200.distributed import CompartmentExchange, migration_matrix

MIG_WILDTYPE = -0x80000000  # 0x80000000 as int32


class FakeCompartment:
    """Population = list of haplotypes (entries tuple, raw fitness tuple); subsample = the first floor(f * n) infectants."""

    def __init__(self, rank, n, n_providers=2):
        rng = np.random.default_rng(100 + rank)
        self.P = n_providers
        self.pop = []
        for i in range(n):
            if i % 3 == 0:
                self.pop.append(((), tuple([1.0] * n_providers), True))  # the wildtype
            else:
                k = int(rng.integers(0, 4))  # a mutant may have an empty list (back mutation) and still not be the wildtype
                ent = tuple(int(x) for x in rng.integers(0, 1 << 20, size=k) + (rank << 24))
                self.pop.append((ent, tuple(float(x) for x in rng.random(n_providers)), False))
        self.generation = 0

    def begin_generation(self):
        self.generation += 1

    def subsample_parts(self, factors):
        return [list(self.pop[: int(f * len(self.pop))]) for f in factors]

    def export_parts(self, parts):
        lens, raw, ent, m, w = [], [], [], [], []
        for part in parts:
            m.append(len(part))
            w.append(sum(len(h[0]) for h in part))
            for e, r, is_wt in part:
                lens.append(MIG_WILDTYPE if is_wt else len(e))
                raw.extend(r)
                ent.extend(e)
        return (torch.tensor(lens, dtype=torch.int32), torch.tensor(raw, dtype=torch.float64), torch.tensor(ent, dtype=torch.int32),
                np.array(m, dtype=np.int64), np.array(w, dtype=np.int64), self.P)

    def import_part(self, lens, raw, entries):
        out, wo = [], 0
        for i, l in enumerate(lens.tolist()):
            r = tuple(raw[i * self.P:(i + 1) * self.P].tolist())
            if l == MIG_WILDTYPE:
                out.append(((), r, True))
            else:
                out.append((tuple(entries[wo:wo + l].tolist()), r, False))
                wo += l
        assert wo == entries.numel()
        return _Part(out)

    def set_population(self, parts):
        self.pop = [h for p in parts for h in p]

    def synchronize_collectives(self):
        pass


class _Part(list):
    def free(self):
        pass


def _worker(rank, world, port, T, sizes, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = FakeCompartment(rank, sizes[rank])
        # parts handed out by subsample_parts are plain lists: give them the free() the exchange calls on shipped parts
        orig = c.subsample_parts
        c.subsample_parts = lambda f: [_Part(p) for p in orig(f)]
        ex = CompartmentExchange(c, rank, world, transfer=T)
        ex.step()
        expect = []
        for o in range(world):
            src = FakeCompartment(o, sizes[o]).pop
            expect.extend(src[: int(T[rank, o] * len(src))])
        assert c.pop == expect, f"rank {rank}: population differs from concat over origins"
        assert c.generation == 1
        torch.save({"n": len(c.pop), "sent": ex.migrants_sent}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_over_gloo(world, tmp_path):
    T = migration_matrix(world, 0.25)
    T[0, world - 1] = 0.0  # a zero entry: nothing travels on that edge
    sizes = [40 + 13 * r for r in range(world)]
    mp.spawn(_worker, args=(world, _free_port(), T, sizes, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), f"r{r}.pt")) for r in range(world)]
    for t in range(world):
        assert res[t]["n"] == sum(int(T[t, o] * sizes[o]) for o in range(world))
    for o in range(world):
        assert res[o]["sent"] == sum(int(T[t, o] * sizes[o]) for t in range(world) if t != o)


def test_migration_matrix_rows_and_shape_check():
    T = migration_matrix(8, 0.1)
    assert T.shape == (8, 8) and np.allclose(T.sum(axis=1), 1.0) and np.allclose(np.diag(T), 0.9)
    assert np.allclose(T[0, 1], 0.1 / 7)
    assert np.array_equal(migration_matrix(1, 0.1), np.eye(1))
    with pytest.raises(ValueError):
        CompartmentExchange(FakeCompartment(0, 4), 0, 2, transfer=np.eye(3))
