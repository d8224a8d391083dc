"""Worker of tests/test_gpu_multidevice.py: one rank per GPU (torchrun).  Every rank runs its compartment through the peer
exchange over NVLink (CUDA IPC mappings of the other ranks' receive buffers) and, on its own GPU only, ALL compartments
through the same exchange with in-process "ranks" (raw pointers instead of IPC handles).  The two must give the same
population for the rank's compartment, infectant by infectant."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from virolution_b200.distributed import PeerExchange, migration_matrix  # noqa: E402
from virolution_b200.simulation import GpuSimulation, PeerExchangeHandle  # noqa: E402
from virolution_b200.workloads import workload  # noqa: E402


def state(g):
    off, pos, base = g.export_mutations()
    return off, pos, base, g.get_raw_fitness().view(np.uint64)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    n, gens, seed = int(sys.argv[1]), int(sys.argv[2]), 17
    wl = workload("k2_hiv", n)
    T = np.asarray(migration_matrix(world, 0.15), dtype=np.float64)
    T2 = np.asarray(migration_matrix(world, 0.05), dtype=np.float64)
    Tcap = np.maximum(T, T2)

    def make(k, device):
        s = GpuSimulation(wl["wildtype"], wl["parameters"], wl["specs"], seed=seed, device=device, compartment_id=k)
        s.set_population_wildtype(wl["n0"])
        return s

    # ---- across devices
    mine = make(rank, dev)
    px = PeerExchange(mine, rank, world, transfer=Tcap)
    px.set_transfer(T)
    px.run(gens)
    px.set_transfer(T2)       # a `transmission` event
    px.step()
    px.run(2)
    mine.synchronize()
    got = state(mine)
    # ---- the same compartments on this GPU alone
    sims = [make(k, dev) for k in range(world)]
    ex = [PeerExchangeHandle(sims[k], k, world, Tcap, 0) for k in range(world)]
    bufs = [e.recv_buffer() for e in ex]
    for k in range(world):
        ex[k].connect(None, [bufs[t] if t != k else 0 for t in range(world)])

    def local(g, Tm):
        for e in ex:
            e.set_transfer(Tm)
        for _ in range(g):
            for e in ex:
                e.push()
            for e in ex:
                e.pull()

    local(gens, T)
    local(3, T2)
    for s in sims:
        s.synchronize()
    want = state(sims[rank])
    ok = all(np.array_equal(x, y) for x, y in zip(got, want)) and mine.population_len() == sims[rank].population_len()
    mutants = int((np.diff(got[0].astype(np.int64)) > 0).sum())
    flag = torch.tensor([1 if ok and mutants > 100 else 0], device=f"cuda:{dev}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"MULTIDEVICE {'OK' if int(flag.item()) == 1 else 'MISMATCH'} world={world} n={n} population={mine.population_len()} mutants={mutants}", flush=True)
    for e in ex:
        e.close()
    for s in sims:
        s.close()
    px.close()
    mine.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
