#!/usr/bin/env python
"""bench.py — infectant-generations/sec of the B200 population-update path (BASELINE.json metric).

A "step" is one generation (infect -> mutate -> fitness -> replicate -> subsample/transmit) of the workload.
    python bench.py --gpus 1 --steps K --warmup W            our arm (CUDA path through the C ABI)
    python bench.py --impl reference ...                     the reference's CPU algorithm (oracle port, host cores)
    torchrun ... bench.py --gpus N ...                       one compartment per GPU, migrants pushed into peer memory (weak scaling)

Keys of the JSON line:
  value         device-resident loop (vl_run), timed with CUDA events on the library's stream, max over ranks; the
                compactions of the record table that fall into the K generations are part of it
  e2e           the trait-level calls a `impl Simulation for GpuSimulation` makes, with HOST offspring vectors:
                replicate_infectants -> Vec<usize> (D2H), subsample_population(&offspring) (H2D), set_population
  roofline      dominant kernel: algorithmic bytes per launch / average CUDA-event duration vs measured HBM peak
                (per-kernel durations come from a separate pass with events around every launch)
  long_run      ms per generation at generations 25, 200 and 1000 of one run (haplotype records grow with time)
  extra         the other BASELINE configs, bounded: K1, K3 (8 compartments on this GPU), K4, K5 (one GPU)
  cpu_baseline  the oracle (CPU restatement of the reference loop) on a bounded sample, rank 0, N=1 only
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "infectant-generations/sec"
UNIT = "infectant-generations/s"
MIN_WARMUP = 20  # haplotype diversity at quasi-steady state before timing

# Algorithmic bytes per infectant for the kernels of the device-resident generation: u32
# ids / hosts / offspring, f64 fitness.  k_mutate works per new record and is accounted from the records it made.
ALGO_BYTES = {
    "k_offspring_sum<true>": 4 + 8 + 8 + 4 + 8,        # host_id, carried fitness, per-host sum (gather), offspring, clearing the other sum buffer
    "k_offspring_sum<false>": 4 + 8 + 8 + 4 + 8,
    # <PREFILL, MULTI>: offspring, hap, fitness in; hap, fitness out; PREFILL: next generation's host out + per-host sum (reduction)
    "(k_resample_pos<true, false>)": 4 + 4 + 8 + 4 + 8 + 4 + 8, "(k_resample_pos<true, true>)": 4 + 4 + 8 + 4 + 8 + 4 + 8,
    "(k_resample_pos<true, false, true>)": 4 + 4 + 8 + 4 + 8 + 4 + 8, "(k_resample_pos<true, true, true>)": 4 + 4 + 8 + 4 + 8 + 4 + 8,
    "(k_resample_pos<false, false>)": 4 + 4 + 8 + 4 + 8, "(k_resample_pos<false, true>)": 4 + 4 + 8 + 4 + 8,
    "k_infect_sum<false>": 4 + 8 + 4 + 8, "k_infect_sum<true>": 4 + 8 + 8 + 4 + 8,
    # general kernels (trait-level path)
    "k_replicate_fused<true>": 4 + 8 + 8 + 4, "k_replicate_fused<false>": 4 + 8 + 8 + 4,
    "(k_infect_bucket<false, false, 1024>)": 4 + 16, "(k_infect_bucket<true, true, 1024>)": 4 + 4 + 16,
    "(k_bucket_sort<1024, BSORT_STAGE_RECS>)": 16 + 8 + 4, "k_resample": 4 + 8 + 4,
}
# the kernels SURVEY.md §8d's 72 B figure covers (fitness gather + per-host sum + offspring + scan/plan + resample); in the
# device-resident generation the resample also draws the next generation's hosts and mutation counts (+ 8 B "infect")
FITNESS_REPLICATION = ("k_offspring_sum<true>", "k_offspring_sum<false>", "k_plan_resample", "(k_resample_pos<true, false>)", "(k_resample_pos<true, true>)", "(k_resample_pos<true, false, true>)", "(k_resample_pos<true, true, true>)",
                       "(k_resample_pos<false, false>)", "(k_resample_pos<false, true>)",
                       "k_infect_sum<false>", "k_infect_sum<true>", "k_replicate_fused<true>", "k_replicate_fused<false>", "k_fitness",
                       "k_host_sum", "k_host_sum_medium", "k_host_sum_heavy", "k_offspring<true>", "k_offspring<false>", "k_resample")
GROUP_BYTES = 72 + 8


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="k2_hiv")
    ap.add_argument("--n", "--population", dest="n", type=int, default=0, help="override N = H = max_population")
    ap.add_argument("--sharded", action="store_true", help="N > 1: ONE compartment of the workload split over the GPUs (strong scaling) "
                    "instead of one compartment per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the long run and the other BASELINE configs")
    ap.add_argument("--profile-steps", type=int, default=10)
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            time.sleep(0.3)  # first sample before the timed region starts
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for k, name in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def pinned_u64(n):
    import torch
    t = torch.empty(max(n, 1), dtype=torch.int64, pin_memory=True)
    return t, t.numpy().view(np.uint64)


def trait_generation(sim, off_host):
    """One generation the way Runner::run drives `dyn Simulation` (src/runner.rs:382-422), host offspring vector."""
    n = sim.population_len()
    sim.increment_generation()
    sim.infect()
    sim.mutate_infectants()
    off = sim.replicate_infectants(out=off_host[:n])           # Vec<usize> back to the host
    pop = sim.subsample_population(off, 1.0)                   # &[usize] from the host
    sim.set_population(pop)
    return n


def timed_run(sim, gens):
    """ms per generation and infectant-generations/s of `gens` device-resident generations (CUDA events on the library's stream)."""
    sim.infectant_generations(reset=True)
    sim.timer_start()
    sim.run(gens)
    ms = sim.timer_stop()
    ig = sim.infectant_generations(reset=True)
    return ms / gens, ig / (ms / 1e3)


def long_run(device):
    """One K2 run to generation 1020: ms per generation around generations 25, 200 and 1000 (record lists grow with time)."""
    from virolution_b200.simulation import GpuSimulation
    from virolution_b200.workloads import workload
    wl = workload("k2_hiv", None)
    sim = GpuSimulation(wl["wildtype"], wl["parameters"], wl["specs"], seed=77, device=device)
    sim.set_population_wildtype(wl["n0"])
    out, at = {}, 0
    for target in (25, 200, 1000):
        sim.run(target - 10 - at)
        ms, _ = timed_run(sim, 20)
        at = target + 10
        c = sim.get_counters()
        out[f"ms_per_generation_at_{target}"] = round(ms, 4)
        out[f"compactions_by_{target}"] = c["gc_runs"]
    out["ratio_1000_over_25"] = round(out["ms_per_generation_at_1000"] / out["ms_per_generation_at_25"], 3)
    out["workload"] = wl["description"]
    out["note"] = ("20 generations timed around each mark, compactions included; a record is a 64-byte head (shared base list + 8 inline delta "
                   "entries) and is flattened once per 8 mutations of its lineage, so the cost per mutation grows with length / 8")
    sim.close()
    return out


def extra_configs(device):
    """The other BASELINE configs, bounded (a fraction of a second each), device-resident loops."""
    from virolution_b200.simulation import GpuSimulation, step
    from virolution_b200.workloads import workload
    from virolution_b200.distributed import migration_matrix
    out = {}

    def single(name, n, warm, gens, label):
        wl = workload(name, n)
        sim = GpuSimulation(wl["wildtype"], wl["parameters"], wl["specs"], seed=5, device=device)
        sim.set_population_wildtype(wl["n0"])
        sim.run(warm)
        ms, v = timed_run(sim, gens)
        r = {"workload": wl["description"], "ms_per_generation": round(ms, 5), "value": v, "unit": UNIT, "generations_timed": gens,
             "warmup": warm, "bound": label}
        if "hbm" in label:
            peak, _ = measured_peak()
            r["achieved_GBps_at_95B_per_infectant_generation"] = round(95.0 * v / 1e9, 1)
            r["frac_of_measured_hbm"] = round(95.0 * v / 1e9 / peak, 4)
        sim.close()
        return r
    out["k1_singlehost"] = single("k1_singlehost", None, 100, 400, "launch/latency-bound (1 MB of traffic per generation): generations/s is the figure")
    out["k1_singlehost"]["generations_per_s"] = round(1e3 / out["k1_singlehost"]["ms_per_generation"], 1)
    out["k4_epistasis"] = single("k4_epistasis", None, 20, 20, "hbm/issue-bound like K2 at a tenth of its size (10^6 infectants, 10^5 sparse pairs)")
    out["k5_sars_1gpu"] = single("k5_sars", None, 10, 10, "hbm: recombination needs the host map, so the general kernels run (bucketed sort + fused replicate)")
    # K3: data/multihost.yaml semantics (two host specs per compartment), 8 compartments, 8 x 8 matrix 0.9 / 0.1 spread,
    # all on this GPU through vl_step (parallel-build rule): what one GPU does when the box has fewer GPUs than compartments
    wl = workload("k3_multihost", None)
    C = 8
    sims = [GpuSimulation(wl["wildtype"], wl["parameters"], wl["specs"], seed=9, device=device, compartment_id=c) for c in range(C)]
    for s in sims:
        s.set_population_wildtype(wl["n0"])
    T = migration_matrix(C, 0.1)
    for _ in range(10):
        step(sims, T, 0)
    for s in sims:
        s.synchronize()
        s.infectant_generations(reset=True)
    t0 = time.perf_counter()
    gens = 30
    for _ in range(gens):
        step(sims, T, 0)
    for s in sims:
        s.synchronize()
    dt = time.perf_counter() - t0
    ig = sum(s.infectant_generations(reset=True) for s in sims)
    out["k3_multihost_8_compartments_1gpu"] = {"workload": wl["description"] + f" x {C} compartments on one GPU, vl_step (parallel-build transmission rule)",
                                               "ms_per_generation": round(dt / gens * 1e3, 4), "value": ig / dt, "unit": UNIT,
                                               "generations_timed": gens, "bound": "launch/host-bound: 8 compartments of 10^5 driven through trait-level "
                                               "enqueues with one synchronisation per generation (wall clock)"}
    # the same eight compartments through the in-process peer exchange (CompartmentGroup): nothing but enqueues
    from virolution_b200.simulation import CompartmentGroup
    group = CompartmentGroup(sims, T)
    group.run(10)
    group.synchronize()
    for s in sims:
        s.infectant_generations(reset=True)
    t0 = time.perf_counter()
    gens = 100
    group.run(gens)
    group.synchronize()
    dt = time.perf_counter() - t0
    ig = sum(s.infectant_generations(reset=True) for s in sims)
    sizes = [s.population_len() for s in sims]
    out["k3_multihost_8_compartments_1gpu_group"] = {"workload": wl["description"] + f" x {C} compartments on one GPU, CompartmentGroup (peer exchange between "
                                                     "the compartments of one process, parallel-build transmission rule)",
                                                     "ms_per_generation": round(dt / gens * 1e3, 4), "value": ig / dt, "unit": UNIT, "generations_timed": gens,
                                                     "population_per_compartment": sizes,
                                                     "bound": "launch-bound: ~40 kernels per compartment and generation on eight streams (wall clock)"}
    group.close()
    for s in sims:
        s.close()
    return out


def run_b200(args, rank, world):
    from virolution_b200 import abi
    from virolution_b200.simulation import GpuSimulation
    from virolution_b200.workloads import workload
    dist = None
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = workload(args.workload, args.n or None)
    flags = 0
    par = wl["parameters"]
    N = wl["n0"]
    sharded = args.sharded and world > 1
    if sharded:
        from virolution_b200.distributed import PeerExchange, sharded_simulation
        sim = sharded_simulation(wl["wildtype"], par, wl["specs"], rank, world, seed=2026, device=local_rank, n0=N)
    else:
        sim = GpuSimulation(wl["wildtype"], par, wl["specs"], seed=2026, device=local_rank, compartment_id=rank, flags=flags)
        sim.set_population_wildtype(N)
    runner, T = None, None
    if sharded:
        runner = PeerExchange(sim, rank, world, sharded=True, shard_seed=2026)
        args.no_e2e = True  # the trait-level calls address one handle; a sharded compartment is driven through the exchange only
    elif world > 1:
        from virolution_b200.distributed import CompartmentExchange, GpuCompartment, PeerExchange, migration_matrix
        T = migration_matrix(world, 0.1)
        collective = CompartmentExchange(GpuCompartment(sim), rank, world, migration=0.1)  # e2e leg: host vectors + NCCL all-to-all-v
        # device-resident loop: migrants are pushed by the kernels into peer memory (VL_EXCHANGE=nccl keeps the collective path)
        runner = collective if os.environ.get("VL_EXCHANGE") == "nccl" else PeerExchange(sim, rank, world, migration=0.1)

    def advance(k):
        if runner is None:
            sim.run(k)
        else:
            if hasattr(runner, "run") and os.environ.get("VL_EXCHANGE_STEPWISE") != "1":
                runner.run(k)
            else:
                for _ in range(k):
                    runner.step()

    def barrier():
        sim.synchronize()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    warm = max(args.warmup, MIN_WARMUP)
    # ---- warm-up (untimed): also brings haplotype diversity towards its quasi-steady state
    advance(warm)
    barrier()
    sim.infectant_generations(reset=True)
    ct0 = sim.get_counters()
    l0 = ct0["kernels_launched"]
    # ---- timed region: exactly K generations, device-resident
    with ClockSampler(local_rank) as clocks:
        barrier()
        t_wall = time.perf_counter()
        sim.timer_start()
        advance(args.steps)
        ms = sim.timer_stop()
        barrier()
        t_wall = time.perf_counter() - t_wall
    ig = sim.infectant_generations(reset=True)
    ct1 = sim.get_counters()
    launches = ct1["kernels_launched"] - l0
    n_after = sim.population_len()
    exchange_checked = None
    if dist is not None:
        import torch
        t = torch.tensor([ms, float(ig)], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, ig = float(tmax[0]), int(tsum[1])
        if not sharded and T is not None:
            # invariants of the transmission block (src/runner.rs:401-416, src/simulation.rs:516-527): the population of target t
            # is the sum over origins of floor(T[t][o] * target_size(o)), and target_size = max_population when the cap binds (K2)
            expect = sum(int(T[rank, o] * par.max_population) for o in range(world))
            sizes = torch.tensor([float(n_after), float(expect)], dtype=torch.float64, device="cuda")
            allsz = [torch.zeros_like(sizes) for _ in range(world)]
            dist.all_gather(allsz, sizes)
            ok = all(int(a[0]) == int(a[1]) for a in allsz)
            mf = sim.mean_fitness()
            exchange_checked = {"ok": bool(ok and np.isfinite(mf) and mf > 0), "population_per_rank": [int(a[0]) for a in allsz],
                                "expected_per_rank": [int(a[1]) for a in allsz],
                                "rule": "population[t] == sum_o floor(T[t][o] * max_population) on every rank; mean mapped fitness finite and positive"}
        elif sharded:
            tot = torch.tensor([float(n_after)], dtype=torch.float64, device="cuda")
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            exchange_checked = {"ok": int(tot[0]) == int(par.max_population), "population_total": int(tot[0]), "expected_total": int(par.max_population),
                                "rule": "the shards of the compartment hold (min(W * dilution, max_population)) as usize infectants together"}
    value = ig / (ms / 1e3)
    compaction = None
    if rank == 0:
        cap = sim.get_capacities()
        compaction = {"runs_in_timed_region": int(ct1["gc_runs"] - ct0["gc_runs"]), "records_at_end": ct1["haplotype_records"],
                      "record_capacity": cap["haplotype_records"],
                      "policy": "bit-map compaction when the table passes 2 x the infectant capacity; its kernels are inside the timed region"}

    # ---- per-kernel CUDA-event durations over the same loop (separate pass: events around every launch perturb the step time)
    roofline, kernels = None, {}
    if rank == 0:
        sim.profile(True)
    c0 = sim.get_counters()
    advance(args.profile_steps)
    barrier()
    c1 = sim.get_counters()
    if rank == 0:
        rep = sim.profile_report()
        sim.profile(False)
        n_now = sim.population_len()
        peak, peak_src = measured_peak()
        # k_mutate works per new haplotype record: item 16 + parent head 64 + base-list window 32 + table rows 32 read; head 64 + memo
        # 16 + carried fitness 8 + per-host sum 8 written; flattened lists are read and written once
        mutate_bytes = None
        k_mut = next((k for k in rep if k.lstrip("(").startswith("k_mutate")), None)  # k_mutate<false> / k_mutate<true> (epistatic providers)
        if c1["gc_runs"] == c0["gc_runs"] and k_mut and rep[k_mut][0]:
            d_rec, d_words = c1["haplotype_records"] - c0["haplotype_records"], c1["arena_words"] - c0["arena_words"]
            mutate_bytes = (240.0 * d_rec + 8.0 * max(d_words, 0)) / rep[k_mut][0]
        tot_ms = sum(v[1] for v in rep.values())
        for name, (cnt, tms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
            per_launch_ms = tms / cnt
            ab = ALGO_BYTES.get(name)
            kernels[name] = {"launches_per_step": cnt / args.profile_steps, "ms_per_launch": round(per_launch_ms, 5), "share": round(tms / tot_ms, 4)}
            if ab is not None:
                kernels[name]["GBps"] = round(ab * n_now / (per_launch_ms * 1e-3) / 1e9, 1)
            elif name == k_mut and mutate_bytes is not None:
                kernels[name]["GBps"] = round(mutate_bytes / (per_launch_ms * 1e-3) / 1e9, 1)
                kernels[name]["algorithmic_bytes_per_launch"] = round(mutate_bytes)
        launch_bytes = {k: ALGO_BYTES[k] * n_now for k in rep if k in ALGO_BYTES}
        if mutate_bytes is not None:
            launch_bytes[k_mut] = mutate_bytes
        if launch_bytes:
            dom = max(launch_bytes, key=lambda k: rep[k][1])  # the kernel that takes most of the step
            dcnt, dms = rep[dom]
            achieved = launch_bytes[dom] / (dms / dcnt * 1e-3) / 1e9
            fr_ms = sum(rep[k][1] / args.profile_steps for k in FITNESS_REPLICATION if k in rep)
            fr_bytes = float(GROUP_BYTES) * n_now
            traffic = None
            try:  # DRAM bytes per launch of the same kernel from the committed ncu --set full capture (same workload size only)
                tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
                if tj.get("n") == n_now and dom in tj:
                    traffic = tj[dom]
            except Exception:
                pass
            roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                        "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": round(launch_bytes[dom]), "infectants_per_launch": n_now,
                        "durations": "CUDA events around every launch, separate pass of %d generations after the timed region" % args.profile_steps,
                        "by_kernel": {k: {"ms_per_launch": round(rep[k][1] / rep[k][0], 5),
                                          "achieved": round(launch_bytes[k] / (rep[k][1] / rep[k][0] * 1e-3) / 1e9, 1),
                                          "frac": round(launch_bytes[k] / (rep[k][1] / rep[k][0] * 1e-3) / 1e9 / peak, 4)}
                                      for k in sorted(launch_bytes, key=lambda k: -rep[k][1])[:4]},
                        "fitness_replication_group": {"kernels": [k for k in FITNESS_REPLICATION if k in rep],
                                                      "ms_per_generation": round(fr_ms, 5), "algorithmic_bytes_per_infectant": GROUP_BYTES,
                                                      "bytes_note": "SURVEY.md §8d: 72 B fitness + replication + scan + resample, + 8 B infect (the resample "
                                                                    "of this loop also draws the next generation's hosts and mutation counts)",
                                                      "achieved": round(fr_bytes / (fr_ms * 1e-3) / 1e9, 1),
                                                      "frac": round(fr_bytes / (fr_ms * 1e-3) / 1e9 / peak, 4)} if fr_ms > 0 else None}

    # ---- end to end through the trait-level C-ABI calls with host buffers
    e2e = None
    if not args.no_e2e and runner is None:
        keep, off_host = pinned_u64(max(N, par.max_population) * 2)
        for _ in range(3):
            trait_generation(sim, off_host)
        sim.synchronize()
        k_e2e = max(5, min(args.steps, 20))
        t0 = time.perf_counter()
        total, bytes_d2h, bytes_h2d = 0, 0, 0
        for _ in range(k_e2e):
            n = trait_generation(sim, off_host)
            total += n
            bytes_d2h += 8 * n + 8
            bytes_h2d += 8 * n
        sim.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": total / dt, "unit": UNIT, "h2d_bytes_per_step": bytes_h2d // k_e2e, "d2h_bytes_per_step": bytes_d2h // k_e2e,
               "steps": k_e2e, "api": "vl_infect/vl_mutate_infectants/vl_replicate_infectants(host Vec<usize>)/vl_subsample_population(host &[usize])/vl_set_population"}
    elif runner is not None and not args.no_e2e:
        # the same trait-level generation per rank (offspring Vec<usize> to the host and back), then the NCCL exchange
        keep, off_host = pinned_u64(max(N, par.max_population) * 2)

        def dist_generation():
            n = sim.population_len()
            sim.increment_generation()
            sim.infect()
            sim.mutate_infectants()
            off = sim.replicate_infectants(out=off_host[:n])
            sim.target_size(off)      # &[usize] from the host: uploads the map the subsamples below draw from
            collective.transmit()
            return n
        for _ in range(2):
            dist_generation()
        barrier()
        k_e2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        total = 0
        for _ in range(k_e2e):
            total += dist_generation()
        barrier()
        dt = time.perf_counter() - t0
        import torch
        t = torch.tensor([float(total), dt], dtype=torch.float64, device="cuda")
        tsum, tmax = t.clone(), t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e = {"value": float(tsum[0]) / float(tmax[1]), "unit": UNIT, "h2d_bytes_per_step": int(8 * total // k_e2e) * world,
               "d2h_bytes_per_step": int((8 * total + 8 * k_e2e) // k_e2e) * world, "steps": k_e2e,
               "api": "per rank: vl_infect/vl_mutate_infectants/vl_replicate_infectants(host Vec<usize>)/vl_target_size(host &[usize])/"
                      "vl_subsample_populations + NCCL all-to-all-v of migrant images + vl_set_population"}

    out = None
    if rank == 0:
        clk = clocks.summary()
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "impl": "b200",
               "config": {"workload": wl["description"] + ("" if world == 1 or sharded else f" per compartment x {world} compartments"), "compartments": 1 if sharded else world,
                          "parallelism": "1 compartment" if world == 1 else (f"1 compartment sharded over {world} GPUs (hosts split in {world} ranges; the (target, origin) cells of the "
                                                                              "resampling are pushed into peer memory over NVLink)" if sharded else
                                                                              f"{world} compartments, 1 per GPU, migrants pushed into peer memory over NVLink by the kernels (0.9 stay / 0.1 migrate uniformly)"),
                          "l2_policy": "working set (per-infectant arrays, 32 B x N) larger than L2; no flush" if N * 32 > 126e6 else "working set fits L2 (latency-bound config); no flush",
                          "host_sum": "per-host f64 sums by atomics (device-resident loop); reference-ordered CSR on the trait-level path"},
               "gpu_launches": int(launches), "wall_s_timed_region": t_wall, "clocks": clk, "roofline": roofline, "kernels": kernels,
               "e2e": e2e}
        if compaction is not None:
            out["compaction"] = compaction
        if exchange_checked is not None:
            out["exchange_checked"] = exchange_checked
    if runner is not None and hasattr(runner, "close"):
        runner.close()
    sim.close()
    if rank == 0 and world == 1 and not args.no_extra:
        out["long_run"] = long_run(local_rank)
        out["extra"] = extra_configs(local_rank)
    if world > 1 and not args.no_extra and not sharded:
        blk = extra_sharded_k5(rank, world, local_rank, dist)
        if rank == 0:
            out["extra"] = {"k5_sharded": blk}
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(args, wl)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return out


def extra_sharded_k5(rank, world, device, dist):
    """BASELINE configs[4] shape, bounded: ONE K5 compartment (30 kb genome, recombination, R0 = 100, dilution 0.02) of 10^7
    infectants sharded over all ranks, device-resident, a dozen generations."""
    import torch
    from virolution_b200.distributed import PeerExchange, sharded_simulation
    from virolution_b200.workloads import workload
    wl = workload("k5_sars", 10_000_000)
    par = wl["parameters"]
    sim = sharded_simulation(wl["wildtype"], par, wl["specs"], rank, world, seed=77, device=device, n0=wl["n0"])
    px = PeerExchange(sim, rank, world, sharded=True, shard_seed=77)
    warm, gens = 6, 12
    px.run(warm)
    sim.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sim.infectant_generations(reset=True)
    sim.timer_start()
    px.run(gens)
    ms = sim.timer_stop()
    work = sim.infectant_generations()
    t = torch.tensor([float(work), ms, float(sim.population_len())], dtype=torch.float64, device="cuda")
    tsum, tmax = t.clone(), t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    px.close()
    sim.close()
    return {"workload": wl["description"] + f", ONE compartment sharded over {world} GPUs", "ms_per_generation": round(float(tmax[1]) / gens, 5),
            "value": float(tsum[0]) / (float(tmax[1]) * 1e-3), "unit": UNIT, "generations_timed": gens, "warmup": warm, "scaling": "strong",
            "population_total_at_end": int(tsum[2]), "expected_total": int(par.max_population),
            "bound": "general kernels (recombination needs the host map) + exchange of (world-1)/world of the population per generation"}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def oracle_rate(wl, N, threads, warm, steps, fully_parallel=False, compartments=1, T=None):
    """infectant-generations/s of the oracle (CPU restatement of the reference loop) after `warm` generations."""
    from oracle.oracle import OracleWorld
    from virolution_b200.workloads import workload
    w = OracleWorld(wl["wildtype"], wl["parameters"], wl["specs"], n_compartments=compartments, seed=7, n_threads=threads)
    w.set_mode(no_logs=True, fully_parallel=fully_parallel)  # the reference keeps no draw logs
    for c in range(compartments):
        w.set_population_wildtype(N, c)

    def gen():
        if compartments == 1:
            w.next_generation()
        else:
            w.step(T, 0)  # parallel-build rule, compartments over the threads like rayon (src/runner.rs:388-416)
    for _ in range(warm):
        gen()
    ig0 = sum(w.infectant_generations(c) for c in range(compartments))
    t0 = time.perf_counter()
    for _ in range(steps):
        gen()
    dt = time.perf_counter() - t0
    ig = sum(w.infectant_generations(c) for c in range(compartments)) - ig0
    return ig / dt, dt, w.num_threads()


def cpu_baseline(args, wl):
    """The oracle (kind "port") on a bounded sample of the same workload: a 10^6-infectant slice (the work per infectant does
    not depend on N; the smaller population is kinder to the CPU's caches), advanced 10 generations before 4 are timed."""
    from virolution_b200.workloads import workload
    N = min(wl["n0"], 1_000_000)
    wls = workload(args.workload, N)
    cores = host_cores()
    one, _, _ = oracle_rate(wls, N, 1, 10, 4)
    allc, dt, used = oracle_rate(wls, N, cores, 10, 4) if cores > 1 else (one, 0.0, 1)
    full, _, _ = oracle_rate(wls, N, cores, 10, 4, fully_parallel=True) if cores > 1 else (one, 0.0, 1)
    best, threads = (allc, used) if allc >= one else (one, 1)
    return {"value": best, "unit": UNIT, "cores": threads, "kind": "port", "host_cores": cores,
            "sample": f"4 generations of a {N}-infectant slice of the workload after 10 warm-up generations; oracle/vl_oracle.c parallel where the "
                      f"reference's rayon build is (per-host mutation; infect / replicate / sample serial, SURVEY.md D12), no draw logs",
            "one_thread": one, "all_cores_rayon_equivalent": allc,
            "fully_parallel_cpu": {"value": full, "cores": cores, "note": "separately labelled variant (BASELINE.md §3.1): infect, replicate and the "
                                   "sample draws over all cores too — more than the reference's rayon build does"}}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm for this path.  The Rust crate cannot be built in this image
    (no cargo/rustc, nightly-only), so this is the oracle port with every host core, parallel where the reference's rayon
    build is: compartments (src/runner.rs:388-416) and per-host mutation (src/simulation.rs:389-408).  At --gpus N it runs
    the arm's N compartments with the arm's transfer matrix (parallel-build rule)."""
    if rank != 0:
        return None
    from virolution_b200.workloads import workload
    from virolution_b200.distributed import migration_matrix
    wl = workload(args.workload, args.n or None)
    N = wl["n0"]
    cores = host_cores()
    steps = max(1, min(args.steps, 5 if world == 1 else 2))
    warm = max(1, min(args.warmup, 2 if world == 1 else 1))
    T = migration_matrix(world, 0.1) if world > 1 else None
    value, dt, used = oracle_rate(wl, N, cores, warm, steps, compartments=world, T=T)
    sample = (f"{steps} generations of the workload (N={N} per compartment, {world} compartment(s)) after {warm} warm-up, {used} OpenMP thread(s) of "
              f"{cores} host cores; parallel over compartments and over hosts inside mutate like the reference's rayon build, infect / replicate / "
              f"sample serial per compartment like the reference; no draw logs")
    return {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": wl["description"] + ("" if world == 1 else f" per compartment x {world} compartments"), "compartments": world},
            "gpu_launches": 0, "host_cores": cores,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = run_reference(args, rank, world) if args.impl == "reference" else run_b200(args, rank, world)
    if rank == 0 and out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
