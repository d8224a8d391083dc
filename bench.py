#!/usr/bin/env python
"""bench.py — infectant-generations/sec of the B200 population-update path (BASELINE.json metric).

A "step" is one generation (infect -> mutate -> fitness -> replicate -> subsample/transmit) of the workload.
    python bench.py --gpus 1 --steps K --warmup W            our arm (CUDA path through the C ABI)
    python bench.py --impl reference ...                     the reference's CPU algorithm (oracle port, host cores)
    torchrun ... bench.py --gpus N ...                       one compartment per GPU + NCCL migrant exchange (weak scaling)

Lines of the JSON object:
  value         device-resident loop (vl_run), timed with CUDA events on the library's stream, max over ranks
  e2e           the trait-level calls a `impl Simulation for GpuSimulation` makes, with HOST offspring vectors:
                replicate_infectants -> Vec<usize> (D2H), subsample_population(&offspring) (H2D), set_population
  roofline      dominant kernel: algorithmic bytes per launch / average CUDA-event duration vs measured HBM peak
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

# Algorithmic bytes per unit for each kernel of the fused generation: u32 ids/hosts/offspring,
# f64 fitness, 16-byte bucket records, 8-byte CSR records.  "per host" terms are multiplied by H/N (= 1 for the K configs).
# Unit = one infectant at generation start, except k_mutate (one mutated infectant, measured separately below).
ALGO_BYTES = {
    "k_replicate_fused<true>": 4 + 8 + 8 + 4,   # offsets[h] (per host), rec[p], gather raw fitness, write offspring
    "k_replicate_fused<false>": 4 + 8 + 8 + 4,
    "(k_infect_bucket<false, false, 1024>)": 4 + 16,  # read hap_id, append (index, haplotype, host) to the host bucket
    "(k_infect_bucket<true, true, 1024>)": 4 + 4 + 16,  # ... + host_id[] kept for the trait-level phases
    "(k_bucket_sort<1024, BSORT_STAGE_RECS>)": 16 + 8 + 4,  # read bucket record, write rec[p], offsets[h] (per host)
    "k_resample": 4 + 8 + 4,                    # offspring[p], rec[p] (tile staging), write next hap_id
    "k_fitness": 8 + 8 + 8, "k_host_sum": 8 + 8 + 8 + 8, "k_offspring<true>": 8 + 8 + 4, "k_offspring<false>": 8 + 8 + 4,
    "k_infect<true, true>": 4 + 4 + 8, "k_scatter": 4 + 4 + 4 + 4 + 8,
}
# the kernels SURVEY.md §8d's 72 B figure covers (fitness gather + per-host sum + offspring + scan/plan + resample)
FITNESS_REPLICATION = ("k_replicate_fused<true>", "k_replicate_fused<false>", "k_fitness", "k_host_sum", "k_host_sum_medium",
                       "k_host_sum_heavy", "k_offspring<true>", "k_offspring<false>", "k_plan_resample", "k_resample")


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
    ap.add_argument("--profile-steps", type=int, default=10)
    ap.add_argument("--cpu-steps", type=int, default=0, help="generations of the cpu_baseline sample (0 = auto)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
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
    runner = None
    if sharded:
        runner = PeerExchange(sim, rank, world, sharded=True, shard_seed=2026)
        args.no_e2e = True  # the trait-level calls address one handle; a sharded compartment is driven through the exchange only
    elif world > 1:
        from virolution_b200.distributed import CompartmentExchange, GpuCompartment, PeerExchange
        collective = CompartmentExchange(GpuCompartment(sim), rank, world, migration=0.1)  # e2e leg: host vectors + NCCL all-to-all-v
        # device-resident loop: migrants are pushed by the kernels into peer memory (VL_EXCHANGE=nccl keeps the collective path)
        runner = collective if os.environ.get("VL_EXCHANGE") == "nccl" else PeerExchange(sim, rank, world, migration=0.1)

    def advance(k):
        if runner is None:
            sim.run(k)
        else:
            for _ in range(k):
                runner.step()

    def barrier():
        sim.synchronize()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    if runner is not None and os.environ.get("VL_TRACE") and hasattr(runner, "transmit"):
        runner.trace = {}
    # ---- warm-up (untimed): also brings haplotype diversity towards its quasi-steady state
    advance(max(args.warmup, 3))
    barrier()
    sim.infectant_generations(reset=True)
    if runner is not None and runner.trace is not None:
        runner.trace = {}
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
    if runner is not None and runner.trace is not None and rank == 0:
        print("exchange host seconds per phase over the timed steps:", {k: round(v, 4) for k, v in runner.trace.items()}, file=sys.stderr)
    ig = sim.infectant_generations(reset=True)
    ct1 = sim.get_counters()
    launches = ct1["kernels_launched"] - l0
    if dist is not None:
        import torch
        t = torch.tensor([ms, float(ig)], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, ig = float(tmax[0]), int(tsum[1])
    value = ig / (ms / 1e3)

    # ---- compaction of the record tables (the stand-in for Rc/Arc drops): how many ran inside the timed region, what one
    # costs (forced once here, on the table as the timed region left it, timed with CUDA events) and how often a long run
    # needs one, so that the amortised share can be read off whether or not the timed window happened to contain one
    compaction = None
    if rank == 0 and runner is None:
        cap = sim.get_capacities()
        runs = ct1["gc_runs"] - ct0["gc_runs"]
        rec_per_gen = (ct1["haplotype_records"] - ct0["haplotype_records"]) / args.steps if runs == 0 else None
        sim.timer_start()
        sim.collect_garbage()
        gc_ms = sim.timer_stop()
        live = sim.get_counters()["haplotype_records"]
        compaction = {"runs_in_timed_region": int(runs), "ms_per_run": round(gc_ms, 4), "records_before": ct1["haplotype_records"],
                      "live_records_after": int(live), "record_capacity": cap["haplotype_records"],
                      "trigger": "record table or arena half full"}
        if rec_per_gen and rec_per_gen > 0:
            period = max(1.0, (cap["haplotype_records"] / 2 - live) / rec_per_gen)
            # a run at the trigger point passes over capacity/2 records: scale the measured run to that table size
            at_trigger = gc_ms * max(1.0, (cap["haplotype_records"] / 2) / max(ct1["haplotype_records"], 1))
            compaction.update({"new_records_per_generation": round(rec_per_gen), "period_generations": round(period, 1),
                               "ms_per_run_at_trigger": round(at_trigger, 4), "amortised_ms_per_step": round(at_trigger / period, 5),
                               "value_with_amortised_compaction": value * (ms / args.steps) / (ms / args.steps + at_trigger / period)})

    # ---- per-kernel CUDA-event durations over the same loop (separate pass: events perturb the step time)
    roofline, kernels = None, {}
    # (every rank runs the generations — the exchange is collective — rank 0 alone records events)
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
        # k_mutate works per new haplotype record, not per infectant: its algorithmic bytes come from the records and arena
        # words the profiled generations created (item 16 + parent meta 16 + parent raw 8 + table row 16 read; meta 16 + raw 8 +
        # mapped 8 + hap_id 4 written; every arena word of the new record is read from the parent and written once)
        mutate_bytes = None
        if c1["gc_runs"] == c0["gc_runs"] and "k_mutate" in rep:
            d_rec, d_words = c1["haplotype_records"] - c0["haplotype_records"], c1["arena_words"] - c0["arena_words"]
            mutate_bytes = (92.0 * d_rec + 8.0 * d_words) / rep["k_mutate"][0]
        peak, peak_src = measured_peak()
        tot_ms = sum(v[1] for v in rep.values())
        for name, (cnt, tms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
            per_launch_ms = tms / cnt
            ab = ALGO_BYTES.get(name)
            kernels[name] = {"launches_per_step": cnt / args.profile_steps, "ms_per_launch": round(per_launch_ms, 5),
                             "share": round(tms / tot_ms, 4)}
            if ab is not None:
                kernels[name]["GBps"] = round(ab * n_now / (per_launch_ms * 1e-3) / 1e9, 1)
            elif name == "k_mutate" and mutate_bytes is not None:
                kernels[name]["GBps"] = round(mutate_bytes / (per_launch_ms * 1e-3) / 1e9, 1)
                kernels[name]["algorithmic_bytes_per_launch"] = round(mutate_bytes)
        launch_bytes = {k: ALGO_BYTES[k] * n_now for k in rep if k in ALGO_BYTES}
        if mutate_bytes is not None:
            launch_bytes["k_mutate"] = mutate_bytes
        dom = max(launch_bytes, key=lambda k: rep[k][1])  # the kernel that takes most of the step
        dcnt, dms = rep[dom]
        achieved = launch_bytes[dom] / (dms / dcnt * 1e-3) / 1e9
        fr_ms = sum(rep[k][1] / args.profile_steps for k in FITNESS_REPLICATION if k in rep)
        fr_bytes = 72.0 * n_now  # fitness + replication + scan/resample
        traffic = None
        try:  # DRAM bytes per launch of the same kernel from the committed ncu --set full capture (same workload size only)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            if tj.get("n") == n_now and dom in tj:
                traffic = tj[dom]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": round(launch_bytes[dom]), "infectants_per_launch": n_now,
                    "by_kernel": {k: {"ms_per_launch": round(rep[k][1] / rep[k][0], 5),
                                      "achieved": round(launch_bytes[k] / (rep[k][1] / rep[k][0] * 1e-3) / 1e9, 1),
                                      "frac": round(launch_bytes[k] / (rep[k][1] / rep[k][0] * 1e-3) / 1e9 / peak, 4)}
                                  for k in sorted(launch_bytes, key=lambda k: -rep[k][1])[:4]},
                    "fitness_replication_group": {"kernels": [k for k in FITNESS_REPLICATION if k in rep],
                                                  "ms_per_generation": round(fr_ms, 5), "algorithmic_bytes_per_infectant": 72,
                                                  "achieved": round(fr_bytes / (fr_ms * 1e-3) / 1e9, 1),
                                                  "frac": round(fr_bytes / (fr_ms * 1e-3) / 1e9 / peak, 4)}}

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
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "impl": "b200",
               "config": {"workload": wl["description"] + ("" if world == 1 or sharded else f" per compartment x {world} compartments"), "compartments": 1 if sharded else world,
                          "parallelism": "1 compartment" if world == 1 else (f"1 compartment sharded over {world} GPUs (hosts split in {world} ranges; the (target, origin) cells of the "
                                                                              "resampling are pushed into peer memory over NVLink)" if sharded else
                                                                              f"{world} compartments, 1 per GPU, migrants pushed into peer memory over NVLink by the kernels (0.9 stay / 0.1 migrate uniformly)"),
                          "l2_policy": "working set (per-infectant arrays, 32 B x N) larger than L2; no flush" if N * 32 > 126e6 else "working set fits L2 (latency-bound config); no flush",
                          "host_sum": "reference-ordered CSR"},
               "gpu_launches": int(launches), "wall_s_timed_region": t_wall, "clocks": clk, "roofline": roofline, "kernels": kernels,
               "e2e": e2e}
        if compaction is not None:
            out["compaction"] = compaction
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(args, wl, threads=1)
    if runner is not None and hasattr(runner, "close"):
        runner.close()
    sim.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return out


def cpu_baseline(args, wl, threads):
    """The oracle (CPU restatement of the reference loop; kind "port") on a bounded sample of the same workload."""
    from oracle.oracle import OracleWorld
    N = wl["n0"]
    steps = args.cpu_steps or max(2, min(10, int(2.0e7 // max(N, 1)) + 2))
    w = OracleWorld(wl["wildtype"], wl["parameters"], wl["specs"], n_compartments=1, seed=7, n_threads=threads)
    w.set_population_wildtype(N)
    w.next_generation()  # first touch of every buffer
    ig0 = w.infectant_generations()
    t0 = time.perf_counter()
    for _ in range(steps):
        w.next_generation()
    dt = time.perf_counter() - t0
    ig = w.infectant_generations() - ig0
    return {"value": ig / dt, "unit": UNIT, "cores": w.num_threads(), "kind": "port",
            "sample": f"{steps} generations of the same workload (N={N}) after 1 warm-up generation; oracle/vl_oracle.c, "
                      f"{w.num_threads()} thread(s), parallel only where the reference's rayon build is (per-host mutation)",
            "seconds": dt}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm for this path.  The Rust crate cannot be built in this image
    (no cargo/rustc, nightly-only), so this is the oracle port, parallel where the reference's rayon build is (per-host
    mutation; infect / replicate / sample are serial there too).  The thread count that is faster on this
    box is used (the parallel section pays for its locks on node allocation, like the reference's DashMap + Arc)."""
    if rank != 0:
        return None
    from virolution_b200.workloads import workload
    from oracle.oracle import OracleWorld
    wl = workload(args.workload, args.n or None)
    N = wl["n0"]

    def timed(threads, steps, warm):
        w = OracleWorld(wl["wildtype"], wl["parameters"], wl["specs"], n_compartments=1, seed=7, n_threads=threads)
        w.set_population_wildtype(N)
        for _ in range(warm):
            w.next_generation()
        ig0 = w.infectant_generations()
        t0 = time.perf_counter()
        for _ in range(steps):
            w.next_generation()
        dt = time.perf_counter() - t0
        return (w.infectant_generations() - ig0) / dt, dt, w.num_threads()

    all_threads = os.cpu_count() or 1
    probe = {t: timed(t, 1, 1)[0] for t in sorted({1, all_threads})}
    threads = max(probe, key=probe.get)
    steps = max(1, min(args.steps, 6))
    warm = max(1, min(args.warmup, 2))
    value, dt, used = timed(threads, steps, warm)
    sample = (f"{steps} generations of the workload (N={N}) after {warm} warm-up, {used} OpenMP thread(s) "
              f"(probe: " + ", ".join(f"{t} thread(s) {v / 1e6:.2f} M/s" for t, v in probe.items()) + "); parallel only where the "
              "reference's rayon build is (per-host mutation; infect/replicate/sample are serial like the reference)")
    return {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference", "config": {"workload": wl["description"], "compartments": 1},
            "gpu_launches": 0,
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
