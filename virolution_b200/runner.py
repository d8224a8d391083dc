"""Host-side mirror of `Runner` (src/runner.rs) for the GPU path: the generation loop of the parallel build
(src/runner.rs:336-423) over `GpuSimulation` handles, with the outputs the reference's integration tests look for —
`<provider>_table.npy`, `virolution.log`, `samples/sample_<generation>_<compartment>.{fasta,csv}`,
`samples/barcodes.csv`, `final.<i>.csv`
(src/main.rs:63-154, src/runner.rs:137-152,230-240, src/readwrite/sample_writer.rs:29-61,243-263).

Everything that costs time stays on the device: per generation one enqueue per compartment (infect, mutate, replicate and
the transmission block, migrants written by the kernels into the other compartments' buffers), populations are only
materialised on the host at sample events and at the end.

    python -m virolution_b200.runner --settings tests/singlehost/singlehost.yaml --sequence tests/singlehost/ref.fasta \\
        --generations 200 --n-compartments 2 --outdir out
"""
from __future__ import annotations

import argparse
import csv
import os
from typing import List, Optional, Sequence

import numpy as np

from .config import load_parameters, load_settings, read_fasta_wildtype
from .simulation import CompartmentGroup, GpuSimulation, step

LETTERS = "ATCG"


def block_id(haplotype_id: int) -> str:
    """Stand-in for `HaplotypeRef::get_block_id` (the reference encodes the node's heap address as letters,
    src/references/desc.rs): a stable label per device record."""
    s, x = "", int(haplotype_id)
    while True:
        s = chr(ord("a") + x % 26) + s
        x //= 26
        if x == 0:
            return s


def haplotype_string(wildtype: np.ndarray, positions: Sequence[int], bases: Sequence[int]) -> str:
    """Haplotype::get_string (src/core/haplotype.rs:437-453)."""
    text = ";".join(f"{int(p)}:{LETTERS[int(wildtype[int(p)])]}->{LETTERS[int(b)]}" for p, b in zip(positions, bases))
    return text or "wt"


class GpuRunner:
    """Runner::new + Runner::start (src/runner.rs:43-112,278-423), parallel-build transmission rule."""

    def __init__(self, settings: str, sequence, generations: int = 200, n_compartments: int = 1, outdir: str = ".",
                 initial_population_size: Optional[int] = None, seed: int = 0, device: int = 0, table_seed: int = 0,
                 logfile: str = "virolution.log", name: str = "simulation", sampling_format: str = "fasta"):
        self.wildtype = read_fasta_wildtype(sequence) if isinstance(sequence, str) else np.asarray(sequence, dtype=np.uint8)
        self.parameters, self.host_models, self.schedule = load_settings(settings)
        if not self.schedule.check_transfer_table_sizes(n_compartments):
            # Runner::new (src/runner.rs:78-85)
            raise ValueError("Transfer tables are too small for the number of compartments.")
        self.generations, self.n_compartments, self.outdir = generations, n_compartments, outdir
        self.settings_dir = os.path.dirname(os.path.abspath(settings))
        if sampling_format not in ("fasta", "csv"):
            raise ValueError("Unknown sampling format.")  # src/runner.rs:329-332
        self.name, self.sampling_format = name, sampling_format
        os.makedirs(outdir, exist_ok=True)
        # SampleWriter::init_barcode_file (src/readwrite/sample_writer.rs:75-86): <outdir>/samples/barcodes.csv, truncated
        self.samples_dir = os.path.join(outdir, "samples")
        os.makedirs(self.samples_dir, exist_ok=True)
        with open(os.path.join(self.samples_dir, "barcodes.csv"), "w", newline="") as f:
            csv.writer(f, lineterminator="\n").writerow(["barcode", "experiment", "time", "replicate", "compartment"])
        par = self.parameters[0]
        base = os.path.dirname(os.path.abspath(settings))
        # the reference's table generator is unseeded; here the tables come from `table_seed`
        self.specs = self.host_models[0].make_definitions(par, self.wildtype, base_path=base, rng=np.random.default_rng(table_seed))
        self._write_fitness_tables()
        n0 = par.max_population if initial_population_size is None else initial_population_size
        self.sims: List[GpuSimulation] = []
        for c in range(n_compartments):
            sim = GpuSimulation(self.wildtype, par, self.specs, seed=seed, device=device, compartment_id=c)
            sim.set_population_wildtype(n0)  # population![wildtype; n] (src/runner.rs:263)
            self.sims.append(sim)
        self.log = open(os.path.join(outdir, logfile), "w")
        self.infectant_generations = 0

    def _write_fitness_tables(self):
        # AttributeProvider::write (src/providers/fitness.rs:66-71,104-113): <name>_table.npy (L x 4 f64) [+ epistasis]
        for s, spec in enumerate(self.specs):
            for kind, prov in (("reproductive", spec.reproductive), ("infective", spec.infective)):
                if prov is None or prov.table is None:
                    continue
                name = f"{kind}_fitness_{s}"
                np.save(os.path.join(self.outdir, f"{name}_table.npy"), np.asarray(prov.table, dtype=np.float64).reshape(-1, 4))
                if prov.epistasis is not None:
                    epi = np.asarray(prov.epistasis)
                    if not epi.dtype.names:
                        rec = np.zeros(len(epi), dtype=[("pos1", "<u8"), ("base1", "u1"), ("pos2", "<u8"), ("base2", "u1"), ("value", "<f8")])
                        for k, f in enumerate(("pos1", "base1", "pos2", "base2", "value")):
                            rec[f] = epi[:, k]
                        epi = rec
                    np.save(os.path.join(self.outdir, f"{name}_epistasis_table.npy"), epi)

    # ------------------------------------------------------------------ sample / final writers
    def _rows(self, sim: GpuSimulation, indices: Optional[np.ndarray] = None):
        """(record id, haplotype string, count) per haplotype record among the chosen infectants (all if None); the
        reference counts by reference identity, not by string."""
        ids = sim.get_haplotype_ids()
        off, pos, base = sim.export_mutations()
        chosen = np.arange(ids.size) if indices is None else np.asarray(indices, dtype=np.int64)
        uniq, first, counts = np.unique(ids[chosen], return_index=True, return_counts=True)
        rows = []
        for k in np.argsort(first):
            i = int(chosen[first[k]])
            a, b = int(off[i]), int(off[i + 1])
            rows.append((int(uniq[k]), haplotype_string(self.wildtype, pos[a:b], base[a:b]), int(counts[k])))
        return rows

    def write_sample(self, generation: int, sample_size: int, rng: np.random.Generator):
        """SampleWriter::sample_from_simulations (src/readwrite/sample_writer.rs:29-61): `choose_multiple` = a uniform
        sample without replacement of min(size, n) infectants (src/core/population.rs:353-370), written as
        samples/sample_<generation>_<compartment>.{fasta,csv} plus one line of samples/barcodes.csv (src/barcode.rs)."""
        for c, sim in enumerate(self.sims):
            n = sim.population_len()
            idx = rng.choice(n, size=min(sample_size, n), replace=False) if n else np.zeros(0, dtype=np.int64)
            # The reference's compartments share ONE generation attribute and each of them increments it
            # (src/runner.rs:72,271,382-384, src/providers/generation.rs:31-33): after i loop iterations it reads
            # i * n_compartments, and that is what file names, rows and barcodes carry.  Every handle here
            # counts its own generations, so the label is rebuilt from it.
            g = sim.get_generation() * len(self.sims)
            barcode = f"sample_{g}_{c}"
            if self.sampling_format == "csv":
                # CsvSampleWriter::write (:228-266): one row per distinct haplotype of the sample
                with open(os.path.join(self.samples_dir, f"{barcode}.csv"), "w", newline="") as f:
                    w = csv.writer(f, lineterminator="\n")
                    w.writerow(["block_id", "compartment_id", "generation", "haplotype", "count"])
                    for hid, text, count in self._rows(sim, idx):
                        w.writerow([block_id(hid), c, g, text, count])
            else:
                # FastaSampleWriter::write (:161-208): one record per sampled infectant, the whole sequence on one line
                # (Haplotype::get_sequence, src/core/haplotype.rs:410-419: wildtype overwritten by get_mutations())
                ids = sim.get_haplotype_ids()
                off, pos, base = sim.export_mutations()
                letters = np.frombuffer(LETTERS.encode(), dtype=np.uint8)
                wt = letters[self.wildtype]
                with open(os.path.join(self.samples_dir, f"{barcode}.fasta"), "wb") as f:
                    for k, i in enumerate(idx):
                        seq = wt.copy()
                        a, b = int(off[i]), int(off[i + 1])
                        seq[pos[a:b]] = letters[base[a:b]]
                        f.write(f">sequence_id={k};block_id={block_id(int(ids[i]))};compartment_id={c};generation={g}\n".encode())
                        f.write(seq.tobytes() + b"\n")
            with open(os.path.join(self.samples_dir, "barcodes.csv"), "a", newline="") as f:
                csv.writer(f, lineterminator="\n").writerow([barcode, self.name, g, 0, c])

    def finish(self):
        """Runner::finish (src/runner.rs:137-152): final.<compartment>.csv with header haplotype,count."""
        for c, sim in enumerate(self.sims):
            with open(os.path.join(self.outdir, f"final.{c}.csv"), "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(["haplotype", "count"])
                for _, text, count in self._rows(sim):
                    w.writerow([text, count])
        self.log.close()

    # ------------------------------------------------------------------ schedule events other than transmission / sample
    def adjust_settings(self, generation: int):
        """`settings` event (src/runner.rs:373-379, src/config/schedule.rs:210-222): the value is the path of a YAML file
        holding one parameter block; every compartment gets it.  A file that cannot be read is reported and skipped, as
        `get_settings` does (`.ok()`)."""
        value = self.schedule.get_event_value("settings", generation)
        if value is None:
            return None
        path = value if os.path.isabs(value) or os.path.exists(value) else os.path.join(self.settings_dir, value)
        try:
            parameters = load_parameters(path)
        except Exception as err:
            self.log.write(f"Error reading settings file `{value}`: {err!r}\n")
            return None
        self.log.write(f"Adjust settings to:\n{parameters}\n")
        for sim in self.sims:
            sim.set_parameters(parameters)
        self._drop_group()  # the exchange buffers were sized for the old max_population
        return parameters

    # ------------------------------------------------------------------ the transmission block of several compartments
    def _drop_group(self):
        if getattr(self, "_group", None) is not None:
            self._group.close()
        self._group = None

    def _transmit(self, T):
        """One generation of every compartment.  Several compartments are stepped through the peer exchange between the
        compartments of this process (`CompartmentGroup`: nothing is synchronised or copied to the host), sized for the
        entry-wise maximum of the schedule's matrices; a single compartment through `vl_step`."""
        C = self.n_compartments
        if C < 2:
            step(self.sims, T, rule=0)
            return
        if getattr(self, "_group", None) is None:
            cap = np.array(T, dtype=np.float64)
            for M in self.schedule.transfers.values():
                if M.shape[0] >= C and M.shape[1] >= C:
                    cap = np.maximum(cap, M[:C, :C])
            self._group = CompartmentGroup(self.sims, cap)
        self._group.step(T)

    def collect_stats(self, generation: int):
        """`stats` event (src/runner.rs:634-673; only the serial build of the reference has it): comma-separated list of
        diversity / average-fitness / distance, computed on the offspring-producing population of this generation with
        device reductions (src/stats/population.rs:13-75) and written to the log in the reference's wording."""
        stats = self.schedule.get_event_value("stats", generation)
        out = {}
        for stat in (stats.split(",") if stats else []):
            if stat == "diversity":
                out["frequencies"] = [sim.frequencies().reshape(-1, 4) for sim in self.sims]
                for idx, freq in enumerate(out["frequencies"]):
                    self.log.write(f"({idx})frequencies={np.array2string(freq, precision=3, threshold=64)}\n")
            elif stat == "average-fitness":
                out["mean"] = []
                for idx, sim in enumerate(self.sims):
                    self.log.write(f"compartment={idx}\n")
                    for s, spec in enumerate(self.specs):
                        for which, (name, prov) in enumerate((("reproductive_fitness", spec.reproductive), ("infective_fitness", spec.infective))):
                            if prov is None:
                                continue
                            mean = sim.mean_fitness(s, which)
                            out["mean"].append((idx, s, name, mean))
                            self.log.write(f"mean({name}_{s})={mean}\n")
            elif stat == "distance":
                out["distance"] = [a.distance(b) for a in self.sims for b in self.sims]
                self.log.write(f"distance={out['distance']}\n")
        return out

    # ------------------------------------------------------------------ the loop
    def start(self, sample_seed: int = 0):
        rng = np.random.default_rng(sample_seed)
        C = self.n_compartments
        for generation in range(self.generations + 1):
            sizes = [s.population_len() for s in self.sims]
            self.log.write(f"\n        generation={generation}\n        population_sizes={sizes}\n")  # src/runner.rs:345-349
            sample_size = self.schedule.get_sample_size(generation)
            if sample_size > 0:
                self.write_sample(generation, sample_size, rng)
            if generation == self.generations:
                break
            self.adjust_settings(generation)
            self.collect_stats(generation)
            self.infectant_generations += sum(sizes)
            # increment_generation, infect, mutate_infectants, replicate_infectants and the transmission block of every
            # compartment (src/runner.rs:382-422) as one call; T = schedule.get_transfer_matrix(generation), row = target
            T = self.schedule.get_transfer_matrix(generation)[:C, :C]
            self._transmit(T)
        self.finish()
        return self

    def close(self):
        self._drop_group()
        for s in self.sims:
            s.close()
        self.sims = []


def main(argv=None):
    ap = argparse.ArgumentParser(description="virolution's generation loop on B200 (mirror of src/args.rs)")
    ap.add_argument("--settings", required=True)
    ap.add_argument("--sequence", required=True)
    ap.add_argument("--generations", type=int, default=200)
    ap.add_argument("--n-compartments", type=int, default=1)
    ap.add_argument("--outdir", default="./")
    ap.add_argument("--initial-population-size", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0, help="the reference has no seed (rand::rng()); this one is ours")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--name", default="simulation")
    ap.add_argument("--sampling-format", default="fasta", choices=["fasta", "csv"])
    ap.add_argument("--logfile", default="virolution.log")
    a = ap.parse_args(argv)
    r = GpuRunner(a.settings, a.sequence, a.generations, a.n_compartments, a.outdir, a.initial_population_size, a.seed, a.device,
                  logfile=a.logfile, name=a.name, sampling_format=a.sampling_format)
    r.start()
    print(f"{r.infectant_generations} infectant-generations; outputs in {a.outdir}")
    r.close()


if __name__ == "__main__":
    main()
