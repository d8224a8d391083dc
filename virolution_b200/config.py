"""Host-side configuration resolution: what crosses the C ABI is numbers, not YAML.

Restates the pieces of the reference that turn a settings file into the resolved inputs of the hot path:
  * Nucleotide index order A,T,C,G = 0..3                      (src/encoding.rs:31-73)
  * FASTA wildtype loading                                      (src/readwrite/haplotype.rs:24-39)
  * FitnessModel -> table generation / .npy loading             (src/init/fitness.rs:109-215)
  * HostModel::{SingleHost,MultiHost}::make_definitions          (src/config/parameters.rs:112-145)
  * Settings / Schedule (transmission matrix per generation)     (src/config/settings.rs, schedule.rs:196-266)
The reference's table generator is unseeded (rand::rng()); here it takes an explicit numpy Generator.
"""
from __future__ import annotations

import ast
import os
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from .abi import FitnessProvider, HostSpec, Parameters, UtilityFunction

_DECODE = {}
for _chars, _idx in (("Aa0", 0), ("Tt1", 1), ("Cc2", 2), ("Gg3", 3)):
    for _ch in _chars:
        _DECODE[ord(_ch)] = _idx
_ENCODE = "ATCG"


def decode_sequence(seq: str) -> np.ndarray:
    """Nucleotide::decode over a string (src/encoding.rs:34-43); invalid symbols raise like the panic."""
    out = np.empty(len(seq), dtype=np.uint8)
    for i, ch in enumerate(seq.encode()):
        if ch not in _DECODE:
            raise ValueError(f"Invalid nucleotide symbol: {ch}")
        out[i] = _DECODE[ch]
    return out


def encode_sequence(idx: np.ndarray) -> str:
    return "".join(_ENCODE[int(i)] for i in idx)


def read_fasta_wildtype(path: str) -> np.ndarray:
    """First record of a FASTA file as base indices (src/readwrite/haplotype.rs:24-39)."""
    seq: List[str] = []
    started = False
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                if started:
                    break
                started = True
                continue
            seq.append(line)
    return decode_sequence("".join(seq))


# ------------------------------------------------------------------------- fitness models (init/fitness.rs)

@dataclass
class FitnessModel:
    """FitnessModel{distribution, utility} (src/init/fitness.rs:15-60)."""
    distribution: str = "Neutral"  # Neutral | Exponential | Lognormal | File | Epistatic
    params: Dict = field(default_factory=dict)
    utility: UtilityFunction = field(default_factory=UtilityFunction)


def exponential_table(wildtype: np.ndarray, weights=(0.29, 0.51, 0.2, 0.0), lambda_beneficial=0.03,
                      lambda_deleterious=0.21, rng: Optional[np.random.Generator] = None) -> np.ndarray:
    """ExponentialParameters::create_table (src/init/fitness.rs:109-147): wildtype entries 1; every other entry
    drawn in row-major order from {beneficial 1+Exp(scale lambda_b), deleterious max(1-Exp(scale lambda_d),0),
    lethal 0, neutral 1} with the category weights (beneficial, deleterious, lethal, neutral)."""
    rng = rng or np.random.default_rng(0)
    L = len(wildtype)
    table = np.full((L, 4), -1.0)
    table[np.arange(L), wildtype] = 1.0
    mask = table < 0
    n = int(mask.sum())
    w = np.asarray(weights, dtype=np.float64)
    cat = rng.choice(4, size=n, p=w / w.sum())
    vals = np.ones(n)
    ben = cat == 0
    dele = cat == 1
    vals[ben] = 1.0 + rng.exponential(lambda_beneficial, size=int(ben.sum()))
    vals[dele] = np.maximum(1.0 - rng.exponential(lambda_deleterious, size=int(dele.sum())), 0.0)
    vals[cat == 2] = 0.0
    table[mask] = vals
    return table


def lognormal_table(wildtype: np.ndarray, lethal=0.0, mu=0.0, sigma=1.0,
                    rng: Optional[np.random.Generator] = None) -> np.ndarray:
    """LognormalParameters::create_table (src/init/fitness.rs:151-178)."""
    rng = rng or np.random.default_rng(0)
    L = len(wildtype)
    table = np.full((L, 4), -1.0)
    table[np.arange(L), wildtype] = 1.0
    mask = table < 0
    n = int(mask.sum())
    vals = rng.lognormal(mu, sigma, size=n)
    vals[rng.random(n) < lethal] = 0.0
    table[mask] = vals
    return table


def load_epistasis_npy(path: str) -> np.ndarray:
    """EpiEntry records (src/core/fitness/epistasis.rs:15-21; loader src/init/fitness.rs:208-215)."""
    return np.load(path)


def provider_from_model(model: Optional[FitnessModel], wildtype: np.ndarray, base_path: str = ".",
                        rng: Optional[np.random.Generator] = None) -> Optional[FitnessProvider]:
    """FitnessProvider::from_model (src/providers/fitness.rs:176-206) + FitnessTable::from_model
    (src/core/fitness/table.rs:25-47, including the L x 4 size check)."""
    if model is None:
        return None
    d, p = model.distribution, model.params
    if d == "Neutral":
        return FitnessProvider("Neutral", utility=model.utility)
    if d == "Exponential":
        w = p.get("weights", {})
        weights = (w.get("beneficial", 0.0), w.get("deleterious", 0.0), w.get("lethal", 0.0), w.get("neutral", 0.0))
        table = exponential_table(wildtype, weights, p["lambda_beneficial"], p["lambda_deleterious"], rng)
        return FitnessProvider("NonEpistatic", table=table, utility=model.utility)
    if d == "Lognormal":
        table = lognormal_table(wildtype, p.get("lethal", 0.0), p["mu"], p["sigma"], rng)
        return FitnessProvider("NonEpistatic", table=table, utility=model.utility)
    if d == "File":
        table = np.load(os.path.join(base_path, p["path"])).astype(np.float64)
        _check_table(table, wildtype)
        return FitnessProvider("NonEpistatic", table=table, utility=model.utility)
    if d == "Epistatic":
        table = np.load(os.path.join(base_path, p["path"])).astype(np.float64)
        _check_table(table, wildtype)
        epi = load_epistasis_npy(os.path.join(base_path, p["epi_path"]))
        return FitnessProvider("SimpleEpistatic", table=table, epistasis=epi, utility=model.utility)
    raise ValueError(f"unknown fitness distribution {d}")


def _check_table(table, wildtype):
    if table.size != len(wildtype) * 4:
        raise ValueError(f"Fitness table has wrong size: {table.size} instead of {len(wildtype) * 4}")


@dataclass
class HostFitness:
    """HostFitness{reproductive, infective} (src/config/parameters.rs:53-57)."""
    reproductive: Optional[FitnessModel] = None
    infective: Optional[FitnessModel] = None


@dataclass
class HostModel:
    """HostModel::{SingleHost(HostFitness), MultiHost(Vec<FitnessModelFraction>)} (parameters.rs:105-109)."""
    kind: str = "SingleHost"
    single: Optional[HostFitness] = None
    fractions: Sequence = ()  # [(fraction, HostFitness)]

    def make_definitions(self, parameters: Parameters, wildtype: np.ndarray, base_path: str = ".",
                         rng: Optional[np.random.Generator] = None) -> List[HostSpec]:
        """HostModel::make_definitions (src/config/parameters.rs:112-145), including the MultiHost range rule
        n = round(fraction*H); lower = i*n; upper = min(lower+n, H)."""
        H = int(parameters.host_population_size)
        if self.kind == "SingleHost":
            hf = self.single or HostFitness()
            return [HostSpec(0, H, provider_from_model(hf.reproductive, wildtype, base_path, rng),
                             provider_from_model(hf.infective, wildtype, base_path, rng))]
        specs = []
        for i, (fraction, hf) in enumerate(self.fractions):
            n_hosts = int(_rust_round(fraction * H))
            lower = i * n_hosts
            upper = min(lower + n_hosts, H)
            specs.append(HostSpec(lower, max(upper, lower), provider_from_model(hf.reproductive, wildtype, base_path, rng),
                                  provider_from_model(hf.infective, wildtype, base_path, rng)))
        return specs


def _rust_round(x: float) -> float:
    """f64::round — half away from zero."""
    return np.floor(x + 0.5) if x >= 0 else -np.floor(-x + 0.5)


# ------------------------------------------------------------------------- schedule (config/schedule.rs)

NAMED_TRANSFERS = {
    "default": [[1., 0., 0., 0.], [0., 1., 0., 0.], [0., 0., 1., 0.], [0., 0., 0., 1.]],
    "migration_fwd": [[1., 0., 0., 0.], [0.2, 0.8, 0., 0.], [0., 0.2, 0.8, 0.], [0., 0., 0., 0.]],
    "migration_rev": [[0.8, 0.2, 0., 0.], [0., 1., 0., 0.], [0., 0., 1., 0.], [0., 0., 0., 0.]],
    "root_ab": [[1., 0., 0., 0.], [1., 1., 0., 0.], [0., 0., 1., 0.], [0., 0., 0., 1.]],
    "root_bc": [[1., 0., 0., 0.], [0., 1., 0., 0.], [0., 1., 1., 0.], [0., 0., 0., 1.]],
    "root_cd": [[1., 0., 0., 0.], [0., 1., 0., 0.], [0., 0., 1., 0.], [0., 0., 1., 1.]],
}


@dataclass
class ScheduleRecord:
    generation: str
    event: str
    value: str


def _eval_generation_expr(text: str, generation: int) -> int:
    """`eval_int_with_context` of the schedule's generation column (src/config/schedule.rs:250-266): integer arithmetic over
    x / generation / {} with + - * / % ^ and parentheses.  `/` truncates toward zero and `%` takes the sign of the dividend
    (Rust i64), `^` is the power operator; anything else, a division by zero or a non-integer result is the reference's
    "Invalid generation" panic.  The expression is parsed, never executed."""
    bad = ValueError(f"Invalid generation `{text}`.")
    expr = text.replace("{}", "x").replace("^", "**")
    if len(expr) > 200:
        raise bad
    try:
        tree = ast.parse(expr.strip(), mode="eval")
    except SyntaxError:
        raise bad from None

    def ev(node) -> int:
        if isinstance(node, ast.Expression):
            return ev(node.body)
        if isinstance(node, ast.Constant) and type(node.value) is int:
            return node.value
        if isinstance(node, ast.Name) and node.id in ("x", "generation"):
            return generation
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
            v = ev(node.operand)
            return -v if isinstance(node.op, ast.USub) else v
        if isinstance(node, ast.BinOp):
            a, b = ev(node.left), ev(node.right)
            if isinstance(node.op, ast.Add):
                r = a + b
            elif isinstance(node.op, ast.Sub):
                r = a - b
            elif isinstance(node.op, ast.Mult):
                r = a * b
            elif isinstance(node.op, (ast.Div, ast.FloorDiv)):
                if b == 0:
                    raise bad
                r = abs(a) // abs(b) * (1 if (a >= 0) == (b >= 0) else -1)
            elif isinstance(node.op, ast.Mod):
                if b == 0:
                    raise bad
                r = abs(a) % abs(b) * (1 if a >= 0 else -1)
            elif isinstance(node.op, ast.Pow):
                if b < 0 or b > 64:
                    raise bad
                r = a ** b
            else:
                raise bad
            if not -(1 << 63) <= r < (1 << 63):  # i64 overflow is an evaluation error in the reference
                raise bad
            return r
        raise bad

    return ev(tree)


class Schedule:
    """Schedule (src/config/schedule.rs:110-266): first record whose generation matches (literal number, or an
    expression in x / generation / {} that evaluates to 0) and whose event name matches."""

    def __init__(self, records: Sequence[ScheduleRecord]):
        self.table = list(records)
        self.transfers = {k: np.asarray(v, dtype=np.float64) for k, v in NAMED_TRANSFERS.items()}
        for r in self.table:
            if r.event == "transmission" and r.value not in self.transfers:
                self.transfers[r.value] = np.asarray(ast.literal_eval(r.value), dtype=np.float64)

    @staticmethod
    def _match(record: ScheduleRecord, generation: int) -> bool:
        try:
            return int(record.generation) == generation
        except ValueError:
            return _eval_generation_expr(record.generation, generation) == 0

    def get_event_value(self, event: str, generation: int) -> Optional[str]:
        for r in self.table:
            if self._match(r, generation) and r.event == event:
                return r.value
        return None

    def get_transfer_matrix(self, generation: int) -> np.ndarray:
        name = self.get_event_value("transmission", generation)
        return self.transfers[name if name is not None else "default"]

    def get_sample_size(self, generation: int) -> int:
        v = self.get_event_value("sample", generation)
        return int(v) if v is not None else 0

    def check_transfer_table_sizes(self, minimum: int) -> bool:
        return all(t.shape[0] >= minimum for t in self.transfers.values())


# ------------------------------------------------------------------------- YAML settings (config/settings.rs)

def load_parameters(path: str) -> Parameters:
    """Parameters::read_from_file (src/config/parameters.rs:182-186): ONE parameter block as a YAML document — what a
    `settings` event of the schedule points to (src/config/schedule.rs:210-222).  Only the numeric parameters reach the
    simulation (set_parameters, src/simulation.rs:355-363): host specs and providers are built once, at start-up."""
    return _load(path, single=True)


def load_settings(path: str):
    """Settings{parameters: Vec<Parameters>, schedule} from the reference's YAML dialect (serde tags
    !SingleHost / !MultiHost / !Exponential / !Lognormal / !File / !Epistatic / !Algebraic / !Hill / !Neutral)."""
    return _load(path, single=False)


def _load(path: str, single: bool):
    import yaml

    class _Tagged(dict):
        tag = ""

    class _Loader(yaml.SafeLoader):
        pass

    def _multi(loader, suffix, node):
        if isinstance(node, yaml.MappingNode):
            v = _Tagged(loader.construct_mapping(node, deep=True))
        elif isinstance(node, yaml.SequenceNode):
            v = _Tagged(items=loader.construct_sequence(node, deep=True))
        else:
            v = _Tagged()
        v.tag = suffix
        return v

    _Loader.add_multi_constructor("!", _multi)
    with open(path) as fh:
        doc = yaml.load(fh, Loader=_Loader)

    def utility(u) -> UtilityFunction:
        if isinstance(u, str):
            return UtilityFunction(u)
        if u.tag == "Algebraic":
            return UtilityFunction("Algebraic", upper=float(u["upper"]))
        if u.tag == "Hill":
            if u["n"] < 0.0 or u["p"] > 1.0:
                raise ValueError("p must be between 0 and 1")
            return UtilityFunction("Hill", p=float(u["p"]), n=float(u["n"]))
        return UtilityFunction(u.tag or "Linear")

    def model(m) -> Optional[FitnessModel]:
        if m is None:
            return None
        d = m["distribution"]
        if isinstance(d, str):
            return FitnessModel(d, {}, utility(m.get("utility", "Linear")))
        return FitnessModel(d.tag, dict(d), utility(m.get("utility", "Linear")))

    def host_fitness(h) -> HostFitness:
        return HostFitness(model(h.get("reproductive")), model(h.get("infective")))

    params, host_models = [], []
    plist = [doc] if single else doc["parameters"]
    for p in plist:
        hm = p["host_model"]
        if hm.tag == "SingleHost":
            host_models.append(HostModel("SingleHost", single=host_fitness(hm)))
        else:
            host_models.append(HostModel("MultiHost", fractions=[(float(f["fraction"]), host_fitness(f["host_fitness"]))
                                                                 for f in hm["items"]]))
        params.append(Parameters(float(p["mutation_rate"]), float(p["recombination_rate"]),
                                 int(p["host_population_size"]), float(p["basic_reproductive_number"]),
                                 int(p["max_population"]), float(p["dilution"]), p["substitution_matrix"],
                                 int(p.get("n_hits", 1))))
    if single:
        return params[0]
    schedule = Schedule([ScheduleRecord(str(r["generation"]), str(r["event"]), str(r["value"]))
                         for r in doc.get("schedule", [])])
    return params, host_models, schedule
