"""Synthetic benchmark configurations K1..K5: seeded scale-ups of the reference's
shipped settings (data/singlehost.yaml, data/hiv.yaml, data/multihost.yaml, tests/epistasis).  The reference's own
table generator is unseeded; here every table comes from a fixed numpy seed so both arms see identical inputs."""
from __future__ import annotations

import numpy as np

from .abi import FitnessProvider, HostSpec, Parameters, UtilityFunction
from .config import exponential_table

UNIFORM_SUBST = ((0., 1., 1., 1.), (1., 0., 1., 1.), (1., 1., 0., 1.), (1., 1., 1., 0.))
# data/hiv.yaml:9-25 substitution matrix (rows = from, A,T,C,G order)
HIV_SUBST = ((0.0, 0.24752475, 0.11371137, 0.09370937), (0.38363836, 0.0, 0.08910891, 0.09150915),
             (0.21922192, 0.11081108, 0.0, 0.24252425), (0.18228177, 0.11488851, 0.24477553, 0.0))
ALGEBRAIC = UtilityFunction("Algebraic", upper=1.5)


def hiv_like_sequence(L: int, seed: int = 1234) -> np.ndarray:
    """iid A .37 / G .24 / T .19 / C .20 (composition of data/hiv_generated.fasta), as indices A,T,C,G = 0..3."""
    rng = np.random.default_rng(seed)
    return rng.choice(4, size=L, p=[0.37, 0.19, 0.20, 0.24]).astype(np.uint8)


def uniform_sequence(L: int, seed: int = 4321) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 4, size=L).astype(np.uint8)


def synthetic_pairs(wt: np.ndarray, n_pairs: int, seed: int = 99) -> np.ndarray:
    rng = np.random.default_rng(seed)
    L = len(wt)
    p = np.sort(np.stack([rng.integers(0, L, n_pairs), rng.integers(0, L, n_pairs)], axis=1), axis=1)
    p[p[:, 0] == p[:, 1], 1] = (p[p[:, 0] == p[:, 1], 1] + 1) % L
    b1 = (wt[p[:, 0]] + rng.integers(1, 4, n_pairs)) % 4
    b2 = (wt[p[:, 1]] + rng.integers(1, 4, n_pairs)) % 4
    v = rng.lognormal(0.0, 0.1, n_pairs)
    return np.stack([p[:, 0], b1, p[:, 1], b2, v], axis=1).astype(np.float64)


def workload(name: str, n: int | None = None):
    """-> dict(name, description, wildtype, parameters, specs, n0)."""
    if name == "k1_singlehost":
        N = n or 10_000
        wt = uniform_sequence(5386, 174)  # phiX174 length; the real sequence is only in the test fixtures
        par = Parameters(1e-6, 0.0, N, 100.0, N, 0.02, UNIFORM_SUBST)
        tab = exponential_table(wt, rng=np.random.default_rng(1))
        specs = [HostSpec(0, N, FitnessProvider("NonEpistatic", table=tab, utility=ALGEBRAIC))]
        desc = f"K1 singlehost phiX174-length (5386 nt), N=H=max_population={N}, mu=1e-6, R0=100, dilution=0.02"
    elif name == "k2_hiv":
        N = n or 10_000_000
        wt = hiv_like_sequence(9700)
        par = Parameters(2.16e-5, 0.0, N, 6.0, N, 1.0, HIV_SUBST)
        tab = exponential_table(wt, rng=np.random.default_rng(2))
        specs = [HostSpec(0, N, FitnessProvider("NonEpistatic", table=tab, utility=ALGEBRAIC))]
        desc = (f"K2 HIV-length (9700 nt) single host, Exponential DFE + Algebraic(1.5), mu=2.16e-5, R0=6, dilution=1, "
                f"N=H=max_population={N}")
    elif name == "k4_epistasis":
        N = n or 1_000_000
        wt = hiv_like_sequence(9700)
        par = Parameters(2.16e-5, 0.0, N, 100.0, N, 0.02, UNIFORM_SUBST)
        tab = exponential_table(wt, rng=np.random.default_rng(2))
        pairs = synthetic_pairs(wt, 100_000)
        specs = [HostSpec(0, N, FitnessProvider("SimpleEpistatic", table=tab, epistasis=pairs, utility=ALGEBRAIC))]
        desc = f"K4 epistasis HIV-length, 1e5 sparse pairs, N=H=max_population={N}"
    elif name == "k3_multihost":
        N = n or 100_000
        wt = uniform_sequence(5386, 174)
        par = Parameters(1e-6, 0.0, N, 100.0, N, 0.02, UNIFORM_SUBST)
        t0 = exponential_table(wt, rng=np.random.default_rng(3))
        t1 = exponential_table(wt, rng=np.random.default_rng(4))
        half = int(np.floor(0.5 * N + 0.5))
        specs = [HostSpec(0, half, FitnessProvider("NonEpistatic", table=t0, utility=ALGEBRAIC)),
                 HostSpec(half, min(2 * half, N), FitnessProvider("NonEpistatic", table=t1, utility=ALGEBRAIC))]
        desc = f"K3 multihost (2 host specs x 0.5) phiX174-length, N=H=max_population={N} per compartment"
    elif name == "k5_sars":
        N = n or 10_000_000
        wt = uniform_sequence(30_000, 2019)
        par = Parameters(1e-5, 1e-4, N, 100.0, N, 0.02, UNIFORM_SUBST)
        tab = exponential_table(wt, rng=np.random.default_rng(5))
        specs = [HostSpec(0, N, FitnessProvider("NonEpistatic", table=tab, utility=ALGEBRAIC))]
        desc = f"K5 30 kb genome, mu=1e-5, rho=1e-4, R0=100, dilution=0.02, N=H=max_population={N}"
    else:
        raise KeyError(name)
    return dict(name=name, description=desc, wildtype=wt, parameters=par, specs=specs, n0=N)
