"""Regenerates the committed fixtures under tests/golden/ (run in the build container, where /root/reference exists).

  reference_fixtures.npz   inputs the reference's own integration tests use for this path, re-encoded:
      phix174_wildtype      tests/{singlehost,multihost,epistasis}/ref.fasta (phiX174, 5386 nt), as base indices
                            A,T,C,G = 0..3 (src/encoding.rs:65-72)
      hiv_wildtype          data/hiv_generated.fasta (2600 nt), as base indices
      epistasis_fitness     tests/epistasis/fitness_table.npy   (5386 x 4 f64)
      epistasis_pairs       tests/epistasis/epistasis_table.npy (pos1, base1, pos2, base2, value) as an (n,5) f64 array
  oracle_replay_*.npz      draw logs + outputs of the CPU oracle for small seeded runs (regression pin of the oracle,
                            and committed replay vectors for the GPU parity tests)

The reference itself cannot be executed here (no Rust toolchain), so there are no outputs *of the
reference* to record; its known-answer tests are restated literally in tests/test_oracle_kat.py instead.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
REF = "/root/reference"


def reference_fixtures():
    from virolution_b200.config import read_fasta_wildtype
    phix = read_fasta_wildtype(os.path.join(REF, "tests/epistasis/ref.fasta"))
    for other in ("tests/singlehost/ref.fasta", "tests/multihost/ref.fasta"):
        assert np.array_equal(phix, read_fasta_wildtype(os.path.join(REF, other)))
    hiv = read_fasta_wildtype(os.path.join(REF, "data/hiv_generated.fasta"))
    fit = np.load(os.path.join(REF, "tests/epistasis/fitness_table.npy"))
    epi = np.load(os.path.join(REF, "tests/epistasis/epistasis_table.npy"))
    pairs = np.array([[e["pos1"], e["base1"], e["pos2"], e["base2"], e["value"]] for e in epi], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "reference_fixtures.npz"), phix174_wildtype=phix, hiv_wildtype=hiv,
                        epistasis_fitness=fit, epistasis_pairs=pairs)
    print("phix174", phix.size, "hiv", hiv.size, "fitness", fit.shape, "pairs", pairs.shape)


def oracle_replays():
    import helpers
    for name in helpers.GOLDEN_CASES:
        log = helpers.record_oracle_run(name)
        np.savez_compressed(os.path.join(HERE, f"oracle_replay_{name}.npz"), **log)
        print(name, {k: v.shape for k, v in log.items() if k.startswith("g0_")})


if __name__ == "__main__":
    if os.path.isdir(REF):
        reference_fixtures()
    oracle_replays()
