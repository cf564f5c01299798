"""Generates tests/golden/config3_n<N>_d100.npz: BASELINE config 3 (random layered circuit, depth 100, seed 30;
SURVEY.md 8d) run through the CPU oracle's dense mode (oracle/quantr_oracle.cpp, which restates
/root/reference/src/circuit/simulation.rs:64-135) at sizes the oracle needs minutes to tens of minutes for, sampled at
4,096 seeded indices plus the norm.

    python tests/golden/make_config3_fixtures.py 26 28

The GPU parity test (tests/test_gpu_parity.py::test_config3_depth100_against_oracle_fixture) compares the device
result with these samples; a smaller size runs against the live oracle in the same test file.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import OracleCircuit, encode_gates, orc, qb, random_layered_circuit  # noqa: E402


def sample_indices(n, count=4096):
    return np.random.default_rng(1000 + n).integers(0, 1 << n, size=count, dtype=np.uint64)


def main():
    for n in [int(a) for a in sys.argv[1:]] or [26]:
        c = random_layered_circuit(OracleCircuit, qb.Gate, n, 100, seed=30)
        enc = encode_gates(c.circuit_gates, n)
        t0 = time.time()
        amps = orc.simulate(n, enc.ops, enc.n_ops, None, mode="dense", threads=int(os.environ.get("ORACLE_THREADS", os.cpu_count() or 1)))
        idx = sample_indices(n)
        out = os.path.join(HERE, f"config3_n{n}_d100.npz")
        np.savez_compressed(out, n=n, depth=100, seed=30, n_gates=enc.n_ops, indices=idx, amps=amps[idx.astype(np.int64)],
                            norm_sqr=float(np.sum(amps.real ** 2 + amps.imag ** 2)))
        print(f"n={n}: {enc.n_ops} gates, oracle {time.time() - t0:.0f} s -> {out}", flush=True)


if __name__ == "__main__":
    main()
