#!/usr/bin/env python
"""Developer probe: BASELINE config 3's generator at n qubits / depth d through qsv_apply (for ncu captures)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import quantr_b200 as qb
from quantr_b200.circuit import encode_gates
from helpers import random_layered_circuit, OracleCircuit
n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
d = int(sys.argv[2]) if len(sys.argv) > 2 else 10
c = random_layered_circuit(OracleCircuit, qb.Gate, n, d, seed=30)
enc = encode_gates(list(c.circuit_gates), n)
s = qb.DeviceState(n)
s.set_option("timing", 1)
if os.environ.get("QSV_LOW_BITS"): s.set_option("low_bits", int(os.environ["QSV_LOW_BITS"]))
for rep in range(2):
    s.init_basis(0)
    st = s.apply(enc)
    s.synchronize()
    print(f"rep {rep}: passes {st['n_passes']} rounds {st['n_rounds']} device {st['device_ms']:.1f} ms", flush=True)
s.close()
