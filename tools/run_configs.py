#!/usr/bin/env python
"""Times BASELINE.json configs 1-3 through the public host API (developer tool; results go to profiles/).

  config 1: examples/grovers.rs 3-qubit Grover + measure_all(500)
  config 2: QFT-16 from |0xACE1> (+ the variant whose last three wires are one Gate::Custom QFT, tests/qft.rs)
  config 3: random layered circuit, 30 qubits, depth 100 (4,000 gates), seed 30
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import quantr_b200 as qb
from quantr_b200 import states as st
from quantr_b200.circuit import encode_gates
from golden import reference_vectors as rv
from helpers import qft_circuit, qft_expected, random_layered_circuit, OracleCircuit, orc

G = qb.Gate

def wall(fn, reps=5):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    return (time.perf_counter() - t0) / reps, out

# ---- config 1
def c1():
    sim = rv.build_example_grovers(qb.Circuit, G, st).simulate()
    bins = sim.measure_all(500).take()
    return sim, bins
t, (sim, bins) = wall(c1, 20)
p = np.abs(sim.get_state().take().get_amplitudes()) ** 2
print(f"config 1  grover-3 simulate + measure_all(500): {t*1e6:9.1f} us wall   passes {sim.stats['n_passes']}  p(110)={p[6]:.6f} p(111)={p[7]:.6f}  bins {{{', '.join(f'{k}:{v}' for k, v in sorted(bins.items(), key=lambda kv: kv[0].to_string()))}}}")

# ---- config 2
n, x = 16, 0xACE1
def c2():
    return qft_circuit(qb.Circuit, G, n, x).simulate()
t, sim = wall(c2, 10)
amps = sim.get_state().take().get_amplitudes()
print(f"config 2  QFT-16 (136 gates) simulate:          {t*1e6:9.1f} us wall   passes {sim.stats['n_passes']}  max-abs err vs closed form {np.max(np.abs(amps - qft_expected(n, x))):.2e}  gates/s {136/t:.3e}")
def c2b():
    c = qb.Circuit.new(n)
    for pos in range(n - 3):
        c.add_gate(G.H, pos)
        for k in range(2, n - pos + 1):
            c.add_gate(G.CRk(k, pos + k - 1), pos)
    c.add_gate(G.Custom(rv.make_qft_closure(qb.Circuit, G), [13, 14], "QFT"), 15)
    c.change_register(st.ProductState.binary_basis(x, n))
    return c.simulate()
t, sim = wall(c2b, 5)
amps2 = sim.get_state().take().get_amplitudes()
print(f"config 2b QFT-16 with 3-wire Gate::Custom QFT:   {t*1e6:9.1f} us wall   passes {sim.stats['n_passes']}  max-abs err vs closed form {np.max(np.abs(amps2 - qft_expected(n, x))):.2e}")

# ---- config 3
n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 30
c = random_layered_circuit(OracleCircuit, G, n3, 100, seed=30)
gates = list(c.circuit_gates)
n_gates = sum(1 for g in gates if g.kind != 0)
enc = encode_gates(gates, n3)
for label, opts in (("auto", {}), ("low_bits=3", {"low_bits": 3}), ("low_bits=4", {"low_bits": 4}), ("low_bits=5", {"low_bits": 5}), ("tile 12 low 4", {"tile_bits": 12, "low_bits": 4})):
    s = qb.DeviceState(n3)
    for k, v in opts.items():
        s.set_option(k, v)
    s.set_option("timing", 1)
    s.init_basis(0); s.apply(enc)  # warm-up
    t0 = time.perf_counter(); s.init_basis(0); stats = s.apply(enc); s.synchronize(); dt = time.perf_counter() - t0
    norm = s.norm_sqr()
    eff = stats["n_passes"] * stats["bytes_per_pass"] / (stats["device_ms"] * 1e-3) / 1e9
    print(f"config 3  random layered n={n3} depth 100 ({n_gates} gates) {label:14s}: {dt*1e3:9.1f} ms wall  device {stats['device_ms']:.1f} ms  passes {stats['n_passes']} "
          f"(passes/gate {stats['n_passes']/n_gates:.3f}, per layer {stats['n_passes']/100:.2f})  rounds {stats['n_rounds']}  {eff:.0f} GB/s per pass = {eff/6544.3:.3f} of HBM peak  norm {norm:.12f}", flush=True)
    s.close()
