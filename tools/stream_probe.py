#!/usr/bin/env python
"""Times single fused passes with (almost) no arithmetic to measure the streaming ceiling of the tile access pattern.

  python tools/stream_probe.py [n_qubits]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import quantr_b200 as qb
from quantr_b200.circuit import encode_gates

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
G = qb.Gate

def time_circuit(label, build, **opts):
    c = qb.Circuit.new(n)
    build(c)
    enc = encode_gates(c.get_gates(), n)
    s = qb.DeviceState(n)
    for k, v in opts.items():
        s.set_option(k, v)
    s.set_option("timing", 1)
    best = None
    for _ in range(4):
        st = s.apply(enc)
        ms = st["device_ms"] / max(1, st["n_passes"])
        best = ms if best is None else min(best, ms)
    gbs = 32.0 * (1 << n) / (best * 1e-3) / 1e9
    print(f"{label:52s} passes {st['n_passes']}  ms/pass {best:8.3f}  {gbs:8.1f} GB/s", flush=True)
    s.close()

time_circuit("Z on last wire (contiguous tile, 1 diag)", lambda c: c.add_gate(G.Z, n - 1))
for tb in (11, 12, 13):
    for lb in (3, 4, 5, 6, 7, 8):
        k = tb - lb
        time_circuit(f"X on top {k} wires, tile_bits={tb} low_bits={lb} ({16 << lb} B runs)", lambda c: [c.add_gate(G.X, w) for w in range(k)], tile_bits=tb, low_bits=lb)
time_circuit("H+CRk ladder on wires 0..6 (7 stages), T=12 L=5", lambda c: [(c.add_gate(G.H, w), [c.add_gate(G.CRk(k, w + k - 1), w) for k in range(2, n - w + 1)]) for w in range(7)], tile_bits=12, low_bits=5)
time_circuit("H+CRk ladder on wires 0..8 (9 stages), T=12 L=3", lambda c: [(c.add_gate(G.H, w), [c.add_gate(G.CRk(k, w + k - 1), w) for k in range(2, n - w + 1)]) for w in range(9)], tile_bits=12, low_bits=3)
time_circuit("H+CRk ladder on last 12 wires (12 stages), T=12", lambda c: [(c.add_gate(G.H, w), [c.add_gate(G.CRk(k, w + k - 1), w) for k in range(2, n - w + 1)]) for w in range(n - 12, n)], tile_bits=12, low_bits=3)
