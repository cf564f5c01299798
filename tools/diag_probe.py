#!/usr/bin/env python
"""Cost of diagonal ops by structure (developer tool): contiguous 2^12 tiles, n = 30."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantr_b200 as qb
from quantr_b200.circuit import encode_gates
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
G = qb.Gate
def t(label, build, **opts):
    c = qb.Circuit.new(n); build(c)
    enc = encode_gates(c.get_gates(), n)
    s = qb.DeviceState(n)
    for k, v in opts.items(): s.set_option(k, v)
    s.set_option("timing", 1)
    best = None
    for _ in range(4):
        st = s.apply(enc); ms = st["device_ms"]
        best = ms if best is None else min(best, ms)
    print(f"{label:64s} passes {st['n_passes']} rounds {st['n_rounds']}  {best:8.3f} ms", flush=True)
    s.close()
lastw = list(range(n - 12, n))
t("base: Z on last wire", lambda c: c.add_gate(G.Z, n - 1))
t("12 H on last 12 wires", lambda c: [c.add_gate(G.H, w) for w in lastw])
t("12 X on last 12 wires", lambda c: [c.add_gate(G.X, w) for w in lastw])
t("12 Rx on last 12 wires (general 2x2)", lambda c: [c.add_gate(G.Rx(0.3), w) for w in lastw])
t("12 Ry on last 12 wires (real 2x2)", lambda c: [c.add_gate(G.Ry(0.3), w) for w in lastw])
# diagonal ops: H between them prevents merging; controls on different sets
def hd(c, ctrl_fn):
    for w in lastw:
        c.add_gate(G.H, w)
        for cw in ctrl_fn(w):
            c.add_gate(G.CRk(3, cw), w)
t("12 x (H + CRk ladder on all later wires)   [QFT stage]", lambda c: hd(c, lambda w: range(w + 1, n)))
t("12 x (H + CRk from wire 0 only)            [ext-only phase]", lambda c: hd(c, lambda w: [0]))
t("12 x (H + CRk from wires 0..17)            [ext-only, 18 terms]", lambda c: hd(c, lambda w: range(0, 18)))
t("12 x (H + CRk from last wire)              [tile-bit control]", lambda c: hd(c, lambda w: [n - 1] if w != n - 1 else [n - 2]))
t("12 x (H + Rz on same wire)                 [cmask none, lin on reg bit]", lambda c: [(c.add_gate(G.H, w), c.add_gate(G.Rz(0.3), w)) for w in lastw])
t("12 x (H + T on same wire)", lambda c: [(c.add_gate(G.H, w), c.add_gate(G.T, w)) for w in lastw])
t("24 H (2 per wire, different)", lambda c: [(c.add_gate(G.H, w), c.add_gate(G.H, lastw[(i + 5) % 12])) for i, w in enumerate(lastw)])
