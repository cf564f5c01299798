"""`Gate`: mirror of quantr's gate enum (src/circuit/gate.rs:18-106).

`Gate.H`, `Gate.X`, ... are constants; parameterised variants are built by calling
`Gate.Rx(angle)`, `Gate.CNot(control)`, `Gate.Toffoli(c1, c2)`,
`Gate.Custom(func, controls, name)`.
"""
from __future__ import annotations

from . import _ffi as F


class Gate:
    __slots__ = ("variant", "kind", "param", "iparam", "controls", "func", "name")

    def __init__(self, variant, kind, param=0.0, iparam=0, controls=(), func=None, name=None):
        self.variant = variant
        self.kind = kind
        self.param = float(param)
        self.iparam = int(iparam)
        self.controls = tuple(int(c) for c in controls)
        self.func = func
        self.name = name

    # -- parameterised variants -----------------------------------------------------------
    @staticmethod
    def Rx(angle): return Gate("Rx", F.GATE_RX, param=angle)
    @staticmethod
    def Ry(angle): return Gate("Ry", F.GATE_RY, param=angle)
    @staticmethod
    def Rz(angle): return Gate("Rz", F.GATE_RZ, param=angle)
    @staticmethod
    def Phase(angle): return Gate("Phase", F.GATE_PHASE, param=angle)
    @staticmethod
    def CR(angle, control): return Gate("CR", F.GATE_CR, param=angle, controls=(control,))
    @staticmethod
    def CRk(k, control): return Gate("CRk", F.GATE_CRK, iparam=k, controls=(control,))
    @staticmethod
    def CZ(control): return Gate("CZ", F.GATE_CZ, controls=(control,))
    @staticmethod
    def CY(control): return Gate("CY", F.GATE_CY, controls=(control,))
    @staticmethod
    def CNot(control): return Gate("CNot", F.GATE_CNOT, controls=(control,))
    @staticmethod
    def Swap(control): return Gate("Swap", F.GATE_SWAP, controls=(control,))
    @staticmethod
    def Toffoli(control1, control2): return Gate("Toffoli", F.GATE_TOFFOLI, controls=(control1, control2))
    @staticmethod
    def Custom(func, controls, name): return Gate("Custom", F.GATE_CUSTOM, controls=tuple(controls), func=func, name=str(name))

    # -- gate.rs:110-138 -------------------------------------------------------------------
    def get_nodes(self):
        return None if self.is_single_gate() else list(self.controls)

    def is_single_gate(self) -> bool:  # gate.rs:173-201
        return self.kind <= F.GATE_PHASE

    def is_custom_gate(self) -> bool:  # gate.rs:203-208
        return self.kind == F.GATE_CUSTOM

    def get_name(self) -> str:  # gate.rs:210-238
        names = {"Id": "", "Sdag": "S*", "Tdag": "T*", "Phase": "P", "MX90": "X90*", "MY90": "Y90*", "Swap": "Sw",
                 "CZ": "Z", "CY": "Y", "CNot": "X", "Toffoli": "X"}
        if self.kind == F.GATE_CUSTOM:
            return self.name
        return names.get(self.variant, self.variant)

    def clone(self):
        return self  # immutable

    def __eq__(self, other):
        return (isinstance(other, Gate) and self.kind == other.kind and self.param == other.param and
                self.iparam == other.iparam and self.controls == other.controls and self.func is other.func and
                self.name == other.name)

    def __hash__(self):
        return hash((self.kind, self.param, self.iparam, self.controls, self.name))

    def __repr__(self):  # Debug formatting of the reference enum
        if self.kind == F.GATE_CUSTOM:
            return f"Custom({getattr(self.func, '__name__', self.func)}, {list(self.controls)}, {self.name!r})"
        if self.kind in (F.GATE_RX, F.GATE_RY, F.GATE_RZ, F.GATE_PHASE):
            return f"{self.variant}({self.param})"
        if self.kind == F.GATE_CR:
            return f"CR({self.param}, {self.controls[0]})"
        if self.kind == F.GATE_CRK:
            return f"CRk({self.iparam}, {self.controls[0]})"
        if self.controls:
            return f"{self.variant}({', '.join(str(c) for c in self.controls)})"
        return self.variant


for _name, _kind in [("Id", F.GATE_ID), ("H", F.GATE_H), ("X", F.GATE_X), ("Y", F.GATE_Y), ("Z", F.GATE_Z),
                     ("S", F.GATE_S), ("Sdag", F.GATE_SDAG), ("T", F.GATE_T), ("Tdag", F.GATE_TDAG),
                     ("X90", F.GATE_X90), ("Y90", F.GATE_Y90), ("MX90", F.GATE_MX90), ("MY90", F.GATE_MY90)]:
    setattr(Gate, _name, Gate(_name, _kind))
