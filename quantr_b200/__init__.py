"""quantr_b200 — B200-native state-vector engine behind quantr's `Circuit::simulate`.

Layout:
  csrc/            hand-written sm_100a CUDA kernels, the fusion scheduler and the C ABI (include/qsv.h)
  libqsv.so        built in-tree by `make lib`; there is no CPU fallback
  circuit.py ...   host-side mirror of quantr's public API (Circuit, Gate, SimulatedCircuit, states)
"""
from . import states
from .circuit import Circuit, DeviceState, EncodedOps, Measurement, Plan, SimulatedCircuit, encode_gates, seed
from .error import QuantrError
from .gate import Gate
from .states import ProductState, Qubit, SuperPosition

__all__ = [
    "Circuit", "Gate", "SimulatedCircuit", "Measurement", "QuantrError", "states", "ProductState", "Qubit",
    "SuperPosition", "DeviceState", "Plan", "EncodedOps", "encode_gates", "seed",
]
