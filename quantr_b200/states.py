"""Host-side state types: mirror of quantr's `states` module.

These are plain host IO types (what Custom closures receive and return, what
`change_register` / `get_state` exchange).  The amplitudes the simulation works on
live in HBM behind a `qsv_state` handle, not here.

Reference: src/circuit/states/{qubit,product_states,super_positions,super_positions_unchecked}.rs
"""
from __future__ import annotations

import enum

import numpy as np

from .error import QuantrError

ZERO_MARGIN = 1e-6  # super_positions.rs:18


class Qubit(enum.IntEnum):
    """qubit.rs:15"""
    Zero = 0
    One = 1

    def kronecker_prod(self, other: "Qubit") -> "ProductState":  # qubit.rs:35
        return ProductState.new_unchecked([self, other])


class ProductState:
    """A computational-basis label; wire 0 first (product_states.rs:18-23)."""

    __slots__ = ("qubits",)

    def __init__(self, qubits):
        self.qubits = [Qubit(q) for q in qubits]

    @staticmethod
    def new(product_state) -> "ProductState":  # product_states.rs:37
        if len(product_state) == 0:
            raise QuantrError("The slice of qubits is empty, it needs to at least have one element.")
        return ProductState(product_state)

    @staticmethod
    def new_unchecked(product_state) -> "ProductState":  # product_states.rs:104
        return ProductState(product_state)

    @staticmethod
    def from_qubit(q: Qubit) -> "ProductState":  # product_states.rs:242
        return ProductState([q])

    def get(self, i):  # product_states.rs:64
        return self.qubits[i] if 0 <= i < len(self.qubits) else None

    def get_qubits(self):  # product_states.rs:80
        return self.qubits

    def get_mut_qubits(self):  # product_states.rs:99
        return self.qubits

    def num_qubits(self) -> int:  # product_states.rs:135
        return len(self.qubits)

    def invert_digit(self, place_num: int) -> "ProductState":  # product_states.rs:154
        if place_num >= len(self.qubits):
            raise QuantrError(
                f"The position of the binary digit, {place_num}, is out of bounds. The product dimension is "
                f"{len(self.qubits)}, and so the position must be strictly less."
            )
        self.qubits[place_num] = Qubit.One if self.qubits[place_num] == Qubit.Zero else Qubit.Zero
        return self

    def kronecker_prod(self, other: Qubit) -> "ProductState":  # product_states.rs:180
        self.qubits.append(Qubit(other))
        return self

    def insert_qubits(self, qubits, pos):  # product_states.rs:112-123
        for e, i in enumerate(pos):
            self.qubits[i] = Qubit(qubits[e])

    def comp_basis(self) -> int:  # product_states.rs:191 (without the u32 overflow)
        v = 0
        for q in self.qubits:
            v = (v << 1) | int(q)
        return v

    @staticmethod
    def binary_basis(index: int, basis_size: int) -> "ProductState":  # product_states.rs:205-215
        return ProductState([(index >> n) & 1 for n in reversed(range(basis_size))])

    def to_string(self) -> str:  # product_states.rs:218-240
        return "".join("1" if q == Qubit.One else "0" for q in self.qubits)

    __str__ = to_string

    def __repr__(self):
        return f"ProductState({self.to_string()})"

    def __eq__(self, other):
        return isinstance(other, ProductState) and self.qubits == other.qubits

    def __hash__(self):
        return hash(tuple(int(q) for q in self.qubits))

    def __iter__(self):
        return iter(self.qubits)

    def clone(self):
        return ProductState(list(self.qubits))

    def into_super_position(self) -> "SuperPosition":  # super_positions.rs:360
        amps = np.zeros(1 << len(self.qubits), dtype=np.complex128)
        amps[self.comp_basis()] = 1.0
        return SuperPosition._raw(amps, len(self.qubits))


class SuperPosition:
    """Dense amplitude vector, canonical order (super_positions.rs:22-25)."""

    __slots__ = ("amplitudes", "product_dim")

    def __init__(self):
        raise TypeError("use SuperPosition.new / new_with_amplitudes / new_with_amplitudes_unchecked")

    @classmethod
    def _raw(cls, amplitudes: np.ndarray, product_dim: int) -> "SuperPosition":
        self = object.__new__(cls)
        self.amplitudes = amplitudes
        self.product_dim = product_dim
        return self

    @classmethod
    def new(cls, prod_dimension: int) -> "SuperPosition":  # super_positions.rs:41
        if prod_dimension == 0:
            raise QuantrError("The number of qubits must be non-zero.")
        amps = np.zeros(1 << prod_dimension, dtype=np.complex128)
        amps[0] = 1.0
        return cls._raw(amps, prod_dimension)

    @classmethod
    def new_unchecked(cls, num_qubits: int) -> "SuperPosition":  # super_positions_unchecked.rs:39-46
        amps = np.zeros(1 << num_qubits, dtype=np.complex128)
        amps[0] = 1.0
        return cls._raw(amps, num_qubits)

    @staticmethod
    def _equal_within_error(num: float, compare_num: float) -> bool:  # super_positions.rs:246-248
        return compare_num - ZERO_MARGIN < num < compare_num + ZERO_MARGIN

    @classmethod
    def _check_probability(cls, amps: np.ndarray):  # super_positions.rs:69-73, 236-240
        if not cls._equal_within_error(float(np.sum(amps.real ** 2 + amps.imag ** 2)), 1.0):
            raise QuantrError("Slice given to set amplitudes in super position does not conserve probability, "
                              "the absolute square sum of the coefficents must be one.")

    @classmethod
    def new_with_amplitudes(cls, amplitudes) -> "SuperPosition":  # super_positions.rs:68-88 (probability first, then length)
        amps = np.array(amplitudes, dtype=np.complex128).reshape(-1)
        cls._check_probability(amps)
        length = amps.shape[0]
        if (length & (length - 1)) != 0:
            raise QuantrError("The length of the array must be of the form 2**n where n is an integer.")
        return cls._raw(amps, length.bit_length() - 1)

    @classmethod
    def new_with_amplitudes_unchecked(cls, amplitudes) -> "SuperPosition":  # super_positions_unchecked.rs:64
        amps = np.array(amplitudes, dtype=np.complex128).reshape(-1)
        length = amps.shape[0]
        tz = (length & -length).bit_length() - 1 if length else 0
        return cls._raw(amps, tz)

    @classmethod
    def _check_hash_amplitudes(cls, hash_amplitudes: dict, product_dim: int, conserve_message: str):
        """super_positions.rs:113-123 / 279-290: every key has `product_dim` qubits, the squares sum to one."""
        total = 0.0
        for states, amplitude in hash_amplitudes.items():
            if states.num_qubits() != product_dim:
                raise QuantrError(f"The first state has product dimension of {product_dim}, whilst the state, |{states}>, "
                                  f"found as a key in the HashMap has dimension {states.num_qubits()}.")
            a = complex(amplitude)
            total += a.real * a.real + a.imag * a.imag
        if not cls._equal_within_error(total, 1.0):
            raise QuantrError(conserve_message.format(total))

    @staticmethod
    def _from_hash_to_array(hash_amplitudes: dict, n: int) -> np.ndarray:  # super_positions.rs:344-357
        amps = np.zeros(1 << n, dtype=np.complex128)
        for k, v in hash_amplitudes.items():
            amps[k.comp_basis()] = v
        return amps

    @classmethod
    def new_with_hash_amplitudes(cls, hash_amplitudes: dict) -> "SuperPosition":  # super_positions.rs:105-131
        if not hash_amplitudes:
            raise QuantrError("An empty HashMap was given. A superposition must have at least one non-zero state.")
        n = next(iter(hash_amplitudes)).num_qubits()
        cls._check_hash_amplitudes(hash_amplitudes, n, "The total sum of the absolute square of all amplitudes, {}, does not equal 1. "
                                                       "That is, the superpositon does not conserve probability.")
        return cls._raw(cls._from_hash_to_array(hash_amplitudes, n), n)

    def get_amplitude(self, pos: int):  # super_positions.rs:145
        return complex(self.amplitudes[pos]) if 0 <= pos < self.amplitudes.shape[0] else None

    def get_num_qubits(self) -> int:  # super_positions.rs:160
        return self.product_dim

    def get_dimension(self) -> int:  # super_positions.rs:175
        return int(self.amplitudes.shape[0])

    def get_amplitudes(self) -> np.ndarray:  # super_positions.rs:191
        return self.amplitudes

    def get_amplitude_from_state(self, prod_state: ProductState) -> complex:  # super_positions.rs:207
        if prod_state.num_qubits() != self.product_dim:
            raise QuantrError(
                f"Unable to retreive product state, |\"{prod_state.to_string()}\"> with dimension {prod_state.num_qubits()}. "  # ({:?} of a String: quoted)
                f"The superposition is a linear combination of states with different dimension. These dimensions should be equal."
            )
        return complex(self.amplitudes[prod_state.comp_basis()])

    def set_amplitudes(self, amplitudes) -> "SuperPosition":  # super_positions.rs:229
        amps = np.array(amplitudes, dtype=np.complex128).reshape(-1)
        if amps.shape[0] != self.amplitudes.shape[0]:
            raise QuantrError(
                f"The slice given to set the amplitudes in the computational basis has length {amps.shape[0]}, "
                f"when it should have length {self.amplitudes.shape[0]}."
            )
        self._check_probability(amps)
        self.amplitudes = amps
        return self

    def set_amplitudes_unchecked(self, amplitudes) -> "SuperPosition":  # super_positions_unchecked.rs:74
        self.amplitudes = np.array(amplitudes, dtype=np.complex128).reshape(-1)
        return self

    def set_amplitudes_from_states(self, amplitudes: dict) -> "SuperPosition":  # super_positions.rs:270-295
        if not amplitudes:
            raise QuantrError("An empty HashMap was given. A superposition must have at least one non-zero state.")
        self._check_hash_amplitudes(amplitudes, self.product_dim, "The total sum of the absolute square of all amplitudes does not equal 1. "
                                                                  "That is, the superpositon does not conserve probability.")
        self.amplitudes = self._from_hash_to_array(amplitudes, self.product_dim)
        return self

    def to_hash_map(self) -> dict:  # super_positions.rs:315-323 (amplitudes with |a|^2 within 1e-6 of zero are left out)
        out = {}
        for i, a in enumerate(self.amplitudes):
            if not self._equal_within_error(a.real * a.real + a.imag * a.imag, 0.0):
                out[ProductState.binary_basis(i, self.product_dim)] = complex(a)
        return out

    def measure(self, dice_roll: float | None = None):  # super_positions.rs:332-342
        """Observes the superposition: the first state whose cumulative probability exceeds the dice roll (strict `<`), or
        None when the squares sum to less than the roll (non-unitary Custom gates).  The reference draws the roll from
        `fastrand::f64()`; here it comes from the package generator (`quantr_b200.seed`) unless given."""
        if dice_roll is None:
            from . import circuit
            dice_roll = float(circuit._rng.random())
        cumulative = np.cumsum(self.amplitudes.real ** 2 + self.amplitudes.imag ** 2)
        hit = np.nonzero(dice_roll < cumulative)[0]
        return ProductState.binary_basis(int(hit[0]), self.product_dim) if hit.size else None

    def __iter__(self):  # super_position_iter.rs:56-72 (zeros included)
        for i, a in enumerate(self.amplitudes):
            yield ProductState.binary_basis(i, self.product_dim), complex(a)

    def clone(self) -> "SuperPosition":
        return SuperPosition._raw(self.amplitudes.copy(), self.product_dim)


def into_super_position(x) -> SuperPosition:
    """`.into()` of the reference: Qubit / ProductState / SuperPosition -> SuperPosition."""
    if isinstance(x, SuperPosition):
        return x
    if isinstance(x, ProductState):
        return x.into_super_position()
    if isinstance(x, Qubit):
        return ProductState([x]).into_super_position()
    raise TypeError(f"cannot convert {type(x).__name__} into a SuperPosition")
