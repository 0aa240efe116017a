"""Host-side circuit compiler: ansatz name -> flat gate program for the CUDA library.

The reference builds its circuit by calling PennyLane gate functions inside a QNode
(``nn/DVQuantumLayer.py:176-214`` and the six builders at ``:246-371``).  Here the same gate
sequence is emitted once as an ``int32[n_ops, 4]`` table ``(kind, a, b, p)`` (see
``include/qcpinn_b200.h``) that the device interprets; ``p`` is a flat index into the row-major
``(L, P)`` angle matrix.  Only the batch-shared part is compiled (L ansatz layers, the optional
Haar blocks, the final Hadamard); the per-sample encoding is a kernel template.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

RX, RY, RZ, CRX, CRZ, CNOT, HAD, U4, RY_IN, RZ_IN, CZ = range(11)
ENC_ANGLE, ENC_AMPLITUDE, ENC_NONE = 0, 1, 2

# ansaetze whose per-sample gates live INSIDE the program (data re-uploading): no separate encoding
# stage, no Haar blocks, no final Hadamard
REUPLOAD_ANSATZ = ("cz_melt",)

ANSATZ_PARAM_COUNT = {
    # reference nn/DVQuantumLayer.py:25-78
    "layered": lambda n: 4 * n,
    "alternate": lambda n: 4 * n - 4,
    "cascade": lambda n: 3 * n,
    "farhi": lambda n: 2 * n - 2,
    "sim_circ_15": lambda n: 2 * n,
    "cross_mesh": lambda n: 4 * n + n * (n - 1),
    # reference hybrid_testing/CG_HQPINN_IBMtest_16qubits.py:236 weight_shapes (n_layers, n_qubits, 3)
    "cz_melt": lambda n: 3 * n,
}


def params_per_layer(ansatz: str, n: int) -> int:
    try:
        return ANSATZ_PARAM_COUNT[ansatz](n)
    except KeyError:
        raise ValueError("Parameters are not initialized. Check the q_ansatz value.") from None


class _Emitter:
    """Collects ops of one layer; angle slots are handed out in emission order."""

    def __init__(self, n: int, offset: int, limit: int):
        self.n, self.offset, self.limit = n, offset, limit
        self.used = 0
        self.ops: list[tuple[int, int, int, int]] = []

    def _slot(self) -> int:
        if self.used >= self.limit:
            # same failure the reference hits when a builder runs past its angle row
            raise IndexError(
                f"index {self.used} is out of bounds for dimension 0 with size {self.limit}")
        self.used += 1
        return self.offset + self.used - 1

    def rot(self, kind: int, wire: int):
        self.ops.append((kind, wire, -1, self._slot()))

    def crot(self, kind: int, control: int, target: int):
        self.ops.append((kind, control, target, self._slot()))

    def cnot(self, control: int, target: int):
        self.ops.append((CNOT, control, target, -1))

    def cz(self, a: int, b: int):
        self.ops.append((CZ, a, b, -1))

    def sample_rot(self, kind: int, wire: int, input_index: int, quarters: int):
        """Per-sample rotation by (quarters / 4) * z[input_index] (kind = RY_IN / RZ_IN)."""
        self.ops.append((kind, wire, input_index, quarters))


def _emit_layer(ansatz: str, e: _Emitter, variant: str = "dv", layer: int = 0) -> None:
    n = e.n
    if ansatz == "cz_melt":
        # reference hybrid_testing/CG_HQPINN_IBMtest_16qubits.py:224-234: per wire the re-upload
        # RZ(0.5 * inputs[(i + layer) % n]) then Rot(phi, theta, omega) = RZ(omega) RY(theta) RZ(phi);
        # CZ on even pairs, odd pairs, and the ring closure (n-1, 0)
        for w in range(n):
            e.sample_rot(RZ_IN, w, (w + layer) % n, 2)
            e.rot(RZ, w)
            e.rot(RY, w)
            e.rot(RZ, w)
        for w in range(0, n - 1, 2):
            e.cz(w, w + 1)
        for w in range(1, n - 1, 2):
            e.cz(w, w + 1)
        e.cz(n - 1, 0)
        return
    if ansatz == "alternate" and variant == "single_file":
        # reference train_hybrid_qpinn.py:273-295: neighbour pairs WITHOUT the wrap-around pair, so
        # even qubit counts work (4n - 4 angles are exactly enough)
        for c in list(range(0, n - 1, 2)) + list(range(1, n - 1, 2)):
            e.rot(RY, c)
            e.rot(RY, c + 1)
            e.cnot(c, c + 1)
            e.rot(RZ, c)
            e.rot(RZ, c + 1)
        return
    if ansatz == "layered":          # RZ,RX per wire | CNOT ring | RX,RZ per wire
        for w in range(n):
            e.rot(RZ, w)
            e.rot(RX, w)
        for w in range(n):
            e.cnot(w, (w + 1) % n)
        for w in range(n):
            e.rot(RX, w)
            e.rot(RZ, w)
    elif ansatz == "alternate":      # RY,RY,CNOT,RZ,RZ blocks on even then odd neighbour pairs
        starts = list(range(0, n - 1, 2)) + list(range(1, n, 2))
        for c in starts:
            t = (c + 1) % n
            e.rot(RY, c)
            e.rot(RY, t)
            e.cnot(c, t)
            e.rot(RZ, c)
            e.rot(RZ, t)
    elif ansatz == "cascade":        # RX^n, RZ^n, CRX(n-1 -> 0), CRX(i-1 -> i) for i = n-1..1
        for w in range(n):
            e.rot(RX, w)
        for w in range(n):
            e.rot(RZ, w)
        e.crot(CRX, n - 1, 0)
        for w in range(n - 1, 0, -1):
            e.crot(CRX, w - 1, w)
    elif ansatz == "farhi":          # CNOT.R(last).CNOT sandwiches: RXX then RZX family
        last = n - 1
        for kind in (RX, RZ):
            for w in range(n - 1):
                e.cnot(last, w)
                e.rot(kind, last)
                e.cnot(last, w)
    elif ansatz == "sim_circ_15":    # RY^n, descending CNOT ring, RY^n, skip-3 CNOT ring
        for w in range(n):
            e.rot(RY, w)
        for w in range(n - 1, -1, -1):
            e.cnot(w, (w + 1) % n)
        for w in range(n):
            e.rot(RY, w)
        for w in range(n):
            c = (w + n - 1) % n
            e.cnot(c, (c + 3) % n)
    elif ansatz == "cross_mesh":     # RX^n, RZ^n, all-to-all CRZ (descending), RX^n, RZ^n
        for kind in (RX, RZ):
            for w in range(n):
                e.rot(kind, w)
        for c in range(n - 1, -1, -1):
            for t in range(n - 1, -1, -1):
                if t != c:
                    e.crot(CRZ, c, t)
        for kind in (RX, RZ):
            for w in range(n):
                e.rot(kind, w)
    else:
        raise ValueError("Parameters are not initialized. Check the q_ansatz value.")


def haar_pair(seed: int):
    """The two fixed Haar 4x4 unitaries of reference ``nn/DVQuantumLayer.py:203-207``."""
    from scipy.stats import unitary_group

    return tuple(
        np.asarray(unitary_group.rvs(4, random_state=np.random.RandomState(s)), dtype=np.complex128)
        for s in (seed, seed + 1)
    )


@dataclass(frozen=True)
class CircuitProgram:
    n_qubits: int
    n_layers: int
    ansatz: str
    ops: np.ndarray       # int32 [n_ops, 4]
    consts: np.ndarray    # complex128 [n_consts, 4, 4]
    n_theta: int

    @property
    def params_per_layer(self) -> int:
        return self.n_theta // max(self.n_layers, 1)

    @property
    def reupload(self) -> bool:
        """True when the per-sample gates are part of the program (no separate encoding stage)."""
        return self.ansatz in REUPLOAD_ANSATZ


def compile_program(ansatz: str, n_qubits: int, n_layers: int, haar_seed=None,
                    variant: str = "dv") -> CircuitProgram:
    """Gate table for: L x ansatz -> [Haar on wires (0,1),(2,3) iff seed given and n >= 4] -> H(n-1).

    Raises ``IndexError`` exactly where the reference does when a builder over-indexes its angle
    row (``alternate`` with an even qubit count).
    """
    per_layer = params_per_layer(ansatz, n_qubits)
    ops: list[tuple[int, int, int, int]] = []
    if ansatz in REUPLOAD_ANSATZ:
        # RY(inputs[i]) on wire i (reference :220-221), the layers, measurement; nothing else
        for w in range(n_qubits):
            ops.append((RY_IN, w, w, 4))
        for layer in range(n_layers):
            em = _Emitter(n_qubits, layer * per_layer, per_layer)
            _emit_layer(ansatz, em, variant, layer)
            ops.extend(em.ops)
        table = np.asarray(ops, dtype=np.int32).reshape(-1, 4)
        return CircuitProgram(n_qubits, n_layers, ansatz, table,
                              np.zeros((0, 4, 4), dtype=np.complex128), n_layers * per_layer)
    for layer in range(n_layers):
        em = _Emitter(n_qubits, layer * per_layer, per_layer)
        _emit_layer(ansatz, em, variant)
        ops.extend(em.ops)
    consts = np.zeros((0, 4, 4), dtype=np.complex128)
    if haar_seed is not None and n_qubits >= 4:
        consts = np.stack(haar_pair(int(haar_seed)))
        ops.append((U4, 0, 1, 0))
        ops.append((U4, 2, 3, 1))
    if n_qubits > 0:
        ops.append((HAD, n_qubits - 1, -1, -1))
    table = np.asarray(ops, dtype=np.int32).reshape(-1, 4)
    return CircuitProgram(n_qubits, n_layers, ansatz, table, consts, n_layers * per_layer)
