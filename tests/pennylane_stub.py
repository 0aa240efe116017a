"""A minimal stand-in for the part of the PennyLane API the reference calls (test infrastructure).

PennyLane is not installable in this container, so the reference's own ``nn/DVQuantumLayer.py``
cannot be imported as is.  This module provides just enough of ``pennylane`` -- ``device``,
``QNode``, the gate constructors the reference uses, ``templates.AngleEmbedding`` /
``AmplitudeEmbedding``, ``expval(PauliZ(i))`` -- for the UNMODIFIED reference source to execute:

* every gate call is recorded on a tape (name, wires, parameter object), so tests can compare the
  reference's real gate order / parameter indexing / control-target order with this repository's
  gate compiler (``program.compile_program``) and oracle;
* the tape is executed by a small broadcasting statevector simulator (torch, complex128,
  differentiable to any order), so the reference's ``DVPDESolver.forward`` and ``nn/pde.py`` run
  end to end.

It is NOT PennyLane: the gate matrices, the wire order (wire 0 = most significant bit) and the
embedding rules below restate PennyLane's published definitions, so fixtures produced through it
pin the reference's own circuit-construction code but not ``default.qubit`` itself
(``meta["simulator"] == "pennylane_stub"`` marks them; real-PennyLane fixtures come from
``tests/golden/make_pennylane_golden.py`` run where PennyLane exists).  Written independently of
``oracle/`` (different state layout: the state is kept as a (B, 2, ..., 2) tensor and gates are
contracted with ``torch.tensordot``), so agreement between the two is a genuine cross-check.
"""

import math
import sys
import types

import numpy as np
import torch

__version__ = "0.0-stub"
_TAPE = None


class Operation:
    def __init__(self, name, wires, param=None, matrix=None):
        self.name = name
        self.wires = [int(w) for w in (wires if isinstance(wires, (list, tuple, range)) else [wires])]
        self.param = param          # the object the reference passed (e.g. a view of self.params)
        self.matrix = matrix        # fixed matrix for QubitUnitary
        if _TAPE is not None:
            _TAPE.append(self)

    def __repr__(self):
        return f"{self.name}{self.wires}"


def _make(name):
    def ctor(*args, wires=None, **kw):
        if name in ("CNOT", "Hadamard", "PauliZ", "CZ"):
            if wires is None and args:
                wires = args[0]
            return Operation(name, wires)
        if name == "QubitUnitary":
            return Operation(name, wires, matrix=np.asarray(args[0], dtype=np.complex128))
        return Operation(name, wires, param=args[0])
    ctor.__name__ = name
    return ctor


RX, RY, RZ, PhaseShift, CRX, CRZ = (_make(k) for k in ("RX", "RY", "RZ", "PhaseShift", "CRX", "CRZ"))
CNOT, Hadamard, QubitUnitary = _make("CNOT"), _make("Hadamard"), _make("QubitUnitary")
CZ = _make("CZ")


def Rot(phi, theta, omega, wires=None):
    """PennyLane: Rot(phi, theta, omega) = RZ(omega) RY(theta) RZ(phi) (RZ(phi) acts first)."""
    Operation("RZ", wires, param=phi)
    Operation("RY", wires, param=theta)
    Operation("RZ", wires, param=omega)


class PauliZ:
    def __init__(self, wires):
        self.wire = int(wires)


class _Expval:
    def __init__(self, obs):
        self.obs = obs


def expval(obs):
    return _Expval(obs)


def _angle_embedding(features, wires, rotation="X"):
    """PennyLane AngleEmbedding: rotation(features[..., i]) on wires[i] (trailing axis = features)."""
    wires = list(wires)
    features = torch.as_tensor(features)
    assert features.shape[-1] <= len(wires)
    for i in range(features.shape[-1]):
        Operation("R" + rotation, wires[i], param=features[..., i])


def _amplitude_embedding(features, wires, normalize=False, pad_with=None):
    """PennyLane AmplitudeEmbedding: pad the trailing axis to 2^n with ``pad_with``, L2-normalise,
    load as the state (basis index = wires in order, wire 0 most significant)."""
    Operation("AmplitudeEmbedding", list(wires), param=(torch.as_tensor(features), normalize, pad_with))


templates = types.SimpleNamespace(AngleEmbedding=_angle_embedding, AmplitudeEmbedding=_amplitude_embedding)
AngleEmbedding, AmplitudeEmbedding = _angle_embedding, _amplitude_embedding


class _Device:
    def __init__(self, name, wires, shots=None, **kw):
        self.name, self.num_wires, self.shots = name, int(wires), shots


def device(name, wires=None, shots=None, **kw):
    if name != "default.qubit":
        raise RuntimeError(f"pennylane_stub only provides default.qubit, not {name}")
    return _Device(name, wires, shots)


def draw_mpl(qnode):
    raise RuntimeError("pennylane_stub has no circuit drawer")


# ---------------------------------------------------------------------------------------------
# tape execution
# ---------------------------------------------------------------------------------------------

def _rot_matrix(name, theta):
    """2x2 matrix of a rotation; theta is a 0-d tensor (shared) or (B,) (per sample).  Computed in the
    parameter's own real dtype, then promoted to complex128 (what qml.math does with a c128 state)."""
    theta = torch.as_tensor(theta)
    c, s = torch.cos(theta / 2), torch.sin(theta / 2)
    zero = torch.zeros_like(c)
    if name == "RX":
        m = [[torch.complex(c, zero), torch.complex(zero, -s)], [torch.complex(zero, -s), torch.complex(c, zero)]]
    elif name == "RY":
        m = [[torch.complex(c, zero), torch.complex(-s, zero)], [torch.complex(s, zero), torch.complex(c, zero)]]
    elif name == "RZ":
        cz = torch.complex(zero, zero)
        m = [[torch.complex(c, -s), cz], [cz, torch.complex(c, s)]]
    elif name == "PhaseShift":
        cz = torch.complex(zero, zero)
        one = torch.complex(torch.ones_like(c), zero)
        m = [[one, cz], [cz, torch.complex(torch.cos(theta), torch.sin(theta))]]
    else:
        raise KeyError(name)
    rows = [torch.stack(r, dim=-1) for r in m]
    return torch.stack(rows, dim=-2).to(torch.complex128)      # (..., 2, 2)


def _apply_1q(state, m, wire):
    """state (B, 2, ..., 2); m (2,2) or (B,2,2) on axis 1+wire."""
    ax = 1 + wire
    st = state.movedim(ax, -1)                                  # (B, ..., 2)
    if m.dim() == 2:
        st = torch.matmul(st, m.transpose(0, 1))
    else:
        shape = st.shape
        st = torch.matmul(st.reshape(shape[0], -1, 2), m.transpose(1, 2)).reshape(shape)
    return st.movedim(-1, ax)


def _apply_controlled(state, m, control, target):
    """|0><0| x I + |1><1| x m  on (control, target)."""
    idx0 = [slice(None)] * state.dim()
    idx1 = [slice(None)] * state.dim()
    idx0[1 + control], idx1[1 + control] = slice(0, 1), slice(1, 2)
    keep = state[tuple(idx0)]
    hit = _apply_1q(state[tuple(idx1)], m, target)
    return torch.cat([keep, hit], dim=1 + control)


def _apply_2q_matrix(state, u, w0, w1):
    """Dense 4x4 on (w0, w1), w0 = most significant bit of the matrix index."""
    u = torch.as_tensor(u, dtype=torch.complex128).reshape(2, 2, 2, 2)       # out0 out1 in0 in1
    st = torch.tensordot(state, u, dims=([1 + w0, 1 + w1], [2, 3]))          # (..., out0, out1) at the end
    return st.movedim((-2, -1), (1 + w0, 1 + w1))


def execute(tape, n, batch):
    """Run a recorded tape on |0...0>; returns the state as (B, 2^n) complex128."""
    state = torch.zeros((batch,) + (2,) * n, dtype=torch.complex128)
    state[(slice(None),) + (0,) * n] = 1.0
    xflip = torch.tensor([[0, 1], [1, 0]], dtype=torch.complex128)
    had = torch.tensor([[1, 1], [1, -1]], dtype=torch.complex128) / math.sqrt(2.0)
    for op in tape:
        if op.name == "AmplitudeEmbedding":
            feats, normalize, pad = op.param
            feats = feats.reshape(-1, feats.shape[-1])
            if feats.shape[-1] < 2 ** n:
                fill = torch.full((feats.shape[0], 2 ** n - feats.shape[-1]), float(pad), dtype=feats.dtype)
                feats = torch.cat([feats, fill], dim=-1)
            if normalize:
                feats = feats / torch.linalg.norm(feats, dim=-1, keepdim=True)
            state = feats.to(torch.complex128).reshape((feats.shape[0],) + (2,) * n)
            if state.shape[0] != batch:
                state = state.expand((batch,) + (2,) * n)
        elif op.name in ("RX", "RY", "RZ", "PhaseShift"):
            state = _apply_1q(state, _rot_matrix(op.name, op.param), op.wires[0])
        elif op.name in ("CRX", "CRZ"):
            state = _apply_controlled(state, _rot_matrix(op.name[1:], op.param), *op.wires)
        elif op.name == "CNOT":
            state = _apply_controlled(state, xflip, *op.wires)
        elif op.name == "CZ":
            state = _apply_controlled(state, torch.tensor([[1, 0], [0, -1]], dtype=torch.complex128),
                                      *op.wires)
        elif op.name == "Hadamard":
            state = _apply_1q(state, had, op.wires[0])
        elif op.name == "QubitUnitary":
            state = _apply_2q_matrix(state, op.matrix, *op.wires)
        else:
            raise KeyError(op.name)
    return state.reshape(batch, -1)


class QNode:
    """``qml.QNode(func, dev, interface="torch", diff_method=...)``: calling it records ``func`` on
    a fresh tape, executes it and returns the measurements in the structure ``func`` returned
    (the reference returns a list of n expvals; each comes back with the broadcast shape (B,) or ()
    as float64, like default.qubit does)."""

    last_tape = None

    def __init__(self, func, dev, interface="torch", diff_method="backprop", **kw):
        self.func, self.device, self.interface, self.diff_method = func, dev, interface, diff_method

    def __call__(self, *args, **kw):
        global _TAPE
        outer, _TAPE = _TAPE, []
        try:
            measurements = self.func(*args, **kw)
            tape = _TAPE
        finally:
            _TAPE = outer
        QNode.last_tape = tape
        n = self.device.num_wires
        x = torch.as_tensor(args[0] if args else kw.get("inputs"))
        batched = x.dim() == 2
        batch = x.shape[0] if batched else 1
        state = execute(tape, n, batch)
        probs = (state.real ** 2 + state.imag ** 2)                        # (B, 2^n) float64
        index = torch.arange(2 ** n)
        out = []
        for m in measurements:
            sign = 1.0 - 2.0 * ((index >> (n - 1 - m.obs.wire)) & 1).to(torch.float64)
            val = probs @ sign
            out.append(val if batched else val[0])
        return out


def qnode(dev, interface="torch", diff_method="backprop", **kw):
    """``@qml.qnode(device, ...)`` decorator form."""
    def wrap(func):
        return QNode(func, dev, interface=interface, diff_method=diff_method, **kw)
    return wrap


class _TorchLayer(torch.nn.Module):
    """Minimal ``qml.qnn.TorchLayer``: owns one Parameter per entry of ``weight_shapes`` (uniform in
    [0, 2 pi) like PennyLane's default init) and calls the QNode with ``inputs=`` plus the weights
    as keyword arguments; the measurement list comes back stacked."""

    def __init__(self, qnode_obj, weight_shapes):
        super().__init__()
        self.qnode = qnode_obj
        for name, shape in weight_shapes.items():
            self.register_parameter(name, torch.nn.Parameter(2 * math.pi * torch.rand(shape)))
        self._names = list(weight_shapes)

    def forward(self, inputs):
        out = self.qnode(inputs=inputs, **{k: getattr(self, k) for k in self._names})
        return torch.stack(list(out), dim=-1) if inputs.dim() == 2 else torch.stack(list(out))


qnn = types.SimpleNamespace(TorchLayer=_TorchLayer)


def install():
    """Register this module as ``pennylane`` (only when the real one is absent)."""
    try:
        import pennylane  # noqa: F401
        return False
    except Exception:
        sys.modules["pennylane"] = sys.modules[__name__]
        return True
