"""``DVQuantumLayer`` -- drop-in for reference nn/DVQuantumLayer.py backed by sm_100a kernels.

Same constructor, same ``params`` parameter ((L, P) float32, state-dict key ``params``), same
``forward(x: (B, n)) -> (n, B)`` orientation, same error behaviour for bad ansatz names, the
``alternate``/even-n ``IndexError`` and the Haar rule (``seed`` present and n >= 4).  What is
different: the circuit is never handed to PennyLane; it is compiled to a gate table
(:mod:`..program`) and evaluated by the CUDA library.  The IBM-hardware branch of the reference
(``:97-141``) is out of scope and raises ``NotImplementedError``.

New optional ``args`` keys (defaults keep a reference dict working unchanged):
``dtype`` = ``"float64"`` (default; complex128-equivalent arithmetic like ``default.qubit``) or
``"float32"``; ``diff_mode`` = ``"kernels"`` (default) or ``"autograd"`` (opt-in slow path that is
differentiable to any order, :mod:`.torch_circuit`).
"""

from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as F
from ..program import compile_program, params_per_layer

_DTYPES = {
    "float64": torch.float64, "double": torch.float64, "f64": torch.float64,
    torch.float64: torch.float64,
    "float32": torch.float32, "float": torch.float32, "f32": torch.float32,
    torch.float32: torch.float32,
}


def resolve_dtype(value) -> torch.dtype:
    try:
        return _DTYPES[value]
    except KeyError:
        raise ValueError(f"unsupported dtype {value!r}; use 'float64' or 'float32'") from None


class DVQuantumLayer(nn.Module):
    def __init__(self, args, diff_method="parameter-shift"):
        super().__init__()
        self.num_qubits = args["num_qubits"]
        self.num_quantum_layers = args["num_quantum_layers"]
        self.shots = args.get("shots", 1024)
        self.q_ansatz = args["q_ansatz"]
        self.problem = args["problem"]
        self.encoding = args.get("encoding", "angle")
        self.use_ibm_hardware = args.get("use_ibm_hardware", False)
        self.compute_dtype = resolve_dtype(args.get("dtype", "float64"))
        self.diff_mode = args.get("diff_mode", "kernels")
        if self.diff_mode not in ("kernels", "autograd"):
            raise ValueError(f"diff_mode must be 'kernels' or 'autograd', got {self.diff_mode!r}")
        if self.use_ibm_hardware:
            raise NotImplementedError(
                "use_ibm_hardware=True (remote QPU execution) is outside the B200 hot path")

        count = params_per_layer(self.q_ansatz, self.num_qubits)   # ValueError on unknown ansatz
        self.params = nn.Parameter(
            torch.empty(self.num_quantum_layers, count, dtype=torch.float32))
        self._initialize_weights()

        seed = args.get("seed", None) if self.num_qubits >= 4 else None
        if self.q_ansatz == "cz_melt":
            seed = None          # the re-uploading family has no Haar blocks (reference 16-qubit script)
        self.haar_seed1 = seed
        self.haar_seed2 = seed + 1 if seed is not None else None
        self.use_batch_processing = True
        self._program = None
        self._plans = {}
        self._epoch = 0     # bumped by DVPDESolver after every optimizer step (cache keys)

    def _initialize_weights(self):
        torch.nn.init.xavier_normal_(self.params)

    # -- compiled circuit / plans --------------------------------------------------------------
    @property
    def program(self):
        if self._program is None:   # compiled lazily so construction never fails (like the QNode)
            self._program = compile_program(
                self.q_ansatz, self.num_qubits, self.num_quantum_layers, self.haar_seed1,
                getattr(self, "program_variant", "dv"))
        return self._program

    def plan(self, device, hidden=1, io_dtype=None) -> F.Plan:
        key = (str(device), int(hidden), self.compute_dtype, io_dtype)
        plan = self._plans.get(key)
        if plan is None:
            plan = F.Plan(self.program, F.encoding_code(self.encoding), self.compute_dtype,
                          hidden, torch.device(device), io_dtype=io_dtype)
            self._plans[key] = plan
        return plan

    def theta_key(self):
        return (self._epoch, self.params.data_ptr(), self.params._version, self.params.device.index)

    def mark_updated(self):
        """Call after parameters changed in a way that may not bump ``_version`` counters."""
        self._epoch += 1

    def describe(self) -> str:
        names = ["RX", "RY", "RZ", "CRX", "CRZ", "CNOT", "H", "U4"]
        lines = [f"{self.encoding if self.encoding == 'amplitude' else 'angle'} encoding on "
                 f"{self.num_qubits} wires"]
        for kind, a, b, p in self.program.ops.tolist():
            wires = f"[{a}]" if b < 0 else f"[{a},{b}]"
            lines.append(f"  {names[kind]}{wires}" + (f" theta[{p}]" if p >= 0 and kind != 7 else ""))
        lines.append("  measure <Z_i> on every wire")
        return "\n".join(lines)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        single = x.dim() == 1
        z = x.reshape(1, -1) if single else x
        if z.dim() != 2 or z.shape[1] != self.num_qubits:
            raise ValueError(
                f"Features must be of length {self.num_qubits}; got shape {tuple(x.shape)}")
        if self.params.device != z.device:
            raise RuntimeError(
                f"quantum_layer.params on {self.params.device} but input on {z.device}; "
                "move the module with .to(device)")
        if self.diff_mode == "autograd":
            from .torch_circuit import run_layer

            if z.device.type != "cuda":        # same rule as the kernels: this package has no CPU path
                raise RuntimeError(
                    f"qcpinn_b200 runs on CUDA devices only (got {z.device}); there is no CPU path")

            cdtype = torch.complex128 if self.compute_dtype == torch.float64 else torch.complex64
            out = run_layer(self.program, self.encoding, z, self.params, cdtype)
        else:
            plan = self.plan(z.device)
            out = F.layer_apply(plan, z, self.params, self.theta_key())
        return out[:, 0] if single else out
