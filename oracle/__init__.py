"""CPU oracle for the QCPINN training hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``qcpinn-convection-diffusion-qiskit_b200``) never does and fails loudly without its CUDA library.

What it restates (citations are relative to the reference checkout):

* ``nn/DVQuantumLayer.py:25-78,88-94,176-214,246-371``  -> :mod:`oracle.circuits`
* ``nn/DVPDESolver.py:28-51,69-76,81-110``               -> :mod:`oracle.solver`
* ``nn/pde.py:53-72``                                    -> :func:`oracle.solver.diffusion_operator`
* ``data/diffusion_dataset.py:5-38``                     -> :mod:`oracle.dataset`
* ``trainer/diffusion_train.py:30-49,81-90``             -> :func:`oracle.solver.train_step`

The arithmetic of the reference lives in a third-party dependency that is neither vendored in the
reference nor installable here: PennyLane ``default.qubit`` (pulled in by
``pennylane-qiskit==0.44.1``, reference ``requirements-dev.txt:1``) with ``interface="torch"``,
``diff_method="backprop"``.  This oracle restates PennyLane's *published* gate definitions
(RX/RY/RZ/CRX/CRZ/CNOT/Hadamard/QubitUnitary/AngleEmbedding/AmplitudeEmbedding/expval(PauliZ),
wire 0 = most significant bit) as a gate-by-gate batched statevector simulation in torch with plain
nested autograd -- the same algorithmic structure ``default.qubit`` backprop executes.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path
 and cannot be run here, so the oracle is pinned only by hand-derivable
known-answer tests (tests/test_oracle_kats.py), finite-difference derivative checks, unitarity /
norm invariants and scipy reproducibility of the Haar unitaries.
"""

from . import circuits, dataset, solver  # noqa: F401
