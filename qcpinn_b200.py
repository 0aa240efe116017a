"""Import shim: ``import qcpinn_b200`` -> the package in ``qcpinn-convection-diffusion-qiskit_b200/``
(whose directory name, fixed by the project layout, is not a valid Python identifier)."""

import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("qcpinn-convection-diffusion-qiskit_b200")
sys.modules[__name__] = _pkg
