"""qcpinn_b200 -- B200 (sm_100a) implementation of the QCPINN convection-diffusion training hot
path: ``DVQuantumLayer`` / ``DVPDESolver`` / ``diffusion_operator`` behind the reference's API.

The directory is named ``qcpinn-convection-diffusion-qiskit_b200`` (not an identifier); import it
through the ``qcpinn_b200`` shim at the repository root, or call
:func:`install_reference_aliases` to expose the reference's own top-level module names
(``nn``, ``data``, ``utils``, ``trainer``) so reference scripts run unchanged.
"""

import importlib
import sys

from . import _lib, functional, program  # noqa: F401
from .data import diffusion_dataset  # noqa: F401
from .nn.DVPDESolver import DVPDESolver  # noqa: F401
from .nn.DVQuantumLayer import DVQuantumLayer  # noqa: F401
from .nn.pde import (diffusion_operator, helmholtz_operator, klein_gordon_operator,  # noqa: F401
                     navier_stokes_2D_operator, wave_operator)
from .utils.logger import Logging  # noqa: F401

__version__ = "0.1.0"

_ALIASES = {
    "nn": "nn", "nn.DVQuantumLayer": "nn.DVQuantumLayer", "nn.DVPDESolver": "nn.DVPDESolver",
    "nn.pde": "nn.pde",
    "data": "data", "data.diffusion_dataset": "data.diffusion_dataset",
    "utils": "utils", "utils.logger": "utils.logger",
    "trainer": "trainer", "trainer.diffusion_train": "trainer.diffusion_train",
    "trainer.diffusion_eval": "trainer.diffusion_eval",
}


def install_reference_aliases(force=False):
    """Register this package's sub-modules under the reference's top-level import names, e.g.
    ``from nn.pde import diffusion_operator`` (reference trainer/diffusion_train.py:3-4)."""
    for alias, rel in _ALIASES.items():
        if alias in sys.modules and not force:
            continue
        sys.modules[alias] = importlib.import_module(f"{__name__}.{rel}")


def build(force=False, verbose=False):
    from .build import build_library

    return build_library(force=force, verbose=verbose)
