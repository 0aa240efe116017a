"""Import shim: ``import qcpinn_b200`` (and ``qcpinn_b200.<sub>``) resolve to the package in
``qcpinn-convection-diffusion-qiskit_b200/`` -- a directory name fixed by the project layout that
is not a valid Python identifier.  A meta-path finder maps every ``qcpinn_b200.*`` name onto the
one canonical module object, so nothing is ever imported twice."""

import importlib
import importlib.abc
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_REAL = "qcpinn-convection-diffusion-qiskit_b200"
_ALIAS = "qcpinn_b200"
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname == _ALIAS or fullname.startswith(_ALIAS + "."):
            return importlib.util.spec_from_loader(fullname, self)
        return None

    def create_module(self, spec):
        return importlib.import_module(_REAL + spec.name[len(_ALIAS):])

    def exec_module(self, module):
        pass


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
sys.modules[_ALIAS] = importlib.import_module(_REAL)
