"""Run logger with the reference's ``Logging`` surface (reference utils/logger.py:5-52).

``DVPDESolver`` and the trainer only need ``print(*args)`` and ``get_output_dir()``; each instance
writes to ``<log_path>/<timestamp>[_<experiment>]/output.log``.
"""

from __future__ import annotations

import logging
import os
from datetime import datetime

import numpy as np

_FLOATS = (float, np.floating)


class Logging:
    def __init__(self, log_path, experiment_name=None, source_file=None):
        self.log_path = log_path
        self.experiment_name = experiment_name
        stamp = datetime.now().strftime("%Y-%m-%d_%H-%M-%S-%f")
        if experiment_name is not None:
            stamp = f"{stamp}_{experiment_name}"
        self.output_dir = os.path.join(log_path, stamp)
        os.makedirs(self.output_dir, exist_ok=True)
        # one logger object per run directory so two models never share handlers
        self.logger = logging.getLogger(f"qcpinn_b200.{stamp}.{id(self)}")
        self.logger.setLevel(logging.DEBUG)
        self.logger.propagate = False
        self._handler = logging.FileHandler(os.path.join(self.output_dir, "output.log"), mode="w")
        self.logger.addHandler(self._handler)

    def get_output_dir(self):
        return self.output_dir

    @staticmethod
    def _fmt(value):
        return "%.4e" % value if isinstance(value, _FLOATS) else value

    def print(self, *args):
        """Several arguments go on one line (no separator, like the reference); floats as %.4e."""
        if not args:
            return
        if len(args) == 1:
            self.logger.info(args[0])
            return
        self._handler.terminator = ""
        for value in args[:-1]:
            self.logger.info(self._fmt(value))
        self._handler.terminator = "\n"
        self.logger.info(self._fmt(args[-1]))

    def close(self):
        self.logger.removeHandler(self._handler)
        self._handler.close()
