"""scbonita_b200 -- B200-native replacement of scBONITA's rule-inference fitness hot path.

The package is deliberately small: the CUDA library (``csrc/`` -> ``libscbonita_b200.so``, C-ABI in
``include/scbonita_b200.h``) and the host-side mirror of the reference's Python boundary.

* :mod:`scbonita_b200.simulator`   -- ``Simulator`` (context + population scoring, local search, attractors,
  decider stage, importance score) and ``ModelTables``;
* :mod:`scbonita_b200.distributed` -- sharding of a population over GPUs with one fitness all-gather;
* :mod:`scbonita_b200.synth`       -- seeded synthetic problems of the BASELINE shapes;
* :mod:`scbonita_b200._lib`        -- the ctypes binding (no CPU fallback: it raises when the library or a CUDA
  device is missing).
"""
__version__ = "0.1.0"

from .simulator import MODE_CELLWISE, MODE_NODEWISE, MODE_SUM, ModelTables, ScbError, Simulator  # noqa: F401
