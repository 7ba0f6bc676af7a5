"""B200-native pressure-correction path of freeCappuccino (assembly + Krylov solve).

Host-side mirror of the reference interface lives in :mod:`freecappuccino_b200.solver`;
the compute path is the C-ABI CUDA library built from ``csrc/`` (``libfcapp_cuda.so``).
There is no CPU fallback: importing :mod:`freecappuccino_b200.lib` fails loudly when the
library is missing.
"""
__version__ = "0.1.0"
