"""trx2dyn: B200-native folding hot path of trRosettaX2-Dynamics.

Host side (Python, as the reference's folding/ is Python) over a C-ABI CUDA
library (csrc/ -> libtrx2dyn.so, declared in include/trx2dyn.h).  No CPU fallback:
anything that computes goes through the CUDA library and fails loudly without it.
"""
__version__ = "0.1.0"
