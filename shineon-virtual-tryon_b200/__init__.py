"""shineon-virtual-tryon_b200 — B200-native (sm_100a) try-on hot path of ShineOn-Virtual-Tryon.

Layout:
  csrc/        hand-written CUDA kernels + the C ABI (include/shineon_b200.h) -> libshineon_b200.so
  build.py     nvcc recipe (in-tree .so, no JIT cache)
  _lib.py      ctypes binding of the C ABI (fails loudly when the library is missing)
  ops.py       torch-tensor wrappers around the C ABI (allocation + stream plumbing only)
  networks/    host-side mirrors of the reference nn.Modules (same names / ctor args / state_dict keys)
  models/      WarpModel / UnetMaskModel mirrors (forward / step signatures of the reference)
"""
__version__ = "0.1.0"
