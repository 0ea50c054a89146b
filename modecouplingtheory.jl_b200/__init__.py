"""B200-native hot path of ModeCouplingTheory.jl (memory-kernel evaluation + time-doubling step).

Host-side mirror of the reference interface; all numerics run in libmct_sm100a.so (hand-written sm_100a CUDA
behind a C ABI, include/mct.h).  There is no CPU fallback: calling into the library without the built
extension or without a GPU raises.
"""
from . import structure_factors  # noqa: F401
