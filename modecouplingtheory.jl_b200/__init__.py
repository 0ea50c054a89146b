"""B200-native hot path of ModeCouplingTheory.jl (memory-kernel evaluation + time-doubling step).

Host-side mirror of the reference interface; all numerics run in libmct_sm100a.so (hand-written sm_100a CUDA
behind a C ABI, include/mct.h).  There is no CPU fallback: calling into the library without the built
extension or without a B200 raises MCTDeviceError.
"""
from . import api, structure_factors, sweep  # noqa: F401
from .api import (  # noqa: F401
    DomainError, MCTDeviceError, MemoryEquation, MemoryEquationSolution, MemoryKernel, ModeCouplingKernel,
    MultiComponentModeCouplingKernel, dDimModeCouplingKernel, ddim_tables, ActiveMCTKernel, active_tables, TaggedActiveMCTKernel, tagged_active_tables,
    TaggedModeCouplingKernel, MSDModeCouplingKernel, TaggedMultiComponentModeCouplingKernel,
    MSDMultiComponentModeCouplingKernel, TimeDoublingSolver, device_count, launch_count, evaluate_kernel, evaluate_kernel_, evaluate_kernel_many,
    find_relaxation_time, get_F, get_K, get_t, library_path, load_library, solve, solve_batch, solve_steady_state,
    vertex_tables, EXPORTS,
)
