"""One Nk=1000 hard-sphere solve (eta=0.51, defaults) for ncu captures: warm-up solve + one profiled solve."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from __graft_entry__ import load_pkg
pkg = load_pkg()
sf = pkg.structure_factors
Nk, eta = 1000, float(os.environ.get("ETA", "0.51"))
k = sf.wavenumber_grid(Nk); Sk = sf.percus_yevick_sk(k, eta)
kern = pkg.ModeCouplingKernel(6 * eta / np.pi, 1.0, 1.0, k, Sk)
eq = pkg.MemoryEquation(0.0, 1.0, k ** 2 / Sk, 0.0, Sk, np.zeros(Nk), kern)
solver = pkg.TimeDoublingSolver(max_iterations=10 ** 6)
dev = pkg.api.create_equation(eq, solver)
for i in range(2):
    ev, ms = pkg.api.solve_resident(dev)
    print(f"solve {i}: kernel_evals={ev} device_ms={ms:.2f} us/eval={1e3 * ms / ev:.3f}", flush=True)
