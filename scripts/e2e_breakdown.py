"""Wall-clock breakdown of the host-facing solve (create / solve / download / destroy) at the bench configuration."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from conftest import load_pkg
pkg = load_pkg()
sf = pkg.structure_factors
Nk, eta = 1000, 0.5155
k = sf.wavenumber_grid(Nk); Sk = sf.percus_yevick_sk(k, eta)
kern = pkg.ModeCouplingKernel(6 * eta / np.pi, 1.0, 1.0, k, Sk)
eq = pkg.MemoryEquation(0.0, 1.0, k ** 2 / Sk, 0.0, Sk, np.zeros(Nk), kern)
solver = pkg.TimeDoublingSolver(max_iterations=10 ** 6)
for rep in range(4):
    t0 = time.perf_counter(); q = pkg.api.create_equation(eq, solver); t1 = time.perf_counter()
    ev, ms = pkg.api.solve_resident(q); t2 = time.perf_counter()
    sol = pkg.api.download_solution(q, solver, Nk); t3 = time.perf_counter()
    q.close(); t4 = time.perf_counter()
    print(f"rep {rep}: create {1e3*(t1-t0):.1f} ms, solve wall {1e3*(t2-t1):.1f} ms (device {ms:.1f}), download(pageable) {1e3*(t3-t2):.1f} ms, destroy {1e3*(t4-t3):.1f} ms")
