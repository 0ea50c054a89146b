"""Where the host time of a sweep goes (development aid): 128 points at Nk=500."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from __graft_entry__ import load_pkg
pkg = load_pkg(); sf = pkg.structure_factors
Nk = 500; k = sf.wavenumber_grid(Nk); etas = np.linspace(0.50, 0.512, 128)
T = {}
def tick(name, t0): T[name] = T.get(name, 0.0) + time.perf_counter() - t0
eqs = []
for eta in etas:
    t0 = time.perf_counter(); Sk = sf.percus_yevick_sk(k, eta); tick("S(k)", t0)
    t0 = time.perf_counter(); kern = pkg.ModeCouplingKernel(6 * eta / np.pi, 1.0, 1.0, k, Sk); tick("kernel create", t0)
    t0 = time.perf_counter(); eqs.append(pkg.MemoryEquation(0.0, 1.0, k ** 2 / Sk, 0.0, Sk, np.zeros(Nk), kern)); tick("MemoryEquation", t0)
solver = pkg.TimeDoublingSolver(max_iterations=10 ** 6)
t0 = time.perf_counter(); devs = [pkg.api._create_equation(e, solver) for e in eqs]; tick("equation create (device)", t0)
import ctypes as C
L = pkg.load_library(); n = len(devs)
arr = (C.c_void_p * n)(*[d._q for d in devs]); status = (C.c_int * n)(); evals = (C.c_long * n)(); ms = C.c_float(0)
t0 = time.perf_counter(); L.mct_equation_solve_batch(n, arr, status, evals, C.byref(ms)); tick("solve_batch call", t0)
nt = solver.output_len()
t0 = time.perf_counter()
for d, e in zip(devs, eqs): pkg.api.download_slice(d, e.kernel.dof, [0, nt - 1], [87], False)
tick("download slices", t0)
t0 = time.perf_counter()
for d in devs: d.close()
tick("equation destroy", t0)
t0 = time.perf_counter()
for e in eqs: e.kernel.close()
tick("kernel destroy", t0)
print(f"device ms {ms.value:.0f}")
for k_, v in T.items(): print(f"{k_:28s} {1e3 * v:9.1f} ms total  {1e3 * v / n:7.2f} ms per equation")
