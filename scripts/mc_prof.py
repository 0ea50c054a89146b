"""Config 4 (Ns=4, Nk=2000) on one GPU, short time axis: evaluations/s and the per-phase cycle counts of td_grid_kernel
(MCT_GRID_PROF=1: slots 0 G matrices, 1 line sums, 2 prefix, 3 K (evaluate only), 4 convolution, 5 update).
usage: MCT_GRID_PROF=1 python scripts/mc_prof.py [Nk] [t_max] [N] [dt] [passes]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_pkg
pkg = load_pkg()
sf = pkg.structure_factors
ns = 4
Nk = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
tmax = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-2
N = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dt = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-4
passes = int(sys.argv[5]) if len(sys.argv) > 5 else 4
diam = np.array([0.7, 0.8, 0.9, 1.0]); x = np.full(ns, 0.25)
rho = 6 * 0.50 / (np.pi * np.sum(x * diam ** 3)) * x
m = np.ones(ns)
k = sf.wavenumber_grid(Nk)
Sk = sf.percus_yevick_mixture_sk(k, diam, rho)
xx = rho / rho.sum()
gamma = np.einsum("kab,kbc->kac", np.array([np.diag(k[i] ** 2 * xx / m) for i in range(Nk)]), np.linalg.inv(Sk))
kern = pkg.MultiComponentModeCouplingKernel(rho, 1.0, m, k, Sk, device=0)
eq = pkg.MemoryEquation(1.0, 0.0, gamma, np.zeros((Nk, ns, ns)), Sk, np.zeros((Nk, ns, ns)), kern)
solver = pkg.TimeDoublingSolver(N=N, dt=dt, t_max=tmax, tolerance=1e-10, max_iterations=10 ** 6)
h = pkg.api.create_equation(eq, solver)
for i in range(passes):
    evals, ms = pkg.api.solve_resident(h)
    tot = evals + 2 * solver.N + 1
    print(f"pass {i}: {tot} evaluations, {ms:.3f} ms, {1e3 * ms / tot:.2f} us per evaluation, {tot / ms:.2f} k evaluations/s", flush=True)
F = pkg.api.download_solution(h, solver, kern.dof).F
print("checksum", float(np.sum(F[-1])), float(np.abs(F[-1]).max()))
