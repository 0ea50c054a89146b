"""Kernel evaluations per sweep point (config 5: 512 packing fractions in [0.50, 0.53], Nk=500) -- the cost table behind
modecouplingtheory.jl_b200/sweep.py:expected_cost.  Writes profiles/sweep_costs_r02.json."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from __graft_entry__ import load_pkg
pkg = load_pkg()
sf = pkg.structure_factors
Nk, points = 500, int(sys.argv[1]) if len(sys.argv) > 1 else 512
etas = np.linspace(0.50, 0.53, points)
k = sf.wavenumber_grid(Nk)
out = {"etas": etas.tolist(), "evals": [], "f_end": []}
t0 = time.time()
for b0 in range(0, points, 64):
    eqs = []
    for eta in etas[b0:b0 + 64]:
        Sk = sf.percus_yevick_sk(k, eta)
        kern = pkg.ModeCouplingKernel(6 * eta / np.pi, 1.0, 1.0, k, Sk)
        eqs.append(pkg.MemoryEquation(0.0, 1.0, k ** 2 / Sk, 0.0, Sk, np.zeros(Nk), kern))
    nt = pkg.TimeDoublingSolver().output_len()
    sols, status, evals, ms = pkg.solve_batch(eqs, pkg.TimeDoublingSolver(max_iterations=10 ** 6), download=dict(times=[0, nt - 1], dofs=[87], K=False))
    out["evals"] += [int(e) for e in evals]
    out["f_end"] += [float(s.F[-1, 0] / s.F[0, 0]) for s in sols]
    print(f"batch {b0}: {ms:.0f} ms, evals {min(evals)}..{max(evals)}", flush=True)
    for e in eqs: e.kernel.close()
out["wall_s"] = time.time() - t0
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep_costs_r02.json", "w"))
ev = np.array(out["evals"])
print("total evals", ev.sum(), "max", ev.max(), "at eta", etas[ev.argmax()], "wall", out["wall_s"])
