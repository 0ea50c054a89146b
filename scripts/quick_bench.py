"""Quick device timing of single solves (development aid; bench.py is the contract benchmark)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from conftest import load_pkg
import problems as P
pkg = load_pkg()

def run(Nk, eta, **kw):
    p = P.hard_spheres(Nk, eta)
    e = P.overdamped(p)
    kern = pkg.ModeCouplingKernel(p["rho"], 1.0, 1.0, p["k"], p["Sk"])
    eq = pkg.MemoryEquation(e["alpha"], e["beta"], e["gamma"], e["delta"], e["F0"], e["dF0"], kern)
    solver = pkg.TimeDoublingSolver(max_iterations=10**6, **kw)
    t0 = time.time(); sol = pkg.solve(eq, solver); wall = time.time() - t0
    ev = solver.kernel_evals
    print(f"Nk={Nk} eta={eta} team={os.environ.get('MCT_TEAM_SIZE','auto')} evals={ev} device_ms={solver.device_ms:.2f} wall_ms={wall*1e3:.1f} "
          f"us/eval={solver.device_ms*1e3/ev:.3f} GB/s(24Nk^2)={ev*24*Nk*Nk/solver.device_ms/1e6:.0f} F_end_max={sol.F[-1].max():.3e}", flush=True)

if __name__ == "__main__":
    for Nk, eta in ((100, 0.51), (500, 0.51), (1000, 0.51), (1000, 0.515)):
        run(Nk, eta)
    for team in (8, 32, 74, 148):
        os.environ["MCT_TEAM_SIZE"] = str(team)
        run(1000, 0.51)
    for team in (1, 2, 4, 8):
        os.environ["MCT_TEAM_SIZE"] = str(team)
        run(100, 0.51)
