"""The kernels of libmct_sm100a.so that really stream HBM, one launch each after an L2 flush (for ncu captures and the
GB/s figures in profiles/README.md):
  sc3d_line_partials + sc3d_finalize   stand-alone evaluate_kernel! at Nk = 1000 and 2000 (24 Nk^2 bytes of tables)
  td_grid_kernel (dDim mode)           dDimModeCouplingKernel evaluation at Nk = 192 (8 Nk^3 bytes of fused table)
  msd_table_kernel                     MSD kernel table from the resident collective + tagged trajectories at config-3 size
Prints device milliseconds where the ABI returns them (mct_kernel_eval_many)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from __graft_entry__ import load_pkg
import torch
pkg = load_pkg(); sf = pkg.structure_factors
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
REPS = 1 if os.environ.get("NCU") else 5        # under ncu one launch per case is enough
out = {}
rng = np.random.default_rng(0)
for Nk in (1000, 2000):
    eta = 0.51; k = sf.wavenumber_grid(Nk); Sk = sf.percus_yevick_sk(k, eta)
    kern = pkg.ModeCouplingKernel(6 * eta / np.pi, 1.0, 1.0, k, Sk)
    F = rng.random((1, Nk))
    pkg.evaluate_kernel_many(kern, F)                    # builds the raw tables, warm-up
    ms = []
    for _ in range(REPS):
        flush.fill_(1); torch.cuda.synchronize()
        ms.append(pkg.evaluate_kernel_many(kern, F)[1])
    bytes_ = 24.0 * Nk * Nk
    out[f"sc3d_eval_Nk{Nk}"] = {"ms_median": float(np.median(ms)), "algorithmic_bytes": bytes_, "GB_per_s": bytes_ / (np.median(ms) * 1e-3) / 1e9}
    F8 = rng.random((8, Nk))
    m8 = pkg.evaluate_kernel_many(kern, F8)[1]
    out[f"sc3d_eval_Nk{Nk}_x8_L2_resident"] = {"ms": m8, "GB_per_s_algorithmic": 8 * bytes_ / (m8 * 1e-3) / 1e9}
    kern.close()
# dDim
Nk = 192; eta = 0.45; k = sf.wavenumber_grid(Nk); Sk = sf.percus_yevick_sk(k, eta)
kd = pkg.dDimModeCouplingKernel(6 * eta / np.pi, 1.0, 1.0, k, Sk, 3)
F = rng.random(Nk)
pkg.evaluate_kernel(kd, F)
for _ in range(1 if os.environ.get("NCU") else 3):     # timed by ncu (gpu__time_duration): the ABI returns no device time here
    flush.fill_(1); torch.cuda.synchronize()
    pkg.evaluate_kernel(kd, F)
out[f"ddim_eval_Nk{Nk}"] = {"algorithmic_bytes": 8.0 * Nk ** 3}
kd.close()
# MSD table at config-3 size (collective -> tagged -> MSD, everything resident)
Nk = 1000; eta = 0.5155; k = sf.wavenumber_grid(Nk); Sk = sf.percus_yevick_sk(k, eta); rho = 6 * eta / np.pi
kern = pkg.ModeCouplingKernel(rho, 1.0, 1.0, k, Sk)
solver = pkg.TimeDoublingSolver(max_iterations=10 ** 6)
sol = pkg.solve(pkg.MemoryEquation(0.0, 1.0, k ** 2 / Sk, 0.0, Sk, np.zeros(Nk), kern), solver)
tk = pkg.TaggedModeCouplingKernel(rho, 1.0, 1.0, k, Sk, sol)
sols = pkg.solve(pkg.MemoryEquation(0.0, 1.0, k ** 2, 0.0, np.ones(Nk), np.zeros(Nk), tk), pkg.TimeDoublingSolver(max_iterations=10 ** 6))
flush.fill_(1); torch.cuda.synchronize()
mk = pkg.MSDModeCouplingKernel(rho, 1.0, 1.0, k, Sk, sol, sols)
msd = pkg.solve(pkg.MemoryEquation(0.0, 1.0, 0.0, -6.0, 0.0, 0.0, mk), pkg.TimeDoublingSolver(max_iterations=10 ** 6))
out["msd_table_config3"] = {"nt": len(sol.t), "algorithmic_bytes": 2.0 * len(sol.t) * Nk * 8, "msd_end": float(np.ravel(msd.F)[-1])}
print(json.dumps(out, indent=1))
