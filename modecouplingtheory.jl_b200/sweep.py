"""Sharding of many independent memory equations (a packing-fraction sweep) over ranks / GPUs.

The path shards over independent equations only: no collective during the solves, one final
gather of reduced observables.  `partition` is pure host logic (tested on CPU with gloo, world_size 2); `run_sweep`
drives one GPU per rank through `solve_batch` (one persistent launch per batch).
"""
import numpy as np


def partition(costs, world, rank):
    """Indices of the sweep points handled by `rank`: points are sorted by decreasing expected cost and dealt in
    snake order (0..w-1, w-1..0, ...), so every rank gets the same number of points (+-1) and a similar total cost."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    mine = []
    for pos, idx in enumerate(order):
        lap, off = divmod(pos, world)
        owner = off if lap % 2 == 0 else world - 1 - off
        if owner == rank:
            mine.append(int(idx))
    return sorted(mine)


def expected_cost(etas, eta_c=0.5159):
    """Fixed-point iterations per time step grow like 1/|eta - eta_c| near the glass transition (measured: 3.7 per step
    at eta=0.51, 53 at eta=0.5159); above eta_c the solution never decays but every block still runs."""
    etas = np.asarray(etas, dtype=np.float64)
    return 1.0 / np.maximum(np.abs(etas - eta_c), 2e-4)


def gather(local_indices, local_values, n_points, world, dist=None):
    """Final gather of per-point observables (1-D float arrays of equal length) to every rank."""
    local_values = np.asarray(local_values, dtype=np.float64).reshape(len(local_indices), -1)
    width = local_values.shape[1] if len(local_indices) else 1
    if world == 1 or dist is None:
        parts = [(local_indices, local_values)]
    else:
        parts = [None] * world
        dist.all_gather_object(parts, (list(local_indices), local_values))
        width = max([p[1].shape[1] for p in parts if len(p[0])] + [width])
    out = np.full((n_points, width), np.nan)
    for idx, vals in parts:
        for i, v in zip(idx, vals):
            out[i] = v
    return out


def run_sweep(pkg, etas, Nk, device, solver_kwargs=None, batch=None, observable=None, download=True):
    """Solve the hard-sphere MCT equation at every packing fraction in `etas` on one GPU, `batch` equations per
    persistent launch.  `download` is passed to solve_batch: True = whole trajectories, a dict(times=, dofs=) = only that
    slice (the observable then sees the sliced F, K).  Returns (values [len(etas), w], total kernel evaluations, device
    milliseconds)."""
    sf = pkg.structure_factors
    solver_kwargs = dict(solver_kwargs or {})
    solver_kwargs.setdefault("max_iterations", 10 ** 6)
    observable = observable or (lambda sol: np.array([sol.F[-1].max()]))
    k = sf.wavenumber_grid(Nk)
    batch = batch or len(etas)
    values, evals_tot, ms_tot = [], 0, 0.0
    for b0 in range(0, len(etas), batch):
        eqs = []
        for eta in etas[b0:b0 + batch]:
            Sk = sf.percus_yevick_sk(k, eta)
            kern = pkg.ModeCouplingKernel(6 * eta / np.pi, 1.0, 1.0, k, Sk, device=device)
            eqs.append(pkg.MemoryEquation(0.0, 1.0, k ** 2 / Sk, 0.0, Sk, np.zeros(Nk), kern))
        sols, status, evals, ms = pkg.solve_batch(eqs, pkg.TimeDoublingSolver(**solver_kwargs), download=download)
        if any(status):
            raise pkg.DomainError(f"sweep: solver status {status}")
        values += [observable(s) for s in sols]
        evals_tot += int(sum(evals)); ms_tot += ms
        for e in eqs:
            e.kernel.close()
    return np.array(values), evals_tot, ms_tot
