"""Sharding of many independent memory equations (a packing-fraction sweep) over ranks / GPUs.

The path shards over independent equations only: no collective during the solves, one final
gather of reduced observables.  `partition` / `expected_cost` are pure host logic (tested on CPU with gloo, world_size 2);
`run_sweep` drives one GPU per rank through `solve_batch` (persistent launches, teams pull equations from a queue).

Load balance (round 2).  The cost of a point is its number of kernel evaluations, which varies by a factor 150 over the
sweep of BASELINE config 5 (profiles/sweep_costs_r02.json, measured on B200): 11 k evaluations deep in the liquid, 140 k
just below the glass transition eta_c, 1.7 M just above it and 200-300 k deep in the glass.  Three measures:
  * eta_c is LOCATED on the grid at hand (bisection with the device's steady-state solver, ~20 cheap solves) instead of
    being hard-coded, and the cost model is a log-log interpolation of the measured evaluations against |eta - eta_c|;
  * points are dealt to ranks longest-first onto the least loaded rank (LPT), and inside a rank the queue of a launch is
    sorted by decreasing cost, so the long solves start first and the short ones fill the gaps;
  * points whose cost exceeds what a rank's fair share of the whole sweep leaves room for ("outliers") are solved one at
    a time by a whole-GPU team (lowest latency per evaluation) instead of by one of the small throughput teams.
"""
import numpy as np

# measured kernel evaluations of TimeDoublingSolver() defaults at Nk = 500 against |eta - eta_c| (profiles/sweep_costs_r02.json)
_LIQUID = [(2.94e-05, 138482), (8.81e-05, 91140), (1.737e-04, 68675), (3.46e-04, 51133), (6.913e-04, 38034), (1.4088e-03, 28222),
           (2.8774e-03, 21156), (5.7882e-03, 16179), (1.16293e-02, 12622), (3e-02, 11000)]
_GLASS = [(2.94e-05, 1725546), (8.81e-05, 1147266), (1.737e-04, 906073), (3.46e-04, 713534), (6.647e-04, 567963), (1.3224e-03, 446552),
          (2.6516e-03, 349480), (5.2698e-03, 273932), (1.04866e-02, 214130), (3e-02, 190000)]
ETA_C_NK500 = 0.515822      # located on the Nk = 500 grid (between 0.515793 and 0.515851); only the default of expected_cost


def expected_cost(etas, eta_c=ETA_C_NK500):
    """Expected kernel evaluations of the time-doubling solve at every packing fraction (log-log interpolation of the
    measured table, clamped at its ends; closer to eta_c than the first knot the cost keeps growing like the last slope)."""
    etas = np.asarray(etas, dtype=np.float64)
    eps = etas - eta_c
    out = np.empty_like(etas)
    for side, table in ((eps < 0, _LIQUID), (eps >= 0, _GLASS)):
        if not side.any():
            continue
        lx, ly = np.log([a for a, _ in table]), np.log([b for _, b in table])
        x = np.log(np.maximum(np.abs(eps[side]), 1e-7))
        y = np.interp(x, lx, ly)
        slope = (ly[1] - ly[0]) / (lx[1] - lx[0])
        y = np.where(x < lx[0], ly[0] + slope * (x - lx[0]), y)
        out[side] = np.exp(y)
    return out


def partition(costs, world, rank):
    """Indices of the sweep points handled by `rank`: longest processing time first onto the least loaded rank (ties: the
    lowest rank), so the total expected cost per rank differs by less than one point's cost.  Deterministic: every rank
    computes the same assignment."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world)
    mine = []
    for idx in order:
        r = int(np.argmin(load))
        load[r] += costs[idx]
        if r == rank:
            mine.append(int(idx))
    return sorted(mine)


# measured on B200 at Nk = 500 (profiles/README.md, round 2): a GPU full of small throughput teams finishes 2.3 M kernel
# evaluations per second in total, each team needs ~5 us per evaluation; one equation alone on a whole-GPU team needs ~3.1 us
THROUGHPUT_EVALS_PER_S = 2.3e6
LATENCY_SMALL_TEAM_S = 5.0e-6
LATENCY_WHOLE_GPU_S = 3.1e-6


def outliers(costs, world):
    """Points too expensive for a throughput team: a point that alone on a small team would take longer than the whole
    sweep's fair share of a rank is solved by a whole-GPU team instead (fewer microseconds per evaluation, but the GPU does
    nothing else meanwhile).  Only worthwhile when the sweep is spread over several ranks."""
    costs = np.asarray(costs, dtype=np.float64)
    if world <= 1:
        return []
    makespan = costs.sum() / (world * THROUGHPUT_EVALS_PER_S)
    return [int(i) for i in np.argsort(-costs, kind="stable") if costs[i] * LATENCY_SMALL_TEAM_S > 1.1 * makespan][:world]


def weighted_costs(costs, world):
    """Costs in units of GPU time for the partition: a whole-GPU-team point occupies its GPU alone, i.e. it costs as much as
    LATENCY_WHOLE_GPU_S * THROUGHPUT_EVALS_PER_S (~7) throughput-team evaluations per evaluation."""
    w = np.asarray(costs, dtype=np.float64).copy()
    for i in outliers(costs, world):
        w[i] *= LATENCY_WHOLE_GPU_S * THROUGHPUT_EVALS_PER_S
    return w


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


def locate_eta_c(pkg, Nk, device, lo=0.50, hi=0.53, steps=16):
    """Glass transition of the hard-sphere Percus-Yevick MCT on this wavenumber grid: bisection on the non-ergodicity
    parameter f(k_peak) = F_inf/S of solve_steady_state (src/SteadyStateMemoryEquation.jl:59-100), on the device."""
    sf = pkg.structure_factors
    k = sf.wavenumber_grid(Nk)

    def is_glass(eta):
        Sk = sf.percus_yevick_sk(k, eta)
        kern = pkg.ModeCouplingKernel(6 * eta / np.pi, 1.0, 1.0, k, Sk, device=device)
        try:
            ss = pkg.solve_steady_state(k ** 2 / Sk, Sk, kern, tolerance=1e-9, max_iterations=10 ** 6)
        finally:
            kern.close()
        return float(np.max(ss.F[0] / Sk)) > 0.1

    for _ in range(steps):
        mid = 0.5 * (lo + hi)
        if is_glass(mid):
            hi = mid
        else:
            lo = mid
    return 0.5 * (lo + hi)


def run_sweep(pkg, etas, Nk, device, solver_kwargs=None, batch=None, observable=None, download=True, costs=None, big_team=()):
    """Solve the hard-sphere MCT equation at every packing fraction in `etas` on one GPU.  The points go into persistent
    launches of `batch` equations (default: all of them in one), sorted by decreasing `costs`; the positions listed in
    `big_team` are solved first, one at a time, by a whole-GPU team.  `download` is passed to solve_batch: True = whole
    trajectories, a dict(times=, dofs=) = only that slice.  Returns (values [len(etas), w], kernel evaluations per point,
    device milliseconds)."""
    import os
    sf = pkg.structure_factors
    solver_kwargs = dict(solver_kwargs or {})
    solver_kwargs.setdefault("max_iterations", 10 ** 6)
    observable = observable or (lambda sol: np.array([sol.F[-1].max()]))
    k = sf.wavenumber_grid(Nk)
    n = len(etas)
    costs = np.ones(n) if costs is None else np.asarray(costs, dtype=np.float64)
    values, evals_pt, ms_tot = [None] * n, np.zeros(n, dtype=np.int64), 0.0

    def equation(eta):
        Sk = sf.percus_yevick_sk(k, eta)
        kern = pkg.ModeCouplingKernel(6 * eta / np.pi, 1.0, 1.0, k, Sk, device=device)
        return pkg.MemoryEquation(0.0, 1.0, k ** 2 / Sk, 0.0, Sk, np.zeros(Nk), kern)

    import time
    timing = run_sweep.last_timing = {"build_equations_s": 0.0, "solve_batch_call_s": 0.0, "release_s": 0.0}

    def launch(positions):
        nonlocal ms_tot
        t0 = time.perf_counter()
        eqs = [equation(etas[i]) for i in positions]
        t1 = time.perf_counter()
        sols, status, evals, ms = pkg.solve_batch(eqs, pkg.TimeDoublingSolver(**solver_kwargs), download=download)
        t2 = time.perf_counter()
        if any(status):
            raise pkg.DomainError(f"sweep: solver status {status}")
        for i, s, ev in zip(positions, sols, evals):
            values[i] = observable(s); evals_pt[i] = int(ev)
        ms_tot += ms
        for e in eqs:
            e.kernel.close()
        t3 = time.perf_counter()
        timing["build_equations_s"] += t1 - t0; timing["solve_batch_call_s"] += t2 - t1; timing["release_s"] += t3 - t2

    big = [i for i in big_team if 0 <= i < n]
    for i in big:                                   # a single equation per launch gets the latency geometry (whole GPU)
        launch([i])
    rest = [int(i) for i in np.argsort(-costs, kind="stable") if int(i) not in set(big)]
    batch = batch or max(1, len(rest))
    for b0 in range(0, len(rest), batch):
        launch(rest[b0:b0 + batch])
    return np.array(values), evals_pt, ms_tot
