"""Parity of the CUDA path (through the C ABI of libmct_sm100a.so) against the CPU oracle and the reference goldens.

Tolerances (BASELINE.json north_star): max|dF(k,t)| <= 1e-10 away from the critical point for the same solver
settings; relative <= 1e-6 for tau_alpha and f(k) near eta_c.  Kernel evaluations differ from the reference only by
summation order (the reference's own sequential recursion carries ~2e-12 relative error), so
single evaluations are compared at rtol 1e-11.
"""
import os

import numpy as np
import pytest

import oracle as O
import problems as P
from conftest import load_pkg

pytestmark = pytest.mark.gpu
pkg = load_pkg()

KERNEL_RTOL = 1e-11


def gpu_kernel(p, **kw):
    return pkg.ModeCouplingKernel(p["rho"], p["kBT"], p["m"], p["k"], p["Sk"], **kw)


def gpu_solve(p, e, kernel=None, **solver_kw):
    kern = kernel or gpu_kernel(p)
    eq = pkg.MemoryEquation(e["alpha"], e["beta"], e["gamma"], e["delta"], e["F0"], e["dF0"], kern)
    solver = pkg.TimeDoublingSolver(**solver_kw)
    return pkg.solve(eq, solver), solver


def oracle_solve(p, e, kernel=None, **kw):
    kern = kernel or O.sc3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"])
    return O.td_solve(kern, e["alpha"], e["beta"], e["gamma"], e["delta"], e["F0"], e["dF0"], **kw)


# ------------------------------------------------------------------------------------------ kernel evaluation
def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.abs(b)))


@pytest.mark.parametrize("Nk", [2, 3, 20, 100, 257, 500, 1000])
@pytest.mark.parametrize("smooth", [True, False])
def test_eval_matches_oracle_within_reference_rounding(Nk, smooth):
    """The CUDA sums and the reference-order recursion are two float64 summation orders of the same numbers.  An
    extended-precision sum arbitrates: the CUDA result must be accurate to 1e-12, and may differ from the
    reference order by no more than the reference order's own rounding error (which grows to ~1e-9 at Nk=1000)."""
    p = P.hard_spheres(Nk, 0.51)
    F = p["Sk"] if smooth else np.random.default_rng(0).random(Nk)
    okern = O.sc3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"])
    ref = okern.evaluate(F)
    exact = P.exact_sc3d_eval(okern.tables(), p["k"], F, F)
    got = pkg.evaluate_kernel(gpu_kernel(p), F, 0.0)
    err_gpu, err_ref = _rel(got, exact), _rel(ref, exact)
    print(f"Nk={Nk} smooth={smooth}: |gpu-exact|={err_gpu:.2e} |oracle-exact|={err_ref:.2e} |gpu-oracle|={_rel(got, ref):.2e}")
    assert err_gpu < 1e-12
    assert _rel(got, ref) <= 2.0 * err_ref + 1e-12


def test_eval_from_host_tables_equals_device_built_tables():
    p = P.hard_spheres(200, 0.50)
    V = pkg.vertex_tables(p["rho"], 1.0, 1.0, p["k"], p["Sk"])
    F = p["Sk"]
    a = pkg.evaluate_kernel(gpu_kernel(p, tables=V), F)
    b = pkg.evaluate_kernel(gpu_kernel(p), F)
    assert np.array_equal(a, b)     # same tables bit for bit => same sums


def test_eval_many_and_linearity():
    """K is a quadratic form in F: K(sF) = s^2 K(F) exactly for s a power of two (size-independent property)."""
    p = P.hard_spheres(1000, 0.515)
    kern = gpu_kernel(p)
    F = np.random.default_rng(1).random((3, 1000))
    F[1] = 2.0 * F[0]
    out, _ = pkg.evaluate_kernel_many(kern, F)
    assert np.array_equal(out[1], 4.0 * out[0])
    okern = O.sc3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"])
    exact = P.exact_sc3d_eval(okern.tables(), p["k"], F[2], F[2])
    assert _rel(out[2], exact) < 1e-12


def test_eval_nonsymmetric_tables_tagged():
    p = P.hard_spheres(150, 0.50)
    rng = np.random.default_rng(2)
    Fc = rng.random((4, 150)); Fs = rng.random(150)
    ref = O.tagged3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"], Fc).evaluate(Fs, 2)

    class Sol: F = Fc
    kern = pkg.TaggedModeCouplingKernel(p["rho"], 1.0, 1.0, p["k"], p["Sk"], Sol)
    got = pkg.evaluate_kernel(kern, Fs, 2)
    assert np.max(np.abs(got - ref) / np.abs(ref)) < KERNEL_RTOL
    with pytest.raises(AssertionError):
        pkg.evaluate_kernel(kern, Fs, 4)        # time not in the collective solution: tDict KeyError in the reference


def test_large_grid_Nk2000_eval_and_short_solve():
    """Largest supported grid of the single-component path (8 wavenumbers per thread in the update phase)."""
    p = P.hard_spheres(2000, 0.51)
    kern = gpu_kernel(p)
    okern = O.sc3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"])
    got = pkg.evaluate_kernel(kern, p["Sk"], 0.0)
    ref = okern.evaluate(p["Sk"])
    assert _rel(got, ref) < 1e-10            # the reference-order recursion itself carries ~5e-12 at this size
    e = P.overdamped(p)
    kw = dict(N=4, dt=1e-6, t_max=1e-3, tolerance=1e-10)
    sol, solver = gpu_solve(p, e, kernel=kern, **kw)
    refs = oracle_solve(p, e, kernel=okern, **kw)
    assert np.max(np.abs(sol.F - refs["F"])) <= 1e-10
    assert solver.kernel_evals == refs["kernel_evals"]


def test_bad_grid_is_rejected():
    p = P.hard_spheres(32, 0.5)
    k = p["k"].copy(); k[5] *= 1.001
    with pytest.raises(AssertionError):
        pkg.ModeCouplingKernel(p["rho"], 1.0, 1.0, k, p["Sk"])       # @assert all(diff(k_array) .≈ Δk)
    with pytest.raises(AssertionError):
        pkg.ModeCouplingKernel(p["rho"], 1.0, 1.0, p["k"] + 0.1, p["Sk"])   # @assert k_array[1] ≈ Δk/2


# ------------------------------------------------------------------------------------------ time-doubling solve
def test_golden_test_MCT_56_and_oracle():
    p = P.hard_spheres(100, 0.52, rho_roundtrip=True)
    e = P.overdamped(p)
    kw = dict(N=2, dt=1e-10, t_max=1e10, tolerance=1e-8, max_iterations=10 ** 8)
    sol, solver = gpu_solve(p, e, **kw)
    assert sol.F.sum() == pytest.approx(13740.912031583604, rel=1.5e-8)          # test/test_MCT.jl:56
    ref = oracle_solve(p, e, **kw)
    assert np.array_equal(sol.t, ref["t"])
    assert sol.F.shape == ref["F"].shape
    # tolerance 1e-8 => iterates may stop one iteration apart; the bound is a few times the solver tolerance
    assert np.max(np.abs(sol.F - ref["F"])) < 5e-8
    assert abs(solver.kernel_evals - ref["kernel_evals"]) <= 0.01 * ref["kernel_evals"]


@pytest.mark.parametrize("team", [0, 1, 3, 16, 148])
def test_config2_Nk100_eta051_defaults_vs_oracle(team, monkeypatch):
    """BASELINE config 2 with the default solver (N=32, tol 1e-10) for several team sizes (CTAs per equation)."""
    if team:
        monkeypatch.setenv("MCT_TEAM_SIZE", str(team))
    p = P.hard_spheres(100, 0.51)
    e = P.overdamped(p)
    sol, solver = gpu_solve(p, e)
    ref = oracle_solve(p, e)
    assert np.array_equal(sol.t, ref["t"])
    dF = np.max(np.abs(sol.F - ref["F"]))
    dK = np.max(np.abs(sol.K - ref["K"]) / (np.abs(ref["K"]) + 1e-300))
    print(f"team={team} evals gpu={solver.kernel_evals} oracle={ref['kernel_evals']} max|dF|={dF:.2e} max rel dK={dK:.2e}")
    assert dF <= 1e-10
    assert abs(solver.kernel_evals - ref["kernel_evals"]) <= 0.01 * ref["kernel_evals"]


def test_tight_tolerance_separates_solver_noise_from_kernel_noise():
    p = P.hard_spheres(100, 0.51)
    e = P.overdamped(p)
    kw = dict(N=16, dt=1e-8, t_max=1e4, tolerance=1e-13, max_iterations=10 ** 6)
    sol, _ = gpu_solve(p, e, **kw)
    ref = oracle_solve(p, e, **kw)
    assert np.max(np.abs(sol.F - ref["F"])) <= 2e-12


def test_newtonian_golden_test_tagged_59_and_tagged_chain():
    p = P.hard_spheres(100, 0.514)
    kw = dict(N=8, dt=1e-5, t_max=1e5, tolerance=1e-8)
    e = P.newtonian(p)
    sol, _ = gpu_solve(p, e, **kw)
    assert sol.F.sum() == pytest.approx(25492.648222645254, rel=1.5e-8)          # test/test_tagged.jl:59
    ref = oracle_solve(p, e, **kw)
    assert np.max(np.abs(sol.F - ref["F"])) < 5e-8
    # tagged particle chained on the device-resident collective trajectory (test/test_tagged.jl:44-55)
    et = P.tagged_newtonian(p)
    tk = pkg.TaggedModeCouplingKernel(p["rho"], 1.0, 1.0, p["k"], p["Sk"], sol)
    sols, _ = gpu_solve(p, et, kernel=tk, **kw)
    otk = O.tagged3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"], ref["F"])
    refs = oracle_solve(p, et, kernel=otk, **kw)
    assert np.max(np.abs(sols.F - refs["F"])) < 5e-8
    # same through host tables + host trajectory (the path the Julia glue uses)
    V = pkg.vertex_tables(p["rho"], 1.0, 1.0, p["k"], p["Sk"], tagged=True)
    sol.F_host_only = True
    class HostSol: F = sol.F
    tk2 = pkg.TaggedModeCouplingKernel(p["rho"], 1.0, 1.0, p["k"], p["Sk"], HostSol, tables=V)
    sols2, _ = gpu_solve(p, et, kernel=tk2, **kw)
    assert np.max(np.abs(sols2.F - sols.F)) < 1e-12
    # MSD from the two device solutions, integrated by the oracle's scalar solver (test/test_tagged.jl:62-71)
    Kmsd = O.msd3d_table(p["rho"], 1.0, 1.0, p["k"], p["Sk"], sol.F, sols.F)
    msd = O.td_solve(O.tabulated(Kmsd), 1.0, 0.0, 0.0, -6.0, [0.0], [0.0], scalar_mode=True, **kw)
    assert msd["F"].sum() == pytest.approx(51.75329405084479, rel=1.5e-8)
    # the same on the device: kernel table reduced from the two resident trajectories, the scalar equation
    # integrated by one device thread in the reference's immutable-branch order
    mk = pkg.MSDModeCouplingKernel(p["rho"], 1.0, 1.0, p["k"], p["Sk"], sol, sols)
    msolver = pkg.TimeDoublingSolver(**kw)
    msol = pkg.solve(pkg.MemoryEquation(1.0, 0.0, 0.0, -6.0, 0.0, 0.0, mk), msolver)
    assert msol.F.sum() == pytest.approx(51.75329405084479, rel=1.5e-8)            # test/test_tagged.jl:71
    assert np.allclose(msol.K, Kmsd, rtol=1e-12, atol=0)
    assert np.allclose(msol.t, msd["t"], rtol=0, atol=0)
    assert np.max(np.abs(msol.F - np.ravel(msd["F"]))) <= 1e-10 * max(1.0, np.max(np.abs(msd["F"])))
    assert msolver.kernel_evals == msd["kernel_evals"]


def test_relaxation_time_near_critical_point_docs_Scope_230():
    """tau_alpha at eta = 0.51591 (eps ~ 1e-4 below eta_c), the set-up of the example in docs/src/Scope.md:212-231.  The
    acceptance bar is a relative 1e-6 against the reference's own solve, for which the oracle stands in (same operation
    order, pinned on every CI-tested golden incl. the Newtonian N=8 solve of test/test_tagged.jl:59).  The value printed in
    the docs (4.232796654132995e11, not covered by the reference's tests) is NOT what the algorithm of v0.8.9 gives for
    these inputs: oracle and device agree on 2.9483873e11 with identical iteration counts; tau_alpha ~ |eps|^-2.46 this
    close to eta_c turns a 1e-5 change of eta into such a factor, so the doc figure is reported, not asserted."""
    p = P.hard_spheres(100, 0.51591)
    e = P.newtonian(p)
    kw = dict(N=8, dt=1e-5, t_max=1e15, tolerance=1e-8, max_iterations=10 ** 6)
    sol, solver = gpu_solve(p, e, **kw)
    ref = oracle_solve(p, e, **kw)
    tau_gpu = pkg.find_relaxation_time(sol.t, sol.F[:, 17])
    tau_ref = O.relaxation_time(ref["t"], ref["F"][:, 17])
    print(f"tau_alpha(k18): gpu={tau_gpu:.9e} oracle={tau_ref:.9e} (docs/src/Scope.md:231 prints 4.232796654132995e11)  "
          f"evals gpu={solver.kernel_evals} oracle={ref['kernel_evals']}")
    assert tau_gpu == pytest.approx(tau_ref, rel=1e-6)
    assert solver.kernel_evals == ref["kernel_evals"]
    for ik in range(0, 100, 7):
        a, b = pkg.find_relaxation_time(sol.t, sol.F[:, ik]), O.relaxation_time(ref["t"], ref["F"][:, ik])
        if np.isfinite(b):
            assert a == pytest.approx(b, rel=1e-6)


def test_not_converged_raises_domain_error():
    p = P.hard_spheres(64, 0.52)
    e = P.overdamped(p)
    with pytest.raises(pkg.DomainError):
        gpu_solve(p, e, N=4, dt=1e-6, t_max=1e6, tolerance=1e-14, max_iterations=2)


@pytest.mark.parametrize("team", [0, 3])
def test_batch_equals_individual_solves(team, monkeypatch):
    """team=3: same team geometry for both => bitwise identical.  team=0: the library picks a latency-oriented team for
    the single solves and a throughput-oriented one for the batch => different summation order, same tolerance."""
    if team:
        monkeypatch.setenv("MCT_TEAM_SIZE", str(team))
    etas = [0.40, 0.45, 0.50, 0.51, 0.515]
    kw = dict(N=8, dt=1e-8, t_max=1e6, tolerance=1e-10)
    eqs, singles = [], []
    for eta in etas:
        p = P.hard_spheres(96, eta)
        e = P.overdamped(p)
        kern = gpu_kernel(p)
        eqs.append(pkg.MemoryEquation(e["alpha"], e["beta"], e["gamma"], e["delta"], e["F0"], e["dF0"], kern))
        singles.append(gpu_solve(p, e, kernel=kern, **kw)[0])
    sols, status, evals, ms = pkg.solve_batch(eqs, pkg.TimeDoublingSolver(**kw))
    assert status == [0] * len(etas)
    for a, b in zip(sols, singles):
        if team:
            assert np.array_equal(a.F, b.F) and np.array_equal(a.K, b.K)     # same geometry => bitwise identical
        else:
            assert np.max(np.abs(a.F - b.F)) <= 1e-10


# ------------------------------------------------------------------------------------------ steady state
def test_steady_state_golden_test_steady_state_64():
    p = P.hard_spheres(100, 0.51595, rho_roundtrip=True)
    sol = pkg.solve_steady_state(p["k"] ** 2 / p["Sk"], p["Sk"], gpu_kernel(p), tolerance=1e-8)
    assert sol.F.sum() == pytest.approx(19.285439116785284, rel=1.5e-8)          # test/test_steady_state.jl:64
    ref = O.steady_state(O.sc3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"]), p["k"] ** 2 / p["Sk"], p["Sk"], tolerance=1e-8)
    assert np.max(np.abs(sol.F[0] - ref["F"]) / np.maximum(np.abs(ref["F"]), 1e-3)) < 1e-6
    assert abs(sol.iterations - ref["iterations"]) <= 2


def test_steady_state_below_critical_point_is_zero():
    p = P.hard_spheres(100, 0.50)
    sol = pkg.solve_steady_state(p["k"] ** 2 / p["Sk"], p["Sk"], gpu_kernel(p), tolerance=1e-10)
    assert sol.F.sum() < 1e-7


def test_slice_download_and_relaxation_times_match_the_full_trajectory():
    """get_F(sol, its, iks) and find_relaxation_time on the device-resident solution must give exactly what
    the whole downloaded trajectory gives (src/HelperFunctions.jl:169-210, src/RelaxationTime.jl:14-39)."""
    p = P.hard_spheres(100, 0.51)
    sol, solver = gpu_solve(p, P.overdamped(p), N=8, dt=1e-8, t_max=1e6, tolerance=1e-10)
    nt = sol.F.shape[0]
    times = np.array([0, 1, 17, nt // 2, nt - 1])
    dofs = np.array([0, 18, 19, 99])
    t, F, K = pkg.api.download_slice(sol._equation, 100, times, dofs)
    assert np.array_equal(t, sol.t[times])
    assert np.array_equal(F, sol.F[np.ix_(times, dofs)])
    assert np.array_equal(K, sol.K[np.ix_(times, dofs)])
    t, F, K = pkg.api.download_slice(sol._equation, 100, [nt - 1], None, want_K=False)        # f(k) at the last time
    assert K is None and np.array_equal(F[0], sol.F[-1])
    for mode in ("log", "lin"):
        tau = pkg.api.relaxation_times(sol._equation, dofs, mode=mode)
        host = [pkg.find_relaxation_time(sol.t, sol.F[:, ik], mode=mode) for ik in dofs]
        assert np.allclose(tau, host, rtol=1e-14, atol=0)
    assert np.isinf(pkg.api.relaxation_times(sol._equation, [18], threshold=-1.0)[0])           # never crossed
    with pytest.raises(Exception):
        pkg.api.download_slice(sol._equation, 100, [nt], [0])
    # batch solve that returns only the slice (what the packing-fraction sweep needs)
    eq = pkg.MemoryEquation(0.0, 1.0, p["k"] ** 2 / p["Sk"], 0.0, p["Sk"], np.zeros(100), gpu_kernel(p))
    sols, status, _, _ = pkg.solve_batch([eq], pkg.TimeDoublingSolver(N=8, dt=1e-8, t_max=1e6, tolerance=1e-10),
                                         download=dict(times=[0, nt - 1], dofs=[18], K=False))
    assert status == [0] and sols[0].F.shape == (2, 1)
    assert np.allclose(sols[0].F[:, 0], sol.F[[0, nt - 1], 18], rtol=0, atol=1e-10)


@pytest.mark.parametrize("Nk,eta", [(100, 0.51), (1000, 0.50)])
def test_streamed_download_of_mct_td_solve_equals_the_resident_solution(Nk, eta):
    """mct_td_solve copies the trajectory to the host buffers WHILE the persistent kernel produces it (sentinel-validated
    time blocks); the result must be bit-identical to solving resident and downloading afterwards."""
    p = P.hard_spheres(Nk, eta)
    e = P.overdamped(p)
    kw = dict(N=16, dt=1e-8, t_max=1e4, tolerance=1e-10)
    kern = gpu_kernel(p)
    eq = pkg.MemoryEquation(e["alpha"], e["beta"], e["gamma"], e["delta"], e["F0"], e["dF0"], kern)
    ref = pkg.solve(eq, pkg.TimeDoublingSolver(**kw))
    nt = ref.F.shape[0]
    for _ in range(3):
        t = np.full(nt, np.nan); F = np.full((nt, Nk), np.nan); K = np.full((nt, Nk), np.nan)
        s2 = pkg.TimeDoublingSolver(**kw)
        pkg.api.solve_into(eq, s2, t, F, K)
        assert np.array_equal(t, ref.t) and np.array_equal(F, ref.F) and np.array_equal(K, ref.K)
        assert s2.kernel_evals == ref.solver.kernel_evals
    # a solve that does not converge must return its error, not wait for blocks that never come
    t = np.empty(nt); F = np.empty((nt, Nk)); K = np.empty((nt, Nk))
    with pytest.raises(pkg.DomainError):
        pkg.api.solve_into(eq, pkg.TimeDoublingSolver(max_iterations=1, **kw), t, F, K)


def test_tagged_solve_on_another_time_grid_is_a_keyerror_and_handles_are_refcounted():
    """The reference looks the collective solution up with tDict[t] (src/Kernels/ModeCouplingKernels.jl:394, 436): a tagged
    solve with another N or dt hits a KeyError.  The device pairs Fs(t) with Fc by output index, so mct_equation_create
    must check the two time grids.  Also: handle lifetimes (a kernel cannot go while its equations live, a collective
    equation cannot go while a tagged kernel reads its trajectory)."""
    p = P.hard_spheres(64, 0.50)
    kw = dict(N=8, dt=1e-6, t_max=1e-2, tolerance=1e-10)
    sol, _ = gpu_solve(p, P.overdamped(p), **kw)
    et = dict(alpha=0.0, beta=1.0, gamma=p["k"] ** 2, delta=0.0, F0=np.ones(64), dF0=np.zeros(64))

    class HostSol: F = sol.F; t = sol.t
    for tk in (pkg.TaggedModeCouplingKernel(p["rho"], 1.0, 1.0, p["k"], p["Sk"], sol),          # chained on the device solution
               pkg.TaggedModeCouplingKernel(p["rho"], 1.0, 1.0, p["k"], p["Sk"], HostSol)):     # host trajectory + get_t(sol)
        good, _ = gpu_solve(p, et, kernel=tk, **kw)
        assert np.all(np.isfinite(good.F))
        shorter, _ = gpu_solve(p, et, kernel=tk, **dict(kw, t_max=1e-4))     # a prefix of the collective grid is fine
        assert np.array_equal(shorter.F, good.F[:shorter.F.shape[0]])
        for bad in (dict(kw, dt=2e-6), dict(kw, N=4), dict(kw, t_max=1.0)):
            with pytest.raises(AssertionError, match="tDict"):
                gpu_solve(p, et, kernel=tk, **bad)
        del good, shorter
    # lifetimes
    tk = pkg.TaggedModeCouplingKernel(p["rho"], 1.0, 1.0, p["k"], p["Sk"], sol)
    with pytest.raises(pkg.MCTDeviceError, match="tagged kernel"):
        sol._equation.close()
    ts, _ = gpu_solve(p, et, kernel=tk, **kw)
    with pytest.raises(pkg.MCTDeviceError, match="still alive"):
        tk.close()
    ts._equation.close()
    tk.close()
    sol._equation.close()
    # a multi-component solution cannot feed the single-component tagged constructor
    b = P.binary_mixture(Nk=64)
    mk = pkg.MultiComponentModeCouplingKernel(b["rho"], 1.0, b["m"], b["k"], b["Sk"])
    z = np.zeros((64, 2, 2))
    msol = pkg.solve(pkg.MemoryEquation(1.0, 0.0, b["gamma"], z, b["Sk"], z, mk), pkg.TimeDoublingSolver(N=4, dt=1e-4, t_max=1e-3))
    with pytest.raises(AssertionError, match="single-component"):
        pkg.TaggedModeCouplingKernel(p["rho"], 1.0, 1.0, p["k"], p["Sk"], msol)
