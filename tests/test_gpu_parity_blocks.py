"""GPU parity of the block-valued and dense kernels (MultiComponentModeCouplingKernel3D, dDimModeCouplingKernel) and the
solvers that consume them, through the C ABI, against the CPU oracle and the reference's goldens.

Tolerances: the device sums the same numbers in a different order (line sums + prefix instead of the reference's
sequential recursion; factorised species contraction, which the reference itself evaluates under @fastmath), so single
evaluations are compared at rtol 1e-11 and solves at the north-star bound max|dF| <= 1e-10.
"""
import os

import numpy as np
import pytest

import oracle as O
import problems as P
from conftest import ROOT, load_pkg

pytestmark = pytest.mark.gpu
pkg = load_pkg()
RTOL = 1.5e-8          # Julia's isapprox default, used by the reference's golden comparisons


def mixture(ns, Nk=40, phi=0.45):
    diam = np.linspace(0.7, 1.0, ns) if ns > 1 else np.array([1.0])
    cr = np.arange(1, ns + 1, dtype=float); cr /= cr.sum()
    rho = 6 * phi / (np.pi * np.sum(cr * diam ** 3)) * cr
    k = P.sf.wavenumber_grid(Nk)
    Sk = P.sf.percus_yevick_mixture_sk(k, diam, rho)
    return rho, np.linspace(1.0, 1.5, ns), k, Sk


@pytest.mark.parametrize("ns", [1, 2, 3, 4])
def test_mc_eval_matches_oracle_and_brute_force(ns):
    rho, m, k, Sk = mixture(ns)
    Nk = len(k)
    rng = np.random.default_rng(ns)
    F = rng.random((Nk, ns, ns))
    kern = pkg.MultiComponentModeCouplingKernel(rho, 1.3, m, k, Sk)
    got = pkg.evaluate_kernel(kern, F)
    ref = O.mc3d(rho, 1.3, m, k, Sk).evaluate(O._blocks(F))
    scale = np.abs(ref).max()
    assert np.max(np.abs(got - ref)) <= 1e-11 * scale
    if ns <= 2:      # the reference's own cross-check: fast kernel == brute-force triangle sum (test/test_MCMCT.jl:192)
        naive = O.mc3d(rho, 1.3, m, k, Sk, naive=True).evaluate(O._blocks(F))
        assert np.allclose(got, naive, rtol=RTOL, atol=1e-12 * scale)


def test_mc_time_doubling_golden_test_MCMCT_211():
    b = P.binary_mixture()
    Nk = b["Nk"]
    kw = dict(N=4, dt=1e-4, t_max=1e3, tolerance=1e-12, max_iterations=20000)
    kern = pkg.MultiComponentModeCouplingKernel(b["rho"], 1.0, b["m"], b["k"], b["Sk"])
    eq = pkg.MemoryEquation(1.0, 0.0, b["gamma"], np.zeros((Nk, 2, 2)), b["Sk"], np.zeros((Nk, 2, 2)), kern)
    solver = pkg.TimeDoublingSolver(**kw)
    sol = pkg.solve(eq, solver)
    Fend = sol.F[-1].reshape(Nk, 4)[18].reshape(2, 2, order="F")
    gold = np.array([[0.13881069002073976, 0.04192527228732725], [0.041925272287327266, 0.52419581715369]])
    assert np.allclose(Fend, gold, rtol=RTOL, atol=0)                                   # test/test_MCMCT.jl:211
    ref = O.td_solve(O.mc3d(b["rho"], 1.0, b["m"], b["k"], b["Sk"]), 1.0, 0.0, O._blocks(b["gamma"]), np.zeros(Nk * 4),
                     O._blocks(b["Sk"]), np.zeros(Nk * 4), **kw)
    assert np.array_equal(sol.t, ref["t"])
    assert np.max(np.abs(sol.F - ref["F"])) <= 1e-10
    assert abs(solver.kernel_evals - ref["kernel_evals"]) <= 0.01 * ref["kernel_evals"]


def test_mc_steady_state_golden_test_steady_state_104():
    Nk = 100
    diam = np.array([0.8, 1.0]); cr = np.array([0.2, 0.8]); cr /= cr.sum()
    rho = 6 * 0.55 / (np.pi * np.sum(cr * diam ** 3)) * cr
    x = rho / rho.sum()
    k = P.sf.wavenumber_grid(Nk)
    data = np.loadtxt(os.path.join(ROOT, "tests", "golden", "Sk_MC2.txt")).reshape(-1)
    Sk = np.transpose(data.reshape((2, 2, 100), order="F") * np.sqrt(x[:, None, None] * x[None, :, None]), (2, 0, 1)).copy()
    gam = np.einsum("kab,kbc->kac", np.array([np.diag(k[i] ** 2 * x) for i in range(Nk)]), np.linalg.inv(Sk))
    kern = pkg.MultiComponentModeCouplingKernel(rho, 1.0, np.ones(2), k, Sk)
    sol = pkg.solve_steady_state(gam, Sk, kern, tolerance=1e-12)
    assert sol.F.sum() == pytest.approx(45.50146332163797, rel=RTOL)                     # test/test_steady_state.jl:104
    ref = O.steady_state(O.mc3d(rho, 1.0, np.ones(2), k, Sk), O._blocks(gam), O._blocks(Sk), tolerance=1e-12)
    assert np.max(np.abs(sol.F[0] - ref["F"]) / np.maximum(np.abs(ref["F"]), 1e-3)) < 1e-6


def test_ddim_eval_matches_oracle_and_bengtzelius_at_d3():
    p = P.hard_spheres(20, 0.50)
    kd = pkg.dDimModeCouplingKernel(p["rho"], 1.0, 1.0, p["k"], p["Sk"], 3)
    got = pkg.evaluate_kernel(kd, p["Sk"])
    ref = O.ddim(p["rho"], 1.0, 1.0, p["k"], p["Sk"], 3).evaluate(p["Sk"])
    assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-12
    ks = pkg.evaluate_kernel(pkg.ModeCouplingKernel(p["rho"], 1.0, 1.0, p["k"], p["Sk"]), p["Sk"])
    assert np.abs(got - ks).sum() < 1e-8                                                 # test/test_dDimMCT_consistency.jl:69
    F = np.random.default_rng(5).random(20)
    assert np.max(np.abs(pkg.evaluate_kernel(kd, F) - O.ddim(p["rho"], 1.0, 1.0, p["k"], p["Sk"], 3).evaluate(F)) / np.abs(ref)) < 1e-12


def test_ddim_host_tables_are_bit_identical_to_oracle():
    p = P.hard_spheres(16, 0.45)
    for d in (2, 3, 5):
        pref, V, J = pkg.ddim_tables(p["rho"], 1.0, 1.0, p["k"], p["Sk"], d)
        Vo, Jo, po = O.ddim_tables(O.ddim(p["rho"], 1.0, 1.0, p["k"], p["Sk"], d))
        assert np.allclose(V, Vo, rtol=1e-15, atol=0) and np.allclose(J, Jo, rtol=1e-14, atol=0) and pref == pytest.approx(po, rel=1e-15)


def test_ddim_critical_point_2d_goldens_test_dDimMCT_crit_pt_96_97():
    from scipy.special import j0, j1
    k = P.sf.wavenumber_grid(100)

    def ck2d(phi):
        t1 = -5 * (1 - phi) ** 2 * k ** 2 * j0(k / 2) ** 2 / 4
        t2 = (4 * ((phi - 20) * phi + 7) + 5 * (1 - phi) ** 2 * k ** 2 / 4) * j1(k / 2) ** 2
        t3 = 2 * (phi - 13) * (1 - phi) * k * j1(k / 2) * j0(k / 2)
        return np.pi * (t1 + t2 + t3) / (6 * (1 - phi) ** 3 * k ** 2)

    out = {}
    for phi in (0.69, 0.70):
        rho = 4 * phi / np.pi
        ck = ck2d(phi)
        Sk = 1 + rho * ck / (1 - rho * ck)
        kern = pkg.dDimModeCouplingKernel(rho, 1.0, 1.0, k, Sk, 2)
        out[phi] = pkg.solve_steady_state(k ** 2 / Sk, Sk, kern, tolerance=1e-8).F.sum()
    assert out[0.70] == pytest.approx(17.987734457716535, rel=RTOL)                      # test/test_dDimMCT_crit_pt.jl:96
    assert out[0.69] < 1e-7                                                              # :97


def test_ddim_critical_point_5d_goldens_test_dDimMCT_crit_pt_133_134():
    """d = 5 with Leutheusser's Percus-Yevick c(k) (test/test_dDimMCT_crit_pt.jl:25-47, :100-134) through the device's
    steady-state solver."""
    k = P.sf.wavenumber_grid(100)

    def ck5d(eta):
        T = 1 + 18 * eta + 6 * eta ** 2
        Q0 = 1 / (120 * eta * (1 - eta) ** 3) * (1 - 33 * eta - 87 * eta ** 2 - 6 * eta ** 3 - T ** 1.5)
        Q1 = -1 / (12 * (1 - eta) ** 3) * ((3 + 2 * eta) * T ** 0.5 + 3 + 19 * eta + 3 * eta ** 2)
        Q2 = -T ** 0.5 / (24 * (1 - eta) ** 3) * (2 + 3 * eta + T ** 0.5)
        c0 = -(-8 * Q2) ** 2
        c1 = 120 * eta * Q0 ** 2
        c3 = 20 * eta * (8 * Q0 * Q2 - 3 * Q1 ** 2)
        c5 = -3 / 8 * eta * c0
        return -(1 / k ** 10) * 8 * np.pi ** 2 * (
            8 * (720 * c5 - 18 * c3 * k ** 2 + c1 * k ** 4)
            + (-5760 * c5 + 144 * (c3 + 20 * c5) * k ** 2 - 8 * (c1 + 9 * c3 + 30 * c5) * k ** 4 + (3 * c0 + 4 * c1 + 6 * c3 + 8 * c5) * k ** 6) * np.cos(k)
            + k * (-5760 * c5 + 48 * (3 * c3 + 20 * c5) * k ** 2 - (3 * c0 + 8 * (c1 + 3 * c3 + 6 * c5)) * k ** 4 + (c0 + c1 + c3 + c5) * k ** 6) * np.sin(k))

    out = {}
    for phi in (0.25, 0.26):
        rho = 60 * phi / np.pi ** 2
        ck = ck5d(phi)
        Sk = 1 + rho * ck / (1 - rho * ck)
        kern = pkg.dDimModeCouplingKernel(rho, 1.0, 1.0, k, Sk, 5)
        out[phi] = pkg.solve_steady_state(k ** 2 / Sk, Sk, kern, tolerance=1e-8).F.sum()
    assert out[0.26] == pytest.approx(33.463940782589255, rel=RTOL)                      # test/test_dDimMCT_crit_pt.jl:133
    assert out[0.25] < 1e-7                                                              # :134


def test_ddim_time_doubling_vs_oracle():
    p = P.hard_spheres(24, 0.50)
    e = P.overdamped(p)
    kw = dict(N=8, dt=1e-6, t_max=1e2, tolerance=1e-10)
    kern = pkg.dDimModeCouplingKernel(p["rho"], 1.0, 1.0, p["k"], p["Sk"], 3)
    eq = pkg.MemoryEquation(e["alpha"], e["beta"], e["gamma"], e["delta"], e["F0"], e["dF0"], kern)
    solver = pkg.TimeDoublingSolver(**kw)
    sol = pkg.solve(eq, solver)
    ref = O.td_solve(O.ddim(p["rho"], 1.0, 1.0, p["k"], p["Sk"], 3), e["alpha"], e["beta"], e["gamma"], e["delta"], e["F0"], e["dF0"], **kw)
    assert np.max(np.abs(sol.F - ref["F"])) <= 1e-10
    assert abs(solver.kernel_evals - ref["kernel_evals"]) <= max(3, 0.01 * ref["kernel_evals"])


def test_mc_tagged_and_msd_chain_goldens_test_MCMCT_212_226():
    """Mixture -> tagged particle of species 2 -> its MSD, every step chained on the device-resident solution of the step
    before (test/test_MCMCT.jl:196-226)."""
    b = P.binary_mixture()
    Nk = b["Nk"]
    kw = dict(N=4, dt=1e-4, t_max=1e3, tolerance=1e-12, max_iterations=20000)
    kern = pkg.MultiComponentModeCouplingKernel(b["rho"], 1.0, b["m"], b["k"], b["Sk"])
    eq = pkg.MemoryEquation(1.0, 0.0, b["gamma"], np.zeros((Nk, 2, 2)), b["Sk"], np.zeros((Nk, 2, 2)), kern)
    sol = pkg.solve(eq, pkg.TimeDoublingSolver(**kw))
    s = 2                                                                                # the reference's 1-based species
    tk = pkg.TaggedMultiComponentModeCouplingKernel(s, b["rho"], 1.0, b["m"], b["k"], b["Sk"], sol)
    teq = pkg.MemoryEquation(1.0, 0.0, b["k"] ** 2 / b["m"][s - 1], 0.0, np.ones(Nk), np.zeros(Nk), tk)
    tsol = pkg.solve(teq, pkg.TimeDoublingSolver(**kw))
    assert tsol.F[-1][18] == pytest.approx(0.7678799482793754, rel=RTOL)                # test/test_MCMCT.jl:212
    traj = sol.F.reshape(-1, Nk, 4)
    otk = O.tagged_mc3d(s - 1, b["rho"], 1.0, b["m"], b["k"], b["Sk"], traj)
    ots = O.td_solve(otk, 1.0, 0.0, b["k"] ** 2 / b["m"][s - 1], 0.0, np.ones(Nk), np.zeros(Nk), **kw)
    # Newtonian dynamics amplify rounding differences of the kernel (the oracle sums per species pair, the device
    # contracts the species weight first): same bar as the single-component Newtonian chain (test_gpu_parity.py)
    assert np.max(np.abs(tsol.F - ots["F"])) < 5e-8
    mk = pkg.MSDMultiComponentModeCouplingKernel(s, b["rho"], 1.0, b["m"], b["k"], b["Sk"], sol, tsol)
    msol = pkg.solve(pkg.MemoryEquation(1.0, 0.0, 0.0, -6.0 / b["m"][s - 1], 0.0, 0.0, mk), pkg.TimeDoublingSolver(**kw))
    assert msol.F.sum() == pytest.approx(0.6057207573375025, rel=RTOL)                  # test/test_MCMCT.jl:226
    Kt = O.msd_mc3d_table(s - 1, b["rho"], 1.0, b["m"], b["k"], b["Sk"], traj, tsol.F)
    assert np.allclose(msol.K, Kt, rtol=1e-11, atol=0)
    oms = O.td_solve(O.tabulated(Kt), 1.0, 0.0, 0.0, -6.0 / b["m"][s - 1], [0.0], [0.0], scalar_mode=True, **kw)
    assert np.max(np.abs(msol.F - np.ravel(oms["F"]))) <= 1e-9 * max(1.0, np.max(np.abs(oms["F"])))


def test_active_mct_kernel_eval_solve_and_steady_state_test_ActiveMCT():
    """ActiveMCTKernel is the dDim triple sum over another table (output index last).  Kernel evaluation
    against the oracle's loop restatement; the reference's own checks (test/test_ActiveMCT.jl:16-32): with w(k)=w0=1 the
    solve and the steady state equal the passive 3D kernel's (rtol 1e-5), and w0 = 0.8 changes the result."""
    Nk = 20
    p = P.hard_spheres(Nk, 0.51 * np.pi / 6, kmax=20.0)
    rho, k, Sk = 0.51, p["k"], p["Sk"]
    rng = np.random.default_rng(0)
    for dim, wk, w0 in ((3, np.ones(Nk), 1.0), (3, np.linspace(0.8, 1.2, Nk), 0.9), (2, np.linspace(0.8, 1.2, Nk), 0.9)):
        ka = pkg.ActiveMCTKernel(rho, k, wk, w0, Sk, dim)
        pref, J, V2 = O.active_tables(rho, k, wk, w0, Sk, dim)
        F = rng.random(Nk)
        assert np.allclose(pkg.evaluate_kernel(ka, F), O.active_evaluate(pref, J, V2, F), rtol=1e-12, atol=0)
    wk = np.ones(Nk)
    gamma = k ** 2 * wk / Sk
    kw = dict(N=8, dt=1e-6, t_max=1e4, tolerance=1e-10)
    passive = pkg.solve(pkg.MemoryEquation(1.0, 0.0, gamma, 0.0, Sk, np.zeros(Nk), pkg.ModeCouplingKernel(rho, 1.0, 1.0, k, Sk)),
                        pkg.TimeDoublingSolver(**kw))
    ka = pkg.ActiveMCTKernel(rho, k, wk, 1.0, Sk, 3)
    active = pkg.solve(pkg.MemoryEquation(1.0, 0.0, gamma, 0.0, Sk, np.zeros(Nk), ka), pkg.TimeDoublingSolver(**kw))
    assert active.F.sum() == pytest.approx(passive.F.sum(), rel=1e-5)                   # test/test_ActiveMCT.jl:22
    assert pkg.find_relaxation_time(active.t, active.F[:, 5]) == pytest.approx(
        pkg.find_relaxation_time(passive.t, passive.F[:, 5]), rel=1e-5)                  # :23
    ss_p = pkg.solve_steady_state(gamma, Sk, pkg.ModeCouplingKernel(rho, 1.0, 1.0, k, Sk))
    ss_a = pkg.solve_steady_state(gamma, Sk, ka)
    assert ss_a.F.sum() == pytest.approx(ss_p.F.sum(), rel=1e-5)                        # :24
    kb = pkg.ActiveMCTKernel(rho, k, wk, 0.8, Sk, 3)
    other = pkg.solve(pkg.MemoryEquation(1.0, 0.0, gamma, 0.0, Sk, np.zeros(Nk), kb), pkg.TimeDoublingSolver(**kw))
    assert abs(other.F.sum() - active.F.sum()) > 1e-3 * abs(active.F.sum())              # :32


def test_tagged_active_mct_kernel_test_ActiveMCT_57_64():
    """TaggedActiveMCTKernel chained on the device-resident collective solution: evaluation against the oracle's loops at
    several times, and the reference's check that with w(k) = w0 = 1 it equals the passive tagged kernel (rtol 1e-5)."""
    Nk = 20
    p = P.hard_spheres(Nk, 0.51 * np.pi / 6, kmax=20.0)
    rho, k, Sk = 0.51, p["k"], p["Sk"]
    wk, w0 = np.ones(Nk), 1.0
    gamma = k ** 2 * wk / Sk
    kw = dict(N=8, dt=1e-6, t_max=1e4, tolerance=1e-10)
    sol_act = pkg.solve(pkg.MemoryEquation(1.0, 0.0, gamma, 0.0, Sk, np.zeros(Nk), pkg.ActiveMCTKernel(rho, k, wk, w0, Sk, 3)),
                        pkg.TimeDoublingSolver(**kw))
    tk = pkg.TaggedActiveMCTKernel(rho, k, wk, w0, Sk, sol_act, 3)
    pref, J, V2 = O.tagged_active_tables(rho, k, wk, w0, Sk, 3)
    Fs = np.random.default_rng(1).random(Nk)
    for tix in (0, 5, sol_act.F.shape[0] - 1):
        ref = O.tagged_active_evaluate(pref, J, V2, sol_act.F[tix], Fs)
        assert np.allclose(pkg.evaluate_kernel(tk, Fs, tix), ref, rtol=1e-12, atol=1e-300)
    gamma2 = k ** 2 * w0
    tsol_act = pkg.solve(pkg.MemoryEquation(1.0, 0.0, gamma2, 0.0, np.ones(Nk), np.zeros(Nk), tk), pkg.TimeDoublingSolver(**kw))
    sol = pkg.solve(pkg.MemoryEquation(1.0, 0.0, gamma, 0.0, Sk, np.zeros(Nk), pkg.ModeCouplingKernel(rho, 1.0, 1.0, k, Sk)),
                    pkg.TimeDoublingSolver(**kw))
    tkp = pkg.TaggedModeCouplingKernel(rho, 1.0, 1.0, k, Sk, sol)
    tsol = pkg.solve(pkg.MemoryEquation(1.0, 0.0, gamma2, 0.0, np.ones(Nk), np.zeros(Nk), tkp), pkg.TimeDoublingSolver(**kw))
    assert tsol_act.F.sum() == pytest.approx(tsol.F.sum(), rel=1e-5)                     # test/test_ActiveMCT.jl:64


@pytest.mark.parametrize("Nk", [2, 3, 5, 33, 63, 130])
@pytest.mark.parametrize("ns", [2, 4])
def test_mc_eval_ragged_grid_sizes(ns, Nk):
    """The line sums run on 4-wavenumber tiles with zero padding and only the d >= 0 / iq < ip halves are walked: sizes that
    are not multiples of 4 (and the smallest grids) against the oracle's reference-order evaluation, with a random
    NON-symmetric F so that nothing but the mirror property of the entries themselves is relied on."""
    rho, m, k, Sk = mixture(ns, Nk=Nk)
    rng = np.random.default_rng(100 * ns + Nk)
    F = rng.random((Nk, ns, ns)) - 0.3
    kern = pkg.MultiComponentModeCouplingKernel(rho, 0.7, m, k, Sk)
    got = pkg.evaluate_kernel(kern, F)
    ref = O.mc3d(rho, 0.7, m, k, Sk).evaluate(O._blocks(F))
    assert np.max(np.abs(got - ref)) <= 1e-11 * np.abs(ref).max()
