"""Pins the CPU oracle (oracle/mct_oracle.c) on the reference's own regression goldens.
Tolerance: the reference's `≈` (rtol = sqrt(eps) ~ 1.5e-8).  Runs on the CPU."""
import os

import numpy as np
import pytest

import oracle as O
import problems as P

RTOL = 1.5e-8


def test_sc_time_doubling_golden_test_MCT_56():
    p = P.hard_spheres(100, 0.52, rho_roundtrip=True)
    e = P.overdamped(p)
    kern = O.sc3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"])
    sol = O.td_solve(kern, e["alpha"], e["beta"], e["gamma"], 0.0, e["F0"], e["dF0"], N=2, dt=1e-10, t_max=1e10, tolerance=1e-8,
                     max_iterations=10 ** 8)
    assert sol["status"] == 0
    assert sol["F"].sum() == pytest.approx(13740.912031583604, rel=RTOL)


@pytest.fixture(scope="module")
def tagged_chain():
    p = P.hard_spheres(100, 0.514)
    e = P.newtonian(p)
    kern = O.sc3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"])
    kw = dict(N=8, dt=1e-5, t_max=1e5, tolerance=1e-8)
    sol = O.td_solve(kern, 1.0, 0.0, e["gamma"], 0.0, e["F0"], e["dF0"], **kw)
    et = P.tagged_newtonian(p)
    tk = O.tagged3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"], sol["F"])
    sols = O.td_solve(tk, 1.0, 0.0, et["gamma"], 0.0, et["F0"], et["dF0"], **kw)
    return p, sol, sols, kw


def test_collective_golden_test_tagged_59(tagged_chain):
    _, sol, _, _ = tagged_chain
    assert sol["F"].sum() == pytest.approx(25492.648222645254, rel=RTOL)


def test_msd_chain_golden_test_tagged_71(tagged_chain):
    p, sol, sols, kw = tagged_chain
    Kmsd = O.msd3d_table(p["rho"], 1.0, 1.0, p["k"], p["Sk"], sol["F"], sols["F"])
    msd = O.td_solve(O.tabulated(Kmsd), 1.0, 0.0, 0.0, -6.0, [0.0], [0.0], scalar_mode=True, **kw)
    assert msd["F"].sum() == pytest.approx(51.75329405084479, rel=RTOL)


def test_steady_state_golden_test_steady_state_64():
    p = P.hard_spheres(100, 0.51595, rho_roundtrip=True)
    kern = O.sc3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"])
    ss = O.steady_state(kern, p["k"] ** 2 / p["Sk"], p["Sk"], tolerance=1e-8)
    assert ss["status"] == 0
    assert ss["F"].sum() == pytest.approx(19.285439116785284, rel=RTOL)


def test_scalar_f2_steady_state_test_steady_state_8():
    ss = O.steady_state(O.f2(4.00000001), [1.0], [1.0], tolerance=1e-10)
    assert abs(ss["F"][0] - 0.5) < 1e-3


@pytest.mark.skipif(not os.path.exists("/root/reference/test/Sk_MC2.txt"), reason="reference fixture not on this box")
def test_mc_steady_state_golden_test_steady_state_104():
    Nk = 100
    diam = np.array([0.8, 1.0]); cr = np.array([0.2, 0.8]); cr /= cr.sum()
    rho = 6 * 0.55 / (np.pi * np.sum(cr * diam ** 3)) * cr
    x = rho / rho.sum()
    k = P.sf.wavenumber_grid(Nk)
    data = np.loadtxt("/root/reference/test/Sk_MC2.txt").reshape(-1)
    Sk = np.transpose(data.reshape((2, 2, 100), order="F") * np.sqrt(x[:, None, None] * x[None, :, None]), (2, 0, 1)).copy()
    gam = np.einsum("kab,kbc->kac", np.array([np.diag(k[i] ** 2 * x) for i in range(Nk)]), np.linalg.inv(Sk))
    ss = O.steady_state(O.mc3d(rho, 1.0, np.ones(2), k, Sk), O._blocks(gam), O._blocks(Sk), tolerance=1e-12)
    assert ss["F"].sum() == pytest.approx(45.50146332163797, rel=RTOL)


@pytest.fixture(scope="module")
def mixture_chain():
    b = P.binary_mixture()
    Nk = b["Nk"]
    kern = O.mc3d(b["rho"], 1.0, b["m"], b["k"], b["Sk"])
    kw = dict(N=4, dt=1e-4, t_max=1e3, tolerance=1e-12, max_iterations=20000)
    sol = O.td_solve(kern, 1.0, 0.0, O._blocks(b["gamma"]), np.zeros(Nk * 4), O._blocks(b["Sk"]), np.zeros(Nk * 4), **kw)
    return b, kern, sol, kw


def test_mc_fast_kernel_equals_brute_force_test_MCMCT_192(mixture_chain):
    b, kern, _, _ = mixture_chain
    naive = O.mc3d(b["rho"], 1.0, b["m"], b["k"], b["Sk"], naive=True)
    F = O._blocks(np.random.default_rng(0).random((b["Nk"], 2, 2)))
    a, c = kern.evaluate(F), naive.evaluate(F)
    assert np.allclose(a, c, rtol=RTOL, atol=0)


def test_mc_time_doubling_golden_test_MCMCT_211(mixture_chain):
    b, _, sol, _ = mixture_chain
    Fend = sol["F"][-1].reshape(b["Nk"], 4)[18].reshape(2, 2, order="F")
    gold = np.array([[0.13881069002073976, 0.04192527228732725], [0.041925272287327266, 0.52419581715369]])
    assert np.allclose(Fend, gold, rtol=RTOL, atol=0)


def test_mc_tagged_and_msd_goldens_test_MCMCT_212_226(mixture_chain):
    b, _, sol, kw = mixture_chain
    Nk = b["Nk"]
    traj = sol["F"].reshape(-1, Nk, 4)
    tk = O.tagged_mc3d(1, b["rho"], 1.0, b["m"], b["k"], b["Sk"], traj)
    ts = O.td_solve(tk, 1.0, 0.0, b["k"] ** 2 / b["m"][1], 0.0, np.ones(Nk), np.zeros(Nk), **kw)
    assert ts["F"][-1][18] == pytest.approx(0.7678799482793754, rel=RTOL)
    Kt = O.msd_mc3d_table(1, b["rho"], 1.0, b["m"], b["k"], b["Sk"], traj, ts["F"])
    ms = O.td_solve(O.tabulated(Kt), 1.0, 0.0, 0.0, -6.0, [0.0], [0.0], scalar_mode=True, **kw)
    assert ms["F"].sum() == pytest.approx(0.6057207573375025, rel=RTOL)


def test_ddim_matches_bengtzelius_test_dDimMCT_consistency_69():
    p = P.hard_spheres(20, 0.50)
    kd = O.ddim(p["rho"], 1.0, 1.0, p["k"], p["Sk"], 3)
    ks = O.sc3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"])
    assert np.abs(kd.evaluate(p["Sk"]) - ks.evaluate(p["Sk"])).sum() < 1e-8


def test_ddim_critical_point_2d_test_dDimMCT_crit_pt_96_97():
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
        out[phi] = O.steady_state(O.ddim(rho, 1.0, 1.0, k, Sk, 2), k ** 2 / Sk, Sk, tolerance=1e-8)["F"].sum()
    assert out[0.70] == pytest.approx(17.987734457716535, rel=RTOL)
    assert out[0.69] < 1e-7


def test_relaxation_time_test_relaxationtime_6_8():
    t = 10 ** np.linspace(-10, 10, 10000)
    F = np.exp(-t / 1e5) * 6
    assert abs(O.relaxation_time(t, F, mode="log") - 1e5) < 0.1
    assert abs(O.relaxation_time(t, F, mode="lin") - 1e5) < 1.0


def test_active_kernel_equals_passive_kernel_test_ActiveMCT_16_22():
    """ActiveMCTKernel with w(k) = w0 = 1 is the passive 3D kernel (test/test_ActiveMCT.jl:1-24 checks that through the
    solves); here: the oracle's loop restatement against the Bengtzelius oracle, and the host table builder of the
    product against the loops."""
    p = P.hard_spheres(20, 0.51 * np.pi / 6, kmax=20.0)
    rho = 0.51
    pref, J, V2 = O.active_tables(rho, p["k"], np.ones(20), 1.0, p["Sk"], 3)
    F = p["Sk"] * np.exp(-0.1 * p["k"])
    Ka = O.active_evaluate(pref, J, V2, F)
    Kp = O.sc3d(rho, 1.0, 1.0, p["k"], p["Sk"]).evaluate(F)
    assert np.abs(Ka - Kp).sum() < 1e-8 * np.abs(Kp).sum()
    from conftest import load_pkg
    pkg = load_pkg()
    for dim in (3, 2):
        a = pkg.active_tables(rho, p["k"], np.linspace(0.8, 1.2, 20), 0.9, p["Sk"], dim)
        b = O.active_tables(rho, p["k"], np.linspace(0.8, 1.2, 20), 0.9, p["Sk"], dim)
        assert a[0] == pytest.approx(b[0], rel=1e-15)
        assert np.allclose(a[1], b[1], rtol=1e-13, atol=0) and np.allclose(a[2], b[2], rtol=1e-13, atol=0)


def test_ddim_critical_point_5d_test_dDimMCT_crit_pt_133_134():
    """d = 5, Percus-Yevick direct correlation of Leutheusser (test/test_dDimMCT_crit_pt.jl:25-47, :100-134)."""
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
        out[phi] = O.steady_state(O.ddim(rho, 1.0, 1.0, k, Sk, 5), k ** 2 / Sk, Sk, tolerance=1e-8)["F"].sum()
    assert out[0.26] == pytest.approx(33.463940782589255, rel=RTOL)
    assert out[0.25] < 1e-7


def test_mc_with_one_species_equals_sc_kernel_test_MCTvsMCMCT_25():
    p = P.hard_spheres(100, 0.50, rho_roundtrip=True)
    F = np.random.default_rng(5).random(100)
    sc = O.sc3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"]).evaluate(F)
    Sk1 = p["Sk"].reshape(100, 1, 1)
    mc = O.mc3d([p["rho"]], 1.0, [1.0], p["k"], Sk1).evaluate(F)
    nv = O.mc3d([p["rho"]], 1.0, [1.0], p["k"], Sk1, naive=True).evaluate(F)
    assert np.allclose(mc, sc, rtol=RTOL, atol=0) and np.allclose(nv, sc, rtol=RTOL, atol=0)


@pytest.mark.parametrize("ns", [3, 4])
def test_mc_fast_kernel_equals_brute_force_for_3_and_4_species(ns):
    """The reference pins its multi-component kernel with a brute-force triangle sum at Ns = 2 only
    (test/test_MCMCT.jl:6-113, :192); the same cross-check here at the species counts of BASELINE config 4."""
    diam = np.linspace(0.7, 1.0, ns)
    cr = np.arange(1, ns + 1, dtype=float); cr /= cr.sum()
    rho = 6 * 0.45 / (np.pi * np.sum(cr * diam ** 3)) * cr
    m = np.linspace(1.0, 1.5, ns)
    k = P.sf.wavenumber_grid(24)
    Sk = P.sf.percus_yevick_mixture_sk(k, diam, rho)
    F = O._blocks(np.random.default_rng(ns).random((24, ns, ns)))
    a = O.mc3d(rho, 1.3, m, k, Sk).evaluate(F)
    b = O.mc3d(rho, 1.3, m, k, Sk, naive=True).evaluate(F)
    assert np.allclose(a, b, rtol=RTOL, atol=1e-13 * np.abs(b).max())
