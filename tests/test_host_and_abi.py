"""CPU-side checks: the C-ABI library loads and exports every symbol include/mct.h declares (no compute call
without a GPU), and the host logic of the package agrees with the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle as O
import problems as P
from conftest import ROOT, load_pkg

pkg = load_pkg()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mct.h")).read()
    return sorted(set(re.findall(r"MCT_API[^;(]*?\b(mct_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import subprocess, sys
    subprocess.check_call([sys.executable, os.path.join(ROOT, "modecouplingtheory.jl_b200", "build.py")])
    lib = ctypes.CDLL(pkg.library_path())
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(pkg.EXPORTS) == syms


def test_no_cpu_fallback_without_gpu():
    """On a box without a B200 the product must fail loudly instead of computing on the CPU."""
    try:
        n = pkg.device_count()
    except pkg.MCTDeviceError:
        n = 0
    if n > 0:
        pytest.skip("GPU present")
    p = P.hard_spheres(16, 0.5)
    with pytest.raises(pkg.MCTDeviceError):
        pkg.ModeCouplingKernel(p["rho"], 1.0, 1.0, p["k"], p["Sk"])


@pytest.mark.parametrize("tagged", [False, True])
def test_host_vertex_tables_are_bit_identical_to_oracle(tagged):
    p = P.hard_spheres(37, 0.45)
    V = pkg.vertex_tables(p["rho"], 1.0, 1.0, p["k"], p["Sk"], tagged=tagged)
    kern = O.tagged3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"], np.zeros((1, 37))) if tagged else O.sc3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"])
    for a, b in zip(V, kern.tables()):
        assert np.array_equal(a, b)


def test_output_len_matches_reference_loop():
    # 1 + 2N + 2N * #(dt < 2 t_max doublings): defaults give 4417
    assert O.output_len(32, 1e-10, 1e10) == 4417
    assert O.output_len(2, 1e-10, 1e10) == 1 + 4 + 4 * 68


def test_relaxation_time_host():
    t = 10 ** np.linspace(-10, 10, 10000)
    F = np.exp(-t / 1e5) * 6
    assert pkg.find_relaxation_time(t, F, mode="log") == pytest.approx(O.relaxation_time(t, F, mode="log"), rel=1e-14)
    assert pkg.find_relaxation_time(t, F, mode="lin") == pytest.approx(O.relaxation_time(t, F, mode="lin"), rel=1e-14)
    assert pkg.find_relaxation_time(t, np.ones_like(t)) == float("inf")


def test_mixture_structure_factor_closed_form_matches_quadrature():
    """The Baxter q-integral is taken in closed form; check it against scipy quadrature (the reference uses quadgk)."""
    from scipy.integrate import quad
    sf = pkg.structure_factors
    a, b, d, s, K = 1.7, -0.9, 0.9, -0.1, 7.3
    q = lambda r: 0.5 * a * (r * r - d * d) + b * (r - d)
    re_ = quad(lambda r: q(r) * np.cos(K * r), s, d, epsabs=1e-14, epsrel=1e-14)[0]
    im_ = quad(lambda r: q(r) * np.sin(K * r), s, d, epsabs=1e-14, epsrel=1e-14)[0]
    z = sf._baxter_q_integral(K, a, b, d, s)
    assert abs(z.real - re_) < 1e-13 and abs(z.imag - im_) < 1e-13
