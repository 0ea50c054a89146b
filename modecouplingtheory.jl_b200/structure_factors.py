"""Synthetic structure-factor inputs (host-side input generators, not part of the accelerated path).

* Percus-Yevick hard spheres (Wertheim): the closed form the reference's own tests use,
  /root/reference/test/test_MCT.jl:7-31.
* Percus-Yevick hard-sphere mixtures (Baxter factorisation): /root/reference/test/test_MCMCT.jl:116-154.
"""
import numpy as np


def wavenumber_grid(Nk, kmax=40.0):
    """k_i = (i - 1/2) * dk, dk = kmax/Nk  (test/test_MCT.jl:40-41)."""
    dk = kmax / Nk
    return dk * (np.arange(1, Nk + 1) - 0.5)


def percus_yevick_ck(k, eta):
    """Direct correlation function c(k) of PY hard spheres (test/test_MCT.jl:7-18)."""
    k = np.asarray(k, dtype=np.float64)
    A = -(1 - eta) ** -4 * (1 + 2 * eta) ** 2
    B = (1 - eta) ** -4 * 6 * eta * (1 + eta / 2) ** 2
    D = -(1 - eta) ** -4 * 1 / 2 * eta * (1 + 2 * eta) ** 2
    k2 = k * k
    k4 = k2 * k2
    k6 = k4 * k2
    return 4 * np.pi / k6 * (
        24 * D - 2 * B * k2 - (24 * D - 2 * (B + 6 * D) * k2 + (A + B + D) * k4) * np.cos(k)
        + k * (-24 * D + (A + 2 * B + 4 * D) * k2) * np.sin(k)
    )


def percus_yevick_sk(k, eta):
    """Static structure factor S(k) of PY hard spheres at packing fraction eta (test/test_MCT.jl:26-31)."""
    ck = percus_yevick_ck(k, eta)
    rho = 6 / np.pi * eta
    return 1 + rho * ck / (1 - rho * ck)


def _baxter_q_integral(K, a, b, dik, sik):
    """int_{sik}^{dik} q(r) exp(iKr) dr with q(r) = a/2 (r^2 - dik^2) + b (r - dik), in closed form."""
    def anti(r):
        e = np.exp(1j * K * r)
        i0 = e / (1j * K)
        i1 = e * (r / (1j * K) + 1 / K ** 2)
        i2 = e * (r * r / (1j * K) + 2 * r / K ** 2 - 2 / (1j * K ** 3))
        return 0.5 * a * (i2 - dik ** 2 * i0) + b * (i1 - dik * i0)
    return anti(dik) - anti(sik)


def percus_yevick_mixture_sk(k_array, diameters, rho):
    """S_ab(k) (a Nk x Ns x Ns array) of a PY hard-sphere mixture via Baxter's factorisation.

    Restates find_direct_correlation_function_PY / find_structure_factor_PY (test/test_MCMCT.jl:130-154); the
    reference integrates q(r) cis(Kr) with quadgk, here the polynomial-times-phase integral is taken in closed form.
    """
    diameters = np.asarray(diameters, dtype=np.float64)
    rho = np.asarray(rho, dtype=np.float64)
    p = len(diameters)
    d = (diameters[:, None] + diameters[None, :]) / 2
    s = (diameters[:, None] - diameters[None, :]) / 2
    xi = [np.pi / 6 * np.sum(rho * diameters ** nu) for nu in (1, 2, 3)]
    a = (1 - xi[2]) ** -2 * (1 - xi[2] + 3 * xi[1] * diameters)
    b = -3 / 2 * diameters ** 2 * (1 - xi[2]) ** -2 * xi[1]
    x = rho / rho.sum()
    out = np.empty((len(k_array), p, p))
    for n, K in enumerate(np.asarray(k_array, dtype=np.float64)):
        Q = np.empty((p, p), dtype=np.complex128)
        for i in range(p):
            for kk in range(p):
                Q[i, kk] = (i == kk) - 2 * np.pi * np.sqrt(rho[i] * rho[kk]) * _baxter_q_integral(K, a[i], b[i], d[i, kk], s[i, kk])
        Cm = np.eye(p) - np.real(Q.conj().T @ Q)
        Cm = Cm / np.sqrt(rho[:, None] * rho[None, :])
        out[n] = np.linalg.inv(np.eye(p) / x[:, None] - rho.sum() * Cm)
    return out
