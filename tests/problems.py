"""Shared problem set-ups (inputs only) for the parity tests; each mirrors a reference test or BASELINE config."""
import numpy as np

from conftest import load_pkg

sf = load_pkg().structure_factors


def hard_spheres(Nk, eta, kmax=40.0, rho_roundtrip=False):
    """PY hard spheres on the reference grid (test/test_MCT.jl:34-49).  rho_roundtrip reproduces
    `find_analytical_S_k(k_array, ρ*π/6)` with ρ = η*6/π as written in test_MCT.jl:35,42."""
    rho = eta * 6 / np.pi
    k = sf.wavenumber_grid(Nk, kmax)
    Sk = sf.percus_yevick_sk(k, rho * np.pi / 6 if rho_roundtrip else eta)
    return dict(Nk=Nk, eta=eta, rho=rho, kBT=1.0, m=1.0, k=k, Sk=Sk)


def overdamped(p):
    """alpha=0, beta=1, gamma=k^2 kBT/(m S), delta=0, F0=S, dF0=0 (test/test_MCT.jl:44-49)."""
    return dict(alpha=0.0, beta=1.0, gamma=p["k"] ** 2 * p["kBT"] / (p["m"] * p["Sk"]), delta=0.0, F0=p["Sk"].copy(),
                dF0=np.zeros(p["Nk"]))


def newtonian(p):
    """alpha=1, beta=0 (test/test_tagged.jl:39-41)."""
    d = overdamped(p)
    d.update(alpha=1.0, beta=0.0)
    return d


def tagged_newtonian(p):
    """F0=1, gamma=k^2 kBT/m (test/test_tagged.jl:49-53)."""
    return dict(alpha=1.0, beta=0.0, gamma=p["k"] ** 2 * p["kBT"] / p["m"], delta=0.0, F0=np.ones(p["Nk"]), dF0=np.zeros(p["Nk"]))


def binary_mixture(Nk=50, phi=0.55):
    """test/test_MCMCT.jl:156-185."""
    diam = np.array([0.8, 1.0]); cr = np.array([0.2, 0.8]); cr = cr / cr.sum()
    rho_all = 6 * phi / (np.pi * np.sum(cr * diam ** 3))
    rho = rho_all * cr
    k = sf.wavenumber_grid(Nk)
    x = rho / rho.sum()
    m = np.ones(2)
    Sk = sf.percus_yevick_mixture_sk(k, diam, rho)
    Sinv = np.linalg.inv(Sk)
    J = np.array([np.diag(k[i] ** 2 * x / m) for i in range(Nk)])
    gamma = np.einsum("kab,kbc->kac", J, Sinv)
    return dict(Nk=Nk, Ns=2, rho=rho, kBT=1.0, m=m, k=k, Sk=Sk, gamma=gamma, x=x)


def exact_sc3d_eval(V, k, F1, F2):
    """Extended-precision (x87 long double) evaluation of the Bengtzelius recursion
    T_m[ik] = T_m[ik-1] + sum_iq (A[iq,iq+ik-1] + A[iq+ik-1,iq]) - sum_{iq<ik} A[iq,ik-iq]   (docs/src/MCT.md:266-274)
    with A_m = V_m .* (F1 F2'), K = k T1 + T2/k^3 + T3/k.  Arbiter between two float64 summation orders."""
    Nk = len(k)
    ld = np.longdouble
    Ts = []
    for Vm in V:
        A = np.asarray(Vm).astype(ld) * (np.asarray(F1).astype(ld)[:, None] * np.asarray(F2).astype(ld)[None, :])
        Aflip = A[:, ::-1]
        T = np.zeros(Nk, dtype=ld)
        T[0] = np.trace(A)
        for ik in range(2, Nk + 1):
            d = ik - 1
            inc = np.trace(A, offset=d) + np.trace(A, offset=-d)
            # antidiagonal iq + ip = ik (1-based), i.e. a + b = ik - 2 (0-based): offset in the column-flipped matrix
            inc -= np.trace(Aflip, offset=(Nk - 1) - (ik - 2))
            T[ik - 1] = T[ik - 2] + inc
        Ts.append(T)
    kl = np.asarray(k).astype(ld)
    return kl * Ts[0] + Ts[1] / kl ** 3 + Ts[2] / kl
