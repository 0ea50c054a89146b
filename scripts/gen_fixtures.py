#!/usr/bin/env python
"""Generates tests/golden/*.npz: CPU-oracle solutions of the BENCHMARKED configurations (BASELINE.json configs 3, 4, 5)
at sizes the oracle needs minutes to hours for, so that the GPU tests and bench.py can check the CUDA path against them
without running the oracle at those sizes.

    python scripts/gen_fixtures.py c3            # config 3: Nk=1000, eta=0.5155 collective solve + tagged solve + tau_alpha  (~2.5 h, 1 core)
    python scripts/gen_fixtures.py c5 [eta ...]  # config 5: three sweep points at Nk=500 (below, near and above eta_c)
    python scripts/gen_fixtures.py c4            # config 4: Ns=4 mixture at Nk=2000 on a short time axis

The oracle (oracle/mct_oracle.c) is the C restatement of the reference in its own operation order; it is pinned on the
reference's regression goldens by tests/test_oracle_goldens.py.  The reference itself (Julia) cannot run in this image.

Every fixture stores, for the solution arrays X in {F, K} of shape [nt, dof]:
    X_cols   X[:, cols]     whole time series of a few degrees of freedom
    X_rows   X[rows, :]     whole wavenumber profiles at a few times
    X_sum_t  sum_k X[t, k]  one checksum per time      (localises a wrong time step)
    X_sum_k  sum_t X[t, k]  one checksum per dof       (localises a wrong wavenumber)
plus t, kernel_evals, the inputs' parameters and (single component) tau_alpha at the peak wavenumber.
"""
import importlib.util
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O   # noqa: E402

spec = importlib.util.spec_from_file_location("mct_sf", os.path.join(ROOT, "modecouplingtheory.jl_b200", "structure_factors.py"))
sf = importlib.util.module_from_spec(spec); spec.loader.exec_module(sf)
GOLD = os.path.join(ROOT, "tests", "golden")


def digest(res, cols, rows, extra):
    F, K = res["F"], res["K"]
    d = dict(t=res["t"], kernel_evals=np.int64(res["kernel_evals"]), status=np.int64(res["status"]), cols=np.asarray(cols),
             rows=np.asarray(rows))
    for name, X in (("F", F), ("K", K)):
        d[name + "_cols"] = X[:, cols].copy(); d[name + "_rows"] = X[rows, :].copy()
        d[name + "_sum_t"] = X.sum(axis=1); d[name + "_sum_k"] = X.sum(axis=0)
    d.update(extra)
    return d


def pick_rows(nt):
    # t = 0, the bootstrap, early / middle / late blocks, the last time
    return sorted(set([0, 1, 17, 64, 65, 100, nt // 4, nt // 2, (3 * nt) // 4, nt - 130, nt - 2, nt - 1]))


def sc_problem(Nk, eta):
    k = sf.wavenumber_grid(Nk)
    Sk = sf.percus_yevick_sk(k, eta)
    return dict(Nk=Nk, eta=eta, rho=6 * eta / np.pi, k=k, Sk=Sk, gamma=k ** 2 / Sk)


def sc_solve(p, **kw):
    kern = O.sc3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"])
    return O.td_solve(kern, 0.0, 1.0, p["gamma"], 0.0, p["Sk"], np.zeros(p["Nk"]), max_iterations=10 ** 6, **kw)


def sc_cols(p):
    kpeak = int(np.argmax(p["Sk"]))
    Nk = p["Nk"]
    return sorted(set([0, 1, Nk // 20, kpeak - 1, kpeak, kpeak + 1, Nk // 3, Nk // 2, (3 * Nk) // 4, Nk - 2, Nk - 1])), kpeak


def gen_c3():
    p = sc_problem(1000, 0.5155)
    cols, kpeak = sc_cols(p)
    t0 = time.time()
    res = sc_solve(p)
    print(f"c3 collective: {res['kernel_evals']} evals, status {res['status']}, {time.time() - t0:.0f} s", flush=True)
    nt = len(res["t"])
    tau = O.relaxation_time(res["t"], res["F"][:, kpeak])
    np.savez_compressed(os.path.join(GOLD, "c3_nk1000_eta0.5155.npz"),
             **digest(res, cols, pick_rows(nt), dict(Nk=1000, eta=0.5155, kpeak=kpeak, tau_alpha=tau)))
    # tagged particle: F0 = 1, gamma = k^2 kBT/m, overdamped like the collective solve (test/test_tagged.jl:49-53 set-up)
    t0 = time.time()
    tk = O.tagged3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"], res["F"])
    rs = O.td_solve(tk, 0.0, 1.0, p["k"] ** 2, 0.0, np.ones(1000), np.zeros(1000), max_iterations=10 ** 6,
                    kernel_tmap=np.arange(nt, dtype=np.int64))
    print(f"c3 tagged: {rs['kernel_evals']} evals, status {rs['status']}, {time.time() - t0:.0f} s", flush=True)
    taus = O.relaxation_time(rs["t"], rs["F"][:, kpeak])
    np.savez_compressed(os.path.join(GOLD, "c3_tagged_nk1000_eta0.5155.npz"),
             **digest(rs, cols, pick_rows(nt), dict(Nk=1000, eta=0.5155, kpeak=kpeak, tau_alpha=taus)))


def gen_c5(etas=(0.50, 0.5155, 0.52)):
    for eta in etas:
        p = sc_problem(500, eta)
        cols, kpeak = sc_cols(p)
        t0 = time.time()
        res = sc_solve(p)
        print(f"c5 eta={eta}: {res['kernel_evals']} evals, status {res['status']}, {time.time() - t0:.0f} s", flush=True)
        tau = O.relaxation_time(res["t"], res["F"][:, kpeak])
        ss = O.steady_state(O.sc3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"]), p["gamma"], p["Sk"], tolerance=1e-12, max_iterations=10 ** 6)
        np.savez_compressed(os.path.join(GOLD, f"c5_nk500_eta{eta}.npz"),
                 **digest(res, cols, pick_rows(len(res["t"])), dict(Nk=500, eta=eta, kpeak=kpeak, tau_alpha=tau, f_steady=ss["F"],
                                                                   steady_iterations=np.int64(ss["iterations"]))))


def mixture4(Nk):
    """BASELINE config 4: diameters 0.7/0.8/0.9/1.0, equal number fractions, phi = 0.50, Newtonian."""
    ns = 4
    diam = np.array([0.7, 0.8, 0.9, 1.0]); x = np.full(ns, 0.25)
    rho = 6 * 0.50 / (np.pi * np.sum(x * diam ** 3)) * x
    m = np.ones(ns)
    k = sf.wavenumber_grid(Nk)
    Sk = sf.percus_yevick_mixture_sk(k, diam, rho)
    xx = rho / rho.sum()
    gamma = np.einsum("kab,kbc->kac", np.array([np.diag(k[i] ** 2 * xx / m) for i in range(Nk)]), np.linalg.inv(Sk))
    return dict(Nk=Nk, ns=ns, rho=rho, m=m, k=k, Sk=Sk, gamma=gamma)


def gen_c4(Nk=2000, N=4, dt=1e-4, t_max=1e-2):
    p = mixture4(Nk)
    ns = p["ns"]
    kern = O.mc3d(p["rho"], 1.0, p["m"], p["k"], p["Sk"])
    F0 = O._blocks(p["Sk"]); g = O._blocks(p["gamma"])
    t0 = time.time()
    res = O.td_solve(kern, 1.0, 0.0, g, 0.0, F0, np.zeros_like(F0), N=N, dt=dt, t_max=t_max, tolerance=1e-10, max_iterations=10 ** 6)
    print(f"c4 Nk={Nk}: {res['kernel_evals']} evals, status {res['status']}, {time.time() - t0:.0f} s", flush=True)
    nt = len(res["t"])
    kk = [0, 1, Nk // 20, Nk // 6, Nk // 3, Nk // 2, Nk - 1]
    cols = sorted(set(int(i * ns * ns + e) for i in kk for e in (0, 1, 6, 15)))
    rows = sorted(set([0, 1, 2 * N, 2 * N + 1, nt // 2, nt - 2, nt - 1]))
    np.savez_compressed(os.path.join(GOLD, f"c4_ns4_nk{Nk}_short.npz"),
             **digest(res, cols, rows, dict(Nk=Nk, ns=ns, N=N, dt=dt, t_max=t_max)))


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "c3":
        gen_c3()
    elif what == "c5":
        gen_c5(*([tuple(float(a) for a in sys.argv[2:])] if len(sys.argv) > 2 else []))
    elif what == "c4":
        gen_c4(*(json.loads(a) for a in sys.argv[2:]))
    else:
        raise SystemExit(__doc__)
