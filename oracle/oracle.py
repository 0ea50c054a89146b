"""ctypes front-end of the CPU oracle (oracle/mct_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
``--impl reference`` legs of bench.py -- never by the product package.  Parity status: pinned against the
reference's regression goldens, see tests/test_oracle_goldens.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

c_dp = C.POINTER(C.c_double)
c_lp = C.POINTER(C.c_long)


def build(force=False):
    src = os.path.join(_HERE, "mct_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        try:
            _lib = C.CDLL(_SO)
        except OSError:
            build(force=True)
            _lib = C.CDLL(_SO)
        L = _lib
        for name in ("ok_sc3d_new", "ok_sc3d_from_tables", "ok_tagged3d_new", "ok_tagged3d_from_tables", "ok_mc3d_new",
                     "ok_mc3d_naive_new", "ok_tagged_mc3d_new", "ok_ddim_new", "ok_f2_new", "ok_tabulated_new"):
            getattr(L, name).restype = C.c_void_p
        L.ok_td_output_len.restype = C.c_long
        L.ok_relaxation_time.restype = C.c_double
        L.ok_kernel_eval_repeat.restype = C.c_double
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


class Kernel:
    """Owns an ok_kernel*.  ``dof`` = Nk*ns*ns doubles in and out of evaluate()."""

    def __init__(self, ptr, Nk, ns, keep=()):
        self.ptr = C.c_void_p(ptr)
        self.Nk, self.ns = Nk, ns
        self._keep = keep

    def __del__(self):
        if getattr(self, "ptr", None) and _lib is not None:
            _lib.ok_kernel_free(self.ptr)
            self.ptr = None

    def evaluate(self, F, t_index=0):
        F, Fp = _d(F)
        out = np.empty(self.Nk * self.ns * self.ns)
        lib().ok_kernel_eval(self.ptr, Fp, C.c_long(t_index), out.ctypes.data_as(c_dp))
        return out

    def tables(self):
        n = self.Nk
        V = [np.empty((n, n), order="F") for _ in range(3)]
        lib().ok_get_tables(self.ptr, *[v.ctypes.data_as(c_dp) for v in V])
        return V

    def eval_repeat(self, F, n):
        F, Fp = _d(F)
        out = np.empty(self.Nk * self.ns * self.ns)
        return lib().ok_kernel_eval_repeat(self.ptr, Fp, C.c_int(n), out.ctypes.data_as(c_dp))


def sc3d(rho, kBT, m, k, Sk):
    k, kp = _d(k); Sk, Sp = _d(Sk)
    return Kernel(lib().ok_sc3d_new(len(k), C.c_double(rho), C.c_double(kBT), C.c_double(m), kp, Sp), len(k), 1)


def sc3d_from_tables(k, V1, V2, V3):
    k, kp = _d(k)
    Vs = [np.asfortranarray(v, dtype=np.float64) for v in (V1, V2, V3)]
    return Kernel(lib().ok_sc3d_from_tables(len(k), kp, *[v.ctypes.data_as(c_dp) for v in Vs]), len(k), 1)


def tagged3d(rho, kBT, m, k, Sk, Fc_traj):
    k, kp = _d(k); Sk, Sp = _d(Sk); Fc, Fp = _d(Fc_traj)
    return Kernel(lib().ok_tagged3d_new(len(k), C.c_double(rho), C.c_double(kBT), C.c_double(m), kp, Sp,
                                        C.c_long(Fc.shape[0]), Fp), len(k), 1)


def tagged3d_from_tables(k, V1, V2, V3, Fc_traj):
    k, kp = _d(k); Fc, Fp = _d(Fc_traj)
    Vs = [np.asfortranarray(v, dtype=np.float64) for v in (V1, V2, V3)]
    return Kernel(lib().ok_tagged3d_from_tables(len(k), kp, *[v.ctypes.data_as(c_dp) for v in Vs],
                                                C.c_long(Fc.shape[0]), Fp), len(k), 1)


def _blocks(Sk):
    """[Nk, ns, ns] numpy (S[k][a,b]) -> flat [Nk][ns*ns] column-major blocks."""
    Sk = np.asarray(Sk, dtype=np.float64)
    return np.ascontiguousarray(np.transpose(Sk, (0, 2, 1))).reshape(Sk.shape[0], -1)


def mc3d(rho, kBT, m, k, Sk, naive=False):
    k, kp = _d(k); rho, rp = _d(rho); m, mp = _d(m)
    ns = len(rho)
    Sb, Sp = _d(_blocks(Sk))
    fn = lib().ok_mc3d_naive_new if naive else lib().ok_mc3d_new
    return Kernel(fn(len(k), ns, rp, C.c_double(kBT), mp, kp, Sp), len(k), ns)


def tagged_mc3d(s0, rho, kBT, m, k, Sk, Fc_traj_blocks):
    """Fc_traj_blocks: [nt][Nk][ns*ns] column-major blocks of the collective multi-component solution."""
    k, kp = _d(k); rho, rp = _d(rho); m, mp = _d(m)
    ns = len(rho)
    Sb, Sp = _d(_blocks(Sk)); Fc, Fp = _d(Fc_traj_blocks)
    return Kernel(lib().ok_tagged_mc3d_new(s0, len(k), ns, rp, C.c_double(kBT), mp, kp, Sp, C.c_long(Fc.shape[0]), Fp),
                  len(k), 1)


def ddim(rho, kBT, m, k, Sk, d, Fc_traj=None):
    k, kp = _d(k); Sk, Sp = _d(Sk)
    if Fc_traj is None:
        ptr = lib().ok_ddim_new(len(k), C.c_double(rho), C.c_double(kBT), C.c_double(m), kp, Sp, int(d), 0, C.c_long(0), None)
    else:
        Fc, Fp = _d(Fc_traj)
        ptr = lib().ok_ddim_new(len(k), C.c_double(rho), C.c_double(kBT), C.c_double(m), kp, Sp, int(d), 1,
                                C.c_long(Fc.shape[0]), Fp)
    return Kernel(ptr, len(k), 1)


def ddim_tables(kern):
    n = kern.Nk
    V = np.empty((n, n, n), order="F"); J = np.empty((n, n, n), order="F")
    pf = C.c_double()
    lib().ok_ddim_get_tables(kern.ptr, V.ctypes.data_as(c_dp), J.ctypes.data_as(c_dp), C.byref(pf))
    return V, J, pf.value


def f2(nu):
    return Kernel(lib().ok_f2_new(C.c_double(nu)), 1, 1)


def tabulated(Ktab):
    Kt, Kp = _d(Ktab)
    return Kernel(lib().ok_tabulated_new(C.c_long(len(Kt)), Kp), 1, 1)


def msd3d_table(rho, kBT, m, k, Sk, F, Fs):
    k, kp = _d(k); Sk, Sp = _d(Sk); F, Fp = _d(F); Fs, Fsp = _d(Fs)
    out = np.empty(F.shape[0])
    lib().ok_msd3d_table(len(k), C.c_double(rho), C.c_double(kBT), C.c_double(m), kp, Sp, C.c_long(F.shape[0]), Fp, Fsp,
                         out.ctypes.data_as(c_dp))
    return out


def msd_mc3d_table(s0, rho, kBT, m, k, Sk, F_blocks, Fs):
    k, kp = _d(k); rho, rp = _d(rho); m, mp = _d(m)
    Sb, Sp = _d(_blocks(Sk)); F, Fp = _d(F_blocks); Fs, Fsp = _d(Fs)
    out = np.empty(F.shape[0])
    lib().ok_msd_mc3d_table(s0, len(k), len(rho), rp, C.c_double(kBT), mp, kp, Sp, C.c_long(F.shape[0]), Fp, Fsp,
                            out.ctypes.data_as(c_dp))
    return out


def output_len(N, dt, tmax):
    return lib().ok_td_output_len(int(N), C.c_double(dt), C.c_double(tmax))


def _coef(c, Nk, ns):
    """scalar -> (0, lambda, NULL); array [Nk] (ns=1) or [Nk][ns*ns] -> (1, 0, ptr)."""
    if np.isscalar(c):
        return 0, float(c), None, None
    a = np.ascontiguousarray(c, dtype=np.float64).reshape(Nk, ns * ns)
    return 1, 0.0, a, a.ctypes.data_as(c_dp)


def td_solve(kernel, alpha, beta, gamma, delta, F0, dF0, N=32, dt=1e-10, t_max=1e10, max_iterations=10**4,
             tolerance=1e-10, scalar_mode=False, init_with_memory=True, kernel_tmap=None, max_evals=0):
    """solve(MemoryEquation(alpha,beta,gamma,delta,F0,dF0,kernel), TimeDoublingSolver(...)).

    Returns dict(t, F [nt, dof], K [nt, dof], kernel_evals, status)."""
    Nk, ns = kernel.Nk, kernel.ns
    dof = Nk * ns * ns
    F0, F0p = _d(np.asarray(F0, dtype=np.float64).reshape(dof))
    dF0, dF0p = _d(np.asarray(dF0, dtype=np.float64).reshape(dof))
    if np.isscalar(delta):
        delta = np.full(dof, float(delta))
    delta, dp = _d(np.asarray(delta, dtype=np.float64).reshape(dof))
    ak, al, aa, ap = _coef(alpha, Nk, ns)
    bk, bl, ba, bp = _coef(beta, Nk, ns)
    gk, gl, ga, gp = _coef(gamma, Nk, ns)
    nt = output_len(N, dt, t_max)
    t = np.zeros(nt); F = np.zeros((nt, dof)); K = np.zeros((nt, dof))
    evals = C.c_long(0)
    tm = None
    if kernel_tmap is not None:
        tm_arr = np.ascontiguousarray(kernel_tmap, dtype=np.int64)
        tm = tm_arr.ctypes.data_as(c_lp)
    st = lib().ok_td_solve(kernel.ptr, Nk, ns, ak, C.c_double(al), ap, bk, C.c_double(bl), bp, gk, C.c_double(gl), gp,
                           dp, F0p, dF0p, int(N), C.c_double(dt), C.c_double(t_max), C.c_long(max_iterations),
                           C.c_double(tolerance), int(bool(scalar_mode)), int(bool(init_with_memory)), tm,
                           t.ctypes.data_as(c_dp), F.ctypes.data_as(c_dp), K.ctypes.data_as(c_dp), C.byref(evals),
                           C.c_long(max_evals))
    return dict(t=t, F=F, K=K, kernel_evals=evals.value, status=st)


def steady_state(kernel, gamma, F0, tolerance=1e-8, max_iterations=10**6):
    Nk, ns = kernel.Nk, kernel.ns
    dof = Nk * ns * ns
    g, gp = _d(np.asarray(gamma, dtype=np.float64).reshape(dof))
    F0, F0p = _d(np.asarray(F0, dtype=np.float64).reshape(dof))
    F = np.empty(dof); K = np.empty(dof); it = C.c_long(0)
    st = lib().ok_steady_state(kernel.ptr, Nk, ns, gp, F0p, C.c_double(tolerance), C.c_long(max_iterations),
                               F.ctypes.data_as(c_dp), K.ctypes.data_as(c_dp), C.byref(it))
    return dict(F=F, K=K, iterations=it.value, status=st)


def relaxation_time(t, F, threshold=np.exp(-1), mode="log"):
    t, tp = _d(t); F, Fp = _d(F)
    return lib().ok_relaxation_time(C.c_long(len(t)), tp, Fp, C.c_double(threshold), 0 if mode == "log" else 1)


# ---- ActiveMCTKernel (src/Kernels/ActiveMCTKernel.jl:28-61): pure-Python restatement in the reference's loop order, for
#      small Nk only (test infrastructure like everything in this file).
def active_tables(rho, k, wk, w0, Sk, dim=3):
    from math import gamma, pi
    k = np.asarray(k, dtype=np.float64); wk = np.asarray(wk, dtype=np.float64); Sk = np.asarray(Sk, dtype=np.float64)
    Nk = len(k)
    dk = k[1] - k[0]
    surf = 2 * pi ** ((dim - 1) / 2) / gamma((dim - 1) / 2)            # surface_d_dim_unit_sphere(dim-1), HelperFunctions.jl:258
    prefactor = dk ** 2 * rho / (2 * (2 * pi) ** dim) * surf             # :30
    mCk = 1 / rho * (1 - wk / (w0 * Sk))                               # :33
    V2 = np.zeros((Nk, Nk, Nk)); J = np.zeros((Nk, Nk, Nk))           # indexed [l, j, i]
    for i in range(1, Nk + 1):
        for j in range(1, Nk + 1):
            for l in range(abs(j - i) + 1, min(i + j - 1, Nk) + 1):   # :38
                kk, q, p = k[i - 1], k[j - 1], k[l - 1]
                cq, cp = mCk[j - 1], mCk[l - 1]
                V = cq / (2 * kk) * (kk ** 2 + q ** 2 - p ** 2) + cp / (2 * kk) * (kk ** 2 + p ** 2 - q ** 2)   # :46
                V2[l - 1, j - 1, i - 1] = wk[i - 1] * V ** 2
                J[l - 1, j - 1, i - 1] = 2 * p * q / ((2 * kk) ** (dim - 2)) * ((q + p - kk) * (kk + p - q) * (kk + q - p) * (kk + p + q)) ** ((dim - 3) / 2)
    return prefactor, J, V2


def active_evaluate(prefactor, J, V2, F):
    Nk = len(F)
    out = np.zeros(Nk)
    for i in range(1, Nk + 1):                                         # :56-58
        for j in range(1, Nk + 1):
            for l in range(abs(j - i) + 1, min(i + j - 1, Nk) + 1):
                out[i - 1] += J[l - 1, j - 1, i - 1] * V2[l - 1, j - 1, i - 1] * F[j - 1] * F[l - 1]
    return out * prefactor                                             # :59


# ---- TaggedActiveMCTKernel (src/Kernels/ActiveMCTKernel.jl:96-136), loop restatement for small Nk
def tagged_active_tables(rho, k, wk, w0, Sk, dim=3):
    from math import gamma, pi
    k = np.asarray(k, dtype=np.float64); wk = np.asarray(wk, dtype=np.float64); Sk = np.asarray(Sk, dtype=np.float64)
    Nk = len(k)
    dk = k[1] - k[0]
    surf = 2 * pi ** ((dim - 1) / 2) / gamma((dim - 1) / 2)
    prefactor = rho * dk ** 2 / ((2 * pi) ** dim) * surf               # :98
    mCk = 1 / rho * (1 - wk / (w0 * Sk))                               # :102
    V2 = np.zeros((Nk, Nk, Nk)); J = np.zeros((Nk, Nk, Nk))
    for i in range(1, Nk + 1):
        for j in range(1, Nk + 1):
            for l in range(abs(j - i) + 1, min(i + j - 1, Nk) + 1):   # :107
                kk, q, p = k[i - 1], k[j - 1], k[l - 1]
                V2[l - 1, j - 1, i - 1] = (mCk[l - 1] / (2 * kk) * (kk ** 2 + p ** 2 - q ** 2)) ** 2 * w0          # :112
                J[l - 1, j - 1, i - 1] = 2 * p * q / ((2 * kk) ** (dim - 2)) * ((q + p - kk) * (kk + p - q) * (kk + q - p) * (kk + p + q)) ** ((dim - 3) / 2)
    return prefactor, J, V2


def tagged_active_evaluate(prefactor, J, V2, F_col, Fs):
    Nk = len(Fs)
    out = np.zeros(Nk)
    for i in range(1, Nk + 1):                                         # :131-133
        for j in range(1, Nk + 1):
            for l in range(abs(j - i) + 1, min(i + j - 1, Nk) + 1):
                out[i - 1] += J[l - 1, j - 1, i - 1] * V2[l - 1, j - 1, i - 1] * F_col[l - 1] * Fs[j - 1]
    return out * prefactor
