"""Host-side mirror of the reference's interface for the accelerated path.

The reference is Julia (not installed in this image), so the host side that would be the Julia glue
(`julia/MCTDevice.jl`) is mirrored here in Python with the same names, argument meaning and
error behaviour, on top of the C ABI of libmct_sm100a.so (include/mct.h):

    ModeCouplingKernel, TaggedModeCouplingKernel           src/Kernels/ModeCouplingKernels.jl:94, :390
    evaluate_kernel, evaluate_kernel_                       src/Kernels.jl:10-26   (`evaluate_kernel!`)
    MemoryEquation                                          src/MemoryEquation.jl:44-52
    TimeDoublingSolver, solve                               src/TimeDoublingSolver.jl:48-50, :512-532
    solve_steady_state                                      src/SteadyStateMemoryEquation.jl:49-56
    get_t, get_F, get_K, find_relaxation_time               src/HelperFunctions.jl:169-253, src/RelaxationTime.jl:14-39

Every numerical operation happens on the GPU.  There is no CPU fallback: if the CUDA library is missing or no
sm_100 device is present the calls raise `MCTDeviceError`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MCT_LIB selects another build of the same library (the instrumented MCT_PROFILE build of build.py, development only)
_LIB_PATH = os.environ.get("MCT_LIB") or os.path.join(_HERE, "lib", "libmct_sm100a.so")
_lib = None

c_dp = C.POINTER(C.c_double)

MCT_OK, MCT_NOT_CONVERGED, MCT_BAD_ARGUMENT, MCT_UNSUPPORTED, MCT_CUDA_ERROR = 0, 1, 2, 3, 4


class MCTDeviceError(RuntimeError):
    """CUDA failure, missing library or missing B200 (MCT_CUDA_ERROR) -- never silently replaced by a CPU path."""


class DomainError(ArithmeticError):
    """Mirrors Julia's DomainError thrown at src/TimeDoublingSolver.jl:406-408."""


class TdOptions(C.Structure):
    _fields_ = [("N", C.c_int), ("dt", C.c_double), ("t_max", C.c_double), ("max_iterations", C.c_long),
                ("tolerance", C.c_double), ("verbose", C.c_int), ("init_with_memory", C.c_int)]


class Coeffs(C.Structure):
    _fields_ = [("alpha_kind", C.c_int), ("alpha_scalar", C.c_double), ("alpha", c_dp),
                ("beta_kind", C.c_int), ("beta_scalar", C.c_double), ("beta", c_dp),
                ("gamma_kind", C.c_int), ("gamma_scalar", C.c_double), ("gamma", c_dp),
                ("delta", c_dp)]


EXPORTS = [
    "mct_last_error", "mct_device_count", "mct_version", "mct_launch_count", "mct_sc3d_create", "mct_sc3d_create_from_sk",
    "mct_sc3d_tagged_create", "mct_sc3d_tagged_create_from_sk", "mct_sc3d_tagged_create_from_equation", "mct_mc3d_create",
    "mct_ddim_create", "mct_kernel_eval", "mct_kernel_eval_many", "mct_kernel_destroy", "mct_kernel_set_collective_times", "mct_td_output_len", "mct_td_solve",
    "mct_equation_create", "mct_equation_solve", "mct_equation_download", "mct_equation_destroy", "mct_equation_solve_batch", "mct_equation_shard_prepare", "mct_equation_shard_connect",
    "mct_steady_state", "mct_msd3d_solve", "mct_equation_download_slice", "mct_equation_relaxation_times",
    "mct_mc3d_tagged_create_from_equation", "mct_msd_mc3d_solve", "mct_active_create", "mct_active_tagged_create_from_equation",
]


def library_path():
    return _LIB_PATH


def load_library():
    """dlopen libmct_sm100a.so (no compute call is made here)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise MCTDeviceError(f"{_LIB_PATH} is missing: run `python __graft_entry__.py build` (there is no CPU fallback)")
        L = C.CDLL(_LIB_PATH)
        L.mct_last_error.restype = C.c_char_p
        L.mct_launch_count.restype = C.c_long
        _lib = L
    return _lib


def _check(rc):
    if rc == MCT_OK:
        return
    msg = load_library().mct_last_error().decode()
    if rc == MCT_NOT_CONVERGED:
        raise DomainError(msg)
    if rc == MCT_BAD_ARGUMENT:
        raise AssertionError(msg)
    if rc == MCT_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise MCTDeviceError(msg)


def _arr(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _p(a):
    return a.ctypes.data_as(c_dp)


def launch_count():
    """CUDA kernels launched by libmct_sm100a.so in this process so far."""
    return load_library().mct_launch_count()


def device_count():
    n = C.c_int(0)
    _check(load_library().mct_device_count(C.byref(n)))
    return n.value


# --------------------------------------------------------------------------------------------- kernels
class MemoryKernel:
    """Opaque device kernel handle (replaces a `MemoryKernel` object, src/Kernels.jl:1)."""

    def __init__(self, handle, Nk, Ns=1, device=0, keep=()):
        self._h = handle
        self.Nk, self.Ns, self.device = Nk, Ns, device
        self._keep = keep

    @property
    def dof(self):
        return self.Nk * self.Ns * self.Ns

    def close(self):
        """mct_kernel_destroy.  Refused (the handle stays valid) while equations created on this kernel are alive."""
        if self._h is not None and _lib is not None:
            if _lib.mct_kernel_destroy(self._h) != MCT_OK:
                raise MCTDeviceError(_lib.mct_last_error().decode())
            self._h = None
            self._keep = ()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def vertex_tables(rho, kBT, m, k_array, Sk, tagged=False):
    """The reference constructor's V1, V2, V3 (Nk x Nk, [iq, ip]) in its own operation order
    (src/Kernels/ModeCouplingKernels.jl:105-126 collective, :402-422 tagged).  Host-side, like the Julia constructor
    that stays untouched; the `from_sk` constructors build the same tables on the device instead."""
    k = _arr(k_array)
    Sk = _arr(Sk)
    dk = k[1] - k[0]
    Ck = (Sk - 1.0) / (rho * Sk)
    D0 = kBT / m
    q, p = k[:, None], k[None, :]
    cq, cp = Ck[:, None], Ck[None, :]
    dd = q * q - p * p
    if not tagged:
        c8 = 8.0 * (np.pi * np.pi)
        s = cp + cq
        dc = cq - cp
        V1 = p * q * (s * s) / 4.0 * D0 * rho / c8 * dk * dk
        V2 = p * q * (dd * dd) * (dc * dc) / 4.0 * D0 * rho / c8 * dk * dk
        V3 = p * q * dd * (cq * cq - cp * cp) / 2.0 * D0 * rho / c8 * dk * dk
    else:
        c4 = 4.0 * (np.pi * np.pi)
        dk2 = dk * dk
        c2 = cq * cq + 0.0 * cp
        V1 = p * q * c2 / 4.0 * D0 * rho / c4 * dk2
        V2 = p * q * (dd * dd) * c2 / 4.0 * D0 * rho / c4 * dk2
        V3 = p * q * dd * c2 / 2.0 * D0 * rho / c4 * dk2
    return [np.asfortranarray(v) for v in (V1, V2, V3)]


def ModeCouplingKernel(rho, kBT, m, k_array, Sk, dims=3, device=0, tables=None):
    """ModeCouplingKernel(ρ, kBT, m, k_array, Sₖ; dims=3)  -- src/Kernels/ModeCouplingKernels.jl:94-129.

    `tables=(V1, V2, V3)` passes tables made elsewhere (what the Julia glue does with `kernel.V1..V3`); by default
    the tables are built on the device from S(k)."""
    if dims != 3:
        raise NotImplementedError("dDimModeCouplingKernel is not implemented on the device yet")
    L = load_library()
    k = _arr(k_array)
    h = C.c_void_p()
    if tables is None:
        Sk = _arr(Sk)
        _check(L.mct_sc3d_create_from_sk(C.byref(h), device, len(k), C.c_double(rho), C.c_double(kBT), C.c_double(m), _p(k), _p(Sk)))
    else:
        V = [np.asfortranarray(v, dtype=np.float64) for v in tables]
        _check(L.mct_sc3d_create(C.byref(h), device, len(k), _p(k), _p(V[0]), _p(V[1]), _p(V[2])))
    return MemoryKernel(h, len(k), 1, device)


def TaggedModeCouplingKernel(rho, kBT, m, k_array, Sk, sol, dims=3, device=0, tables=None):
    """TaggedModeCouplingKernel(ρ, kBT, m, k_array, Sₖ, sol; dims=3) -- src/Kernels/ModeCouplingKernels.jl:390-426.
    `sol` is the collective MemoryEquationSolution; if it still holds its device trajectory the tagged kernel is
    chained on it without a host round trip."""
    if dims != 3:
        raise NotImplementedError("dDimTaggedModeCouplingKernel is not implemented on the device yet")
    L = load_library()
    k = _arr(k_array)
    Sk = _arr(Sk)
    h = C.c_void_p()
    if tables is None and getattr(sol, "_equation", None) is not None:
        _check(L.mct_sc3d_tagged_create_from_equation(C.byref(h), len(k), C.c_double(rho), C.c_double(kBT), C.c_double(m),
                                                      _p(k), _p(Sk), sol._equation._q))
        return MemoryKernel(h, len(k), 1, device, keep=(sol._equation,))
    Fc = _arr(sol.F)
    if tables is None:
        _check(L.mct_sc3d_tagged_create_from_sk(C.byref(h), device, len(k), C.c_double(rho), C.c_double(kBT), C.c_double(m),
                                                _p(k), _p(Sk), C.c_long(Fc.shape[0]), _p(Fc)))
    else:
        V = [np.asfortranarray(v, dtype=np.float64) for v in tables]
        _check(L.mct_sc3d_tagged_create(C.byref(h), device, len(k), _p(k), _p(V[0]), _p(V[1]), _p(V[2]),
                                        C.c_long(Fc.shape[0]), _p(Fc)))
    kern = MemoryKernel(h, len(k), 1, device)
    tc = getattr(sol, "t", None)
    if tc is not None:                       # the keys of the reference's tDict (:394): checked against the tagged solve's grid
        tc = _arr(tc)
        _check(L.mct_kernel_set_collective_times(h, C.c_long(len(tc)), _p(tc)))
    return kern


class MSDKernel:
    """MSDModeCouplingKernel(ρ, kBT, m, k_array, Sₖ, sol, taggedsol; dims=3) -- src/Kernels/ModeCouplingKernels.jl:552-570.
    Holds the two device-resident solutions; `solve` of a scalar MemoryEquation with this kernel runs
    `mct_msd3d_solve` (kernel table reduced on the device, scalar time doubling in one device thread)."""
    Ns = 1
    dof = 1

    def __init__(self, rho, kBT, m, k_array, Sk, sol, taggedsol, species=None, full_species_sum=False):
        if getattr(sol, "_equation", None) is None or getattr(taggedsol, "_equation", None) is None:
            raise MCTDeviceError("MSDModeCouplingKernel needs the device-resident solutions (solve(..., keep_device=True))")
        self.species = species                    # None: single component; else 0-based species of a mixture
        self.full_species_sum = bool(full_species_sum)
        if species is None:
            self.rho, self.kBT, self.m = float(rho), float(kBT), float(m)
            self.Sk = _arr(Sk)
        else:
            self.rho, self.kBT, self.m = float(np.sum(rho)), float(kBT), float(np.asarray(m, dtype=np.float64)[species])
            self.Sk = None
        self.k = _arr(k_array)
        self.sol, self.taggedsol = sol, taggedsol


def MSDModeCouplingKernel(rho, kBT, m, k_array, Sk, sol, taggedsol, dims=3):
    if dims != 3:
        raise NotImplementedError("dDimMSDModeCouplingKernel is not implemented on the device yet")
    return MSDKernel(rho, kBT, m, k_array, Sk, sol, taggedsol)


def TaggedMultiComponentModeCouplingKernel(s, rho, kBT, m, k_array, Sk, sol, dims=3):
    """TaggedMultiComponentModeCouplingKernel(s, ρ, kBT, m, k_array, Sₖ, sol; dims=3) --
    src/Kernels/MultiComponentKernels.jl:268-333.  `s` is the reference's 1-based species index; `sol` must be the
    device-resident solution of the multi-component equation (the species weight is contracted from it on the device)."""
    if dims != 3:
        raise NotImplementedError("MSD is not implemented for dimentions other than 3")
    if getattr(sol, "_equation", None) is None:
        raise MCTDeviceError("TaggedMultiComponentModeCouplingKernel needs the device-resident mixture solution")
    L = load_library()
    m = np.asarray(m, dtype=np.float64)
    if not (1 <= s <= len(m)):
        raise AssertionError("species index out of range")
    h = C.c_void_p()
    _check(L.mct_mc3d_tagged_create_from_equation(C.byref(h), int(s) - 1, C.c_double(float(np.sum(rho))), C.c_double(kBT),
                                                  C.c_double(float(m[s - 1])), sol._equation._q))
    return MemoryKernel(h, len(k_array), 1, sol._equation._kernel.device, keep=(sol._equation,))


def MSDMultiComponentModeCouplingKernel(s, rho, kBT, m, k_array, Sk, sol, taggedsol, dims=3, full_species_sum=False):
    """MSDMultiComponentModeCouplingKernel(s, ρ, kBT, m, k_array, Sₖ, sol, taggedsol) -- MultiComponentKernels.jl:400-468;
    `taggedsol` must come from a TaggedMultiComponentModeCouplingKernel of the same species.  The reference's species
    sum only visits the first species (`Ns = length(Fs[1])` is 1, :477); that behaviour -- and its golden -- is the
    default, full_species_sum=True evaluates the documented sum over all pairs."""
    if dims != 3:
        raise NotImplementedError("MSD is not implemented for dimentions other than 3")
    return MSDKernel(rho, kBT, m, k_array, Sk, sol, taggedsol, species=int(s) - 1, full_species_sum=full_species_sum)


def _solve_msd(equation, solver):
    L = load_library()
    kern = equation.kernel
    for name in ("alpha", "beta", "gamma"):
        if not isinstance(equation.coeffs[name], float):
            raise MCTDeviceError("the MSD equation is scalar: coefficients must be numbers")
    q = kern.sol._equation
    nt = kern.sol.F.shape[0]
    if solver.output_len() != nt:
        raise MCTDeviceError("the MSD solve must use the solver options of the collective solve (same time grid)")
    t = np.empty(nt); F = np.empty(nt); K = np.empty(nt)
    evals = C.c_long(0)
    delta = 0.0 if equation.delta is None else float(equation.delta[0])
    tail = (C.c_double(equation.coeffs["alpha"]), C.c_double(equation.coeffs["beta"]), C.c_double(equation.coeffs["gamma"]),
            C.c_double(delta), C.c_double(float(equation.F0[0])), C.c_double(float(equation.dF0[0])), _p(t), _p(F), _p(K),
            C.byref(evals))
    if kern.species is None:
        rc = L.mct_msd3d_solve(q._q, kern.taggedsol._equation._q, C.c_double(kern.rho), C.c_double(kern.kBT), C.c_double(kern.m),
                               _p(kern.Sk), *tail)
    else:
        rc = L.mct_msd_mc3d_solve(q._q, kern.taggedsol._equation._q, int(kern.species), int(kern.full_species_sum),
                                  C.c_double(kern.rho), C.c_double(kern.kBT), C.c_double(kern.m), *tail)
    solver.kernel_evals = evals.value
    _check(rc)
    return MemoryEquationSolution(t, F, K, solver)


def _to_blocks(X):
    """[Nk, ns, ns] arrays X[k][a, b] (numpy row-major) -> flat [Nk][ns*ns] column-major blocks, the memory of a Julia
    Vector{SMatrix{Ns,Ns}}.  Flat input passes through."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 3:
        return np.ascontiguousarray(np.transpose(X, (0, 2, 1))).reshape(-1)
    return np.ascontiguousarray(X).reshape(-1)


def _from_blocks(flat, Nk, ns):
    """inverse of _to_blocks for one or many time points: [..., Nk*ns*ns] -> [..., Nk, ns, ns] with [a, b] indexing."""
    a = np.asarray(flat).reshape(flat.shape[:-1] + (Nk, ns, ns))
    return np.swapaxes(a, -1, -2)


def MultiComponentModeCouplingKernel(rho, kBT, m, k_array, Sk, dims=3, device=0):
    """MultiComponentModeCouplingKernel(ρ, kBT, m, k_array, Sₖ; dims=3) -- src/Kernels/MultiComponentKernels.jl:84-131.
    `Sk[k]` is the Ns x Ns matrix S_αβ(k)/sqrt(x_α x_β) as in the reference.  The constructor's host work (direct
    correlation blocks C[k] = (δ./x - inv(S[k]))/ρ_all :103-109, prefactor :122-127) is done here like in Julia; the
    kernel evaluation and everything that consumes it runs on the device."""
    if dims != 3:
        raise NotImplementedError("MSD is not implemented for dimentions other than 3")
    L = load_library()
    k = _arr(k_array)
    rho = _arr(rho); m = _arr(m)
    Sk = np.asarray(Sk, dtype=np.float64)
    Nk, ns = len(k), len(rho)
    assert Sk.shape == (Nk, ns, ns) and m.shape == rho.shape
    dk = k[1] - k[0]
    rho_all = rho.sum()
    x = rho / rho.sum()
    Sinv = np.linalg.inv(Sk)
    Ck = (np.eye(ns)[None, :, :] / x[None, :, None] - Sinv) / rho_all
    pref = np.empty((ns, ns))
    for a in range(ns):
        for b in range(ns):
            pref[a, b] = rho_all * kBT / (8 * m[a] * x[b]) * (dk / 2 / np.pi) ** 2
    Cb = _to_blocks(Ck)
    pb = np.ascontiguousarray(pref.T).reshape(-1)            # column-major
    h = C.c_void_p()
    _check(L.mct_mc3d_create(C.byref(h), device, Nk, ns, _p(k), _p(Cb), _p(pb)))
    return MemoryKernel(h, Nk, ns, device)


def _surface_d_dim_unit_sphere(d):
    from math import gamma, pi
    return 2 * pi ** (d / 2) / gamma(d / 2)


def ddim_tables(rho, kBT, m, k_array, Sk, d):
    """prefactor, V[iq, ik, ip], J[iq, ik, ip] of dDimModeCouplingKernel -- src/Kernels/ModeCouplingKernels.jl:34-71."""
    k = _arr(k_array); Sk = _arr(Sk)
    Nk = len(k)
    dk = k[1] - k[0]
    Ck = (1 - 1.0 / Sk) / rho
    prefactor = (kBT / m) * rho * dk ** 2 * _surface_d_dim_unit_sphere(d - 1) / (4 * np.pi) ** d
    iq, ik, ip = np.meshgrid(np.arange(1, Nk + 1), np.arange(1, Nk + 1), np.arange(1, Nk + 1), indexing="ij")
    q, kk, p = k[iq - 1], k[ik - 1], k[ip - 1]
    c_p, c_k = Ck[ip - 1], Ck[ik - 1]
    inside = (np.abs(iq - ik) + 1 <= ip) & (ip <= np.minimum(Nk, ik + iq - 1))
    base = np.where(inside, 4 * q ** 2 * kk ** 2 - (q ** 2 + kk ** 2 - p ** 2) ** 2, 1.0)
    J = np.where(inside, kk * p * base ** ((d - 3) / 2) / q ** d, 0.0)
    V = np.where(inside, ((q ** 2 + kk ** 2 - p ** 2) * c_k + (q ** 2 - kk ** 2 + p ** 2) * c_p) ** 2, 0.0)
    return prefactor, np.asfortranarray(V), np.asfortranarray(J)


def dDimModeCouplingKernel(rho, kBT, m, k_array, Sk, d, device=0):
    """dDimModeCouplingKernel(ρ, kBT, m, k_array, Sₖ, d) -- src/Kernels/ModeCouplingKernels.jl:34-71.  The O(Nk^3)
    tables are built on the host like in the Julia constructor and shipped once."""
    L = load_library()
    k = _arr(k_array)
    prefactor, V, J = ddim_tables(rho, kBT, m, k, Sk, d)
    h = C.c_void_p()
    _check(L.mct_ddim_create(C.byref(h), device, len(k), _p(k), C.c_double(prefactor), _p(V), _p(J)))
    return MemoryKernel(h, len(k), 1, device)


def active_tables(rho, k_array, wk, w0, Sk, dim=3):
    """prefactor, J[l, j, i], V2[l, j, i] of ActiveMCTKernel -- src/Kernels/ActiveMCTKernel.jl:28-49 (i: output wavenumber
    k, j: q, l: p; zero outside abs(j-i)+1 <= l <= min(i+j-1, Nk))."""
    k = _arr(k_array); Sk = _arr(Sk); wk = _arr(wk)
    Nk = len(k)
    dk = k[1] - k[0]
    prefactor = dk ** 2 * rho / (2 * (2 * np.pi) ** dim) * _surface_d_dim_unit_sphere(dim - 1)
    mCk = 1 / rho * (1 - wk / (w0 * Sk))
    il, ij, ii = np.meshgrid(np.arange(1, Nk + 1), np.arange(1, Nk + 1), np.arange(1, Nk + 1), indexing="ij")
    kk, q, p = k[ii - 1], k[ij - 1], k[il - 1]
    cq, cp = mCk[ij - 1], mCk[il - 1]
    inside = (np.abs(ij - ii) + 1 <= il) & (il <= np.minimum(ii + ij - 1, Nk))
    V = cq / (2 * kk) * (kk ** 2 + q ** 2 - p ** 2) + cp / (2 * kk) * (kk ** 2 + p ** 2 - q ** 2)
    V2 = np.where(inside, wk[ii - 1] * V ** 2, 0.0)
    base = np.where(inside, (q + p - kk) * (kk + p - q) * (kk + q - p) * (kk + p + q), 1.0)
    J = np.where(inside, 2 * p * q / ((2 * kk) ** (dim - 2)) * base ** ((dim - 3) / 2), 0.0)
    return prefactor, np.asfortranarray(J), np.asfortranarray(V2)


def ActiveMCTKernel(rho, k_array, wk, w0, Sk, dim=3, device=0):
    """ActiveMCTKernel(ρ, k_array, wk, w0, Sk, dim=3) -- src/Kernels/ActiveMCTKernel.jl:28-49.  The tables are built on the
    host like in the Julia constructor; the evaluation (:51-61) is the device's dense triple sum."""
    L = load_library()
    k = _arr(k_array)
    prefactor, J, V2 = active_tables(rho, k, wk, w0, Sk, dim)
    h = C.c_void_p()
    _check(L.mct_active_create(C.byref(h), device, len(k), _p(k), C.c_double(prefactor), _p(J), _p(V2)))
    return MemoryKernel(h, len(k), 1, device)


def tagged_active_tables(rho, k_array, wk, w0, Sk, dim=3):
    """prefactor, J[l, j, i], V2[l, j, i] of TaggedActiveMCTKernel -- src/Kernels/ActiveMCTKernel.jl:96-117."""
    k = _arr(k_array); Sk = _arr(Sk); wk = _arr(wk)
    Nk = len(k)
    dk = k[1] - k[0]
    prefactor = rho * dk ** 2 / ((2 * np.pi) ** dim) * _surface_d_dim_unit_sphere(dim - 1)
    mCk = 1 / rho * (1 - wk / (w0 * Sk))
    il, ij, ii = np.meshgrid(np.arange(1, Nk + 1), np.arange(1, Nk + 1), np.arange(1, Nk + 1), indexing="ij")
    kk, q, p = k[ii - 1], k[ij - 1], k[il - 1]
    inside = (np.abs(ij - ii) + 1 <= il) & (il <= np.minimum(ii + ij - 1, Nk))
    V2 = np.where(inside, (mCk[il - 1] / (2 * kk) * (kk ** 2 + p ** 2 - q ** 2)) ** 2 * w0, 0.0)
    base = np.where(inside, (q + p - kk) * (kk + p - q) * (kk + q - p) * (kk + p + q), 1.0)
    J = np.where(inside, 2 * p * q / ((2 * kk) ** (dim - 2)) * base ** ((dim - 3) / 2), 0.0)
    return prefactor, np.asfortranarray(J), np.asfortranarray(V2)


def TaggedActiveMCTKernel(rho, k_array, wk, w0, Sk, Fsol, dim=3):
    """TaggedActiveMCTKernel(ρ, k_array, wk, w0, Sk, Fsol, dim) -- src/Kernels/ActiveMCTKernel.jl:96-117; `Fsol` is the
    device-resident collective solution (the tDict lookup becomes the output index)."""
    if getattr(Fsol, "_equation", None) is None:
        raise MCTDeviceError("TaggedActiveMCTKernel needs the device-resident collective solution")
    L = load_library()
    k = _arr(k_array)
    prefactor, J, V2 = tagged_active_tables(rho, k, wk, w0, Sk, dim)
    h = C.c_void_p()
    _check(L.mct_active_tagged_create_from_equation(C.byref(h), len(k), _p(k), C.c_double(prefactor), _p(J), _p(V2),
                                                    Fsol._equation._q))
    return MemoryKernel(h, len(k), 1, Fsol._equation._kernel.device, keep=(Fsol._equation,))


def evaluate_kernel(kernel, F, t=0):
    """out = evaluate_kernel(kernel, F, t) -- src/Kernels.jl:18-26.  For tagged kernels `t` is the index of the time
    in the collective solution's `t` array."""
    out = np.empty(kernel.dof)
    evaluate_kernel_(out, kernel, F, t)
    return out


def evaluate_kernel_(out, kernel, F, t=0):
    """evaluate_kernel!(out, kernel, F, t) -- src/Kernels.jl:10-16."""
    F = _arr(_to_blocks(F) if kernel.Ns > 1 else F, (kernel.dof,))
    assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == kernel.dof
    _check(load_library().mct_kernel_eval(kernel._h, _p(F), C.c_long(int(t)), _p(out)))
    return out


def evaluate_kernel_many(kernel, F):
    F = _arr(F, (-1, kernel.dof))
    out = np.empty_like(F)
    ms = C.c_float(0)
    _check(load_library().mct_kernel_eval_many(kernel._h, F.shape[0], _p(F), _p(out), C.byref(ms)))
    return out, ms.value


# --------------------------------------------------------------------------------------------- equation + solver
class MemoryEquation:
    """MemoryEquation(α, β, γ, δ, F₀, ∂ₜF₀, kernel) -- src/MemoryEquation.jl:44-52.

    α F̈ + β Ḟ + γ F + δ + ∫ K(τ) Ḟ(t-τ) dτ = 0.  Scalars become UniformScaling, vectors become Diagonal
    (src/HelperFunctions.jl:73-84).  A custom `update_coefficients!` is not supported on the device."""

    def __init__(self, alpha, beta, gamma, delta, F0, dF0, kernel, update_coefficients=None):
        if update_coefficients is not None:
            raise NotImplementedError("update_coefficients! callbacks are not supported by the device solver")
        self.kernel = kernel
        blk = _to_blocks if kernel.Ns > 1 else (lambda v: v)
        self.F0 = _arr(blk(F0), (kernel.dof,))
        self.dF0 = _arr(blk(dF0), (kernel.dof,))
        self.coeffs = {}
        for name, v in (("alpha", alpha), ("beta", beta), ("gamma", gamma)):
            self.coeffs[name] = float(v) if np.isscalar(v) else _arr(blk(v), (kernel.dof,))
        self.delta = None if (np.isscalar(delta) and float(delta) == 0.0) else (
            np.full(kernel.dof, float(delta)) if np.isscalar(delta) else _arr(blk(delta), (kernel.dof,)))

    def _c_coeffs(self):
        c = Coeffs()
        for name in ("alpha", "beta", "gamma"):
            v = self.coeffs[name]
            if isinstance(v, float):
                setattr(c, name + "_kind", 0); setattr(c, name + "_scalar", v); setattr(c, name, None)
            else:
                setattr(c, name + "_kind", 1); setattr(c, name + "_scalar", 0.0); setattr(c, name, _p(v))
        c.delta = _p(self.delta) if self.delta is not None else None
        return c


class TimeDoublingSolver:
    """TimeDoublingSolver(; N=32, Δt=1e-10, t_max=1e10, max_iterations=10^4, tolerance=1e-10, verbose=false,
    ismutable=true, init_with_memory=true) -- src/TimeDoublingSolver.jl:48-50 (Δt is spelled `dt`)."""

    def __init__(self, N=32, dt=1e-10, t_max=1e10, max_iterations=10 ** 4, tolerance=1e-10, verbose=False, ismutable=True,
                 init_with_memory=True):
        if not ismutable:
            raise NotImplementedError("ismutable=false is not supported by the device solver")
        self.N, self.dt, self.t_max = int(N), float(dt), float(t_max)
        self.max_iterations, self.tolerance = int(max_iterations), float(tolerance)
        self.verbose, self.init_with_memory = bool(verbose), bool(init_with_memory)
        self.kernel_evals = 0
        self.device_ms = None

    def _c_opts(self):
        return TdOptions(self.N, self.dt, self.t_max, self.max_iterations, self.tolerance, int(self.verbose),
                         int(self.init_with_memory))

    def output_len(self):
        nt = C.c_long(0)
        o = self._c_opts()
        _check(load_library().mct_td_output_len(C.byref(o), C.byref(nt)))
        return nt.value


class _DeviceEquation:
    def __init__(self, q, kernel=None):
        self._q = q
        self._kernel = kernel      # the equation handle refers to the kernel handle: keep it alive as long as the equation

    def close(self):
        """mct_equation_destroy.  Refused while a tagged kernel chained on this solution is alive."""
        if self._q is not None and _lib is not None:
            if _lib.mct_equation_destroy(self._q) != MCT_OK:
                raise MCTDeviceError(_lib.mct_last_error().decode())
            self._q = None
        self._kernel = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MemoryEquationSolution:
    """MemoryEquationSolution(t, F, K, solver) -- src/Solvers.jl:19-24.  F and K are [nt, dof] arrays: row `it` is the
    reference's `sol.F[it]` (and the diagonal of `sol.K[it]`)."""

    def __init__(self, t, F, K, solver, equation=None):
        self.t, self.F, self.K, self.solver = t, F, K, solver
        self._equation = equation

    def __getitem__(self, ik):   # getindex(sol, I...) = getindex.(sol.F, I...)  src/Solvers.jl:50
        return self.F[:, ik]


def get_t(sol):
    return sol.t


def get_F(sol, *idx):
    return sol.F[idx] if idx else sol.F


def get_K(sol, *idx):
    return sol.K[idx] if idx else sol.K


def _create_equation(equation, solver):
    L = load_library()
    q = C.c_void_p()
    c = equation._c_coeffs()
    o = solver._c_opts()
    _check(L.mct_equation_create(C.byref(q), equation.kernel._h, C.byref(c), _p(equation.F0), _p(equation.dF0), C.byref(o)))
    return _DeviceEquation(q, equation.kernel)


def _download(dev_eq, solver, dof, keep_device):
    L = load_library()
    nt = solver.output_len()
    t = np.empty(nt); F = np.empty((nt, dof)); K = np.empty((nt, dof))
    _check(L.mct_equation_download(dev_eq._q, _p(t), _p(F), _p(K)))
    return MemoryEquationSolution(t, F, K, solver, dev_eq if keep_device else None)


def solve(equation, solver=None, keep_device=True):
    """solve(equation, solver) -- src/Solvers.jl:42-46, src/TimeDoublingSolver.jl:512-532.
    Writes solver.kernel_evals like the reference (:416, :520)."""
    solver = solver or TimeDoublingSolver()
    if isinstance(equation.kernel, MSDKernel):
        return _solve_msd(equation, solver)
    L = load_library()
    dev_eq = _create_equation(equation, solver)
    evals = C.c_long(0)
    ms = C.c_float(0)
    rc = L.mct_equation_solve(dev_eq._q, C.byref(evals), C.byref(ms))
    solver.kernel_evals = evals.value
    solver.device_ms = ms.value
    _check(rc)
    sol = _download(dev_eq, solver, equation.kernel.dof, keep_device)
    if not keep_device:
        dev_eq.close()
    return sol


def create_equation(equation, solver):
    """Upload an equation once (coefficients, initial conditions; allocates rings + trajectory on the device)."""
    return _create_equation(equation, solver)


def solve_resident(dev_eq):
    """Run the persistent kernel on an equation that is already resident (no host<->device traffic).
    Returns (kernel_evals, device_ms)."""
    evals = C.c_long(0)
    ms = C.c_float(0)
    _check(load_library().mct_equation_solve(dev_eq._q, C.byref(evals), C.byref(ms)))
    return evals.value, ms.value


def shard_equation(dev_eq, rank, world, all_gather):
    """Shard a resident multi-component equation over `world` GPUs (one process per GPU).  `all_gather(bytes) -> list of
    bytes` is the host transport for the 64-byte CUDA IPC handles (e.g. torch.distributed.all_gather_object)."""
    L = load_library()
    mine = (C.c_ubyte * 64)()
    _check(L.mct_equation_shard_prepare(dev_eq._q, rank, world, mine))
    handles = all_gather(bytes(mine))
    assert len(handles) == world and all(len(h) == 64 for h in handles)
    buf = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
    _check(L.mct_equation_shard_connect(dev_eq._q, buf))


def download_solution(dev_eq, solver, dof):
    return _download(dev_eq, solver, dof, False)


def download_slice(dev_eq, dof, times=None, dofs=None, want_K=True):
    """get_F(sol, its, iks) / get_K on a device-resident solution (src/HelperFunctions.jl:169-210): only the requested
    [times, dofs] slice crosses PCIe.  Returns (t, F, K) with F, K of shape [len(times), len(dofs)]."""
    L = load_library()
    nt = C.c_long(0)
    ti = None if times is None else np.ascontiguousarray(times, dtype=np.int64)
    di = None if dofs is None else np.ascontiguousarray(dofs, dtype=np.int32)
    if ti is None:
        raise ValueError("pass the time indices (use download_solution for whole trajectories)")
    n_t, n_d = len(ti), (dof if di is None else len(di))
    t = np.empty(n_t); F = np.empty((n_t, n_d)); K = np.empty((n_t, n_d)) if want_K else None
    _check(L.mct_equation_download_slice(dev_eq._q, C.c_long(n_t), ti.ctypes.data_as(C.POINTER(C.c_long)), n_d,
                                         None if di is None else di.ctypes.data_as(C.POINTER(C.c_int)),
                                         _p(t), _p(F), None if K is None else _p(K)))
    return t, F, K


def relaxation_times(dev_eq, ik, threshold=np.exp(-1), mode="log"):
    """find_relaxation_time(t, F(k_ik); threshold, mode) for the wavenumbers `ik` of a device-resident solution."""
    if mode not in ("log", "lin"):
        raise ValueError(f"the mode ':{mode}' is not know. it must be either :lin, :log.")
    L = load_library()
    idx = np.ascontiguousarray(np.atleast_1d(ik), dtype=np.int32)
    tau = np.empty(len(idx))
    _check(L.mct_equation_relaxation_times(dev_eq._q, len(idx), idx.ctypes.data_as(C.POINTER(C.c_int)), C.c_double(threshold),
                                           0 if mode == "log" else 1, _p(tau)))
    return tau


def solve_into(equation, solver, t, F, K):
    """solve(equation, solver) through the one-call C entry point `mct_td_solve` with caller-owned HOST buffers
    (t[nt], F[nt, dof], K[nt, dof]) -- what the Julia glue ccalls."""
    c = equation._c_coeffs()
    o = solver._c_opts()
    evals = C.c_long(0)
    rc = load_library().mct_td_solve(equation.kernel._h, C.byref(c), _p(equation.F0), _p(equation.dF0), C.byref(o), _p(t), _p(F),
                                     _p(K), C.byref(evals))
    solver.kernel_evals = evals.value
    _check(rc)
    return MemoryEquationSolution(t, F, K, solver)


def solve_batch(equations, solver=None, download=True):
    """Many independent equations with identical (Nk, solver options) in ONE persistent launch
    (e.g. a packing-fraction sweep).  Returns (solutions, statuses, kernel_evals, device_ms)."""
    solver = solver or TimeDoublingSolver()
    L = load_library()
    devs = [_create_equation(e, solver) for e in equations]
    n = len(devs)
    arr = (C.c_void_p * n)(*[d._q for d in devs])
    status = (C.c_int * n)()
    evals = (C.c_long * n)()
    ms = C.c_float(0)
    rc = L.mct_equation_solve_batch(n, arr, status, evals, C.byref(ms))
    if rc not in (MCT_OK, MCT_NOT_CONVERGED):
        _check(rc)
    sols = []
    if isinstance(download, dict):            # {"times": [...], "dofs": [...]}: only that slice of every solution comes back
        for d, e in zip(devs, equations):
            t, F, K = download_slice(d, e.kernel.dof, download.get("times"), download.get("dofs"), download.get("K", True))
            sols.append(MemoryEquationSolution(t, F, K, solver))
    elif download:
        for d, e in zip(devs, equations):
            sols.append(_download(d, solver, e.kernel.dof, False))
    for d in devs:
        d.close()
    solver.kernel_evals = int(sum(evals))
    solver.device_ms = ms.value
    return sols, list(status), list(evals), ms.value


def solve_steady_state(gamma, F0, kernel, tolerance=1e-8, max_iterations=10 ** 6, verbose=False):
    """solve_steady_state(γ, F₀, kernel; tolerance, max_iterations) -- src/SteadyStateMemoryEquation.jl:49-100."""
    blk = _to_blocks if kernel.Ns > 1 else (lambda v: v)
    g = _arr(blk(gamma), (kernel.dof,)) if not np.isscalar(gamma) else np.full(kernel.dof, float(gamma))
    F0 = _arr(blk(F0), (kernel.dof,))
    F = np.empty(kernel.dof); K = np.empty(kernel.dof)
    it = C.c_long(0)
    rc = load_library().mct_steady_state(kernel._h, _p(g), _p(F0), C.c_double(tolerance), C.c_long(max_iterations), _p(F), _p(K),
                                         C.byref(it))
    if rc == MCT_NOT_CONVERGED:
        raise RuntimeError(load_library().mct_last_error().decode())   # `error(...)` :76-78
    _check(rc)
    sol = MemoryEquationSolution(np.array([np.inf]), F[None, :], K[None, :],
                                 dict(tolerance=tolerance, max_iterations=max_iterations, verbose=verbose))
    sol.iterations = it.value
    return sol


def find_relaxation_time(t, F, threshold=np.exp(-1), mode="log"):
    """find_relaxation_time(t, F; threshold=exp(-1), mode=:log) -- src/RelaxationTime.jl:14-39 (O(nt) host
    post-processing of a finished solution, kept on the host like in the reference)."""
    t = np.asarray(t, dtype=np.float64)
    F = np.asarray(F, dtype=np.float64)
    for it in range(len(t)):
        if F[it] / F[0] > threshold:
            continue
        if it == 0:
            return 0.0
        y0, y1, y = F[it] / F[0], F[it - 1] / F[0], threshold
        if mode == "log":
            x0, x1 = np.log(t[it]), np.log(t[it - 1])
            return float(np.exp(((y - y0) * (x1 - x0) + x0 * (y1 - y0)) / (y1 - y0)))
        if mode == "lin":
            x0, x1 = t[it], t[it - 1]
            return float(((y - y0) * (x1 - x0) + x0 * (y1 - y0)) / (y1 - y0))
        raise ValueError(f"the mode ':{mode}' is not know. it must be either :lin, :log.")
    return float("inf")
