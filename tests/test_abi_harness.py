"""The C ABI driven from plain C (tests/abi_harness.c) -- the executable stand-in for julia/MCTDevice.jl, whose ccalls use the
same entry points and struct layouts.  CPU part: the harness builds with gcc against include/mct.h, links the library, and
its struct layouts are the ones the Julia `struct`s declare.  GPU part: create -> mct_td_solve -> mct_kernel_eval -> destroy
from C gives bit for bit what the Python mirror gives, and matches the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

import problems as P
from conftest import ROOT, load_pkg

pkg = load_pkg()
HERE = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.path.dirname(pkg.api.library_path())
EXE = os.path.join(HERE, "_build", "abi_harness")


def build():
    pkg.load_library()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-o", EXE, os.path.join(HERE, "abi_harness.c"), "-L" + LIBDIR,
                           "-lmct_sm100a", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath,/usr/local/cuda/lib64", "-Wl,--allow-shlib-undefined"])
    return EXE


def env():
    e = dict(os.environ)
    e["LD_LIBRARY_PATH"] = "/usr/local/cuda/lib64:" + e.get("LD_LIBRARY_PATH", "")
    return e


def julia_struct_layout(name):
    """Field offsets of an isbits Julia struct with C layout, from the declaration in julia/MCTDevice.jl."""
    src = open(os.path.join(ROOT, "julia", "MCTDevice.jl")).read()
    body = re.search(r"struct " + name + r"\b(.*?)\nend", src, re.S).group(1)
    size = {"Cint": 4, "Cdouble": 8, "Clong": 8, "Ptr{Cdouble}": 8}
    off, out, align = 0, {}, 1
    for fname, ftype in re.findall(r"(\w+)::([\w{}]+)", body):
        s = size[ftype]
        off = (off + s - 1) // s * s
        out[fname] = off
        off += s
        align = max(align, s)
    return out, (off + align - 1) // align * align


def test_harness_builds_links_and_struct_layouts_match_the_julia_declaration():
    exe = build()
    out = subprocess.check_output([exe, "--layout"], env=env(), text=True)
    for line, jl in ((out.splitlines()[0], "MctTdOptions"), (out.splitlines()[1], "MctCoeffs")):
        tok = line.split()
        c_size, c_off = int(tok[1]), {tok[i]: int(tok[i + 1]) for i in range(2, len(tok), 2)}
        j_off, j_size = julia_struct_layout(jl)
        assert c_size == j_size, (jl, c_size, j_size)
        assert c_off == j_off, (jl, c_off, j_off)


@pytest.mark.gpu
def test_c_harness_solve_equals_python_mirror_and_oracle(tmp_path):
    import oracle as O
    exe = build()
    Nk, N, dt, t_max, tol = 64, 8, 1e-6, 1e2, 1e-10
    p = P.hard_spheres(Nk, 0.50)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(np.array([Nk, N], dtype=np.int32).tobytes()); f.write(np.array([p["rho"], dt, t_max, tol]).tobytes())
        f.write(p["k"].tobytes()); f.write(p["Sk"].tobytes())
    print(subprocess.check_output([exe, str(fin), str(fout)], env=env(), text=True).strip())
    raw = open(fout, "rb").read()
    nt, evals = np.frombuffer(raw[:16], dtype=np.int64)
    body = np.frombuffer(raw[16:], dtype=np.float64)
    t, F, K, K0 = body[:nt], body[nt:nt + nt * Nk].reshape(nt, Nk), body[nt + nt * Nk:nt + 2 * nt * Nk].reshape(nt, Nk), body[nt + 2 * nt * Nk:]
    e = P.overdamped(p)
    kern = pkg.ModeCouplingKernel(p["rho"], 1.0, 1.0, p["k"], p["Sk"])
    solver = pkg.TimeDoublingSolver(N=N, dt=dt, t_max=t_max, tolerance=tol, max_iterations=10 ** 6)
    sol = pkg.solve(pkg.MemoryEquation(e["alpha"], e["beta"], e["gamma"], e["delta"], e["F0"], e["dF0"], kern), solver)
    assert evals == solver.kernel_evals and np.array_equal(t, sol.t)
    assert np.array_equal(F, sol.F) and np.array_equal(K, sol.K)                      # same library, same calls: bit for bit
    ref = O.td_solve(O.sc3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"]), 0.0, 1.0, e["gamma"], 0.0, e["F0"], e["dF0"], N=N, dt=dt, t_max=t_max,
                     tolerance=tol, max_iterations=10 ** 6)
    assert np.max(np.abs(F - ref["F"])) <= 1e-10 and evals == ref["kernel_evals"]
    assert np.allclose(K0, ref["K"][0], rtol=1e-12, atol=0)
