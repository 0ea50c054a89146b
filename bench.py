#!/usr/bin/env python
"""Benchmark of the B200 hot path of ModeCouplingTheory.jl (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): kernel evaluations per second of the hard-sphere MCT solve at Nk=1000 to t=1e10
(config 3: ModeCouplingKernel, Percus-Yevick S(k), eta=0.5155, overdamped, TimeDoublingSolver defaults).
One "step" = one complete solve(equation, TimeDoublingSolver()) = one launch of the persistent kernel.

  value      evals/s with the equation already resident in HBM (coefficients, F0, vertex tables); device time of
             the launch measured with CUDA events on the library's own stream, max over ranks.
  e2e        the same solve through the host-facing call with HOST buffers (mct_td_solve semantics: upload
             coefficients, solve, download t, F, K): wall clock around the call, copies inside.
  roofline   algorithmic table bytes (24 Nk^2 + 16 Nk per evaluation) / launch duration against the
             measured HBM copy peak.  The tables never leave the chip during a solve (registers + shared memory),
             which is why the figure can exceed the HBM peak; DESIGN.md discusses what really bounds the kernel.
  cpu_baseline   the oracle (C restatement of the reference in its operation order, 1 thread: the reference hot path
             is single-threaded) on a bounded prefix of the same solve.  The reference itself is Julia and cannot
             run here.

N > 1: the contract line shards over independent equations only -- every rank solves its own equation of
the same configuration (weak scaling, no data-path collective); NCCL is used for the barrier and the max over ranks.

Besides the contract keys the line carries (all measured by this run, at every --gpus N):
  parity_checked / parity   the timed host-buffer solve compared with the oracle's solution of the same equation
             (tests/golden/c3_nk1000_eta0.5155.npz: F and K at 11 wavenumbers x all times and 12 times x all wavenumbers,
             kernel_evals): max|dF| <= 1e-10 or the run fails.
  sweep      BASELINE config 5 -- the 512-point packing-fraction sweep at Nk=500 sharded over the N ranks (strong scaling:
             solves/s, wall, device ms per rank, kernel evaluations per rank, located eta_c).   --no-sweep skips it.
  kshard     BASELINE config 4 -- the Ns=4, Nk=2000 mixture: evaluations/s of one equation k-sharded over the N ranks and,
             for N >= 2, bitwise equality of the sharded solution with the single-GPU one.       --no-kshard skips it.
  configs    config 1 (F2 schematic, scalar, CPU: reported only) and config 2 (Nk=100 on the device, microseconds/eval).
  roofline_fp64, roofline.bound_actual   what the kernel is really bound by (it does not touch HBM during a solve).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NK = 1000
ETA = 0.5155
WORKLOAD = ("config3: hard-sphere MCT, Percus-Yevick S(k), Nk=1000, eta=0.5155 (near eta_c), overdamped, "
            "TimeDoublingSolver(N=32, dt=1e-10, t_max=1e10, tol=1e-10)")
NCU_TRAFFIC_BYTES_PER_LAUNCH = 30113536 + 25608448      # dram__bytes_read.sum + dram__bytes_write.sum of one launch
NCU_TRAFFIC_SOURCE = "profiles/ncu_full_r02_td_sc_kernel.csv"
METRIC = "hs_mct_nk1000_kernel_evals_per_s"
UNIT = "kernel evals/s"


def problem(pkg_sf, eta=ETA, Nk=NK):
    k = pkg_sf.wavenumber_grid(Nk)
    Sk = pkg_sf.percus_yevick_sk(k, eta)
    rho = 6 * eta / np.pi
    return dict(Nk=Nk, k=k, Sk=Sk, rho=rho, gamma=k ** 2 / Sk, F0=Sk.copy(), dF0=np.zeros(Nk))


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (profiling recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, f"/tmp/mct_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _oracle_problem():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import importlib.util
    spec = importlib.util.spec_from_file_location("mct_sf", os.path.join(ROOT, "modecouplingtheory.jl_b200", "structure_factors.py"))
    sf = importlib.util.module_from_spec(spec); spec.loader.exec_module(sf)
    p = problem(sf)
    return O, p, O.sc3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"])


_BOOT = {}


def cpu_baseline_sample(n_evals=200):
    """The oracle on a bounded prefix of the SAME solve, 1 thread: `n_evals` fixed-point kernel evaluations of the main loop
    (what 53 197 of the solve's 53 262 evaluations are), timed as (prefix with n_evals) - (prefix with 1), i.e. without the
    2N-step Euler bootstrap, whose O(step) convolution would dominate a short prefix; the bootstrap is reported beside it."""
    O, p, kern = _oracle_problem()
    solve = lambda n: O.td_solve(kern, 0.0, 1.0, p["gamma"], 0.0, p["F0"], p["dF0"], max_iterations=10 ** 6, max_evals=n)
    if "s" not in _BOOT:                      # bootstrap (65 evaluations + Euler convolutions) + 1 evaluation: timed once
        t0 = time.perf_counter(); solve(1); _BOOT["s"] = time.perf_counter() - t0
    t0 = time.perf_counter(); res = solve(1 + n_evals); wall = time.perf_counter() - t0
    done = res["kernel_evals"] - 1
    rate = done / max(wall - _BOOT["s"], 1e-9)
    return rate, (f"{done} fixed-point kernel evaluations of the same Nk=1000 solve in {wall - _BOOT['s']:.1f} s "
                  f"(bootstrap of 65 evaluations timed separately: {_BOOT['s']:.1f} s; whole solve extrapolated: "
                  f"{53197 / rate + _BOOT['s']:.0f} s)")


def config1_f2_on_cpu():
    """BASELINE config 1 (README.md:30-47): SchematicF2Kernel(3.999), scalar MemoryEquation, TimeDoublingSolver defaults --
    stays on the CPU, reported for completeness: oracle wall time and evaluations."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    t0 = time.perf_counter()
    r = O.td_solve(O.f2(3.999), 1.0, 0.0, 1.0, 0.0, np.array([1.0]), np.array([0.0]), t_max=1e10, scalar_mode=True)
    wall = time.perf_counter() - t0
    return {"workload": "config1: SchematicF2Kernel(nu=3.999) scalar, TimeDoublingSolver defaults, CPU oracle (reported, not accelerated)",
            "wall_ms": 1e3 * wall, "kernel_evals": int(r["kernel_evals"]), "F_end": float(r["F"][-1, 0])}


def run_reference(args, rank):
    if rank != 0:
        return
    vals, sample = [], ""
    for i in range(args.warmup + args.steps):       # every step: one bounded sample (200 fixed-point evaluations, ~13 s)
        v, sample = cpu_baseline_sample(200)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * 200.0 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic (analytic Percus-Yevick hard-sphere S(k))",
            "config": {"workload": WORKLOAD},
            "details": {"note": "reference = ModeCouplingTheory.jl (Julia, not installable here: no julia binary, no network); timed "
                                "instead: oracle/mct_oracle.c, the C restatement in the reference's operation order"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample +
                             "; 1 thread because the reference hot path is single-threaded (@turbo, no @threads)",
                             "spread": {"min": float(np.min(vals)), "max": float(np.max(vals)), "samples": len(vals)}},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def sweep_record(args, pkg, torch, dist, rank, world, dev):
    """Config 5: `points` packing fractions in [0.50, 0.53] at Nk = sweep_nk, sharded over the ranks (strong scaling).
    eta_c is located on the grid (steady-state bisection on the device), the measured cost model deals the points to ranks
    longest-first and sorts every rank's queue, outliers get a whole-GPU team (sweep.py); no data-path collective, one final
    gather of the observable.  The wall clock includes everything: eta_c location, table construction, uploads, solves,
    downloads of the observable, the gather."""
    sw = pkg.sweep
    etas = np.linspace(0.50, 0.53, args.points)
    kpeak = int(np.argmin(np.abs(pkg.structure_factors.wavenumber_grid(args.sweep_nk) - 7.0)))
    nt = pkg.TimeDoublingSolver().output_len()
    obs = lambda sol: np.array([sol.F[-1, 0] / sol.F[0, 0]])
    slice_ = dict(times=[0, nt - 1], dofs=[kpeak], K=False)        # two numbers per solution cross PCIe
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    eta_c = sw.locate_eta_c(pkg, args.sweep_nk, dev)
    t_loc = time.perf_counter() - t0
    cost = sw.expected_cost(etas, eta_c)
    big = sw.outliers(cost, world)
    mine = sw.partition(sw.weighted_costs(cost, world), world, rank)
    pos_big = [mine.index(i) for i in big if i in mine]
    vals, evals_pt, ms = sw.run_sweep(pkg, etas[mine], args.sweep_nk, dev, observable=obs, download=slice_, costs=cost[mine], big_team=pos_big)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    per_rank = torch.zeros(world, 3, dtype=torch.float64, device=f"cuda:{dev}")
    per_rank[rank, 0], per_rank[rank, 1], per_rank[rank, 2] = wall, ms, float(evals_pt.sum())
    if world > 1:
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
    allv = sw.gather(mine, vals, len(etas), world, dist if world > 1 else None)
    wall_all = time.perf_counter() - t0
    w = torch.tensor([wall_all], dtype=torch.float64, device=f"cuda:{dev}")
    if world > 1:
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
    pr = per_rank.cpu().numpy()
    nonerg = allv[:, 0] > 0.1
    return {"workload": f"config5: {args.points}-point packing-fraction sweep eta in [0.50, 0.53], Nk={args.sweep_nk}, overdamped, "
                        "TimeDoublingSolver defaults (max_iterations 1e6)",
            "solves_per_s": args.points / w.item(), "wall_s": w.item(), "n_gpus": world, "scaling": "strong",
            "device_ms_per_rank": [round(x, 1) for x in pr[:, 1].tolist()], "wall_s_per_rank": [round(x, 3) for x in pr[:, 0].tolist()],
            "kernel_evals_per_rank": [int(x) for x in pr[:, 2].tolist()], "kernel_evals_total": int(pr[:, 2].sum()),
            "device_ms_max_over_min": float(pr[:, 1].max() / max(pr[:, 1].min(), 1e-9)),
            "eta_c_located_by_steady_state": eta_c, "eta_c_location_s": t_loc, "host_timing_rank0": {k: round(v, 3) for k, v in sw.run_sweep.last_timing.items()},
            "first_nonergodic_point": float(etas[np.argmax(nonerg)]) if nonerg.any() else None,
            "whole_gpu_team_points": [float(etas[i]) for i in big],
            "sharding": "longest-first onto the least loaded rank by the measured cost model, sorted queue per rank, no data-path "
                        "collective, final gather of one observable"}


def kshard_record(args, pkg, torch, dist, rank, world, dev):
    """Config 4: ONE multi-component equation (Ns=4 hard-sphere mixture, Nk=2000) on a shortened time axis.  N = 1: evaluations/s
    of the single-GPU solve.  N >= 2: the same equation k-sharded over the ranks (rank r evaluates its share of the lines of
    the vertex sum, the sums travel by device-initiated NVLink loads from CUDA-IPC-mapped peer buffers, everything else is
    replicated) -- evaluations/s, and the sharded trajectory compared BITWISE with the single-GPU trajectory of the same
    library on every rank."""
    sf = pkg.structure_factors
    ns, Nk = 4, args.mc_nk
    diam = np.array([0.7, 0.8, 0.9, 1.0]); x = np.full(ns, 0.25)
    rho = 6 * 0.50 / (np.pi * np.sum(x * diam ** 3)) * x
    m = np.ones(ns)
    k = sf.wavenumber_grid(Nk)
    Sk = sf.percus_yevick_mixture_sk(k, diam, rho)
    xx = rho / rho.sum()
    gamma = np.einsum("kab,kbc->kac", np.array([np.diag(k[i] ** 2 * xx / m) for i in range(Nk)]), np.linalg.inv(Sk))
    kern = pkg.MultiComponentModeCouplingKernel(rho, 1.0, m, k, Sk, device=dev)
    eq = pkg.MemoryEquation(1.0, 0.0, gamma, np.zeros((Nk, ns, ns)), Sk, np.zeros((Nk, ns, ns)), kern)
    solver = pkg.TimeDoublingSolver(N=4, dt=1e-4, t_max=args.mc_tmax, tolerance=1e-10, max_iterations=10 ** 6)

    def timed(dev_eq):
        ms_all = []
        for _ in range(args.mc_warmup + args.mc_steps):      # every pass is the whole solve again from F0
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            evals, ms = pkg.api.solve_resident(dev_eq)
            ms_all.append(ms)
        t = torch.tensor([float(np.mean(ms_all[args.mc_warmup:]))], dtype=torch.float64, device=f"cuda:{dev}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return evals + 2 * solver.N + 1, t.item()

    single = pkg.api.create_equation(eq, solver)
    evals1, ms1 = timed(single)
    F1 = pkg.api.download_solution(single, solver, kern.dof).F
    single.close()
    rec = {"workload": f"config4: MultiComponentModeCouplingKernel Ns=4 (diameters 0.7/0.8/0.9/1.0, equal fractions, phi=0.50), Nk={Nk}, "
                       f"Newtonian, TimeDoublingSolver(N=4, dt=1e-4, t_max={args.mc_tmax}, tol=1e-10)",
           "n_gpus": world, "scaling": "strong", "kernel_evals": int(evals1),
           "single_gpu_evals_per_s": evals1 / (ms1 * 1e-3), "single_gpu_ms": ms1,
           "fp64_ginstr_per_s_single_gpu": 12.0 * ns * ns * Nk * Nk * 0.75 * evals1 / (ms1 * 1e-3) / 1e9, "fp64_peak_ginstr_per_s": 17700.0}
    if world > 1:
        def all_gather(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        sharded = pkg.api.create_equation(eq, solver)
        pkg.api.shard_equation(sharded, rank, world, all_gather)
        evalsN, msN = timed(sharded)
        FN = pkg.api.download_solution(sharded, solver, kern.dof).F
        same = torch.tensor([1.0 if (evalsN == evals1 and np.array_equal(FN, F1)) else 0.0], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        sharded.close()
        # the collective the north star names for this step, for comparison: NCCL all-gather of F(k) (dof doubles in total) plus
        # a MAX all-reduce of the 8-byte error word, launched from the host once per fixed-point iteration (what a solver that
        # is NOT resident on the device would pay per evaluation besides its kernels); CUDA events, max over ranks
        dof = kern.dof
        shard = torch.zeros((dof + world - 1) // world, dtype=torch.float64, device=f"cuda:{dev}")
        full = torch.zeros(shard.numel() * world, dtype=torch.float64, device=f"cuda:{dev}")
        err = torch.zeros(1, dtype=torch.float64, device=f"cuda:{dev}")
        for _ in range(20):
            dist.all_gather_into_tensor(full, shard); dist.all_reduce(err, op=dist.ReduceOp.MAX)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 200
        e0.record()
        for _ in range(reps):
            dist.all_gather_into_tensor(full, shard); dist.all_reduce(err, op=dist.ReduceOp.MAX)
        e1.record(); torch.cuda.synchronize()
        nccl_us = torch.tensor([1e3 * e0.elapsed_time(e1) / reps], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(nccl_us, op=dist.ReduceOp.MAX)
        rec.update({"nccl_allgather_plus_maxreduce_us_per_iteration": nccl_us.item(), "us_per_evaluation_sharded": 1e3 * msN / evalsN,
                    "us_per_evaluation_single_gpu": 1e3 * ms1 / evals1})
        rec.update({"evals_per_s": evalsN / (msN * 1e-3), "ms": msN, "speedup_over_single_gpu": ms1 / msN,
                    "bitwise_equal_to_single_gpu": bool(same.item() == 1.0),
                    "exchange": "device-initiated pull of the line sums from CUDA-IPC-mapped peer buffers over NVLink (no collective call)"})
    else:
        rec.update({"evals_per_s": evals1 / (ms1 * 1e-3), "ms": ms1, "bitwise_equal_to_single_gpu": None})
    kern.close()
    return rec


def config2_record(pkg, dev):
    """BASELINE config 2: Nk=100, eta=0.51, defaults -- the launch-bound regime: one persistent launch, microseconds per evaluation."""
    sf = pkg.structure_factors
    p = problem(sf, eta=0.51, Nk=100)
    kern = pkg.ModeCouplingKernel(p["rho"], 1.0, 1.0, p["k"], p["Sk"], device=dev)
    eq = pkg.MemoryEquation(0.0, 1.0, p["gamma"], 0.0, p["F0"], p["dF0"], kern)
    solver = pkg.TimeDoublingSolver(max_iterations=10 ** 6)
    d = pkg.api.create_equation(eq, solver)
    ms = []
    for _ in range(4):
        ev, t = pkg.api.solve_resident(d)
        ms.append(t)
    d.close(); kern.close()
    return {"workload": "config2: hard-sphere MCT, Nk=100, eta=0.51, TimeDoublingSolver defaults, one persistent launch (thread-block cluster team)",
            "kernel_evals": int(ev), "solve_ms": float(np.mean(ms[1:])), "us_per_eval": 1e3 * float(np.mean(ms[1:])) / ev,
            "evals_per_s": ev / (float(np.mean(ms[1:])) * 1e-3)}


def parity_against_fixture(t, F, K, evals):
    """The host buffers of the timed end-to-end solve against the oracle's solution of the same equation (scripts/gen_fixtures.py)."""
    path = os.path.join(ROOT, "tests", "golden", "c3_nk1000_eta0.5155.npz")
    if not os.path.exists(path):
        return False, {"error": "tests/golden/c3_nk1000_eta0.5155.npz missing"}
    fx = np.load(path)
    cols, rows = fx["cols"], fx["rows"]
    dF = max(np.max(np.abs(F[:, cols] - fx["F_cols"])), np.max(np.abs(F[rows, :] - fx["F_rows"])))
    scaleK = np.max(np.abs(fx["K_rows"]))
    dK = max(np.max(np.abs(K[:, cols] - fx["K_cols"])), np.max(np.abs(K[rows, :] - fx["K_rows"]))) / scaleK
    ds = float(np.max(np.abs(F.sum(axis=1) - fx["F_sum_t"])) / F.shape[1])
    ok = bool(np.array_equal(t, fx["t"]) and dF <= 1e-10 and int(evals) == int(fx["kernel_evals"]))
    return ok, {"against": "tests/golden/c3_nk1000_eta0.5155.npz (CPU oracle, full solve)", "max_abs_dF": float(dF), "tolerance": 1e-10,
                "max_abs_dK_over_maxK": float(dK), "checksum_dev_per_entry": ds, "kernel_evals": int(evals),
                "kernel_evals_oracle": int(fx["kernel_evals"]), "time_grid_identical": bool(np.array_equal(t, fx["t"]))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the config-5 sub-record")
    ap.add_argument("--no-kshard", action="store_true", help="skip the config-4 sub-record")
    ap.add_argument("--points", type=int, default=512)
    ap.add_argument("--sweep-nk", type=int, default=500)
    ap.add_argument("--mc-nk", type=int, default=2000)
    ap.add_argument("--mc-steps", type=int, default=3)
    ap.add_argument("--mc-warmup", type=int, default=1)
    ap.add_argument("--mc-tmax", type=float, default=1e-2)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_pkg
    pkg = load_pkg()
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank

    p = problem(pkg.structure_factors)
    kern = pkg.ModeCouplingKernel(p["rho"], 1.0, 1.0, p["k"], p["Sk"], device=dev)
    eq = pkg.MemoryEquation(0.0, 1.0, p["gamma"], 0.0, p["F0"], p["dF0"], kern)
    solver = pkg.TimeDoublingSolver(max_iterations=10 ** 6)
    nt = solver.output_len()
    # pinned host buffers for the end-to-end leg (outputs t, F, K of one solve)
    pin = lambda n: torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()
    host_out = (pin(nt), pin(nt * NK).reshape(nt, NK), pin(nt * NK).reshape(nt, NK))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{dev}")      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(dev_eq):
        flush.fill_(1)                      # L2 flush between timed iterations
        torch.cuda.synchronize()
        evals, ms = pkg.api.solve_resident(dev_eq)
        return evals, ms

    # ---- device-resident leg: equation created (uploaded) once, the timed region is the persistent launch only
    dev_eq = pkg.api.create_equation(eq, solver)
    for _ in range(args.warmup):
        device_step(dev_eq)
    barrier()
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    launches0 = pkg.launch_count()
    evals_tot, ms_tot, ms_list = 0, 0.0, []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        ev, ms = device_step(dev_eq)
        evals_tot += ev; ms_tot += ms; ms_list.append(ms)
    barrier()
    wall_region = time.perf_counter() - t_wall0
    launches = pkg.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    dev_eq.close()

    # ---- end-to-end leg: host buffers in, host buffers out, every step
    def e2e_step():
        flush.fill_(1); torch.cuda.synchronize()
        t0 = time.perf_counter()
        sol = pkg.api.solve_into(eq, solver, *host_out)
        return solver.kernel_evals, time.perf_counter() - t0
    for _ in range(args.warmup):
        e2e_step()
    barrier()
    e2e_evals, e2e_s = 0, 0.0
    for _ in range(args.steps):
        ev, s = e2e_step()
        e2e_evals += ev; e2e_s += s
    barrier()
    parity_ok, parity = parity_against_fixture(host_out[0], host_out[1], host_out[2], ev)      # the LAST timed solve's host buffers
    kern.close()

    # ---- the other BASELINE configurations, measured by every run (all ranks take part in the sharded ones)
    sweep_rec = None if args.no_sweep else sweep_record(args, pkg, torch, dist, rank, world, dev)
    kshard_rec = None if args.no_kshard else kshard_record(args, pkg, torch, dist, rank, world, dev)

    # ---- max over ranks, whole-job aggregate
    stats = torch.tensor([ms_tot, e2e_s], dtype=torch.float64, device=f"cuda:{dev}")
    sums = torch.tensor([float(evals_tot), float(e2e_evals), float(launches)], dtype=torch.float64, device=f"cuda:{dev}")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms_max, e2e_max = stats.tolist()
    evals_all, e2e_evals_all, launches_all = sums.tolist()

    if rank == 0:
        peak, peak_src = measured_peak()
        bytes_per_eval = 24.0 * NK * NK + 16.0 * NK
        launch_ms = ms_tot / args.steps
        evals_per_solve = evals_tot / args.steps
        achieved = evals_per_solve * bytes_per_eval / (launch_ms * 1e-3) / 1e9
        # FP64 instructions the kernel really issues per evaluation: every warp walks U units of 32 positions x 4 rows,
        # 4 DMUL + 12 DFMA per lane and unit (padding units included) -- G*8*U*32*16 lane-instructions
        geo = pkg.api.last_geometry() if hasattr(pkg.api, "last_geometry") else None
        fp64_lane_instr = (148 * 8 * 8 * 32 * 16) if not geo else geo["G"] * 8 * geo["U"] * 32 * 16
        fp64_rate = fp64_lane_instr * evals_per_solve / (launch_ms * 1e-3)
        h2d = 6 * NK * 8
        d2h = (2 * nt * NK + nt) * 8
        line = {
            "metric": METRIC, "value": evals_all / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (analytic Percus-Yevick hard-sphere S(k))",
            "config": {"workload": WORKLOAD, "l2": "256 MiB buffer written between timed iterations (L2 flush)"},
            "details": {"kernel_evals_per_solve": evals_per_solve, "output_times": nt, "solve_wall_ms_device": launch_ms,
                        "sharding": "one independent equation per GPU, no data-path collective" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_evals_all / e2e_max, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "solve_wall_ms": 1e3 * e2e_max / args.steps},
            "gpu_launches": int(launches_all),
            "parity_checked": parity_ok, "parity": parity,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH,
                         "kernel": "td_sc_kernel (one persistent cooperative launch per solve)",
                         "algorithmic_bytes_per_eval": bytes_per_eval, "peak_source": peak_src,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one td_sc_kernel launch, ncu --set full capture of "
                                           "the Nk=1000 solve (" + NCU_TRAFFIC_SOURCE + "); not re-measured by this run",
                         "bound_actual": "latency: one inter-SM all-gather through L2 per evaluation (arrival counter + one load round trip, "
                                         "~1.1 us) and per-warp instruction issue at 8 warps/SM; neither HBM nor FP64 nor shared memory binds",
                         "note": "the vertex tables stay in tensor memory (TMEM) for the whole solve, so 'achieved' is an HBM-equivalent "
                                 "figure of the algorithmic bytes (SURVEY 8d); DRAM traffic per launch is one table load + the trajectory "
                                 "store, ~10^4 x less than the algorithmic bytes"},
            "roofline_fp64": {"achieved": fp64_rate / 1e12, "peak": 17.7, "unit": "T FP64 instr/s (an FMA counts once)", "frac": fp64_rate / 17.7e12,
                              "peak_source": "measured (profiles/fp64_bench.txt: 60.9 DFMA lanes/clk/SM)",
                              "fp64_lane_instructions_per_eval": fp64_lane_instr},
            "clocks": clocks,
        }
        if sweep_rec is not None:
            line["sweep"] = sweep_rec
        if kshard_rec is not None:
            line["kshard"] = kshard_rec
        line["configs"] = {"config1": config1_f2_on_cpu(), "config2": config2_record(pkg, dev)}
        if not args.no_cpu_baseline:
            v, sample = cpu_baseline_sample()
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
        if not parity_ok:
            raise SystemExit("bench.py: the timed solve does not match the oracle fixture: " + json.dumps(parity))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
