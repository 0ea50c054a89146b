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

N > 1: the path shards over independent equations only -- every rank solves its own equation of the
same configuration (weak scaling, no data-path collective); NCCL is used for the barrier and the max over ranks.
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
NCU_TRAFFIC_BYTES_PER_LAUNCH = 34611712 + 26191872      # profiles/ncu_full_r01_td_sc_kernel.csv
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


def cpu_baseline_sample(seconds=15.0):
    """The oracle on a bounded prefix of the SAME solve (first max_evals kernel evaluations of the Nk=1000 solve,
    bootstrap included), 1 thread."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import importlib.util
    spec = importlib.util.spec_from_file_location("mct_sf", os.path.join(ROOT, "modecouplingtheory.jl_b200", "structure_factors.py"))
    sf = importlib.util.module_from_spec(spec); spec.loader.exec_module(sf)
    p = problem(sf)
    kern = O.sc3d(p["rho"], 1.0, 1.0, p["k"], p["Sk"])
    # calibrate with a few evaluations, then run the real solver for a bounded number of evaluations
    t0 = time.perf_counter(); kern.eval_repeat(p["F0"], 3); per_eval = (time.perf_counter() - t0) / 3
    n = max(8, int(seconds / per_eval) - (2 * 32 + 1))
    t0 = time.perf_counter()
    res = O.td_solve(kern, 0.0, 1.0, p["gamma"], 0.0, p["F0"], p["dF0"], max_iterations=10 ** 6, max_evals=n)
    wall = time.perf_counter() - t0
    evals = res["kernel_evals"] + 2 * 32 + 1          # + bootstrap (2N) and K0 evaluations, not counted by solver.kernel_evals
    return evals / wall, f"first {res['kernel_evals']} fixed-point kernel evaluations (+{2 * 32 + 1} bootstrap) of the same Nk=1000 solve, {wall:.1f} s"


def run_reference(args, rank):
    if rank != 0:
        return
    vals = []
    sample = ""
    for i in range(args.warmup + args.steps):
        v, sample = cpu_baseline_sample(seconds=max(3.0, min(20.0, 60.0 / max(1, args.warmup + args.steps))))
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic (analytic Percus-Yevick hard-sphere S(k))",
            "config": {"workload": WORKLOAD, "note": "reference = ModeCouplingTheory.jl (Julia, not installable here: no julia binary, "
                       "no network); timed instead: oracle/mct_oracle.c, the C restatement in the reference's operation order"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample +
                             "; 1 thread because the reference hot path is single-threaded (@turbo, no @threads)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_mc_workload(args, pkg, torch, dist, rank, world, dev):
    """Config 4: ONE multi-component equation (Ns=4 hard-sphere mixture, Nk=2000), k-sharded over the ranks: rank r
    evaluates the lines L = r (mod world) of the vertex sum and stores their sums into every rank's buffer over NVLink
    from inside its persistent kernel.  Strong scaling; the time axis is shortened (t_max) so that a step stays bounded."""
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
    dev_eq = pkg.api.create_equation(eq, solver)
    if world > 1:
        def all_gather(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        pkg.api.shard_equation(dev_eq, rank, world, all_gather)
    ms_all = []
    for _ in range(args.mc_warmup + args.mc_steps):      # every pass is the whole solve again from F0
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        evals, ms = pkg.api.solve_resident(dev_eq)
        ms_all.append(ms)
    ms = float(np.mean(ms_all[args.mc_warmup:]))
    stats = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{dev}")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms_max = stats.item()
        n_out = solver.output_len()
        total_evals = evals + 2 * solver.N + 1
        print(json.dumps({
            "metric": "mc_ns4_kernel_evals_per_s", "value": total_evals / (ms_max * 1e-3), "unit": "kernel evals/s", "n_gpus": world,
            "steps": args.mc_steps, "warmup": args.mc_warmup, "ms_per_step": ms_max, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (Percus-Yevick hard-sphere mixture, Baxter factorisation)",
            "config": {"workload": f"config4: MultiComponentModeCouplingKernel Ns=4 (diameters 0.7/0.8/0.9/1.0, equal fractions, phi=0.50), "
                       f"Nk={Nk}, Newtonian, TimeDoublingSolver(N=4, dt=1e-4, t_max={args.mc_tmax}, tol=1e-10); k-sharded over {world} GPU(s)",
                       "kernel_evals": total_evals, "output_times": n_out,
                       "fp64_instr_per_eval": 14.0 * ns * ns * Nk * Nk * 0.75,      # 14 per (iq,ip,a,b) entry, 3/4 Nk^2 entries walked
                       "achieved_fp64_ginstr_per_s": 14.0 * ns * ns * Nk * Nk * 0.75 * total_evals / (ms_max * 1e-3) / 1e9,
                       "fp64_peak_ginstr_per_s": 17700.0,                   # profiles/fp64_bench.txt (an FMA counts once)
                       "exchange": "device-initiated stores of the line sums into CUDA-IPC-mapped peer buffers (no collective call)" if world > 1 else "none"}}),
            flush=True)


def run_sweep_workload(args, pkg, torch, dist, rank, world, dev):
    """Config 5: `points` packing fractions in [0.50, 0.53] at Nk = sweep_nk, sharded over the ranks (cost-aware snake
    partition), `solve_batch` per GPU (one persistent launch, teams pull equations from a queue), final gather of
    F(k_peak, t_end) to locate eta_c.  One step = the whole sweep."""
    sw = pkg.sweep
    etas = np.linspace(0.50, 0.53, args.points)
    mine = sw.partition(sw.expected_cost(etas), world, rank)
    kpeak = int(np.argmin(np.abs(pkg.structure_factors.wavenumber_grid(args.sweep_nk) - 7.0)))
    nt = pkg.TimeDoublingSolver().output_len()
    # only F(k_peak, t=0) and F(k_peak, t_end) of every solution cross PCIe (mct_equation_download_slice)
    obs = lambda sol: np.array([sol.F[-1, 0] / sol.F[0, 0]])
    slice_ = dict(times=[0, nt - 1], dofs=[kpeak], K=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    vals, evals, ms = sw.run_sweep(pkg, etas[mine], args.sweep_nk, dev, batch=64, observable=obs, download=slice_)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    stats = torch.tensor([wall, ms], dtype=torch.float64, device=f"cuda:{dev}")
    tot = torch.tensor([float(evals)], dtype=torch.float64, device=f"cuda:{dev}")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    allv = sw.gather(mine, vals, len(etas), world, dist if world > 1 else None)
    if rank == 0:
        wall_max, ms_max = stats.tolist()
        nonerg = allv[:, 0] > 0.1
        eta_c = float(etas[np.argmax(nonerg)]) if nonerg.any() else None
        print(json.dumps({
            "metric": "eta_sweep_solves_per_s", "value": args.points / wall_max, "unit": "solves/s", "n_gpus": world, "steps": 1,
            "warmup": 0, "ms_per_step": 1e3 * wall_max, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (analytic Percus-Yevick hard-sphere S(k))",
            "config": {"workload": f"config5: {args.points}-point packing-fraction sweep eta in [0.50, 0.53], Nk={args.sweep_nk}, "
                       "overdamped, TimeDoublingSolver defaults; wall clock incl. table construction, uploads and trajectory downloads",
                       "device_ms_max_over_ranks": ms_max, "kernel_evals_total": tot.item(), "eta_c_located": eta_c,
                       "sharding": "cost-aware snake partition over ranks, no data-path collective, final gather of one observable"}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="solve", choices=["solve", "sweep", "mc"],
                    help="solve: config 3 (the contract line).  sweep: config 5, the packing-fraction sweep sharded over the ranks")
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

    if args.workload == "mc":
        run_mc_workload(args, pkg, torch, dist, rank, world, dev)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload == "sweep":
        run_sweep_workload(args, pkg, torch, dist, rank, world, dev)
        if world > 1:
            dist.destroy_process_group()
        return

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
        achieved = (evals_tot / args.steps) * bytes_per_eval / (launch_ms * 1e-3) / 1e9
        h2d = 6 * NK * 8
        d2h = (2 * nt * NK + nt) * 8
        line = {
            "metric": METRIC, "value": evals_all / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (analytic Percus-Yevick hard-sphere S(k))",
            "config": {"workload": WORKLOAD, "kernel_evals_per_solve": evals_tot / args.steps, "output_times": nt,
                       "solve_wall_ms_device": launch_ms, "l2": "256 MiB buffer written between timed iterations (L2 flush)",
                       "sharding": "one independent equation per GPU, no data-path collective" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_evals_all / e2e_max, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "solve_wall_ms": 1e3 * e2e_max / args.steps},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH,
                         "kernel": "td_sc_kernel (one persistent cooperative launch per solve)",
                         "algorithmic_bytes_per_eval": bytes_per_eval, "peak_source": peak_src,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one td_sc_kernel launch, ncu --set full capture of "
                                           "the Nk=1000 solve (profiles/ncu_full_r01_td_sc_kernel.csv); not re-measured by this run",
                         "note": "vertex tables stay in registers + shared memory for the whole solve; DRAM traffic per launch is "
                                 "one table load + the trajectory store, ~6000x less than the algorithmic bytes"},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            v, sample = cpu_baseline_sample()
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
