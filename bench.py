#!/usr/bin/env python
"""bench.py — solver iterations/s of the parallel-krylov hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Headline workload (all N, strong scaling): CG on the 3-D 7-point Poisson system 512^3 (n = 134 217 728,
nnz = 937 951 232), fp64, x0 = 0, tol = 1e-8, one *step* = one full solve through the public entry point (iteration
cap 2500; it converges earlier) (BASELINE.json: "CG iters/s & achieved HBM GB/s, 3D Poisson fp64 at 1/2/4/8 B200"; the
north-star scaling target is quoted on 512^3).  Inputs are far larger than L2 (11.8 GB of CSR + 1 GiB vectors), so no L2 flush is needed.

ONE JSON line on stdout (rank 0).  Everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("PK_QUIET", "1")

import numpy as np  # noqa: E402

# name -> (solver, k, matrix kind, dims, iteration cap per step)
WORKLOADS = {
    "cg_p3d512": ("cg", None, "stencil", (512, 512, 512), 2500),   # full solve: converges in ~1.9k iterations
    "cg_p3d256": ("cg", None, "stencil", (256, 256, 256), 500),
    "cg_p3d128": ("cg", None, "stencil", (128, 128, 128), 500),
    "cg_p2d256": ("cg", None, "stencil", (256, 256, 1), 763),
    "mrr_p3d128": ("mrr", None, "stencil", (128, 128, 128), 500),
    "mrr_p3d256": ("mrr", None, "stencil", (256, 256, 256), 500),
    "kskipcg4_p3d256": ("kskipcg", 4, "stencil", (256, 256, 256), 500),
    "kskipmrr8_p3d256": ("kskipmrr", 8, "stencil", (256, 256, 256), 495),
    "kskipmrr4_p3d256": ("kskipmrr", 4, "stencil", (256, 256, 256), 500),
    "kskipmrr8_band32m": ("kskipmrr", 8, "banded", (1 << 25, 13), 46),
    "cg_band32m": ("cg", None, "banded", (1 << 25, 13), 42),
    "adaptive8_p3d512": ("adaptivekskipmrr", 8, "stencil", (512, 512, 512), 297),
    # opt-in extensions: single-reduction CG, Chebyshev-basis k-skip MrR
    "kskipmrr8_p3d512": ("kskipmrr", 8, "stencil", (512, 512, 512), 297),
    "cgcg_p3d512": ("cgcg", None, "stencil", (512, 512, 512), 2500),
    "cgcg_p3d256": ("cgcg", None, "stencil", (256, 256, 256), 500),
}
DEFAULT_WORKLOAD = "cg_p3d512"


# stdout carries exactly ONE JSON line: anything a library prints to fd 1 (NCCL prints its version there) is sent to
# stderr instead; the JSON goes to a private duplicate of the original stdout.
_JSON_OUT = None


def _claim_stdout():
    global _JSON_OUT
    if _JSON_OUT is None:
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def algorithmic_bytes(solver, k, n, nnz):
    """minimal-traffic model, per solver iteration (and for one operator application).
    MrR: §8d's primary figure B_spmv + 80n (the literal three-phase form moves 104n, see actual_bytes)."""
    b_spmv = 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n
    if solver == "cg":
        per_it = b_spmv + 72.0 * n
    elif solver == "cgcg":
        per_it = b_spmv + 80.0 * n           # p, s, x, r updated from u(=r), w: 6R 4W
    elif solver == "mrr":
        per_it = b_spmv + 80.0 * n
    elif solver == "kskipcg":
        per_it = ((3 * k + 2) * b_spmv + 8.0 * n * (2 * k + 3) + 56.0 * n * (k + 1)) / (k + 1)
    else:
        per_it = ((3 * k + 2) * b_spmv + 8.0 * n * (2 * k + 3) + 72.0 * n * (k + 1)) / (k + 1)
    return b_spmv, per_it


def one_pass_basis(name):
    """Does the workload's operator qualify for the one-pass matrix-powers kernel (csrc/pk_matpow.cu: half bandwidth bw
    with W - 2(k-1)bw >= W/2; the banded workloads are full bands of <= 27 diagonals, whose kernel has a window of W =
    768 rows)?  Only the banded systems do; 3-D stencils have bw = nx*ny."""
    solver, k, kind, dims, _ = WORKLOADS[name]
    if kind != "banded" or not k or k < 2 or os.environ.get("PK_MATPOW", "1") in ("0", ""):
        return False
    return 768 - 2 * (k - 1) * dims[1] >= 384 and 2 * dims[1] + 1 <= 27


def actual_bytes(solver, k, n, nnz, matpow=False):
    """Bytes the code REALLY moves per solver iteration with the default kernels (every array pass counted once, x
    gathers counted as one pass): the k-skip basis reads A once for both chains and the step SpMVs consume A·v in
    registers, so a trip makes 2k+1 passes over A, not the 3k+2 of §8d's formula — and k+2 passes when the one-pass
    matrix-powers kernel generates the basis (matpow: the dense-band kernel reads the values only — no column indices,
    no row pointers; ghost-row re-reads of A are served by L2 and not counted), and 2 passes per trip when the steps of
    a k-skip MrR trip are fused into one pass as well (dense band, one GPU)."""
    b_a = 12.0 * nnz + 4.0 * (n + 1)                     # one pass over the CSR arrays
    if solver == "cg":                                   # SpMV (A, p, v) + xr (4R 2W) + p (2R 1W)
        return b_a + 16.0 * n + 72.0 * n
    if solver == "mrr":                                  # SpMV (A, r, Ar, y for the dots) + s-phase (3R) + update (5R 4W)
        return b_a + 24.0 * n + 24.0 * n + 72.0 * n
    if solver == "cgcg":                                 # SpMV (A, u, w) + fused update (6R 4W, u == r without M)
        return b_a + 16.0 * n + 80.0 * n
    first = 56.0 if solver == "kskipcg" else 72.0        # un-fused first step of a trip
    fused = 48.0 if solver == "kskipcg" else 64.0        # step fused into the SpMV epilogue: vectors read + written
    if matpow:
        a_once = b_a if matpow == "general" else 8.0 * nnz
        basis = a_once + (2 + 2 * k) * 8.0 * n           # one pass over A, 2 inputs read, 2k level vectors written
    else:
        basis = k * (b_a + 32.0 * n)                     # k two-chain passes: A + 2 gathers + 2 stores each
    if matpow == "dense" and solver == "kskipmrr" and os.environ.get("PK_KSTEPS", "1") not in ("0", ""):
        # dense band, one GPU: the k+1 steps + closing mat-vec are ONE pass (k_mrr_steps_band): A's values once, the five
        # vectors read and written once, then r, A r, y copied home (3R 3W)
        steps = 8.0 * nnz + 80.0 * n + 48.0 * n
    else:
        steps = (k + 1) * b_a + first * n + k * fused * n + 16.0 * n     # k fused steps + closing SpMV
    trip = basis + steps + 8.0 * n * (2 * k + 3)         # + Gram: every basis vector once
    if solver == "adaptivekskipmrr":
        trip += 16.0 * n                                 # best-x snapshot per trip
    return trip / (k + 1)


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_solvers():
    """The CPU implementation the reference arm / cpu_baseline time: the reference's own v3/cpu files staged under
    oracle/_ref by build() (kind "reference"), else the oracle port (kind "port")."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    fns = ref_loader.load()
    if fns is not None:
        return fns, "reference", "oracle/_ref = the reference's own v3/cpu/*.py, unmodified (np.int shim, np.dot proxy for CSR)"
    import krylov_oracle as oracle
    return oracle.SOLVERS, "port", "oracle/krylov_oracle.py (port of v3/cpu; oracle/_ref not staged)"


def cpu_solver_for(solver):
    """(callable, kind, description) for one solver: the staged reference when it has that method (the five v3 methods),
    else the oracle port (cgcg: the reference's sketch cannot be imported)."""
    fns, kind, what = cpu_solvers()
    if solver in fns:
        return fns[solver], kind, what
    import krylov_oracle as oracle
    return oracle.SOLVERS[solver], "port", "oracle/krylov_oracle.py (repaired restatement; the reference file cannot be imported)"


# ------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, timed on this box's host cores.
    The reference is pure Python (numpy/scipy): build() stages its v3/cpu files under the git-ignored oracle/_ref and
    this arm runs THOSE (kind "reference"); without them it runs the oracle port (oracle/krylov_oracle.py —
    bit-identical to /root/reference/v3/cpu on the golden vectors).  Either way: scipy's serial csr_matvec +
    OpenBLAS-threaded dots, i.e. all the host threads the reference itself can use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    solver, k, kind, dims, cap = WORKLOADS[args.workload]
    cpu_fn, cpu_kind, cpu_what = cpu_solver_for(solver)          # also puts oracle/ on sys.path
    import host_kernels as hk
    from parallel_krylov_b200 import problems
    solvers_cpu = {solver: cpu_fn}
    t0 = time.time()
    if kind == "stencil":
        rowptr, col, val, n = hk.stencil_csr(*dims)
    else:
        rowptr, col, val, n = problems.banded_spd(*dims)
    A = problems.to_scipy(rowptr, col, val, n)
    b = problems.hash_normal(0, n)
    log(f"[reference] built {args.workload} on host in {time.time() - t0:.1f}s (n={n}, nnz={len(val)})")
    # bounded sample: a few iterations per step so the whole run stays within minutes
    kk = (k or 0) + 1
    probe_it = kk + (1 if solver != "cg" and solver != "kskipcg" else 0)
    kw = {"k": k} if k is not None else {}
    t1 = time.perf_counter()
    _, info = solvers_cpu[solver](A, b, tol=1e-8, maxiter=probe_it, **kw)
    per_it = max(info["time"] / max(int(info["nosl"][-1]), 1), 1e-9)
    total_steps = args.steps + args.warmup
    budget_s = 150.0
    it_per_step = int(max(kk, min(cap, budget_s / total_steps / per_it)))
    it_per_step = max(kk, (it_per_step // kk) * kk)
    log(f"[reference] probe: {per_it:.3f} s/iteration -> {it_per_step} iterations per step")
    for _ in range(args.warmup):
        solvers_cpu[solver](A, b, tol=1e-8, maxiter=it_per_step, **kw)
    its, secs = 0, 0.0
    for _ in range(args.steps):
        _, info = solvers_cpu[solver](A, b, tol=1e-8, maxiter=it_per_step, **kw)
        its += int(info["nosl"][-1])
        secs += info["time"]
    value = its / secs
    try:
        from threadpoolctl import threadpool_info
        blas_threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        blas_threads = os.cpu_count()
    sample = (f"{it_per_step} iterations per step of the same system (x0=0), timed like the reference "
              f"(perf_counter around the loop, after the initial residual)")
    line = {
        "impl": "reference", "metric": "solver_iterations_per_s", "value": value, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, n, len(val), it_per_step),
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": int(blas_threads), "kind": cpu_kind,
                         "implementation": cpu_what, "sample": sample, "host_cpus": os.cpu_count(),
                         "note": "scipy csr_matvec is serial; numpy dots use OpenBLAS threads"},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_config(name, n, nnz, cap):
    solver, k, kind, dims, _ = WORKLOADS[name]
    if kind == "stencil":
        desc = f"{'3-D 7-point' if dims[2] > 1 else '2-D 5-point'} Poisson {'x'.join(map(str, dims))}"
    else:
        desc = f"banded SPD, {2 * dims[1] + 1} diagonals"
    return {"workload": f"{solver}{'' if k is None else ' k=' + str(k)} on {desc} (n={n}, nnz={nnz}) fp64, "
                        f"x0=0, tol=1e-8, iteration cap {cap} per step",
            "name": name, "l2": "inputs >> L2 (no flush needed)" if nnz * 12 > 4e8 else "L2 flushed between steps",
            "partition": "contiguous block rows, sharded vectors, halo exchange + all-reduced dots"}


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import parallel_krylov_b200 as pk
    from parallel_krylov_b200 import device_problems as dp
    from parallel_krylov_b200 import _lib
    from parallel_krylov_b200._core import Context, Operator, solve
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = Context.get(local_rank)

    # ---- parity, where the driver measures: five solvers x two systems through parallel_krylov_b200.mpi.* at THIS GPU
    # count against the committed outputs of the unmodified reference (tests/golden); see parallel_krylov_b200/selfcheck.py
    parity = None
    if not args.no_parity:
        from parallel_krylov_b200 import selfcheck
        if world == 1:
            selfcheck.init_single_rank_group()
        t_par = time.time()
        parity = selfcheck.run(None, verbose=bool(os.environ.get("PK_BENCH_VERBOSE")))
        parity["seconds"] = round(time.time() - t_par, 2)
        if rank == 0:
            log(f"[parity] {parity['cases']} cases on {world} GPU(s): ok={parity['ok']} max_dev50={parity['max_dev50']:.2e} "
                f"(k=4: {parity['max_dev50_k4']:.2e}) in {parity['seconds']}s")

    solver, k, kind, dims, cap = WORKLOADS[args.workload]
    if args.maxiter:
        cap = args.maxiter

    def make_problem(name):
        """This rank's row block of a workload, generated directly in HBM -> (operator, b, csr tensors, sizes)."""
        _, _, kind_, dims_, _ = WORKLOADS[name]
        n_ = int(np.prod(dims_)) if kind_ == "stencil" else dims_[0]
        base_ = n_ // world
        row0_ = rank * base_
        n_loc_ = base_ if rank < world - 1 else n_ - row0_
        if kind_ == "stencil":
            rp_, cg_, va_, _ = dp.stencil_csr(*dims_, row0=row0_, n_rows=n_loc_, ctx=ctx)
        else:
            rp_, cg_, va_, _ = dp.banded_csr(dims_[0], dims_[1], 0, row0=row0_, n_rows=n_loc_, ctx=ctx)
        nnz_loc_ = int(va_.numel())
        b_ = dp.hash_normal(0, n_loc_, offset=row0_, ctx=ctx)
        if world > 1:
            from parallel_krylov_b200.mpi import DistOperator
            offs_ = [r * base_ for r in range(world)] + [n_]
            op_ = DistOperator.from_local_csr(rp_, cg_, va_, n_, None, ctx, row_offsets=offs_)
            t_ = torch.tensor([nnz_loc_], dtype=torch.int64, device=dev)
            dist.all_reduce(t_)
            nnz_ = int(t_.item())
        else:
            op_ = Operator.from_csr_tensors(rp_, cg_, va_, n_, ctx)
            nnz_ = nnz_loc_
        return op_, b_, (rp_, cg_, va_), (n_, n_loc_, nnz_, nnz_loc_)

    t0 = time.time()
    op, b, (rowptr, colg, val), (n, n_loc, nnz, nnz_loc) = make_problem(args.workload)
    torch.cuda.synchronize()
    if rank == 0:
        log(f"[ours] {args.workload}: n={n} nnz={nnz} on {world} GPU(s), generated in HBM in {time.time() - t0:.1f}s; "
            f"kernel {op.kernel_info()}, halo {op.n_halo}")
    kw = {"k": k} if k is not None else {}

    def one_solve(graph):
        return solve(solver, op, b, tol=1e-8, maxiter=cap, use_graph=graph, ctx=ctx, **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The timed region replays captured CUDA graphs (use_graph=True is the default of the public entry points); the
    # per-launch duration of the operator kernel for the roofline comes from a separate un-graphed pass of one full solve
    # right after it (event pairs around every launch cannot be recorded inside a graph replay).  --no-graph times plain
    # stream launches in both.
    use_graph = not args.no_graph
    for _ in range(args.warmup):
        one_solve(use_graph)
    # ---- timed region 1: inputs resident in HBM -----------------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = launches = spmvs = 0
    loop_s = 0.0
    for _ in range(args.steps):
        _, info = one_solve(use_graph)
        iters += info["iterations"]
        launches += info["gpu_launches"]
        spmvs += info["spmv"]
        loop_s += info["time"]
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    elapsed_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = iters / (elapsed_ms * 1e-3)

    # ---- profiling pass: one un-graphed solve, CUDA event pairs around every operator application -------------------
    prof_ms, prof_n = C.c_double(), C.c_int64()
    _lib.check(ctx.lib.pk_prof_begin(ctx.handle, 16384))
    barrier()
    e0.record()
    _, pinfo = one_solve(False)
    e1.record()
    barrier()
    _lib.check(ctx.lib.pk_prof_end(ctx.handle, C.byref(prof_ms), C.byref(prof_n)))
    prof_pass_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([prof_pass_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        prof_pass_ms = float(t.item())
    plain_launch_its = pinfo["iterations"] / (prof_pass_ms * 1e-3)

    # ---- exposed communication per iteration (N > 1): same kernels with halo exchange and all-reduces switched off ---
    exposed = None
    if world > 1:
        _lib.check(ctx.lib.pk_ctx_set_nocomm(ctx.handle, 1))
        cap_nc = min(cap, 200)
        solve(solver, op, b, tol=0.0, maxiter=cap_nc, use_graph=use_graph, ctx=ctx, **kw)
        barrier()
        e0.record()
        _, inc = solve(solver, op, b, tol=0.0, maxiter=cap_nc, use_graph=use_graph, ctx=ctx, **kw)
        e1.record()
        barrier()
        _lib.check(ctx.lib.pk_ctx_set_nocomm(ctx.handle, 0))
        t = torch.tensor([e0.elapsed_time(e1) / max(inc["iterations"], 1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        compute_only_ms = float(t.item())
        iter_ms = elapsed_ms / max(iters, 1)
        exposed = {"iteration_us": 1e3 * iter_ms, "compute_only_us": 1e3 * compute_only_ms,
                   "exposed_comm_us_per_iteration": 1e3 * (iter_ms - compute_only_ms),
                   "how": "same solve, same kernels, with the halo push / flag waits / all-reduces disabled "
                          "(pk_ctx_set_nocomm), max over ranks"}

    # ---- timed region 2: end to end through the public entry point with HOST buffers ---------------------------
    # (A, b in pinned host memory -> H2D every step, solve, x -> D2H every step)
    band_ext_active = getattr(op, "band_ext_op", None) is not None
    del op
    h_rowptr = rowptr.cpu().pin_memory()
    h_col = colg.cpu().pin_memory()
    h_val = val.cpu().pin_memory()
    h_b = b.cpu().pin_memory()
    h_x = torch.empty(n_loc, dtype=torch.float64).pin_memory()
    del rowptr, colg, val
    torch.cuda.empty_cache()

    def e2e_step():
        if world > 1:
            from parallel_krylov_b200 import mpi as pkm
            x, info = getattr(pkm, solver)(None, (h_rowptr, h_col, h_val, n), h_b, tol=1e-8, maxiter=cap,
                                           gather_x=False, use_graph=use_graph, **kw)
        else:
            x, info = getattr(pk, solver)((h_rowptr, h_col, h_val, n), h_b, tol=1e-8, maxiter=cap, use_graph=use_graph, **kw)
        h_x.copy_(x, non_blocking=False)
        return info

    e2e_warm = min(args.warmup, 1) if nnz > 2e8 else args.warmup
    for _ in range(e2e_warm):
        e2e_step()
    barrier()
    e0.record()
    it2 = 0
    h2d = 0
    e2e_loop_s = 0.0
    for _ in range(args.steps):
        info = e2e_step()
        it2 += info["iterations"]
        h2d = info["h2d_bytes"]
        e2e_loop_s += info["time"]
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        t = torch.tensor([h2d, n_loc * 8], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        h2d_total, d2h_total = int(t[0].item()), int(t[1].item())
    else:
        h2d_total, d2h_total = h2d, n_loc * 8
    e2e_value = it2 / (e2e_ms * 1e-3)

    # ---- the other BASELINE.json configs, measured briefly in the same run (default workload only) -------------------
    # N = 1: configs[1], [2], the north-star's k-skip MrR 256^3 target, CG 256^3, and configs[3], [4] on one GPU;
    # N > 1: configs[3] (kskipmrr k=8, banded 32M) and configs[4] (adaptivekskipmrr k=8, 512^3) row-partitioned over N,
    #        plus kskipmrr k=8 and the single-reduction CG (cgcg, opt-in) on the headline 512^3 system for comparison.
    other = None
    peak, peak_src = measured_peak_gbs()
    if args.workload == DEFAULT_WORKLOAD and not args.no_other:
        other = {}
        torch.cuda.empty_cache()
        names = (("mrr_p3d128", "kskipcg4_p3d256", "kskipmrr8_p3d256", "cg_p3d256", "mrr_p3d256", "kskipmrr8_band32m",
                  "adaptive8_p3d512") if world == 1 else ("kskipmrr8_band32m", "adaptive8_p3d512", "kskipmrr8_p3d512",
                                                          "cgcg_p3d512"))
        for name in names:
            s2, k2, _, _, cap2 = WORKLOADS[name]
            op2, b2, csr2, (n2, _, nnz2, _) = make_problem(name)
            del csr2
            kw2 = {"k": k2} if k2 is not None else {}
            solve(s2, op2, b2, tol=1e-8, maxiter=cap2, use_graph=True, ctx=ctx, **kw2)
            barrier()
            e0.record()
            its2 = 0
            for _ in range(2):
                _, i2 = solve(s2, op2, b2, tol=1e-8, maxiter=cap2, use_graph=True, ctx=ctx, **kw2)
                its2 += i2["iterations"]
            e1.record()
            barrier()
            ms2 = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms2], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms2 = float(t.item())
            v_its = its2 / (ms2 * 1e-3)
            _, pib = algorithmic_bytes(s2, k2 or 0, n2, nnz2)
            act = actual_bytes(s2, k2 or 0, n2, nnz2, one_pass_basis(name) and ("dense" if (world == 1 or getattr(op2, "band_ext_op", None) is not None) else "general"))
            other[name] = {"iterations_per_s": v_its, "iterations_per_solve": its2 / 2, "converged": bool(i2["converged"]),
                           "final_k": i2.get("final_k"),
                           "frac_formula": pib * v_its / 1e9 / (peak * world),
                           "frac_actual_bytes": act * v_its / 1e9 / (peak * world),
                           "bytes_per_iteration_formula": pib, "bytes_per_iteration_actual": act,
                           "workload": workload_config(name, n2, nnz2, cap2)["workload"]}
            if rank == 0:
                log(f"[other] {name}: {v_its:.1f} it/s, frac formula {other[name]['frac_formula']:.3f}, "
                    f"actual bytes {other[name]['frac_actual_bytes']:.3f}")
            del op2, b2
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0 if (parity is None or parity["ok"]) else 1

    # ---- roofline of the dominant kernel (operator application) -------------------------------------------------
    b_spmv_loc, _ = algorithmic_bytes(solver, k or 0, n_loc, nnz_loc)      # per launch on this rank
    _, per_it_bytes = algorithmic_bytes(solver, k or 0, n, nnz)
    per_it_actual = actual_bytes(solver, k or 0, n, nnz, one_pass_basis(args.workload) and ("dense" if (world == 1 or band_ext_active) else "general"))
    spmv_avg_ms = prof_ms.value / max(prof_n.value, 1)
    spmv_gbs = b_spmv_loc / (spmv_avg_ms * 1e-3) / 1e9 if spmv_avg_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_spmv_tma (y = A x with fused dots; TMA bulk-copy pipeline)", "achieved": spmv_gbs, "peak": peak,
                "unit": "GB/s", "frac": spmv_gbs / peak,
                # the ncu capture is of the single-GPU launch; a rank of an N-GPU run moves 1/N of it, so no number is claimed there
                "traffic": load_traffic(args.workload) if world == 1 else None,
                "algorithmic_bytes_per_launch": b_spmv_loc, "avg_launch_ms": spmv_avg_ms,
                "launches_timed": int(prof_n.value), "peak_source": peak_src,
                "timing": "CUDA event pairs around every operator application of one full un-graphed solve of the same "
                          "workload, run right after the timed region (rank 0's launches)",
                "share_of_iteration": (spmv_avg_ms * prof_n.value) / prof_pass_ms if prof_pass_ms > 0 else None,
                "whole_iteration": {"bytes_per_iteration": per_it_bytes,
                                    "achieved_gbs": per_it_bytes * value / 1e9 / 1.0,
                                    "frac_of_peak_x_gpus": per_it_bytes * value / 1e9 / (peak * world),
                                    "bytes_per_iteration_actual": per_it_actual,
                                    "frac_actual_bytes": per_it_actual * value / 1e9 / (peak * world)}}

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample (N=1 only) -----------------------
    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        cpu_fn, cpu_kind, cpu_what = cpu_solver_for(solver)
        solvers_cpu = {solver: cpu_fn}
        import scipy.sparse as sp
        A = sp.csr_matrix((h_val.numpy(), h_col.numpy(), h_rowptr.numpy()), shape=(n, n))
        kk = (k or 0) + 1
        m = kk * max(1, int(round(3 / kk))) if nnz > 2e8 else kk * max(1, 20 // kk)
        t_cpu = time.perf_counter()
        _, ci = solvers_cpu[solver](A, h_b.numpy(), tol=1e-8, maxiter=m, **kw)
        # extend the sample to ~10-30 s if the first probe was short
        if ci["time"] < 5.0:
            m2 = int(min(cap, m * max(2, int(12.0 / max(ci["time"], 1e-3))))) // kk * kk
            _, ci = solvers_cpu[solver](A, h_b.numpy(), tol=1e-8, maxiter=max(m2, kk), **kw)
        try:
            from threadpoolctl import threadpool_info
            blas_threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
        except Exception:
            blas_threads = os.cpu_count()
        cpu_its = int(ci["nosl"][-1])
        cpu_baseline = {"value": cpu_its / ci["time"], "unit": "iterations/s", "cores": int(blas_threads),
                        "kind": cpu_kind, "implementation": cpu_what, "host_cpus": os.cpu_count(),
                        "sample": f"{cpu_its} iterations of the same system on the host (scipy serial csr_matvec + "
                                  f"OpenBLAS dots), {ci['time']:.1f}s, timed like the reference"}
        log(f"[cpu_baseline] {cpu_baseline['value']:.4f} it/s ({time.perf_counter() - t_cpu:.1f}s total)")

    line = {
        "metric": "solver_iterations_per_s", "value": value, "unit": "iterations/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, n, nnz, cap),
        "iterations_per_step": iters / max(args.steps, 1), "loop_only_iterations_per_s": iters / loop_s,
        "launch_mode": "cuda-graph replay per batch of iterations" if use_graph else "plain stream launches",
        "plain_launch_iterations_per_s": plain_launch_its,
        "gpu_launches": int(launches), "spmv_launch_count": int(spmvs),
        "e2e": {"value": e2e_value, "unit": "iterations/s", "h2d_bytes_per_step": int(h2d_total),
                "d2h_bytes_per_step": int(d2h_total), "ms_per_step": e2e_ms / max(args.steps, 1),
                "solver_loop_ms_per_step": 1e3 * e2e_loop_s / max(args.steps, 1)},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks, "exposed_comm": exposed,
        "other_baseline_configs": other, "parity": parity,
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        log("PARITY FAILED: see the 'parity' object of the JSON line")
        return 1
    return 0


def load_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed
    `ncu --set full` capture (profiles/traffic.json), or null when none has been taken for this workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(workload, {}).get("spmv_dram_bytes_per_launch")
    except Exception:
        return None


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--maxiter", type=int, default=0, help="override the per-step iteration cap")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the golden-vector parity solves before the timed region")
    ap.add_argument("--no-other", action="store_true", help="skip the brief runs of the other BASELINE configs")
    ap.add_argument("--no-graph", action="store_true",
                    help="plain stream launches in the timed region instead of CUDA-graph replay")
    ap.add_argument("--graph", action="store_true", help="(default now; kept for compatibility)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("note: timing rules ask for >= 3 warm-up steps")
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
