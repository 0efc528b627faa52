#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config, one JSON line on stdout.

  metric   CG iterations/s, Jacobi-preconditioned CG, 3-D 7-point Poisson 512^3 fp64
           (configs[2]: "Jacobi-preconditioned CG on 3D Poisson 512^3 fp64 row-partitioned across
           1/2/4/8 B200"; it fits one GPU, so N=1 runs the same 512^3 system on one device).
  step     one ab_cg_fixed / ab_dist_cg_fixed call = ITERS_PER_STEP CG iterations from x0 = 0 on
           b = A*1 (the solve restarts every step, so every step does identical work).
  value    device-resident: f and u live in HBM when the timed region starts.
  e2e      the same call through the C-ABI with HOST (pinned) f and u: the H2D copy of f and the
           D2H copy of u are inside the timed region every step.
  N > 1    one process per GPU (torchrun); z-slab row partition of the SAME 512^3 system
           (strong scaling); halos + inner products move inside libadcme_b200 (NCCL over NVLink).
  --impl reference   the CPU arm: the oracle's port of the same recurrence on the box's host
           threads (the reference has no iterative solver and its direct Eigen solve is
           infeasible at 512^3 — DESIGN.md §measurement); rank 0 only.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

ITERS_PER_STEP = 200      # "time >= 200 CG iterations"
REF_ITERS_PER_STEP = 10
METRIC = "cg_iters_per_s_poisson3d_512_fp64"
UNIT = "CG iterations/s"


# ----------------------------------------------------------------------------- host helpers
def slab_rows(n: int, world: int, rank: int):
    """Inclusive global row range of `rank`'s z-slab of the n^3 grid (whole xy-planes)."""
    planes = n
    lo = (planes * rank // world) * n * n
    hi = (planes * (rank + 1) // world) * n * n if rank < world - 1 else n ** 3
    return lo, hi - 1


def _reduce(x: float, op, device):
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(x: float, device="cuda"):
    import torch.distributed as dist

    return _reduce(x, dist.ReduceOp.MAX, device)


def sum_over_ranks(x: float, device="cuda"):
    import torch.distributed as dist

    return _reduce(x, dist.ReduceOp.SUM, device)


def measured_peak_gbs():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def spmv_bytes(N: int, nnz: int) -> float:
    return 12.0 * nnz + 4.0 * (N + 1) + 16.0 * N


def cg_iter_bytes(N: int, nnz: int) -> float:
    return 12.0 * nnz + 4.0 * (N + 1) + 13.0 * 8.0 * N


def poisson_nnz(n: int) -> int:
    return 7 * n ** 3 - 6 * n * n


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(gpu_index)],
                stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(",") for r in open(self.tmp.name) if r.strip()]
        os.unlink(self.tmp.name)
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); out["sm_max_mhz"] = float(r[1])
                for nm, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            out["sm_mhz"] = statistics.median(sm)
        out["reasons"], out["samples"] = sorted(reasons), len(sm)
        return out


# ----------------------------------------------------------------------------- CPU arm
def cpu_cg_sample(n_req: int, iters: int, reps: int = 1, warm: int = 0):
    """Times the oracle's Jacobi-PCG port (all host threads) on the n^3 system; falls back to a
    256^3 sample (rescaled by rows) when host memory is short.  Returns a dict."""
    import numpy as np
    import __graft_entry__ as ge

    orc = ge.load_oracle()
    avail_gb = 0.0
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                avail_gb = float(line.split()[1]) / 1e6
    except Exception:
        pass
    n = n_req
    need_gb = (12.0 * poisson_nnz(n) + 8.0 * 7 * n ** 3) / 1e9
    scale = 1.0
    if avail_gb and need_gb * 1.5 > avail_gb:
        n = 256 if n_req > 256 else n_req
        scale = (n ** 3) / float(n_req ** 3)      # iterations/s shrink with the rows touched
    rp, ci, va = orc.poisson3d_csr(n, n, n)
    b = orc.csr_spmm(rp, ci, va, np.ones(n ** 3))
    times = []
    for r in range(warm + reps):
        t0 = time.perf_counter()
        x, it, rel = orc.pcg_jacobi(rp, ci, va, b, tol=-1.0, maxit=iters)
        dt = time.perf_counter() - t0
        if r >= warm:
            times.append(dt)
    its = iters / (sum(times) / len(times)) * scale
    sample = f"{iters} Jacobi-PCG iterations x {len(times)} on the {n}^3 system, {orc.num_threads()} OpenMP threads"
    if scale != 1.0:
        sample += f" (host RAM {avail_gb:.0f} GB < needed; iterations/s rescaled by rows {n}^3/{n_req}^3)"
    return {"value": its, "unit": UNIT, "cores": orc.num_threads(), "kind": "port", "sample": sample,
            "times": times, "relres": rel, "n": n}


def run_reference(args, rank: int):
    if rank != 0:
        return
    n = args.grid
    res = cpu_cg_sample(n, REF_ITERS_PER_STEP, reps=args.steps, warm=args.warmup)
    ms = 1e3 * sum(res["times"]) / len(res["times"])
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n, args.gpus, REF_ITERS_PER_STEP),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference arm = oracle port of the same Jacobi-PCG recurrence on host threads; the reference itself "
                "has no CG and its Eigen SparseLU `\\` is infeasible at 512^3 (SURVEY F2-F3)",
    }
    print(json.dumps(line), flush=True)


def workload_config(n: int, gpus: int, iters_per_step: int):
    return {
        "workload": f"Jacobi-PCG on 3-D 7-point Poisson {n}^3 fp64 (BASELINE.json configs[2]), b = A*1, x0 = 0",
        "rows": n ** 3, "nnz": poisson_nnz(n), "iters_per_step": iters_per_step,
        "partition": f"z-slab row blocks over {gpus} GPU(s)" if gpus > 1 else "single GPU",
        "l2": "inputs larger than L2 (one CG iteration streams %.1f GB)" % (cg_iter_bytes(n ** 3, poisson_nnz(n)) / 1e9),
    }


# ----------------------------------------------------------------------------- GPU arm
def run_gpu(args, rank: int, world: int, local_rank: int):
    import numpy as np
    import torch
    import __graft_entry__ as ge

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    pkg = ge.load_package()
    lib = pkg.lib
    for kv in filter(None, os.environ.get("ADCME_B200_OPTS", "").split(",")):   # e.g. dist_p2p=0
        key, val = kv.split("=")
        pkg.check(lib.ab_set_option(key.encode(), float(val)))
    n = args.grid
    N, nnz = n ** 3, poisson_nnz(n)
    pkg.mpi_init(rank, world, local_rank)
    for kv in filter(None, os.environ.get("AB_OPTS", "").split(",")):       # A/B switches, e.g. AB_OPTS=dist_fused_wait=0
        k_, v_ = kv.split("=")
        pkg.check(lib.ab_set_option(k_.strip().encode(), float(v_)))
    A = pkg.mpi_SparseTensor.poisson3d(n, n, n, rank, world)
    nloc = A.nloc
    ones = torch.ones(nloc, dtype=torch.float64, device=dev)
    b = A @ ones                                    # device-resident right-hand side
    u = torch.empty_like(b)
    b_host = torch.empty(nloc, dtype=torch.float64).pin_memory()
    u_host = torch.empty(nloc, dtype=torch.float64).pin_memory()
    b_host.copy_(b)
    torch.cuda.synchronize()
    rel = C.c_double(0)

    def step(dst, src):
        pkg.check(lib.ab_dist_cg_fixed(A._dh, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()),
                                       ITERS_PER_STEP, C.byref(rel)))

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def timed(dst, src, sample_clocks):
        for _ in range(args.warmup):
            step(dst, src)
        barrier()
        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        l0 = lib.ab_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step(dst, src)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.ab_launch_count() - l0
        clocks = sampler.stop() if sampler else None
        ms = max_over_ranks(ms, dev)
        return ms, launches, clocks

    ms_dev, launches, clocks = timed(u, b, True)
    relres_dev = rel.value
    # sanity guard inside the bench: the timed iterations did real work — the true residual of the
    # returned iterate matches the recurrence residual the solver reports, and it has dropped
    true_res = float(((A @ u) - b).norm().item())
    if world > 1:
        true_res = sum_over_ranks(true_res ** 2, dev) ** 0.5
        bnorm = sum_over_ranks(float(b.norm().item()) ** 2, dev) ** 0.5
    else:
        bnorm = float(b.norm().item())
    err = true_res / bnorm
    if not (err < 0.5 and abs(err - relres_dev) <= 1e-6 * max(err, 1e-300) + 1e-9):
        raise SystemExit(f"bench sanity check failed: true relres {err} vs reported {relres_dev}")
    ms_e2e, _, _ = timed(u_host, b_host, False)
    total_iters = ITERS_PER_STEP * args.steps
    value = total_iters / (ms_dev * 1e-3)
    e2e_value = total_iters / (ms_e2e * 1e-3)

    # ---- per-kernel durations (CUDA events on the launching stream), rank 0's local block
    peak, peak_src = measured_peak_gbs()
    roof, kernels, spmv_info = None, None, None
    if rank == 0:
        k1, k2, k3 = C.c_double(0), C.c_double(0), C.c_double(0)
        hloc = A.local.handle()
        if world == 1:
            pkg.check(lib.ab_bench_cg_kernels(hloc, 20, C.byref(k1), C.byref(k2), C.byref(k3)))
            m_, n_, nnz_loc, _ = A.local.query()
            b_spmv = spmv_bytes(nloc, nnz_loc)
            ach = b_spmv / (k1.value * 1e-3) / 1e9
            traffic = None
            tp = os.path.join(REPO, "profiles", "spmv_traffic.json")
            if os.path.exists(tp):
                try:
                    traffic = json.load(open(tp)).get("dram_bytes_per_launch")
                except Exception:
                    traffic = None
            info = C.c_double(0)
            lib.ab_csr_get_info(hloc, b"ncode", C.byref(info)); ncode = int(info.value)
            lib.ab_csr_get_info(hloc, b"noff", C.byref(info)); noff = int(info.value)
            lib.ab_csr_get_info(hloc, b"npat", C.byref(info)); npat = int(info.value)
            # bytes the kernel that actually ran has to move: index+value bytes per stored entry differ by form
            per_entry = 0 if npat > 0 else 1 if ncode > 0 else (9 if noff > 0 else 12)
            kname = ("spmv_rowpat_kernel<256> (1 byte per ROW names its {offset,value} pattern)" if npat > 0 else
                     "spmv_code_kernel<256,8> (1-byte {offset,value} pattern codes)" if ncode > 0 else
                     "spmv_tile8_kernel<256,8> (1-byte offset dictionary + fp64 values)" if noff > 0 else
                     "spmv_tile_kernel<256,8> (int32 colind + fp64 values)")
            streamed = (17.0 * nloc if npat > 0 else per_entry * nnz_loc + 4.0 * (nloc + 1) + 16.0 * nloc)
            if traffic is not None and json.load(open(tp)).get("bytes_per_entry") != per_entry:
                traffic = None          # the committed ncu capture is of a different kernel form
            roof = {"bound": "hbm", "kernel": kname + " + fused <p,Ap>,<Ap,Ap>", "achieved": ach, "peak": peak,
                    "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": b_spmv, "ms_per_launch": k1.value,
                    "streamed_bytes_per_launch": streamed, "streamed_GBps": streamed / (k1.value * 1e-3) / 1e9,
                    "note": "achieved = SURVEY §8d algorithmic bytes (12 B/entry CSR) / measured time; this kernel form "
                            "streams %d B/entry (lossless dictionary), so frac > 1 is traffic removed, not work skipped" % per_entry}
            kernels = {
                "spmv_dot": {"ms": k1.value, "GBps": ach},
                # symmetrically scaled CG with the x update riding in K3: K2 streams 3 vectors
                # (Ap, r in; r out), K3 streams 5 (r, p, x in; p, x out)
                "cg_update": {"ms": k2.value, "GBps": 3 * 8.0 * nloc / (k2.value * 1e-3) / 1e9},
                "cg_direction": {"ms": k3.value, "GBps": 5 * 8.0 * nloc / (k3.value * 1e-3) / 1e9},
            }
            x = torch.rand(N, dtype=torch.float64, device=dev)
            y = torch.empty(nloc, dtype=torch.float64, device=dev)
            t = C.c_double(0)
            pkg.check(lib.ab_bench_spmv(hloc, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), 5, 0, C.byref(t)))
            pkg.check(lib.ab_bench_spmv(hloc, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), 20, 0, C.byref(t)))
            spmv_info = {"ms": t.value, "GBps": b_spmv / (t.value * 1e-3) / 1e9, "frac_of_peak": b_spmv / (t.value * 1e-3) / 1e9 / peak}
            del x, y
    iter_bytes = cg_iter_bytes(N, nnz)
    if world == 1:
        exchange = "none (single GPU)"
    else:
        act = C.c_double(0)
        lib.ab_get_option(b"dist_p2p_active", C.byref(act))
        exchange = ("peer-memory kernels over NVLink (CUDA IPC): halo push + 2-value all-reduce, no NCCL in the iteration"
                    if act.value else "NCCL send/recv halos + ncclAllReduce")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(n, world, ITERS_PER_STEP),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * N, "d2h_bytes_per_step": 8 * N,
                "ms_per_step": ms_e2e / args.steps,
                "call": "ab_dist_cg_fixed(dh, f_host_pinned, u_host_pinned, iters) — matrix handle resident, like the reference's factorize/solve reuse (src/sparse.jl:671-699)"},
        "gpu_launches": int(launches), "clocks": clocks,
        "cg_iteration": {"algorithmic_GB": iter_bytes / 1e9, "achieved_GBps": iter_bytes * value / 1e9,
                         "frac_of_hbm_peak": iter_bytes * value / 1e9 / (peak * world), "peak_GBps_per_gpu": peak,
                         "peak_source": peak_src,
                         "note": "algorithmic bytes = SURVEY §8d (12 B/entry CSR + 13 vector passes); the solver streams a "
                                 "lossless dictionary form of the matrix (roofline.streamed_bytes_per_launch) and 10 vector "
                                 "passes (scaled PCG, x update fused into the direction kernel), so fractions above 1 are "
                                 "traffic removed, not skipped work"},
        "relres_after_step": relres_dev, "true_relres_after_step": err,
        "exchange": exchange,
    }
    if roof:
        line["roofline"] = roof
        line["kernels"] = kernels
        line["spmv"] = spmv_info
    if rank == 0 and world == 1 and not args.no_cpu:
        cb = cpu_cg_sample(n, 4, reps=2, warm=1)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        print(json.dumps(line), flush=True)
    A.close()
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        pkg.mpi_finalize()
        dist.destroy_process_group()


def cpu_config_samples(path: str):
    """cpu_baseline leg for BASELINE.json's secondary configs (C2 assembly + SpMV, C4 direct solve, C5
    SpMM): the oracle / scipy timed on this box's host cores on BOUNDED samples of the same workloads.
    Single-threaded like the reference ops (Eigen setFromTriplets / SparseLU and the TF sparse matmul
    CPU functor do not shard).  Writes one JSON object; rank 0 only, no GPU needed."""
    import numpy as np
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    import __graft_entry__ as ge

    orc = ge.load_oracle()
    out = {"cores": 1, "kind": "port", "host_threads_available": os.cpu_count()}

    def best(fn, reps=2):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
        return min(ts), r

    n = 96                                                     # C2 sample: 96^3 face-wise stream
    ii, jj, vv = orc.facewise_triplets_3d(n, seed=233)
    N = n ** 3
    t_asm, (rp, ci, va, _) = best(lambda: orc.coo_to_csr(ii, jj, vv, N, N, base=1))
    x = np.random.default_rng(233).uniform(-1, 1, N)
    t_spmv, _ = best(lambda: orc.csr_spmm(rp, ci, va, x), reps=3)
    out["c2"] = {"sample": f"{n}^3 face-wise stream, {len(vv)} triplets -> nnz {len(va)} (config: 256^3)",
                 "assembly_s": t_asm, "assembly_Mtriplets_per_s": len(vv) / t_asm / 1e6,
                 "spmv_s": t_spmv, "spmv_GBps": (12.0 * len(va) + 4.0 * (N + 1) + 16.0 * N) / t_spmv / 1e9}
    m = 512                                                    # C4 sample: the reference's direct solve on 512^2
    h_ = 1.0 / (m + 1)
    xs = np.arange(m + 2) * h_
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    kext = 1 + X ** 2 + Y ** 2
    cols, val, ncols = orc.poisson2d_matrix(m, kext)
    rowptr = np.concatenate([[0], np.cumsum(ncols)])
    A = -sp.csr_matrix((val, cols, rowptr), shape=(m * m, m * m))
    f = np.random.default_rng(233).uniform(size=m * m)
    t_fac, lu = best(lambda: spla.splu(A.tocsc()), reps=1)
    t_sol, u = best(lambda: lu.solve(f), reps=2)
    t_adj, _ = best(lambda: spla.splu(A.T.tocsc()).solve(u), reps=1)     # SparseSolverGrad refactorises A' (SparseSolver.h:46-70)
    out["c4"] = {"sample": f"{m}^2 variable-coefficient 5-point system, SuperLU direct solve (config: 4096^2 = 64x the unknowns)",
                 "factorize_s": t_fac, "solve_s": t_sol, "adjoint_factorize_and_solve_s": t_adj,
                 "relres": float(np.linalg.norm(A @ u - f) / np.linalg.norm(f))}
    rows, k = 1_000_000, 32                                    # C5 sample: 1 M power-law rows
    rng = np.random.default_rng(233)
    deg = np.clip(np.floor(np.clip(rng.uniform(size=rows), 1e-12, None) ** (-1.0 / 1.1)), 1, 1e5)
    deg = np.clip(np.round(deg * (16 / deg.mean())), 1, 1e5).astype(np.int64)
    rp5 = np.concatenate([[0], np.cumsum(deg)])
    nnz5 = int(rp5[-1])
    idx2 = np.stack([np.repeat(np.arange(rows), deg), rng.integers(0, rows, nnz5)], axis=1).astype(np.int64)
    v5 = rng.uniform(-1, 1, nnz5)
    X5 = rng.uniform(-1, 1, (rows, k))
    t_spmm, Y5 = best(lambda: orc.spmm_coo(idx2, v5, rows, X5), reps=2)
    t_sddmm, _ = best(lambda: orc.spmm_grad_values(idx2, Y5, X5), reps=2)
    out["c5"] = {"sample": f"{rows} power-law rows, nnz {nnz5}, k = {k} (config: 50 M rows)",
                 "spmm_s": t_spmm, "spmm_Mnnz_per_s": nnz5 / t_spmm / 1e6,
                 "value_grad_s": t_sddmm, "value_grad_Mnnz_per_s": nnz5 / t_sddmm / 1e6}
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=512, help="grid edge n of the n^3 Poisson system (BASELINE: 512)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-configs", metavar="OUT.json", default=None,
                    help="cpu_baseline leg of the secondary configs (C2/C4/C5 bounded samples); writes OUT.json and exits")
    args = ap.parse_args()
    if args.cpu_configs:
        if int(os.environ.get("RANK", 0)) == 0:
            cpu_config_samples(args.cpu_configs)
        return
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3          # timing rules: W >= 3
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            raise SystemExit(f"--gpus {args.gpus} needs `python -m torch.distributed.run --nproc-per-node {args.gpus} bench.py ...`")
        args.gpus = world
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
