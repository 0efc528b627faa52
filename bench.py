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
  extra keys of the JSON line (all measured inside this run):
    roofline / kernels / cg_iteration   per-kernel CUDA-event times; frac = bytes the running kernel form
                     streams / time / measured peak; the SURVEY §8d (plain CSR) normalisation is reported
                     separately as vs_plain_csr_roofline / vs_section8d_roofline
    general_csr      the same CG on the variable-coefficient 7-point operator of the same grid (no value
                     dictionary applies: the general CSR kernels run), with its own roofline block
    parity           adcme.jl_b200/selfcheck.py run BEFORE timing (N > 1: every multi-rank data path vs the
                     single-GPU solver / scipy / a matrix-free stencil); the run aborts when it fails
    configs          N = 1: C2 assembly + SpMV grads, C4 solve + adjoint, C5 SpMM + value gradient
    cpu_baseline     oracle port on all host threads + single-thread figure
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
def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_cg_sample(n_req: int, iters: int, reps: int = 1, warm: int = 0, single_thread_iters: int = 2):
    """Times the oracle's Jacobi-PCG port on the n^3 system with ALL host threads (set explicitly: torchrun
    exports OMP_NUM_THREADS=1) and, on a shorter sample, with ONE thread (the reference's ops are
    single-threaded); falls back to a 256^3 sample (rescaled by rows) when host memory is short."""
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
    nthr = orc.set_num_threads(host_threads())
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
    single = None
    if single_thread_iters > 0:
        orc.set_num_threads(1)
        t0 = time.perf_counter()
        orc.pcg_jacobi(rp, ci, va, b, tol=-1.0, maxit=single_thread_iters)
        single = single_thread_iters / (time.perf_counter() - t0) * scale
        orc.set_num_threads(nthr)
    sample = f"{iters} Jacobi-PCG iterations x {len(times)} on the {n}^3 system, {nthr} OpenMP threads (set explicitly)"
    if single is not None:
        sample += f"; single-thread figure from {single_thread_iters} iterations"
    if scale != 1.0:
        sample += f" (host RAM {avail_gb:.0f} GB < needed; iterations/s rescaled by rows {n}^3/{n_req}^3)"
    return {"value": its, "unit": UNIT, "cores": nthr, "kind": "port", "sample": sample, "single_thread_value": single,
            "times": times, "relres": rel, "n": n}


CPU_KEYS = ("value", "unit", "cores", "kind", "sample", "single_thread_value")


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
        "config": workload_config(n, args.gpus), "iters_per_step": REF_ITERS_PER_STEP,
        "cpu_baseline": {k: res[k] for k in CPU_KEYS},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference arm = oracle port of the same Jacobi-PCG recurrence on ALL host threads (value) and on one "
                "(cpu_baseline.single_thread_value: the reference's ops are single-threaded); the reference itself has "
                "no CG and its Eigen SparseLU `\\` is infeasible at 512^3 (SURVEY F2-F3)",
    }
    print(json.dumps(line), flush=True)


def workload_config(n: int, gpus: int):
    return {
        "workload": f"Jacobi-PCG on 3-D 7-point Poisson {n}^3 fp64 (BASELINE.json configs[2]), b = A*1, x0 = 0",
        "rows": n ** 3, "nnz": poisson_nnz(n),
        "partition": f"z-slab row blocks over {gpus} GPU(s)" if gpus > 1 else "single GPU",
        "l2": "inputs larger than L2 (one CG iteration streams >= 10.9 GB at 512^3; L2 is 126 MB)",
    }


# ----------------------------------------------------------------------------- GPU arm
class near_gpu_affinity:
    """Context manager: run the block with the CPU affinity NVML recommends for the GPU (its NUMA node), then restore
    the previous affinity (the CPU baseline leg wants all cores back).  Silently a no-op when NVML is unavailable."""

    def __init__(self, index: int):
        self.index, self.saved = index, None

    def __enter__(self):
        try:
            import pynvml
            import torch

            self.saved = os.sched_getaffinity(0)
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(self.index)
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
            cpus &= self.saved
            if cpus:
                os.sched_setaffinity(0, cpus)
        except Exception:
            pass
        return self

    def __exit__(self, *exc):
        if self.saved:
            try:
                os.sched_setaffinity(0, self.saved)
            except Exception:
                pass
        return False


def spmv_form(lib, h, nloc: int, nnz_loc: int):
    """Which SpMV kernel the CG solver runs for this handle and the bytes that form has to move per launch."""
    info = C.c_double(0)

    def q(key):
        lib.ab_csr_get_info(h, key, C.byref(info))
        return int(info.value)

    ncode, noff, npat, rowwin = q(b"ncode"), q(b"noff"), q(b"npat"), q(b"rowwin")
    if npat > 0 and rowwin == 1 and q(b"zmarch") == 1:
        return {"kernel": f"spmv_zmarch_kernel<G={q(b'zmarch_g')}> (1 byte per ROW names its {{offset,value}} pattern; x windows of {q(b'zmarch_g')} x 1024 rows + "
                          "margins streamed by TMA through a 4-slot mbarrier ring, planes -M/+M carried in registers along z-chains)",
                "index_value_bytes": "1 per row", "streamed_bytes": 17.0 * nloc}
    if npat > 0 and rowwin == 1 and q(b"rowmarch") == 1:
        return {"kernel": "spmv_rowmarch_kernel (1 byte per ROW names its {offset,value} pattern; ring of x windows marched along z-chains, staged by TMA)",
                "index_value_bytes": "1 per row", "streamed_bytes": 17.0 * nloc}
    if npat > 0 and rowwin == 1:
        return {"kernel": "spmv_rowwin_kernel (1 byte per ROW names its {offset,value} pattern; x intervals staged by TMA)",
                "index_value_bytes": "1 per row", "streamed_bytes": 17.0 * nloc}
    if npat > 0:
        return {"kernel": "spmv_rowpat_kernel<256> (1 byte per ROW names its {offset,value} pattern)",
                "index_value_bytes": "1 per row", "streamed_bytes": 17.0 * nloc}
    per = 1 if ncode > 0 else (9 if noff > 0 else 12)
    name = ("spmv_code_kernel<256,8> (1-byte {offset,value} pattern codes)" if ncode > 0 else
            "spmv_tile8_kernel<256,8> (1-byte offset dictionary + fp64 values)" if noff > 0 else
            "spmv_tile_kernel<256,8> (int32 colind + fp64 values)")
    return {"kernel": name, "index_value_bytes": f"{per} per stored entry", "streamed_bytes": per * nnz_loc + 4.0 * (nloc + 1) + 16.0 * nloc}


def cg_leg(pkg, A, args, rank, world, dev, e2e: bool, sample_clocks: bool, constant_diagonal: bool = True):
    """Times `steps` calls of ab_dist_cg_fixed (ITERS_PER_STEP iterations each) on the distributed handle A:
    device-resident f/u, and (e2e) pinned-host f/u with both copies inside the timed region."""
    import torch

    lib = pkg.lib
    nloc = A.nloc
    ones = torch.ones(nloc, dtype=torch.float64, device=dev)
    b = A @ ones                                    # device-resident right-hand side
    u = torch.empty_like(b)
    rel = C.c_double(0)

    def step(dst, src):
        pkg.check(lib.ab_dist_cg_fixed(A._dh, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()),
                                       ITERS_PER_STEP, C.byref(rel)))

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def timed(dst, src, clocks_on):
        for _ in range(args.warmup):
            step(dst, src)
        barrier()
        sampler = ClockSampler(dev.index) if (clocks_on and rank == 0) else None
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
        return max_over_ranks(ms, dev), launches, clocks

    ms_dev, launches, clocks = timed(u, b, sample_clocks)
    relres_dev = rel.value
    # sanity guard inside the bench: the timed iterations did real work — the TRUE residual of the returned iterate
    # (computed with the plain A @ u product, not the solver's compressed form) matches the recurrence residual the
    # solver reports, and it has dropped
    true_res = float(((A @ u) - b).norm().item())
    if world > 1:
        true_res = sum_over_ranks(true_res ** 2, dev) ** 0.5
        bnorm = sum_over_ranks(float(b.norm().item()) ** 2, dev) ** 0.5
    else:
        bnorm = float(b.norm().item())
    err = true_res / bnorm
    # (constant diagonal: the solver's scaled-norm residual IS the true one; variable coefficients: same order of magnitude)
    same = (abs(err - relres_dev) <= 1e-5 * max(err, 1e-300) + 1e-9) if constant_diagonal else (0.05 * relres_dev <= err <= 20 * relres_dev)
    if not (err < 0.9 and same):
        raise RuntimeError(f"bench sanity check failed: true relres {err} vs reported {relres_dev}")
    out = {"ms_dev": ms_dev, "launches": int(launches), "clocks": clocks, "relres": relres_dev, "true_relres": err,
           "value": ITERS_PER_STEP * args.steps / (ms_dev * 1e-3)}
    if e2e:
        # host buffers on the NUMA node next to this rank's GPU: with 8 ranks staging at once, buffers that all sit on
        # one socket make every copy cross the inter-socket link (measured N=8 e2e before: 15 GB/s per rank)
        with near_gpu_affinity(dev.index):
            b_host = torch.empty(nloc, dtype=torch.float64).pin_memory()
            u_host = torch.empty(nloc, dtype=torch.float64).pin_memory()
            b_host.fill_(0.0)
            u_host.fill_(0.0)                       # first touch under the node-local affinity
        b_host.copy_(b)
        torch.cuda.synchronize()
        ms_e2e, _, _ = timed(u_host, b_host, False)
        out["ms_e2e"] = ms_e2e
        out["e2e_value"] = ITERS_PER_STEP * args.steps / (ms_e2e * 1e-3)
        del b_host, u_host
    del b, u, ones
    return out


def kernel_rooflines(pkg, hloc, nloc, nnz_loc, N, nnz, value, world, peak, peak_src):
    """Per-kernel CUDA-event durations of one CG iteration on this handle (ab_bench_cg_kernels) and the roofline
    blocks derived from them.  frac = bytes the running kernel form has to move / time / measured peak."""
    lib = pkg.lib
    k1, k2, k3 = C.c_double(0), C.c_double(0), C.c_double(0)
    pkg.check(lib.ab_bench_cg_kernels(hloc, 20, C.byref(k1), C.byref(k2), C.byref(k3)))
    form = spmv_form(lib, hloc, nloc, nnz_loc)
    info = C.c_double(0)
    lib.ab_csr_get_info(hloc, b"scaled", C.byref(info))
    scaled = int(info.value) == 1
    streamed = form["streamed_bytes"]
    b_alg = spmv_bytes(nloc, nnz_loc)
    ach = streamed / (k1.value * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": form["kernel"] + " + fused <p,Ap>, <Ap,Ap>", "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": ach / peak, "traffic": None, "peak_source": peak_src, "ms_per_launch": k1.value,
            "bytes_per_launch": streamed, "bytes_model": "what this kernel form must move: " + form["index_value_bytes"] +
            " of index/value data + 8 B/row x + 8 B/row y (each vector element once)",
            "section8d_algorithmic_bytes_per_launch": b_alg,
            "vs_plain_csr_roofline": b_alg / (k1.value * 1e-3) / 1e9 / peak,
            "note": "frac uses the bytes the running kernel streams; vs_plain_csr_roofline divides SURVEY §8d's plain-CSR bytes "
                    "(12 B/entry) by the same time — above 1 there means traffic removed by the lossless dictionary form, not "
                    "a bandwidth claim.  traffic: not measured in-run (ncu captures of this kernel are under profiles/)"}
    k2_bytes = (3 if scaled else 7) * 8.0 * nloc
    k3_bytes = (5 if scaled else 4) * 8.0 * nloc
    kernels = {
        "spmv_dot": {"ms": k1.value, "bytes": streamed, "GBps": ach, "frac": ach / peak},
        "cg_update": {"ms": k2.value, "bytes": k2_bytes, "GBps": k2_bytes / (k2.value * 1e-3) / 1e9,
                      "frac": k2_bytes / (k2.value * 1e-3) / 1e9 / peak},
        "cg_direction": {"ms": k3.value, "bytes": k3_bytes, "GBps": k3_bytes / (k3.value * 1e-3) / 1e9,
                         "frac": k3_bytes / (k3.value * 1e-3) / 1e9 / peak},
    }
    # whole iteration: bytes streamed per iteration over the whole system, at the measured iterations/s
    per_row_vec = (8 if scaled else 11) * 8.0
    it_streamed = streamed / nloc * N + per_row_vec * N if nloc else 0.0
    it_alg = cg_iter_bytes(N, nnz)
    iteration = {"streamed_GB": it_streamed / 1e9, "achieved_GBps": it_streamed * value / 1e9,
                 "frac": it_streamed * value / 1e9 / (peak * world), "peak_GBps_per_gpu": peak, "peak_source": peak_src,
                 "vector_passes": 10 if scaled else 13,
                 "section8d_algorithmic_GB": it_alg / 1e9, "vs_section8d_roofline": it_alg * value / 1e9 / (peak * world),
                 "note": "frac = bytes the iteration streams (SpMV form above + %d vector passes) x iterations/s / peak; "
                         "vs_section8d_roofline = SURVEY §8d bytes (12 B/entry CSR + 13 vector passes) at the same rate "
                         "(the north-star's >= 0.70 target is on that figure)" % (10 if scaled else 13)}
    return roof, kernels, iteration


def run_gpu(args, rank: int, world: int, local_rank: int):
    import torch
    import __graft_entry__ as ge

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    pkg = ge.load_package()
    lib = pkg.lib
    n = args.grid
    N, nnz = n ** 3, poisson_nnz(n)
    pkg.mpi_init(rank, world, local_rank)
    for env in ("ADCME_B200_OPTS", "AB_OPTS"):                                  # A/B switches, e.g. AB_OPTS=dist_fused_wait=0
        for kv in filter(None, os.environ.get(env, "").split(",")):
            k_, v_ = kv.split("=")
            pkg.check(lib.ab_set_option(k_.strip().encode(), float(v_)))
    peak, peak_src = measured_peak_gbs()

    # ---- parity first: a fast kernel whose results differ is not done (multi-rank paths are only reachable here)
    parity = None
    if not args.no_parity:
        from adcme_jl_b200 import selfcheck

        t0 = time.time()
        parity = selfcheck.dist_parity(pkg, rank, world, dev, n=64 if world > 1 else 40)
        parity["seconds"] = time.time() - t0
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "parity check failed", "parity": parity}), flush=True)
            raise SystemExit("parity check failed: " + ", ".join(parity["failed"]))

    # ---- main leg: constant-coefficient system of BASELINE.json (the solver's row-pattern form applies)
    A = pkg.mpi_SparseTensor.poisson3d(n, n, n, rank, world)
    nloc = A.nloc
    main = cg_leg(pkg, A, args, rank, world, dev, e2e=True, sample_clocks=True)
    value = main["value"]
    roof = kernels = iteration = spmv_info = None
    if rank == 0 and world == 1:
        hloc = A.local.handle()
        _, _, nnz_loc, _ = A.local.query()
        roof, kernels, iteration = kernel_rooflines(pkg, hloc, nloc, nnz_loc, N, nnz, value, world, peak, peak_src)
        x = torch.rand(N, dtype=torch.float64, device=dev)
        y = torch.empty(nloc, dtype=torch.float64, device=dev)
        t = C.c_double(0)
        pkg.check(lib.ab_bench_spmv(hloc, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), 5, 0, C.byref(t)))
        pkg.check(lib.ab_bench_spmv(hloc, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), 20, 0, C.byref(t)))
        info = C.c_double(0)
        lib.ab_csr_get_info(hloc, b"noff", C.byref(info))
        gb = (9.0 if info.value > 0 else 12.0) * nnz_loc + 4.0 * (nloc + 1) + 16.0 * nloc      # int8 offset dictionary or int32 colind
        spmv_info = {"what": "plain A*x (ab_spmm, k = 1) on the UNSCALED values: general-CSR kernel", "ms": t.value,
                     "bytes": gb, "GBps": gb / (t.value * 1e-3) / 1e9, "frac": gb / (t.value * 1e-3) / 1e9 / peak,
                     "vs_plain_csr_roofline": spmv_bytes(nloc, nnz_loc) / (t.value * 1e-3) / 1e9 / peak}
        del x, y
    if world == 1:
        exchange = "none (single GPU)"
    else:
        act = C.c_double(0)
        lib.ab_get_option(b"dist_p2p_active", C.byref(act))
        exchange = ("peer-memory kernels over NVLink (CUDA IPC): halo push + 2-value all-reduce, no NCCL in the iteration"
                    if act.value else "NCCL send/recv halos + ncclAllReduce")
    A.close()
    if A.local is not None:
        A.local.close()
    torch.cuda.empty_cache()

    # ---- general-CSR leg: the SAME grid with variable coefficients (kappa from a seeded field): no value
    # dictionary applies, the solver streams int8 offsets + fp64 values (or int32 colind) — the number any
    # variable-coefficient FEM matrix gets
    general = None
    if not args.no_general:
        try:
            g = torch.Generator(device=dev).manual_seed(233)
            kappa = torch.exp(torch.rand(N, dtype=torch.float64, device=dev, generator=g) * 2 - 1)    # exp(U(-1,1)), same on every rank
            G = pkg.mpi_SparseTensor.poisson3d_kappa(kappa, n, n, n, rank, world)
            del kappa
            gl = cg_leg(pkg, G, args, rank, world, dev, e2e=False, sample_clocks=False, constant_diagonal=False)
            general = {"workload": f"Jacobi-PCG on the variable-coefficient 7-point operator -div(kappa grad u), {n}^3, "
                                   "kappa = exp(U(-1,1)) per cell (seed 233), face weights (kappa_i+kappa_j)/2, b = A*1, x0 = 0",
                       "value": gl["value"], "unit": UNIT, "ms_per_step": gl["ms_dev"] / args.steps, "gpu_launches": gl["launches"],
                       "relres_after_step": gl["relres"], "true_relres_after_step": gl["true_relres"]}
            if rank == 0 and world == 1:
                hg = G.local.handle()
                _, _, nnz_g, _ = G.local.query()
                groof, gk, git = kernel_rooflines(pkg, hg, G.nloc, nnz_g, N, nnz, gl["value"], world, peak, peak_src)
                general.update(roofline=groof, kernels=gk, cg_iteration=git)
            else:
                it_alg = cg_iter_bytes(N, nnz)
                general["cg_iteration"] = {"vs_section8d_roofline": it_alg * gl["value"] / 1e9 / (peak * world)}
            G.close()
            G.local.close()
            torch.cuda.empty_cache()
        except Exception as e:          # the headline must survive a failure of the secondary leg
            general = {"error": repr(e)}

    ms_dev, ms_e2e = main["ms_dev"], main["ms_e2e"]
    it_alg = cg_iter_bytes(N, nnz)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(n, world), "iters_per_step": ITERS_PER_STEP,
        "e2e": {"value": main["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": 8 * N, "d2h_bytes_per_step": 8 * N,
                "ms_per_step": ms_e2e / args.steps,
                "call": "ab_dist_cg_fixed(dh, f_host_pinned, u_host_pinned, iters) — matrix handle resident, like the reference's factorize/solve reuse (src/sparse.jl:671-699)"},
        "gpu_launches": main["launches"], "clocks": main["clocks"],
        "cg_iteration": iteration if iteration else {"vs_section8d_roofline": it_alg * value / 1e9 / (peak * world),
                                                     "section8d_algorithmic_GB": it_alg / 1e9, "peak_GBps_per_gpu": peak},
        "relres_after_step": main["relres"], "true_relres_after_step": main["true_relres"],
        "exchange": exchange,
    }
    if roof:
        line["roofline"] = roof
        line["kernels"] = kernels
        line["spmv"] = spmv_info
    if general is not None:
        line["general_csr"] = general
    if parity is not None:
        line["parity"] = {"ok": parity["ok"], "n_checks": len(parity["checks"]), "grid": parity["grid"], "seconds": parity["seconds"],
                          "max_error": max([v for k, v in parity["checks"].items() if "exact" not in k and "auto_is_cg" not in k] or [0.0]),
                          "checks": parity["checks"]}
    if rank == 0 and world == 1 and not args.no_legs:
        line["configs"] = secondary_configs(args)
    if rank == 0 and world == 1 and not args.no_cpu:
        cb = cpu_cg_sample(n, 4, reps=2, warm=1)
        line["cpu_baseline"] = {k: cb[k] for k in CPU_KEYS}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        pkg.mpi_finalize()
        dist.destroy_process_group()


def secondary_configs(args):
    """BASELINE.json's other configs, driver-run: C2 assembly + SpMV fwd/grad at 256^3, C4 forward solve + adjoint
    gradient at 4096^2, C5 power-law SpMM + value gradient (profiles/bench_configs.py holds the legs; every leg
    verifies its result through a size-independent property before its time counts)."""
    import torch

    out = {}
    try:
        sys.path.insert(0, os.path.join(REPO, "profiles"))
        import bench_configs as bc

        scale = args.legs_scale
        for name, fn, arg in (("c2", bc.c2, int(round(256 * scale))), ("c4", bc.c4, int(round(4096 * scale))),
                              ("c5", bc.c5, int(50_000_000 * scale))):
            torch.cuda.empty_cache()
            t0 = time.time()
            try:
                out[name] = fn(arg)
            except Exception as e:
                out[name] = {"error": repr(e)}
            out[name]["wall_s"] = time.time() - t0
    except Exception as e:
        out["error"] = repr(e)
    return out


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
    ap.add_argument("--no-general", action="store_true", help="skip the variable-coefficient (general CSR) leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity block that runs before timing")
    ap.add_argument("--no-legs", action="store_true", help="skip the C2/C4/C5 legs (N = 1 only)")
    ap.add_argument("--legs-scale", type=float, default=1.0, help="size factor of the C2/C4/C5 legs (1.0 = BASELINE sizes)")
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
