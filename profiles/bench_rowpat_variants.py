"""A/B of the solver SpMV forms on the 512^3 Poisson system: per-kernel CUDA-event times of the CG
iteration (ab_bench_cg_kernels) for the row-pattern kernel at several run lengths, the per-entry
pattern-code kernel, the offset-dictionary kernel and the plain int32 kernel.
    python profiles/bench_rowpat_variants.py [grid=512]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402,F401

import __graft_entry__ as ge  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
pkg = ge.load_package()
lib = pkg.lib
h = C.c_int64(0)
pkg.check(lib.ab_poisson3d_csr(n, n, n, 0, n ** 3, C.byref(h)))
out = {}


def run(tag, **opts):
    for k, v in opts.items():
        pkg.check(lib.ab_set_option(k.encode(), float(v)))
    k1, k2, k3 = C.c_double(0), C.c_double(0), C.c_double(0)
    pkg.check(lib.ab_bench_cg_kernels(h.value, 5, C.byref(k1), C.byref(k2), C.byref(k3)))
    pkg.check(lib.ab_bench_cg_kernels(h.value, 30, C.byref(k1), C.byref(k2), C.byref(k3)))
    out[tag] = {"spmv_ms": round(k1.value, 4), "update_ms": round(k2.value, 4), "direction_ms": round(k3.value, 4),
                "iter_ms": round(k1.value + k2.value + k3.value, 4)}
    print(tag, out[tag], flush=True)


for r in (1, 2, 4, 8):
    run(f"rowpat_run{r}", spmv_rowpat=1, spmv_code=1, spmv_off8=1, spmv_rowpat_run=r)
run("code", spmv_rowpat=0, spmv_code=1, spmv_off8=1)
run("off8", spmv_rowpat=0, spmv_code=0, spmv_off8=1)
run("int32", spmv_rowpat=0, spmv_code=0, spmv_off8=0)
run("rowpat_nodefer", spmv_rowpat=1, spmv_code=1, spmv_off8=1, spmv_rowpat_run=4, cg_defer_x=0)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "rowpat_variants.json"), "w"), indent=1)
