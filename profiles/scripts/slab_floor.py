"""Single-GPU floor of one rank's share at N ranks: the 512 x 512 x (512/N) slab as a standalone Dirichlet problem
(no halo, no all-reduce).  Prints CG it/s and per-kernel CUDA-event times: the slab's it/s is what the N-rank job would
reach if exchange and the extra launches cost nothing (every rank iterates on its slab in lock-step); the difference to
the measured N-rank number is halo push, boundary tiles, the two all-reduces and their launch gaps."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import __graft_entry__ as ge
pkg = ge.load_package(); lib = pkg.lib
for kv in filter(None, os.environ.get("AB_OPTS", "").split(",")):
    k_, v_ = kv.split("="); pkg.check(lib.ab_set_option(k_.strip().encode(), float(v_)))
for N in (int(a) for a in (sys.argv[1:] or ["8", "4", "2", "1"])):
    nz = 512 // N
    n = 512 * 512 * nz
    h = C.c_int64(0)
    pkg.check(lib.ab_poisson3d_csr(512, 512, nz, 0, n, C.byref(h)))
    A = pkg.SparseTensor.from_handle(h.value)
    b = A @ torch.ones(n, dtype=torch.float64, device="cuda")
    u = torch.empty_like(b)
    rel = C.c_double(0)
    pkg.check(lib.ab_cg_fixed(h.value, b.data_ptr(), u.data_ptr(), 50, C.byref(rel)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(3):
        pkg.check(lib.ab_cg_fixed(h.value, b.data_ptr(), u.data_ptr(), 200, C.byref(rel)))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 600
    k1, k2, k3 = C.c_double(0), C.c_double(0), C.c_double(0)
    pkg.check(lib.ab_bench_cg_kernels(h.value, 20, C.byref(k1), C.byref(k2), C.byref(k3)))
    info = C.c_double(0); lib.ab_csr_get_info(h.value, b"zmarch_items", C.byref(info))
    print(f"slab 512x512x{nz} (rank share at N={N}): {1e3/ms:9.1f} it/s ({ms*1e3:7.1f} us/it);  kernels us: "
          f"spmv+dots {k1.value*1e3:6.1f} update {k2.value*1e3:6.1f} direction {k3.value*1e3:6.1f} sum {(k1.value+k2.value+k3.value)*1e3:6.1f}; zmarch items {int(info.value)}", flush=True)
    A.close()
