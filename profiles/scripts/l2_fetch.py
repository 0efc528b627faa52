"""Does the L2 fetch-granularity hint (cudaLimitMaxL2FetchGranularity) change what a random 8-byte gather costs?
Times ab_csr_update_values and ab_spmm_grad_values (k = 1) on a shuffled face-wise 7-point stream (default 192^3) and a
short CG run (streaming kernels) under 128 / 64 / 32."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import __graft_entry__ as ge
pkg = ge.load_package(); lib = pkg.lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
dev = torch.device("cuda")
N = n ** 3
idx = torch.arange(N, device=dev, dtype=torch.int64)
i, j, k = idx // (n * n), (idx // n) % n, idx % n
rows, cols, vals = [idx], [idx], [torch.full((N,), 6.0, device=dev, dtype=torch.float64)]
for d, stride, coord in ((1, 1, k), (n, n, j), (n * n, n * n, i)):
    m = coord < n - 1
    a, b = idx[m], idx[m] + stride
    for r, c in ((a, b), (b, a)):
        rows.append(r); cols.append(c); vals.append(torch.full((r.numel(),), -1.0, device=dev, dtype=torch.float64))
ii, jj, vv = torch.cat(rows) + 1, torch.cat(cols) + 1, torch.cat(vals)
g = torch.Generator(device=dev).manual_seed(233)
perm = torch.randperm(ii.numel(), device=dev, generator=g)
ii, jj, vv = ii[perm].contiguous(), jj[perm].contiguous(), vv[perm].contiguous()
T = ii.numel()
h = C.c_int64(0)
pkg.check(lib.ab_csr_from_coo(T, ii.data_ptr(), jj.data_ptr(), vv.data_ptr(), N, N, 1, C.byref(h)))
x = torch.rand(N, device=dev, dtype=torch.float64); dy = torch.rand(N, device=dev, dtype=torch.float64)
dvv = torch.empty(T, device=dev, dtype=torch.float64)
def ev(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
h3 = C.c_int64(0)
pkg.check(lib.ab_poisson3d_csr(256, 256, 256, 0, 256 ** 3, C.byref(h3)))
A3 = pkg.SparseTensor.from_handle(h3.value)
b3 = A3 @ torch.ones(256 ** 3, dtype=torch.float64, device=dev); u3 = torch.empty_like(b3); rel = C.c_double(0)
cur = C.c_double(0); lib.ab_get_option(b"l2_fetch_granularity", C.byref(cur)); print("default granularity", cur.value)
for gran in (128, 64, 32, int(cur.value)):
    pkg.check(lib.ab_set_option(b"l2_fetch_granularity", float(gran)))
    t_upd = ev(lambda: pkg.check(lib.ab_csr_update_values(h.value, vv.data_ptr())))
    t_grad = ev(lambda: pkg.check(lib.ab_spmm_grad_values(h.value, dy.data_ptr(), x.data_ptr(), 1, dvv.data_ptr())))
    t_cg = ev(lambda: pkg.check(lib.ab_cg_fixed(h3.value, b3.data_ptr(), u3.data_ptr(), 50, C.byref(rel))), reps=3) / 50
    print(f"granularity {gran:3d}: T={T/1e6:.1f} M triplets  update_values {t_upd:.3f} ms  grad_values {t_grad:.3f} ms  CG 256^3 {t_cg*1e3:.1f} us/it", flush=True)
