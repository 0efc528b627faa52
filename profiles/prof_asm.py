"""One COO->CSR assembly of the face-wise 3-D 7-pt stream (C2) for ncu launch lists.

    python profiles/prof_asm.py [grid=256]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402

import bench_configs as bc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ii, jj, vv = bc.facewise_3d_device(n)
torch.cuda.synchronize()
h = C.c_int64(0)
bc.pkg.check(bc.lib.ab_csr_from_coo(ii.numel(), ii.data_ptr(), jj.data_ptr(), vv.data_ptr(), n ** 3, n ** 3, 1, C.byref(h)))
torch.cuda.synchronize()
print("T", ii.numel(), "nnz", bc.pkg.SparseTensor.from_handle(h.value).query()[2], "assembly s", bc.lib.ab_timer_get(2))
