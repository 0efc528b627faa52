"""The C-ABI library loads on a CPU-only box, exports every symbol include/adcme_b200.h declares
(and nothing torch-typed), and fails loudly — no CPU fallback — when asked to compute without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(REPO, "include", "adcme_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ab_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_all_exported(pkg):
    declared = _declared_symbols()
    assert len(declared) >= 50
    out = subprocess.run(["nm", "-D", "--defined-only", pkg.LIB_PATH], check=True, capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    # the Python binding covers exactly the header
    assert sorted(pkg.SIGNATURES) == declared
    assert sorted(pkg.get_library_symbols()) and set(declared) <= set(pkg.get_library_symbols())


def test_library_has_sm100a_code_and_tma(pkg):
    out = subprocess.run(["cuobjdump", "-lelf", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass, "the SpMV tile kernel must stage its slices with the TMA bulk-copy unit"


def test_no_cpu_fallback_without_gpu(pkg):
    if pkg.lib.ab_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_int64(0)
    ii = np.array([1, 2], np.int64); vv = np.array([1.0, 2.0])
    rc = pkg.lib.ab_csr_from_coo(2, ii.ctypes.data, ii.ctypes.data, vv.ctypes.data, 2, 2, 1, C.byref(h))
    assert rc == -3 and b"no CPU fallback" in pkg.lib.ab_last_error()
    with pytest.raises(pkg.ADCMEError):
        pkg.SparseTensor(ii, ii, vv, 2, 2).csr()
    rc = pkg.lib.ab_poisson3d_csr(4, 4, 4, 0, 64, C.byref(h))
    assert rc == -3


def test_pure_host_entry_points(pkg):
    assert pkg.lib.ab_version().startswith(b"adcme_b200")
    assert pkg.lib.ab_set_option(b"tol", 1e-9) == 0
    v = C.c_double(0)
    assert pkg.lib.ab_get_option(b"tol", C.byref(v)) == 0 and v.value == 1e-9
    assert pkg.lib.ab_get_option(b"nope", C.byref(v)) == -1
    assert pkg.lib.ab_set_option(b"tol", 1e-12) == 0
    assert pkg.lib.ab_timer_reset() == 0 and pkg.lib.ab_timer_get(0) == 0.0 and pkg.lib.ab_timer_get(9) == -1.0
    assert pkg.lib.ab_csr_query(12345, None, None, None, None) == -2
    # assembler bookkeeping is host-side until the dump
    assert pkg.lib.ab_asm_create(7, 3, 0.5) == 0
    cols = np.array([1, 2, 3], np.int32); vals = np.array([2.0, 0.5, -0.7])
    assert pkg.lib.ab_asm_add(7, 2, cols.ctypes.data, vals.ctypes.data, 3) == 0
    assert pkg.lib.ab_asm_nnz(7) == 2                      # 0.5 is not > tol
    assert pkg.lib.ab_asm_add(7, 4, cols.ctypes.data, vals.ctypes.data, 3) == -4   # row out of bounds
    assert pkg.lib.ab_asm_nnz(7) == 2
    assert pkg.lib.ab_asm_create(7, 99, 0.0) == 0 and pkg.lib.ab_asm_nnz(7) == 2   # existing handle returned
    assert pkg.lib.ab_asm_destroy(7) == 0 and pkg.lib.ab_asm_nnz(7) == -2


def test_op_registry_mirrors_reference_names(pkg):
    for name in ("sparse_solver", "sparse_compress", "sparse_accumulator", "sparse_accumulator_add",
                 "sparse_accumulator_copy", "sparse_factorization", "solve", "mpi_create_matrix"):
        op = pkg.load_system_op(name)
        assert callable(op)
    assert pkg.load_op_and_grad(pkg.libadcme, "sparse_solver").grad is not None
    assert pkg.load_op(pkg.libadcme, "sparse_solver").grad is None
    with pytest.raises(KeyError):
        pkg.load_system_op("sinkhorn_knopp")
    assert pkg.options.sparse.solver == "SparseLU" and pkg.options.sparse.auto_reorder is True
    assert pkg.options.mpi.solver == "BoomerAMG"


# ---------------------------------------------------------------------------------------------
# The three descriptions of the boundary must agree argument by argument: the header prototypes,
# the ctypes table the tests bind with, and the ccall signatures of the Julia shim a maintainer adds.
def _header_prototypes():
    txt = open(os.path.join(REPO, "include", "adcme_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"\b(int64_t|int|double|const char\*)\s+(ab_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        args = " ".join(args.split())
        kinds = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    kinds.append("str" if re.match(r"const char\s*\*", a) else "ptr")
                elif re.match(r"(u?int64_t)\b", a):
                    kinds.append("i64")
                elif re.match(r"(u?int32_t)\b", a):
                    kinds.append("i32")
                elif re.match(r"double\b", a):
                    kinds.append("dbl")
                elif re.match(r"(unsigned\s+)?int\b", a):
                    kinds.append("int")
                else:
                    raise AssertionError(f"unclassified parameter '{a}' of {name}")
        protos[name] = (ret, kinds)
    return protos


def test_ctypes_table_matches_header_prototypes(pkg):
    protos = _header_prototypes()
    assert set(protos) == set(pkg.SIGNATURES)
    cmap = {C.c_int64: "i64", C.c_int32: "i32", C.c_int: "int", C.c_double: "dbl", C.c_void_p: "ptr", C.c_char_p: "str"}
    rmap = {"int": C.c_int, "int64_t": C.c_int64, "double": C.c_double, "const char*": C.c_char_p}
    for name, (ret, kinds) in protos.items():
        res, args = pkg.SIGNATURES[name]
        assert res is rmap[ret], name
        got = [cmap[a] for a in args]
        # int32_t and int are both 4-byte signed here; ctypes declares either
        norm = lambda ks: ["int" if k == "i32" else k for k in ks]          # noqa: E731
        assert norm(got) == norm(kinds), f"{name}: ctypes {got} vs header {kinds}"


def test_julia_shim_ccalls_match_header_prototypes():
    protos = _header_prototypes()
    src = open(os.path.join(REPO, "julia", "ADCMEB200.jl")).read()
    calls = re.findall(r"ccall\(\(:(ab_[a-z0-9_]+),\s*LIB\),\s*([A-Za-z0-9_{}]+),\s*\(([^)]*)\)", src)
    assert len(calls) >= 20
    jl = {"Int64": "i64", "Int32": "int", "Cint": "int", "Cdouble": "dbl", "Cstring": "str"}
    rj = {"Cint": "int", "Int64": "int64_t", "Cdouble": "double", "Cstring": "const char*"}
    seen = set()
    for name, ret, args in calls:
        assert name in protos, f"{name} is not declared in the header"
        seen.add(name)
        hret, kinds = protos[name]
        assert rj[ret] == hret, f"{name}: Julia returns {ret}, header {hret}"
        toks = [t.strip() for t in args.split(",") if t.strip()]
        got = ["ptr" if t.startswith(("Ptr{", "Ref{")) else jl[t] for t in toks]
        norm = lambda ks: ["int" if k == "i32" else k for k in ks]          # noqa: E731
        assert got == norm(kinds), f"{name}: Julia {got} vs header {kinds}"
    # the shim covers the hot path and the widened rows
    for must in ("ab_csr_from_coo", "ab_spmm", "ab_solve", "ab_solve_grad", "ab_asm_copy", "ab_compress", "ab_factorize",
                 "ab_zero_out_row", "ab_scatter_update", "ab_sparse_indexing", "ab_spgemm", "ab_csr_axpby", "ab_newton_step"):
        assert must in seen, must
