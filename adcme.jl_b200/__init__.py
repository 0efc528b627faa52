"""adcme.jl_b200 — B200-native drop-in for ADCME.jl's sparse linear-algebra custom-op hot path.

The product is libadcme_b200.so (csrc/*.cu behind include/adcme_b200.h).  This package is the
host-side mirror of the reference's Julia API for that path (src/sparse.jl, src/extra.jl op
loading, src/mpi.jl mpi_SparseTensor) used by tests and benchmarks; julia/ADCMEB200.jl is the
ccall shim for ADCME itself.

The directory name contains a dot, so import it through `__graft_entry__.load_package()`
(importlib with an explicit module name) rather than an `import` statement.
"""
from ._lib import ADCMEError, LIB_PATH, SIGNATURES, lib, check   # noqa: F401  (raises if the .so is missing)
from .options import options, reset_default_options             # noqa: F401
from .sparse import (                                           # noqa: F401
    SparseTensor, SparseAssembler, accumulate, assemble, backslash, compress, compress_grad,
    compress_indices, factorize, find, rows, cols, values, mul, rmul, solve,
)
from .structure import (                                        # noqa: F401
    zero_out_row, scatter_update, scatter_add, getindex, add, matmul, spdiag, spzero, hcat, vcat, hvcat, to_dense, dense_to_sparse, spsum, trisolve,
)
from .optim import NRResult, backtracking, newton_raphson, newton_raphson_with_grad   # noqa: F401
from .extra import (                                            # noqa: F401
    load_op, load_op_and_grad, load_system_op, get_library_symbols, libadcme,
)
from .mpi import mpi_init, mpi_finalize, mpi_rank, mpi_size, mpi_SparseTensor, row_partition  # noqa: F401

__all__ = [
    "ADCMEError", "SparseTensor", "SparseAssembler", "accumulate", "assemble", "backslash", "compress",
    "factorize", "find", "rows", "cols", "values", "solve", "load_op", "load_op_and_grad", "load_system_op",
    "mpi_init", "mpi_finalize", "mpi_SparseTensor", "options", "zero_out_row", "scatter_update", "scatter_add",
    "getindex", "spdiag", "spzero", "newton_raphson", "newton_raphson_with_grad", "NRResult",
]
