"""ADCME.options, restricted to the fields the sparse hot path reads
(/root/reference/src/options.jl:9-13,54-58,70-88)."""
from __future__ import annotations

from dataclasses import dataclass, field


@dataclass
class OptionsSparse:
    auto_reorder: bool = True   # options.jl:12 — find() returns row-major sorted triplets
    solver: str = "SparseLU"    # options.jl:11 — mapped: SparseLU/SparseQR -> BiCGSTAB, LLT/LDLT -> CG
    tol: float = 1e-12          # new: relative residual target of the iterative solvers
    maxiter: int = 0            # new: 0 = library default (10 n + 100, capped)


@dataclass
class OptionsMPI:
    solver: str = "BoomerAMG"   # options.jl:56 — accepted and mapped to Jacobi-PCG
    printlevel: int = 2         # options.jl:57 — kept for signature compatibility


@dataclass
class Options:
    sparse: OptionsSparse = field(default_factory=OptionsSparse)
    mpi: OptionsMPI = field(default_factory=OptionsMPI)


options = Options()


def reset_default_options() -> None:
    """ADCME.reset_default_options (options.jl:80-88)."""
    global options
    options.sparse = OptionsSparse()
    options.mpi = OptionsMPI()
