"""ADCME.options, restricted to the fields the sparse hot path reads
(/root/reference/src/options.jl:9-13,14-44,54-58,70-88)."""
from __future__ import annotations

from dataclasses import dataclass, field


@dataclass
class OptionsSparse:
    auto_reorder: bool = True   # options.jl:12 — find() returns row-major sorted triplets
    solver: str = "SparseLU"    # options.jl:11 — mapped: LLT/LDLT -> CG; SparseLU/SparseQR/auto -> CG first when the matrix is bitwise symmetric with a positive diagonal, else BiCGSTAB (+ dense partial-pivot LU on the device as the last resort for n <= 8192)
    tol: float = 1e-12          # new: relative residual target of the iterative solvers
    maxiter: int = 0            # new: 0 = library default (10 n + 100, capped)


@dataclass
class OptionsNewtonRaphsonLineSearch:
    """options.jl:20-28"""
    c1: float = 1e-4
    rho_hi: float = 0.5
    rho_lo: float = 0.1
    iterations: int = 1000
    maxstep: int = 9999999
    alpha_initial: float = 1.0


@dataclass
class OptionsNewtonRaphson:
    """options.jl:35-44"""
    max_iter: int = 100
    verbose: bool = False
    rtol: float = 1e-12
    tol: float = 1e-12
    LM: float = 0.0
    linesearch: bool = False
    linesearch_options: OptionsNewtonRaphsonLineSearch = field(default_factory=OptionsNewtonRaphsonLineSearch)


@dataclass
class OptionsMPI:
    solver: str = "BoomerAMG"   # options.jl:56 — BoomerAMG/auto -> CG on symmetric positive-diagonal systems, BiCGSTAB otherwise; GMRES -> BiCGSTAB
    printlevel: int = 2         # options.jl:57 — kept for signature compatibility


@dataclass
class Options:
    sparse: OptionsSparse = field(default_factory=OptionsSparse)
    newton_raphson: OptionsNewtonRaphson = field(default_factory=OptionsNewtonRaphson)
    mpi: OptionsMPI = field(default_factory=OptionsMPI)


options = Options()


def reset_default_options() -> None:
    """ADCME.reset_default_options (options.jl:80-88)."""
    global options
    options.sparse = OptionsSparse()
    options.newton_raphson = OptionsNewtonRaphson()
    options.mpi = OptionsMPI()
