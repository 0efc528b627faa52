# ADCMEB200.jl — ccall shim that re-points ADCME.jl's sparse hot path at libadcme_b200.so.
#
# Drop this file next to /root/reference/src/sparse.jl and `include` it after it: the methods below
# replace the bodies that build TensorFlow graph nodes (PyCall) with direct C-ABI calls, exactly the
# way the reference already calls its own C symbols (src/mpi.jl:16-94, src/pcl.jl:104-114).
# Julia is not available in the build image, so this file is NOT exercised by the tests; the same
# symbols are exercised through the Python/ctypes mirror (adcme.jl_b200/).  Signatures mirror
# include/adcme_b200.h one to one.
module ADCMEB200

using SparseArrays
export B200SparseTensor, b200_spmm, b200_solve, b200_solve_grad, B200Assembler, b200_accumulate!,
       b200_assemble, b200_compress, b200_factorize, b200_fsolve, b200_zero_out_row, b200_scatter_update,
       b200_getindex, b200_spgemm, b200_add, b200_trisolve, b200_newton_step

const LIB = get(ENV, "ADCME_B200_LIB", joinpath(@__DIR__, "..", "adcme.jl_b200", "libadcme_b200.so"))

struct B200Error <: Exception
    code::Cint
    msg::String
end
Base.showerror(io::IO, e::B200Error) = print(io, "adcme_b200 [", e.code, "]: ", e.msg)

@inline function chk(rc::Cint)
    rc == 0 && return nothing
    # the reference raises PyCall.PyError from a TF Status (test/sparse.jl:335-339)
    throw(B200Error(rc, unsafe_string(ccall((:ab_last_error, LIB), Cstring, ()))))
end

init(device::Integer=0) = chk(ccall((:ab_init, LIB), Cint, (Cint,), device))

# ---- SparseTensor(ii, jj, vv, m, n)  — src/sparse.jl:62-76 ---------------------------------------
mutable struct B200SparseTensor
    h::Int64
    m::Int64
    n::Int64
    T::Int64
    function B200SparseTensor(ii::Vector{Int64}, jj::Vector{Int64}, vv::Vector{Float64}, m::Integer, n::Integer)
        h = Ref{Int64}(0)
        chk(ccall((:ab_csr_from_coo, LIB), Cint,
                  (Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Cdouble}, Int64, Int64, Cint, Ref{Int64}),
                  length(vv), ii, jj, vv, m, n, 1, h))
        s = new(h[], m, n, length(vv))
        finalizer(x -> ccall((:ab_csr_destroy, LIB), Cint, (Int64,), x.h), s)
        s
    end
end
B200SparseTensor(A::SparseMatrixCSC) = (t = findnz(A); B200SparseTensor(Int64.(t[1]), Int64.(t[2]), Float64.(t[3]), size(A)...))

"New values on the cached pattern (no re-sort): the factorize-style reuse of src/sparse.jl:671-699."
update_values!(s::B200SparseTensor, vv::Vector{Float64}) =
    chk(ccall((:ab_csr_update_values, LIB), Cint, (Int64, Ptr{Cdouble}), s.h, vv))

"run(sess, S) — src/sparse.jl:175-178: the assembled matrix with duplicates summed."
function Base.convert(::Type{SparseMatrixCSC}, s::B200SparseTensor)
    m, n, nnz, T = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    chk(ccall((:ab_csr_query, LIB), Cint, (Int64, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Int64}), s.h, m, n, nnz, T))
    rp, ci, va = Vector{Int32}(undef, m[] + 1), Vector{Int32}(undef, nnz[]), Vector{Float64}(undef, nnz[])
    chk(ccall((:ab_csr_export, LIB), Cint, (Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Cdouble}), s.h, rp, ci, va))
    rows = vcat([fill(r, rp[r+1] - rp[r]) for r = 1:m[]]...)
    sparse(rows, Int.(ci) .+ 1, va, m[], n[])
end

# ---- s * o and o * s  — src/sparse.jl:227-253 (row-major [n,k] like TensorFlow) -------------------
function b200_spmm(s::B200SparseTensor, X::Array{Float64}; trans::Bool=false)
    k = ndims(X) == 1 ? 1 : size(X, 2)
    Xr = ndims(X) == 1 ? X : permutedims(X)            # Julia is column-major; the ABI is row-major
    Y = Array{Float64}(undef, k, trans ? s.n : s.m)
    chk(ccall((:ab_spmm, LIB), Cint, (Int64, Cint, Ptr{Cdouble}, Int64, Ptr{Cdouble}), s.h, trans ? 1 : 0, Xr, k, Y))
    ndims(X) == 1 ? vec(Y) : permutedims(Y)
end
Base.:*(s::B200SparseTensor, o::Array{Float64}) = b200_spmm(s, o)
Base.:*(o::Array{Float64,2}, s::B200SparseTensor) = permutedims(b200_spmm(s, permutedims(o); trans=true))

"Gradients of Y = A*X w.r.t. the triplet values (one per input triplet) and the dense operand."
function b200_spmm_grad(s::B200SparseTensor, dY::Array{Float64}, X::Array{Float64})
    k = ndims(X) == 1 ? 1 : size(X, 2)
    dYr, Xr = (ndims(X) == 1 ? dY : permutedims(dY)), (ndims(X) == 1 ? X : permutedims(X))
    dv = Vector{Float64}(undef, s.T)
    chk(ccall((:ab_spmm_grad_values, LIB), Cint, (Int64, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}), s.h, dYr, Xr, k, dv))
    dX = Array{Float64}(undef, k, s.n)
    chk(ccall((:ab_spmm_grad_dense, LIB), Cint, (Int64, Ptr{Cdouble}, Int64, Ptr{Cdouble}), s.h, dYr, k, dX))
    dv, (ndims(X) == 1 ? vec(dX) : permutedims(dX))
end

# ---- s \ o  — src/sparse.jl:413-445, SparseSolver.h:13-70 -----------------------------------------
function b200_solve(s::B200SparseTensor, f::Vector{Float64}, method::String="SparseLU"; tol=0.0, maxit=0)
    u = Vector{Float64}(undef, s.n)
    it, rr = Ref{Cint}(0), Ref{Cdouble}(0)
    chk(ccall((:ab_solve, LIB), Cint,
              (Int64, Cstring, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cint, Ref{Cint}, Ref{Cdouble}),
              s.h, method, f, u, tol, maxit, it, rr))
    u
end
Base.:\(s::B200SparseTensor, f::Vector{Float64}) = b200_solve(s, f)

"(grad_vv, grad_f) — the SparseSolverGrad outputs that are defined (SparseSolver.h:46-70)."
function b200_solve_grad(s::B200SparseTensor, grad_u::Vector{Float64}, u::Vector{Float64}, method::String="SparseLU")
    gvv, gf = Vector{Float64}(undef, s.T), Vector{Float64}(undef, s.n)
    chk(ccall((:ab_solve_grad, LIB), Cint,
              (Int64, Cstring, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cint, Ptr{Cint}, Ptr{Cdouble}),
              s.h, method, grad_u, u, gvv, gf, 0.0, 0, C_NULL, C_NULL))
    gvv, gf
end

# register(forward, backward) shape of src/extra.jl:516-541:
#   forward  = (vv, f) -> (update_values!(A, vv); b200_solve(A, f))
#   backward = (du, u, vv, f) -> b200_solve_grad(A, du, u)

# ---- SparseAssembler / accumulate / assemble  — src/sparse.jl:480-534 ------------------------------
struct B200Assembler
    handle::Int32
end
function B200Assembler(handle::Integer, n::Integer, tol::Real=0.0)
    chk(ccall((:ab_asm_create, LIB), Cint, (Int32, Int32, Cdouble), handle, n, tol))
    B200Assembler(Int32(handle))
end
b200_accumulate!(a::B200Assembler, row::Integer, cols::Vector{Int32}, vals::Vector{Float64}) =
    chk(ccall((:ab_asm_add, LIB), Cint, (Int32, Int32, Ptr{Int32}, Ptr{Cdouble}, Int32), a.handle, row, cols, vals, length(vals)))
function b200_assemble(m::Integer, n::Integer, a::B200Assembler)
    k = ccall((:ab_asm_nnz, LIB), Int64, (Int32,), a.handle)
    k < 0 && chk(Cint(k))
    r, c, v = Vector{Int32}(undef, k), Vector{Int32}(undef, k), Vector{Float64}(undef, k)
    chk(ccall((:ab_asm_copy, LIB), Cint, (Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Cdouble}), a.handle, r, c, v))
    B200SparseTensor(Int64.(r), Int64.(c), v, m, n)
end

# ---- compress  — src/sparse.jl:776-781 (indices: 0-based [N,2], row-major) -------------------------
function b200_compress(indices::Array{Int64,2}, v::Vector{Float64})
    idx = permutedims(indices)                        # -> row-major [N,2]
    N = length(v); nout = Ref{Int64}(0)
    chk(ccall((:ab_compress_count, LIB), Cint, (Int64, Ptr{Int64}, Ref{Int64}), N, idx, nout))
    oi, ov = Array{Int64}(undef, 2, nout[]), Vector{Float64}(undef, nout[])
    chk(ccall((:ab_compress, LIB), Cint, (Int64, Ptr{Int64}, Ptr{Cdouble}, Ptr{Int64}, Ptr{Cdouble}), N, idx, v, oi, ov))
    permutedims(oi), ov
end

# ---- factorize / solve  — src/sparse.jl:671-699 ----------------------------------------------------
function b200_factorize(ii::Vector{Int64}, jj::Vector{Int64}, vv::Vector{Float64}, d::Integer, max_cache_size::Integer=999999)
    id = Ref{Int64}(0)
    chk(ccall((:ab_factorize, LIB), Cint, (Ptr{Int64}, Ptr{Int64}, Ptr{Cdouble}, Int64, Int64, Int64, Ref{Int64}),
              ii, jj, vv, length(vv), d, max_cache_size, id))
    id[]
end
function b200_fsolve(id::Int64, rhs::Vector{Float64})
    out = similar(rhs)
    chk(ccall((:ab_factorized_solve, LIB), Cint, (Int64, Ptr{Cdouble}, Int64, Ptr{Cdouble}), id, rhs, length(rhs), out))
    out
end

# ---- structural ops between assembly and solve (SURVEY §8f N2/N4) — csrc/structure.cu ---------------
# Variable-length triplet results come back through a triplet handle: size -> allocate -> fetch.
function fetch_trip(th::Int64)
    k = ccall((:ab_trip_size, LIB), Int64, (Int64,), th)
    k < 0 && chk(Cint(k))
    ii, jj, vv = Vector{Int64}(undef, k), Vector{Int64}(undef, k), Vector{Float64}(undef, k)
    chk(ccall((:ab_trip_fetch, LIB), Cint, (Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Cdouble}), th, ii, jj, vv))
    ii, jj, vv
end

"zero_out_row(A, bd) — src/sparse.jl:790-797.  `indices` is the 0-based [N,2] matrix of the tf.SparseTensor."
function b200_zero_out_row(indices::Array{Int64,2}, vv::Vector{Float64}, bd::Vector{Int64})
    idx = permutedims(indices); th = Ref{Int64}(0)
    chk(ccall((:ab_zero_out_row, LIB), Cint, (Int64, Ptr{Int64}, Ptr{Cdouble}, Ptr{Int64}, Int64, Ref{Int64}),
              length(vv), idx, vv, bd, length(bd), th))
    ii, jj, ov = fetch_trip(th[])
    hcat(ii, jj), ov
end

"scatter_update(A, i1, i2, B) — src/sparse.jl:340-364 (all 1-based; (uii,ujj) are positions inside the block)."
function b200_scatter_update(oii, ojj, ovv, m, n, uii, ujj, uvv, ii::Vector{Int64}, jj::Vector{Int64})
    th = Ref{Int64}(0)
    chk(ccall((:ab_scatter_update, LIB), Cint,
              (Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Cdouble}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Cdouble}, Int64, Int64,
               Ptr{Int64}, Int64, Ptr{Int64}, Int64, Ref{Int64}),
              length(ovv), oii, ojj, ovv, length(uvv), uii, ujj, uvv, m, n, ii, length(ii), jj, length(jj), th))
    fetch_trip(th[])
end

"A[i1, i2] — src/sparse.jl:303-327: 1-based triplets of the sub-block, in the reference op's order."
function b200_getindex(s::B200SparseTensor, ii::Vector{Int64}, jj::Vector{Int64})
    th = Ref{Int64}(0)
    chk(ccall((:ab_sparse_indexing, LIB), Cint, (Int64, Ptr{Int64}, Int64, Ptr{Int64}, Int64, Ref{Int64}),
              s.h, ii, length(ii), jj, length(jj), th))
    fetch_trip(th[])
end

"SparseTensor * SparseTensor — src/sparse.jl:579-595; returns the handle of C = A*B (kept on the device)."
function b200_spgemm(a::B200SparseTensor, b::B200SparseTensor)
    h = Ref{Int64}(0)
    chk(ccall((:ab_spgemm, LIB), Cint, (Int64, Int64, Cint, Ref{Int64}), a.h, b.h, 0, h))
    h[]
end

"A + B (tf.sparse.add, src/sparse.jl:217): handle of alpha*A + beta*B on the union pattern."
function b200_add(a::B200SparseTensor, b::B200SparseTensor, alpha::Real=1.0, beta::Real=1.0)
    h = Ref{Int64}(0)
    chk(ccall((:ab_csr_axpby, LIB), Cint, (Int64, Cdouble, Int64, Cdouble, Ref{Int64}), a.h, alpha, b.h, beta, h))
    h[]
end

"trisolve(a, b, c, d) — src/sparse.jl:731-739."
function b200_trisolve(a::Vector{Float64}, b::Vector{Float64}, c::Vector{Float64}, d::Vector{Float64})
    x = similar(d)
    chk(ccall((:ab_tri_solve, LIB), Cint, (Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
              length(b), a, b, c, d, x))
    x
end

# ---- Newton-Raphson inner loop (SURVEY §8f N3) — csrc/newton.cu; src/optim.jl:509-539 --------------
"δ = J \\ r ; u_new = u - step*δ ; returns (u_new, δ, ‖r‖).  The residual/Jacobian callback stays in Julia."
function b200_newton_step(J::B200SparseTensor, r::Vector{Float64}, u::Vector{Float64}; step::Real=1.0, method::String="SparseLU")
    unew, delta, rn = similar(u), similar(u), Ref{Cdouble}(0.0)
    chk(ccall((:ab_newton_step, LIB), Cint,
              (Int64, Cstring, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}, Cdouble, Cint, Ptr{Cint}),
              J.h, method, r, u, step, unew, delta, rn, 0.0, 0, C_NULL))
    unew, delta, rn[]
end

end # module
