# FirstOrderSolversB200.jl -- the Julia side of the drop-in: thin `ccall` glue that replaces the
# hot loop of FirstOrderSolvers.jl with the B200 library (libfos_b200.so, include/fos_b200.h).
#
# NOT RUNNABLE IN THE BUILD ENVIRONMENT (no julia binary; SURVEY.md F2).  It is kept deliberately
# thin -- every arithmetic decision lives behind the C ABI, which is what the parity tests drive
# (through ctypes) -- so that there is little here to get wrong.  A maintainer applies it by
# `include`-ing this file after `src/solverwrapper.jl` in src/FirstOrderSolvers.jl: it overrides
#   init_algorithm!(alg, model::FOSMathProgModel)   (src/solvers/gap.jl:23-28 and siblings)
#   iterate(alg, data::B200Data, status, x, max_iters)  (src/solverwrapper.jl:20-41)
#   getcgiter(data::B200Data)                        (src/solvers/defaults.jl:27-29)
# and leaves the public API (GAP/DR/AP/GAPA/FISTA/Dykstra/GAPP constructors, MathProgBase
# methods, kwargs, model.history, printed table) untouched.

const libfos = get(ENV, "FOS_B200_LIB", "libfos_b200.so")

const CONE_CODE = Dict(:Free => 0, :Zero => 1, :NonNeg => 2, :NonPos => 3, :SOC => 4,
                       :SOCRotated => 5, :SDP => 6, :ExpPrimal => 7, :ExpDual => 8)
const STATUS_SYMBOL = Dict(0 => :Continue, 1 => :Optimal, 2 => :Unbounded, 3 => :Infeasible, 4 => :Indeterminate)
const FOS_REC_LEN = 10

mutable struct B200Data <: FOSSolverData
    handle::Ptr{Cvoid}
    cgiter::Int64
end

function fos_check(h::Ptr{Cvoid}, rc::Int32)
    rc == 0 && return
    msg = unsafe_string(ccall((:fos_last_error, libfos), Cstring, (Ptr{Cvoid},), h))
    error("fos_b200 error $rc: $msg")
end

algparams(a::GAP)     = (Int32(0), a.α, a.α1, a.α2, 0.0, Int64(100))
algparams(a::GAPA)    = (Int32(1), a.α, 0.0, 0.0, a.β, Int64(100))
algparams(a::FISTA)   = (Int32(2), a.α, 0.0, 0.0, 0.0, Int64(100))
algparams(a::Dykstra) = (Int32(3), 0.0, 0.0, 0.0, 0.0, Int64(100))
algparams(a::GAPP)    = (Int32(4), a.α, a.α1, a.α2, 0.0, Int64(a.iproj))
algparams(a::LineSearchWrapper) = algparams(a.alg)          # wrappers/linesearch.jl:3-7: the inner algorithm's step
inneralg(a::FOSAlgorithm) = a
inneralg(a::LineSearchWrapper) = a.alg

conearrays(K::ConeProduct, names) =
    (Int32[CONE_CODE[s] for s in names], Int64[length(r) for r in K.ranges])

# replaces init_algorithm! for every algorithm: the device handle takes the place of GAPData etc.
function init_algorithm_b200!(alg::FOSAlgorithm, model::FOSMathProgModel, constr_cones, var_cones)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:fos_create, libfos), Int32, (Ref{Ptr{Cvoid}}, Int32), href, Int32(get(model.options, :device, 0)))
    rc == 0 || error(unsafe_string(ccall((:fos_last_error, libfos), Cstring, (Ptr{Cvoid},), C_NULL)))
    h = href[]
    A = model.A                                   # SparseMatrixCSC{Float64,Int64}, passed as is (1-based)
    m, n = size(A)
    t1 = Int32[CONE_CODE[c[1]] for c in constr_cones]; l1 = Int64[length(c[2]) for c in constr_cones]
    t2 = Int32[CONE_CODE[c[1]] for c in var_cones];    l2 = Int64[length(c[2]) for c in var_cones]
    fos_check(h, ccall((:fos_load_conic_csc, libfos), Int32,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64},
         Int64, Ptr{Int32}, Ptr{Int64}, Int64, Ptr{Int32}, Ptr{Int64}, Int32),
        h, m, n, A.colptr, A.rowval, A.nzval, 1, model.b, model.c,
        length(t1), t1, l1, length(t2), t2, l2, Int32(0)))
    # HSDE(model, direct=alg.direct) (FOSSolverInterface.jl:77, HSDE.jl:10-15): exact projection on the device
    inneralg(alg).direct && fos_check(h, ccall((:fos_set_direct, libfos), Int32, (Ptr{Cvoid}, Int32), h, Int32(1)))
    code, a, a1, a2, b, ip = algparams(alg)
    fos_check(h, ccall((:fos_set_algorithm, libfos), Int32,
        (Ptr{Cvoid}, Int32, Float64, Float64, Float64, Float64, Int64), h, code, a, a1, a2, b, ip))
    # LineSearchWrapper(alg; lsinterval) (wrappers/linesearch.jl:19-33): GAP / GAPA only (support_linesearch)
    alg isa LineSearchWrapper && fos_check(h, ccall((:fos_set_linesearch, libfos), Int32, (Ptr{Cvoid}, Int64),
                                                    h, Int64(alg.lsinterval)))
    data = B200Data(h, 0)
    finalizer(d -> ccall((:fos_destroy, libfos), Int32, (Ptr{Cvoid},), d.handle), data)
    m2, n2 = size(model.A)
    status_generator = (mo, checki, eps, verbose, debug) ->
        HSDEStatus(m2, n2, 0, mo, :Continue, checki, eps, verbose, false, inneralg(alg).direct, time_ns(), model.init_duration, debug)
    return data, status_generator
end

getcgiter(data::B200Data) = data.cgiter

# One status record of the library -> what checkstatus(::HSDEStatus, z) does on the Julia side
# (history push, printed row, "Found solution"): src/problemforms/HSDE/HSDEStatus.jl:27-71.
function absorb_record!(status::HSDEStatus, data::B200Data, rec::AbstractVector{Float64})
    i = Int(rec[1]); p, d, g, ctx, bty, κ, τ = rec[2:8]
    data.cgiter = Int(rec[9])
    t = time_ns() - status.init_time
    model = status.model
    if status.debug > 0
        for (k, v) in ((:p, p), (:d, d), (:g, g), (:ctx, ctx), (:bty, bty), (:κ, κ), (:τ, τ), (:t, t))
            push!(model.history, k, i, v)
        end
    end
    if status.verbose > 0
        push!(model.history, :cgiter, i, data.cgiter)
        printstatusiter(i, p, d, g, ctx, bty, κ/τ, data.cgiter, t)
        Int(rec[10]) == 1 && println("Found solution i=$i")
    end
    status.status = STATUS_SYMBOL[Int(rec[10])]
    status.checked = true
end

# replaces iterate() (src/solverwrapper.jl:20-41): `checki` iterations per ccall
function iterate(alg::FOSAlgorithm, data::B200Data, status::HSDEStatus, x, max_iters)
    h = data.handle
    t1 = time()
    printstatusheader(status)
    fos_check(h, ccall((:fos_set_iterate, libfos), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64), h, x, length(x)))
    fos_check(h, ccall((:fos_begin_solve, libfos), Int32, (Ptr{Cvoid},), h))
    rec = zeros(Float64, FOS_REC_LEN, 2)
    done = Ref{Int64}(0); st = Ref{Int32}(0); nrec = Ref{Int64}(0)
    i = 1
    while i <= max_iters && status.status == :Continue
        chunk = min(status.checki - ((i - 1) % status.checki), max_iters - i + 1)
        fos_check(h, ccall((:fos_run, libfos), Int32,
            (Ptr{Cvoid}, Int64, Int64, Int64, Float64, Ref{Int64}, Ref{Int32}, Ptr{Float64}, Int64, Ref{Int64}, Ptr{Float64}),
            h, i, chunk, status.checki, status.eps, done, st, rec, 2, nrec, C_NULL))
        status.i = i + done[] - 1
        status.checked = false
        for k in 1:min(nrec[], 2)
            absorb_record!(status, data, view(rec, :, k))
        end
        i += done[]
        done[] < chunk && break
    end
    guess = similar(x)
    fos_check(h, ccall((:fos_finish, libfos), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ref{Int64}, Ref{Int32}),
        h, guess, length(guess), rec, nrec, st))
    nrec[] > 0 && absorb_record!(status, data, view(rec, :, 1))     # forced final check, solverwrapper.jl:32-34
    warned = Ref{Float64}(0.0)
    ccall((:fos_get_info, libfos), Int32, (Ptr{Cvoid}, Int32, Ref{Float64}), h, Int32(4), warned)
    warned[] != 0 && @warn "CG reached max iterations, result may be inaccurate"
    if status.verbose > 0
        println("Time for iterations: ")
        println("$(time() - t1) s")
    end
    return guess
end

# ---------------------------------------------------------------------------------------------
# Batch mode (config 5): `[solve!(m) for m in models]` for models that share (m, n, cones), solved
# by ONE kernel launch -- one persistent CTA per problem, no host round trips (fos_*_batch).
#   As :: Array{Float64,3} of size (n, m, B): Julia is column-major, so As[:, :, j] is the ROW-major
#   m x n matrix of problem j that the library expects (lda = n, problem stride = m*n);
#   bs :: (m, B), cs :: (n, B).
# Returns (guess (2(m+n+1), B), status::Vector{Symbol}, iterations, records (10, ncheck, B)).
# ---------------------------------------------------------------------------------------------
function solve_batch_b200(alg::FOSAlgorithm, cs::Matrix{Float64}, As::Array{Float64,3}, bs::Matrix{Float64},
                          constr_cones, var_cones; device = 0)
    n, m, B = size(As)
    opts = Dict{Symbol,Any}(alg.options)
    max_iters = get(opts, :max_iters, 10000); eps = get(opts, :eps, 1e-5); checki = get(opts, :checki, 100)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:fos_create, libfos), Int32, (Ref{Ptr{Cvoid}}, Int32), href, Int32(device))
    rc == 0 || error(unsafe_string(ccall((:fos_last_error, libfos), Cstring, (Ptr{Cvoid},), C_NULL)))
    h = href[]
    t1 = Int32[CONE_CODE[c[1]] for c in constr_cones]; l1 = Int64[length(c[2]) for c in constr_cones]
    t2 = Int32[CONE_CODE[c[1]] for c in var_cones];    l2 = Int64[length(c[2]) for c in var_cones]
    try
        fos_check(h, ccall((:fos_load_conic_dense_batch, libfos), Int32,
            (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Float64}, Int64, Int64, Int32, Ptr{Float64}, Ptr{Float64},
             Int64, Ptr{Int32}, Ptr{Int64}, Int64, Ptr{Int32}, Ptr{Int64}),
            h, B, m, n, As, n, m * n, Int32(0), bs, cs, length(t1), t1, l1, length(t2), t2, l2))
        code, a, a1, a2, b, ip = algparams(alg)
        fos_check(h, ccall((:fos_set_algorithm, libfos), Int32,
            (Ptr{Cvoid}, Int32, Float64, Float64, Float64, Float64, Int64), h, code, a, a1, a2, b, ip))
        N = 2 * (m + n + 1)
        cap = div(max_iters, checki) + 2
        guess = zeros(Float64, N, B); rec = zeros(Float64, FOS_REC_LEN, cap, B)
        done = zeros(Int64, B); st = zeros(Int32, B); nrec = zeros(Int64, B)
        fos_check(h, ccall((:fos_solve_batch, libfos), Int32,
            (Ptr{Cvoid}, Int64, Int64, Float64, Ptr{Float64}, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}, Int64, Ptr{Int64}),
            h, max_iters, checki, eps, guess, done, st, rec, cap, nrec))
        syms = (:Continue, :Optimal, :Unbounded, :Infeasible, :Indeterminate)
        return guess, [syms[s + 1] for s in st], done, rec
    finally
        ccall((:fos_destroy, libfos), Int32, (Ptr{Cvoid},), h)
    end
end

# ---------------------------------------------------------------------------------------------
# Multi-GPU (one Julia process per GPU, e.g. MPI.jl): row-sharded dense A.
#   fos_comm_unique_id on rank 0 -> MPI.Bcast -> fos_comm_init -> fos_load_conic_dense(row block)
#   -> fos_comm_p2p_export -> MPI.Allgather of the 64-byte handles -> fos_comm_p2p_import.
# After the import every pass over A exchanges its partial sums through CUDA-IPC peer memory inside
# the fused kernels (k1_exchange_p2p / k_cg_tail_hsde); without it the library falls back to
# fold + ncclAllReduce.
# ---------------------------------------------------------------------------------------------
function enable_p2p_exchange_b200(h::Ptr{Cvoid}, nranks::Integer, allgather::Function)
    mine = zeros(UInt8, 64)
    fos_check(h, ccall((:fos_comm_p2p_export, libfos), Int32, (Ptr{Cvoid}, Ptr{UInt8}), h, mine))
    table = allgather(mine)::Vector{UInt8}          # nranks * 64 bytes, rank order
    length(table) == 64 * nranks || error("handle table has the wrong size")
    fos_check(h, ccall((:fos_comm_p2p_import, libfos), Int32, (Ptr{Cvoid}, Ptr{UInt8}), h, table))
end
