# FirstOrderSolversB200.jl -- the Julia side of the drop-in: thin `ccall` glue that replaces the
# hot loop of FirstOrderSolvers.jl with the B200 library (libfos_b200.so, include/fos_b200.h).
#
# UNTESTED: NOT RUNNABLE IN THE BUILD ENVIRONMENT (no julia binary; SURVEY.md F2).  It is kept deliberately
# thin -- every arithmetic decision lives behind the C ABI, which is what the parity tests drive
# (through ctypes and through the plain-C harness tests/c_abi_harness.c) -- so that there is little here to
# get wrong.  A maintainer applies it by `include`-ing this file at the end of src/FirstOrderSolvers.jl (after
# src/problemforms/Feasibility/FeasibilityStatus.jl).  It adds MORE SPECIFIC methods, so nothing of the reference
# is edited:
#   init_algorithm!(alg::GAP|GAPA|FISTA|Dykstra|GAPP|LineSearchWrapper, model::FOSMathProgModel)
#                                  beats init_algorithm!(alg, model::AbstractFOSModel)   (src/solvers/gap.jl:23-28 ...)
#   init_algorithm!(alg::..., model::FeasibilityModel{<:AffinePlusLinear,<:ConeProduct})   (Feasibility.jl:42)
#   iterate(alg, data::B200Data, status, x, max_iters)      beats iterate(alg, data::FOSSolverData, ...)
#                                                                               (src/solverwrapper.jl:20-41)
#   getcgiter(data::B200Data)                                                    (src/solvers/defaults.jl:27-29)
# and leaves the public API (GAP/DR/AP/GAPA/FISTA/Dykstra/GAPP constructors, MathProgBase methods,
# solve!(::Feasibility, alg; kw...), kwargs, model.history, printed table) untouched.

const libfos = get(ENV, "FOS_B200_LIB", "libfos_b200.so")

const STATUS_SYMBOL = Dict(0 => :Continue, 1 => :Optimal, 2 => :Unbounded, 3 => :Infeasible, 4 => :Indeterminate)
const FOS_REC_LEN = 10

# cone objects of `conemap` (src/cones.jl:4-14) -> FOS_CONE_* codes of include/fos_b200.h
conecode(::ProximalOperators.IndFree) = Int32(0)
conecode(::ProximalOperators.IndZero) = Int32(1)
conecode(::ProximalOperators.IndNonnegative) = Int32(2)
conecode(::ProximalOperators.IndNonpositive) = Int32(3)
conecode(::ProximalOperators.IndSOC) = Int32(4)
conecode(::ProximalOperators.IndRotatedSOC) = Int32(5)
conecode(::ProximalOperators.IndPSD) = Int32(6)
conecode(::ProximalOperators.IndExpPrimal) = Int32(7)
conecode(::ProximalOperators.IndExpDual) = Int32(8)
conecode(c) = error("fos_b200: cone $(typeof(c)) cannot cross the C ABI")
conearrays(K::ConeProduct) = (Int32[conecode(c) for c in K.cones], Int64[length(r) for r in K.ranges])

mutable struct B200Data <: FOSSolverData
    handle::Ptr{Cvoid}
    cgiter::Int64
    N::Int64            # iterate length: 2(m+n+1) (HSDE) or an+am (Feasibility)
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

function new_handle(device)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:fos_create, libfos), Int32, (Ref{Ptr{Cvoid}}, Int32), href, Int32(device))
    rc == 0 || error(unsafe_string(ccall((:fos_last_error, libfos), Cstring, (Ptr{Cvoid},), C_NULL)))
    return href[]
end

function set_algorithm_b200!(h, alg)
    code, a, a1, a2, b, ip = algparams(alg)
    fos_check(h, ccall((:fos_set_algorithm, libfos), Int32,
        (Ptr{Cvoid}, Int32, Float64, Float64, Float64, Float64, Int64), h, code, a, a1, a2, b, ip))
    # LineSearchWrapper(alg; lsinterval) (wrappers/linesearch.jl:19-33): GAP / GAPA only (support_linesearch)
    alg isa LineSearchWrapper && fos_check(h, ccall((:fos_set_linesearch, libfos), Int32, (Ptr{Cvoid}, Int64),
                                                    h, Int64(alg.lsinterval)))
end

function finish_data(h)
    N = ccall((:fos_iterate_length, libfos), Int64, (Ptr{Cvoid},), h)
    data = B200Data(h, 0, N)
    finalizer(d -> ccall((:fos_destroy, libfos), Int32, (Ptr{Cvoid},), d.handle), data)
    return data
end

# ---- MathProgBase / Convex.jl front door: HSDE(model; direct) on the device (HSDE.jl:7-29) ----------------
function init_algorithm_b200!(alg::FOSAlgorithm, model::FOSMathProgModel)
    h = new_handle(get(model.options, :device, 0))
    A = model.A                                   # SparseMatrixCSC{Float64,Int64}, passed as is (1-based)
    m, n = size(A)
    t1, l1 = conearrays(model.K1)                 # built by loadproblem! (FOSSolverInterface.jl:44-50)
    t2, l2 = conearrays(model.K2)
    fos_check(h, ccall((:fos_load_conic_csc, libfos), Int32,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64},
         Int64, Ptr{Int32}, Ptr{Int64}, Int64, Ptr{Int32}, Ptr{Int64}, Int32),
        h, m, n, A.colptr, A.rowval, A.nzval, 1, model.b, model.c,
        length(t1), t1, l1, length(t2), t2, l2, Int32(0)))
    # HSDE(model, direct=alg.direct) (FOSSolverInterface.jl:77, HSDE.jl:10-15): exact projection on the device
    direct = inneralg(alg).direct
    direct && fos_check(h, ccall((:fos_set_direct, libfos), Int32, (Ptr{Cvoid}, Int32), h, Int32(1)))
    set_algorithm_b200!(h, alg)
    # same closure as HSDE.jl:26-27 (it captures model.init_duration at construction time, SURVEY a-Q 11)
    status_generator = (mo, checki, eps, verbose, debug) ->
        HSDEStatus(m, n, 0, mo, :Continue, checki, eps, verbose, false, direct, time_ns(), model.init_duration, debug)
    return finish_data(h), status_generator
end

# ---- Feasibility front door (Feasibility.jl:51-81): only the GPU-able pair crosses the ABI ---------------
# S1 = AffinePlusLinear(A, b, q, beta; decreasing_accuracy) on [x; z], S2 = ConeProduct over the an+am entries.
# Any other pair of ProximableFunctions keeps the reference's CPU path (the generic init_algorithm! methods).
function init_algorithm_b200!(alg::FOSAlgorithm, model::FeasibilityModel{<:AffinePlusLinear,<:ConeProduct})
    h = new_handle(get(model.options, :device, 0))
    S1, S2 = model.S1, model.S2
    A = S1.A isa SparseMatrixCSC{Float64,Int64} ? S1.A : sparse(Matrix{Float64}(S1.A))
    am, an = size(A)
    t, l = conearrays(S2)
    fos_check(h, ccall((:fos_load_affine_csc, libfos), Int32,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64},
         Int32, Int32, Int64, Ptr{Int32}, Ptr{Int64}, Int32),
        h, am, an, A.colptr, A.rowval, A.nzval, 1, S1.b, S1.q, Int32(S1.β), Int32(S1.decreasing_accuracy),
        length(t), t, l, Int32(0)))
    set_algorithm_b200!(h, alg)
    # Feasibility.jl:75-81: direct = true is forced for printing (no `cg` column, no :cgiter key)
    status_generator = (mo, checki, eps, verbose, debug) ->
        FeasibilityStatus(mo.n, 0, mo, fill(NaN, mo.n), Array{Array{Float64,1},1}(), :Continue, checki, eps, verbose,
                          false, true, time_ns(), mo.init_duration, debug)
    return finish_data(h), status_generator
end
# IndBox(lo, hi) as S2 (test/testfeasibility.jl:10): load with a Free cone over the box range, then
#   ccall((:fos_set_box, libfos), Int32, (Ptr{Cvoid}, Int64, Int64, Float64, Float64), h, start0, len, lo, hi)

for T in (:GAP, :GAPA, :FISTA, :Dykstra, :GAPP, :LineSearchWrapper)
    @eval init_algorithm!(alg::$T, model::FOSMathProgModel) = init_algorithm_b200!(alg, model)
    @eval init_algorithm!(alg::$T, model::FeasibilityModel{<:AffinePlusLinear,<:ConeProduct}) =
        init_algorithm_b200!(alg, model)
end

getcgiter(data::B200Data) = data.cgiter

# debug > 1: x, y, s of the checked point (HSDEStatus.jl:133-135) = the last unrelaxed S2 projection
function fetch_checked_point(data::B200Data)
    z = Vector{Float64}(undef, data.N)
    fos_check(data.handle, ccall((:fos_get_state, libfos), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64),
                                 data.handle, Int32(5), z, length(z)))
    return z
end

# One status record of the library -> what checkstatus(::HSDEStatus, z) does on the Julia side
# (history push, printed row, "Found solution"): src/problemforms/HSDE/HSDEStatus.jl:27-71.
function absorb_record!(status::HSDEStatus, data::B200Data, rec::AbstractVector{Float64})
    i = Int(rec[1]); p, d, g, ctx, bty, κ, τ = rec[2:8]
    data.cgiter = Int(rec[9])
    t = time_ns() - status.init_time
    model = status.model
    if status.debug > 0                                            # savedata, :125-139
        for (k, v) in ((:p, p), (:d, d), (:g, g), (:ctx, ctx), (:bty, bty), (:κ, κ), (:τ, τ), (:t, t))
            push!(model.history, k, i, v)
        end
        if status.debug > 1                                        # copies, not views (superset of a-Q 4)
            z = fetch_checked_point(data)
            m, n = status.m, status.n; nu = n + m + 1
            push!(model.history, :x, i, z[1:n]); push!(model.history, :y, i, z[n+1:n+m])
            push!(model.history, :s, i, z[nu+n+1:nu+n+m])
        end
    end
    if status.verbose > 0                                          # :42-51
        if !status.direct
            push!(model.history, :cgiter, i, data.cgiter)
            printstatusiter(i, p, d, g, ctx, bty, κ/τ, data.cgiter, t)
        else
            printstatusiter(i, p, d, g, ctx, bty, κ/τ, t)          # no cg column, no :cgiter key
        end
        Int(rec[10]) == 1 && println("Found solution i=$i")
    end
    status.status = STATUS_SYMBOL[Int(rec[10])]
    status.checked = true
end

# FeasibilityStatus.jl:32-72 (direct = true is forced by Feasibility.jl:76, so the 3-argument row is printed)
function absorb_record!(status::FeasibilityStatus, data::B200Data, rec::AbstractVector{Float64})
    i = Int(rec[1]); err = rec[2]
    data.cgiter = Int(rec[9])
    t = time_ns() - status.init_time
    model = status.model
    if status.debug > 0                                            # savedata, :94-103
        push!(model.history, :err, i, err); push!(model.history, :t, i, t)
        if status.debug > 1
            push!(model.history, :z, i, fetch_checked_point(data))
            push!(model.history, :extra, i, Array{Array{Float64,1},1}())   # logextra copies are not kept on the device
        end
    end
    if status.verbose > 0
        if !status.direct
            push!(model.history, :cgiter, i, data.cgiter)
            printstatusiter(i, err, data.cgiter, t)
        else
            printstatusiter(i, err, t)
        end
        Int(rec[10]) == 1 && println("Found solution i=$i")
    end
    status.status = STATUS_SYMBOL[Int(rec[10])]
    status.checked = true
end

# replaces iterate() (src/solverwrapper.jl:20-41): `checki` iterations per ccall
function iterate(alg::FOSAlgorithm, data::B200Data, status::AbstractStatus, x, max_iters)
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
    t1 = Int32[conecode(conemap[c[1]]) for c in constr_cones]; l1 = Int64[length(c[2]) for c in constr_cones]
    t2 = Int32[conecode(conemap[c[1]]) for c in var_cones];    l2 = Int64[length(c[2]) for c in var_cones]
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
