# PathMatFacB200.jl -- Julia shim that drops libpmf (include/pmf.h) in for the fit-loop hot path
# of PathMatFac.jl.  `using PathMatFacB200` REDEFINES the method the reference itself calls,
# `PathMatFac.mf_fit!(model::PathMatFacModel; kwargs...)` (src/fit.jl:9-38; its only caller is
# `mf_fit_adapt_lr!`, src/fit.jl:58), so the staging code (`fit!`, `mf_fit_adapt_lr!`, `basic_fit!`, ...,
# `transform`) keeps running in Julia unchanged and every inner fit lands on the GPU.  A model is bound to a
# device handle with `to_device(model)` (the `gpu(model)` of fit_matfac.jl:325-340); a model without a handle
# gets a transient one for the duration of the call.  There is no CPU fallback.
#
# WRITE-ONLY in this repository: the build image has no Julia runtime, so this file is exercised
# only indirectly -- the Python mirror (pathmatfac.jl_b200/fit.py) drives the very same C ABI
# with the very same marshalling rules (INTEGRATION.md, SURVEY.md Appendix C).
module PathMatFacB200

using SparseArrays
import PathMatFac
const PM = PathMatFac

const LIBPMF = get(ENV, "LIBPMF", joinpath(@__DIR__, "..", "pathmatfac.jl_b200", "libpmf.so"))
const TERM_CODES = ("max_epochs", "abs_tol", "rel_tol", "loss_increase", "nonfinite")
const DIST_CODE = Dict("NormalNoise" => 0, "BernoulliNoise" => 1, "PoissonNoise" => 2, "OrdinalNoise" => 3,
                       "SquaredHingeNoise" => 4, "OrdinalSqHingeNoise" => 5)

struct PmfDims
    M::Int32; N::Int32; K::Int32; device::Int32
end

mutable struct PmfFitOpts            # mirrors pmf_fit_opts
    max_epochs::Int32; epoch::Int32; lr::Float32; adagrad_eps::Float32
    rel_tol::Float64; abs_tol::Float64
    update_X::Int32; update_Y::Int32; update_col_layers::Int32
    kernel::Int32; precision::Int32; check_every::Int32; no_terminate::Int32
    update_noise_models::Int32; alternating::Int32
end

mutable struct PmfHistory            # mirrors pmf_history
    term_code::Int32; epochs::Int32; n_recorded::Int32; capacity::Int32
    loss_total::Ptr{Float64}; loss_data::Ptr{Float64}; loss_x_reg::Ptr{Float64}
    loss_y_reg::Ptr{Float64}; loss_layer_reg::Ptr{Float64}
    device_ms::Float32; kernel_launches::Int64
end

const Handle = Ptr{Cvoid}

function check(h::Handle, rc::Integer)
    rc == 0 && return
    msg = ccall((:pmf_last_error, LIBPMF), Cstring, (Handle,), h)
    error("libpmf error $rc: ", msg == C_NULL ? "?" : unsafe_string(msg))
end

unwrap(l) = isa(l, PM.FrozenLayer) ? l.layer : l
unwrap_reg(r) = isa(r, PM.FrozenRegularizer) ? r.reg : r
f32(a) = convert(Array{Float32}, a)
i32(a) = convert(Vector{Int32}, a)

# model => handle.  Keyed by object identity: the stage functions mutate the model in place and pass the same
# object down to mf_fit! (src/fit.jl:58), which looks its handle up here.
const HANDLES = IdDict{PM.PathMatFacModel,Handle}()

function create_handle(model::PM.PathMatFacModel, device::Integer)
    M, N = size(model.data); K = size(model.matfac.X, 1)
    h = Ref{Handle}(C_NULL)
    check(C_NULL, ccall((:pmf_create, LIBPMF), Cint, (Ref{PmfDims}, Ref{Handle}), PmfDims(M, N, K, device), h))
    A = f32(model.data)                                    # M x N column-major, NaN = missing
    check(h[], ccall((:pmf_set_data, LIBPMF), Cint, (Handle, Ptr{Float32}), h[], A))
    push_layout!(h[], model)                               # once per handle: it re-allocates the batch tables
    return h[]
end

"""gpu(model) equivalent (fit_matfac.jl:325-340): create the handle, upload the data once, bind it to the model."""
function to_device(model::PM.PathMatFacModel; device::Integer=0)
    haskey(HANDLES, model) && return HANDLES[model]
    HANDLES[model] = create_handle(model, device)
end

"""cpu(model) equivalent: the host model is current after every fit, so this only frees the device side."""
function release(model::PM.PathMatFacModel)
    h = pop!(HANDLES, model, C_NULL)
    h == C_NULL || release(h)
    return model
end
function release(h::Handle)
    for (opt, hs) in OPT_OWNERS; filter!(x -> x != h, hs); end
    ccall((:pmf_destroy, LIBPMF), Cint, (Handle,), h)
end
release_cached_memory() = ccall((:pmf_release_cached_memory, LIBPMF), Cint, ())

function push_params!(h::Handle, model)
    mf = model.matfac
    check(h, ccall((:pmf_set_factors, LIBPMF), Cint, (Handle, Ptr{Float32}, Ptr{Float32}), h, f32(mf.X), f32(mf.Y)))
    layers = mf.col_transform.layers
    check(h, ccall((:pmf_set_col_params, LIBPMF), Cint, (Handle, Ptr{Float32}, Ptr{Float32}), h,
                   f32(unwrap(layers[1]).logsigma), f32(unwrap(layers[3]).mu)))
    if isa(unwrap(layers[2]), PM.BatchScale)
        ld = unwrap(layers[2]).logdelta; th = unwrap(layers[4]).theta
        for v in 1:length(ld.values)
            check(h, ccall((:pmf_set_batch_values, LIBPMF), Cint, (Handle, Int32, Ptr{Float32}, Ptr{Float32}),
                           h, v - 1, f32(ld.values[v]), f32(th.values[v])))
        end
    end
end

# Batch layout (views, batch counts, batch of every sample): structure, not values.  Sent when the handle is created;
# sending an identical layout again is a no-op in the library (parameters, AdaGrad accumulators and the tensor-core
# batch plan survive), a different one re-allocates the batch tables.
function push_layout!(h::Handle, model)
    layers = model.matfac.col_transform.layers
    if isa(unwrap(layers[2]), PM.BatchScale)
        ld = unwrap(layers[2]).logdelta
        nv = length(ld.col_ranges); M = size(model.data, 1)
        cs = i32([r.start - 1 for r in ld.col_ranges]); ce = i32([r.stop for r in ld.col_ranges])
        nb = i32([size(v, 1) for v in ld.values])
        bos = fill(Int32(-1), M, nv)                       # column v = batch ordinal (0-based) of every sample; a row of
                                                           # row_batches[v] without a nonzero stays -1, which the library refuses
        for v in 1:nv
            I, J, _ = findnz(sparse(ld.row_batches[v]))
            bos[I, v] .= Int32.(J .- 1)
        end
        check(h, ccall((:pmf_set_batch_layout, LIBPMF), Cint, (Handle, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                       h, nv, cs, ce, nb, bos))
    else
        check(h, ccall((:pmf_set_batch_layout, LIBPMF), Cint, (Handle, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                       h, 0, C_NULL, C_NULL, C_NULL, C_NULL))
    end
end

function pull_params!(h::Handle, model)
    mf = model.matfac
    X = Matrix{Float32}(undef, size(mf.X)...); Y = Matrix{Float32}(undef, size(mf.Y)...)
    check(h, ccall((:pmf_get_factors, LIBPMF), Cint, (Handle, Ptr{Float32}, Ptr{Float32}), h, X, Y))
    mf.X .= X; mf.Y .= Y
    layers = mf.col_transform.layers
    ls = Vector{Float32}(undef, size(mf.Y, 2)); mu = similar(ls)
    check(h, ccall((:pmf_get_col_params, LIBPMF), Cint, (Handle, Ptr{Float32}, Ptr{Float32}), h, ls, mu))
    unwrap(layers[1]).logsigma .= ls; unwrap(layers[3]).mu .= mu
    if isa(unwrap(layers[2]), PM.BatchScale)
        ld = unwrap(layers[2]).logdelta; th = unwrap(layers[4]).theta
        for v in 1:length(ld.values)
            a = Matrix{Float32}(undef, size(ld.values[v])...); b = similar(a)
            check(h, ccall((:pmf_get_batch_values, LIBPMF), Cint, (Handle, Int32, Ptr{Float32}, Ptr{Float32}), h, v - 1, a, b))
            ld.values[v] .= a; th.values[v] .= b
        end
    end
    # interior ordinal thresholds, trained when update_noise_models (src/fit.jl:14)
    nm = mf.noise_model
    if any(hasproperty(n, :ext_thresholds) for n in nm.noises)
        th4 = zeros(Float32, 4, length(nm.noises))
        check(h, ccall((:pmf_get_thresholds, LIBPMF), Cint, (Handle, Int32, Ptr{Float32}), h, length(nm.noises), th4))
        for (i, n) in enumerate(nm.noises); hasproperty(n, :ext_thresholds) && (n.ext_thresholds[2:3] .= th4[2:3, i]); end
    end
end

function recognise_closure(r::Function, K::Integer)
    probes = (ones(Float32, K, 3), Float32.(reshape(sin.(1:5K), K, 5)), 3f0 .* Float32.(reshape(cos.(1:2K), K, 2)))
    vals = [Float64(r(P)) for P in probes]
    all(==(0.0), vals) && return nothing
    ssq = [sum(Float64.(P) .^ 2) for P in probes]
    w = vals[1] / (0.5 * ssq[1])
    (w > 0 && all(abs(v - 0.5 * w * q) <= 1e-4 * abs(0.5 * w * q) for (v, q) in zip(vals, ssq))) && return w
    error("PathMatFacB200: unsupported regulariser closure (only `x -> 0` and `x -> 0.5*w*sum(x.*x)`, the closures ",
          "src/fit.jl installs, can be moved to the device); use a regulariser object")
end

const model_K = Ref{Any}(zeros(Float32, 1, 1))     # X of the model being marshalled (size(., 1) = K), set by push_regs!

# one regulariser object -> ABI calls; `which` 0 = X_reg, 1 = Y_reg; p = mixture weight
function install_reg!(h::Handle, which::Integer, r, p::Real)
    p == 0 && return
    if isa(r, PM.L2Regularizer)
        check(h, ccall((:pmf_set_reg_l2, LIBPMF), Cint, (Handle, Int32, Ptr{Float32}, Float32), h, which, f32(r.weights), p))
    elseif isa(r, PM.GroupRegularizer)
        st = i32([g.start - 1 for g in r.group_idx]); en = i32([g.stop for g in r.group_idx])
        w = f32(hcat(r.group_weights...))                  # K x n_groups column-major == n_groups x K row-major
        check(h, ccall((:pmf_set_reg_group, LIBPMF), Cint, (Handle, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float32}, Float32),
                       h, which, length(st), st, en, w, p))
    elseif isa(r, PM.SelectiveL1Reg)
        check(h, ccall((:pmf_set_reg_sel_l1, LIBPMF), Cint, (Handle, Int32, Ptr{UInt8}, Ptr{Float32}, Float32),
                       h, which, convert(Matrix{UInt8}, r.l1_idx), f32(r.weight), p))
    elseif isa(r, PM.ARDRegularizer)
        st = i32([g.start - 1 for g in r.col_ranges]); en = i32([g.stop for g in r.col_ranges])
        check(h, ccall((:pmf_set_reg_ard, LIBPMF), Cint, (Handle, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float32}, Ptr{Float32}),
                       h, which, length(st), st, en, f32(collect(r.alpha)), f32(collect(r.beta))))
    elseif isa(r, PM.FeatureSetARDReg)
        check(h, ccall((:pmf_set_reg_fsard, LIBPMF), Cint, (Handle, Int32, Ptr{Float32}, Ptr{Float32}), h, which, f32(r.alpha), f32(r.beta)))
    elseif isa(r, PM.NetworkRegularizer)
        csr(m) = (mt = SparseMatrixCSC(transpose(m)); (i32(mt.colptr .- 1), i32(mt.rowval .- 1), f32(mt.nzval)))
        cat3(ms) = (t = [csr(m) for m in ms]; (vcat(first.(t)...), vcat(getindex.(t, 2)...), vcat(last.(t)...)))
        aa = cat3(r.AA); ab = cat3(r.AB); bb = cat3(r.BB)
        nv = i32([size(b, 1) for b in r.BB]); xv = f32(vcat(r.x_virtual...))
        check(h, ccall((:pmf_set_reg_network, LIBPMF), Cint,
                       (Handle, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float32},
                        Ptr{Int32}, Ptr{Int32}, Ptr{Float32}, Ptr{Float32}, Float32, Float32, Float32, Int32),
                       h, which, nv, aa[1], aa[2], aa[3], ab[1], ab[2], ab[3], bb[1], bb[2], bb[3], xv, p, 0f0, 0f0, 0))
    elseif isa(r, Function)
        # The stages install anonymous functions: `X -> 0.5f0*sum(X.*X)` (src/fit.jl:686, the DEFAULT fit! path),
        # `0.05 .* sum(x.^2)`-style quadratics (:266-267) and the zero closures `X -> 0f0` / `y -> 0` / `x -> 0.0`
        # (:415, :769, src/transform.jl:61,70).  A closure has no fields to marshal, so it is identified by probing:
        # identically zero -> nothing to install; 0.5 w sum(x^2) with one scalar w -> L2 with uniform weight w.
        # Anything else is an error: a penalty is never dropped silently.
        K = size(model_K[], 1)
        w = recognise_closure(r, K)
        w === nothing || check(h, ccall((:pmf_set_reg_l2, LIBPMF), Cint, (Handle, Int32, Ptr{Float32}, Float32),
                                        h, which, fill(Float32(w), K), p))
    else
        error("PathMatFacB200: unsupported regulariser ", typeof(r))
    end
end

function push_regs!(h::Handle, model)
    mf = model.matfac
    model_K[] = mf.X
    for (which, reg) in ((0, mf.X_reg), (1, mf.Y_reg))
        check(h, ccall((:pmf_clear_reg, LIBPMF), Cint, (Handle, Int32), h, which))
        if isa(reg, PM.CompositeRegularizer)
            for (r, p) in zip(reg.regularizers, reg.mixture_p); install_reg!(h, which, r, p); end
        else
            install_reg!(h, which, reg, 1.0)
        end
    end
    frozen_layers = UInt32(0); frozen_regs = UInt32(0)
    N = size(mf.Y, 2)
    for (s, l) in enumerate(mf.col_transform.layers); isa(l, PM.FrozenLayer) && (frozen_layers |= UInt32(1) << (s - 1)); end
    for (s, r0) in enumerate(mf.col_transform_reg.regs)
        isa(r0, PM.FrozenRegularizer) && (frozen_regs |= UInt32(1) << (s - 1))
        r = unwrap_reg(r0)
        if isa(r, PM.ColParamReg)
            w = zeros(Float32, N); c = zeros(Float32, N)
            for (rng, wi, ci) in zip(r.col_ranges, r.weights, r.centers); w[rng] .= wi; c[rng] .= ci; end
            check(h, ccall((:pmf_set_layer_reg_col, LIBPMF), Cint, (Handle, Int32, Ptr{Float32}, Ptr{Float32}), h, s, w, c))
        elseif isa(r, PM.BatchArrayReg)
            check(h, ccall((:pmf_set_layer_reg_batch, LIBPMF), Cint, (Handle, Int32, Ptr{Float32}, Ptr{Float32}),
                           h, s, f32(vcat(r.weights...)), f32(vcat(r.centers...))))
        end
    end
    check(h, ccall((:pmf_set_frozen, LIBPMF), Cint, (Handle, UInt32, UInt32), h, frozen_layers, frozen_regs))
    nm = mf.noise_model
    cs = i32([r.start - 1 for r in nm.col_ranges]); ce = i32([r.stop for r in nm.col_ranges])
    dc = i32([DIST_CODE[string(nameof(typeof(n)))] for n in nm.noises])
    th = zeros(Float32, 4, length(nm.noises))
    for (i, n) in enumerate(nm.noises); hasproperty(n, :ext_thresholds) && (th[:, i] .= n.ext_thresholds); end
    w = f32(vcat([n.weight for n in nm.noises]...))
    check(h, ccall((:pmf_set_noise, LIBPMF), Cint, (Handle, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float32}, Ptr{Float32}),
                   h, length(cs), cs, ce, dc, th, w))
end

# A fresh Flux AdaGrad object means fresh accumulators (src/fit.jl:55); the same object across the LR-halving
# restarts of mf_fit_adapt_lr! keeps them (:55-64).  Per optimiser the handles whose accumulators belong to it.
const OPT_OWNERS = IdDict{Any,Vector{Handle}}()

"""
The method the reference calls (src/fit.jl:58: `mf_fit!(model; opt=opt, capacity=..., max_epochs=..., epoch=...,
keep_history=true, kwargs...)`), same keyword surface as src/fit.jl:9-21.  `capacity` is accepted and ignored (the
fused pass never materialises Z), `scale_column_losses=true` is refused (never used by the reference).
"""
function PM.mf_fit!(model::PM.PathMatFacModel; scale_column_losses=false, update_X=false, update_Y=false,
                    update_row_layers=false, update_col_layers=false, update_noise_models=true,
                    reg_relative_weighting=false, update_X_reg=false, update_Y_reg=false,
                    update_row_layers_reg=false, update_col_layers_reg=false, keep_history=true,
                    opt=PM.construct_optimizer(model, 1.0), max_epochs=1000, epoch=1, rel_tol=1e-5, abs_tol=1e-5,
                    capacity=nothing, verbosity=1, print_prefix="", alternating=false, device=0, kwargs...)
    scale_column_losses && error("PathMatFacB200: scale_column_losses=true is not implemented (unused by the reference)")
    transient = !haskey(HANDLES, model)
    h = transient ? create_handle(model, device) : HANDLES[model]
    try
        push_params!(h, model); push_regs!(h, model)
        owners = get!(OPT_OWNERS, opt, Handle[])
        if !(h in owners)
            check(h, ccall((:pmf_reset_opt_state, LIBPMF), Cint, (Handle, Float32), h, opt.epsilon)); push!(owners, h)
        end
        cap = max(1, max_epochs - epoch + 1)
        losses = [zeros(Float64, cap) for _ in 1:5]
        hist = PmfHistory(0, 0, 0, cap, pointer.(losses)..., 0f0, 0)
        opts = PmfFitOpts(max_epochs, epoch, opt.eta, opt.epsilon, rel_tol, abs_tol, update_X, update_Y, update_col_layers,
                          0, 0, 8, 0, update_noise_models, alternating)
        GC.@preserve losses check(h, ccall((:pmf_fit, LIBPMF), Cint, (Handle, Ref{PmfFitOpts}, Ref{PmfHistory}), h, opts, hist))
        pull_params!(h, model)
        n = hist.n_recorded
        return Dict("term_code" => TERM_CODES[hist.term_code + 1], "epochs" => Int(hist.epochs), "loss" => losses[1][1:n],
                    "data_loss" => losses[2][1:n], "X_reg" => losses[3][1:n], "Y_reg" => losses[4][1:n],
                    "layer_reg" => losses[5][1:n])
    finally
        transient && release(h)
    end
end

# ---- the staging code's other streaming passes (SURVEY 8f rank 1) ------------------------------------

"""`MF.link_col_sqerr` + `MF.column_nonnan` (src/fit.jl:138-140) at the handle's current parameters."""
function link_col_sqerr(h::Handle, N::Integer)
    sq = zeros(Float32, N); cnt = zeros(Float32, N)
    check(h, ccall((:pmf_link_col_sqerr, LIBPMF), Cint, (Handle, Ptr{Float32}, Ptr{Float32}), h, sq, cnt))
    return sq, cnt
end

"""`MF.batched_column_ssq_grads` + `MF.column_nonnan` (src/fit.jl:166-168, :140)."""
function column_stats(h::Handle, N::Integer)
    ssq = zeros(Float32, N); cnt = zeros(Float32, N)
    check(h, ccall((:pmf_column_stats, LIBPMF), Cint, (Handle, Ptr{Float32}, Ptr{Float32}), h, ssq, cnt))
    return ssq, cnt
end

"""`ba_map(d->isfinite.(d), theta, data)` and `ba_map(MF.sqerr_func, theta, model, data)`
(src/batch_array.jl:320-334): one n_b x N_v matrix per batched view."""
function batch_stats(h::Handle, theta::PM.BatchArray)
    cnt = [zeros(Float32, size(v)) for v in theta.values]; sq = [zeros(Float32, size(v)) for v in theta.values]
    GC.@preserve cnt sq check(h, ccall((:pmf_batch_stats, LIBPMF), Cint, (Handle, Int32, Ptr{Ptr{Float32}}, Ptr{Ptr{Float32}}),
                                       h, length(cnt), pointer.(cnt), pointer.(sq)))
    return cnt, sq
end

# ---- sample-sharded multi-GPU: one Julia process per GPU (INTEGRATION.md section 3) ---------------------

"""Rank 0 creates the 128-byte NCCL id; broadcast it with MPI.jl / Distributed.jl, then `comm_init!`."""
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(C_NULL, ccall((:pmf_comm_unique_id, LIBPMF), Cint, (Ptr{UInt8},), id))
    return id
end
comm_init!(h::Handle, nranks::Integer, rank::Integer, id::Vector{UInt8}) =
    check(h, ccall((:pmf_comm_init_rank, LIBPMF), Cint, (Handle, Int32, Int32, Ptr{UInt8}), h, nranks, rank, id))
comm_destroy!(h::Handle) = check(h, ccall((:pmf_comm_destroy, LIBPMF), Cint, (Handle,), h))

end # module
