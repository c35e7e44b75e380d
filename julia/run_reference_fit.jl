# Run the UNMODIFIED reference on a problem exported by pathmatfac.jl_b200/simulate.py::export_problem and write what
# the pinning tool (oracle/pin_against_julia.py) compares: the parameters after `epochs` epochs of one `mf_fit!` call and
# the history MatFac.jl returned.  This is how the "parity unpinned" part of the oracle (noise-model losses, epoch order,
# termination: the un-vendored MatFac.jl) gets pinned the day a Julia runtime is available.  Usage:
#     julia --project=/path/to/PathMatFac.jl run_reference_fit.jl <export dir> <K> [lr=0.05] [epochs=5] [key=value ...]
# The extra key=value pairs go to the PathMatFacModel constructor (e.g. lambda_X_l2=1.0), the same ones must be given to
# the pinning tool.  WRITE-ONLY here: the build image has no Julia.
include(joinpath(@__DIR__, "load_exported_problem.jl"))
import JSON

function write_bin(path, a)
    open(path, "w") do io
        write(io, convert(Array{Float32}, a))                 # column-major, little endian
    end
end

function main(args)
    dir, K = args[1], parse(Int, args[2])
    lr = length(args) >= 3 ? parse(Float64, args[3]) : 0.05
    epochs = length(args) >= 4 ? parse(Int, args[4]) : 5
    kw = Dict{Symbol,Any}()
    for a in args[5:end]
        k, v = split(a, "=")
        kw[Symbol(k)] = parse(Float64, v)
    end
    model = load_exported_problem(dir; K=K, kw...)
    opt = PM.construct_optimizer(model, lr)
    # tolerances 0: every epoch runs, so the parameters after `epochs` epochs do not depend on the termination test
    h = PM.mf_fit!(model; opt=opt, max_epochs=epochs, update_X=true, update_Y=true, update_col_layers=true,
                   keep_history=true, verbosity=0, rel_tol=0.0, abs_tol=0.0)
    mf = model.matfac
    write_bin(joinpath(dir, "ref_X.bin"), mf.X); write_bin(joinpath(dir, "ref_Y.bin"), mf.Y)
    write_bin(joinpath(dir, "ref_logsigma.bin"), mf.col_transform.layers[1].logsigma)
    write_bin(joinpath(dir, "ref_mu.bin"), mf.col_transform.layers[3].mu)
    layer4 = mf.col_transform.layers[4]
    if !isa(layer4, Function)
        for (v, name) in enumerate(layer4.theta.col_range_ids)
            write_bin(joinpath(dir, "ref_theta__" * String(name) * ".bin"), layer4.theta.values[v])
            write_bin(joinpath(dir, "ref_logdelta__" * String(name) * ".bin"), mf.col_transform.layers[2].logdelta.values[v])
        end
    end
    open(joinpath(dir, "ref_history.json"), "w") do io
        JSON.print(io, Dict("lr" => lr, "epochs" => epochs, "history" => h))
    end
end

main(ARGS)
