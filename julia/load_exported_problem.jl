# Rebuild the inputs written by pathmatfac.jl_b200/simulate.py::export_problem (raw little-endian column-major arrays +
# manifest.txt) as a PathMatFacModel of the reference, so that a Julia run of the unmodified reference
# (`fit!(model; ...)`, CPU or `gpu(model)`) and libpmf see the same bytes.  Usage:
#     include("load_exported_problem.jl"); model = load_exported_problem("/path/to/dir"; K=64)
import PathMatFac
const PM = PathMatFac

function load_exported_problem(dir::AbstractString; K::Integer, kwargs...)
    arrays = Dict{String,Array}()
    for line in eachline(joinpath(dir, "manifest.txt"))
        name, ty, r, c = split(line)
        T = ty == "Float32" ? Float32 : Int32
        arrays[name] = open(io -> read!(io, Matrix{T}(undef, parse(Int, r), parse(Int, c))), joinpath(dir, name * ".bin"))
    end
    views = readlines(joinpath(dir, "feature_views.txt"))
    dists = readlines(joinpath(dir, "feature_distributions.txt"))
    conds = filter(!isempty, readlines(joinpath(dir, "sample_conditions.txt")))
    batch_dict = Dict(String(split(k, "__")[2]) => vec(v) for (k, v) in arrays if startswith(k, "batch_of_sample__"))
    model = PM.PathMatFacModel(arrays["data"]; K=K, feature_views=views, feature_distributions=dists,
                               sample_conditions=(isempty(conds) ? nothing : conds),
                               batch_dict=(isempty(batch_dict) ? nothing : batch_dict), kwargs...)
    # the exported columns are already in the constructor's (distribution, view) order: data_idx is the identity
    model.matfac.X .= arrays["X"]; model.matfac.Y .= arrays["Y"]
    model.matfac.col_transform.layers[1].logsigma .= vec(arrays["logsigma"])
    model.matfac.col_transform.layers[3].mu .= vec(arrays["mu"])
    for (slot, key) in ((2, "logdelta__"), (4, "theta__"))
        layer = model.matfac.col_transform.layers[slot]
        isa(layer, Function) && continue
        ba = slot == 2 ? layer.logdelta : layer.theta
        for (v, name) in enumerate(ba.col_range_ids); ba.values[v] .= arrays[key * String(name)]; end
    end
    return model
end
