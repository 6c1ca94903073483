# Julia-side binding of libsimplemcmc_cuda.so (include/simplemcmc.h) -- the ccall stubs a SimpleMCMC.jl maintainer adds.
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  The same call sequence is exercised from Python
# through ctypes (simplemcmc.jl_b200/capi.py), which is what the test-suite verifies.
#
# Where it plugs into the reference (paths relative to the SimpleMCMC.jl tree):
#   src/parsing.jl:488-492  instead of `eval`-ing the generated closure, generateModelFunction serialises
#                           `m.exprs ++ m.dexprs` (+ shapes from vhint) into a tape (include/simplemcmc_tape.h) and
#                           returns  beta -> smc_eval(model, beta)
#   src/samplers.jl:146-167 simpleHMC's loop  -> one call of hmc_run
#   src/samplers.jl:89-113  simpleRWM's loop  -> one call of rwm_run
#   src/samplers.jl:263-311 simpleNUTS's loop -> one call of nuts_run
#   src/distribs.jl:37-81   the dlopen("libRmath") block disappears: the densities are device functions
module SimpleMCMCCuda

const lib = get(ENV, "SIMPLEMCMC_CUDA_LIB", "libsimplemcmc_cuda")

struct SMCError <: Exception
    status::Cint
    msg::String
end

lasterror() = unsafe_string(ccall((:smc_last_error, lib), Cstring, ()))
check(rc::Cint) = rc == 0 ? nothing : throw(SMCError(rc, lasterror()))

mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer=0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:smc_ctx_create, lib), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h))
        c = new(h[])
        finalizer(c -> ccall((:smc_ctx_destroy, lib), Cint, (Ptr{Cvoid},), c.h), c)
        c
    end
end

# ---- multi-GPU plumbing of row-sharded models (one Julia process per GPU; the host exchanges the byte strings) ----------
"128-byte NCCL unique id (made by the first rank of a communicator, broadcast by the host)"
function comm_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:smc_comm_unique_id, lib), Cint, (Ptr{UInt8}, Cint), id, 128))
    id
end
comm_init!(ctx::Context, id::Vector{UInt8}, rank::Integer, nranks::Integer) =
    check(ccall((:smc_ctx_comm_init, lib), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint, Cint), ctx.h, id, length(id), rank, nranks))
"optional, one node: 64-byte IPC handle of this rank's peer-exchange buffer / map the peers' buffers (handles in rank order)"
function p2p_export(ctx::Context, nranks::Integer)
    h = Vector{UInt8}(undef, 64)
    check(ccall((:smc_ctx_p2p_export, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{UInt8}, Cint), ctx.h, nranks, h, 64))
    h
end
p2p_import!(ctx::Context, handles::Vector{UInt8}, nranks::Integer, rank::Integer) =
    check(ccall((:smc_ctx_p2p_import, lib), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), ctx.h, handles, nranks, rank))

mutable struct Model
    h::Ptr{Cvoid}
    nparams::Int
    ctx::Context
end

"compile a serialised tape, bind the external variables (Dict name => Array{Float64}), finalize for `chains` chains"
function Model(ctx::Context, tape::Vector{UInt8}, data::Dict{String,<:Any}; chains::Integer=1, row_shard::Bool=false)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:smc_model_compile, lib), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Csize_t, Ref{Ptr{Cvoid}}), ctx.h, tape, length(tape), h))
    for (name, val) in data
        a = Array{Float64}(val isa Number ? fill(Float64(val)) : val)        # column major already
        dims = Int64[size(a)...]
        check(ccall((:smc_model_bind_data, lib), Cint, (Ptr{Cvoid}, Cstring, Ptr{Float64}, Ptr{Int64}, Cint),
                    h[], name, a, dims, ndims(a)))
    end
    check(ccall((:smc_model_finalize, lib), Cint, (Ptr{Cvoid}, Int64, Cint), h[], chains, row_shard))
    n = Ref{Int64}(0)
    check(ccall((:smc_model_nparams, lib), Cint, (Ptr{Cvoid}, Ref{Int64}), h[], n))
    m = Model(h[], n[], ctx)
    finalizer(m -> ccall((:smc_model_destroy, lib), Cint, (Ptr{Cvoid},), m.h), m)
    m
end

"the closure generateModelFunction returns: beta::Vector -> (logp, grad)   (C chains: beta is M x C, column = chain)"
function evaluate(m::Model, beta::VecOrMat{Float64}; gradient::Bool=true)
    C = size(beta, 2)
    logp = Vector{Float64}(undef, C)
    grad = gradient ? similar(beta) : Float64[]
    # a Julia M x C matrix is [C][M] row-major in C terms: exactly the ABI layout
    check(ccall((:smc_eval, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}),
                m.h, beta, C, logp, gradient ? pointer(grad) : C_NULL))
    beta isa Vector ? (gradient ? (logp[1], grad) : logp[1]) : (gradient ? (logp, grad) : logp)
end

struct SamplerConfig            # smc_sampler_config
    steps::Int64; burnin::Int64; isteps::Int64; stepsize::Float64; seed::UInt64; chain_offset::Int64
    store_draws::Int32; init_per_chain::Int32
    inject_normals::Ptr{Float64}; inject_uniforms::Ptr{Float64}; inject_uniforms_per_step::Int64
end

mutable struct RunOutput        # smc_run_output
    draws::Ptr{Float64}; loglik::Ptr{Float64}; accept::Ptr{Float64}; final_state::Ptr{Float64}; accept_rate::Ptr{Float64}
    mean::Ptr{Float64}; m2::Ptr{Float64}; epsilon::Ptr{Float64}; jmax::Ptr{Float64}
    device_ms::Float64; n_evals::Int64; n_leapfrogs::Int64; n_kernel_launches::Int64
    first::Ptr{Float64}; lag1::Ptr{Float64}       # ABI 2: lag-1 statistics for calcStats! (ESS) without the draws
end

for (jl, sym) in ((:hmc_run, :smc_hmc_run), (:rwm_run, :smc_rwm_run), (:nuts_run, :smc_nuts_run))
    @eval function $jl(m::Model, init::Vector{Float64}; steps=1000, burnin=100, isteps=2, stepsize=1e-3, chains=1, seed=1)
        M, S = m.nparams, steps - burnin
        draws  = Array{Float64}(undef, M, chains, S)        # = [samples][C][M] in C order
        loglik = Array{Float64}(undef, chains, S)
        accept = Array{Float64}(undef, chains, S)
        rate   = Vector{Float64}(undef, chains)
        eps    = Array{Float64}(undef, chains, steps)
        jmax   = Array{Float64}(undef, chains, steps)
        final, mean, m2, first, lag1 = (Array{Float64}(undef, M, chains) for _ in 1:5)   # = [C][M] in C order
        cfg = Ref(SamplerConfig(steps, burnin, isteps, stepsize, seed, 0, 1, 0, C_NULL, C_NULL, 0))
        out = RunOutput(pointer(draws), pointer(loglik), pointer(accept), pointer(final), pointer(rate), pointer(mean), pointer(m2),
                        pointer(eps), pointer(jmax), 0.0, 0, 0, 0, pointer(first), pointer(lag1))
        GC.@preserve draws loglik accept rate eps jmax final mean m2 first lag1 begin
            check(ccall(($(QuoteNode(sym)), lib), Cint, (Ptr{Cvoid}, Ref{SamplerConfig}, Ptr{Float64}, Int64, Ref{RunOutput}),
                        m.h, cfg, init, chains, out))
        end
        # calcStats! (src/samplers.jl:352-374) needs only mean, m2, first, lag1 and final (see api.py::_finish for the formula)
        (draws=draws, loglik=loglik, accept=accept, accept_rate=rate, epsilon=eps, jmax=jmax, final_state=final,
         mean=mean, m2=m2, first=first, lag1=lag1, device_ms=out.device_ms, n_leapfrogs=out.n_leapfrogs)
    end
end

end # module
