#=
DiffusionMCMCToolsB200.jl — Julia glue over libdmt.so (include/dmt.h).

NOT EXECUTED IN THIS REPOSITORY'S CI: neither the build container nor the GPU box has Julia (SURVEY.md §0.7), so this file
is reviewed by eye only; the identical sequence of C calls is exercised from Python (diffusionmcmctools.jl_b200/_lib.py,
host.py) in tests/.  It adds device-backed methods to the reference's own generic functions, so a user loop written against
DiffusionMCMCTools.jl (docs/src/tutorials/block_ensemble/inference.md:44-75) runs unchanged on a `DeviceBlockEnsemble`.

Layout conversion: the reference stores one Trajectory per interval with cumulative Wiener paths; the library takes
X[point, dim, chain] and Wiener INCREMENTS W[step, dw, chain] with the chain index fastest (Julia arrays of size
(M, d, NP) / (M, dw, S), column-major).
=#
module DiffusionMCMCToolsB200

using DiffusionMCMCTools
import DiffusionMCMCTools: draw_proposal_path!, accept_reject_proposal_path!, swap_paths!, swap_XX!, swap_WW!, swap_PP!,
    swap_ll!, loglikhd!, loglikhd°!, fetch_ll, fetch_ll°, save_ll!, accpt_rate, ll_of_accepted, find_W_for_X!, set_proposal_law!
import GuidedProposals
const GP = GuidedProposals

const libdmt = get(ENV, "DMT_LIB", joinpath(@__DIR__, "..", "diffusionmcmctools.jl_b200", "libdmt.so"))

# mirrors `dmt_config` (include/dmt.h)
struct DmtConfig
    model::Int32; n_chains::Int32; n_psets::Int32; n_intervals::Int32; obs_dim::Int32; device::Int32
    two_sided_laws::Int32; ll_hist_len::Int32; n_layouts::Int32; chain_offset::Int32; seed::UInt64; artificial_noise::Float64
end

struct DmtError <: Exception
    code::Int32
    msg::String
end

function check(ctx::Ptr{Cvoid}, rc::Int32)
    rc == 0 && return
    msg = unsafe_string(ccall((:dmt_last_error, libdmt), Cstring, (Ptr{Cvoid},), ctx))
    throw(DmtError(rc, msg))
end

"""
    DeviceEnsemble(model_id, theta, L, Σ, v, xbar, x0, n_pts, tt; kwargs...)

Device-resident `SamplingEnsemble` (src/sampling_ensemble.jl:17-41).  `v::Array{Float64,3}` is (P, m, K), `xbar` (P, d, K),
`x0` (M, d): chain / pset index fastest.
"""
mutable struct DeviceEnsemble
    ctx::Ptr{Cvoid}
    M::Int; P::Int; K::Int; d::Int; dw::Int
    next_layout::Int32
    theta::Matrix{Float64}      # (P, npar) accepted
    theta°::Matrix{Float64}
    xbar::Array{Float64,3}
end

function DeviceEnsemble(model::Integer, theta::Matrix{Float64}, L::Matrix{Float64}, Σ::Matrix{Float64}, v::Array{Float64,3},
                        xbar::Array{Float64,3}, x0::Matrix{Float64}, n_pts::Vector{Int32}, tt::Vector{Float64};
                        device=0, seed=UInt64(0), two_sided_laws=true, max_layouts=8, chain_offset=0, artificial_noise=1e-11)
    M, d = size(x0); P = size(v, 1); m = size(L, 1); K = length(n_pts)
    cfg = Ref(DmtConfig(model, M, P, K, m, device, two_sided_laws, 0, max_layouts, chain_offset, seed, artificial_noise))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:dmt_create, libdmt), Int32, (Ref{DmtConfig}, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}, Ref{Ptr{Cvoid}}),
               cfg, n_pts, tt, C_NULL, out)
    rc == 0 || throw(DmtError(rc, unsafe_string(ccall((:dmt_last_error, libdmt), Cstring, (Ptr{Cvoid},), C_NULL))))
    ctx = out[]
    dw = Ref{Int32}(0)
    ccall((:dmt_model_dims, libdmt), Int32, (Int32, Ptr{Int32}, Ref{Int32}, Ptr{Int32}, Ptr{Int32}), model, C_NULL, dw, C_NULL, C_NULL)
    # per-interval broadcast of L and Σ: [k][m*d][P] with P fastest
    Lk = repeat(reshape(permutedims(L), 1, d * m, 1), P, 1, K)          # row-major L flattened, (P, m*d, K)
    Σk = repeat(reshape(permutedims(Σ), 1, m * m, 1), P, 1, K)
    for side in (two_sided_laws ? (0, 1) : (0,))
        check(ctx, ccall((:dmt_set_params, libdmt), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Int32, Ptr{Float64}), ctx, side, 3, 0, K - 1, theta))
        check(ctx, ccall((:dmt_set_obs, libdmt), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                         ctx, side, 0, K - 1, Lk, Σk, v))
        for store in (0, 1)   # aux_laws_blocking = aux_laws (src/sampling_unit.jl:57)
            check(ctx, ccall((:dmt_set_aux_linearised, libdmt), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Int32, Ptr{Float64}),
                             ctx, side, store, 0, K - 1, xbar))
        end
    end
    check(ctx, ccall((:dmt_set_start, libdmt), Int32, (Ptr{Cvoid}, Ptr{Float64}), ctx, x0))
    se = DeviceEnsemble(ctx, M, P, K, d, dw[], 0, copy(theta), copy(theta), xbar)
    finalizer(s -> ccall((:dmt_destroy, libdmt), Int32, (Ptr{Cvoid},), s.ctx), se)
    se
end

"""
    DeviceBlockEnsemble(se, block_ranges, ρρ=0.0, ll_hist_len=0)

`BlockEnsemble` (src/block_ensemble.jl:17-34) as one registered layout; `block_ranges` are the reference's 1-based UnitRanges.
"""
struct DeviceBlockEnsemble
    se::DeviceEnsemble
    layout::Int32
    n_blocks::Int
    ll_hist_len::Int
end

function DeviceBlockEnsemble(se::DeviceEnsemble, block_ranges, ρρ=0.0, ll_hist_len=0)
    nb = length(block_ranges)
    i0 = Int32[first(r) - 1 for r in block_ranges]; i1 = Int32[last(r) - 1 for r in block_ranges]
    ρ = ρρ isa Number ? fill(Float64(ρρ), nb) : Vector{Float64}(ρρ)
    layout = se.next_layout; se.next_layout += 1
    check(se.ctx, ccall((:dmt_set_blocks, libdmt), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{UInt8}, Int32),
                        se.ctx, layout, nb, i0, i1, ρ, C_NULL, ll_hist_len))
    DeviceBlockEnsemble(se, layout, nb, ll_hist_len)
end

const DBE = DeviceBlockEnsemble

# ---- imputation (src/block_ensemble.jl:50,63-67).  mcmciter is 1-based in the reference, 0-based in the library.
draw_proposal_path!(be::DBE, mcmciter::Integer=1) =
    check(be.se.ctx, ccall((:dmt_draw_proposal_path, libdmt), Int32, (Ptr{Cvoid}, Int32, UInt32, Ptr{Float64}), be.se.ctx, be.layout, mcmciter - 1, C_NULL))
accept_reject_proposal_path!(be::DBE, mcmciter::Integer) =
    check(be.se.ctx, ccall((:dmt_accept_reject_path, libdmt), Int32, (Ptr{Cvoid}, Int32, UInt32, Ptr{Float64}), be.se.ctx, be.layout, mcmciter - 1, C_NULL))

# ---- swaps (src/block_ensemble.jl:79-112)
_swap(be::DBE, what) = check(be.se.ctx, ccall((:dmt_swap, libdmt), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), be.se.ctx, be.layout, what, C_NULL))
swap_XX!(be::DBE) = _swap(be, 1)
swap_WW!(be::DBE) = _swap(be, 2)
swap_paths!(be::DBE) = _swap(be, 3)
swap_ll!(be::DBE) = _swap(be, 8)
function swap_PP!(be::DBE)
    _swap(be, 4)
    be.se.theta, be.se.theta° = be.se.theta°, be.se.theta
    nothing
end

# ---- utility (src/block_ensemble.jl:121-179)
loglikhd!(be::DBE; skip=0) = check(be.se.ctx, ccall((:dmt_loglikhd, libdmt), Int32, (Ptr{Cvoid}, Int32, Int32, Int32), be.se.ctx, be.layout, 0, skip))
loglikhd°!(be::DBE; skip=0) = check(be.se.ctx, ccall((:dmt_loglikhd, libdmt), Int32, (Ptr{Cvoid}, Int32, Int32, Int32), be.se.ctx, be.layout, 1, skip))
function _stats(be::DBE)
    out = Vector{Float64}(undef, 2 + be.n_blocks)
    check(be.se.ctx, ccall((:dmt_allreduce_stats, libdmt), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}), be.se.ctx, be.layout, out))
    out
end
fetch_ll(be::DBE) = _stats(be)[1]
fetch_ll°(be::DBE) = _stats(be)[2]
save_ll!(be::DBE, i::Integer) = check(be.se.ctx, ccall((:dmt_save_ll, libdmt), Int32, (Ptr{Cvoid}, Int32, UInt32), be.se.ctx, be.layout, i - 1))
function accpt_rate(be::DBE, range)
    counts = Vector{Int64}(undef, be.n_blocks)
    check(be.se.ctx, ccall((:dmt_accept_counts, libdmt), Int32, (Ptr{Cvoid}, Int32, UInt32, UInt32, Ptr{Int64}),
                           be.se.ctx, be.layout, first(range) - 1, last(range) - 1, counts))
    counts ./ (length(range) * be.se.M)
end
function ll_of_accepted(be::DBE, i::Integer)
    n = be.n_blocks * be.se.M
    acc = Vector{UInt8}(undef, n); ll = Vector{Float64}(undef, n); ll° = Vector{Float64}(undef, n)
    check(be.se.ctx, ccall((:dmt_get_accept_history, libdmt), Int32, (Ptr{Cvoid}, Int32, UInt32, UInt32, Ptr{UInt8}), be.se.ctx, be.layout, i - 1, i - 1, acc))
    check(be.se.ctx, ccall((:dmt_get_ll_history, libdmt), Int32, (Ptr{Cvoid}, Int32, Int32, UInt32, UInt32, Ptr{Float64}), be.se.ctx, be.layout, 0, i - 1, i - 1, ll))
    check(be.se.ctx, ccall((:dmt_get_ll_history, libdmt), Int32, (Ptr{Cvoid}, Int32, Int32, UInt32, UInt32, Ptr{Float64}), be.se.ctx, be.layout, 1, i - 1, i - 1, ll°))
    reshape(ifelse.(acc .!= 0, ll°, ll), be.se.M, be.n_blocks)
end

# ---- blocking (src/block_ensemble.jl:192-221)
GP.set_obs!(be::DBE) = check(be.se.ctx, ccall((:dmt_set_artificial_obs, libdmt), Int32, (Ptr{Cvoid}, Int32), be.se.ctx, be.layout))
_rgt(be::DBE, which) = check(be.se.ctx, ccall((:dmt_recompute_guiding_term, libdmt), Int32, (Ptr{Cvoid}, Int32, Int32), be.se.ctx, be.layout, which))
GP.recompute_guiding_term!(be::DBE) = _rgt(be, 3)
GP.recompute_guiding_term!(be::DBE, ::Val{:P_only}) = _rgt(be, 1)
GP.recompute_guiding_term!(be::DBE, ::Val{:P°_only}) = _rgt(be, 2)
find_W_for_X!(be::DBE) = check(be.se.ctx, ccall((:dmt_find_W_for_X, libdmt), Int32, (Ptr{Cvoid}, Int32), be.se.ctx, be.layout))

# ---- the blocking step of docs/src/tutorials/block_collection/inference_with_blocking.md as one call:
# set_obs!; recompute_guiding_term!(Val(:P_only)); find_W_for_X!; loglikhd!; draw_proposal_path!
blocking_sweep!(be::DBE, mcmciter::Integer) =
    check(be.se.ctx, ccall((:dmt_blocking_sweep, libdmt), Int32, (Ptr{Cvoid}, Int32, UInt32), be.se.ctx, be.layout, mcmciter - 1))
# while θ, the auxiliary laws and the real observations stay fixed, K1 after set_obs! is F = F⁰ + Ψv (exact); see include/dmt.h
enable_guiding_cache!(be::DBE, on::Bool=true) =
    check(be.se.ctx, ccall((:dmt_enable_guiding_cache, libdmt), Int32, (Ptr{Cvoid}, Int32, Int32), be.se.ctx, be.layout, on))

# lanes per (chain, block) in the forward kernel: 0 = automatic (results never depend on it)
set_fwd_lanes!(se::DeviceEnsemble, lanes::Integer) =
    check(se.ctx, ccall((:dmt_set_fwd_lanes, libdmt), Int32, (Ptr{Cvoid}, Int32), se.ctx, lanes))
# multi-GPU, one process per GPU: export the 64-byte handle, all-gather it (MPI.Allgather, Distributed, ...), map the peers
p2p_export(se::DeviceEnsemble) = (h = zeros(UInt8, 64);
    check(se.ctx, ccall((:dmt_p2p_export, libdmt), Int32, (Ptr{Cvoid}, Ptr{UInt8}), se.ctx, h)); h)
p2p_init!(se::DeviceEnsemble, n_ranks::Integer, rank::Integer, handles::Matrix{UInt8}) =   # handles: 64 x n_ranks
    check(se.ctx, ccall((:dmt_p2p_init, libdmt), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), se.ctx, n_ranks, rank, handles))

# ---- parameters (src/block_ensemble.jl:242-255 -> src/biblock.jl:334-371).  The name translation of
# src/param_names_collections.jl stays here on the host: `pnames` is its result for the target law,
# a vector of (index into θ°) => (index into the model's parameter vector) pairs.
function set_proposal_law!(be::DBE, θ°, pnames::Vector{Pair{Int,Int}}, critical_change::Bool; skip=0)
    se = be.se
    th = copy(se.theta)
    for (i, j) in pnames
        th[:, j] .= θ°[i]
    end
    se.theta° = th
    check(se.ctx, ccall((:dmt_equalize_laws, libdmt), Int32, (Ptr{Cvoid}, Int32, Int32, Int32), se.ctx, 3, 0, se.K - 1))
    check(se.ctx, ccall((:dmt_set_params, libdmt), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Int32, Ptr{Float64}), se.ctx, 1, 3, 0, se.K - 1, th))
    if critical_change
        for store in (0, 1)
            check(se.ctx, ccall((:dmt_set_aux_linearised, libdmt), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Int32, Ptr{Float64}),
                                se.ctx, 1, store, 0, se.K - 1, se.xbar))
        end
    end
    check(se.ctx, ccall((:dmt_set_proposal_law, libdmt), Int32, (Ptr{Cvoid}, Int32, Int32, Int32), se.ctx, be.layout, critical_change, skip))
end

# ---- paths back into the reference's containers: X[point, dim, chain] -> Vector{Vector{SVector}} per interval
function get_paths(se::DeviceEnsemble, side::Integer=0)
    NP = Ref{Int}(0)
    error("get_paths: allocate (M, d, NP) and call dmt_get_X; conversion to Trajectory left to the caller's container types")
end

export DeviceEnsemble, DeviceBlockEnsemble

end # module
