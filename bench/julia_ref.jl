#=
bench/julia_ref.jl — times the UNMODIFIED reference (DiffusionMCMCTools.jl + GuidedProposals / DiffusionDefinition /
ObservationSchemes at the versions of its Manifest.toml) on BASELINE config C1: FitzHugh–Nagumo, one chain, partial observations
of the first coordinate, dt = 1e-3, pCN path update with rho = 0.96.

NEVER EXECUTED IN THIS REPOSITORY: neither the build container nor the GPU box has Julia or network access (BASELINE.md §3).
It exists so that anyone with Julia 1.4 and the pinned packages can produce the true reference number that `bench.py`'s
`cpu_baseline` (a C restatement of the same algorithm, `"kind": "port"`) stands in for.  The set-up and the loop are the
reference's own tutorial code (docs/src/tutorials/preamble.md, docs/src/tutorials/biblock/smoothing.md:25-58), minus plotting.

usage:  julia --project=/path/to/DiffusionMCMCTools.jl bench/julia_ref.jl [num_obs=10] [num_steps=2000]
prints one JSON line in the unit of bench.py: guided EM steps per second.
=#
using GuidedProposals, DiffusionDefinition, ObservationSchemes, DiffusionMCMCTools
const GP = GuidedProposals
const DD = DiffusionDefinition
const OBS = ObservationSchemes
using StaticArrays, Random, Distributions

Random.seed!(100)
@load_diffusion FitzHughNagumo
@load_diffusion FitzHughNagumoAux
DD.var_parameter_names(::FitzHughNagumo) = (:γ,)
DD.var_parameter_names(::FitzHughNagumoAux) = (:γ,)

num_obs = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 10
num_steps = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 2000
dt = 0.001

# data: the tutorial's, cut to `num_obs` observations 0.1 apart (configs[0] of BASELINE.json)
θ = [0.1, -0.8, 1.5, 0.0, 0.3]
P = FitzHughNagumo(θ...)
tt, y1 = 0.0:0.0001:(0.1 * num_obs), @SVector [-0.9, -1.0]
X = rand(P, tt, y1)
obs_scheme = ObsScheme(LinearGsnObs(0.0, (@SVector [0.0]); L=(@SMatrix [1.0 0.0]), Σ=(@SMatrix [0.01])))
data = collect(obs_scheme, X, 1000)
recording = build_recording(P, data, 0.0, KnownStartingPt(y1))

function timed_smoothing(AuxLaw, recording, dt; ρ, num_steps, warmup)
    tts = OBS.setup_time_grids(recording, dt, standard_guid_prop_time_transf)
    sp = SamplingPair(AuxLaw, recording, tts)
    bb = BiBlock(sp, 1:length(recording.obs), ρ, true, num_steps + warmup)
    loglikhd!(bb)
    em_steps_per_sweep = sum(length(t) - 1 for t in tts)
    for i in 1:warmup                       # compile + warm caches
        draw_proposal_path!(bb)
        accept_reject_proposal_path!(bb, i)
    end
    t0 = time_ns()
    for i in (warmup + 1):(warmup + num_steps)
        draw_proposal_path!(bb)
        accept_reject_proposal_path!(bb, i)
    end
    secs = (time_ns() - t0) / 1e9
    em_steps_per_sweep * num_steps / secs, secs / num_steps, accpt_rate(bb, (warmup + 1):(warmup + num_steps))
end

rate, spt, acc = timed_smoothing(FitzHughNagumoAux, recording, dt; ρ=0.96, num_steps=num_steps, warmup=200)
println("{\"impl\": \"reference (Julia)\", \"metric\": \"guided path updates/sec (chains x EM steps/s, FP64)\", \"value\": $rate, ",
        "\"unit\": \"guided EM steps/s\", \"ms_per_step\": $(1e3 * spt), \"threads\": 1, \"accept_rate\": $acc, ",
        "\"config\": {\"workload\": \"C1: FitzHugh-Nagumo, 1 chain, $num_obs observations, dt=$dt, rho=0.96\"}}")
