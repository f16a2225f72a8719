# make_fixtures.jl — reference-held golden vectors for pgopt_b200's parity tests.
#
#   julia --project=/path/to/PGopt/Julia julia/make_fixtures.jl [outdir]        (default: tests/golden)
#
# Runs the UNMODIFIED reference (`using PGopt`; nothing under PGopt/Julia/src is touched) with every random draw of
# the particle filter replaced by pre-drawn numbers, and writes inputs, streams and the reference's outputs as
# tests/golden/ref_<case>/*.npy.  tests/test_independent_sweep.py::test_oracle_against_reference_held_vectors (CPU
# oracle) and tests/test_gpu_parity.py::test_gpu_against_reference_held_vectors (CUDA path) pick the directories up
# automatically.  This is the only route from "parity unpinned" to pinned: the build image has no Julia, so THIS
# SCRIPT HAS NEVER BEEN EXECUTED THERE — if a Julia release changes the Random / Distributions internals it hooks,
# the assertion at the end of `run_case` (all scheduled draws consumed, none left over) fails loudly instead of
# writing wrong vectors.
#
# How the streams get in without editing the reference:
#   * `rand()` / `rand(mvn, n)` in PGopt use Random.default_rng().  The script overrides `Random.default_rng()` to
#     return a ReplayRNG whose `rand(Float64)` / `randn()` pop the pre-drawn uniforms / normals while a sweep's
#     *schedule* is armed, and fall through to an ordinary Xoshiro otherwise (the MNIW draws between sweeps, whose
#     consumption pattern belongs to Distributions.jl's gamma sampler and is not part of the injected-stream contract:
#     the (A, Q) each sweep ran with are read back from PG_samples instead).
#   * `x_init_dist` is a user-supplied argument of particle_Gibbs (particle_Gibbs.jl:114,164).  The script passes an
#     `InitDist` whose `rand` serves the initial-state normals AND arms the schedule of the sweep that is starting — the
#     first `rand(x_init_dist)` of sweep k is the first random call of that sweep (:163-165).
#   * Observable outputs of the unmodified function: PG_samples[k] = (A, Q, w[T,:], x_pf[:,:,T], u[:,T]) for every kept
#     sweep (K_b = 0, k_d = 0: all of them) and the in-place mutated `x_prim` kwarg after the last sweep (:214-217).
#     x_pf[:,:,T] depends on every resampling index of the sweep, so bit-exact index parity is what these pin.
#   * Stream layout = the oracle's (SURVEY.md §8c): Z0[N-1, n_x], Ures[T], Z[T, N, n_x], Uanc[T], Ustar per sweep.
using Random, LinearAlgebra, Distributions, Printf
using PGopt

# ------------------------------------------------------------------------------------------------ replay RNG
mutable struct ReplayRNG <: AbstractRNG
    uni::Vector{Float64}
    nor::Vector{Float64}
    iu::Int
    in_::Int
    fallback::Xoshiro
    served_u::Int
    served_n::Int
end
ReplayRNG() = ReplayRNG(Float64[], Float64[], 0, 0, Xoshiro(12345), 0, 0)
const REPLAY = ReplayRNG()

armed_u(r::ReplayRNG) = r.iu < length(r.uni)
armed_n(r::ReplayRNG) = r.in_ < length(r.nor)
function next_uniform!(r::ReplayRNG)
    r.iu += 1; r.served_u += 1
    return r.uni[r.iu]
end
function next_normal!(r::ReplayRNG)
    r.in_ += 1; r.served_n += 1
    return r.nor[r.in_]
end
# rand() -> rand(default_rng(), Float64) -> this sampler
Random.rand(r::ReplayRNG, ::Random.SamplerTrivial{Random.CloseOpen01{Float64}}) =
    armed_u(r) ? next_uniform!(r) : rand(r.fallback)
Random.randn(r::ReplayRNG) = armed_n(r) ? next_normal!(r) : randn(r.fallback)
Random.randn(r::ReplayRNG, ::Type{Float64}) = randn(r)
# everything else (gamma / chi-square samplers of the MNIW step: randexp, raw bits) comes from the fallback generator
Random.rng_native_52(::ReplayRNG) = UInt64
Random.rand(r::ReplayRNG, ::Random.SamplerType{UInt64}) = rand(r.fallback, UInt64)
Random.randexp(r::ReplayRNG) = randexp(r.fallback)
@eval Random default_rng() = Main.REPLAY
@eval Random default_rng(::Int) = Main.REPLAY

# ------------------------------------------------------------------------------------------------ x_init_dist
mutable struct InitDist
    mean::Vector{Float64}
    L::Matrix{Float64}          # lower Cholesky factor of the covariance
    N::Int
    calls::Int
    arm::Function               # arm(k): load sweep k's streams into REPLAY
end
function Base.rand(d::InitDist)
    if d.calls % (d.N - 1) == 0
        d.arm(div(d.calls, d.N - 1) + 1)
    end
    d.calls += 1
    z = [next_normal!(REPLAY) for _ in 1:length(d.mean)]
    return d.mean + d.L * z
end

# ------------------------------------------------------------------------------------------------ .npy writer
npy_descr(::Type{Float64}) = "<f8"
npy_descr(::Type{Int64}) = "<i8"
function write_npy(path::String, A::AbstractArray{T}) where {T<:Union{Float64,Int64}}
    A = Array(A)
    shape = ndims(A) == 1 ? "($(length(A)),)" : "(" * join(size(A), ", ") * ")"
    hdr = "{'descr': '$(npy_descr(T))', 'fortran_order': True, 'shape': $shape, }"
    total = 10 + length(hdr) + 1
    pad = (64 - total % 64) % 64
    hdr = hdr * " "^pad * "\n"
    open(path, "w") do io
        write(io, UInt8[0x93]); write(io, "NUMPY"); write(io, UInt8[1, 0]); write(io, UInt16(length(hdr)))
        write(io, hdr); write(io, A)
    end
end
write_npy(path::String, x::Real) = write_npy(path, [Float64(x)])

# ------------------------------------------------------------------------------------------------ models
phi_known(x, u) = [0.1 * x[1, :] 0.1 * x[2, :] u[1, :] 0.01 * cos.(3 * x[1, :]) .* x[2, :] 0.1 * sin.(2 * x[2, :]) .* u[1, :]]'  # known example :34
g_meas(x, u) = [1 0] * x                                                                                                     # :54

# phi_sampling of examples/PG_OCP_generic_basis_functions_Altro.jl:52-99, copied from the EXAMPLE script (user code,
# not part of the PGopt package), parametrised by n_phi_dims and L
function make_phi_hilbert(n_phi_dims::Vector{Int}, Lvec::Vector{Float64})
    n_z = length(n_phi_dims); n_phi = prod(n_phi_dims)
    L = zeros(1, 1, n_z); L[1, 1, :] = Lvec
    j_vec = zeros(n_phi, 1, n_z)
    variants = [1; cumprod(n_phi_dims[1:end-1])]
    subscript_values = Array{Int64}(undef, n_z)
    for i in 1:n_phi
        remaining = i - 1
        for j in n_z:-1:1
            subscript_values[j] = floor(remaining / variants[j]) + 1
            remaining = mod(remaining, variants[j])
        end
        j_vec[i, 1, :] = subscript_values
    end
    j_vec = reverse(j_vec, dims=3)                                   # :82
    L_sqrt_inv = 1 ./ sqrt.(L)                                       # :85
    pi_j_over_2L = pi .* j_vec ./ (2 .* L)                           # :86
    function phi_sampling(x, u)                                      # :89-99
        z = vcat(u, x)
        phi = ones(n_phi, size(z, 2))
        for k in axes(z, 1)
            phi .= phi .* (L_sqrt_inv[:, :, k] .* sin.(pi_j_over_2L[:, :, k] * (z[k, :] .+ L[:, :, k])'))
        end
        return phi
    end
    return phi_sampling
end

# ------------------------------------------------------------------------------------------------ one fixture
function run_case(outdir, name; basis, N, T, K, seed, n_phi_dims=Int[], Lvec=Float64[])
    n_x, n_u, n_y = 2, 1, 1
    gen = Xoshiro(seed)
    # data as in the known example :69-98
    f_true(x, u) = [0.8 * x[1] - 0.5 * x[2] + 0.1 * cos(3 * x[1]) * x[2]; 0.4 * x[1] + 0.5 * x[2] + (1 + 0.3 * sin(2 * x[2])) * u]
    Q_true = [0.03 -0.004; -0.004 0.01]; Lt = cholesky(Q_true).L
    u_tr = 3 .* randn(gen, 1, T); y_tr = zeros(1, T); x = [2.0, 2.0] + randn(gen, 2)
    for t in 1:T
        y_tr[1, t] = x[1] + 0.1 * randn(gen)
        x = f_true(x, u_tr[1, t]) + Lt * randn(gen, 2)
    end
    R = 0.1
    if basis == "known"
        phi = phi_known; n_phi = 5
        V = Diagonal(10.0 * ones(n_phi)); A_init = [8.0 -5.0 0.0 10.0 0.0; 4.0 5.0 1.0 0.0 3.0]; Q_init = copy(Q_true)
    else
        phi = make_phi_hilbert(n_phi_dims, Lvec); n_phi = prod(n_phi_dims)
        V = Diagonal(ones(n_phi)); A_init = 2.0 .* randn(gen, n_x, n_phi); Q_init = copy(Q_true)
    end
    Lambda_Q = 100.0 * Matrix(I, n_x, n_x); ell_Q = 10.0
    x0_mean = [2.0, 2.0]; x0_chol = Matrix(1.0I, n_x, n_x)
    # streams, oracle layout, one set per sweep
    Z0 = [randn(gen, n_x, N - 1) for _ in 1:K]          # [:, n] = particle n
    Ures = [rand(gen, T) for _ in 1:K]; Uanc = [rand(gen, T) for _ in 1:K]; Ustar = [rand(gen) for _ in 1:K]
    Z = [randn(gen, n_x, N, T) for _ in 1:K]            # [:, j, t]
    function arm(k)
        @assert !armed_u(REPLAY) && !armed_n(REPLAY) "sweep $(k - 1) left scheduled draws unconsumed"
        first = (k == 1)
        nd = first ? N : N - 1
        uni = Float64[]; nor = Float64[]
        append!(nor, vec(Z0[k]))
        for t in 2:T
            push!(uni, Ures[k][t]); append!(nor, vec(Z[k][:, 1:nd, t]))
            first || push!(uni, Uanc[k][t])
        end
        push!(uni, Ustar[k])
        REPLAY.uni = uni; REPLAY.nor = nor; REPLAY.iu = 0; REPLAY.in_ = 0
    end
    init = InitDist(x0_mean, x0_chol, N, 0, arm)
    x_prim = zeros(n_x, T)
    samples = particle_Gibbs(u_tr, y_tr, K, 0, 0, N, phi, Lambda_Q, ell_Q, Q_init, V, A_init, init, g_meas, R; x_prim=x_prim)
    @assert !armed_u(REPLAY) && !armed_n(REPLAY) "the last sweep left scheduled draws unconsumed"
    @assert init.calls == K * (N - 1)
    dir = joinpath(outdir, "ref_" * name); mkpath(dir)
    w(n, a) = write_npy(joinpath(dir, n * ".npy"), a)
    w("N", N); w("T", T); w("K", K); w("n_x", n_x); w("n_u", n_u); w("n_y", n_y); w("n_phi", n_phi)
    w("basis_is_known", basis == "known" ? 1 : 0)
    w("n_phi_dims", Float64.(n_phi_dims)); w("L", Lvec)
    w("u", u_tr); w("y", y_tr); w("C", [1.0 0.0]); w("R", R); w("x0_mean", x0_mean); w("x0_chol", x0_chol)
    w("V_diag", Vector(diag(V))); w("Lambda_Q", Lambda_Q); w("ell_Q", ell_Q)
    for k in 1:K
        w("Z0_$k", permutedims(Z0[k], (2, 1)))                    # (N-1, n_x)
        w("Z_$k", permutedims(Z[k], (3, 2, 1)))                   # (T, N, n_x)
        w("Ures_$k", Ures[k]); w("Uanc_$k", Uanc[k]); w("Ustar_$k", Ustar[k])
        w("A_$k", samples[k].A); w("Q_$k", samples[k].Q); w("w_m1_$k", samples[k].w_m1); w("x_m1_$k", samples[k].x_m1)
    end
    w("x_prim_final", x_prim)
    @printf("wrote %s (N = %d, T = %d, K = %d; %d uniforms, %d normals replayed)\n", dir, N, T, K, REPLAY.served_u, REPLAY.served_n)
end

outdir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden")
run_case(outdir, "known_N30_T200"; basis="known", N=30, T=200, K=3, seed=1)
run_case(outdir, "hilbert_N30_T100"; basis="hilbert", N=30, T=100, K=3, seed=2, n_phi_dims=[5, 5, 5], Lvec=[10.0, 20.0, 20.0])
run_case(outdir, "known_N4096_T25"; basis="known", N=4096, T=25, K=2, seed=3)
run_case(outdir, "hilbert_N20000_T12"; basis="hilbert", N=20000, T=12, K=2, seed=4, n_phi_dims=[5, 5, 5], Lvec=[10.0, 20.0, 20.0])
