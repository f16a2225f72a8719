# PGoptB200.jl — thin `ccall` shim that keeps PGopt's Julia signatures for the hot path and runs them on
# libpgopt_b200.so (hand-written sm_100a CUDA behind a C ABI; include/pgopt_b200.h).
#
# NOT EXECUTED IN THE BUILD ENVIRONMENT: neither Julia nor a network exists there (SURVEY.md §1 fact 2), so this
# file is the binding a PGopt maintainer would add; the same entry points are exercised by the Python ctypes
# mirror pgopt_b200/api.py, which the test-suite runs on a B200.
#
# Drop-in use:   using PGoptB200   instead of   using PGopt   for
#   particle_Gibbs, systematic_resampling, MNIW_sample, PG_sample, scenario_rollout
# and hand scenario_rollout's outputs to PGopt.solve_PG_OCP_Ipopt(...; x_vec_0, v_vec, e_vec, u_init).
module PGoptB200

using LinearAlgebra
using Distributions
using Printf

export PG_sample, KnownBasisVB, HilbertGPBasis, LinearMeasurement, particle_Gibbs, systematic_resampling,
       MNIW_sample, scenario_rollout, test_prediction, scenario_constraint_max

const LIB = get(ENV, "PGOPT_B200_LIB", "libpgopt_b200.so")
const PGB_MAX_Z = 8

# Julia/src/PGopt.jl:23-29
mutable struct PG_sample
    A::Matrix{Float64}
    Q::Matrix{Float64}
    w_m1::Array{Float64}
    x_m1::Array{Float64}
    u_m1::Array{Float64}
end

# Julia closures cannot cross a C ABI: phi and g are descriptors.
struct KnownBasisVB end                      # examples/PG_OCP_known_basis_functions_Altro.jl:34
struct HilbertGPBasis                        # examples/PG_OCP_generic_basis_functions_Altro.jl:52-99
    n_phi_dims::Vector{Int}                  # [n_phi_u..., n_phi_x...]
    L::Vector{Float64}                       # [L_u..., L_x...]
    n_u::Int
end
struct LinearMeasurement                     # g(x,u) = C*x   (known example :54)
    C::Matrix{Float64}
end

# mirror of pgb_config (include/pgopt_b200.h)
struct PgbConfig
    device::Int32; n_x::Int32; n_u::Int32; n_y::Int32; n_phi::Int32; basis::Int32
    hg_count::NTuple{PGB_MAX_Z,Int32}; hg_stride::NTuple{PGB_MAX_Z,Int32}; hg_L::NTuple{PGB_MAX_Z,Float64}
    N::Int64; T::Int64; n_chains::Int32; keep_w_history::Int32
    x_history::Int32   # 0 auto, 1 store x_pf, 2 replay (two rows + recomputed lineage)
    precision::Int32
end

pad8(v, T) = ntuple(i -> i <= length(v) ? T(v[i]) : zero(T), PGB_MAX_Z)

function make_config(phi, n_y, N, T; device=0, n_chains=1, precision=0)   # precision: 0 = Float64 (reference), 1 = PGB_PRECISION_F32
    if phi isa KnownBasisVB
        return PgbConfig(device, 2, 1, n_y, 5, 0, pad8(Int[], Int32), pad8(Int[], Int32), pad8(Float64[], Float64),
                         N, T, n_chains, 0, 0, precision), 2, 1, 5
    elseif phi isa HilbertGPBasis
        nz = length(phi.L)
        # the reference reverses the index vectors (:82): dimension k uses slot nz+1-k
        count = [phi.n_phi_dims[nz+1-k] for k in 1:nz]
        stride = [prod(phi.n_phi_dims[1:nz-k]) for k in 1:nz]
        n_phi = prod(phi.n_phi_dims)
        return PgbConfig(device, nz - phi.n_u, phi.n_u, n_y, n_phi, 1, pad8(count, Int32), pad8(stride, Int32),
                         pad8(phi.L, Float64), N, T, n_chains, 0, 0, precision), nz - phi.n_u, phi.n_u, n_phi
    else
        throw(ArgumentError("phi must be a KnownBasisVB or HilbertGPBasis descriptor: arbitrary closures cannot " *
                            "cross the C ABI and there is no CPU fallback"))
    end
end

function check(rc, h=C_NULL)
    rc == 0 && return
    msg = unsafe_string(ccall((:pgb_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    rc == 3 && throw(PosDefException(0))
    error("pgopt_b200 error $rc: $msg")
end

"""
    systematic_resampling(W, N)          # Julia/src/particle_Gibbs.jl:17-34
"""
function systematic_resampling(W, N; rnd=rand(), device=0)
    Wv = Vector{Float64}(W)
    idx = Vector{Int64}(undef, N)
    st = Ref{Int32}(0)
    check(ccall((:pgb_systematic_resampling, LIB), Cint,
                (Int32, Ptr{Float64}, Int64, Int64, Float64, Ptr{Int64}, Ref{Int32}),
                device, Wv, length(Wv), N, rnd, idx, st))
    (st[] & 1) != 0 && throw(BoundsError(Wv, length(Wv) + 1))   # what :28 raises
    return idx
end

# V of the MN prior (particle_Gibbs.jl:56, :64): a Diagonal / diagonal matrix goes down the diagonal path (1 ./ V.diag, as
# both examples), any other symmetric positive definite matrix down the dense one (pgb_*_dense: I / V by Cholesky, once)
function prior_V(V)
    V isa Diagonal && return false, Vector{Float64}(V.diag)
    Vm = Matrix{Float64}(V)
    isdiag(Vm) && return false, Vector{Float64}(diag(Vm))
    return true, vec(Vm)
end

"""
    MNIW_sample(Phi, Psi, Sigma, V, Lambda_Q, ell_Q, T)      # Julia/src/particle_Gibbs.jl:56-81
"""
function MNIW_sample(Phi, Psi, Sigma, V, Lambda_Q, ell_Q, T; seed=rand(UInt64), chain=0, sweep_index=1, device=0)
    n_x, n_phi = size(Psi)
    dense, Vd = prior_V(V)
    A = Matrix{Float64}(undef, n_x, n_phi); Q = Matrix{Float64}(undef, n_x, n_x)
    argt = (Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Int64,
            Ptr{Float64}, Ptr{Float64}, UInt64, UInt32, UInt32, Ptr{Float64}, Ptr{Float64})
    call(args...) = dense ? ccall((:pgb_mniw_sample_dense, LIB), Cint, argt, args...) :
                            ccall((:pgb_mniw_sample, LIB), Cint, argt, args...)
    check(call(device, n_x, n_phi, Matrix{Float64}(Phi), Matrix{Float64}(Psi), Matrix{Float64}(Sigma), Vd,
                Matrix{Float64}(Lambda_Q * I(n_x)), Float64(ell_Q), T, C_NULL, C_NULL, seed, chain, sweep_index, A, Q))
    return A, Q
end

"""
    particle_Gibbs(u_training, y_training, K, K_b, k_d, N, phi, Lambda_Q, ell_Q, Q_init, V, A_init, x_init_dist, g, R; x_prim=nothing)

Same signature as Julia/src/particle_Gibbs.jl:114; `phi` / `g` are descriptors.
"""
function particle_Gibbs(u_training, y_training, K, K_b, k_d, N, phi, Lambda_Q, ell_Q, Q_init, V, A_init, x_init_dist,
                        g::LinearMeasurement, R; x_prim=nothing, seed=rand(UInt64), chain=0, device=0, precision=0)
    n_y, T = size(y_training)
    cfg, n_x, n_u, n_phi = make_config(phi, n_y, N, T; device=device, precision=precision)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pgb_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ref{PgbConfig}), h, cfg))
    try
        u = Matrix{Float64}(u_training); y = Matrix{Float64}(y_training)
        check(ccall((:pgb_set_data, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h[], u, y), h[])
        Rm = R isa Number ? fill(Float64(R), 1, 1) : Matrix{Float64}(R)
        check(ccall((:pgb_set_measurement, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h[], g.C, Rm), h[])
        mu = Vector{Float64}(mean(x_init_dist)); Lc = Matrix{Float64}(cholesky(Matrix(cov(x_init_dist))).L)
        check(ccall((:pgb_set_init_dist, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h[], mu, Lc), h[])
        dense, Vd = prior_V(V)
        if dense
            check(ccall((:pgb_set_prior_dense, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64),
                        h[], Vd, Matrix{Float64}(Lambda_Q * I(n_x)), Float64(ell_Q)), h[])
        else
            check(ccall((:pgb_set_prior, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64),
                        h[], Vd, Matrix{Float64}(Lambda_Q * I(n_x)), Float64(ell_Q)), h[])
        end
        xp = x_prim === nothing ? C_NULL : Matrix{Float64}(x_prim)
        check(ccall((:pgb_set_state, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64),
                    h[], 0, Matrix{Float64}(A_init), Matrix{Float64}(Q_init * I(n_x)), xp, 1), h[])
        check(ccall((:pgb_set_rng_philox, LIB), Cint, (Ptr{Cvoid}, UInt64, UInt32), h[], seed, chain), h[])
        A_s = Array{Float64}(undef, n_x, n_phi, K); Q_s = Array{Float64}(undef, n_x, n_x, K)
        w_s = Array{Float64}(undef, N, K); x_s = Array{Float64}(undef, n_x, N, K); u_m1 = Vector{Float64}(undef, n_u)
        println("### Started learning algorithm")
        t0 = time()
        check(ccall((:pgb_particle_gibbs, LIB), Cint,
                    (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    h[], K, K_b, k_d, A_s, Q_s, w_s, x_s, u_m1), h[])
        if x_prim !== nothing       # the reference mutates x_prim in place (:214-217)
            xo = Matrix{Float64}(undef, n_x, T)
            check(ccall((:pgb_get_state, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
                        h[], 0, C_NULL, C_NULL, xo, C_NULL), h[])
            x_prim .= xo
        end
        println("### Learning complete\nRuntime: ", round(time() - t0; digits=2), " s")
        return [PG_sample(A_s[:, :, k], Q_s[:, :, k], w_s[:, k], x_s[:, :, k], copy(u_m1)) for k in 1:K]
    finally
        ccall((:pgb_destroy, LIB), Cvoid, (Ptr{Cvoid},), h[])
    end
end

"""
    scenario_rollout(PG_samples, phi, g, R, H; u_init=zeros(n_u, H))

The sampling + forward rollout solve_PG_OCP_Ipopt does before building the NLP
(Julia/src/optimal_control_Ipopt.jl:47-83,114-125); returns (x_vec_0, v_vec, e_vec, X_init, Y_init) in the
reference's layouts — pass the first three back as keyword arguments of solve_PG_OCP_Ipopt.
"""
function scenario_rollout(PG_samples::Vector{PG_sample}, phi, g::LinearMeasurement, R, H; u_init=nothing,
                          x_vec_0=nothing, v_vec=nothing, e_vec=nothing,   # the reference's kwargs (:37, :48, :67, :77)
                          seed=rand(UInt64), k_offset=0, device=0)
    K = length(PG_samples); N = length(PG_samples[1].w_m1); n_y = size(g.C, 1)
    cfg, n_x, n_u, n_phi = make_config(phi, n_y, N, 2; device=device)
    u0 = u_init === nothing ? zeros(n_u, H) : Matrix{Float64}(u_init)
    # MvNormal(zeros(n_y), R::Real) reads a scalar as a standard deviation (reference quirk, :79)
    Le = R isa Number ? fill(Float64(R), 1, 1) : Matrix{Float64}(cholesky(Matrix(R)).L)
    A = cat([s.A for s in PG_samples]...; dims=3); Q = cat([s.Q for s in PG_samples]...; dims=3)
    w = hcat([s.w_m1 for s in PG_samples]...); x = cat([s.x_m1 for s in PG_samples]...; dims=3)
    um = hcat([s.u_m1 for s in PG_samples]...)
    x0 = Array{Float64}(undef, n_x, K); v = Array{Float64}(undef, n_x, H, K); e = Array{Float64}(undef, n_y, H, K)
    X = Array{Float64}(undef, n_x, H + 1, K); Y = Array{Float64}(undef, n_y, H, K)
    opt(a) = a === nothing ? C_NULL : Array{Float64}(a)   # (n_x*K,), (n_x,H,K), (n_y,H,K): column-major = the ABI's layouts
    check(ccall((:pgb_rollout_supplied, LIB), Cint,
                (Ref{PgbConfig}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, UInt64, UInt32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                cfg, K, H, A, Q, w, x, um, g.C, Le, u0, seed, k_offset, opt(x_vec_0), opt(v_vec), opt(e_vec), x0, v, e, X, Y))
    X_init = reshape(permutedims(X, (1, 3, 2)), n_x * K, H + 1)
    Y_init = reshape(permutedims(Y, (1, 3, 2)), n_y * K, H)
    return vec(x0), v, e, X_init, Y_init
end

"""
    test_prediction(PG_samples, phi, g, R, k_n, u_test, y_test)

Julia/src/particle_Gibbs.jl:253-314 on the rollout kernel: every model simulated `k_n` times over the test input.
Returns (x_test_sim (n_x, T_test+1, K*k_n), y_test_sim (n_y, T_test, K*k_n), mean_rmse) — the arrays the reference
builds at :301-302 and the value it prints at :311-312.  Plotting (`plot_predictions`, :308) stays with the caller.
"""
function test_prediction(PG_samples::Vector{PG_sample}, phi, g::LinearMeasurement, R, k_n, u_test, y_test;
                         seed=rand(UInt64), device=0)
    K = length(PG_samples); N = length(PG_samples[1].w_m1); n_y = size(y_test, 1); T_test = size(y_test, 2)
    cfg, n_x, n_u, n_phi = make_config(phi, n_y, N, 2; device=device)
    Le = R isa Number ? fill(Float64(R), 1, 1) : Matrix{Float64}(cholesky(Matrix(R)).L)   # scalar R: a std (:261)
    A = cat([s.A for s in PG_samples]...; dims=3); Q = cat([s.Q for s in PG_samples]...; dims=3)
    w = hcat([s.w_m1 for s in PG_samples]...); x = cat([s.x_m1 for s in PG_samples]...; dims=3)
    um = hcat([s.u_m1 for s in PG_samples]...)
    ut = Matrix{Float64}(u_test[:, 1:T_test]); yt = Matrix{Float64}(y_test)
    x_sim = Array{Float64}(undef, n_x, T_test + 1, K * k_n); y_sim = Array{Float64}(undef, n_y, T_test, K * k_n)
    rmse = Ref{Float64}(0.0)
    check(ccall((:pgb_test_prediction, LIB), Cint,
                (Ref{PgbConfig}, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, UInt64, Ptr{Float64}, Ptr{Float64}, Ref{Float64}),
                cfg, K, k_n, T_test, A, Q, w, x, um, g.C, Le, ut, yt, seed, x_sim, y_sim, rmse))
    @printf("Mean rmse: %.2f\n", rmse[])
    return x_sim, y_sim, rmse[]
end

"""
    scenario_constraint_max(y_opt, y_min, y_max)

The scenario ordering of the greedy support-sub-sample search (Julia/src/optimal_control_Ipopt.jl:334-338,
optimal_control_Altro.jl:496-500) for the examples' box output constraints: returns
(h_scenario_max::Vector{Float64}, scenarios_sorted::Vector{Int}) with `scenarios_sorted = sortperm(h_scenario_max; rev=true)`.
`y_opt` is (n_y, H, K); `y_min` / `y_max` are (n_y, H) with +-Inf where there is no bound.
"""
function scenario_constraint_max(y_opt::Array{Float64,3}, y_min, y_max; device=0)
    n_y, H, K = size(y_opt)
    lo = Matrix{Float64}(y_min); hi = Matrix{Float64}(y_max)
    h_max = Vector{Float64}(undef, K); order = Vector{Int64}(undef, K)
    # column-major (n_y, H, K) is exactly the scenario-major K x H x n_y layout of the C ABI
    check(ccall((:pgb_scenario_constraint_max, LIB), Cint,
                (Int32, Int64, Int64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
                device, K, H, n_y, y_opt, lo, hi, h_max, order))
    return h_max, order
end

# ---------------------------------------------------------------------------------------------------------------------
# several GPUs (`devices = [0, 1, ...]`), device-resident samples, stacked dynamics — include/pgopt_b200.h: pgb_multi_*,
# pgb_samples_*, pgb_dynamics_dev.  Thin stubs over the C ABI, in the order a maintainer would wire them.
# ---------------------------------------------------------------------------------------------------------------------
"""
    particle_Gibbs_multi(u, y, K, K_b, k_d, N, phi, Lambda_Q, ell_Q, Q_init, V, A_init, x_init_dist, g, R; n_chains, devices)

`n_chains` independent chains of `particle_Gibbs` dealt to `devices` (one process, in-process NCCL).  The kept samples
stay on the GPU that produced them; returns the packed records `[A | Q | x_T | u_m1]` of all chains, a
`(record, K, n_chains)` array (x_T = x_m1[:, star] drawn on the owning GPU, optimal_control_Ipopt.jl:58-59), and the handle
for `multi_rollout`.
"""
function particle_Gibbs_multi(u_training, y_training, K, K_b, k_d, N, phi, Lambda_Q, ell_Q, Q_init, V, A_init, x_init_dist,
                              g::LinearMeasurement, R; n_chains, devices=[0], seed=rand(UInt64), roll_seed=rand(UInt64))
    n_y, T = size(y_training)
    cfg, n_x, n_u, n_phi = make_config(phi, n_y, N, T; n_chains=n_chains)
    m = Ref{Ptr{Cvoid}}(C_NULL)
    dv = Vector{Int32}(devices)
    mcheck(rc) = rc == 0 ? nothing : error("pgopt_b200 error $rc: " * unsafe_string(ccall((:pgb_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), m[])))
    mcheck(ccall((:pgb_multi_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ref{PgbConfig}, Ptr{Int32}, Int32), m, cfg, dv, length(dv)))
    mcheck(ccall((:pgb_multi_set_data, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), m[], Matrix{Float64}(u_training), Matrix{Float64}(y_training)))
    Rm = R isa Number ? fill(Float64(R), 1, 1) : Matrix{Float64}(R)
    mcheck(ccall((:pgb_multi_set_measurement, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), m[], g.C, Rm))
    mu = Vector{Float64}(mean(x_init_dist)); Lc = Matrix{Float64}(cholesky(Matrix(cov(x_init_dist))).L)
    mcheck(ccall((:pgb_multi_set_init_dist, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), m[], mu, Lc))
    dense, Vd = prior_V(V)
    if dense
        mcheck(ccall((:pgb_multi_set_prior_dense, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64), m[], Vd, Matrix{Float64}(Lambda_Q * I(n_x)), Float64(ell_Q)))
    else
        mcheck(ccall((:pgb_multi_set_prior, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64), m[], Vd, Matrix{Float64}(Lambda_Q * I(n_x)), Float64(ell_Q)))
    end
    mcheck(ccall((:pgb_multi_set_state, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64),
                 m[], -1, Matrix{Float64}(A_init), Matrix{Float64}(Q_init * I(n_x)), C_NULL, 1))
    mcheck(ccall((:pgb_multi_set_rng_philox, LIB), Cint, (Ptr{Cvoid}, UInt64, UInt32), m[], seed, 0))
    R_ = ccall((:pgb_packed_record_doubles, LIB), Int64, (Int32, Int32, Int32), n_x, n_phi, n_u)
    packed = Array{Float64}(undef, R_, K, n_chains)
    mcheck(ccall((:pgb_multi_particle_gibbs, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, UInt64, Ptr{Float64}), m[], K, K_b, k_d, roll_seed, packed))
    return packed, m[]          # release with ccall((:pgb_multi_destroy, LIB), Cvoid, (Ptr{Cvoid},), m)
end

"""
    multi_rollout(m, g, R, H, n_x, K_total; u_init, seed)   ->  (x_vec_0, X_init, Y_init)

Scenario rollout of every resident sample, each GPU its own, outputs gathered (NCCL) and returned in the reference's
layouts.  `seed` must be the `roll_seed` of `particle_Gibbs_multi`.
"""
function multi_rollout(m::Ptr{Cvoid}, g::LinearMeasurement, R, H, n_x, n_u, K_total; u_init=nothing, seed)
    n_y = size(g.C, 1)
    Le = R isa Number ? fill(Float64(R), 1, 1) : Matrix{Float64}(cholesky(Matrix(R)).L)
    u0 = u_init === nothing ? zeros(n_u, H) : Matrix{Float64}(u_init)
    x0 = Array{Float64}(undef, n_x, K_total); X = Array{Float64}(undef, n_x, H + 1, K_total); Y = Array{Float64}(undef, n_y, H, K_total)
    rc = ccall((:pgb_multi_rollout, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, UInt64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
               m, H, g.C, Le, u0, seed, x0, X, Y)
    rc == 0 || error("pgopt_b200 error $rc: " * unsafe_string(ccall((:pgb_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), m)))
    return vec(x0), reshape(permutedims(X, (1, 3, 2)), n_x * K_total, H + 1), reshape(permutedims(Y, (1, 3, 2)), n_y * K_total, H)
end

"""
    discrete_dynamics_jacobians(samples, x_vec, u, t; phi_variant=1)  ->  (x_vec_n, dfdx, dfdu)

RobotDynamics.discrete_dynamics of PGS_model_obj (Julia/src/optimal_control_Altro.jl:79-107) and the Jacobians @autodiff
(:51) derives from it, on a device-resident sample set (`samples::Ptr{Cvoid}` from pgb_samples_create / pgb_particle_gibbs_dev
after a pgb_rollout_dev that left v_vec on the device).  dfdx[:, :, k] is the k-th diagonal block, dfdu[:, :, k] rows
n_x*(k-1)+1 .. n_x*k of the (K n_x) x n_u Jacobian.  phi_variant 1 = the rounding of the examples' phi_opt.
"""
function discrete_dynamics_jacobians(samples::Ptr{Cvoid}, x_vec::Vector{Float64}, u::Vector{Float64}, t::Integer, n_x, n_u; phi_variant=1)
    K = div(length(x_vec), n_x)
    xn = Vector{Float64}(undef, K * n_x); jx = Array{Float64}(undef, n_x, n_x, K); ju = Array{Float64}(undef, n_x, n_u, K)
    rc = ccall((:pgb_dynamics_dev, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
               samples, x_vec, u, t, phi_variant, xn, jx, ju)
    rc == 0 || error("pgopt_b200 error $rc: " * unsafe_string(ccall((:pgb_samples_last_error, LIB), Cstring, (Ptr{Cvoid},), samples)))
    return xn, jx, ju
end

"""
    multi_dynamics(m, x_vec, u, t, n_x, n_u; phi_variant=1)  ->  (x_vec_n, dfdx, dfdu)

The same step on the samples a `pgb_multi` handle keeps sharded over its GPUs (after `multi_rollout`): every device
evaluates its own scenarios from / into its block of the host arrays (`pgb_multi_dynamics`); no collective.
"""
function multi_dynamics(m::Ptr{Cvoid}, x_vec::Vector{Float64}, u::Vector{Float64}, t::Integer, n_x, n_u; phi_variant=1)
    K = div(length(x_vec), n_x)
    xn = Vector{Float64}(undef, K * n_x); jx = Array{Float64}(undef, n_x, n_x, K); ju = Array{Float64}(undef, n_x, n_u, K)
    rc = ccall((:pgb_multi_dynamics, LIB), Cint,
               (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
               m, x_vec, u, t, phi_variant, xn, jx, ju, C_NULL)
    rc == 0 || error("pgopt_b200 error $rc: " * unsafe_string(ccall((:pgb_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), m)))
    return xn, jx, ju
end

end # module
