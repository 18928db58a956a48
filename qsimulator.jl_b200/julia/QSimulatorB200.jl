# QSimulatorB200.jl -- Julia shim over libqsim_b200.so (include/qsim_b200.h).
#
# Drop-in for QSimulator.jl's time-evolution path: the same four functions, same argument order,
#   unitary_propagator(cqs, ts) / (cqs, t)        src/time_evolution.jl:29-47
#   unitary_state(cqs, ts, psi0) / (cqs, t, psi0) src/time_evolution.jl:63-79
#   me_propagator(cqs, ts) / (cqs, t)             src/time_evolution.jl:96-123
#   me_state(cqs, ts, rho0) / (cqs, t, rho0)      src/time_evolution.jl:140-164
# plus batched methods taking a parameter table (one trajectory per row) -- the reference's sweeps
# are user-level loops (doc/Examples.ipynb cells 11, 14, 19, 27, 35).
#
# Julia closures cannot cross the C ABI, so the drive factories below return callable structs
# `<: Function` that carry an analytic descriptor: `add_hamiltonian!(cqs, ham::Function, acting_on)`
# (src/composite_systems.jl:40) still dispatches on them and `ham(t)` still returns the matrix the
# reference closure would, so the stock CPU solver keeps working on the same CompositeQSystem.
#
# NOTE: Julia is not installed in the build/test image of this repository; this file is the binding a
# maintainer adds on the reference side (see INTEGRATION.md) and has been checked by reading only.
module QSimulatorB200

using QSimulator
using LinearAlgebra: I
import QSimulator: CompositeQSystem, QSystem, DuffingTransmon, PerturbativeTransmon, RotatingFrameSystem,
                   dimension, spec, hamiltonian, raising, lowering, X, Y, embed

export Param, Signal, CosSignal, SinSignal, ConstSignal, gaussian, erf_squared
export dipole_drive_b200, rwa_dipole_b200, parametric_drive_b200, xy_exchange_drive_b200, ScaledLindblad
export unitary_propagator, unitary_state, me_propagator, me_state

const libqsim = get(ENV, "QSIM_B200_LIB", joinpath(@__DIR__, "..", "lib", "libqsim_b200.so"))

# ---- mirrors of the C structs (include/qsim_b200.h) ---------------------------------------
struct Slot
    param::Int32
    _pad::Int32
    value::Float64
end
struct Param
    index::Int      # 1-based column of the sweep's parameter table
end
Slot(x::Real) = Slot(Int32(-1), Int32(0), Float64(x))
Slot(p::Param) = Slot(Int32(p.index - 1), Int32(0), 0.0)
const Scalar = Union{Real,Param}

struct CSignal
    carrier::Int32; envelope::Int32; blend::Int32; table::Int32   # table: id from qsim_problem_add_table (0: none)
    offset::Slot; amp::Slot; freq::Slot; phase::Slot; t1::Slot; t2::Slot; sigma::Slot; rest::Slot
end
struct CTerm
    kind::Int32; num_terms::Int32
    signal::CSignal
    rate::Slot; anharm::Slot; rot::Slot; EC::Slot; EJ1::Slot; EJ2::Slot
end
struct SolverOpts
    reltol::Float64; abstol::Float64; gamma::Float64; qmin::Float64; qmax::Float64
    beta1::Float64; beta2::Float64; qoldinit::Float64; dtmax::Float64
    maxiters::Int64; flags::Int32; _pad::Int32; stream::Ptr{Cvoid}
end
struct TrajStats
    retcode::Int32; n_accept::Int32; n_reject::Int32; n_rhs::Int32
    dt_init::Float64; t_final::Float64
end
# the literals of src/time_evolution.jl:43 plus OrdinaryDiffEq's Tsit5 defaults
default_opts() = SolverOpts(1e-6, 1e-8, 0.9, 0.2, 10.0, 7 / 50, 2 / 25, 1e-4, 0.0, 1_000_000, 0, 0, C_NULL)

function check(rc::Integer)
    rc == 0 && return
    msg = unsafe_string(ccall((:qsim_last_error, libqsim), Cstring, ()))
    error("qsim_b200 error $rc: $msg")
end

# ---- drive descriptors ---------------------------------------------------------------------
const CARRIER_NONE, CARRIER_COS, CARRIER_SIN = Int32(0), Int32(1), Int32(2)
const ENV_NONE, ENV_GAUSSIAN, ENV_ERF_SQUARED = Int32(0), Int32(1), Int32(2)

"Real scalar drive s(t); callable like the closures the reference passes to its drive factories."
struct Signal <: Function
    carrier::Int32; envelope::Int32; blend::Int32
    offset::Scalar; amp::Scalar; freq::Scalar; phase::Scalar
    t1::Scalar; t2::Scalar; sigma::Scalar; rest::Scalar
end
CosSignal(amp, freq; phase=0.0, offset=0.0) = Signal(CARRIER_COS, ENV_NONE, 0, offset, amp, freq, phase, 0.0, 0.0, 1.0, 0.0)
SinSignal(amp, freq; phase=0.0, offset=0.0) = Signal(CARRIER_SIN, ENV_NONE, 0, offset, amp, freq, phase, 0.0, 0.0, 1.0, 0.0)
ConstSignal(v) = Signal(CARRIER_NONE, ENV_NONE, 0, 0.0, v, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0)
gaussian(s::Signal, center, sigma) =
    Signal(s.carrier, ENV_GAUSSIAN, s.blend, s.offset, s.amp, s.freq, s.phase, center, s.t2, sigma, s.rest)
erf_squared(s::Signal, t1, t2, sigma) =
    Signal(s.carrier, ENV_ERF_SQUARED, 0, s.offset, s.amp, s.freq, s.phase, t1, t2, sigma, s.rest)
erf_squared(s::Signal, t1, t2, sigma, rest) =
    Signal(s.carrier, ENV_ERF_SQUARED, 1, s.offset, s.amp, s.freq, s.phase, t1, t2, sigma, rest)

val(x::Real, row) = Float64(x)
val(p::Param, row) = row === nothing ? error("Param($(p.index)) needs a parameter row") : Float64(row[p.index])

function (s::Signal)(t::Real, row=nothing)
    car = s.carrier == CARRIER_NONE ? 1.0 :
          (arg = (2π * val(s.freq, row)) * t + val(s.phase, row); s.carrier == CARRIER_COS ? cos(arg) : sin(arg))
    env = 1.0
    if s.envelope == ENV_GAUSSIAN
        x = (t - val(s.t1, row)) / val(s.sigma, row)
        env = exp(-0.5 * (x * x))
    elseif s.envelope == ENV_ERF_SQUARED
        sg = val(s.sigma, row)
        env = 0.5 * (erf_((t - val(s.t1, row)) / sg) - erf_((t - val(s.t2, row)) / sg))
    end
    s.blend == 0 && return val(s.offset, row) + val(s.amp, row) * env * car
    return env * (val(s.offset, row) + val(s.amp, row) * car) + (1 - env) * val(s.rest, row)
end
erf_(x) = ccall(:erf, Float64, (Float64,), x)   # libm, avoids a SpecialFunctions dependency

csignal(s::Signal) = CSignal(s.carrier, s.envelope, s.blend, 0, Slot(s.offset), Slot(s.amp), Slot(s.freq),
                             Slot(s.phase), Slot(s.t1), Slot(s.t2), Slot(s.sigma), Slot(s.rest))
as_signal(s::Signal) = s
as_signal(x::Scalar) = ConstSignal(x)
# Arbitrary closures reach the library as tabulated envelopes (qsim_problem_add_table + ENV_TABLE = 3; the Python
# mirror's `tabulated(f, t0, t1)` shows the sampling).  This shim does not wire that path yet.
as_signal(::Function) = error("opaque closures cannot cross the C ABI: pass a Signal (CosSignal/SinSignal/...)")

# One lowered qsim_term before embedding
struct Lowered
    kind::Int32
    signal::Signal
    atom_a::Union{Nothing,Matrix{ComplexF64}}
    atom_b::Union{Nothing,Matrix{ComplexF64}}
    rate::Scalar; anharm::Scalar; rot::Scalar; EC::Scalar; EJ1::Scalar; EJ2::Scalar
    num_terms::Int
end
const TERM_SCALAR, TERM_ROTATING, TERM_DUFFING, TERM_TRANSMON = Int32(0), Int32(1), Int32(2), Int32(3)

"Time-dependent subsystem Hamiltonian: `ham(t)` like the reference's closures, plus descriptors."
struct ParametricHamiltonian <: Function
    f::Function
    lowered::Vector{Lowered}
end
(h::ParametricHamiltonian)(t::Real) = h.f(t)

"dipole_drive (src/operators.jl:220-226) with a descriptor drive."
function dipole_drive_b200(q::QSystem, drive, rotation_rate::Real=0.0)
    s = as_signal(drive)
    f = t -> s(t) * X(q, rotation_rate * t)
    lo = rotation_rate == 0 ?
        Lowered(TERM_SCALAR, s, ComplexF64.(raising(q) + lowering(q)), nothing, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0) :
        Lowered(TERM_ROTATING, s, ComplexF64.(raising(q)), ComplexF64.(lowering(q)), rotation_rate, 0.0, 0.0, 0.0, 0.0, 0.0, 0)
    ParametricHamiltonian(f, [lo])
end

"rwa_dipole (src/operators.jl:242-250); `drive_im` is the quadrature coupling to Y."
function rwa_dipole_b200(q::QSystem, drive_re, drive_im=nothing)
    re = as_signal(drive_re)
    x_ham, y_ham = ComplexF64.(X(q)), ComplexF64.(Y(q))
    los = [Lowered(TERM_SCALAR, re, x_ham, nothing, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0)]
    if drive_im === nothing
        return ParametricHamiltonian(t -> re(t) * x_ham, los)
    end
    im_ = as_signal(drive_im)
    push!(los, Lowered(TERM_SCALAR, im_, y_ham, nothing, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0))
    ParametricHamiltonian(t -> re(t) * x_ham + im_(t) * y_ham, los)
end

"parametric_drive (src/operators.jl:265-268) on a DuffingTransmon / PerturbativeTransmon, optionally in a RotatingFrameSystem."
function parametric_drive_b200(q::QSystem, drive)
    s = as_signal(drive)
    base, rot = q isa RotatingFrameSystem ? (q.system, q.rotation_rate) : (q, 0.0)
    lo = if base isa DuffingTransmon
        Lowered(TERM_DUFFING, s, nothing, nothing, 0.0, spec(base).anharmonicity, rot, 0.0, 0.0, 0.0, 0)
    elseif base isa PerturbativeTransmon
        sp = spec(base)
        Lowered(TERM_TRANSMON, s, nothing, nothing, 0.0, 0.0, rot, sp.EC, sp.EJ1, sp.EJ2, base.num_terms)
    else
        error("parametric_drive_b200: no kernel form for $(typeof(base))")
    end
    ParametricHamiltonian(t -> hamiltonian(q, s(t)), [lo])
end

"t -> coupling * X_Y([qa, qb], [ra*t, rb*t]) (benchmark/benchmarks.jl:184,193) = 2c(e^{2πi(ra-rb)t} a'b + h.c.)."
function xy_exchange_drive_b200(qa::QSystem, qb::QSystem, coupling::Scalar, ra::Real, rb::Real)
    up = ComplexF64.(2.0 * kron(raising(qa), lowering(qb)))
    dn = ComplexF64.(2.0 * kron(lowering(qa), raising(qb)))
    f = t -> val(coupling, nothing) * QSimulator.X_Y([qa, qb], [ra * t, rb * t])
    ParametricHamiltonian(f, [Lowered(TERM_ROTATING, ConstSignal(coupling), up, dn, ra - rb, 0.0, 0.0, 0.0, 0.0, 0.0, 0)])
end

"L(t) = scale(t) * op for add_lindblad!(cqs, ::Function, acting_on) (src/composite_systems.jl:59-62)."
struct ScaledLindblad <: Function
    op::Matrix{ComplexF64}
    scale::Signal
end
ScaledLindblad(op::AbstractMatrix, scale) = ScaledLindblad(ComplexF64.(op), as_signal(scale))
(l::ScaledLindblad)(t::Real) = l.scale(t) * l.op

# ---- lowering a CompositeQSystem to a qsim_problem ------------------------------------------
# expansion indices -> embedded d x d matrix, exactly embed_add! on zeros (src/composite_systems.jl:65-72)
function embedded(cqs::CompositeQSystem, op::AbstractMatrix, idxs)
    d = dimension(cqs)
    m = zeros(ComplexF64, d, d)
    QSimulator.embed_add!(m, op, idxs)
    return m
end

# level of the acted-on subsystem for every composite basis state, from the expansion indices of a
# diagonal single-subsystem operator: the (n,n) entry of the small operator lands on those states
function levels_from_indices(cqs::CompositeQSystem, idxs)
    d = dimension(cqs)
    da = isqrt(length(idxs))
    lv = zeros(Int32, d)
    for n in 0:da-1
        for lin in idxs[n * da + n + 1]
            lv[(lin - 1) % d + 1] = n
        end
    end
    return lv
end

mutable struct Problem
    handle::Ptr{Cvoid}
    d::Int
    n_params::Int
end

function Problem(cqs::CompositeQSystem, n_params::Integer)
    d = dimension(cqs)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:qsim_problem_create, libqsim), Cint, (Cint, Cint, Ref{Ptr{Cvoid}}), d, n_params, h))
    prob = Problem(h[], d, n_params)
    finalizer(p -> ccall((:qsim_problem_destroy, libqsim), Cint, (Ptr{Cvoid},), p.handle), prob)
    h0 = Matrix{ComplexF64}(hamiltonian(cqs))                       # src/composite_systems.jl:151-157
    check(ccall((:qsim_problem_set_h0, libqsim), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), prob.handle, h0))
    for (ham, idxs) in cqs.parametric_Hs                             # src/composite_systems.jl:40-50
        ham isa ParametricHamiltonian || error("parametric Hamiltonian is an opaque closure; use the *_b200 drive factories")
        for lo in ham.lowered
            term = Ref(CTerm(lo.kind, Int32(lo.num_terms), csignal(lo.signal), Slot(lo.rate), Slot(lo.anharm),
                             Slot(lo.rot), Slot(lo.EC), Slot(lo.EJ1), Slot(lo.EJ2)))
            if lo.kind == TERM_SCALAR || lo.kind == TERM_ROTATING
                a = embedded(cqs, lo.atom_a, idxs)
                b = lo.atom_b === nothing ? nothing : embedded(cqs, lo.atom_b, idxs)
                check(ccall((:qsim_problem_add_term, libqsim), Cint,
                            (Ptr{Cvoid}, Ref{CTerm}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Int32}),
                            prob.handle, term, a, b === nothing ? C_NULL : b, C_NULL))
            else
                lv = levels_from_indices(cqs, idxs)
                check(ccall((:qsim_problem_add_term, libqsim), Cint,
                            (Ptr{Cvoid}, Ref{CTerm}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Int32}),
                            prob.handle, term, C_NULL, C_NULL, lv))
            end
        end
    end
    for (op, idxs) in cqs.fixed_Ls                                   # src/composite_systems.jl:53-56
        l = embedded(cqs, op, idxs)
        check(ccall((:qsim_problem_add_lindblad, libqsim), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Cvoid}),
                    prob.handle, l, C_NULL))
    end
    for (lf, idxs) in cqs.parametric_Ls                              # src/composite_systems.jl:59-62
        lf isa ScaledLindblad || error("parametric Lindblad operator is an opaque closure; use ScaledLindblad")
        l = embedded(cqs, lf.op, idxs)
        sc = Ref(csignal(lf.scale))
        check(ccall((:qsim_problem_add_lindblad, libqsim), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ref{CSignal}),
                    prob.handle, l, sc))
    end
    check(ccall((:qsim_problem_finalize, libqsim), Cint, (Ptr{Cvoid},), prob.handle))
    return prob
end

# ---- context ---------------------------------------------------------------------------------
const CTX = Ref{Ptr{Cvoid}}(C_NULL)
function context()
    if CTX[] == C_NULL
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:qsim_create, libqsim), Cint, (Cint, Ref{Ptr{Cvoid}}), parse(Int, get(ENV, "QSIM_B200_DEVICE", "0")), h))
        CTX[] = h[]
    end
    return CTX[]
end

# ---- solver calls ------------------------------------------------------------------------------
function solve(sym::Symbol, cqs::CompositeQSystem, ts::AbstractVector{<:Real}, params::AbstractMatrix{<:Real},
               init::Union{Nothing,Array{ComplexF64}}, shape::Tuple)
    @assert issorted(ts)                                             # src/time_evolution.jl:30,64,97,141
    n_traj, n_params = size(params)
    prob = Problem(cqs, n_params)
    tsv = collect(Float64, ts)
    ptab = Matrix{Float64}(transpose(Float64.(params)))              # row-major rows for the C side
    n = prod(shape)
    out = Array{ComplexF64}(undef, shape..., length(tsv), n_traj)    # out[traj][k][col][row], column-major
    stats = Vector{TrajStats}(undef, n_traj)
    opts = Ref(default_opts())
    GC.@preserve prob tsv ptab out stats init begin
        rc = if init === nothing
            ccall((sym, libqsim), Cint,
                  (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Cint, Int64, Ref{SolverOpts},
                   Ptr{ComplexF64}, Ptr{TrajStats}),
                  context(), prob.handle, n_traj, ptab, tsv, length(tsv), 0, opts, out, stats)
        else
            ccall((sym, libqsim), Cint,
                  (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Cint, Int64, Ptr{ComplexF64}, Int64,
                   Ref{SolverOpts}, Ptr{ComplexF64}, Ptr{TrajStats}),
                  context(), prob.handle, n_traj, ptab, tsv, length(tsv), 0, init, 0, opts, out, stats)
        end
        check(rc)
    end
    for (i, s) in enumerate(stats)
        s.retcode == 0 || @warn "trajectory $i: solver retcode $(s.retcode) at t = $(s.t_final)"   # OrdinaryDiffEq warns too
    end
    return out, stats
end

split_results(out::Array{ComplexF64,4}, traj) = [out[:, :, k, traj] for k in 1:size(out, 3)]
split_results(out::Array{ComplexF64,3}, traj) = [out[:, k, traj] for k in 1:size(out, 2)]
no_params(cqs) = zeros(Float64, 1, 0)

# single-system methods: identical signatures and return types to the reference
function unitary_propagator(cqs::CompositeQSystem, ts::AbstractVector{<:Real})
    d = dimension(cqs)
    out, _ = solve(:qsim_unitary_propagator, cqs, ts, no_params(cqs), nothing, (d, d))
    return split_results(out, 1)
end
unitary_propagator(cqs::CompositeQSystem, t::Real) = unitary_propagator(cqs, [0.0, t])[end]

function unitary_state(cqs::CompositeQSystem, ts::AbstractVector{<:Real}, ψ0::Vector{<:Number})
    out, _ = solve(:qsim_unitary_state, cqs, ts, no_params(cqs), Vector{ComplexF64}(ψ0), (dimension(cqs),))
    return split_results(out, 1)
end
unitary_state(cqs::CompositeQSystem, t::Real, ψ0::Vector{<:Number}) = unitary_state(cqs, [0.0, t], ψ0)[end]

function me_propagator(cqs::CompositeQSystem, ts::AbstractVector{<:Real})
    d2 = dimension(cqs)^2
    out, _ = solve(:qsim_me_propagator, cqs, ts, no_params(cqs), nothing, (d2, d2))
    return split_results(out, 1)
end
me_propagator(cqs::CompositeQSystem, t::Real) = me_propagator(cqs, [0.0, t])[end]

function me_state(cqs::CompositeQSystem, ts::AbstractVector{<:Real}, ρ0::Matrix{<:Number})
    d = dimension(cqs)
    out, _ = solve(:qsim_me_state, cqs, ts, no_params(cqs), Matrix{ComplexF64}(ρ0), (d, d))
    return split_results(out, 1)
end
me_state(cqs::CompositeQSystem, t::Real, ρ0::Matrix{<:Number}) = me_state(cqs, [0.0, t], ρ0)[end]

# batched methods: one trajectory per row of `params`; result[traj][k]
function unitary_propagator(cqs::CompositeQSystem, ts::AbstractVector{<:Real}, params::AbstractMatrix{<:Real})
    d = dimension(cqs)
    out, _ = solve(:qsim_unitary_propagator, cqs, ts, params, nothing, (d, d))
    return [split_results(out, i) for i in 1:size(params, 1)]
end
function me_propagator(cqs::CompositeQSystem, ts::AbstractVector{<:Real}, params::AbstractMatrix{<:Real})
    d2 = dimension(cqs)^2
    out, _ = solve(:qsim_me_propagator, cqs, ts, params, nothing, (d2, d2))
    return [split_results(out, i) for i in 1:size(params, 1)]
end
function unitary_state(cqs::CompositeQSystem, ts::AbstractVector{<:Real}, ψ0::Vector{<:Number}, params::AbstractMatrix{<:Real})
    out, _ = solve(:qsim_unitary_state, cqs, ts, params, Vector{ComplexF64}(ψ0), (dimension(cqs),))
    return [split_results(out, i) for i in 1:size(params, 1)]
end
function me_state(cqs::CompositeQSystem, ts::AbstractVector{<:Real}, ρ0::Matrix{<:Number}, params::AbstractMatrix{<:Real})
    d = dimension(cqs)
    out, _ = solve(:qsim_me_state, cqs, ts, params, Matrix{ComplexF64}(ρ0), (d, d))
    return [split_results(out, i) for i in 1:size(params, 1)]
end

end # module
