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
# Beyond the reference's surface (all through the same C ABI):
#   tabulated(f, t0, t1)                      any real closure as a drive envelope (qsim_problem_add_table)
#   select = [(row, col), ...], abs2 = true   only those entries (or their squared moduli) of every result
#                                             leave the device (qsim_solve_selected)
#   devices = [0, 1, ...]                     one call fans the sweep out over several GPUs, block-cyclically
#                                             (qsim_solve_multi; no Julia threads involved)
#   floquet_propagator(unitary_propagator, tau) / floquet_rise_fall_propagator(...)
#                                             src/time_evolution.jl:188-346 as ONE batched call (qsim_floquet_propagator)
#
# NOTE: Julia is not installed in the build/test image of this repository; this file is the binding a
# maintainer adds on the reference side (see INTEGRATION.md) and has been checked by reading only.
module QSimulatorB200

using QSimulator
using Libdl
using LinearAlgebra: I, diag
import QSimulator: CompositeQSystem, QSystem, DuffingTransmon, PerturbativeTransmon, DiagonalChargeBasisTransmon, RotatingFrameSystem,
                   dimension, spec, hamiltonian, raising, lowering, X, Y, embed

export Param, Signal, CosSignal, SinSignal, ConstSignal, gaussian, erf_squared
export dipole_drive_b200, rwa_dipole_b200, parametric_drive_b200, xy_exchange_drive_b200, ScaledLindblad
export unitary_propagator, unitary_state, me_propagator, me_state
export tabulated, floquet_propagator, floquet_rise_fall_propagator

const libqsim = get(ENV, "QSIM_B200_LIB", joinpath(@__DIR__, "..", "lib", "libqsim_b200.so"))
# ccall needs a constant (symbol, library) pair; entry points chosen at run time go through dlsym instead
const LIBHANDLE = Ref{Ptr{Cvoid}}(C_NULL)
function fptr(name::Symbol)
    LIBHANDLE[] == C_NULL && (LIBHANDLE[] = Libdl.dlopen(libqsim))
    return Libdl.dlsym(LIBHANDLE[], name)
end

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
const ENV_NONE, ENV_GAUSSIAN, ENV_ERF_SQUARED, ENV_TABLE = Int32(0), Int32(1), Int32(2), Int32(3)

"A real function of time as a piecewise Chebyshev series: n_seg equal segments over [t_start, t_end], n_coef
coefficients each (qsim_problem_add_table); `f` is kept so that the signal still calls the closure on the host."
struct Table
    t_start::Float64
    t_end::Float64
    coef::Matrix{Float64}      # n_coef x n_seg (column = one segment: row-major [n_seg][n_coef] for the C side)
    f::Function
end

"Real scalar drive s(t); callable like the closures the reference passes to its drive factories."
struct Signal <: Function
    carrier::Int32; envelope::Int32; blend::Int32
    offset::Scalar; amp::Scalar; freq::Scalar; phase::Scalar
    t1::Scalar; t2::Scalar; sigma::Scalar; rest::Scalar
    table::Union{Nothing,Table}
end
Signal(c, e, b, offset, amp, freq, phase, t1, t2, sigma, rest) = Signal(c, e, b, offset, amp, freq, phase, t1, t2, sigma, rest, nothing)
CosSignal(amp, freq; phase=0.0, offset=0.0) = Signal(CARRIER_COS, ENV_NONE, 0, offset, amp, freq, phase, 0.0, 0.0, 1.0, 0.0)
SinSignal(amp, freq; phase=0.0, offset=0.0) = Signal(CARRIER_SIN, ENV_NONE, 0, offset, amp, freq, phase, 0.0, 0.0, 1.0, 0.0)
ConstSignal(v) = Signal(CARRIER_NONE, ENV_NONE, 0, 0.0, v, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0)
gaussian(s::Signal, center, sigma) =
    Signal(s.carrier, ENV_GAUSSIAN, s.blend, s.offset, s.amp, s.freq, s.phase, center, s.t2, sigma, s.rest)
erf_squared(s::Signal, t1, t2, sigma) =
    Signal(s.carrier, ENV_ERF_SQUARED, 0, s.offset, s.amp, s.freq, s.phase, t1, t2, sigma, s.rest)
erf_squared(s::Signal, t1, t2, sigma, rest) =
    Signal(s.carrier, ENV_ERF_SQUARED, 1, s.offset, s.amp, s.freq, s.phase, t1, t2, sigma, rest)

"Clenshaw evaluation of a Table at time t (the same lookup the device does; constant continuation outside)."
function (tb::Table)(t::Real)
    n_coef, n_seg = size(tb.coef)
    u = (t - tb.t_start) * (n_seg / (tb.t_end - tb.t_start))
    k = clamp(floor(Int, u), 0, n_seg - 1)
    x = clamp(2.0 * (u - k) - 1.0, -1.0, 1.0)
    b1 = 0.0; b2 = 0.0
    for j in n_coef:-1:2
        b1, b2 = 2.0 * x * b1 + (tb.coef[j, k + 1] - b2), b1
    end
    return x * b1 + (tb.coef[1, k + 1] - b2)
end

"""
    tabulated(f, t_start, t_end; amp=1.0, offset=0.0, n_coef=12, tol=1e-14)

`t -> offset + amp*f(t)` for an arbitrary real closure `f` on [t_start, t_end] -- the generic path for the opaque
drive closures of src/composite_systems.jl:40-50.  `f` is sampled into a piecewise Chebyshev series whose segment
count doubles until it reproduces `f` at 2*n_coef off-node points per segment to `tol * max|f|`.
"""
function tabulated(f::Function, t_start::Real, t_end::Real; amp=1.0, offset=0.0, n_coef::Int=12, tol::Real=1e-14,
                   atol::Real=0.0, max_seg::Int=1 << 16)
    N = n_coef
    nodes = [cos(pi * (j + 0.5) / N) for j in 0:N-1]
    dct = [(k == 0 ? 1.0 : 2.0) / N * cos(pi * k * (j + 0.5) / N) for k in 0:N-1, j in 0:N-1]
    chk = [cos(pi * (j + 0.25) / (2N)) for j in 0:2N-1]
    n_seg = 1
    while true
        h = (t_end - t_start) / n_seg
        coef = Matrix{Float64}(undef, N, n_seg)
        for s in 1:n_seg
            left = t_start + h * (s - 1)
            coef[:, s] = dct * [Float64(f(left + 0.5 * h * (1.0 + x))) for x in nodes]
        end
        tb = Table(Float64(t_start), Float64(t_end), coef, f)
        err = 0.0; scale = 1e-300
        for s in 1:n_seg, x in chk
            t = t_start + h * (s - 1) + 0.5 * h * (1.0 + x)
            want = Float64(f(t))
            err = max(err, abs(tb(t) - want)); scale = max(scale, abs(want))
        end
        if err <= max(tol * scale, atol) || n_seg >= max_seg
            err <= max(1e-9 * scale, atol) || error("could not tabulate the function to 1e-9 with $n_seg segments")
            return Signal(CARRIER_NONE, ENV_TABLE, 0, offset, amp, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, tb)
        end
        n_seg *= 2
    end
end

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
    elseif s.envelope == ENV_TABLE
        env = Float64(s.table.f(t))     # the closure itself on the host, like the reference
    end
    s.blend == 0 && return val(s.offset, row) + val(s.amp, row) * env * car
    return env * (val(s.offset, row) + val(s.amp, row) * car) + (1 - env) * val(s.rest, row)
end
erf_(x) = ccall(:erf, Float64, (Float64,), x)   # libm, avoids a SpecialFunctions dependency

# table_ids: objectid(Table) -> id returned by qsim_problem_add_table for the problem being built
csignal(s::Signal, table_ids=nothing) =
    CSignal(s.carrier, s.envelope, s.blend, s.table === nothing ? Int32(0) : Int32(table_ids[objectid(s.table)]),
            Slot(s.offset), Slot(s.amp), Slot(s.freq), Slot(s.phase), Slot(s.t1), Slot(s.t2), Slot(s.sigma), Slot(s.rest))
as_signal(s::Signal) = s
as_signal(x::Scalar) = ConstSignal(x)
# Arbitrary closures reach the library as tabulated envelopes: wrap them with tabulated(f, t_start, t_end).
as_signal(::Function) = error("opaque closures cannot cross the C ABI as they are: pass tabulated(f, t_start, t_end) or a Signal (CosSignal/SinSignal/...)")

# One lowered qsim_term before embedding
struct Lowered
    kind::Int32
    signal::Signal
    atom_a::Union{Nothing,Matrix{ComplexF64}}
    atom_b::Union{Nothing,Matrix{ComplexF64}}
    rate::Scalar; anharm::Scalar; rot::Scalar; EC::Scalar; EJ1::Scalar; EJ2::Scalar
    num_terms::Int
    level_tables::Vector{Table}      # TERM_SPECTRUM: E_k(z), z = cos(pi*flux) on [-1, 1], one table per level
end
Lowered(kind, signal, a, b, rate, anharm, rot, EC, EJ1, EJ2, num_terms) =
    Lowered(kind, signal, a, b, rate, anharm, rot, EC, EJ1, EJ2, num_terms, Table[])
const TERM_SCALAR, TERM_ROTATING, TERM_DUFFING, TERM_TRANSMON, TERM_SPECTRUM = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4)

"""
Level energies of a DiagonalChargeBasisTransmon (src/systems.jl:143-152, 231-237) as tables in z = cos(pi*flux):
the spectrum depends on the flux only through EJ(flux)^2 = (EJ1 - EJ2)^2 + 4 EJ1 EJ2 z^2 and is a smooth even
function of z.  Built once per problem; the reference diagonalises the charge-basis Hamiltonian in every RHS call.
"""
function charge_spectrum_tables(base::DiagonalChargeBasisTransmon)
    sp = spec(base)
    cache = Dict{Float64,Vector{Float64}}()
    levels(z) = get!(cache, z) do
        real.(diag(hamiltonian(base, acos(clamp(z, -1.0, 1.0)) / pi)))
    end
    atol = 4e-15 * max(4 * abs(sp.EC) * (base.num_terms ÷ 2)^2, abs(sp.EJ1) + abs(sp.EJ2), 1.0)
    [tabulated(z -> levels(z)[k], -1.0, 1.0; n_coef=14, tol=1e-14, atol=atol, max_seg=1 << 10).table
     for k in 1:dimension(base)]
end

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
    elseif base isa DiagonalChargeBasisTransmon
        sp = spec(base)
        Lowered(TERM_SPECTRUM, s, nothing, nothing, 0.0, 0.0, rot, sp.EC, sp.EJ1, sp.EJ2, base.num_terms,
                charge_spectrum_tables(base))
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
    # tabulated envelopes: one qsim_problem_add_table per distinct Table, ids handed to the signals below
    table_ids = Dict{UInt,Int32}()
    register(sig::Signal) = register(sig.table)
    function register(tb::Union{Nothing,Table})
        (tb === nothing || haskey(table_ids, objectid(tb))) && return
        n_coef, n_seg = size(tb.coef)
        id = Ref{Int32}(0)
        check(ccall((:qsim_problem_add_table, libqsim), Cint,
                    (Ptr{Cvoid}, Cdouble, Cdouble, Cint, Cint, Ptr{Float64}, Ref{Int32}),
                    prob.handle, tb.t_start, tb.t_end, n_seg, n_coef, tb.coef, id))
        table_ids[objectid(tb)] = id[]
    end
    for (ham, _) in cqs.parametric_Hs
        ham isa ParametricHamiltonian && foreach(lo -> (register(lo.signal); foreach(register, lo.level_tables)), ham.lowered)
    end
    for (lf, _) in cqs.parametric_Ls
        lf isa ScaledLindblad && register(lf.scale)
    end
    for (ham, idxs) in cqs.parametric_Hs                             # src/composite_systems.jl:40-50
        ham isa ParametricHamiltonian || error("parametric Hamiltonian is an opaque closure; use the *_b200 drive factories")
        for lo in ham.lowered
            term = Ref(CTerm(lo.kind, Int32(lo.num_terms), csignal(lo.signal, table_ids), Slot(lo.rate), Slot(lo.anharm),
                             Slot(lo.rot), Slot(lo.EC), Slot(lo.EJ1), Slot(lo.EJ2)))
            if lo.kind == TERM_SCALAR || lo.kind == TERM_ROTATING
                a = embedded(cqs, lo.atom_a, idxs)
                b = lo.atom_b === nothing ? nothing : embedded(cqs, lo.atom_b, idxs)
                check(ccall((:qsim_problem_add_term, libqsim), Cint,
                            (Ptr{Cvoid}, Ref{CTerm}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Int32}),
                            prob.handle, term, a, b === nothing ? C_NULL : b, C_NULL))
            elseif lo.kind == TERM_SPECTRUM
                # flux drive on a DiagonalChargeBasisTransmon: the level tables registered above
                lv = levels_from_indices(cqs, idxs)
                ids = Int32[table_ids[objectid(tb)] for tb in lo.level_tables]
                rot = Ref(Slot(lo.rot))
                check(ccall((:qsim_problem_add_spectrum_term, libqsim), Cint,
                            (Ptr{Cvoid}, Ref{CSignal}, Ref{Slot}, Ptr{Int32}, Cint, Ptr{Int32}),
                            prob.handle, Ref(csignal(lo.signal, table_ids)), rot, lv, length(ids), ids))
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
        sc = Ref(csignal(lf.scale, table_ids))
        check(ccall((:qsim_problem_add_lindblad, libqsim), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ref{CSignal}),
                    prob.handle, l, sc))
    end
    check(ccall((:qsim_problem_finalize, libqsim), Cint, (Ptr{Cvoid},), prob.handle))
    return prob
end

# ---- contexts (one per device, created on first use) --------------------------------------------
const CTXS = Dict{Int,Ptr{Cvoid}}()
function context(device::Integer=parse(Int, get(ENV, "QSIM_B200_DEVICE", "0")))
    get!(CTXS, Int(device)) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:qsim_create, libqsim), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h))
        h[]
    end
end

const WHICH = Dict(:qsim_unitary_propagator => 0, :qsim_unitary_state => 1, :qsim_me_propagator => 2, :qsim_me_state => 3)

# ---- solver calls ------------------------------------------------------------------------------
# One batched call.  `select` (vector of (row, col) tuples, or of rows for unitary_state; 1-based): only those entries
# of every result are produced (qsim_solve_selected); with abs2 their squared moduli.  `devices`: fan the sweep out
# over several GPUs from this one call (qsim_solve_multi, block-cyclic shards, no Julia threads).
function solve(sym::Symbol, cqs::CompositeQSystem, ts::AbstractVector{<:Real}, params::AbstractMatrix{<:Real},
               init::Union{Nothing,Array{ComplexF64}}, shape::Tuple; select=nothing, abs2::Bool=false,
               devices::Union{Nothing,AbstractVector{<:Integer}}=nothing)
    @assert issorted(ts)                                             # src/time_evolution.jl:30,64,97,141
    n_traj, n_params = size(params)
    prob = Problem(cqs, n_params)
    tsv = collect(Float64, ts)
    ptab = Matrix{Float64}(transpose(Float64.(params)))              # row-major rows for the C side
    stats = Vector{TrajStats}(undef, n_traj)
    opts = Ref(default_opts())
    which = WHICH[sym]
    initp = init === nothing ? Ptr{ComplexF64}(C_NULL) : pointer(init)
    if select !== nothing
        devices === nothing || error("select= is not available together with devices=")
        rows = Int32[(s isa Tuple ? s[1] : s) - 1 for s in select]
        cols = Int32[(s isa Tuple ? s[2] : 1) - 1 for s in select]
        out = abs2 ? Array{Float64}(undef, length(rows), length(tsv), n_traj) :
                     Array{ComplexF64}(undef, length(rows), length(tsv), n_traj)
        GC.@preserve prob tsv ptab out stats init rows cols begin
            check(ccall((:qsim_solve_selected, libqsim), Cint,
                        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Cint, Int64, Ptr{ComplexF64}, Int64,
                         Ref{SolverOpts}, Cint, Ptr{Int32}, Ptr{Int32}, Cint, Ptr{Cvoid}, Ptr{TrajStats}),
                        context(), prob.handle, which, n_traj, ptab, tsv, length(tsv), 0, initp, 0, opts, length(rows), rows,
                        cols, abs2 ? 1 : 0, out, stats))
        end
    else
        out = Array{ComplexF64}(undef, shape..., length(tsv), n_traj)    # out[traj][k][col][row], column-major
        GC.@preserve prob tsv ptab out stats init begin
            rc = if devices !== nothing
                ctxs = Ptr{Cvoid}[context(dv) for dv in devices]
                ccall((:qsim_solve_multi, libqsim), Cint,
                      (Ptr{Ptr{Cvoid}}, Cint, Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Cint, Int64, Ptr{ComplexF64},
                       Int64, Ref{SolverOpts}, Ptr{ComplexF64}, Ptr{TrajStats}),
                      ctxs, length(ctxs), prob.handle, which, n_traj, ptab, tsv, length(tsv), 0, initp, 0, opts, out, stats)
            elseif init === nothing
                ccall(fptr(sym), Cint,
                      (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Cint, Int64, Ref{SolverOpts},
                       Ptr{ComplexF64}, Ptr{TrajStats}),
                      context(), prob.handle, n_traj, ptab, tsv, length(tsv), 0, opts, out, stats)
            else
                ccall(fptr(sym), Cint,
                      (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Cint, Int64, Ptr{ComplexF64}, Int64,
                       Ref{SolverOpts}, Ptr{ComplexF64}, Ptr{TrajStats}),
                      context(), prob.handle, n_traj, ptab, tsv, length(tsv), 0, init, 0, opts, out, stats)
            end
            check(rc)
        end
    end
    for (i, s) in enumerate(stats)
        s.retcode == 0 || @warn "trajectory $i: solver retcode $(s.retcode) at t = $(s.t_final)"   # OrdinaryDiffEq warns too
    end
    return out, stats
end

split_results(out::Array{<:Number,4}, traj) = [out[:, :, k, traj] for k in 1:size(out, 3)]
split_results(out::Array{<:Number,3}, traj) = [out[:, k, traj] for k in 1:size(out, 2)]
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

# batched methods: one trajectory per row of `params`; result[traj][k].  Keywords: select / abs2 (entries of every
# result, returned as result[traj][k] vectors), devices (GPUs to spread the sweep over).
function unitary_propagator(cqs::CompositeQSystem, ts::AbstractVector{<:Real}, params::AbstractMatrix{<:Real}; kw...)
    d = dimension(cqs)
    out, _ = solve(:qsim_unitary_propagator, cqs, ts, params, nothing, (d, d); kw...)
    return [split_results(out, i) for i in 1:size(params, 1)]
end
function me_propagator(cqs::CompositeQSystem, ts::AbstractVector{<:Real}, params::AbstractMatrix{<:Real}; kw...)
    d2 = dimension(cqs)^2
    out, _ = solve(:qsim_me_propagator, cqs, ts, params, nothing, (d2, d2); kw...)
    return [split_results(out, i) for i in 1:size(params, 1)]
end
function unitary_state(cqs::CompositeQSystem, ts::AbstractVector{<:Real}, ψ0::Vector{<:Number}, params::AbstractMatrix{<:Real}; kw...)
    out, _ = solve(:qsim_unitary_state, cqs, ts, params, Vector{ComplexF64}(ψ0), (dimension(cqs),); kw...)
    return [split_results(out, i) for i in 1:size(params, 1)]
end
function me_state(cqs::CompositeQSystem, ts::AbstractVector{<:Real}, ρ0::Matrix{<:Number}, params::AbstractMatrix{<:Real}; kw...)
    d = dimension(cqs)
    out, _ = solve(:qsim_me_state, cqs, ts, params, Matrix{ComplexF64}(ρ0), (d, d); kw...)
    return [split_results(out, i) for i in 1:size(params, 1)]
end

# ---- Floquet wrappers (src/time_evolution.jl:188-346) as one batched library call ---------------------------
function floquet_call(which::Integer, cqs::CompositeQSystem, ts::AbstractVector{<:Real}, params::AbstractMatrix{<:Real},
                      periods::Vector{Float64}, rise::Real, fall::Real, rtol::Real)
    @assert issorted(ts)
    n_traj, n_params = size(params)
    prob = Problem(cqs, n_params)
    n = which == 0 ? dimension(cqs) : dimension(cqs)^2
    tsv = collect(Float64, ts)
    ptab = Matrix{Float64}(transpose(Float64.(params)))
    out = Array{ComplexF64}(undef, n, n, length(tsv), n_traj)
    stats = Vector{TrajStats}(undef, n_traj)
    opts = Ref(default_opts())
    GC.@preserve prob tsv ptab out stats periods begin
        check(ccall((:qsim_floquet_propagator, libqsim), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Cint, Int64, Ptr{Float64}, Int64,
                     Cdouble, Cdouble, Cdouble, Ref{SolverOpts}, Ptr{ComplexF64}, Ptr{TrajStats}),
                    context(), prob.handle, which, n_traj, ptab, tsv, length(tsv), 0, periods, length(periods) == 1 ? 0 : 1,
                    rise, fall, rtol, opts, out, stats))
    end
    return out
end
floquet_which(f) = f === unitary_propagator ? 0 : f === me_propagator ? 2 :
    error("the Floquet wrappers take this module's unitary_propagator or me_propagator")

"""
    floquet_propagator(propagator_func, t_period[, rtol])

As src/time_evolution.jl:188-228: a propagator function for systems periodic with `t_period`.  The returned function
also has a batched method `p(cqs, ts, params)`; `t_period` may then be a vector with one period per sweep point.
"""
floquet_propagator(f::Function, t_period, rtol::Real=0.0) = floquet_rise_fall_propagator(f, t_period, 0.0, 0.0, rtol)

"As src/time_evolution.jl:312-346; rtol = 0 selects the reference's default eps(t_period)."
function floquet_rise_fall_propagator(f::Function, t_period, rise_time::Real, fall_time::Real, rtol::Real=0.0)
    which = floquet_which(f)
    periods = Float64[t_period...]
    p(cqs::CompositeQSystem, ts::AbstractVector{<:Real}) =
        split_results(floquet_call(which, cqs, ts, no_params(cqs), periods, rise_time, fall_time, rtol), 1)
    function p(cqs::CompositeQSystem, ts::AbstractVector{<:Real}, params::AbstractMatrix{<:Real})
        out = floquet_call(which, cqs, ts, params, periods, rise_time, fall_time, rtol)
        return [split_results(out, i) for i in 1:size(params, 1)]
    end
    return p
end

end # module
