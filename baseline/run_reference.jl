# Golden vectors from the REAL reference (QSimulator.jl + OrdinaryDiffEq), for anyone with a Julia install.
#
# The build image has no Julia, so the repo's oracle is "parity unpinned" at the 1e-8 level (DESIGN.md section 5).
# This script closes that gap: it runs the reference's own public API on small versions of the benchmark
# configs (same systems as qsimulator.jl_b200/host/qsim_b200/workloads.py) and writes the results as raw
# little-endian Float64 files that `python baseline/compare_with_reference.py <dir>` checks the oracle
# (and, on a GPU box, the CUDA library) against.
#
#   julia --project=/path/to/QSimulator.jl baseline/run_reference.jl out_dir
#
# Files: <name>.params (n_traj x n_params), <name>.ts, <name>.out ([traj][k][col][row] complex, i.e. each
# saved matrix column-major), <name>.meta (text: n_traj n_params n_ts n_rows n_cols).
using QSimulator
using LinearAlgebra

outdir = length(ARGS) >= 1 ? ARGS[1] : "reference_vectors"
mkpath(outdir)

function write_case(name, params::Matrix{Float64}, ts::Vector{Float64}, results)
    # results[i][k] is a Matrix{ComplexF64} (or Vector) for trajectory i, save k
    n_traj = size(params, 1)
    first = results[1][1]
    n_rows = size(first, 1)
    n_cols = ndims(first) == 2 ? size(first, 2) : 1
    open(joinpath(outdir, name * ".params"), "w") do io
        write(io, permutedims(params))          # row-major [traj][param]
    end
    open(joinpath(outdir, name * ".ts"), "w") do io
        write(io, ts)
    end
    open(joinpath(outdir, name * ".out"), "w") do io
        for i in 1:n_traj, k in 1:length(ts)
            write(io, ComplexF64.(results[i][k])) # column-major, interleaved re/im
        end
    end
    open(joinpath(outdir, name * ".meta"), "w") do io
        println(io, n_traj, " ", size(params, 2), " ", length(ts), " ", n_rows, " ", n_cols)
    end
end

# benchmark/benchmarks.jl:125-146 with the drive amp*sin(2 pi f t)
function lab_frame_gate(n, amp, freq)
    q0 = DuffingTransmon("q0", 3, DuffingSpec(3.94015, -0.1807))
    q1 = PerturbativeTransmon("q1", 3, TransmonSpec(0.172, 16.4, 0.55))
    spectators = [DuffingTransmon("q$ct", 3, DuffingSpec(4.0 + 0.1 * ct, -(0.2 + 0.01 * ct))) for ct = 2:(n-1)]
    all_qs = [q0, q1, spectators...]
    cqs = CompositeQSystem(all_qs)
    add_hamiltonian!(cqs, q0)
    add_hamiltonian!(cqs, parametric_drive(q1, t -> amp * sin(2π * freq * t)), q1)
    add_hamiltonian!(cqs, 0.006 * X([q0, q1]), [q0, q1])
    for q in all_qs[3:end]
        add_hamiltonian!(cqs, q)
        add_hamiltonian!(cqs, 0.006 * X([q, q1]), [q, q1])
    end
    return cqs, all_qs
end

# cfg1: the benchmark itself (amp = 1.0: benchmarks.jl:138 never uses `amp`)
let
    params = [1.0 0.1221]
    ts = collect(0.0:200.0)
    cqs, _ = lab_frame_gate(2, 1.0, 0.1221)
    write_case("cfg1", params, ts, [unitary_propagator(cqs, ts)])
end

# cfg2 corners + centre of the amplitude x frequency grid, short span
let
    params = [0.20 0.110; 0.20 0.134; 0.45 0.110; 0.45 0.134; 0.323 0.1221]
    ts = collect(0.0:1.0:20.0)
    res = [unitary_propagator(lab_frame_gate(2, params[i, 1], params[i, 2])[1], ts) for i in 1:size(params, 1)]
    write_case("cfg2_short", params, ts, res)
end

# cfg3: three transmons
let
    params = [0.323 0.1221; 0.25 0.115]
    ts = collect(0.0:1.0:10.0)
    res = [unitary_propagator(lab_frame_gate(3, params[i, 1], params[i, 2])[1], ts) for i in 1:size(params, 1)]
    write_case("cfg3_short", params, ts, res)
end

# cfg4: decay + dephasing on both transmons, me_propagator; parameters are (amp, f, sqrt(g1), sqrt(2 gphi))
let
    T1, Tphi = 20e3, 30e3
    g1, gphi = 1 / (2π * T1), 1 / (2π * Tphi)
    params = [0.323 0.1221 sqrt(g1) sqrt(2 * gphi)]
    ts = [0.0, 1.0, 2.0]
    cqs, qs = lab_frame_gate(2, 0.323, 0.1221)
    for q in qs
        add_lindblad!(cqs, decay(q, g1), [q])
        add_lindblad!(cqs, dephasing(q, gphi), [q])
    end
    write_case("cfg4_short", params, ts, [me_propagator(cqs, ts)])
end
# benchmark/benchmarks.jl:162-196 (rotating frame at each qubit's own frequency) with the drive amp*sin(2 pi f t)
function rotating_frame_gate(n, amp, freq)
    q0 = DuffingTransmon("q0", 3, DuffingSpec(3.94015, -0.1807))
    q1 = PerturbativeTransmon("q1", 3, TransmonSpec(0.172, 16.4, 0.55))
    spectators = [DuffingTransmon("q$ct", 3, DuffingSpec(4.0 + 0.1 * ct, -(0.2 + 0.01 * ct))) for ct = 2:(n-1)]
    all_qs = [q0, q1, spectators...]
    cqs = CompositeQSystem(all_qs)
    add_hamiltonian!(cqs, q0)
    add_hamiltonian!(cqs, -spec(q0).frequency * number(q0), q0)
    q1_freq = hamiltonian(q1, 0.0)[2, 2]
    add_hamiltonian!(cqs, -q1_freq * number(q1), q1)
    d01 = spec(q0).frequency - q1_freq
    add_hamiltonian!(cqs, t -> 0.006 * .5 * X_Y([q0, q1], [d01 * t, 0.0]), [q0, q1])
    add_hamiltonian!(cqs, parametric_drive(q1, t -> amp * sin(2π * freq * t)), q1)
    for q in all_qs[3:end]
        add_hamiltonian!(cqs, q)
        add_hamiltonian!(cqs, -spec(q).frequency * number(q), q)
        dq = q1_freq - spec(q).frequency
        add_hamiltonian!(cqs, t -> 0.006 * .5 * X_Y([q1, q], [dq * t, 0.0]), [q1, q])
    end
    return cqs, all_qs
end

# cfg5: three transmons in the rotating frame, decay + dephasing on all three, me_propagator (729 x 729), short span;
# parameters are (amp, f, sqrt(g1), sqrt(2 gphi)).  One RHS call of the reference is seven 729^3 GEMMs: minutes.
let
    T1, Tphi = 20e3, 30e3
    g1, gphi = 1 / (2π * T1), 1 / (2π * Tphi)
    params = [0.323 0.1221 sqrt(g1) sqrt(2 * gphi)]
    ts = [0.0, 0.1, 0.2]
    cqs, qs = rotating_frame_gate(3, 0.323, 0.1221)
    for q in qs
        add_lindblad!(cqs, decay(q, g1), [q])
        add_lindblad!(cqs, dephasing(q, gphi), [q])
    end
    write_case("cfg5_short", params, ts, [me_propagator(cqs, ts)])
end
println("wrote reference vectors to ", outdir)
