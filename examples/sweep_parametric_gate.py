"""Switching from QSimulator.jl: a lab-frame parametric two-transmon gate (the system of
/root/reference/benchmark/benchmarks.jl:125-146 with the transmon of test/integration_tests.jl:86, whose second
flux-modulation sideband brings |01> and |10> into resonance near 118 MHz) swept over drive amplitude and
frequency in ONE call.

In the reference this is a Julia loop that rebuilds the CompositeQSystem and calls unitary_propagator per point:

    for amp in amps, f in freqs
        cqs = CompositeQSystem([q0, q1]); add_hamiltonian!(cqs, q0)
        add_hamiltonian!(cqs, parametric_drive(q1, t -> amp*sin(2pi*f*t)), q1)
        add_hamiltonian!(cqs, 0.006*X([q0, q1]), [q0, q1])
        us = unitary_propagator(cqs, 0:200)
        pop[amp, f] = abs2(us[end][4, 2])        # |01> -> |10| transfer
    end

Here the drive's amplitude and frequency are Param(0), Param(1): columns of a parameter matrix, one row per
sweep point.  Needs a CUDA device (there is no CPU fallback).

    python examples/sweep_parametric_gate.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "qsimulator.jl_b200", "host")]
import numpy as np

from qsim_b200 import (CompositeQSystem, DuffingSpec, DuffingTransmon, Param, PerturbativeTransmon, Sin, TransmonSpec,
                       X, add_hamiltonian, average_gate_fidelity, parametric_drive, subspace_block, unitary_propagator)

q0 = DuffingTransmon("q0", 3, DuffingSpec(3.94015, -0.1807))
q1 = PerturbativeTransmon("q1", 3, TransmonSpec(0.172, 12.71, 3.69))
cqs = CompositeQSystem([q0, q1])
add_hamiltonian(cqs, q0)
add_hamiltonian(cqs, parametric_drive(q1, Sin(amp=Param(0), freq=Param(1))), q1)
add_hamiltonian(cqs, 0.006 * X([q0, q1]), [q0, q1])

amps = np.linspace(0.28, 0.36, 32)
freqs = np.linspace(0.112, 0.124, 32)
params = np.array([(a, f) for a in amps for f in freqs])
ts = np.arange(0.0, 151.0)

# full propagators: [n_points, n_ts, 9, 9]
us = unitary_propagator(cqs, ts, params=params)
print("propagators:", us.shape)

# or keep only what the sweep needs, reduced on the device: the |01> -> |10> population ...
i01, i10 = 0 * 3 + 1, 1 * 3 + 0
pop = unitary_propagator(cqs, ts, params=params, select=[(i10, i01)], abs2=True)[:, :, 0]
point, k = np.unravel_index(np.argmax(pop), pop.shape)
print(f"largest 01->10 transfer {pop[point, k]:.4f} at t = {ts[k]:.0f} ns, amp = {params[point, 0]:.3f} Phi0, "
      f"f = {1e3 * params[point, 1]:.2f} MHz")

# ... or the qubit-subspace block, for a gate fidelity against iSWAP (up to single-qubit phases this is only a
# rough figure of merit; the point is that 16 numbers per propagator leave the GPU instead of 81)
sel, shape = subspace_block(cqs.dims)
sub = unitary_propagator(cqs, ts[[0, -1]], params=params, select=sel)[:, -1].reshape((-1,) + shape)
iswap = np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]])
print("best |F(iSWAP)| over the sweep:", float(np.max(average_gate_fidelity(sub, iswap))))
