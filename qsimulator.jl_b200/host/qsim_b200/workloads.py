"""Workload definitions: the BASELINE.json configs as CompositeQSystems + sweep grids.

cfg1/cfg2/cfg3 restate /root/reference/benchmark/benchmarks.jl:120-155 ("lab frame parametric
2Q gate", n = 2 or 3 transmons); the flux drive `t -> sin(2 pi f t)` becomes the sweepable
`amp*sin(2 pi f t)` with amp = Param(0), f = Param(1).  cfg4/cfg5 add decay and dephasing on every
transmon (operators.jl:185-203) with rates gamma = 1/(2 pi T), swept through Lindblad scale
parameters.  Grids follow SURVEY.md section 8(d).
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

from .composite import CompositeQSystem, add_hamiltonian, add_lindblad
from .operators import (ParametricLindblad, X, X_Y, lowering, number, parametric_drive, xy_exchange_drive)
from .signals import Const, Param, Sin
from .systems import DuffingSpec, DuffingTransmon, PerturbativeTransmon, TransmonSpec, hamiltonian

MOD_FREQ = 122.1 / 1e3   # benchmarks.jl:133
MOD_AMP = 0.323          # benchmarks.jl:134


def lab_frame_parametric_gate(n: int = 2, sweep: bool = True, amp=1.0, freq=MOD_FREQ):
    """benchmarks.jl:125-146.  With sweep=True the drive is Param(0)*sin(2 pi Param(1) t)."""
    q0 = DuffingTransmon("q0", 3, DuffingSpec(3.94015, -0.1807))
    q1 = PerturbativeTransmon("q1", 3, TransmonSpec(0.172, 16.4, 0.55))
    spectators = [DuffingTransmon(f"q{ct}", 3, DuffingSpec(4.0 + 0.1 * ct, -(0.2 + 0.01 * ct))) for ct in range(2, n)]
    all_qs = [q0, q1] + spectators
    cqs = CompositeQSystem(all_qs)
    add_hamiltonian(cqs, q0)
    drive = Sin(amp=Param(0), freq=Param(1)) if sweep else Sin(amp=amp, freq=freq)
    add_hamiltonian(cqs, parametric_drive(q1, drive), q1)
    add_hamiltonian(cqs, 0.006 * X([q0, q1]), [q0, q1])
    for q in all_qs[2:]:
        add_hamiltonian(cqs, q)
        add_hamiltonian(cqs, 0.006 * X([q, q1]), [q, q1])
    return cqs


def rotating_frame_parametric_gate(n: int = 2, sweep: bool = True, amp=0.245, freq=115.5 / 1e3):
    """benchmarks.jl:162-196 (rotating frame at each qubit's own frequency)."""
    q0 = DuffingTransmon("q0", 3, DuffingSpec(3.94015, -0.1807))
    q1 = PerturbativeTransmon("q1", 3, TransmonSpec(0.172, 16.4, 0.55))
    spectators = [DuffingTransmon(f"q{ct}", 3, DuffingSpec(4.0 + 0.1 * ct, -(0.2 + 0.01 * ct))) for ct in range(2, n)]
    all_qs = [q0, q1] + spectators
    cqs = CompositeQSystem(all_qs)
    add_hamiltonian(cqs, q0)
    add_hamiltonian(cqs, -q0.spec.frequency * number(q0), q0)
    q1_freq = hamiltonian(q1, 0.0)[1, 1]
    add_hamiltonian(cqs, -q1_freq * number(q1), q1)
    diff_freq = q0.spec.frequency - q1_freq
    add_hamiltonian(cqs, xy_exchange_drive([q0, q1], 0.006 * .5, [diff_freq, 0.0]), [q0, q1])
    drive = Sin(amp=Param(0), freq=Param(1)) if sweep else Sin(amp=amp, freq=freq)
    add_hamiltonian(cqs, parametric_drive(q1, drive), q1)
    for q in all_qs[2:]:
        add_hamiltonian(cqs, q)
        add_hamiltonian(cqs, -q.spec.frequency * number(q), q)
        diff = q1_freq - q.spec.frequency
        add_hamiltonian(cqs, xy_exchange_drive([q1, q], 0.006 * .5, [diff, 0.0]), [q1, q])
    return cqs


def add_t1_tphi(cqs: CompositeQSystem, first_param: int):
    """decay(q, g1) and dephasing(q, gphi) on every subsystem; the per-trajectory parameters are the
    operator scales sqrt(g1) and sqrt(2 gphi) (operators.jl:185-203), shared by all subsystems:
    columns first_param, first_param+1."""
    for q in cqs.subsystems:
        add_lindblad(cqs, ParametricLindblad(lowering(q), Const(Param(first_param))), [q])
        add_lindblad(cqs, ParametricLindblad(number(q), Const(Param(first_param + 1))), [q])


def lindblad_scales(T1_ns, Tphi_ns) -> Tuple[np.ndarray, np.ndarray]:
    """(sqrt(gamma1), sqrt(2 gamma_phi)) with gamma = 1/(2 pi T) (operators.jl:180,196)."""
    g1 = 1.0 / (2 * math.pi * np.asarray(T1_ns, dtype=float))
    gp = 1.0 / (2 * math.pi * np.asarray(Tphi_ns, dtype=float))
    return np.sqrt(g1), np.sqrt(2 * gp)


def amp_freq_grid(n_amp: int, n_freq: int) -> np.ndarray:
    """amp in [0.20, 0.45] Phi0 x f in [0.110, 0.134] GHz; row = (amp, f), f fastest."""
    amps = np.linspace(0.20, 0.45, n_amp)
    freqs = np.linspace(0.110, 0.134, n_freq)
    a, f = np.meshgrid(amps, freqs, indexing="ij")
    return np.ascontiguousarray(np.stack([a.ravel(), f.ravel()], axis=1))


def t1_tphi_grid(n_t1: int, n_tphi: int) -> np.ndarray:
    """log-spaced T1 in [5, 100] us x Tphi in [5, 200] us -> (sqrt(g1), sqrt(2 gphi)) rows."""
    t1 = np.geomspace(5e3, 100e3, n_t1)
    tp = np.geomspace(5e3, 200e3, n_tphi)
    a, b = np.meshgrid(t1, tp, indexing="ij")
    s1, sp = lindblad_scales(a.ravel(), b.ravel())
    return np.ascontiguousarray(np.stack([s1, sp], axis=1))


def config(name: str):
    """Returns (cqs, params[n_traj, n_params], ts, which) for a BASELINE.json config.
    which: 'unitary_propagator' | 'me_propagator'."""
    if name == "cfg1":
        cqs = lab_frame_parametric_gate(2, sweep=True)
        return cqs, np.array([[1.0, MOD_FREQ]]), np.arange(0.0, 201.0), "unitary_propagator"
    if name == "cfg2":
        return lab_frame_parametric_gate(2), amp_freq_grid(64, 64), np.arange(0.0, 201.0), "unitary_propagator"
    if name == "cfg3":
        return lab_frame_parametric_gate(3), amp_freq_grid(128, 128), np.arange(0.0, 201.0), "unitary_propagator"
    if name == "cfg4":
        cqs = lab_frame_parametric_gate(2)
        add_t1_tphi(cqs, 2)
        tt = t1_tphi_grid(128, 64)
        params = np.concatenate([np.tile([MOD_AMP, MOD_FREQ], (tt.shape[0], 1)), tt], axis=1)
        return cqs, np.ascontiguousarray(params), np.array([0.0, 20.0]), "me_propagator"
    if name == "cfg5":
        cqs = rotating_frame_parametric_gate(3)
        add_t1_tphi(cqs, 2)
        af = amp_freq_grid(16, 16)
        tt = t1_tphi_grid(16, 16)
        params = np.concatenate([np.repeat(af, tt.shape[0], axis=0), np.tile(tt, (af.shape[0], 1))], axis=1)
        return cqs, np.ascontiguousarray(params), np.array([0.0, 10.0]), "me_propagator"
    raise KeyError(name)
