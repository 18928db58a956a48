"""Host-side mirror vs the reference's own unit tests for operators, systems and embedding.

Mirrors /root/reference/test/test_operators.jl:4-38, test_composite_systems.jl:4-34 and
test_systems.jl:6-14, plus the known values of SURVEY.md section 8 row a12.
"""
import math

import numpy as np
import pytest

from qsim_b200 import (CompositeQSystem, Cos, DuffingSpec, DuffingTransmon, Param, PerturbativeTransmon,
                       Resonator, ResonatorSpec, RotatingFrameSystem, Sin, TransmonSpec, X, X_Y, Y,
                       add_hamiltonian, dipole_drive, embed, embed_indices, flip_flop, hamiltonian, lower,
                       lowering, number, parametric_drive, perturbative_transmon_anharm,
                       perturbative_transmon_freq, raising, rwa_dipole, xy_exchange_drive)
from qsim_b200.composite import add_parametric_hamiltonians, composite_hamiltonian

kron = np.kron


def test_primitives():
    r = Resonator("test", 3, ResonatorSpec(5.5))
    assert np.array_equal(raising(r), np.array([[0, 0, 0], [1, 0, 0], [0, math.sqrt(2), 0]]))
    assert np.array_equal(lowering(r), np.array([[0, 1, 0], [0, 0, math.sqrt(2)], [0, 0, 0]]))
    r = Resonator("test", 7, ResonatorSpec(5.5))
    assert np.allclose(raising(r) @ lowering(r), number(r))
    assert np.allclose(X(r), raising(r) + lowering(r))
    assert np.allclose(X(r, 0.25), Y(r))
    comm = lowering(r) @ raising(r) - raising(r) @ lowering(r)
    assert np.allclose(comm[:-1, :-1], np.eye(6))
    a = Resonator("a", 2, ResonatorSpec(5.5))
    b = Resonator("b", 2, ResonatorSpec(5.5))
    assert np.array_equal(flip_flop(a, b), np.array([[0, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 0]]))


def test_composite_hamiltonian():
    r1 = Resonator("r1", 3, ResonatorSpec(4.0))
    r2 = Resonator("r2", 2, ResonatorSpec(5.0))
    r3 = Resonator("r3", 2, ResonatorSpec(6.0))
    cqs = CompositeQSystem([r1, r2, r3])
    for q in (r1, r2, r3):
        add_hamiltonian(cqs, q)
    i2, i3 = np.eye(2), np.eye(3)
    m = kron(kron(hamiltonian(r1), i2), i2) + kron(kron(i3, hamiltonian(r2)), i2) + kron(kron(i3, i2), hamiltonian(r3))
    assert np.array_equal(hamiltonian(cqs), m)


def test_embed():
    CNOT = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]])
    SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
    Id = np.eye(2, dtype=int)
    assert np.array_equal(embed(CNOT, [0, 1], [2, 2]), CNOT)
    assert np.array_equal(embed(CNOT, [1, 0], [2, 2]), SWAP @ CNOT @ SWAP)
    assert np.array_equal(embed(CNOT, [0, 2], [2, 2, 2]), kron(Id, SWAP) @ kron(CNOT, Id) @ kron(Id, SWAP))
    m = np.array([[1, 2], [3, 4]])
    assert np.array_equal(embed(m, [1], [3, 2, 4]), kron(kron(np.eye(3, dtype=int), m), np.eye(4, dtype=int)))


def test_embed_indices_shapes():
    # 3x3 op in d=27 -> 9 lists x 9 idx; 9x9 op in d=27 -> 81 x 3
    assert [len(x) for x in embed_indices([1], [3, 3, 3])] == [9] * 9
    assert [len(x) for x in embed_indices([2, 1], [3, 3, 3])] == [3] * 81
    # every big index used at most once
    allidx = np.concatenate(embed_indices([2, 0], [3, 2, 4]))
    assert len(np.unique(allidx)) == len(allidx)


def test_duffing_and_rotating_frame():
    q = DuffingTransmon("q", 3, DuffingSpec(4.0, -0.2))
    assert np.allclose(hamiltonian(q), np.diag([0.0, 4.0, 2 * 4.0 - 0.2]))
    assert np.allclose(hamiltonian(q, 4.5), np.diag([0.0, 4.5, 9.0 - 0.2]))
    r = RotatingFrameSystem.wrap(q, 3.9)
    assert np.allclose(hamiltonian(r), np.diag([0.0, 0.1, 0.2 - 0.2]))


def test_perturbative_known_values():
    f, a = perturbative_transmon_freq, perturbative_transmon_anharm
    assert f(0.172, 16.4, 0.55, 0.0) == pytest.approx(4.650623526074448, rel=1e-14)
    assert a(0.172, 16.4, 0.55, 0.0) == pytest.approx(-0.18857115680014172, rel=1e-14)
    assert f(0.172, 12.71, 3.69, 0.0) == pytest.approx(4.5714982940738835, rel=1e-14)
    assert a(0.172, 12.71, 3.69, 0.0) == pytest.approx(-0.18890707185655456, rel=1e-14)
    assert f(0.172, 12.71, 3.69, 0.5) == pytest.approx(3.3412807356024463, rel=1e-14)
    assert a(0.172, 12.71, 3.69, 0.5) == pytest.approx(-0.19673936190192642, rel=1e-14)
    # num_terms == 1: plasma frequency sqrt(8 EJ EC), zero anharmonicity (docstrings :214,246)
    assert f(0.2, 20.0, 0.0, 0.0, num_terms=1) == pytest.approx(math.sqrt(8 * 20.0 * 0.2), rel=1e-14)
    assert a(0.2, 20.0, 0.0, 0.0, num_terms=1) == 0.0
    q = PerturbativeTransmon("q", 3, TransmonSpec(0.172, 16.4, 0.55))
    h = hamiltonian(q, 0.1)
    fq, aq = f(0.172, 16.4, 0.55, 0.1), a(0.172, 16.4, 0.55, 0.1)
    assert np.allclose(np.diag(h), [0, fq, 2 * fq + aq], rtol=1e-14)


def test_xy_exchange_identity():
    qa = DuffingTransmon("a", 3, DuffingSpec(4.0, -0.2))
    qb = DuffingTransmon("b", 3, DuffingSpec(4.1, -0.2))
    drv = xy_exchange_drive([qa, qb], 0.003, [0.37, 0.0])
    t = 1.234
    lt = drv.lowered()[0]
    th = 2 * math.pi * (lt.rate * t)
    sparse = 0.003 * (np.exp(1j * th) * lt.atom_a + np.exp(-1j * th) * lt.atom_b)
    assert np.allclose(drv(t), sparse, atol=1e-15)
    assert np.allclose(drv(t), 0.003 * X_Y([qa, qb], [0.37 * t, 0.0]), atol=0)


def test_lowering_matches_closures():
    """Descriptor lowering (what crosses the C ABI) reproduces the closure semantics."""
    q0 = DuffingTransmon("q0", 3, DuffingSpec(3.94015, -0.1807))
    q1 = RotatingFrameSystem.wrap(PerturbativeTransmon("q1", 3, TransmonSpec(0.172, 12.71, 3.69)), 3.9)
    q2 = DuffingTransmon("q2", 2, DuffingSpec(4.2, -0.22))
    cqs = CompositeQSystem([q0, q1, q2])
    add_hamiltonian(cqs, q0)
    add_hamiltonian(cqs, parametric_drive(q1, Sin(amp=Param(0), freq=0.12, offset=0.05)), q1)
    add_hamiltonian(cqs, parametric_drive(q0, Cos(amp=0.3, freq=0.3, offset=4.3)), q0)
    add_hamiltonian(cqs, dipole_drive(q2, Cos(amp=0.02, freq=4.2), 4.0), q2)
    add_hamiltonian(cqs, rwa_dipole(q2, Cos(amp=0.01, freq=0.01)), q2)
    add_hamiltonian(cqs, xy_exchange_drive([q2, q0], 0.003, [0.2, 0.0]), [q2, q0])
    spec = lower(cqs)
    assert spec.n_params == 1 and spec.d == 18
    # rebuild H(t) from the lowered terms by hand and compare with the closure path
    t, row = 0.777, np.array([0.31])
    h = composite_hamiltonian(cqs)
    add_parametric_hamiltonians(h, cqs, t, row)
    g = np.array(spec.h0)
    for lt, a, b, levels in spec.terms:
        s = lt.signal(t, row)
        if lt.kind == 0:
            g = g + s * a
        elif lt.kind == 1:
            th = 2 * math.pi * (lt.rate * t)
            g = g + s * (np.exp(1j * th) * a + np.exp(-1j * th) * b)
        else:
            if lt.kind == 2:
                f, al = s, lt.anharm
            else:
                f = perturbative_transmon_freq(lt.EC, lt.EJ1, lt.EJ2, s, lt.num_terms)
                al = perturbative_transmon_anharm(lt.EC, lt.EJ1, lt.EJ2, s, lt.num_terms)
            n = levels.astype(float)
            g = g + np.diag((f - 0.5 * al) * n + 0.5 * al * n * n - n * lt.rot)
    assert np.allclose(g, h, atol=1e-13)


def test_opaque_closure_rejected_by_lowering():
    q = DuffingTransmon("q", 3, DuffingSpec(5.0, -0.2))
    cqs = CompositeQSystem([q])
    add_hamiltonian(cqs, q)
    add_hamiltonian(cqs, lambda t: 0.01 * math.cos(t) * X(q), q)
    with pytest.raises(TypeError):
        lower(cqs)
    with pytest.raises(TypeError):
        dipole_drive(q, lambda t: 0.02 * math.cos(t))


def test_fit_transmon_and_spec_conversions():
    """test_systems.jl:15-22 (first part), :49-92 (PerturbativeTransmon model), :94-103."""
    from qsim_b200 import duffing_hamiltonian, fit_transmon, hamiltonian, transmon_spec_from_duffing
    # tunable: frequency and anharmonicity at phi = 0, frequency at phi = 1/2
    t = fit_transmon(4.0, 3.0, -0.2)
    h = hamiltonian(PerturbativeTransmon("q", 3, t), 0.0)
    assert np.allclose(h, np.diag([0.0, 4.0, 2 * 4.0 - 0.2]), rtol=1e-8, atol=1e-10)
    t = fit_transmon(5.0, 4.0, -0.2)
    assert t.EJ2 != 0.0
    assert np.allclose(hamiltonian(PerturbativeTransmon("q", 3, t), 0.0), duffing_hamiltonian(5.0, -0.2, 3), rtol=1e-7, atol=1e-10)
    assert np.allclose(hamiltonian(PerturbativeTransmon("q", 2, t), 0.5), duffing_hamiltonian(4.0, 0.0, 2), rtol=1e-8, atol=1e-10)
    # fixed frequency: EJ2 == 0
    t = fit_transmon(5.0, 5.0, -0.2)
    assert t.EJ2 == 0.0
    assert np.allclose(hamiltonian(PerturbativeTransmon("q", 3, t), 0.0), duffing_hamiltonian(5.0, -0.2, 3), rtol=1e-7, atol=1e-10)
    # DuffingSpec <-> TransmonSpec round trip
    d1 = DuffingSpec.from_transmon(transmon_spec_from_duffing(DuffingSpec(4.0, -0.2)))
    assert d1.frequency == pytest.approx(4.0, rel=1e-8) and d1.anharmonicity == pytest.approx(-0.2, rel=1e-8)
