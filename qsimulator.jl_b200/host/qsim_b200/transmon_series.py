"""Perturbative transmon series (flux -> f01, anharmonicity).

Host-side mirror of the scalar physics the reference evaluates on every
right-hand-side call of a flux-driven transmon:
  xi_effective                  /root/reference/src/perturbative_transmon.jl:187-189
  evalpoly (Horner with muladd) /root/reference/src/perturbative_transmon.jl:193-201
  perturbative_transmon_freq    /root/reference/src/perturbative_transmon.jl:222-231
  perturbative_transmon_anharm  /root/reference/src/perturbative_transmon.jl:253-261
The coefficient tables are the published expansion of Didier et al.,
arXiv:1706.06566, stored as (numerator, log2 denominator) exactly as the
reference stores them (perturbative_transmon.jl:38-105); a coefficient is
float(numerator) / 2**b, which is what `_prepare` (:167) computes.
"""
from __future__ import annotations

import math

PERTURBATIVE_NUM_TERMS = 10  # perturbative_transmon.jl:173

# (numerator, b): coefficient = numerator / 2**b
_FREQ_RATIONALS = (
    (-1, 0), (-1, 2), (-21, 7), (-19, 7), (-5319, 15), (-6649, 15),
    (-1180581, 22), (-446287, 20), (-1489138635, 31), (-648381403, 29),
    (-614557854099, 38), (-75265839129, 34), (-637411859250147, 46),
    (-86690561488017, 42), (-405768570324517701, 53),
    (-15191635582891041, 47), (-2497063196283456607731, 63),
    (-102281923716042917215, 57), (-2292687293949773041433127, 70),
    (-25544408245062216574759, 62), (-4971071120163260007203175705, 78),
    (-59956026877695226936825271, 70),
    (-6299936888270974385982624367587, 85),
    (-20465345194746565030172477629, 75),
    (-36984324599399309412347250837528543, 94),
    (-128862667153189778842334459173303, 84),
    (-62313306363032484263243187605857135455, 101),
    (-57979140436623262897403437875845329, 89),
    (-239123145585215826671902664445701932931163, 109),
    (-236766982175345008541229542667501202161, 97),
    (-518667194120793070334115565427490753019904133, 116),
)

_ANHARM_RATIONALS = (
    (-1, 0), (-9, 4), (-81, 7), (-3645, 12), (-46899, 15), (-1329129, 19),
    (-20321361, 22), (-2648273373, 28), (-45579861135, 31),
    (-1647988255539, 35), (-31160327412879, 38), (-2457206583272505, 43),
    (-50387904068904927, 46), (-2145673984043982897, 50),
    (-47368663010124907041, 53), (-17329540083222030375645, 60),
    (-410048712835835979799431, 63), (-20066784213453521778111375, 67),
    (-507447585299180759749453827, 70),
    (-53019019946496461235728807475, 75),
    (-1429754157181172012054040903645, 78),
    (-79571741391885949104006842758911, 82),
    (-2283773190022904454409743892590327, 85),
    (-540565733415401595950277192471356985, 91),
    (-16479511149218202447739080120870460083, 94),
    (-1034743270413623494225962156243473940687, 98),
    (-33436402163767100825528499521512789823595, 101),
    (-4445866752212713247387096882061024305480817, 106),
    (-151943275517394187333713519962683890787229463, 109),
    (-10671890872277478133986830879693284605566580129, 113),
    (-384886904723357410697832985058012690322303378753, 116),
)


def _prepare(table):
    # float(int) is correctly rounded, like Julia's parse of the Float64 literal;
    # the power-of-two division is exact.
    return tuple(float(a) / (2.0 ** b) for a, b in table)


PT_FREQ_PRE = _prepare(_FREQ_RATIONALS)
PT_ANH_PRE = _prepare(_ANHARM_RATIONALS)
MAX_NUM_TERMS = len(PT_FREQ_PRE) + 1  # 32


def xi_effective(EC, EJ1, EJ2, phi):
    return math.sqrt((2 * EC) / math.sqrt((EJ1 ** 2 + EJ2 ** 2) + (2 * EJ1 * EJ2) * math.cos(2 * math.pi * phi)))


def _horner(x, coeffs):
    if not coeffs:
        return 0.0
    ex = coeffs[-1]
    for c in reversed(coeffs[:-1]):
        ex = math.fma(x, ex, c) if hasattr(math, "fma") else x * ex + c
    return ex


def _check_terms(num_terms):
    if num_terms < 1:
        raise ValueError("num_terms < 1")
    return min(int(num_terms), MAX_NUM_TERMS)


def perturbative_transmon_freq(EC, EJ1, EJ2, phi, num_terms=PERTURBATIVE_NUM_TERMS):
    t = _check_terms(num_terms)
    xi = xi_effective(EC, EJ1, EJ2, phi)
    return EC * (4 / xi + _horner(xi, PT_FREQ_PRE[: t - 1]))


def perturbative_transmon_anharm(EC, EJ1, EJ2, phi, num_terms=PERTURBATIVE_NUM_TERMS):
    t = _check_terms(num_terms)
    xi = xi_effective(EC, EJ1, EJ2, phi)
    return EC * _horner(xi, PT_ANH_PRE[: t - 1])
