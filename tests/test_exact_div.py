"""The kernels divide by constants with an FMA sequence instead of DDIV
(cemc_kernels.cuh: exact_div).  Check the algorithm with exact rational
arithmetic: it must return the correctly rounded quotient, i.e. what the
reference's `/` (ce_updater.cpp:369,:402) computes."""
import math
import random
from fractions import Fraction


def _rn(fr):
    return float(fr)          # Fraction -> float is correctly rounded


def _fma(a, b, c):
    return _rn(Fraction(a) * Fraction(b) + Fraction(c))


def exact_div(a, b):
    y = 1.0 / b
    q0 = a * y
    r0 = _fma(-b, q0, a)
    q1 = _fma(r0, y, q0)
    r1 = _fma(-b, q1, a)
    return _fma(r1, y, q1)


def _rand_double(rng):
    m = rng.getrandbits(52)
    return math.ldexp(1.0 + m / 2 ** 52, rng.randint(-60, 60)) * rng.choice([-1, 1])


def test_exact_division_matches_ieee():
    rng = random.Random(1)
    dens = [1000 * 24.0, 8000 * 12.0, 64 * 6.0, 1728 * 8.0, 262144 * 24.0, 1000.0,
            8000.0, 0.0430866, 0.017, 3.0, 7.0, 1e-3, 123456789.0]
    for it in range(40000):
        b = rng.choice(dens) if it % 2 else abs(_rand_double(rng))
        a = _rand_double(rng)
        if it % 5 == 0:
            a = _rand_double(rng) * b        # quotients that are (nearly) exact
        assert exact_div(a, b) == a / b
