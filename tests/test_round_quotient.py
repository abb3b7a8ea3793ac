"""Host model of `round_quotient` (csrc/current.cuh): the sampler rounds a / b - sub through a * (1 / b) and falls back to the
exact division within 1e-9 of a tie.  IEEE float64 multiply / subtract / rint are the same on the host, so the claim -- whenever
the fast path is taken its integer equals round-half-even of the exactly divided expression (detsim.py:208-218, 333) -- can be
checked here on millions of values, adversarial ones included."""
import numpy as np


def fast_path(a, b, sub):
    inv_b = 1.0 / b
    with np.errstate(invalid="ignore"):
        v = a * inv_b - sub
        r = np.rint(v)
        slow = (np.abs(np.abs(v - r) - 0.5) < 1e-9) | ~(np.abs(v) < 4.0e9)
    return r, slow


def exact(a, b, sub):
    return np.rint(a / b - sub)          # np.rint == round half to even == Python round() on a float64


def check(a, b, sub):
    r, slow = fast_path(a, b, sub)
    want = exact(a, b, sub)
    assert np.array_equal(r[~slow], want[~slow])
    return slow.mean()


def test_random_quotients_match_exact_division():
    rng = np.random.default_rng(11)
    for b in (0.04434, 0.038, 0.05, 0.1, 0.0443400000001, 1.0 / 3.0):
        for sub in (0.5, 0.0):
            a = rng.uniform(0.0, 45.0 * b, 2_000_000)
            frac_slow = check(a, b, sub)
            assert frac_slow < 1e-6
            a = rng.uniform(0.0, 4000.0 * b, 2_000_000)          # tick differences / response sampling
            assert check(a, b, sub) < 1e-5


def test_values_at_and_next_to_ties():
    # a chosen so that a / b - sub sits on or within a few ulp of k + 0.5: every one of them must either take the exact path or agree
    for b in (0.04434, 0.05, 0.1):
        for sub in (0.5, 0.0):
            k = np.arange(0, 4000, dtype=np.float64)
            centre = (k + 0.5 + sub) * b
            cases = [centre]
            x = centre.copy()
            for _ in range(4):
                x = np.nextafter(x, np.inf); cases.append(x.copy())
            x = centre.copy()
            for _ in range(4):
                x = np.nextafter(x, -np.inf); cases.append(x.copy())
            a = np.concatenate(cases)
            r, slow = fast_path(a, b, sub)
            assert np.array_equal(r[~slow], exact(a, b, sub)[~slow])
            assert slow.mean() > 0.9                              # (and nearly all of these are recognised as ties)


def test_non_finite_and_huge_values_take_the_exact_path():
    a = np.array([np.inf, np.nan, 1e300, -1e300, 5e9 * 0.05])
    _, slow = fast_path(a, 0.05, 0.0)
    assert slow.all()
