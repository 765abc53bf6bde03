"""Host-side helpers of the ACML wrappers that need no GPU: the Poisson pmf table (the reference uses scipy.stats.poisson.pmf,
car_rental.py:165-198; the package must not depend on scipy) and the Gambler win threshold."""
import math

import numpy as np
import pytest

from rl_envs_forge_b200 import acml


@pytest.mark.parametrize("lam", [0, 0.3, 2, 3, 4, 9.5])
def test_poisson_pmf_matches_scipy(lam):
    from scipy.stats import poisson
    for k in range(0, 25):
        want = float(poisson.pmf(k, lam))
        got = acml._poisson_pmf(k, lam)
        assert got == pytest.approx(want, rel=1e-12, abs=1e-300)


@pytest.mark.parametrize("p", [0.0, 0.25, 0.4, 0.5, 1.0 / 3.0, 0.999999999, 1.0])
def test_gambler_integer_threshold_equals_float_comparison(p):
    """`np.random.rand() < p` on u = r * 2**-32 is exactly `r < ceil(p * 2**32)` (what VecGamblersProblem passes down)."""
    thr = min(int(math.ceil(p * 4294967296.0)), 1 << 32)
    rng = np.random.default_rng(0)
    r = np.concatenate([rng.integers(0, 2 ** 32, 20000, dtype=np.uint64),
                        np.clip(np.array([thr - 2, thr - 1, thr, thr + 1, 0, 2 ** 32 - 1], dtype=np.int64), 0, 2 ** 32 - 1).astype(np.uint64)])
    assert np.array_equal(r.astype(np.float64) * 2.0 ** -32 < p, r < thr)
