"""The oracle's Rmath substitutes against mpmath."""
import numpy as np
import pytest

from oracle import np_oracle


@pytest.mark.parametrize("df", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("x", [1e-12, 1e-6, 1e-3, 0.1, 0.5, 1.0, 2.0, 3.9, 10.0, 37.0, 100.0, 1500.0, 5000.0, 20000.0])
def test_pchisq_upper_log_vs_mpmath(oracle, x, df):
    got = oracle.pchisq_upper_log(x, df)
    want = np_oracle.pchisq_upper_log_mp(x, df)
    assert np.isfinite(got)
    # p-value relative error = absolute error of log p; also demand relative accuracy of log p itself
    assert abs(got - want) <= 1e-13 * max(1.0, abs(want)) or abs(got - want) <= 2e-15 * abs(want) + 1e-300


def test_pchisq_edges(oracle):
    assert oracle.pchisq_upper_log(0.0, 1) == 0.0
    assert oracle.pchisq_upper_log(-3.0, 1) == 0.0
    assert oracle.pchisq_upper_log(np.inf, 1) == -np.inf
    assert np.isnan(oracle.pchisq_upper_log(np.nan, 1))
    # far beyond double underflow of p itself: log p must stay finite (log10 p ~ -1088)
    lp = oracle.pchisq_upper_log(5000.0, 1)
    assert abs(lp / np.log(10) - np_oracle.pchisq_upper_log_mp(5000.0, 1) / np.log(10)) < 1e-9


@pytest.mark.parametrize("x", [-40.0, -37.6, -37.0, -20.0, -8.3, -3.0, -1e-3, 0.0, 0.7, 3.0, 8.2, 8.3, 20.0, 37.0, 37.6])
@pytest.mark.parametrize("lower", [True, False])
def test_pnorm_vs_mpmath(oracle, x, lower):
    got = oracle.pnorm(x, lower)
    want = np_oracle.pnorm_mp(x, lower)
    # R returns exactly 0 past |x| = 37.5193 on the small side (nmath/pnorm.c); mpmath gives ~1e-309
    if want < 1e-307:
        assert got == 0.0 or abs(got - want) <= 1e-10 * want
    else:
        # 0.5*erfc(x/sqrt2): the rounding of x/sqrt2 is amplified by ~x^2 in the far tail (1.5e-13 at |x|=37);
        # four orders below the 1e-8 p-value tolerance of BASELINE.json
        assert abs(got - want) <= 1e-12 * want
