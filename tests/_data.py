"""Seeded synthetic inputs and the parity comparator shared by the tests.

Parity bar (BASELINE.json north_star): genotype decode, counts, MAF/missing.rate and routing are
bit-exact; p-values agree to |dlog10 p| <= 1e-6 * |log10 p| (plus 1e-9 absolute in log10 units so
that p ~ 1 is comparable); score statistics to 1e-9 relative.
"""
import numpy as np

from spagxecct_b200.bedio import pack_rows
from spagxecct_b200.harness import get_resid

P_RTOL = 1e-6
P_ATOL_LOG10 = 1e-9
STAT_RTOL = 1e-9


def make_pheno(n, seed, prevalence=0.1, trait="binary", n_extra_cov=0):
    rng = np.random.default_rng(seed)
    X1 = rng.standard_normal(n)
    X2 = rng.binomial(1, 0.5, n).astype(np.float64)
    E = rng.standard_normal(n)
    extra = [rng.standard_normal(n) for _ in range(n_extra_cov)]
    eta = 0.5 * X1 + 0.5 * X2 + 0.5 * E
    if trait == "binary":
        b0 = np.log(prevalence / (1 - prevalence)) - 0.4
        Y = rng.binomial(1, 1 / (1 + np.exp(-(b0 + eta)))).astype(np.float64)
    else:
        Y = 1.0 + eta + rng.standard_normal(n) * 0.5
    covs = [X1, X2] + extra
    R = get_resid(trait, Y, covs + [E])
    return dict(Y=Y, E=E, Cova=np.column_stack(covs), R=R, trait=trait)


def make_geno(n, mafs, seed, miss_rates=None, count_major=None, effect=None, Y=None):
    """(N, M) float matrix with NaN for missing. count_major[j]: code the major allele (MAF_raw > .5).
    effect[j] != 0 shifts allele frequency with Y to create marginal genetic signals (branch B)."""
    rng = np.random.default_rng(seed)
    m = len(mafs)
    G = np.empty((n, m), dtype=np.float64)
    for j, p in enumerate(mafs):
        pj = np.full(n, p)
        if effect is not None and effect[j] != 0 and Y is not None:
            yc = (Y - Y.mean()) / (Y.std() + 1e-12)
            pj = np.clip(p * np.exp(effect[j] * yc), 1e-6, 0.95)
        g = rng.binomial(2, pj).astype(np.float64)
        if count_major is not None and count_major[j]:
            g = 2 - g
        if miss_rates is not None and miss_rates[j] > 0:
            g[rng.random(n) < miss_rates[j]] = np.nan
        G[:, j] = g
    return G


def to_bed_rows(G):
    return pack_rows(np.where(np.isnan(G), -1, G))


def assert_parity(got, ref, p_cols, stat_cols, exact_cols=(0, 1), label=""):
    got = np.asarray(got); ref = np.asarray(ref)
    assert got.shape == ref.shape, f"{label}: shape {got.shape} vs {ref.shape}"
    assert np.array_equal(np.isnan(got), np.isnan(ref)), f"{label}: NA pattern differs"
    for c in exact_cols:
        a, b = got[:, c], ref[:, c]
        ok = (a == b) | (np.isnan(a) & np.isnan(b))
        assert ok.all(), f"{label}: column {c} not bit-exact at rows {np.where(~ok)[0][:5]}: {a[~ok][:3]} vs {b[~ok][:3]}"
    worst = 0.0
    for c in p_cols:
        a, b = got[:, c], ref[:, c]
        msk = ~np.isnan(b)
        z = msk & (b == 0)
        assert (a[z] == 0).all(), f"{label}: column {c}: reference p == 0 but got {a[z][:3]}"
        msk &= b != 0
        assert (a[msk] > 0).all(), f"{label}: column {c}: non-positive p-value"
        la, lb = np.log10(a[msk]), np.log10(b[msk])
        err = np.abs(la - lb)
        tol = P_RTOL * np.abs(lb) + P_ATOL_LOG10
        bad = err > tol
        assert not bad.any(), (f"{label}: p column {c}: {bad.sum()} rows out of tolerance, worst "
                               f"|dlog10p|={err.max():.3e} at p={b[msk][err.argmax()]:.3e}")
        if err.size:
            worst = max(worst, float(np.max(err / np.maximum(np.abs(lb), 1e-3))))
    for c in stat_cols:
        a, b = got[:, c], ref[:, c]
        msk = ~np.isnan(b)
        assert np.allclose(a[msk], b[msk], rtol=STAT_RTOL, atol=1e-12), f"{label}: stat column {c} differs"
    return worst


def routes_from_cct_output(out, epsilon):
    """(filtered, branchB, spa) boolean vectors recovered from the output matrix alone."""
    filtered = np.isnan(out[:, 6])
    branch_b = ~filtered & (out[:, 6] <= epsilon)
    spa = ~filtered & (out[:, 2] != out[:, 5]) & ~np.isnan(out[:, 2])
    return filtered, branch_b, spa
