"""CPU tests: the oracle against the reference's own golden vectors (decode) and against
independent high-precision restatements (everything else is 'parity unpinned')."""
import numpy as np
import pytest

from tests._data import make_geno, make_pheno


@pytest.mark.parametrize("name", ["SPAGxE", "SPAGxEmix", "SPAGxE_Plus"])
def test_bed_decode_matches_plink_raw(oracle, golden_dir, name):
    """inst/GenoMat_*.raw (plink --recode A) is the reference's own golden decode of inst/GenoMat_*.bed."""
    d = np.load(f"{golden_dir}/bed_{name}.npz")
    bed, raw, n = d["bed"], d["raw"], int(d["n"])
    assert bytes(bed[:3]) == bytes([0x6C, 0x1B, 0x01])
    G = oracle.bed_decode(bed[3:], n)
    assert G.shape == raw.shape
    want = raw.astype(np.float64)
    want[raw < 0] = np.nan
    assert np.array_equal(np.isnan(G), np.isnan(want))
    assert np.array_equal(np.nan_to_num(G, nan=-1), np.nan_to_num(want, nan=-1))
    cnt = oracle.bed_counts(bed[3:], n)
    assert np.array_equal(cnt[:, 0], (raw == 2).sum(0))
    assert np.array_equal(cnt[:, 2], (raw == 1).sum(0))
    assert np.array_equal(cnt[:, 3], (raw == 0).sum(0))
    assert np.array_equal(cnt[:, 1], (raw < 0).sum(0))


def test_rda_matrix_pins_a1_count_convention(golden_dir):
    d = np.load(f"{golden_dir}/plus.npz")
    assert bool(d["rda_first100_equals_bed_a1_count"])
    assert bool(d["fam_equals_pheno_iid"])


def test_pnorm_against_mpmath(oracle):
    import mpmath as mp
    mp.mp.dps = 50
    xs = [-37.0, -30.0, -20.5, -10.0, -8.3, -5.7, -5.6, -3.0, -2.0, -1.0, -0.7, -0.67, -0.3, -1e-9, 0.0, 0.5, 1.5, 4.0,
          6.0, 8.0, 12.0, 25.0, 37.0]
    for x in xs:
        for lower in (True, False):
            want = mp.ncdf(x) if lower else mp.ncdf(-x)
            got = oracle.pnorm(x, lower)
            rel = abs((mp.mpf(got) - want) / want)
            assert rel < 5e-15, (x, lower, got, float(want), float(rel))
    assert oracle.pnorm(-37.5, True) > 0 and oracle.pnorm(-37.6, True) == 0.0  # R's hard cut at -37.5193


def test_pt2_against_scipy(oracle):
    from scipy import stats
    for df in [5.0, 30.0, 9990.0, 499990.0]:
        for t in [0.01, 0.5, 1.0, 2.0, 3.5, 6.0, 12.0, 40.0]:
            want = 2 * stats.t.sf(t, df)
            got = oracle.pt2(t, df)
            assert abs(got - want) <= 1e-10 * want, (df, t, got, want)


def test_cct_semantics(oracle):
    import mpmath as mp
    mp.mp.dps = 40
    rc, p, _ = oracle.cct2(0.01, 0.2)
    want = 0.5 - mp.atan((mp.tan((mp.mpf(0.5) - mp.mpf(0.01)) * mp.pi) + mp.tan((mp.mpf(0.5) - mp.mpf(0.2)) * mp.pi)) / 2) / mp.pi
    assert rc == 0 and abs(p - float(want)) < 1e-15
    assert oracle.cct2(0.0, 0.3)[:2] == (0, 0.0)
    assert oracle.cct2(float("nan"), 0.3)[0] == 1        # stop("Cannot have NAs in the p-values!")
    assert oracle.cct2(1.2, 0.3)[0] == 2                 # stop("All p-values must be between 0 and 1!")
    assert oracle.cct2(0.0, 1.0)[0] == 3                 # stop("Cannot have both 0 and 1 p-values!")
    rc, p, warn = oracle.cct2(1.0, 0.3)
    assert rc == 0 and warn == 1
    rc, p, _ = oracle.cct2(1e-20, 0.5)                   # is.small branch and stat > 1e15 branch
    assert rc == 0 and abs(p / 2e-20 - 1) < 1e-12
    rc, p, _ = oracle.cct2(1e-13, 1e-12)                 # `1 - pcauchy` lands on the 2^-53 grid
    assert rc == 0 and (p / 2.0 ** -53) == round(p / 2.0 ** -53)


def _spa_reference_mp(R, maf, s, lower):
    """Independent restatement of fastgetroot_H1_new + GetProb_SPA_G_new in 40-digit arithmetic
    (same Newton path: same starting point, damping and tol = 2^-13 stop-before-update rule)."""
    import mpmath as mp
    R = [mp.mpf(float(r)) for r in R]
    maf = mp.mpf(maf); s = mp.mpf(s)

    def K1(t):
        return sum(r * (2 * maf * mp.e ** (t * r)) / (1 - maf + maf * mp.e ** (t * r)) for r in R)

    def K2(t):
        tot = 0
        for r in R:
            x = maf * mp.e ** (t * r); u = 1 - maf + x
            tot += r * r * 2 * x * (1 - maf) / (u * u)
        return tot

    def K0(t):
        return sum(2 * mp.log(1 - maf + maf * mp.e ** (t * r)) for r in R)

    tol = mp.mpf(2) ** -13
    t = mp.mpf(0); H1 = mp.mpf(0); diff = mp.inf
    for it in range(1, 101):
        old_t, old_diff, old_H1 = t, diff, H1
        H1 = K1(t) - s; H2 = K2(t)
        diff = -H1 / H2
        if mp.sign(H1) != mp.sign(old_H1):
            while abs(diff) > abs(old_diff) - tol:
                diff = diff / 2
        if abs(diff) < tol:
            break
        t = old_t + diff
    w = mp.sign(t) * mp.sqrt(2 * (t * s - K0(t)))
    v = t * mp.sqrt(K2(t))
    x = w + mp.log(v / w) / w
    return float(mp.ncdf(x) if lower else mp.ncdf(-x)), it


def test_spa_tail_against_mpmath(oracle):
    rng = np.random.default_rng(7)
    n = 400
    y = rng.binomial(1, 0.08, n).astype(float)
    R = y - y.mean() + rng.standard_normal(n) * 0.01
    for maf, s, lower in [(0.02, 3.5, False), (0.02, -1.2, True), (0.3, 14.0, False), (0.005, 2.2, False)]:
        got, it = oracle.getprob_spa(R, maf, s, lower)
        want, it_mp = _spa_reference_mp(R, maf, s, lower)
        assert it == it_mp
        assert abs(np.log10(got) - np.log10(want)) < 1e-9 * abs(np.log10(want)) + 1e-12, (maf, s, got, want)


def test_oracle_c1_bundled_example(oracle, golden_dir):
    """config 1: data/Geno.mtx + Pheno.mtx. rs8 is the SPA SNP (z ~ -2.70); values as probed in SURVEY.md section 4."""
    from spagxecct_b200.harness import get_resid
    d = np.load(f"{golden_dir}/c1.npz")
    R = get_resid("binary", d["y"], [d[k] for k in ["Cov1", "Cov2", "E", "PC1", "PC2", "PC3", "PC4"]])
    cova = np.column_stack([d[k] for k in ["PC1", "PC2", "PC3", "PC4", "Cov1", "Cov2"]])
    out, route, iters, rc = oracle.cct_matrix("binary", d["geno"].astype(float), R, d["E"], d["y"], cova, g_is_int=True)
    assert rc == 0
    assert list(np.where(route & 4)[0]) == [7]
    assert tuple(iters[7]) == (3, 4)
    assert abs(out[7, 2] - 7.374e-3) < 5e-6 and abs(out[7, 5] - 6.895e-3) < 5e-6
    assert np.allclose(out[:, 0], d["geno"].mean(0) / 2)


def test_mean_mode_r_faithful_differs_by_at_most_one_ulp(oracle):
    """The MAF contract RN53(A/n_obs)/2 vs R's long-double mean: count the 1-ulp cases."""
    ph = make_pheno(3000, 1)
    G = make_geno(3000, np.linspace(0.01, 0.49, 300), 2, miss_rates=np.linspace(0, 0.05, 300))
    try:
        oracle.set_mean_mode(0)
        a, *_ = oracle.cct_matrix("binary", G, ph["R"], ph["E"], ph["Y"], ph["Cova"])
        oracle.set_mean_mode(1)
        b, *_ = oracle.cct_matrix("binary", G, ph["R"], ph["E"], ph["Y"], ph["Cova"])
    finally:
        oracle.set_mean_mode(0)
    ulp = np.abs(a[:, 0] - b[:, 0]) / np.spacing(a[:, 0])
    assert ulp.max() <= 1.0
    lp = np.abs(np.log10(a[:, 2]) - np.log10(b[:, 2]))
    assert np.nanmax(lp) < 1e-9


def test_glm_and_lm_against_numpy(oracle):
    rng = np.random.default_rng(3)
    n = 4000
    X = np.column_stack([np.ones(n), rng.standard_normal((n, 4))])
    beta = np.array([-2.0, 0.5, -0.3, 0.2, 0.0])
    y = rng.binomial(1, 1 / (1 + np.exp(-X @ beta))).astype(float)
    rc, coef, se, it, cv, mu = oracle.glm_binomial(X, y)
    assert rc == 0 and cv == 1
    from spagxecct_b200.harness import logistic_fit
    b2, mu2 = logistic_fit(X, y)
    assert np.allclose(coef, b2, rtol=1e-6, atol=1e-9)
    W = mu * (1 - mu)
    # summary.glm takes the covariance from the last WLS (weights at the previous iterate): close to, not equal to, the MLE one
    assert np.allclose(se, np.sqrt(np.diag(np.linalg.inv(X.T @ (X * W[:, None])))), rtol=1e-4)
    yq = X @ beta + rng.standard_normal(n)
    rc, coef, se, pv, fit = oracle.lm(X, yq)
    bq, *_ = np.linalg.lstsq(X, yq, rcond=None)
    assert rc == 0 and np.allclose(coef, bq, rtol=1e-10)
    from scipy import stats
    res = yq - X @ bq
    s2 = res @ res / (n - 5)
    se_np = np.sqrt(np.diag(np.linalg.inv(X.T @ X)) * s2)
    assert np.allclose(se, se_np, rtol=1e-10)
    assert np.allclose(pv, 2 * stats.t.sf(np.abs(bq / se_np), n - 5), rtol=1e-9)
