"""traits = "categorical": the Wald test of branch B is ordinal::clm(as.factor(Y) ~ Cova + E + g + g*E) (ref
R/SPAGxECCT.R:664-677).  `ordinal` is a third-party package that the reference neither ships nor imports, so the oracle
restates the model of that call (logit link, flexible thresholds) and its Wald summary.  CPU part: that restatement
against an independent numpy Newton solver written from the general cumulative-link likelihood.  GPU part (-m gpu):
k_clm_wald against the oracle."""
import numpy as np
import pytest

from spagxecct_b200 import _lib
from tests._data import assert_parity, make_geno, routes_from_cct_output, to_bed_rows


def _cumlink(theta, X, ycat, J):
    """log likelihood, gradient, Hessian of P(Y = j) = F(a_j'theta) - F(b_j'theta) with a = (e_j, -x), b = (e_{j-1}, -x):
    the generic two-design-matrix form (independent of the oracle's per-threshold bookkeeping)."""
    n, q = X.shape
    nth = J - 1
    A = np.zeros((n, nth + q)); B = np.zeros((n, nth + q))
    A[:, nth:] = -X; B[:, nth:] = -X
    up = ycat < nth; lo = ycat > 0
    A[np.nonzero(up)[0], ycat[up]] = 1
    B[np.nonzero(lo)[0], ycat[lo] - 1] = 1
    F = lambda t: 1 / (1 + np.exp(-t))
    ea, eb = A @ theta, B @ theta
    Fa = np.where(up, F(ea), 1.0); Fb = np.where(lo, F(eb), 0.0)
    fa = np.where(up, Fa * (1 - Fa), 0.0); fb = np.where(lo, Fb * (1 - Fb), 0.0)
    sa = fa * (1 - 2 * Fa); sb = fb * (1 - 2 * Fb)
    pi = Fa - Fb
    gi = (fa[:, None] * A - fb[:, None] * B) / pi[:, None]
    H = (A.T * (sa / pi)) @ A - (B.T * (sb / pi)) @ B - gi.T @ gi
    return np.log(pi).sum(), gi.sum(0), H


def _fit(X, ycat, J):
    q = X.shape[1]
    theta = np.r_[np.log(np.arange(1, J) / (J - np.arange(1, J))), np.zeros(q)]
    for _ in range(60):
        ll, g, H = _cumlink(theta, X, ycat, J)
        step = np.linalg.solve(-H, g)
        theta = theta + step
        if np.abs(step).max() < 1e-13:
            break
    _, _, H = _cumlink(theta, X, ycat, J)
    return theta, np.sqrt(np.diag(np.linalg.inv(-H)))


def _cat_pheno(n, seed, J, p=2, skew=False):
    rng = np.random.default_rng(seed)
    X1 = rng.standard_normal(n); X2 = rng.binomial(1, 0.5, n).astype(float); E = rng.standard_normal(n)
    eta = 0.5 * X1 - 0.3 * X2 + 0.4 * E
    lat = eta + rng.logistic(size=n)
    probs = np.linspace(0, 1, J + 1)[1:-1] if not skew else 1 - 0.5 ** np.arange(1, J)   # skew: 50 %, 25 %, ... of the samples
    cuts = np.quantile(lat, probs)
    ycat = np.searchsorted(cuts, lat).astype(np.int64)
    cova = np.column_stack([X1, X2])[:, :p]
    # Step-1 residuals stay outside the path: any centred vector works as R here
    R = (ycat - ycat.mean()) / J + 0.1 * rng.standard_normal(n)
    R = R - R.mean()
    return dict(ycat=ycat, E=E, Cova=cova, R=R, levels=np.array([10, 20, 35, 50, 70, 90, 110, 130])[:J])


@pytest.mark.parametrize("J,skew", [(2, False), (3, False), (5, True)])
def test_oracle_clm_matches_independent_newton(oracle, J, skew):
    n = 1500
    ph = _cat_pheno(n, 3 + J, J, skew=skew)
    rng = np.random.default_rng(4)
    g = rng.binomial(2, 0.3, n).astype(float)
    X = np.column_stack([ph["Cova"], ph["E"], g, ph["E"] * g])
    theta, se = _fit(X, ph["ycat"], J)
    coef, se_o, iters, rc = oracle.clm(X, ph["ycat"], J)
    assert rc == 0 and 3 <= iters <= 12
    assert np.allclose(coef, theta, rtol=1e-8, atol=1e-10)
    assert np.allclose(se_o, se, rtol=1e-8)
    if J == 2:   # two levels: logistic regression of 1{Y = upper level}; the threshold is minus the intercept
        Xd = np.column_stack([np.ones(n), X]); yb = (ph["ycat"] == 1).astype(float)
        b = np.zeros(Xd.shape[1])
        for _ in range(40):
            mu = 1 / (1 + np.exp(-(Xd @ b)))
            info = (Xd.T * (mu * (1 - mu))) @ Xd
            b = b + np.linalg.solve(info, Xd.T @ (yb - mu))
        se_l = np.sqrt(np.diag(np.linalg.inv(info)))
        assert np.allclose(coef[1:], b[1:], rtol=1e-7, atol=1e-9) and np.isclose(coef[0], -b[0], rtol=1e-7)
        assert np.allclose(se_o, se_l, rtol=1e-7)


def test_oracle_categorical_branch_b_is_clm_on_the_design(oracle):
    n = 2003; J = 4
    ph = _cat_pheno(n, 7, J)
    G = make_geno(n, np.full(6, 0.25), 8, miss_rates=[0, 0.03, 0, 0.03, 0, 0], count_major=[0, 0, 1, 1, 0, 0])
    out, route, iters, rc = oracle.cct_matrix("categorical", G, ph["R"], ph["E"], ph["ycat"].astype(float), ph["Cova"], epsilon=0.999)
    assert rc == 0 and ((route & 2) > 0).all()
    for j in range(G.shape[1]):
        g = G[:, j].copy(); maf = np.nanmean(g) / 2; g[np.isnan(g)] = 2 * maf
        if maf > 0.5:
            g = 2 - g
        X = np.column_stack([ph["Cova"], ph["E"], g, ph["E"] * g])
        coef, se, _, rc1 = oracle.clm(X, ph["ycat"], J)
        assert rc1 == 0
        assert out[j, 3] == pytest.approx(2 * oracle.pnorm(-abs(coef[-1] / se[-1])), rel=1e-12)


gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("J,p,skew", [(3, 2, False), (5, 2, True), (2, 1, False), (8, 0, False)])
def test_cct_categorical_wald(gpu_ctx, oracle, J, p, skew):
    n = 6007
    ph = _cat_pheno(n, 21 + J, J, p=p, skew=skew)
    mafs = np.r_[np.full(3, 0.0004), np.linspace(0.01, 0.5, 30), np.full(8, 0.2)]
    m = mafs.size
    rng = np.random.default_rng(5)
    miss = np.where(rng.random(m) < 0.5, 0.0, rng.random(m) * 0.06)
    major = rng.random(m) < 0.5
    effect = np.zeros(m); effect[-8:] = np.linspace(0.4, 1.2, 8)
    G = np.ascontiguousarray(make_geno(n, mafs, 12, miss_rates=miss, count_major=major, effect=effect,
                                       Y=(ph["ycat"] >= J // 2).astype(float)))
    rows = to_bed_rows(G)
    y = ph["ycat"].astype(np.float64)
    gpu_ctx.set_vectors(ph["R"], ph["E"])
    gpu_ctx.set_wald(_lib.TRAIT_CATEGORICAL, y, ph["Cova"])
    geno = gpu_ctx.geno_desc(_lib.GENO_BED, rows.ctypes.data, n, m, (n + 3) // 4)
    for eps in (0.001, 0.2):
        out, diag = gpu_ctx.run_cct(geno, dict(epsilon=eps))
        ref, route, iters, rc = oracle.cct_bed("categorical", rows, n, ph["R"], ph["E"], y, ph["Cova"], epsilon=eps, nthreads=8)
        assert rc == 0
        assert_parity(out, ref, [2, 3, 4, 5, 6], [7, 8, 9], label=f"categorical J={J} p={p} eps={eps}")
        f, b, s = routes_from_cct_output(out, eps)
        assert np.array_equal(b, (route & 2) > 0) and np.array_equal(s, (route & 4) > 0)
        assert diag["newton_iters"] == iters.sum() and diag["n_wald_fail"] == 0
        assert diag["clm_evals"] >= 3 * b.sum()
    assert b.sum() >= 8
    # the same call again (cached category order) gives identical bits
    out2, _ = gpu_ctx.run_cct(geno, dict(epsilon=0.2))
    assert np.array_equal(np.nan_to_num(out2, nan=-1), np.nan_to_num(out, nan=-1))


@gpu
def test_categorical_api_and_errors(gpu_ctx, oracle):
    from spagxecct_b200 import SPAGxE_CCT
    from spagxecct_b200._lib import SpagxeError
    n = 3001; J = 4
    ph = _cat_pheno(n, 9, J)
    G = make_geno(n, np.linspace(0.05, 0.4, 12), 10)
    ids = [f"id{i}" for i in range(n)]
    ylab = ph["levels"][ph["ycat"]]                      # arbitrary ordered labels: as.factor() sorts them
    res = SPAGxE_CCT("categorical", Geno_mtx=(G, ids, None), R=ph["R"], E=ph["E"], Phen_mtx={"ID": ids, "Y": ylab},
                     Cova_mtx=ph["Cova"], epsilon=0.5, ctx=gpu_ctx, quiet=True)
    ref, route, _, rc = oracle.cct_matrix("categorical", G, ph["R"], ph["E"], ph["ycat"].astype(float), ph["Cova"], epsilon=0.5)
    assert rc == 0 and res.diag["n_branch_b"] >= 3
    assert_parity(res.values, ref, [2, 3, 4, 5, 6], [7, 8, 9], label="categorical api")
    with pytest.raises(SpagxeError):
        gpu_ctx.set_wald(_lib.TRAIT_CATEGORICAL, np.zeros(n), ph["Cova"])                       # one level
    with pytest.raises(SpagxeError):
        gpu_ctx.set_wald(_lib.TRAIT_CATEGORICAL, np.where(ph["ycat"] == 1, 2, ph["ycat"]).astype(float), ph["Cova"])   # empty level
    with pytest.raises(SpagxeError):
        gpu_ctx.set_wald(_lib.TRAIT_CATEGORICAL, ph["ycat"] + 0.5, ph["Cova"])                  # not an index
    with pytest.raises(SpagxeError):
        gpu_ctx.set_wald(_lib.TRAIT_CATEGORICAL, ph["ycat"].astype(float), np.zeros((n, 3)))    # > 2 covariate columns
