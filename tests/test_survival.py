"""traits = "survival": the Wald test of branch B is coxph(Surv(surv.time, event) ~ Cova + E + g + g*E, iter.max = 1000)
(ref R/SPAGxECCT.R:622-634).  `survival` is a third-party package that is not in the reference tree, so the oracle
restates its published fitter (coxph.fit / coxfit6.c, Efron ties).  CPU part: that restatement against an independent
numpy Newton solver of the Efron partial likelihood.  GPU part (-m gpu): k_cox_wald against the oracle."""
import numpy as np
import pytest

from spagxecct_b200 import _lib
from tests._data import assert_parity, make_geno, routes_from_cct_output, to_bed_rows


def _efron(beta, X, time, event):
    """log partial likelihood, score and information with Efron's approximation, written from the textbook formula."""
    eta = X @ beta
    w = np.exp(eta)
    ll = 0.0; u = np.zeros(X.shape[1]); info = np.zeros((X.shape[1],) * 2)
    for t in np.unique(time[event == 1]):
        risk = time >= t
        dead = (time == t) & (event == 1)
        d = int(dead.sum())
        s0, s1, s2 = w[risk].sum(), (w[risk, None] * X[risk]).sum(0), np.einsum("i,ij,ik->jk", w[risk], X[risk], X[risk])
        d0, d1, d2 = w[dead].sum(), (w[dead, None] * X[dead]).sum(0), np.einsum("i,ij,ik->jk", w[dead], X[dead], X[dead])
        ll += eta[dead].sum(); u += X[dead].sum(0)
        for k in range(d):
            f = k / d
            den = s0 - f * d0; a = s1 - f * d1; c = s2 - f * d2
            ll -= np.log(den); u -= a / den; info += c / den - np.outer(a, a) / den ** 2
    return ll, u, info


def _surv_pheno(n, seed, ties, p=2, event_rate=0.15):
    rng = np.random.default_rng(seed)
    X1 = rng.standard_normal(n); X2 = rng.binomial(1, 0.5, n).astype(float); E = rng.standard_normal(n)
    eta = 0.4 * X1 + 0.3 * X2 + 0.4 * E
    t_event = rng.weibull(2.0, n) * np.exp(-eta / 2.0) * 3.0
    cens = np.quantile(t_event, event_rate) * rng.uniform(0.6, 1.4, n)
    time = np.minimum(t_event, cens); event = (t_event <= cens).astype(float)
    if ties:
        time = np.ceil(time * ties) / ties       # a few dozen distinct times: heavy ties among deaths and censorings
    cova = np.column_stack([X1, X2])[:, :p]
    # Step-1 residuals stay outside the path: any centred vector works as R here (martingale-like)
    R = event - event.mean() * np.exp(eta) / np.exp(eta).mean()
    return dict(time=time, event=event, E=E, Cova=cova, R=R)


@pytest.mark.parametrize("ties", [0, 40])
def test_oracle_coxph_matches_independent_newton(oracle, ties):
    n = 500
    ph = _surv_pheno(n, 3, ties, event_rate=0.4)
    rng = np.random.default_rng(4)
    g = rng.binomial(2, 0.3, n).astype(float)
    X = np.column_stack([ph["Cova"], ph["E"], g, ph["E"] * g])
    beta = np.zeros(X.shape[1])
    for _ in range(30):      # run Newton to machine precision
        ll, u, info = _efron(beta, X, ph["time"], ph["event"])
        beta = beta + np.linalg.solve(info, u)
    se = np.sqrt(np.diag(np.linalg.inv(_efron(beta, X, ph["time"], ph["event"])[2])))
    coef, se_o, iters, rc = oracle.coxph(X, ph["time"], ph["event"], maxiter=1000)
    assert rc == 0 and 3 <= iters <= 8
    assert np.allclose(coef, beta, rtol=1e-7, atol=1e-9)
    assert np.allclose(se_o, se, rtol=1e-7)


def test_oracle_survival_branch_b_is_coxph_on_the_design(oracle):
    n = 2003
    ph = _surv_pheno(n, 7, 25)
    G = make_geno(n, np.full(6, 0.25), 8, miss_rates=[0, 0.03, 0, 0.03, 0, 0], count_major=[0, 0, 1, 1, 0, 0])
    Y2 = np.r_[ph["time"], ph["event"]]
    out, route, iters, rc = oracle.cct_matrix("survival", G, ph["R"], ph["E"], Y2, ph["Cova"], epsilon=0.999)
    assert rc == 0 and ((route & 2) > 0).all()
    for j in range(G.shape[1]):
        g = G[:, j].copy(); maf = np.nanmean(g) / 2; g[np.isnan(g)] = 2 * maf
        if maf > 0.5:
            g = 2 - g
        X = np.column_stack([ph["Cova"], ph["E"], g, ph["E"] * g])
        coef, se, _, rc1 = oracle.coxph(X, ph["time"], ph["event"], maxiter=1000)
        assert rc1 == 0
        assert out[j, 3] == pytest.approx(2 * oracle.pnorm(-abs(coef[-1] / se[-1])), rel=1e-12)


gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("ties,p", [(0, 2), (30, 2), (6, 1), (0, 0)])
def test_cct_survival_wald(gpu_ctx, oracle, ties, p):
    n = 6007
    ph = _surv_pheno(n, 21 + ties, ties, p=p)
    mafs = np.r_[np.full(3, 0.0004), np.linspace(0.01, 0.5, 30), np.full(8, 0.2)]
    m = mafs.size
    rng = np.random.default_rng(5)
    miss = np.where(rng.random(m) < 0.5, 0.0, rng.random(m) * 0.06)
    major = rng.random(m) < 0.5
    effect = np.zeros(m); effect[-8:] = np.linspace(0.4, 1.2, 8)
    G = np.ascontiguousarray(make_geno(n, mafs, 12, miss_rates=miss, count_major=major, effect=effect, Y=ph["event"]))
    rows = to_bed_rows(G)
    gpu_ctx.set_vectors(ph["R"], ph["E"])
    gpu_ctx.set_wald_survival(ph["time"], ph["event"], ph["Cova"])
    geno = gpu_ctx.geno_desc(_lib.GENO_BED, rows.ctypes.data, n, m, (n + 3) // 4)
    Y2 = np.r_[ph["time"], ph["event"]]
    for eps in (0.001, 0.2):
        out, diag = gpu_ctx.run_cct(geno, dict(epsilon=eps))
        ref, route, iters, rc = oracle.cct_bed("survival", rows, n, ph["R"], ph["E"], Y2, ph["Cova"], epsilon=eps, nthreads=8)
        assert rc == 0
        assert_parity(out, ref, [2, 3, 4, 5, 6], [7, 8, 9], label=f"survival ties={ties} p={p} eps={eps}")
        f, b, s = routes_from_cct_output(out, eps)
        assert np.array_equal(b, (route & 2) > 0) and np.array_equal(s, (route & 4) > 0)
        assert diag["newton_iters"] == iters.sum() and diag["n_wald_fail"] == 0
    assert b.sum() >= 8
    # the same call again (cached time order) gives identical bits; new E invalidates the cache
    out2, _ = gpu_ctx.run_cct(geno, dict(epsilon=0.2))
    assert np.array_equal(np.nan_to_num(out2, nan=-1), np.nan_to_num(out, nan=-1))


@gpu
def test_survival_api_and_errors(gpu_ctx):
    from spagxecct_b200 import SPAGxE_CCT
    from spagxecct_b200._lib import SpagxeError
    n = 3001
    ph = _surv_pheno(n, 9, 20)
    G = make_geno(n, np.linspace(0.05, 0.4, 12), 10)
    ids = [f"id{i}" for i in range(n)]
    res = SPAGxE_CCT("survival", Geno_mtx=(G, ids, None), R=ph["R"], E=ph["E"],
                     Phen_mtx={"ID": ids, "surv.time": ph["time"], "event": ph["event"]}, Cova_mtx=ph["Cova"], epsilon=0.5,
                     ctx=gpu_ctx, quiet=True)
    assert res.diag["n_branch_b"] >= 3 and np.isfinite(res.values[:, 2:7]).all()
    with pytest.raises(SpagxeError):
        gpu_ctx.set_wald_survival(ph["time"], np.zeros(n), ph["Cova"])          # no events
    with pytest.raises(SpagxeError):
        gpu_ctx.set_wald_survival(ph["time"], ph["event"], np.zeros((n, 3)))    # > 2 covariate columns
    with pytest.raises(SpagxeError):
        gpu_ctx.set_wald(_lib.TRAIT_SURVIVAL, ph["event"], ph["Cova"])          # wrong entry point for this trait
