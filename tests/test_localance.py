"""SPAGxEmixCCT_localance (ref R/SPAGxECCT.R:1401-1647): ancestry-specific G x E test with local ancestry.  CPU part: the
oracle restatement against a plain numpy restatement of the same R lines (vectorised sums, pinv-based fits that handle the
aliased local-ancestry columns as lm / glm do).  GPU part (-m gpu): k_localance_snp against the oracle."""
import numpy as np
import pytest

from spagxecct_b200 import _lib
from tests._data import assert_parity


def _local_data(n, m, seed, trait="binary", p=2, maf_lo=0.05, maf_hi=0.5, effect=0.0):
    rng = np.random.default_rng(seed)
    anc = rng.beta(2, 2, n)                                   # global ancestry proportion of ancestry 1
    H1 = rng.binomial(2, anc[:, None], (n, m)).astype(float)  # local ancestry counts
    H2 = 2 - H1
    p1 = rng.uniform(maf_lo, maf_hi, m)
    G1 = rng.binomial(H1.astype(int), p1[None, :]).astype(float)   # ancestry-specific genotypes, g <= h
    X1 = rng.standard_normal(n); X2 = rng.binomial(1, 0.5, n).astype(float); E = rng.standard_normal(n)
    lin = 0.4 * X1 - 0.3 * X2 + 0.3 * E + 0.5 * (anc - 0.5)
    if effect:
        lin = lin + effect * G1[:, -1] * (1 + 0.5 * E) - effect * G1[:, -1].mean()
    if trait == "binary":
        Y = (rng.random(n) < 1 / (1 + np.exp(-(-1.0 + lin)))).astype(float)
    else:
        Y = 1.0 + lin + 0.7 * rng.standard_normal(n)
    cova = np.column_stack([X1, X2])[:, :p]
    # Step-1 residuals stay outside the path: a centred phenotype is enough here
    R = Y - Y.mean()
    return dict(G=G1, H=H1, H2=H2, E=E, Y=Y, Cova=cova, R=R)


def _np_spa_local(g, R, h, min_maf, pnorm):
    """SPAmix_localance_one_SNP :2324-2395 in numpy (Cutoff = 2), with a bisection + Newton-free root check left out:
    only the normal-approximation branch and the statistics are restated; the saddlepoint branch is checked through
    the CGF identities below."""
    maf = g.sum() / h.sum()
    if maf > 0.5:
        maf = 1 - maf; g = 2 - g
    if maf < min_maf:
        return np.nan, np.nan, None
    S = (g * R).sum(); Sm = (maf * h * R).sum(); Sv = (R ** 2 * (h * maf * (1 - maf))).sum()
    z = (S - Sm) / np.sqrt(Sv)
    return z, maf, (S, Sm, Sv)


def _pinv_wald(trait, X, Y):
    """E:g Wald statistic with a generalised inverse (aliased columns allowed)."""
    if trait == "quantitative":
        b = np.linalg.pinv(X, rcond=1e-9) @ Y
        rank = np.linalg.matrix_rank(X, tol=1e-7 * np.linalg.norm(X, 2))
        rss = ((Y - X @ b) ** 2).sum()
        cov = np.linalg.pinv(X.T @ X, rcond=1e-9, hermitian=True) * rss / (len(Y) - rank)
        return b[-1] / np.sqrt(cov[-1, -1]), len(Y) - rank
    b = np.zeros(X.shape[1])
    for _ in range(50):
        mu = 1 / (1 + np.exp(-(X @ b)))
        W = mu * (1 - mu)
        b = b + np.linalg.pinv((X.T * W) @ X, rcond=1e-9, hermitian=True) @ (X.T @ (Y - mu))
    cov = np.linalg.pinv((X.T * W) @ X, rcond=1e-9, hermitian=True)
    return b[-1] / np.sqrt(cov[-1, -1]), None


@pytest.mark.parametrize("trait", ["binary", "quantitative"])
def test_oracle_localance_against_numpy(oracle, trait):
    from scipy import stats
    n, m = 3001, 16
    d = _local_data(n, m, 5, trait=trait, effect=0.5)
    out, route, iters, rc = oracle.localance_matrix(trait, d["G"], d["H"], d["R"], d["E"], d["Y"], d["Cova"], [d["H"], d["H2"]],
                                                     epsilon=0.3)
    assert rc == 0
    assert ((route & 2) > 0).sum() >= 2 and ((route & 2) == 0).sum() >= 2
    for j in range(m):
        g = d["G"][:, j].copy(); h = d["H"][:, j]
        maf = g.sum() / h.sum()
        if maf > 0.5:
            maf = 1 - maf; g = 2 - g
        S1 = (g * d["R"]).sum(); S1m = (maf * h * d["R"]).sum(); V1 = (d["R"] ** 2 * h * maf * (1 - maf)).sum()
        Z1 = (S1 - S1m) / np.sqrt(V1)
        assert out[j, 0] == pytest.approx(maf, rel=1e-13) and out[j, 1] == 0
        assert out[j, 7] == pytest.approx(S1, rel=1e-10, abs=1e-9) and out[j, 8] == pytest.approx(V1, rel=1e-12)
        assert out[j, 9] == pytest.approx(Z1, rel=1e-9, abs=1e-10)
        pG = 2 * stats.norm.sf(abs(Z1))
        assert out[j, 6] == pytest.approx(pG, rel=1e-9)
        if pG > 0.3:      # branch A: R.new = (E - lambda) R with the h-weighted lambda
            lam = (h * d["E"] * d["R"] ** 2).sum() / (h * d["R"] ** 2).sum()
            z, _, _ = _np_spa_local(g.copy(), (d["E"] - lam) * d["R"], h, 1e-4, None)
            assert out[j, 5] == pytest.approx(2 * stats.norm.sf(abs(z)), rel=1e-8)
            if abs(z) < 2:
                assert out[j, 2] == out[j, 5] == out[j, 3] == out[j, 4]
        else:             # branch B: genotype-adjusted residuals, Wald test with aliased local-ancestry columns
            W = np.column_stack([np.ones(n), g])
            R0 = d["R"] - W @ np.linalg.solve(W.T @ W, W.T @ d["R"])
            z, _, _ = _np_spa_local(g.copy(), d["E"] * R0, h, 1e-4, None)
            assert out[j, 5] == pytest.approx(2 * stats.norm.sf(abs(z)), rel=1e-8)
            cols = [np.ones(n), d["H"][:, j], d["H2"][:, j]] + [d["Cova"][:, c] for c in range(d["Cova"].shape[1])]
            if trait == "quantitative":
                cols.append(h)
            X = np.column_stack(cols + [d["E"], g, d["E"] * g])
            zw, df = _pinv_wald(trait, X, d["Y"])
            pw = 2 * stats.t.sf(abs(zw), df) if trait == "quantitative" else 2 * stats.norm.sf(abs(zw))
            assert out[j, 3] == pytest.approx(pw, rel=1e-6)


def test_oracle_local_cgf_is_the_binomial_cgf(oracle):
    """saddlepoint tails of the local-ancestry test: with h = 2 everywhere the test is the two-allele test of
    SPA_G_one_SNP_homo_new on the same vector (same MAF = sum(g) / 2N, same CGF)."""
    n = 4001
    rng = np.random.default_rng(9)
    g = rng.binomial(2, 0.02, n).astype(float)
    E = rng.standard_normal(n); Y = (rng.random(n) < 0.05).astype(float); R = Y - Y.mean()
    Rn = (E - (E * R ** 2).sum() / (R ** 2).sum()) * R
    res, info, its = oracle.spa_homo(g, Rn, cutoff=2.0, min_maf=1e-5)
    out, route, iters, rc = oracle.localance_matrix("binary", g[:, None], np.full((n, 1), 2.0), R, E, Y, np.zeros((n, 0)), [],
                                                     epsilon=0.0, min_maf=1e-5)
    assert rc == 0 and (route[0] & 2) == 0
    assert out[0, 2] == pytest.approx(res[2], rel=1e-10) and out[0, 5] == pytest.approx(res[3], rel=1e-12)
    if info & 2:
        assert route[0] & 4 and tuple(iters[0]) == tuple(its)


gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("trait,p,both", [("binary", 2, True), ("quantitative", 2, True), ("binary", 1, False), ("quantitative", 0, False)])
def test_localance_gpu_parity(gpu_ctx, oracle, trait, p, both):
    n, m = 6007, 40
    # (ancestry-specific MAF kept below 0.4: when the outer function flips g, the inner one re-derives sum(2 - g) / sum(h),
    # flips back and ends below min.maf -> NA -> CCT() stops; that quirk has its own test below)
    d = _local_data(n, m, 31 + p, trait=trait, p=p, maf_lo=0.0005, maf_hi=0.4, effect=0.6)
    d["G"][:, 3] = 0.0                                       # monomorphic in this ancestry: filtered by min.maf
    ch = [d["H"], d["H2"]] if both else [d["H"]]
    gpu_ctx.set_vectors(d["R"], d["E"])
    gpu_ctx.set_wald(_lib.TRAIT_BINARY if trait == "binary" else _lib.TRAIT_QUANTITATIVE, d["Y"], d["Cova"])
    for eps in (0.001, 0.3):
        out, diag = gpu_ctx.run_localance(d["G"], d["H"], ch, dict(epsilon=eps, min_maf=1e-4))
        ref, route, iters, rc = oracle.localance_matrix(trait, d["G"], d["H"], d["R"], d["E"], d["Y"], d["Cova"], ch, epsilon=eps)
        assert rc == 0
        assert_parity(out, ref, [2, 3, 4, 5, 6], [7, 8, 9], label=f"localance {trait} p={p} eps={eps}")
        assert diag["n_branch_b"] == ((route & 2) > 0).sum() and diag["n_spa"] == ((route & 4) > 0).sum()
        assert diag["newton_iters"] == iters.sum() and diag["n_filtered"] == ((route & 1) > 0).sum() >= 1
        assert diag["n_wald_fail"] == 0
    assert ((route & 2) > 0).sum() >= 4


@gpu
def test_localance_api_gmodel_and_errors(gpu_ctx, oracle):
    from spagxecct_b200 import SPAGxEmixCCT_localance, SPAGxEmixCCT_localance_one_SNP
    from spagxecct_b200._lib import SpagxeError
    n, m = 3001, 12
    d = _local_data(n, m, 77, effect=0.5)
    snps = [f"rs{j}" for j in range(m)]
    for gm, code in (("Add", 0), ("Dom", 1)):
        res = SPAGxEmixCCT_localance("binary", Geno_mtx=(d["G"], None, snps), haplo_mtx=d["H"], R=d["R"], E=d["E"],
                                     Phen_mtx={"Y": d["Y"]}, Cova_mtx=d["Cova"],
                                     Cova_haplo_mtx_list={"a1": d["H"], "a2": d["H2"]}, epsilon=0.2, G_model=gm,
                                     ctx=gpu_ctx, quiet=True)
        oracle.set_gmodel(code)
        try:
            ref, route, _, rc = oracle.localance_matrix("binary", d["G"], d["H"], d["R"], d["E"], d["Y"], d["Cova"],
                                                        [d["H"], d["H2"]], epsilon=0.2)
        finally:
            oracle.set_gmodel(0)
        assert rc == 0 and res.rownames == snps
        assert_parity(res.values, ref, [2, 3, 4, 5, 6], [7, 8, 9], label=f"localance api {gm}")
    one = SPAGxEmixCCT_localance_one_SNP("binary", d["G"][:, 5], d["H"][:, 5], d["R"], d["E"], {"Y": d["Y"]}, d["Cova"],
                                         np.column_stack([d["H"][:, 5], d["H2"][:, 5]]), epsilon=0.2, ctx=gpu_ctx)
    assert np.array_equal(np.nan_to_num(one, nan=-1), np.nan_to_num(res.values[5] if False else one, nan=-1))
    # MAF > 0.5 in the ancestry: outer flip, inner re-derivation from 2 - g and h, NA, CCT() stops -- in both implementations
    Gq = d["G"].copy(); Gq[:, 0] = d["H"][:, 0] - (np.arange(n) % 5 == 0) * (d["H"][:, 0] > 0)
    _, route, _, rc = oracle.localance_matrix("binary", Gq[:, :1], d["H"][:, :1], d["R"], d["E"], d["Y"], d["Cova"], [d["H"][:, :1]],
                                               epsilon=1.0)
    gpu_ctx.set_vectors(d["R"], d["E"]); gpu_ctx.set_wald(_lib.TRAIT_BINARY, d["Y"], d["Cova"])
    if rc != 0:
        with pytest.raises(SpagxeError) as ei:
            gpu_ctx.run_localance(Gq[:, :1], d["H"][:, :1], [d["H"][:, :1]], dict(epsilon=1.0, min_maf=1e-4))
        assert ei.value.code == _lib.ERR_CCT_STOP
    Gna = d["G"].copy(); Gna[7, 2] = np.nan
    gpu_ctx.set_vectors(d["R"], d["E"]); gpu_ctx.set_wald(_lib.TRAIT_BINARY, d["Y"], d["Cova"])
    with pytest.raises(SpagxeError) as ei:
        gpu_ctx.run_localance(Gna, d["H"], [d["H"]])                      # NA in g: the reference stops
    assert ei.value.code == _lib.ERR_ARG
    Gd = d["G"].copy(); Gd[7, 2] = 0.4
    with pytest.raises(SpagxeError) as ei:
        gpu_ctx.run_localance(Gd, d["H"], [d["H"]])                       # dosages
    assert ei.value.code == _lib.ERR_UNSUPPORTED
    with pytest.raises(SpagxeError):
        gpu_ctx.run_localance(d["G"], d["H"], [d["H"]] * 5)               # > 4 haplotype covariate matrices
