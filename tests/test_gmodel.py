"""G.model = "Dom" / "Rec" (ref R/SPAGxECCT.R:572-574, dup :1045-1047, R/SPAGxEPlus_functions_MYZ.R:285-287).

The recode happens after mean imputation and the MAF > 0.5 flip; MAF and missing.rate stay those of the additive
coding.  SPAGxE_CCT hands the recoded vector to SPA_G_one_SNP_homo_new WITHOUT G.model (:598, :616), so the inner
function re-derives its MAF from the 0/1 vector with its default "Add".
CPU part: the oracle's recode against a numpy composition of its own public pieces.  GPU part (-m gpu): the CUDA path
(second score pass for the hom-A1 class sums, class-value tables in every per-SNP kernel) against the oracle."""
import numpy as np
import pytest

from spagxecct_b200 import _lib
from tests._data import assert_parity, make_geno, make_pheno, routes_from_cct_output, to_bed_rows


def _recode(g, gm):
    """R semantics on one genotype vector with NaN = NA: returns (additive MAF, missing.rate, recoded g)."""
    g = g.copy()
    maf = np.nanmean(g) / 2
    miss = np.isnan(g).mean()
    g[np.isnan(g)] = 2 * maf
    if maf > 0.5:
        maf = 1 - maf; g = 2 - g
    if gm == 1:
        g = np.where(g >= 1, 1.0, 0.0)
    if gm == 2:
        g = np.where(g <= 1, 0.0, 1.0)
    return maf, miss, g


@pytest.fixture
def gmodel(oracle):
    yield oracle.set_gmodel
    oracle.set_gmodel(0)


@pytest.mark.parametrize("gm", [1, 2])
def test_oracle_recode_matches_numpy_composition(oracle, gmodel, gm):
    n = 3001
    ph = make_pheno(n, 5, prevalence=0.1)
    mafs = np.r_[np.linspace(0.02, 0.5, 12), 0.3, 0.4]
    G = make_geno(n, mafs, 6, miss_rates=np.r_[np.zeros(7), np.full(7, 0.04)], count_major=np.arange(14) % 2 == 1)
    gmodel(gm)
    out, route, iters, rc = oracle.cct_matrix("binary", G, ph["R"], ph["E"], ph["Y"], ph["Cova"], epsilon=0.001)
    gmodel(0)
    assert rc == 0
    R, E = ph["R"], ph["E"]
    lam = (E * R * R).sum() / (R * R).sum()
    Rn = (E - lam) * R
    checked = 0
    for j in range(G.shape[1]):
        maf, miss, g = _recode(G[:, j], gm)
        assert out[j, 0] == pytest.approx(maf, rel=1e-15) and out[j, 1] == miss
        S1 = (g * R).sum()
        assert out[j, 7] == pytest.approx(S1, rel=1e-11, abs=1e-12)
        assert out[j, 8] == pytest.approx((R * R).sum() * 2 * maf * (1 - maf), rel=1e-12)
        if route[j] & 2:
            continue
        # branch A: the inner call sees the recoded vector with G.model = "Add" and min.maf = 1e-5
        inner = oracle.spa_homo(g, Rn, cutoff=2.0, min_maf=0.00001)[0]
        assert np.allclose(out[j, 2:5], inner[2], rtol=1e-12, equal_nan=True)
        assert np.allclose(out[j, 5], inner[3], rtol=1e-12, equal_nan=True)
        checked += 1
    assert checked >= 8


pytest_gpu = pytest.mark.gpu


@pytest_gpu
@pytest.mark.parametrize("gm", [1, 2])
@pytest.mark.parametrize("trait", ["binary", "quantitative"])
def test_cct_dom_rec(gpu_ctx, oracle, gmodel, gm, trait):
    n = 6007
    ph = make_pheno(n, 11, prevalence=0.1, trait=trait)
    mafs = np.r_[np.full(4, 0.0004), [0.002, 0.008], np.linspace(0.01, 0.5, 40), np.full(10, 0.15), [0.5, 0.5]]
    m = mafs.size
    rng = np.random.default_rng(5)
    miss = np.where(rng.random(m) < 0.5, 0.0, rng.random(m) * 0.08)
    major = rng.random(m) < 0.5
    effect = np.zeros(m); effect[-12:-2] = np.linspace(0.3, 1.2, 10)
    G = np.ascontiguousarray(make_geno(n, mafs, 12, miss_rates=miss, count_major=major, effect=effect, Y=ph["Y"]))
    # exact MAF_raw == 0.5 with missing calls: the imputed value 2 * MAF = 1 is recoded to 1 by "Dom" (g >= 1)
    G[:, -1] = 1.0; G[:40, -1] = np.nan
    rows = to_bed_rows(G)
    gpu_ctx.set_vectors(ph["R"], ph["E"])
    gpu_ctx.set_wald(_lib.TRAIT_BINARY if trait == "binary" else _lib.TRAIT_QUANTITATIVE, ph["Y"], ph["Cova"])
    geno = gpu_ctx.geno_desc(_lib.GENO_BED, rows.ctypes.data, n, m, (n + 3) // 4)
    for eps in (0.001, 0.05):
        out, diag = gpu_ctx.run_cct(geno, dict(epsilon=eps, g_model=gm))
        gmodel(gm)
        ref, route, iters, rc = oracle.cct_bed(trait, rows, n, ph["R"], ph["E"], ph["Y"], ph["Cova"], epsilon=eps)
        gmodel(0)
        assert rc == 0
        assert_parity(out, ref, [2, 3, 4, 5, 6], [7, 8, 9], label=f"CCT G.model={gm} {trait} eps={eps}")
        f, b, s = routes_from_cct_output(out, eps)
        assert np.array_equal(f, (route & 1) > 0) and np.array_equal(b, (route & 2) > 0) and np.array_equal(s, (route & 4) > 0)
        assert diag["newton_iters"] == iters.sum()
    assert b.sum() >= 5 and f.sum() >= 3
    # the additive run differs (the recode is really applied)
    out0, _ = gpu_ctx.run_cct(geno, dict(epsilon=0.05))
    assert not np.allclose(np.nan_to_num(out0[:, 7]), np.nan_to_num(out[:, 7]))
    assert np.array_equal(np.nan_to_num(out0[:, :2], nan=-1), np.nan_to_num(out[:, :2], nan=-1))   # MAF, missing.rate: additive


@pytest_gpu
@pytest.mark.parametrize("gm", [1, 2])
def test_plus_dom_rec(gpu_ctx, oracle, gmodel, gm):
    n_fam = 2001
    n = 2 * n_fam + 1001
    rng = np.random.default_rng(71)
    ph = make_pheno(n, 72, prevalence=0.05)
    mafs = rng.uniform(0.02, 0.5, 48)
    G = make_geno(n, mafs, 73, miss_rates=np.where(rng.random(48) < 0.5, 0, 0.03), count_major=rng.random(48) < 0.5)
    G[1:2 * n_fam:2] = np.where(rng.random((n_fam, 48)) < 0.5, G[0:2 * n_fam:2], G[1:2 * n_fam:2])
    gi = np.r_[np.arange(n), np.arange(1, 2 * n_fam, 2)].astype(np.int32)
    gj = np.r_[np.arange(n), np.arange(0, 2 * n_fam, 2)].astype(np.int32)
    gv = np.r_[np.ones(n), np.full(n_fam, 0.5)]
    a, b, Rnew, c = gpu_ctx.plus_nullmodel(gi, gj, gv, ph["R"], ph["E"])
    rows = to_bed_rows(G)
    gpu_ctx.set_vectors(ph["R"], ph["E"])
    gpu_ctx.set_grm(gi, gj, gv, a, b, c, Rnew)
    geno = gpu_ctx.geno_desc(_lib.GENO_BED, rows.ctypes.data, n, 48, (n + 3) // 4)
    for eps in (0.001, 0.97):   # "Rec" statistics are small against the additive variance: a high epsilon reaches branch B
        out, diag = gpu_ctx.run_plus(geno, dict(epsilon=eps, cutoff=1.0, g_model=gm))
        gmodel(gm)
        ref, route, iters, _ = oracle.plus_bed(rows, n, ph["R"], ph["E"], Rnew, a, b, c, gi, gj, gv, epsilon=eps, cutoff=1.0)
        gmodel(0)
        assert_parity(out, ref, [2, 3, 4], [5, 6, 7], label=f"plus G.model={gm} eps={eps}")
        assert diag["newton_iters"] == iters.sum() and diag["n_branch_b"] == ((route & 2) > 0).sum()
    assert ((route & 2) > 0).sum() >= 4


@pytest_gpu
@pytest.mark.parametrize("gm", [1, 2])
def test_mix_dom_rec(gpu_ctx, oracle, gmodel, golden_dir, gm):
    from tests.test_gpu_parity import _mix_inputs
    from spagxecct_b200._lib import SpagxeError
    x = _mix_inputs(golden_dir, "binary")
    n = x["n"]
    G = x["G"]
    gpu_ctx.set_vectors(x["R"], x["E"])
    gpu_ctx.set_wald(_lib.TRAIT_BINARY, x["Y"], x["Cova"])
    gpu_ctx.set_pcs(x["PCs"])

    def both(G, eps):
        rows = to_bed_rows(G)
        gmodel(gm)
        ref, route, iters, rc = oracle.mix_bed("binary", rows, n, x["R"], x["E"], x["Y"], x["Cova"], x["PCs"], epsilon=eps,
                                               min_maf=0.0002, nthreads=8)
        gmodel(0)
        geno = gpu_ctx.geno_desc(_lib.GENO_BED, rows.ctypes.data, n, G.shape[1], (n + 3) // 4)
        return ref, route, iters, rc, (lambda keep=rows: gpu_ctx.run_mix(geno, dict(epsilon=eps, min_maf=0.0002, g_model=gm)))   # keep: the packed rows must outlive the descriptor

    for eps in (0.001, 0.2):
        ref, route, iters, rc, run = both(G, eps)
        singular = (route & 32) > 0
        if singular.any():
            # a "Rec" vector that is all zero makes W'W singular in branch B: the reference's solve() stops the whole
            # call (ref :1153), the library reports SPAGXE_ERR_SINGULAR; compare on the remaining SNPs
            with pytest.raises(SpagxeError) as ei:
                run()
            assert ei.value.code == _lib.ERR_SINGULAR and f"{int(singular.sum())} SNP" in str(ei.value)
            ref, route, iters, rc, run = both(np.ascontiguousarray(G[:, ~singular]), eps)
            assert not ((route & 32) > 0).any()
        out, diag = run()
        assert rc == 0
        assert_parity(out, ref, [2, 3, 4, 5, 6], [7, 8, 9], exact_cols=(0, 1, 10), label=f"mix G.model={gm} eps={eps}")
        assert diag["n_branch_b"] == ((route & 2) > 0).sum() and diag["n_spa"] == ((route & 4) > 0).sum()
        assert diag["n_mix_logistic"] == ((route & 256) > 0).sum() and diag["newton_iters"] == iters.sum()
