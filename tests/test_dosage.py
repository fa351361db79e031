"""Imputed (non-integer) dosages in SPAGxE_CCT, SPAGxEmix_CCT and SPAGxEmix_CCT_Plus (ref R/SPAGxECCT.R:543-683: "both hard-called and imputed genotype data are
supported"): a double matrix with values outside {0, 1, 2, NA} leaves the 2-bit pipeline and is tested by the dense
per-SNP kernel k_cct_dosage_snp.  The oracle evaluates the same R lines on the real-valued vectors."""
import numpy as np
import pytest

from spagxecct_b200 import _lib
from tests._data import assert_parity, make_geno, make_pheno, routes_from_cct_output

gpu = pytest.mark.gpu


def _dosages(n, m, seed, Y=None):
    """best-guess dosages of an imputation: hard calls blurred towards the allele-frequency mean, a few NA"""
    rng = np.random.default_rng(seed)
    mafs = np.r_[np.full(2, 0.0003), np.linspace(0.005, 0.5, m - 10), np.full(8, 0.25)]
    effect = np.zeros(m); effect[-8:] = np.linspace(0.3, 1.0, 8)
    G = make_geno(n, mafs, seed, miss_rates=np.where(np.arange(m) % 3 == 0, 0.02, 0.0), count_major=np.arange(m) % 4 == 1,
                  effect=effect, Y=Y)
    info = rng.uniform(0.6, 1.0, m)                      # imputation quality per SNP
    mean = np.nanmean(G, axis=0)
    D = info * G + (1 - info) * mean + 0.0 * rng.standard_normal((n, m))
    D = np.clip(D + (1 - info) * 0.3 * rng.standard_normal((n, m)), 0, 2)
    D[np.isnan(G)] = np.nan
    D[:, 5] = G[:, 5]                                    # one hard-called column inside a dosage matrix
    D[:, 0] = 0.5 * np.nan_to_num(G[:, 0])               # rare and badly imputed: below min.maf
    return np.asfortranarray(D)


@gpu
@pytest.mark.parametrize("trait", ["binary", "quantitative"])
def test_cct_dosage_matrix(gpu_ctx, oracle, trait):
    n, m = 5003, 48
    ph = make_pheno(n, 41, trait=trait)
    D = _dosages(n, m, 42, Y=ph["Y"] if trait == "binary" else (ph["Y"] > np.median(ph["Y"])).astype(float))
    gpu_ctx.set_vectors(ph["R"], ph["E"])
    gpu_ctx.set_wald(_lib.TRAIT_BINARY if trait == "binary" else _lib.TRAIT_QUANTITATIVE, ph["Y"], ph["Cova"])
    geno = gpu_ctx.geno_desc(_lib.GENO_F64, D.ctypes.data, n, m, n)
    for eps in (0.001, 0.3):
        out, diag = gpu_ctx.run_cct(geno, dict(epsilon=eps))
        ref, route, iters, rc = oracle.cct_matrix(trait, D, ph["R"], ph["E"], ph["Y"], ph["Cova"], epsilon=eps, nthreads=8)
        assert rc == 0
        # MAF is a mean of doubles here (no integer allele count): equal to rounding, not bit for bit
        assert np.allclose(out[:, 0], ref[:, 0], rtol=1e-13, atol=0) and np.array_equal(out[:, 1], ref[:, 1])
        assert_parity(out, ref, [2, 3, 4, 5, 6], [7, 8, 9], exact_cols=(), label=f"dosage {trait} eps={eps}")
        f, b, s = routes_from_cct_output(out, eps)
        assert np.array_equal(f, (route & 1) > 0) and np.array_equal(b, (route & 2) > 0) and np.array_equal(s, (route & 4) > 0)
        assert diag["newton_iters"] == iters.sum() and diag["n_wald_fail"] == 0
    assert b.sum() >= 6 and s.sum() >= 1 and f.sum() >= 1


@gpu
def test_dosage_api_gmodel_and_limits(gpu_ctx, oracle):
    from spagxecct_b200 import SPAGxE_CCT, SPAGxE_Plus
    from spagxecct_b200._lib import SpagxeError
    n, m = 3001, 20
    ph = make_pheno(n, 43)
    D = _dosages(n, m, 44, Y=ph["Y"])
    ids = [f"id{i}" for i in range(n)]
    for gm, code in (("Add", 0), ("Dom", 1), ("Rec", 2)):
        res = SPAGxE_CCT("binary", Geno_mtx=(D, ids, None), R=ph["R"], E=ph["E"], Phen_mtx={"ID": ids, "Y": ph["Y"]},
                         Cova_mtx=ph["Cova"], epsilon=0.2, G_model=gm, ctx=gpu_ctx, quiet=True)
        oracle.set_gmodel(code)
        try:
            ref, route, _, rc = oracle.cct_matrix("binary", D, ph["R"], ph["E"], ph["Y"], ph["Cova"], epsilon=0.2)
        finally:
            oracle.set_gmodel(0)
        if rc != 0:      # "Rec" on near-zero dosages: solve() / CCT() stop in the reference
            continue
        assert np.allclose(res.values[:, 0], ref[:, 0], rtol=1e-13, atol=0)
        assert_parity(res.values, ref, [2, 3, 4, 5, 6], [7, 8, 9], exact_cols=(), label=f"dosage api {gm}")
    # the survival / categorical Wald kernels rest on the 2-bit coding and refuse dosages loudly
    gpu_ctx.set_vectors(ph["R"], ph["E"])
    gpu_ctx.set_wald(_lib.TRAIT_CATEGORICAL, (ph["Y"] > 0).astype(float), ph["Cova"])
    gpu_ctx.set_pcs(np.random.default_rng(1).standard_normal((n, 2)))
    with pytest.raises(SpagxeError) as ei:
        gpu_ctx.run_mix(gpu_ctx.geno_desc(_lib.GENO_F64, D.ctypes.data, n, m, n))
    assert ei.value.code == _lib.ERR_UNSUPPORTED


def _mix_dosages(golden_dir, seed):
    """the admixed genotypes of the SPAGxEmix tests blurred into dosages (sib pairs share alleles for the GRM test)"""
    from tests.test_mix_plus import _gpu_inputs
    x = _gpu_inputs(golden_dir)
    rng = np.random.default_rng(seed)
    G = x["G"]
    n, m = G.shape
    info = rng.uniform(0.7, 1.0, m)
    mean = np.nanmean(G, axis=0)
    D = np.clip(info * G + (1 - info) * mean + (1 - info) * 0.3 * rng.standard_normal((n, m)), 0, 2)
    for j in range(0, m, 2):                             # well-imputed columns: hard calls but for a few dozen samples
        D[:, j] = G[:, j]
        idx = rng.choice(n, 40, replace=False)
        D[idx, j] = np.clip(D[idx, j] + rng.uniform(-0.3, 0.3, 40), 0, 2)
    D[np.isnan(G)] = np.nan
    D[:, 7] = G[:, 7]                                    # one hard-called column inside the dosage matrix
    x["D"] = np.asfortranarray(D)
    return x


@gpu
@pytest.mark.parametrize("trait", ["binary", "quantitative"])
def test_mix_dosage_matrix(gpu_ctx, oracle, golden_dir, trait):
    """SPAGxEmix_CCT on imputed dosages (ref R/SPAGxECCT.R:1013-1269 on real-valued g): k_mix_dosage_snp vs the oracle."""
    from spagxecct_b200 import SPAGxEmix_CCT
    from tests.test_gpu_parity import _mix_inputs
    x = _mix_dosages(golden_dir, 51)
    y = _mix_inputs(golden_dir, trait)                   # phenotype / residuals of that trait on the same cohort
    n, m = x["D"].shape
    nb = ns = nl = 0
    for eps, cut in ((0.001, 2.0), (0.2, 1.0)):
        res = SPAGxEmix_CCT(trait, Geno_mtx=x["D"], R=y["R"], E=y["E"], Phen_mtx={"Y": y["Y"]}, Cova_mtx=y["Cova"], topPCs=y["PCs"],
                            epsilon=eps, Cutoff=cut, min_maf=0.0002, ctx=gpu_ctx, quiet=True)
        ref = np.empty((m, 12)); route = np.zeros(m, dtype=int); iters = 0
        for j in range(m):
            ref[j], route[j], it, _, rc = oracle.mix_one_snp(trait, x["D"][:, j], y["R"], y["E"], y["Y"], y["Cova"], y["PCs"],
                                                             epsilon=eps, cutoff=cut, min_maf=0.0002)
            assert rc == 0
            iters += sum(it)
        assert np.allclose(res.values[:, 0], ref[:, 0], rtol=1e-13, atol=0) and np.array_equal(res.values[:, 1], ref[:, 1])
        assert_parity(res.values, ref, [2, 3, 4, 5, 6], [7, 8, 9], exact_cols=(10,), label=f"mix dosage {trait} eps={eps}")
        assert np.allclose(res.values[:, 11], ref[:, 11], rtol=1e-13, equal_nan=True)
        d = res.diag
        assert d["n_branch_b"] == ((route & 2) > 0).sum() and d["n_spa"] == ((route & 4) > 0).sum()
        assert d["n_mix_logistic"] == ((route & 256) > 0).sum() and d["newton_iters"] == iters
        nb += ((route & 2) > 0).sum(); ns += ((route & 4) > 0).sum(); nl += ((route & 256) > 0).sum()
    assert nb >= 4 and ns >= 5 and nl >= 1


@gpu
def test_mixplus_dosage_matrix(gpu_ctx, oracle, golden_dir):
    x = _mix_dosages(golden_dir, 52)
    n, m = x["D"].shape
    gi, gj, gv = x["grm"]
    gpu_ctx.set_vectors(x["R"], x["E"]); gpu_ctx.set_pcs(x["PCs"]); gpu_ctx.set_sparse_grm(gi, gj, gv)
    geno = gpu_ctx.geno_desc(_lib.GENO_F64, x["D"].ctypes.data, n, m, n)
    nb = 0
    for eps, cut in ((0.001, 2.0), (0.2, 1.0)):
        out, diag = gpu_ctx.run_mix_plus(geno, dict(epsilon=eps, cutoff=cut, min_maf=0.0002))
        ref = np.empty((m, 12)); route = np.zeros(m, dtype=int); iters = 0
        for j in range(m):
            ref[j], route[j], it = oracle.mixplus_one_snp(x["D"][:, j], x["R"], x["E"], x["PCs"], gi, gj, gv, epsilon=eps, cutoff=cut, min_maf=0.0002)
            iters += sum(it)
        assert np.allclose(out[:, 0], ref[:, 0], rtol=1e-13, atol=0) and np.array_equal(out[:, 1], ref[:, 1])
        assert_parity(out, ref, [2, 5, 6], [7, 8, 9], exact_cols=(10,), label=f"mix_plus dosage eps={eps}")
        assert diag["n_branch_b"] == ((route & 2) > 0).sum() and diag["newton_iters"] == iters
        nb += ((route & 2) > 0).sum()
    assert nb >= 5


@gpu
def test_plus_dosage_matrix(gpu_ctx, oracle):
    """SPAGxE_Plus on imputed dosages (ref R/SPAGxEPlus_functions_MYZ.R:248-405 on real-valued g): k_plus_dosage_snp vs the oracle."""
    from spagxecct_b200 import SPAGxE_Plus
    n_fam = 1501
    n = 2 * n_fam + 1001
    ph = make_pheno(n, 61, prevalence=0.05)
    m = 40
    D = _dosages(n, m, 62, Y=ph["Y"])
    gi = np.r_[np.arange(n), np.arange(1, 2 * n_fam, 2)].astype(np.int32)
    gj = np.r_[np.arange(n), np.arange(0, 2 * n_fam, 2)].astype(np.int32)
    gv = np.r_[np.ones(n), np.full(n_fam, 0.5)]
    a, b, Rnew, c = gpu_ctx.plus_nullmodel(gi, gj, gv, ph["R"], ph["E"])
    obj = {"R": ph["R"], "R_GRM_R": a, "R_GRM_RE": b, "R.new": Rnew, "R.new_GRM_R.new": c, "sparseGRM": {"i": gi, "j": gj, "Value": gv}}
    nb = ns = 0
    for eps, cut in ((0.001, 2.0), (0.3, 1.0)):
        res = SPAGxE_Plus(Geno_mtx=D, E=ph["E"], Phen_mtx={"ID": np.arange(n)}, Cova_mtx=ph["Cova"], obj_SPAGxE_Plus_Nullmodel=obj,
                          epsilon=eps, Cutoff=cut, ctx=gpu_ctx, quiet=True)
        ref = np.empty((m, 8)); route = np.zeros(m, dtype=int); iters = 0
        for j in range(m):
            ref[j], route[j], it = oracle.plus_one_snp(D[:, j], ph["R"], ph["E"], Rnew, a, b, c, gi, gj, gv, epsilon=eps, cutoff=cut)
            iters += sum(it)
        assert np.allclose(res.values[:, 0], ref[:, 0], rtol=1e-13, atol=0) and np.array_equal(res.values[:, 1], ref[:, 1])
        assert_parity(res.values, ref, [2, 3, 4], [5, 6, 7], exact_cols=(), label=f"plus dosage eps={eps}")
        d = res.diag
        assert d["n_branch_b"] == ((route & 2) > 0).sum() and d["n_spa"] == ((route & 4) > 0).sum() and d["newton_iters"] == iters
        nb += ((route & 2) > 0).sum(); ns += ((route & 4) > 0).sum()
    assert nb >= 5 and ns >= 5
