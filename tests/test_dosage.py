"""Imputed (non-integer) dosages in SPAGxE_CCT (ref R/SPAGxECCT.R:543-683: "both hard-called and imputed genotype data are
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
    # the loops that rest on the 2-bit coding refuse dosages loudly
    gpu_ctx.set_vectors(ph["R"], ph["E"]); gpu_ctx.set_wald(_lib.TRAIT_BINARY, ph["Y"], ph["Cova"])
    gpu_ctx.set_pcs(np.random.default_rng(1).standard_normal((n, 2)))
    with pytest.raises(SpagxeError) as ei:
        gpu_ctx.run_mix(gpu_ctx.geno_desc(_lib.GENO_F64, D.ctypes.data, n, m, n))
    assert ei.value.code == _lib.ERR_UNSUPPORTED
