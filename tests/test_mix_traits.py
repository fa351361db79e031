"""SPAGxEmix_CCT with traits = "survival" / "categorical" (ref R/SPAGxECCT.R:1187-1193, :1244-1250): branch B's Wald test is
coxph / clm on the same design as in SPAGxE_CCT, with SPAGxEmix's extra rule `is.na(pval.wald) -> pval.spaGxE`.  The
per-SNP SPAGxEmix kernel hands those SNPs to k_cox_wald / k_clm_wald; checked against the oracle's SPAGxEmix restatement."""
import numpy as np
import pytest

from spagxecct_b200 import SPAGxEmix_CCT, _lib
from tests._data import assert_parity, to_bed_rows
from tests.test_categorical import _cat_pheno
from tests.test_survival import _surv_pheno

pytestmark = pytest.mark.gpu


def _geno(golden_dir, Yeff, seed):
    from tests.test_gpu_parity import _mix_inputs
    x = _mix_inputs(golden_dir, "binary")
    G = x["G"]
    rng = np.random.default_rng(seed)
    hi = Yeff > np.quantile(Yeff, 0.8)
    for j in range(8):       # a few SNPs associated with the outcome so that branch B is populated at epsilon = 0.001 too
        p = np.where(hi, 0.32, 0.2)
        G[:, j] = rng.binomial(2, p)
    return x


def _check(res, ref, route, iters, label):
    assert_parity(res.values, ref, [2, 3, 4, 5, 6], [7, 8, 9], exact_cols=(0, 1, 10), label=label)
    d = res.diag
    assert d["n_branch_b"] == ((route & 2) > 0).sum() and d["n_spa"] == ((route & 4) > 0).sum()
    assert d["n_mix_logistic"] == ((route & 256) > 0).sum() and d["newton_iters"] == iters.sum()


@pytest.mark.parametrize("per_snp", [0, 1])
def test_mix_survival(gpu_ctx, oracle, golden_dir, per_snp):
    n = 10_000
    ph = _surv_pheno(n, 31, 25)
    x = _geno(golden_dir, ph["event"] - ph["time"] / ph["time"].max(), 5)
    assert x["n"] == n
    rows = to_bed_rows(x["G"])
    Y2 = np.r_[ph["time"], ph["event"]]
    gpu_ctx.set_option(_lib.OPT_MIX_PER_SNP, per_snp)
    try:
        nb = 0
        for eps in (0.001, 0.2):
            res = SPAGxEmix_CCT("survival", Geno_mtx=x["G"], R=ph["R"], E=ph["E"], Phen_mtx={"surv.time": ph["time"], "event": ph["event"]},
                                Cova_mtx=ph["Cova"], topPCs=x["PCs"], epsilon=eps, min_maf=0.0002, ctx=gpu_ctx, quiet=True)
            ref, route, iters, rc = oracle.mix_bed("survival", rows, n, ph["R"], ph["E"], Y2, ph["Cova"], x["PCs"], epsilon=eps,
                                                   min_maf=0.0002, nthreads=8)
            assert rc == 0
            _check(res, ref, route, iters, f"mix survival eps={eps}")
            assert res.diag["cox_evals"] > 0 or not (route & 2).any()
            nb += ((route & 2) > 0).sum()
        assert nb >= 8
    finally:
        gpu_ctx.set_option(_lib.OPT_MIX_PER_SNP, 0)


def test_mix_categorical(gpu_ctx, oracle, golden_dir):
    n = 10_000
    ph = _cat_pheno(n, 33, 4)
    x = _geno(golden_dir, ph["ycat"].astype(float), 6)
    rows = to_bed_rows(x["G"])
    ylab = ph["levels"][ph["ycat"]]
    nb = 0
    for eps in (0.001, 0.2):
        res = SPAGxEmix_CCT("categorical", Geno_mtx=x["G"], R=ph["R"], E=ph["E"], Phen_mtx={"Y": ylab}, Cova_mtx=ph["Cova"],
                            topPCs=x["PCs"], epsilon=eps, min_maf=0.0002, ctx=gpu_ctx, quiet=True)
        ref, route, iters, rc = oracle.mix_bed("categorical", rows, n, ph["R"], ph["E"], ph["ycat"].astype(float), ph["Cova"], x["PCs"],
                                               epsilon=eps, min_maf=0.0002, nthreads=8)
        assert rc == 0
        _check(res, ref, route, iters, f"mix categorical eps={eps}")
        nb += ((route & 2) > 0).sum()
    assert nb >= 8 and res.diag["clm_evals"] > 0
