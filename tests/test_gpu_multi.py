"""Multi-GPU behind the C ABI (-m gpu; needs >= 2 visible GPUs, skipped otherwise): the SNP-sharded result of
spagxe_mgpu_run_cct must be bit-identical to the single-GPU matrix (contiguous SNP ranges, one NCCL
broadcast of [R | E | Y | Cova], disjoint row blocks of the caller's matrix)."""
import numpy as np
import pytest

from spagxecct_b200 import SPAGxE_CCT, _lib
from tests._data import make_geno, make_pheno, to_bed_rows

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")


@needs2
@pytest.mark.parametrize("n_gpus", [2, 3, 4, 8])
def test_sharded_matrix_equals_single_gpu_matrix(gpu_ctx, n_gpus):
    if _n_gpus() < n_gpus:
        pytest.skip(f"{n_gpus} GPUs not visible")
    n, m = 30_011, 1003                      # ragged: M not divisible by the GPU count, N not a multiple of 4
    ph = make_pheno(n, 301, prevalence=0.05)
    rng = np.random.default_rng(302)
    mafs = rng.uniform(0.002, 0.5, m)
    eff = np.where(rng.random(m) < 0.05, 0.8, 0.0)
    G = make_geno(n, mafs, 303, miss_rates=np.where(rng.random(m) < 0.5, 0, 0.03), count_major=rng.random(m) < 0.5,
                  effect=eff, Y=ph["Y"])
    rows = to_bed_rows(G)
    gpu_ctx.set_vectors(ph["R"], ph["E"]); gpu_ctx.set_wald(_lib.TRAIT_BINARY, ph["Y"], ph["Cova"])
    g = gpu_ctx.geno_desc(_lib.GENO_BED, rows.ctypes.data, n, m, (n + 3) // 4)
    one, d1 = gpu_ctx.run_cct(g, dict(epsilon=0.01))
    assert d1["n_branch_b"] >= 20 and d1["n_spa"] >= 20
    with _lib.MultiContext(n_gpus) as mg:
        mg.set_inputs(ph["R"], ph["E"], _lib.TRAIT_BINARY, ph["Y"], ph["Cova"])
        many, dm = mg.run_cct(g, dict(epsilon=0.01))
        # explicit per-GPU shards (unequal sizes) give the same matrix again
        cuts = np.linspace(0, m, n_gpus + 1).astype(int); cuts[1] = max(1, cuts[1] // 2)
        rb = (n + 3) // 4
        shards = [gpu_ctx.geno_desc(_lib.GENO_BED, rows.ctypes.data + int(cuts[k]) * rb, n, int(cuts[k + 1] - cuts[k]), rb)
                  for k in range(n_gpus)]
        many2, _ = mg.run_cct(shards, dict(epsilon=0.01))
    assert np.array_equal(np.nan_to_num(many, nan=-1), np.nan_to_num(one, nan=-1))
    assert np.array_equal(np.nan_to_num(many2, nan=-1), np.nan_to_num(one, nan=-1))
    for k in ("n_filtered", "n_branch_a", "n_branch_b", "n_spa", "newton_iters", "n_snps"):
        assert dm[k] == d1[k], k


@needs2
def test_api_n_gpus_argument(gpu_ctx):
    n, m = 8003, 64
    ph = make_pheno(n, 311)
    G = make_geno(n, np.linspace(0.01, 0.4, m), 312)
    ids = [f"id{i}" for i in range(n)]
    a = SPAGxE_CCT("binary", Geno_mtx=(G, ids, None), R=ph["R"], E=ph["E"], Phen_mtx={"ID": ids, "Y": ph["Y"]},
                   Cova_mtx=ph["Cova"], ctx=gpu_ctx, quiet=True)
    b = SPAGxE_CCT("binary", Geno_mtx=(G, ids, None), R=ph["R"], E=ph["E"], Phen_mtx={"ID": ids, "Y": ph["Y"]},
                   Cova_mtx=ph["Cova"], quiet=True, n_gpus=2)
    assert np.array_equal(np.nan_to_num(a.values, nan=-1), np.nan_to_num(b.values, nan=-1))


@needs2
def test_api_n_gpus_for_mix_plus_and_mixplus(gpu_ctx, golden_dir):
    """n_gpus of SPAGxEmix_CCT / SPAGxE_Plus / SPAGxEmix_CCT_Plus: the 2-GPU matrices equal the single-GPU ones bit for bit."""
    from spagxecct_b200 import SPAGxE_Plus, SPAGxEmix_CCT, SPAGxEmix_CCT_Plus
    from tests.test_mix_plus import _gpu_inputs
    x = _gpu_inputs(golden_dir)
    n = x["n"]
    gi, gj, gv = x["grm"]
    kw = dict(Geno_mtx=x["G"], R=x["R"], E=x["E"], Phen_mtx={"Y": x["Y"]}, Cova_mtx=x["Cova"], topPCs=x["PCs"], epsilon=0.2, min_maf=0.0002, quiet=True)
    one = SPAGxEmix_CCT("binary", ctx=gpu_ctx, **kw)
    two = SPAGxEmix_CCT("binary", n_gpus=2, **kw)
    assert np.array_equal(np.nan_to_num(two.values, nan=-1), np.nan_to_num(one.values, nan=-1))
    kw = dict(Geno_mtx=x["G"], ResidMat={"SubjID": np.arange(n), "Resid": x["R"]}, E=x["E"], Phen_mtx={"Y": x["Y"]}, Cova_mtx=x["Cova"],
              topPCs=x["PCs"], sparseGRM={"i": gi, "j": gj, "Value": gv}, epsilon=0.2, min_maf=0.0002, wald="skip", quiet=True)
    one = SPAGxEmix_CCT_Plus("binary", ctx=gpu_ctx, **kw)
    two = SPAGxEmix_CCT_Plus("binary", n_gpus=2, **kw)
    assert np.array_equal(np.nan_to_num(two.values, nan=-1), np.nan_to_num(one.values, nan=-1)) and one.diag["n_branch_b"] >= 3
    a, b, Rnew, c = gpu_ctx.plus_nullmodel(gi, gj, gv, x["R"], x["E"])
    obj = {"R": x["R"], "R_GRM_R": a, "R_GRM_RE": b, "R.new": Rnew, "R.new_GRM_R.new": c, "sparseGRM": {"i": gi, "j": gj, "Value": gv}}
    kw = dict(Geno_mtx=x["G"], E=x["E"], Phen_mtx={"ID": np.arange(n)}, Cova_mtx=x["Cova"], obj_SPAGxE_Plus_Nullmodel=obj, epsilon=0.2, quiet=True)
    one = SPAGxE_Plus(ctx=gpu_ctx, **kw)
    two = SPAGxE_Plus(n_gpus=2, **kw)
    assert np.array_equal(np.nan_to_num(two.values, nan=-1), np.nan_to_num(one.values, nan=-1))


def test_single_gpu_multicontext_is_the_plain_path(gpu_ctx):
    """n_gpus = 1 needs no NCCL and must equal the single-context result (runs on every box)."""
    n, m = 5003, 40
    ph = make_pheno(n, 321)
    G = make_geno(n, np.linspace(0.01, 0.4, m), 322, miss_rates=np.full(m, 0.01))
    rows = to_bed_rows(G)
    gpu_ctx.set_vectors(ph["R"], ph["E"]); gpu_ctx.set_wald(_lib.TRAIT_BINARY, ph["Y"], ph["Cova"])
    g = gpu_ctx.geno_desc(_lib.GENO_BED, rows.ctypes.data, n, m, (n + 3) // 4)
    one, _ = gpu_ctx.run_cct(g)
    with _lib.MultiContext(1) as mg:
        mg.set_inputs(ph["R"], ph["E"], _lib.TRAIT_BINARY, ph["Y"], ph["Cova"])
        many, _ = mg.run_cct(g)
    assert np.array_equal(np.nan_to_num(many, nan=-1), np.nan_to_num(one, nan=-1))


@needs2
def test_mix_plus_and_survival_shard_like_cct(gpu_ctx, golden_dir):
    """spagxe_mgpu_run_mix / _plus and the survival trait: the 2-GPU matrices equal the single-GPU ones bit for bit."""
    from tests.test_gpu_parity import _mix_inputs
    from tests.test_survival import _surv_pheno
    # SPAGxEmix_CCT
    x = _mix_inputs(golden_dir, "binary")
    rows = to_bed_rows(x["G"]); n = x["n"]; m = x["G"].shape[1]
    gpu_ctx.set_vectors(x["R"], x["E"]); gpu_ctx.set_wald(_lib.TRAIT_BINARY, x["Y"], x["Cova"]); gpu_ctx.set_pcs(x["PCs"])
    g = gpu_ctx.geno_desc(_lib.GENO_BED, rows.ctypes.data, n, m, (n + 3) // 4)
    one, d1 = gpu_ctx.run_mix(g, dict(epsilon=0.2, min_maf=0.0002))
    with _lib.MultiContext(2) as mg:
        mg.set_inputs(x["R"], x["E"], _lib.TRAIT_BINARY, x["Y"], x["Cova"]); mg.set_pcs(x["PCs"])
        two, d2 = mg.run_mix(g, dict(epsilon=0.2, min_maf=0.0002))
    assert np.array_equal(np.nan_to_num(two, nan=-1), np.nan_to_num(one, nan=-1))
    assert d2["n_mix_logistic"] == d1["n_mix_logistic"] and d2["newton_iters"] == d1["newton_iters"]
    # SPAGxEmix_CCT_Plus on the same data with a sib-pair GRM
    from tests.test_mix_plus import _family_grm
    gi, gj, gv = _family_grm(n, 3000)
    gpu_ctx.set_sparse_grm(gi, gj, gv)
    one_mp, dm1 = gpu_ctx.run_mix_plus(g, dict(epsilon=0.2, min_maf=0.0002))
    with _lib.MultiContext(2) as mg:
        mg.set_inputs(x["R"], x["E"]); mg.set_pcs(x["PCs"]); mg.set_sparse_grm(gi, gj, gv)
        two_mp, dm2 = mg.run_mix_plus(g, dict(epsilon=0.2, min_maf=0.0002))
    assert np.array_equal(np.nan_to_num(two_mp, nan=-1), np.nan_to_num(one_mp, nan=-1))
    assert dm2["n_branch_b"] == dm1["n_branch_b"] >= 3 and dm2["newton_iters"] == dm1["newton_iters"]
    # SPAGxE_Plus (sib pairs) and SPAGxE_CCT with a survival trait on the same cohort
    n_fam = 2001; n = 2 * n_fam + 1001
    ph = _surv_pheno(n, 41, 25)
    rng = np.random.default_rng(42)
    G = make_geno(n, rng.uniform(0.02, 0.5, 61), 43, miss_rates=np.where(rng.random(61) < 0.5, 0, 0.03))
    rows = to_bed_rows(G)
    gi = np.r_[np.arange(n), np.arange(1, 2 * n_fam, 2)].astype(np.int32)
    gj = np.r_[np.arange(n), np.arange(0, 2 * n_fam, 2)].astype(np.int32)
    gv = np.r_[np.ones(n), np.full(n_fam, 0.5)]
    a, b, Rnew, c = gpu_ctx.plus_nullmodel(gi, gj, gv, ph["R"], ph["E"])
    g = gpu_ctx.geno_desc(_lib.GENO_BED, rows.ctypes.data, n, 61, (n + 3) // 4)
    gpu_ctx.set_vectors(ph["R"], ph["E"]); gpu_ctx.set_grm(gi, gj, gv, a, b, c, Rnew)
    one_p, _ = gpu_ctx.run_plus(g, dict(epsilon=0.3, cutoff=1.0))
    gpu_ctx.set_wald_survival(ph["time"], ph["event"], ph["Cova"])
    one_s, ds = gpu_ctx.run_cct(g, dict(epsilon=0.3))
    assert ds["n_branch_b"] >= 5
    with _lib.MultiContext(2) as mg:
        mg.set_inputs(ph["R"], ph["E"]); mg.set_grm(gi, gj, gv, a, b, c, Rnew)
        two_p, _ = mg.run_plus(g, dict(epsilon=0.3, cutoff=1.0))
        mg.set_wald_survival(ph["time"], ph["event"], ph["Cova"])
        two_s, _ = mg.run_cct(g, dict(epsilon=0.3))
    assert np.array_equal(np.nan_to_num(two_p, nan=-1), np.nan_to_num(one_p, nan=-1))
    assert np.array_equal(np.nan_to_num(two_s, nan=-1), np.nan_to_num(one_s, nan=-1))
