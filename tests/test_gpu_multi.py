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
