"""BGEN input of the GenoFile argument (GRAB.ReadGeno for `.bgen`, ref R/SPAGxECCT.R:484-487).  CPU: writer -> reader round
trips over bit widths / compression / missing calls, and -- in this container, where /root/reference exists -- the
reference's bundled inst/*.bgen decoded against its own plink .raw golden vectors (tests/golden/bed_*.npz), which pins
the allele convention (GRAB counts the second allele of a BGEN file by default).  GPU: SPAGxE_CCT(GenoFile = x.bgen)
equals SPAGxE_CCT on the matrix, for hard calls and for imputed dosages."""
import os

import numpy as np
import pytest

from spagxecct_b200.bgenio import BgenFile, write_bgen

REF_INST = "/root/reference/inst"


def _probs(n, m, seed, hard):
    rng = np.random.default_rng(seed)
    g = rng.binomial(2, rng.uniform(0.05, 0.5, m), (n, m))          # count of the SECOND allele
    P = np.zeros((n, m, 2))
    if hard:
        P[..., 0] = g == 0; P[..., 1] = g == 1
    else:
        w = rng.dirichlet([6, 1, 1], (n, m))                         # blurred towards the true genotype
        full = np.stack([g == 0, g == 1, g == 2], -1) * 0.7 + 0.3 * w
        full /= full.sum(-1, keepdims=True)
        P[..., 0] = full[..., 0]; P[..., 1] = full[..., 1]
    miss = rng.random((n, m)) < 0.01
    P[miss] = np.nan
    return P, miss


@pytest.mark.parametrize("bits,compress,hard", [(8, True, True), (8, False, False), (16, True, False), (5, True, False), (1, True, True)])
def test_bgen_round_trip(tmp_path, bits, compress, hard):
    n, m = 257, 9
    P, miss = _probs(n, m, bits, hard)
    ids = [f"S{i}" for i in range(n)]; snps = [f"rs{j}" for j in range(m)]
    path = str(tmp_path / "x.bgen")
    write_bgen(path, P, ids, snps, bits=bits, compress=compress)
    bf = BgenFile(path)
    assert (bf.n, bf.m, bf.layout, bf.sample_ids, bf.snp_ids) == (n, m, 2, ids, snps)
    D = bf.read_dosages()
    scale = (1 << bits) - 1
    q = np.rint(np.nan_to_num(P) * scale) / scale                   # what the file can hold
    want = q[..., 1] + 2 * (1 - q[..., 0] - q[..., 1])
    want[miss] = np.nan
    assert np.array_equal(np.isnan(D), miss)
    assert np.allclose(np.nan_to_num(D), np.nan_to_num(want), rtol=0, atol=1e-12)
    if hard and bits > 1:
        assert np.array_equal(np.nan_to_num(D, nan=-1), np.nan_to_num(np.where(miss, np.nan, np.rint(want)), nan=-1))
    D1 = bf.read_dosages(sel=[3, 1], allele_order="alt-first")
    assert np.allclose(np.nan_to_num(D1), np.nan_to_num(2 - want[:, [3, 1]]), atol=1e-12)


@pytest.mark.skipif(not os.path.isdir(REF_INST), reason="the reference tree is only present in the build container")
@pytest.mark.parametrize("name", ["SPAGxE", "SPAGxEmix", "SPAGxE_Plus"])
def test_bundled_bgen_equals_the_reference_raw_vectors(golden_dir, name):
    bf = BgenFile(f"{REF_INST}/GenoMat_{name}.bgen", f"{REF_INST}/GenoMat_{name}.sample")
    g = np.load(os.path.join(golden_dir, f"bed_{name}.npz"), allow_pickle=True)
    raw = g["raw"].astype(np.float64); raw[raw < 0] = np.nan
    D = bf.read_dosages()
    assert D.shape == raw.shape and bf.snp_ids == [str(s) for s in g["snp_ids"]]
    assert bf.sample_ids[: len(g["fam_first"])] == [str(s) for s in g["fam_first"]]
    assert np.array_equal(np.nan_to_num(D, nan=-9), np.nan_to_num(raw, nan=-9))


gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("hard", [True, False])
def test_cct_from_bgen_file(gpu_ctx, oracle, tmp_path, hard):
    from spagxecct_b200 import SPAGxE_CCT
    from tests._data import assert_parity, make_pheno
    n, m = 3001, 24
    ph = make_pheno(n, 51)
    P, miss = _probs(n, m, 52, hard)
    ids = [f"IID-{i + 1}" for i in range(n)]; snps = [f"SNP-{j + 1}" for j in range(m)]
    path = str(tmp_path / "g.bgen")
    write_bgen(path, P, ids, snps, bits=8)
    D = BgenFile(path).read_dosages()
    order = np.random.default_rng(3).permutation(n)[: n - 100]          # SampleIDs: a reordered subset
    sub = [ids[i] for i in order]
    res = SPAGxE_CCT("binary", GenoFile=path, R=ph["R"][order], E=ph["E"][order],
                     Phen_mtx={"ID": sub, "Y": ph["Y"][order]}, Cova_mtx=ph["Cova"][order], epsilon=0.2, ctx=gpu_ctx, quiet=True)
    ref, route, _, rc = oracle.cct_matrix("binary", D[order], ph["R"][order], ph["E"][order], ph["Y"][order], ph["Cova"][order],
                                          epsilon=0.2, nthreads=4)
    assert rc == 0 and res.rownames == snps
    assert np.allclose(res.values[:, 0], ref[:, 0], rtol=1e-13, atol=0)
    assert_parity(res.values, ref, [2, 3, 4, 5, 6], [7, 8, 9], exact_cols=(1,), label=f"bgen hard={hard}")
    assert (res.diag["n_branch_b"] >= 1) and res.diag["n_snps"] == m
