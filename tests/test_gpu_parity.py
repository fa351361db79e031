"""GPU parity tests (-m gpu): the CUDA path, driven through the C ABI, against the CPU oracle on the
same seeded inputs, against the reference's golden decode vectors, and through size-independent
properties at larger sizes."""
import numpy as np
import pytest

from spagxecct_b200 import RStop, SPAGxE_CCT, SPAGxE_CCT_one_SNP, _lib
from spagxecct_b200.harness import get_resid
from tests._data import assert_parity, make_geno, make_pheno, routes_from_cct_output, to_bed_rows

pytestmark = pytest.mark.gpu

P_COLS = [2, 3, 4, 5, 6]
S_COLS = [7, 8, 9]


def _bed_desc(ctx, rows, n, pitch=None):
    pitch = pitch or (n + 3) // 4
    return ctx.geno_desc(_lib.GENO_BED, rows.ctypes.data, n, rows.size // pitch, pitch)


@pytest.mark.parametrize("name", ["SPAGxE", "SPAGxEmix", "SPAGxE_Plus"])
def test_decode_and_counts_bit_exact_vs_plink_raw(gpu_ctx, golden_dir, name):
    d = np.load(f"{golden_dir}/bed_{name}.npz")
    rows = np.ascontiguousarray(d["bed"][3:]); raw = d["raw"]; n = int(d["n"])
    g = _bed_desc(gpu_ctx, rows, n)
    dec = gpu_ctx.bed_decode(g)
    want = raw.astype(np.float64); want[raw < 0] = np.nan
    assert np.array_equal(np.isnan(dec), np.isnan(want))
    assert np.array_equal(np.nan_to_num(dec, nan=-1), np.nan_to_num(want, nan=-1))
    # NA_real_ payload (1954) like R's matrix(NA)
    cnt = gpu_ctx.bed_counts(g)
    assert np.array_equal(cnt[:, 0], (raw == 2).sum(0)) and np.array_equal(cnt[:, 1], (raw < 0).sum(0))
    assert np.array_equal(cnt[:, 2], (raw == 1).sum(0)) and np.array_equal(cnt[:, 3], (raw == 0).sum(0))


def test_decode_missing_is_r_na_real(gpu_ctx):
    G = np.array([[2.0, np.nan], [np.nan, 0.0], [1.0, 1.0], [0.0, 2.0], [2.0, 2.0]])
    rows = to_bed_rows(G)
    dec = gpu_ctx.bed_decode(_bed_desc(gpu_ctx, rows, 5))
    assert np.array_equal(np.nan_to_num(dec, nan=-1), np.nan_to_num(G, nan=-1))
    assert dec.view(np.uint64)[1, 0] == 0x7FF00000000007A2


def test_c1_bundled_example_r_matrix_path(gpu_ctx, oracle, golden_dir):
    """config 1: SPAGxE_CCT on data/Geno.mtx (INTSXP matrix) + Pheno.mtx, binary trait."""
    d = np.load(f"{golden_dir}/c1.npz")
    y = d["y"]
    R = get_resid("binary", y, [d[k] for k in ["Cov1", "Cov2", "E", "PC1", "PC2", "PC3", "PC4"]])
    cova = np.column_stack([d[k] for k in ["PC1", "PC2", "PC3", "PC4", "Cov1", "Cov2"]])
    ids = [f"IID-{i + 1}" for i in range(len(y))]
    G = d["geno"].astype(np.int32)
    for eps in (0.001, 0.6):
        res = SPAGxE_CCT("binary", Geno_mtx=(G, ids, list(d["snp_names"])), R=R, E=d["E"],
                         Phen_mtx={"ID": ids, "Y": y}, Cova_mtx=cova, epsilon=eps, ctx=gpu_ctx, quiet=True)
        ref, route, iters, rc = oracle.cct_matrix("binary", G.astype(float), R, d["E"], y, cova, epsilon=eps, g_is_int=True)
        assert rc == 0
        assert_parity(res.values, ref, P_COLS, S_COLS, label=f"C1 eps={eps}")
        f, b, s = routes_from_cct_output(res.values, eps)
        assert np.array_equal(b, (route & 2) > 0) and np.array_equal(s, (route & 4) > 0)
        assert res.diag["n_branch_b"] == int(((route & 2) > 0).sum()) and res.diag["n_spa"] == int(((route & 4) > 0).sum())
        assert res.diag["newton_iters"] == int(iters.sum())
        assert res.rownames[7] == "rs8" and res.colnames[4] == "p.value.spaGxE.CCT.Wald"
    # the float (REALSXP) matrix path gives identical numbers
    res2 = SPAGxE_CCT("binary", Geno_mtx=(G.astype(np.float64), ids, None), R=R, E=d["E"],
                      Phen_mtx={"ID": ids, "Y": y}, Cova_mtx=cova, epsilon=0.6, ctx=gpu_ctx, quiet=True)
    assert np.array_equal(np.nan_to_num(res2.values, nan=-1), np.nan_to_num(res.values, nan=-1))
    # the reference stops when Phen.mtx has IID instead of ID
    from spagxecct_b200 import RStop
    with pytest.raises(RStop):
        SPAGxE_CCT("binary", Geno_mtx=(G, ids, None), R=R, E=d["E"], Phen_mtx={"IID": ids, "Y": y}, Cova_mtx=cova,
                   ctx=gpu_ctx, quiet=True)


@pytest.mark.parametrize("trait", ["binary", "quantitative"])
def test_c2_bed_path(gpu_ctx, oracle, golden_dir, trait, tmp_path):
    """config 2: SPAGxE_CCT reading inst/GenoMat_SPAGxE.bed, phenotype per roxygen example 6 but seeded."""
    d = np.load(f"{golden_dir}/bed_SPAGxE.npz")
    n = int(d["n"])
    rng = np.random.default_rng(20250101)
    Cov1 = rng.standard_normal(n); Cov2 = rng.binomial(1, 0.5, n).astype(float); E = rng.standard_normal(n)
    Y = rng.binomial(1, 0.5, n).astype(float) if trait == "binary" else rng.normal(1, 0.5, n)
    R = get_resid(trait, Y, [Cov1, Cov2, E])
    cova = np.column_stack([Cov1, Cov2])
    stem = str(tmp_path / "GenoMat_SPAGxE")
    d["bed"].tofile(stem + ".bed")
    ids = [f"IID-{i + 1}" for i in range(n)]
    open(stem + ".fam", "w").write("".join(f"{s} {s} 0 0 1 -9\n" for s in ids))
    open(stem + ".bim", "w").write("".join(f"1\t{s}\t0\t{j + 1}\tA\tG\n" for j, s in enumerate(d["snp_ids"])))
    for eps in (0.001, 0.3):
        res = SPAGxE_CCT(trait, GenoFile=stem + ".bed", R=R, E=E, Phen_mtx={"ID": ids, "Y": Y}, Cova_mtx=cova,
                         epsilon=eps, ctx=gpu_ctx, quiet=True)
        ref, route, iters, rc = oracle.cct_bed(trait, d["bed"][3:], n, R, E, Y, cova, epsilon=eps)
        assert rc == 0
        assert_parity(res.values, ref, P_COLS, S_COLS, label=f"C2 {trait} eps={eps}")
        f, b, s = routes_from_cct_output(res.values, eps)
        assert np.array_equal(b, (route & 2) > 0) and np.array_equal(s, (route & 4) > 0)
        assert res.rownames[:2] == ["SNP-1", "SNP-2"]
    assert b.sum() >= 10  # eps=0.3 exercises branch B (Wald + CCT) on real fixture genotypes


@pytest.mark.parametrize("trait,prev", [("binary", 0.1), ("binary", 0.01), ("quantitative", 0.0)])
def test_synthetic_missing_flip_filter_branches(gpu_ctx, oracle, trait, prev):
    """Edge cases no reference fixture exercises: missing calls, MAF > 0.5 flips,
    min.maf filter incl. an exact tie, ragged N (not a multiple of 4/16/64), SPA-heavy rare variants,
    strong marginal signals (branch B with SPA inside)."""
    n = 6007
    ph = make_pheno(n, 11, prevalence=prev or 0.1, trait=trait)
    mafs = np.r_[np.full(6, 0.0004), [0.001, 0.002, 0.004, 0.008], np.linspace(0.01, 0.5, 40), np.full(10, 0.02)]
    m = mafs.size
    rng = np.random.default_rng(5)
    miss = np.where(rng.random(m) < 0.5, 0.0, rng.random(m) * 0.08)
    major = rng.random(m) < 0.5
    effect = np.zeros(m); effect[-10:] = np.linspace(0.3, 1.2, 10)
    G = make_geno(n, mafs, 12, miss_rates=miss, count_major=major, effect=effect, Y=ph["Y"])
    # exact MAF == min.maf tie: (S/n)/2 == min.maf is NOT filtered (`MAF < min.maf`, ref :576)
    G[:, 6] = 0.0; G[:12, 6] = 1.0
    min_maf = (12.0 / n) / 2
    G = np.ascontiguousarray(G)
    rows = to_bed_rows(G)
    gpu_ctx.set_vectors(ph["R"], ph["E"])
    gpu_ctx.set_wald(_lib.TRAIT_BINARY if trait == "binary" else _lib.TRAIT_QUANTITATIVE, ph["Y"], ph["Cova"])
    for eps in (0.001, 0.05):
        out, diag = gpu_ctx.run_cct(_bed_desc(gpu_ctx, rows, n), dict(epsilon=eps, min_maf=min_maf))
        ref, route, iters, rc = oracle.cct_bed(trait, rows, n, ph["R"], ph["E"], ph["Y"], ph["Cova"], epsilon=eps,
                                               min_maf=min_maf)
        assert rc == 0
        worst = assert_parity(out, ref, P_COLS, S_COLS, label=f"synthetic {trait} eps={eps}")
        f, b, s = routes_from_cct_output(out, eps)
        assert np.array_equal(f, (route & 1) > 0) and np.array_equal(b, (route & 2) > 0) and np.array_equal(s, (route & 4) > 0)
        assert diag["n_filtered"] == f.sum() and diag["newton_iters"] == iters.sum()
    assert f.sum() >= 5 and b.sum() >= 8 and (s.sum() >= 3 or trait == "quantitative")
    assert (out[:, 1] > 0).sum() >= 10          # missing.rate column exercised
    assert (~np.isnan(out[6, 2]))               # the exact tie row is analysed, like `MAF < min.maf`
    # dense double matrix input (R matrix with NA) == packed input
    Gd, gd = np.asfortranarray(G), None
    gd = gpu_ctx.geno_desc(_lib.GENO_F64, Gd.ctypes.data, n, m, n)
    out2, _ = gpu_ctx.run_cct(gd, dict(epsilon=0.05, min_maf=min_maf))
    assert np.array_equal(np.nan_to_num(out2, nan=-1), np.nan_to_num(out, nan=-1))


def test_one_snp_entry_point_and_errors(gpu_ctx, oracle):
    n = 3001
    ph = make_pheno(n, 21)
    G = make_geno(n, [0.2, 0.05], 22, miss_rates=[0.02, 0.0])
    phen = {"Y": ph["Y"]}
    v = SPAGxE_CCT_one_SNP("binary", G[:, 0], ph["R"], ph["E"], phen, ph["Cova"], ctx=gpu_ctx)
    ref, *_ = oracle.cct_matrix("binary", G[:, :1], ph["R"], ph["E"], ph["Y"], ph["Cova"])
    assert_parity(v[None, :], ref, P_COLS, S_COLS, label="one_SNP")
    # imputed (non-integer) dosages are never rounded: the vector goes through the dense kernel (tests/test_dosage.py)
    g = G[:, 1].copy(); g[3] = 0.37
    v = SPAGxE_CCT_one_SNP("binary", g, ph["R"], ph["E"], phen, ph["Cova"], ctx=gpu_ctx)
    ref, *_ = oracle.cct_matrix("binary", g[:, None], ph["R"], ph["E"], ph["Y"], ph["Cova"])
    assert np.isclose(v[0], ref[0, 0], rtol=1e-13)
    assert_parity(v[None, :], ref, P_COLS, S_COLS, exact_cols=(1,), label="one_SNP dosage")
    # sample-count mismatch
    with pytest.raises(_lib.SpagxeError) as ei:
        gpu_ctx.run_cct(gpu_ctx.geno_desc(_lib.GENO_F64, G.ctypes.data, n - 1, 1, n - 1))
    assert ei.value.code == _lib.ERR_ARG
    with pytest.raises(RStop):                   # a trait the reference does not know
        SPAGxE_CCT_one_SNP("ordinal", G[:, 0], ph["R"], ph["E"], phen, ph["Cova"], ctx=gpu_ctx)


def test_extreme_pvalues_down_to_1e300(gpu_ctx, oracle):
    """p-values must agree down to p = 1e-300 (north_star): a quantitative trait with a huge G x E effect."""
    n = 20011
    rng = np.random.default_rng(31)
    ph = make_pheno(n, 32, trait="quantitative")
    g = rng.binomial(2, 0.3, n).astype(float)
    effs = [0.0, 0.05, 0.1, 0.15, 0.19]
    G = np.column_stack([g] * len(effs))
    outs, refs = [], []
    for k, b in enumerate(effs):
        Y = ph["Y"] + b * g * ph["E"] + 0.2 * b * g
        R = get_resid("quantitative", Y, [ph["Cova"][:, 0], ph["Cova"][:, 1], ph["E"]])
        gpu_ctx.set_vectors(R, ph["E"]); gpu_ctx.set_wald(_lib.TRAIT_QUANTITATIVE, Y, ph["Cova"])
        rows = to_bed_rows(G[:, k:k + 1])
        for eps in (0.0, 1.0):  # branch A and branch B
            out, _ = gpu_ctx.run_cct(_bed_desc(gpu_ctx, rows, n), dict(epsilon=eps))
            ref, *_ = oracle.cct_bed("quantitative", rows, n, R, ph["E"], Y, ph["Cova"], epsilon=eps)
            outs.append(out); refs.append(ref)
    out = np.vstack(outs); ref = np.vstack(refs)
    assert_parity(out, ref, P_COLS, S_COLS, label="extreme")
    pv = ref[:, [2, 3, 5]]
    assert np.nanmin(pv[pv > 0]) < 1e-200 and (pv == 0).sum() <= 2, (np.nanmin(pv[pv > 0]), (pv == 0).sum())


def test_properties_at_scale(gpu_ctx):
    """Size-independent properties at a size the oracle would take minutes for (N=100k x 2,048 SNPs):
    (a) device-resident packed input == host input; (b) SNP order permutation permutes the output rows;
    (c) counting the other allele (2-g) gives the same MAF and p-values (flip symmetry, ref :567-570)."""
    n, m = 100_003, 2048
    ph = make_pheno(n, 41, prevalence=0.01)
    rng = np.random.default_rng(42)
    mafs = rng.uniform(0.001, 0.5, m)
    G = (rng.random((n, m)) < mafs).astype(np.int8) + (rng.random((n, m)) < mafs).astype(np.int8)
    miss = rng.random((n, m)) < (rng.random(m) * 0.03 * (rng.random(m) < 0.5))
    Gf = G.astype(np.float64); Gf[miss] = np.nan
    rows = to_bed_rows(Gf)
    gpu_ctx.set_vectors(ph["R"], ph["E"]); gpu_ctx.set_wald(_lib.TRAIT_BINARY, ph["Y"], ph["Cova"])
    base, diag = gpu_ctx.run_cct(_bed_desc(gpu_ctx, rows, n))
    assert diag["n_spa"] > 20 and diag["n_branch_b"] >= 1
    rb = (n + 3) // 4
    perm = rng.permutation(m)
    rows_p = np.ascontiguousarray(rows.reshape(m, rb)[perm]).ravel()
    outp, _ = gpu_ctx.run_cct(_bed_desc(gpu_ctx, rows_p, n))
    assert np.array_equal(np.nan_to_num(outp, nan=-1), np.nan_to_num(base[perm], nan=-1))
    rows_f = to_bed_rows(2 - Gf)
    outf, _ = gpu_ctx.run_cct(_bed_desc(gpu_ctx, rows_f, n))
    ok = ~np.isnan(base[:, 2]) & (np.abs(base[:, 0] - 0.5) > 1e-9)
    assert np.allclose(outf[ok, 0], base[ok, 0], rtol=1e-13) and np.array_equal(outf[:, 1], base[:, 1])
    lp = np.abs(np.log10(outf[ok][:, P_COLS]) - np.log10(base[ok][:, P_COLS]))
    assert np.nanmax(lp / np.maximum(np.abs(np.log10(base[ok][:, P_COLS])), 1e-3)) < 1e-6
    # device-resident rows with a 128-byte aligned pitch
    import torch
    pitch = (rb + 127) // 128 * 128
    dev = torch.full((m, pitch), 0, dtype=torch.uint8, device="cuda")
    dev[:, :rb] = torch.from_numpy(rows.reshape(m, rb)).cuda()
    torch.cuda.synchronize()
    gd = gpu_ctx.geno_desc(_lib.GENO_BED, dev.data_ptr(), n, m, pitch, _lib.LOC_DEVICE)
    outd, _ = gpu_ctx.run_cct(gd)
    assert np.array_equal(np.nan_to_num(outd, nan=-1), np.nan_to_num(base, nan=-1))


def test_spa_quadrature_matches_oracle_and_exact_kernel(oracle):
    """N large enough for the SNP-invariant quadrature of the CGF sums (K4c): p-values and Newton
    iteration counts must match the per-sample oracle, and the exact per-sample kernel."""
    from spagxecct_b200 import Context
    n, m = 120_011, 400
    ph = make_pheno(n, 51, prevalence=0.01)
    rng = np.random.default_rng(52)
    mafs = np.r_[rng.uniform(0.0015, 0.02, 300), rng.uniform(0.02, 0.5, 100)]
    G = make_geno(n, mafs, 53, miss_rates=np.where(rng.random(m) < 0.5, 0, 0.02), count_major=rng.random(m) < 0.5)
    rows = to_bed_rows(G)
    ref, route, iters, rc = oracle.cct_bed("binary", rows, n, ph["R"], ph["E"], ph["Y"], ph["Cova"], nthreads=8)
    assert rc == 0 and ((route & 4) > 0).sum() >= 10
    outs = []
    for exact in (False, True):
        with Context(0) as ctx:
            ctx.set_option(_lib.OPT_EXACT_SPA, int(exact))
            ctx.set_vectors(ph["R"], ph["E"]); ctx.set_wald(_lib.TRAIT_BINARY, ph["Y"], ph["Cova"])
            out, diag = ctx.run_cct(_bed_desc(ctx, rows, n))
        assert_parity(out, ref, P_COLS, S_COLS, label=f"quad exact={exact}")
        assert diag["newton_iters"] == iters.sum() and diag["n_spa"] == ((route & 4) > 0).sum()
        outs.append(out)
    spa = (route & 4) > 0
    d = np.abs(np.log10(outs[0][spa, 2]) - np.log10(outs[1][spa, 2]))
    assert d.max() < 1e-9, d.max()


@pytest.mark.parametrize("n,m", [(6007, 70), (1000, 3), (50_021, 300)])
def test_score_tensor_core_kernel_matches_fp64_kernel(gpu_ctx, n, m):
    """K1+K2: the tcgen05 int8 fixed-point kernel against the fp64 CUDA-core kernel: integer outputs
    (allele sums, missing counts) are identical; the exact-integer contraction agrees with the fp64
    sums to rounding; ragged N and M (not multiples of the 512-sample / 128-row tiles)."""
    ph = make_pheno(n, 61, prevalence=0.05)
    rng = np.random.default_rng(62)
    mafs = rng.uniform(0.001, 0.5, m)
    G = make_geno(n, mafs, 63, miss_rates=np.where(rng.random(m) < 0.5, 0, rng.random(m) * 0.1), count_major=rng.random(m) < 0.5)
    rows = to_bed_rows(G)
    gpu_ctx.set_vectors(ph["R"], ph["E"])
    g = _bed_desc(gpu_ctx, rows, n)
    tc = gpu_ctx.score_stats(g, impl=0)
    v1 = gpu_ctx.score_stats(g, impl=1)
    assert np.array_equal(tc["asum"], v1["asum"]) and np.array_equal(tc["nmiss"], v1["nmiss"])
    Gz = np.nan_to_num(G, nan=0.0)
    assert np.array_equal(tc["asum"], Gz.sum(0).astype(np.int32)) and np.array_equal(tc["nmiss"], np.isnan(G).sum(0))
    scale = np.sqrt((ph["R"] ** 2).sum())
    for k in ("D0", "D1", "T0", "T1"):
        assert np.max(np.abs(tc[k] - v1[k])) <= 1e-11 * scale * 5, (k, np.max(np.abs(tc[k] - v1[k])))
    # the tensor-core result is exact integer arithmetic: compare with a long-double host sum
    want = (Gz.astype(np.longdouble) * ph["R"].astype(np.longdouble)[:, None]).sum(0)
    assert np.max(np.abs(tc["D0"] - want.astype(np.float64))) <= 2e-14 * np.abs(ph["R"]).max() * n
    # run twice: bit-identical (integer accumulation is order independent)
    tc2 = gpu_ctx.score_stats(g, impl=0)
    assert all(np.array_equal(tc[k], tc2[k]) for k in tc)


P_COLS_PLUS = [2, 3, 4]
S_COLS_PLUS = [5, 6, 7]


def _plus_nullmodel(gpu_ctx, oracle, golden_dir):
    """Step-1 object of SPAGxE_Plus: the GRM quadratic forms come from the device (spagxe_plus_nullmodel) and are
    checked against the oracle's restatement of R/SPAGxEPlus_functions_MYZ.R:490-537."""
    d = np.load(f"{golden_dir}/plus.npz")
    R = get_resid("binary", d["Y"], [d["Cov1"], d["Cov2"], d["E"]])
    a, b, Rnew, c = gpu_ctx.plus_nullmodel(d["grm_i"], d["grm_j"], d["grm_v"], R, d["E"])
    a0, b0, Rnew0, c0 = oracle.plus_nullmodel_scalars(d["grm_i"], d["grm_j"], d["grm_v"], R, d["E"])
    assert np.allclose([a, b, c], [a0, b0, c0], rtol=1e-12) and np.allclose(Rnew, Rnew0, rtol=1e-12, atol=1e-15)
    obj = {"R": R, "R_GRM_R": a, "R_GRM_RE": b, "R.new": Rnew, "R.new_GRM_R.new": c,
           "sparseGRM": {"i": d["grm_i"], "j": d["grm_j"], "Value": d["grm_v"]}}
    return d, obj


def test_c5_spagxe_plus_bundled_fixture(gpu_ctx, oracle, golden_dir):
    """SPAGxE_Plus on inst/GenoMat_SPAGxE_Plus.bed + data/sparseGRM.SPAGxEPlus.rda + the bundled binary
    phenotype (prevalence 1%), branch A and (epsilon = 0.5) the sparse-GRM branch B."""
    from spagxecct_b200 import SPAGxE_Plus
    d, obj = _plus_nullmodel(gpu_ctx, oracle, golden_dir)
    bed = np.load(f"{golden_dir}/bed_SPAGxE_Plus.npz")
    n = int(bed["n"])
    rows = np.ascontiguousarray(bed["bed"][3:])
    G = oracle.bed_decode(rows, n)
    for eps, cutoff in ((0.001, 2.0), (0.5, 2.0), (0.5, 0.5)):
        res = SPAGxE_Plus(Geno_mtx=G, E=d["E"], Phen_mtx={"Y": d["Y"]}, Cova_mtx=np.column_stack([d["Cov1"], d["Cov2"]]),
                          obj_SPAGxE_Plus_Nullmodel=obj, epsilon=eps, Cutoff=cutoff, ctx=gpu_ctx, quiet=True)
        ref, route, iters, _ = oracle.plus_bed(rows, n, obj["R"], d["E"], obj["R.new"], obj["R_GRM_R"], obj["R_GRM_RE"],
                                              obj["R.new_GRM_R.new"], d["grm_i"], d["grm_j"], d["grm_v"], epsilon=eps,
                                              cutoff=cutoff)
        assert_parity(res.values, ref, P_COLS_PLUS, S_COLS_PLUS, label=f"plus eps={eps} cutoff={cutoff}")
        assert res.diag["n_branch_b"] == ((route & 2) > 0).sum() and res.diag["n_spa"] == ((route & 4) > 0).sum()
        assert res.diag["newton_iters"] == iters.sum()
        assert res.colnames[2] == "p.value.spaGxE.plus"
    assert ((route & 2) > 0).sum() >= 20 and ((route & 4) > 0).sum() >= 10


def test_spagxe_plus_synthetic_families_with_missing_and_flips(gpu_ctx, oracle):
    """Related samples (sib pairs), missing calls, MAF > 0.5 coding, N not a multiple of anything."""
    n_fam = 3001
    n = 2 * n_fam + 1001
    rng = np.random.default_rng(71)
    ph = make_pheno(n, 72, prevalence=0.05)
    mafs = rng.uniform(0.005, 0.5, 60)
    G = make_geno(n, mafs, 73, miss_rates=np.where(rng.random(60) < 0.5, 0, 0.03), count_major=rng.random(60) < 0.5)
    G[1:2 * n_fam:2] = np.where(rng.random((n_fam, 60)) < 0.5, G[0:2 * n_fam:2], G[1:2 * n_fam:2])  # sibs share alleles
    gi = np.r_[np.arange(n), np.arange(1, 2 * n_fam, 2)].astype(np.int32)
    gj = np.r_[np.arange(n), np.arange(0, 2 * n_fam, 2)].astype(np.int32)
    gv = np.r_[np.ones(n), np.full(n_fam, 0.5)]
    a, b, Rnew, c = gpu_ctx.plus_nullmodel(gi, gj, gv, ph["R"], ph["E"])
    a0, b0, Rnew0, c0 = oracle.plus_nullmodel_scalars(gi, gj, gv, ph["R"], ph["E"])
    assert np.allclose([a, b, c], [a0, b0, c0], rtol=1e-12) and np.allclose(Rnew, Rnew0, rtol=1e-12, atol=1e-15)
    rows = to_bed_rows(G)
    gpu_ctx.set_vectors(ph["R"], ph["E"])
    gpu_ctx.set_grm(gi, gj, gv, a, b, c, Rnew)
    for eps in (0.001, 0.3):
        out, diag = gpu_ctx.run_plus(_bed_desc(gpu_ctx, rows, n), dict(epsilon=eps, cutoff=1.0))
        ref, route, iters, _ = oracle.plus_bed(rows, n, ph["R"], ph["E"], Rnew, a, b, c, gi, gj, gv, epsilon=eps, cutoff=1.0)
        assert_parity(out, ref, P_COLS_PLUS, S_COLS_PLUS, label=f"plus synthetic eps={eps}")
        assert diag["newton_iters"] == iters.sum()
    assert ((route & 2) > 0).sum() >= 5


P_COLS_MIX = [2, 3, 4, 5, 6]
S_COLS_MIX = [7, 8, 9]


def _mix_inputs(golden_dir, trait="binary"):
    """config-4 recipe at the fixture's scale (N = 10,000): two-way admixture proportions a6 and the bundled
    top-10 PCs (simulation_studies/SPAGxEmixCCT/data), ancestry-specific allele frequencies and prevalence."""
    d = np.load(f"{golden_dir}/mix.npz")
    alpha = d["a6"][:, 0]
    PCs = d["top10PCs"]
    n = alpha.size
    rng = np.random.default_rng(81)
    pairs = [(0.3, 0.01), (0.1, 0.01), (0.05, 0.4), (0.01, 0.001), (0.2, 0.0), (0.0, 0.15), (0.0012, 0.0003), (0.25, 0.25)]
    m = 64
    G = np.empty((n, m))
    for j in range(m):
        p1, p2 = pairs[j % len(pairs)]
        p = alpha * p1 + (1 - alpha) * p2
        g = rng.binomial(2, p).astype(float)
        if j % 3 == 0:
            g[rng.random(n) < 0.02] = np.nan
        if j % 5 == 0:
            g = 2 - g
        G[:, j] = g
    X1 = rng.standard_normal(n); X2 = rng.binomial(1, 0.5, n).astype(float); E = rng.standard_normal(n)
    eta = 0.5 * X1 + 0.5 * X2 + 0.5 * E + 1.5 * (alpha - alpha.mean())
    if trait == "binary":
        Y = rng.binomial(1, 1 / (1 + np.exp(-(-2.5 + eta)))).astype(float)
    else:
        Y = eta + rng.standard_normal(n)
    cova = np.column_stack([X1, X2, PCs[:, 0], PCs[:, 1]])
    R = get_resid(trait, Y, [X1, X2, PCs[:, 0], PCs[:, 1], E])
    return dict(G=np.ascontiguousarray(G), R=R, E=E, Y=Y, Cova=cova, PCs=PCs, n=n)


@pytest.mark.parametrize("trait", ["binary", "quantitative"])
def test_c4_spagxemix_admixed(gpu_ctx, oracle, golden_dir, trait):
    from spagxecct_b200 import SPAGxEmix_CCT
    x = _mix_inputs(golden_dir, trait)
    rows = to_bed_rows(x["G"])
    n = x["n"]
    for eps in (0.001, 0.2):
        res = SPAGxEmix_CCT(trait, Geno_mtx=x["G"], R=x["R"], E=x["E"], Phen_mtx={"Y": x["Y"]}, Cova_mtx=x["Cova"],
                            topPCs=x["PCs"], epsilon=eps, Cutoff=2, min_maf=0.0002, ctx=gpu_ctx, quiet=True)
        ref, route, iters, rc = oracle.mix_bed(trait, rows, n, x["R"], x["E"], x["Y"], x["Cova"], x["PCs"], epsilon=eps,
                                               min_maf=0.0002, nthreads=8)
        assert rc == 0
        assert_parity(res.values, ref, P_COLS_MIX, S_COLS_MIX, exact_cols=(0, 1, 10), label=f"mix {trait} eps={eps}")
        assert np.allclose(res.values[:, 11], ref[:, 11], rtol=1e-14, equal_nan=True)   # MAC
        d = res.diag
        assert d["n_branch_b"] == ((route & 2) > 0).sum() and d["n_spa"] == ((route & 4) > 0).sum()
        assert d["n_mix_logistic"] == ((route & 256) > 0).sum() and d["newton_iters"] == iters.sum()
        assert res.colnames[10:] == ["MAF.est.negative.num", "MAC"]
    assert ((route & 256) > 0).sum() >= 3 and ((route & 64) > 0).sum() >= 1 and ((route & 2) > 0).sum() >= 5


def test_multi_chunk_streaming_and_deferred_branch_b(oracle):
    """Many small chunks through the double-buffered host staging path (ragged last chunk): the deferred branch-B
    items of chunk k must still see their rows when chunk k+2 is staged; result == single-chunk result == oracle."""
    from spagxecct_b200 import Context
    n, m = 5003, 333
    ph = make_pheno(n, 91, prevalence=0.1)
    rng = np.random.default_rng(92)
    mafs = rng.uniform(0.002, 0.5, m)
    eff = np.where(rng.random(m) < 0.25, 0.8, 0.0)
    G = make_geno(n, mafs, 93, miss_rates=np.where(rng.random(m) < 0.5, 0, 0.03), count_major=rng.random(m) < 0.5,
                  effect=eff, Y=ph["Y"])
    rows = to_bed_rows(G)
    outs = []
    for chunk in (0, 37):
        with Context(0) as ctx:
            ctx.set_option(_lib.OPT_CHUNK_SNPS, chunk)
            ctx.set_vectors(ph["R"], ph["E"]); ctx.set_wald(_lib.TRAIT_BINARY, ph["Y"], ph["Cova"])
            out, diag = ctx.run_cct(_bed_desc(ctx, rows, n), dict(epsilon=0.05))
            Gd = np.asfortranarray(G)
            out_dense, _ = ctx.run_cct(ctx.geno_desc(_lib.GENO_F64, Gd.ctypes.data, n, m, n), dict(epsilon=0.05))
        outs.append(out)
        assert np.array_equal(np.nan_to_num(out_dense, nan=-1), np.nan_to_num(out, nan=-1))
    assert diag["n_branch_b"] >= 30
    # the score pass is exact integer arithmetic (bit-identical under any chunking); branch-B fp64 sums are grouped by
    # the cluster size chosen per flush, so those columns agree to rounding only
    a, b = np.nan_to_num(outs[0], nan=-1), np.nan_to_num(outs[1], nan=-1)
    assert np.array_equal(a[:, [0, 1, 6, 7, 8, 9]], b[:, [0, 1, 6, 7, 8, 9]])
    assert np.allclose(a, b, rtol=1e-11, atol=0), np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
    ref, route, iters, rc = oracle.cct_bed("binary", rows, n, ph["R"], ph["E"], ph["Y"], ph["Cova"], epsilon=0.05, nthreads=8)
    assert rc == 0
    assert_parity(outs[1], ref, P_COLS, S_COLS, label="multi-chunk")


def test_edge_cases_empty_tiny_and_fatal_inputs(gpu_ctx, oracle):
    """Empty input, N smaller than one 512-sample tile / one 16-sample word, a SNP with every call missing
    (the reference stops at `if(MAF > 0.5)`), monomorphic SNPs (filtered), M not a multiple of 128."""
    n = 13
    rng = np.random.default_rng(101)
    R = rng.standard_normal(n); R -= R.mean(); E = rng.standard_normal(n)
    Y = (rng.random(n) < 0.5).astype(float); cova = rng.standard_normal((n, 1))
    gpu_ctx.set_vectors(R, E); gpu_ctx.set_wald(_lib.TRAIT_BINARY, Y, cova)
    G = rng.integers(0, 3, (n, 5)).astype(float)
    G[:, 2] = 0.0                       # monomorphic -> MAF 0 < min.maf -> NA row
    G[3, 1] = np.nan
    rows = to_bed_rows(G)
    out, diag = gpu_ctx.run_cct(_bed_desc(gpu_ctx, rows, n), dict(epsilon=0.0))
    ref, route, _, rc = oracle.cct_bed("binary", rows, n, R, E, Y, cova, epsilon=0.0)
    assert rc == 0
    assert_parity(out, ref, P_COLS, S_COLS, label="tiny N")
    assert np.isnan(out[2, 2:]).all() and out[2, 0] == 0.0
    # M = 0
    empty = np.zeros(0, dtype=np.uint8)
    g0 = gpu_ctx.geno_desc(_lib.GENO_BED, rows.ctypes.data, n, 0, (n + 3) // 4)
    out0, d0 = gpu_ctx.run_cct(g0)
    assert out0.shape == (0, 10) and d0["n_snps"] == 0
    # all-missing SNP: fatal, like the reference
    Gm = G.copy(); Gm[:, 4] = np.nan
    with pytest.raises(_lib.SpagxeError) as ei:
        gpu_ctx.run_cct(_bed_desc(gpu_ctx, to_bed_rows(Gm), n))
    assert ei.value.code == _lib.ERR_ARG and "missing" in str(ei.value)
    # state errors
    from spagxecct_b200 import Context
    with Context(0) as c2:
        with pytest.raises(_lib.SpagxeError) as ei:
            c2.run_cct(_bed_desc(c2, rows, n))
        assert ei.value.code == _lib.ERR_STATE
        c2.set_vectors(R, E)
        with pytest.raises(_lib.SpagxeError) as ei:
            c2.run_plus(_bed_desc(c2, rows, n))
        assert ei.value.code == _lib.ERR_STATE and "set_grm" in str(ei.value)
        with pytest.raises(_lib.SpagxeError) as ei:
            c2.set_wald(5, Y, cova)            # no such trait
        assert ei.value.code == _lib.ERR_ARG


@pytest.mark.parametrize("e_offset", [0.0, 1950.0])
def test_mix_batched_common_path_matches_per_snp_kernel(gpu_ctx, golden_dir, e_offset):
    """SPAGxEmix_CCT: the batched common path (mix_batch.cu) against the per-SNP cluster kernel on the same inputs
    (tiled 3x so that the sample axis spans several splits and a ragged last tile).  e_offset: an uncentred E (a birth
    year): the batched path expands its statistics in lambda and must not lose digits to cancellation."""
    from spagxecct_b200 import SPAGxEmix_CCT
    x = _mix_inputs(golden_dir, "binary")
    rep = 3
    n = x["n"] * rep - 37
    rng = np.random.default_rng(5)
    def tile(v):
        return np.concatenate([v] * rep, axis=0)[:n]
    G = tile(x["G"]).copy()
    G[rng.random(G.shape) < 0.001] = np.nan
    PCs = tile(x["PCs"]) + 1e-3 * rng.standard_normal((n, x["PCs"].shape[1]))
    args = dict(Geno_mtx=np.ascontiguousarray(G), R=tile(x["R"]) - tile(x["R"]).mean(), E=tile(x["E"]) + e_offset,
                Phen_mtx={"Y": tile(x["Y"])}, Cova_mtx=tile(x["Cova"]), topPCs=PCs, epsilon=0.001, Cutoff=2, min_maf=0.0002,
                ctx=gpu_ctx, quiet=True)
    fast = SPAGxEmix_CCT("binary", **args)
    gpu_ctx.set_option(_lib.OPT_MIX_PER_SNP, 1)
    try:
        slow = SPAGxEmix_CCT("binary", **args)
    finally:
        gpu_ctx.set_option(_lib.OPT_MIX_PER_SNP, 0)
    assert_parity(fast.values, slow.values, P_COLS_MIX, S_COLS_MIX, exact_cols=(0, 1, 10, 11), label="mix batched vs per-SNP")
    for key in ("n_branch_a", "n_branch_b", "n_spa", "n_mix_logistic", "newton_iters", "n_filtered"):
        assert fast.diag[key] == slow.diag[key], key
    # the batched path must actually have served most SNPs: far fewer kernel launches cannot tell, the counters can
    assert fast.diag["n_branch_a"] > fast.diag["n_spa"] + fast.diag["n_branch_b"]


def test_genofile_sample_subset_reorder_marker_filters_and_allele_order(gpu_ctx, oracle, golden_dir, tmp_path):
    """What GRAB.ReadGeno does between the file and the loop (ref R/SPAGxECCT.R:484-487): Phen.mtx$ID is a shuffled
    SUBSET of the .fam samples (device-side 2-bit gather), `control` selects markers (ids / ranges, include or
    exclude) and the allele order; each variant must equal the oracle on the correspondingly decoded matrix."""
    d = np.load(f"{golden_dir}/bed_SPAGxE.npz")
    n_file = int(d["n"])
    stem = str(tmp_path / "GenoMat_SPAGxE")
    d["bed"].tofile(stem + ".bed")
    fam_ids = [f"IID-{i + 1}" for i in range(n_file)]
    open(stem + ".fam", "w").write("".join(f"{s} {s} 0 0 1 -9\n" for s in fam_ids))
    snp_ids = [str(s) for s in d["snp_ids"]]
    open(stem + ".bim", "w").write("".join(f"{1 + j // 50}\t{s}\t0\t{1000 * (j + 1)}\tA\tG\n" for j, s in enumerate(snp_ids)))
    rng = np.random.default_rng(77)
    keep = rng.permutation(n_file)[:7001]                       # shuffled subset, N not a multiple of 4
    ids = [fam_ids[i] for i in keep]
    n = len(ids)
    ph = make_pheno(n, 78, prevalence=0.2)
    Gfull = oracle.bed_decode(d["bed"][3:], n_file)
    Gsub = np.ascontiguousarray(Gfull[keep])
    phen = {"ID": ids, "Y": ph["Y"]}
    # (a) all markers, sample gather only
    res = SPAGxE_CCT("binary", GenoFile=stem + ".bed", R=ph["R"], E=ph["E"], Phen_mtx=phen, Cova_mtx=ph["Cova"], epsilon=0.05,
                     ctx=gpu_ctx, quiet=True)
    ref, *_ = oracle.cct_matrix("binary", Gsub, ph["R"], ph["E"], ph["Y"], ph["Cova"], epsilon=0.05, g_is_int=False)
    assert_parity(res.values, ref, P_COLS, S_COLS, label="sample gather")
    assert res.rownames == snp_ids
    # (b) marker ids to include + ref-first allele order (counts A2: g -> 2 - g)
    inc = tmp_path / "inc.txt"
    inc.write_text("\n".join(snp_ids[3:40:3]) + "\n")
    res_b = SPAGxE_CCT("binary", GenoFile=stem + ".bed", control={"IDsToIncludeFile": str(inc), "AlleleOrder": "ref-first"},
                       R=ph["R"], E=ph["E"], Phen_mtx=phen, Cova_mtx=ph["Cova"], epsilon=0.05, ctx=gpu_ctx, quiet=True)
    cols = list(range(3, 40, 3))
    ref_b, *_ = oracle.cct_matrix("binary", 2 - Gsub[:, cols], ph["R"], ph["E"], ph["Y"], ph["Cova"], epsilon=0.05)
    assert_parity(res_b.values, ref_b, P_COLS, S_COLS, label="ids include + ref-first")
    assert res_b.rownames == [snp_ids[j] for j in cols]
    # (c) range exclusion: chromosome 2 between positions 60,000 and 80,000 (markers 60..80 of the fixture)
    exc = tmp_path / "exc.txt"
    exc.write_text("2 60000 80000\n")
    res_c = SPAGxE_CCT("binary", GenoFile=stem + ".bed", control={"RangesToExcludeFile": str(exc)}, R=ph["R"], E=ph["E"],
                       Phen_mtx=phen, Cova_mtx=ph["Cova"], epsilon=0.05, ctx=gpu_ctx, quiet=True)
    cols_c = [j for j in range(100) if not (59 <= j <= 79)]
    assert res_c.rownames == [snp_ids[j] for j in cols_c]
    assert np.array_equal(np.nan_to_num(res_c.values, nan=-1), np.nan_to_num(res.values[cols_c], nan=-1))
    # include + exclude together is refused like GRAB does; an unknown sample id stops
    from spagxecct_b200 import RStop
    with pytest.raises(RStop):
        SPAGxE_CCT("binary", GenoFile=stem + ".bed", control={"IDsToIncludeFile": str(inc), "RangesToExcludeFile": str(exc)},
                   R=ph["R"], E=ph["E"], Phen_mtx=phen, Cova_mtx=ph["Cova"], ctx=gpu_ctx, quiet=True)
    with pytest.raises(RStop):
        SPAGxE_CCT("binary", GenoFile=stem + ".bed", R=ph["R"], E=ph["E"], Phen_mtx={"ID": ids[:-1] + ["nobody"], "Y": ph["Y"]},
                   Cova_mtx=ph["Cova"], ctx=gpu_ctx, quiet=True)
