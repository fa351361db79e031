"""SPAGxEmix_CCT_Plus (ref R/SPAmixGxECCTPlus_functions_MYZ_2024-09-17.R:38-133, :144-427): oracle restatement against an
independent dense-GRM numpy / scipy restatement (CPU), host logic of the Python mirror (CPU), and the CUDA path
(spagxe_run_mix_plus through the C ABI) against the oracle (GPU)."""
import numpy as np
import pytest

from tests._data import assert_parity, to_bed_rows

P_COLS = [2, 5, 6]
S_COLS = [7, 8, 9]


def _family_grm(n, n_fam):
    """lower triangle incl. diagonal: diag 1, sib pairs (2f, 2f+1) 0.5 -- the layout of data/sparseGRM.SPAGxEPlus.rda"""
    gi = np.r_[np.arange(n), np.arange(1, 2 * n_fam, 2)].astype(np.int32)
    gj = np.r_[np.arange(n), np.arange(0, 2 * n_fam, 2)].astype(np.int32)
    gv = np.r_[np.ones(n), np.full(n_fam, 0.5)]
    return gi, gj, gv


def _small_inputs(n=1500, m=24, k=3, seed=5):
    rng = np.random.default_rng(seed)
    alpha = rng.beta(0.6, 0.6, n)
    n_fam = n // 3
    alpha[1:2 * n_fam:2] = alpha[0:2 * n_fam:2]          # sibs share their ancestry
    PCs = np.column_stack([alpha - alpha.mean()] + [rng.standard_normal(n) for _ in range(k - 1)])
    pairs = [(0.3, 0.02), (0.1, 0.01), (0.05, 0.4), (0.004, 0.001), (0.25, 0.25), (0.0, 0.2)]
    G = np.empty((n, m))
    for j in range(m):
        p1, p2 = pairs[j % len(pairs)]
        p = alpha * p1 + (1 - alpha) * p2
        g = rng.binomial(2, p).astype(float)
        sib = rng.random(n_fam) < 0.5
        g[1:2 * n_fam:2] = np.where(sib, g[0:2 * n_fam:2], g[1:2 * n_fam:2])
        if j % 3 == 0:
            g[rng.random(n) < 0.02] = np.nan
        if j % 5 == 0:
            g = 2 - g
        G[:, j] = g
    E = rng.standard_normal(n)
    Y = rng.binomial(1, 1 / (1 + np.exp(-(-2.0 + 0.5 * E + alpha + 0.6 * np.nan_to_num(G[:, 1]))))).astype(float)
    R = Y - Y.mean() + 0.05 * rng.standard_normal(n)
    return dict(G=G, R=R, E=E, PCs=PCs, n=n, grm=_family_grm(n, n_fam))


def _np_reference_one_snp(g, R, E, PCs, grm, epsilon, cutoff, mafest):
    """dense-matrix restatement of the reference's variance algebra with the allele frequencies `mafest` given
    (t(x) %*% GRM %*% x with the full symmetric matrix), scipy root finder for the saddlepoint."""
    from scipy.optimize import brentq
    from scipy.stats import norm
    n = g.size
    gi, gj, gv = grm
    K = np.zeros((n, n))
    K[gi, gj] = gv
    K = K + K.T - np.diag(np.diag(K))
    maf = np.nanmean(g) / 2
    g = np.where(np.isnan(g), 2 * maf, g)
    if maf > 0.5:
        maf, g = 1 - maf, 2 - g
    m = mafest
    gvar = 2 * m * (1 - m)
    sq = np.sqrt(gvar)
    S1 = g @ R
    VarS1 = (R * sq) @ K @ (R * sq)
    Z1 = (S1 - 2 * (R @ m)) / np.sqrt(VarS1)
    pG = 2 * norm.cdf(-abs(Z1))

    def tails(Rv, S):
        Smean = 2 * (Rv @ m)
        Svar = (Rv * sq) @ K @ (Rv * sq)
        z = (S - Smean) / np.sqrt(Svar)
        if abs(z) < cutoff:
            return 2 * norm.sf(abs(z)), 2 * norm.sf(abs(z))
        ratio = (Rv ** 2 @ gvar) / Svar

        def k0(t): return np.sum(2 * np.log(1 - m + m * np.exp(t * Rv)))
        def k1(t): e = m * np.exp(t * Rv); return np.sum(2 * Rv * e / (1 - m + e))
        def k2(t): e = m * np.exp(t * Rv); return np.sum(2 * Rv ** 2 * e * (1 - m) / (1 - m + e) ** 2)
        ps = 0.0
        for s, lower in ((max(S, 2 * Smean - S), False), (min(S, 2 * Smean - S), True)):
            zeta = brentq(lambda t: k1(t) - s, -50, 50, xtol=1e-14, rtol=1e-14)
            w = np.sign(zeta) * np.sqrt(2 * (zeta * s - k0(zeta)))
            v = zeta * np.sqrt(k2(zeta))
            a = (w + np.log(v / w) / w) * np.sqrt(ratio)
            ps += norm.cdf(a) if lower else norm.sf(a)
        return ps, 2 * norm.sf(abs(z))

    if pG > epsilon:
        S2 = (g * E) @ R
        Cov = (R * sq) @ K @ (R * E * sq)
        lam = Cov / VarS1
        pspa, pn = tails((E - lam) * R, S2 - lam * S1)
    else:
        W = np.column_stack([np.ones(n), g])
        R0 = R - W @ np.linalg.solve(W.T @ W, W.T @ R)
        pspa, pn = tails(E * R0, (g * E) @ R0)
    return pspa, pn, pG, S1, VarS1, Z1


def test_oracle_mixplus_vs_dense_numpy_restatement(oracle):
    x = _small_inputs()
    rows = to_bed_rows(x["G"])
    n = x["n"]
    gi, gj, gv = x["grm"]
    seen_b = seen_spa = 0
    for eps in (0.001, 0.4):
        ref, route, iters, rc = oracle.mixplus_bed(rows, n, x["R"], x["E"], x["PCs"], gi, gj, gv, epsilon=eps, cutoff=1.5,
                                                   min_maf=0.002)
        assert rc == 0
        for j in range(x["G"].shape[1]):
            if route[j] & 1:
                assert np.isnan(ref[j, 2:]).all()
                continue
            # the allele-frequency model is SPAGxEmix_CCT's (pinned elsewhere): take it from the mix oracle
            _, _, _, mafest, _ = oracle.mix_one_snp("quantitative", x["G"][:, j], x["R"], x["E"], x["R"], np.zeros((n, 0)),
                                                    x["PCs"], epsilon=0.0, min_maf=0.002)
            pspa, pn, pG, S1, VarS1, Z1 = _np_reference_one_snp(x["G"][:, j], x["R"], x["E"], x["PCs"], x["grm"], eps, 1.5, mafest)
            assert np.allclose([ref[j, 7], ref[j, 8], ref[j, 9]], [S1, VarS1, Z1], rtol=1e-10)
            assert np.isclose(ref[j, 6], pG, rtol=1e-9) and np.isclose(ref[j, 5], pn, rtol=1e-8)
            assert abs(np.log10(ref[j, 2]) - np.log10(pspa)) <= 2e-5 * max(1.0, abs(np.log10(pspa))), (j, ref[j, 2], pspa)
            if route[j] & 2:
                seen_b += 1
                assert np.isnan(ref[j, 3]) and np.isnan(ref[j, 4])       # the mixed-model Wald test is the caller's
            else:
                assert ref[j, 3] == ref[j, 2] and ref[j, 4] == ref[j, 2]
            seen_spa += bool(route[j] & 4)
    assert seen_b >= 3 and seen_spa >= 4


def test_mixplus_with_identity_grm_equals_mix_oracle(oracle):
    """GRM = identity: every GRM form collapses to SPAGxEmix_CCT's sums and Var.ratio to 1."""
    x = _small_inputs(seed=9)
    n = x["n"]
    rows = to_bed_rows(x["G"])
    gi = np.arange(n, dtype=np.int32)
    ref, route, _, _ = oracle.mixplus_bed(rows, n, x["R"], x["E"], x["PCs"], gi, gi, np.ones(n), epsilon=1e-12, cutoff=1.5, min_maf=0.002)
    mix, mroute, _, _ = oracle.mix_bed("quantitative", rows, n, x["R"], x["E"], x["R"], np.zeros((n, 0)), x["PCs"], epsilon=1e-12,
                                       cutoff=1.5, min_maf=0.002)
    ok = ~np.isnan(ref[:, 2])
    assert ok.sum() >= 15 and not (route & 2).any()
    assert np.allclose(ref[ok][:, [2, 5, 6, 7, 8, 9, 10, 11]], mix[ok][:, [2, 5, 6, 7, 8, 9, 10, 11]], rtol=1e-9)


def test_host_cct_matches_oracle_bitwise_and_stops(oracle):
    import ctypes as C
    from spagxecct_b200.api import CCT, RStop
    L = oracle.lib()
    o = C.c_double(); w = C.c_int()
    rng = np.random.default_rng(3)
    for _ in range(3000):
        a, b = 10 ** (-rng.uniform(0, 40, 2))
        assert L.orc_cct2(C.c_double(a), C.c_double(b), C.byref(o), C.byref(w)) == 0
        assert o.value == CCT([a, b])
    L.orc_cct2(C.c_double(1.0), C.c_double(0.3), C.byref(o), C.byref(w))
    assert o.value == CCT([1.0, 0.3]) and CCT([0.0, 0.3]) == 0.0
    for bad in ([np.nan, 0.1], [1.2, 0.1], [0.0, 1.0]):
        with pytest.raises(RStop):
            CCT(bad)


def test_mixplus_grm_id_handling():
    from spagxecct_b200.api import RStop, _mixplus_grm
    resid = {"SubjID": ["a", "b", "c"], "Resid": [0.1, 0.2, 0.3]}
    grm = {"ID1": ["a", "b", "c", "b", "z", "z"], "ID2": ["a", "b", "c", "a", "z", "a"], "Value": [1, 1, 1, 0.5, 1, 0.25]}
    gi, gj, gv = _mixplus_grm(resid, grm)
    assert gi.tolist() == [0, 1, 2, 1] and gj.tolist() == [0, 1, 2, 0] and gv.tolist() == [1, 1, 1, 0.5]   # rows with z dropped
    with pytest.raises(RStop, match="does not have GRM information"):
        _mixplus_grm({"SubjID": ["a", "q"]}, grm)


# ---------------------------------------------------------------------------------------------------------------------
def _gpu_inputs(golden_dir):
    from tests.test_gpu_parity import _mix_inputs
    x = _mix_inputs(golden_dir, "binary")
    n = x["n"]
    n_fam = 3000
    rng = np.random.default_rng(17)
    G = x["G"]
    sib = rng.random((n_fam, G.shape[1])) < 0.5
    G[1:2 * n_fam:2] = np.where(sib, G[0:2 * n_fam:2], G[1:2 * n_fam:2])
    x["grm"] = _family_grm(n, n_fam)
    return x


@pytest.mark.gpu
def test_mixplus_cuda_vs_oracle(gpu_ctx, oracle, golden_dir):
    x = _gpu_inputs(golden_dir)
    n = x["n"]
    rows = to_bed_rows(x["G"])
    gi, gj, gv = x["grm"]
    gpu_ctx.set_vectors(x["R"], x["E"])
    gpu_ctx.set_pcs(x["PCs"])
    gpu_ctx.set_sparse_grm(gi, gj, gv)
    geno = gpu_ctx.geno_desc(0, rows.ctypes.data, n, rows.size // ((n + 3) // 4), (n + 3) // 4)
    n_b = 0
    for eps, cutoff in ((0.001, 2.0), (0.2, 1.0)):
        out, diag = gpu_ctx.run_mix_plus(geno, dict(epsilon=eps, cutoff=cutoff, min_maf=0.0002))
        ref, route, iters, rc = oracle.mixplus_bed(rows, n, x["R"], x["E"], x["PCs"], gi, gj, gv, epsilon=eps, cutoff=cutoff,
                                                   min_maf=0.0002, nthreads=8)
        assert rc == 0
        assert_parity(out, ref, P_COLS, S_COLS, exact_cols=(0, 1, 10), label=f"mix_plus eps={eps}")
        assert np.allclose(out[:, 11], ref[:, 11], rtol=1e-14, equal_nan=True)
        bb = (route & 2) > 0
        assert np.isnan(out[bb][:, 3:5]).all()
        a = ~bb & ~np.isnan(ref[:, 2])
        assert (out[a, 3] == out[a, 2]).all() and (out[a, 4] == out[a, 2]).all()
        assert diag["n_branch_b"] == bb.sum() and diag["n_spa"] == ((route & 4) > 0).sum()
        assert diag["n_mix_logistic"] == ((route & 256) > 0).sum() and diag["newton_iters"] == iters.sum()
        n_b += bb.sum()
    assert n_b >= 5 and ((route & 256) > 0).sum() >= 2 and ((route & 4) > 0).sum() >= 8


@pytest.mark.gpu
def test_mixplus_identity_grm_equals_run_mix_on_branch_a(gpu_ctx, golden_dir):
    x = _gpu_inputs(golden_dir)
    n = x["n"]
    rows = to_bed_rows(x["G"])
    geno = gpu_ctx.geno_desc(0, rows.ctypes.data, n, rows.size // ((n + 3) // 4), (n + 3) // 4)
    gpu_ctx.set_vectors(x["R"], x["E"])
    gpu_ctx.set_wald(1, x["Y"], x["Cova"])
    gpu_ctx.set_pcs(x["PCs"])
    prm = dict(epsilon=0.001, cutoff=2.0, min_maf=0.0002)
    mix, _ = gpu_ctx.run_mix(geno, prm)
    gi = np.arange(n, dtype=np.int32)
    gpu_ctx.set_sparse_grm(gi, gi, np.ones(n))
    plus, _ = gpu_ctx.run_mix_plus(geno, prm)
    a = ~np.isnan(mix[:, 6]) & (mix[:, 6] > 0.001)
    assert a.sum() >= 30
    assert np.allclose(plus[a], mix[a], rtol=1e-9)
    with pytest.raises(Exception, match="set_grm"):          # the table alone does not arm SPAGxE_Plus
        gpu_ctx.run_plus(geno, prm)


@pytest.mark.gpu
def test_mixplus_api_wald_callback(gpu_ctx, oracle, golden_dir):
    from spagxecct_b200 import CCT, SPAGxEmix_CCT_Plus, SPAGxEmix_CCT_Plus_one_SNP
    x = _gpu_inputs(golden_dir)
    n = x["n"]
    ids = [f"IID-{i + 1}" for i in range(n)]
    gi, gj, gv = x["grm"]
    resid = {"SubjID": ids, "Resid": x["R"]}
    grm = {"ID1": [ids[i] for i in gi] + ["ghost"], "ID2": [ids[j] for j in gj] + ["ghost"], "Value": list(gv) + [1.0]}
    kw = dict(Geno_mtx=x["G"], ResidMat=resid, E=x["E"], Phen_mtx={"Y": x["Y"], "IID": ids}, Cova_mtx=x["Cova"], topPCs=x["PCs"],
              sparseGRM=grm, epsilon=0.2, Cutoff=2, min_maf=0.0002, ctx=gpu_ctx, quiet=True)
    with pytest.raises(NotImplementedError, match="glmer"):
        SPAGxEmix_CCT_Plus("binary", **kw)
    skip = SPAGxEmix_CCT_Plus("binary", wald="skip", **kw)
    seen = {}

    def wald(j, g):
        assert g.shape == (n,) and not np.isnan(g).any() and g.mean() / 2 <= 0.5 + 1e-12
        seen[j] = True
        return np.nan if j % 2 else 0.01 * (1 + j)

    res = SPAGxEmix_CCT_Plus("binary", wald=wald, **kw)
    bb = np.flatnonzero(skip["p.value.betaG"] <= 0.2)
    assert sorted(seen) == bb.tolist() and bb.size >= 5
    for j in bb:
        pw = skip.values[j, 2] if j % 2 else 0.01 * (1 + j)
        assert res.values[j, 3] == pw and res.values[j, 4] == CCT([skip.values[j, 2], pw])
    rest = np.setdiff1d(np.arange(x["G"].shape[1]), bb)
    assert np.array_equal(res.values[rest], skip.values[rest], equal_nan=True)
    assert res.colnames[2:6] == ["p.value.spaGxEmix.plus", "p.value.spaGxEmix.plus.Wald", "p.value.spaGxEmix.plus.CCT.Wald",
                                 "p.value.normGxEmix.plus"]
    j0 = int(rest[~np.isnan(skip.values[rest, 2])][0])
    one = SPAGxEmix_CCT_Plus_one_SNP("binary", x["G"][:, j0], resid, x["E"], {"Y": x["Y"]}, x["Cova"], x["PCs"], grm, epsilon=0.2,
                                     min_maf=0.0002, ctx=gpu_ctx)
    assert np.allclose(one, skip.values[j0], rtol=1e-12, equal_nan=True)


def test_mixplus_host_logic_with_a_recording_context():
    """CPU: the branch-B completion of the Python mirror (genotype recoding handed to `wald`, NaN -> pval.spaGxE, host CCT,
    refusal without a callback) driven through a stand-in for the device context (no compute)."""
    from spagxecct_b200 import CCT, SPAGxEmix_CCT_Plus

    n, m = 50, 4
    rng = np.random.default_rng(2)
    G = rng.binomial(2, 0.7, (n, m)).astype(float)      # MAF > 0.5: the callback must see the flipped coding
    G[3, 1] = np.nan
    canned = np.full((m, 12), 0.5, order="F")
    canned[:, 6] = [0.5, 1e-5, 1e-9, np.nan]             # p.value.betaG: SNPs 1 and 2 take branch B, SNP 3 is filtered
    canned[1:3, 3:5] = np.nan
    canned[1:3, 2] = [0.02, 0.03]
    canned[3, 2:] = np.nan
    calls = []

    class Ctx:
        def set_vectors(self, R, E): calls.append(("vec", R.size))
        def set_pcs(self, pcs): calls.append(("pcs", pcs.shape))
        def set_sparse_grm(self, gi, gj, gv): calls.append(("grm", gi.tolist(), gj.tolist(), gv.tolist()))
        def run_mix_plus(self, geno, prm): return canned.copy(order="F"), {"n_branch_b": 2}

    ids = [f"s{i}" for i in range(n)]
    resid = {"SubjID": ids, "Resid": rng.standard_normal(n)}
    grm = {"ID1": ids + ["s1", "zz"], "ID2": ids + ["s0", "s0"], "Value": [1.0] * n + [0.5, 0.25]}
    kw = dict(Geno_mtx=G, ResidMat=resid, E=rng.standard_normal(n), Phen_mtx={"Y": np.zeros(n)}, Cova_mtx=np.zeros((n, 1)),
              topPCs=rng.standard_normal((n, 2)), sparseGRM=grm, ctx=Ctx(), quiet=True)
    with pytest.raises(NotImplementedError):
        SPAGxEmix_CCT_Plus("binary", **kw)
    assert calls[2][0] == "grm" and len(calls[2][1]) == n + 1 and calls[2][1][-1] == 1 and calls[2][2][-1] == 0   # the row with "zz" is dropped
    seen = {}

    def wald(j, g):
        seen[j] = g.copy()
        return np.nan if j == 1 else 0.004

    res = SPAGxEmix_CCT_Plus("binary", wald=wald, **kw)
    assert sorted(seen) == [1, 2]
    maf1 = np.nanmean(G[:, 1]) / 2
    g1 = np.where(np.isnan(G[:, 1]), 2 * maf1, G[:, 1])
    assert np.array_equal(seen[1], 2 - g1) and np.array_equal(seen[2], 2 - G[:, 2])
    assert res.values[1, 3] == 0.02 and res.values[1, 4] == CCT([0.02, 0.02])            # NaN -> pval.spaGxE (:389)
    assert res.values[2, 3] == 0.004 and res.values[2, 4] == CCT([0.03, 0.004])
    assert np.array_equal(res.values[[0, 3]], canned[[0, 3]], equal_nan=True)
    with pytest.raises(Exception):
        SPAGxEmix_CCT_Plus("survival", **kw)
