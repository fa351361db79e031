#!/usr/bin/env python
"""Throughput of the other two Step-2 loops on synthetic data shaped like BASELINE.json configs 4 and 5
(not the headline metric; reported in DESIGN.md).  Device-timed through the C ABI with host inputs.
    python tools/bench_variants.py [mix|plus|all]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from spagxecct_b200 import Context, _lib  # noqa: E402
from spagxecct_b200.bedio import pack_rows  # noqa: E402
from spagxecct_b200.harness import get_resid  # noqa: E402


def pheno(n, seed, prev):
    rng = np.random.default_rng(seed)
    X1 = rng.standard_normal(n); X2 = rng.binomial(1, 0.5, n).astype(float); E = rng.standard_normal(n)
    eta = 0.5 * X1 + 0.5 * X2 + 0.5 * E
    Y = rng.binomial(1, 1 / (1 + np.exp(-(np.log(prev / (1 - prev)) - 0.4 + eta)))).astype(float)
    return X1, X2, E, Y


def run_mix(n=200_000, m=2048, k=10):
    rng = np.random.default_rng(1)
    alpha = np.clip(np.r_[rng.normal(0.9, 0.05, n // 2), rng.normal(0.1, 0.05, n - n // 2)], 0, 1)
    pairs = [(0.3, 0.01), (0.1, 0.01), (0.05, 0.4), (0.01, 0.001)]
    G = np.empty((n, m), dtype=np.int8)
    for j in range(m):
        p1, p2 = pairs[j % 4]
        G[:, j] = rng.binomial(2, alpha * p1 + (1 - alpha) * p2)
    # PCs: leading left singular vectors of a standardised genotype block generated the same way
    blk = G[:, :256].astype(np.float64)
    blk = (blk - blk.mean(0)) / (blk.std(0) + 1e-12)
    U, _, _ = np.linalg.svd(blk, full_matrices=False)
    PCs = np.ascontiguousarray(U[:, :k])
    X1, X2, E, Y = pheno(n, 2, 0.05)
    R = get_resid("binary", Y, [X1, X2, PCs[:, 0], PCs[:, 1], E])
    rows = pack_rows(G)
    with Context(0) as ctx:
        ctx.set_vectors(R, E); ctx.set_wald(_lib.TRAIT_BINARY, Y, np.column_stack([X1, X2, PCs[:, 0], PCs[:, 1]]))
        ctx.set_pcs(PCs)
        g = ctx.geno_desc(_lib.GENO_BED, rows.ctypes.data, n, m, (n + 3) // 4)
        ctx.run_mix(g)
        t0 = time.perf_counter(); out, d = ctx.run_mix(g); dt = time.perf_counter() - t0
    return {"loop": "SPAGxEmix_CCT", "N": n, "M": m, "k": k, "snps_per_s_device": m / (d["ms_total"] * 1e-3),
            "snps_per_s_wall": m / dt, "ms_score": d["ms_score"], "ms_tests": d["ms_tests"], "n_spa": d["n_spa"],
            "n_branch_b": d["n_branch_b"], "n_mix_logistic": d["n_mix_logistic"]}


def run_plus(n=100_000, m=16384):
    rng = np.random.default_rng(3)
    nfam = 20_000
    mafs = rng.uniform(0.005, 0.5, m)
    G = (rng.random((n, m)) < mafs).astype(np.int8) + (rng.random((n, m)) < mafs).astype(np.int8)
    idx = np.arange(0, 4 * nfam, 4)
    for o in (2, 3):   # two offspring per family inherit from the two founders
        pick = rng.random((nfam, m)) < 0.5
        G[idx + o] = np.where(pick, G[idx], G[idx + 1])
    X1, X2, E, Y = pheno(n, 4, 0.01)
    R = get_resid("binary", Y, [X1, X2, E])
    gi = [np.arange(n)]; gj = [np.arange(n)]; gv = [np.ones(n)]
    for a, b in ((2, 0), (2, 1), (3, 0), (3, 1), (3, 2)):
        gi.append(idx + a); gj.append(idx + b); gv.append(np.full(nfam, 0.5))
    gi = np.concatenate(gi).astype(np.int32); gj = np.concatenate(gj).astype(np.int32); gv = np.concatenate(gv)

    def quad(x, y):
        return 2 * np.sum(gv * x[gi] * y[gj]) - np.sum(x * y)
    RE = R * E
    rgr = quad(R, R)
    rgre = np.sum(gv * R[gi] * RE[gj]) + np.sum(gv * R[gj] * RE[gi]) - np.sum(RE * R)
    Rnew = (E - rgre / rgr) * R
    rows = pack_rows(G)
    with Context(0) as ctx:
        ctx.set_vectors(R, E); ctx.set_grm(gi, gj, gv, rgr, rgre, quad(Rnew, Rnew), Rnew)
        g = ctx.geno_desc(_lib.GENO_BED, rows.ctypes.data, n, m, (n + 3) // 4)
        ctx.run_plus(g)
        t0 = time.perf_counter(); out, d = ctx.run_plus(g); dt = time.perf_counter() - t0
    return {"loop": "SPAGxE_Plus", "N": n, "M": m, "nnz": int(gi.size), "snps_per_s_device": m / (d["ms_total"] * 1e-3),
            "snps_per_s_wall": m / dt, "ms_score": d["ms_score"], "ms_tests": d["ms_tests"], "n_spa": d["n_spa"],
            "n_branch_b": d["n_branch_b"]}


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("mix", "all"):
        print(json.dumps(run_mix()))
    if which in ("plus", "all"):
        print(json.dumps(run_plus()))
