"""ctypes binding of the CPU oracle (oracle/spagxe_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; the product package
(spagxecct_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "spagxe_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        L = _LIB
        L.orc_pnorm.restype = C.c_double
        L.orc_pnorm.argtypes = [C.c_double, C.c_int]
        L.orc_pcauchy.restype = C.c_double
        L.orc_pcauchy.argtypes = [C.c_double]
        L.orc_pt2.restype = C.c_double
        L.orc_pt2.argtypes = [C.c_double, C.c_double]
        L.orc_na.restype = C.c_double
        L.orc_getprob_spa.restype = C.c_double
        L.orc_grm_quad.restype = C.c_double
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _f(a):
    """column-major double matrix"""
    a = np.asfortranarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def set_mean_mode(mode):
    lib().orc_set_mean_mode(C.c_int(mode))


def set_gmodel(mode):
    """G.model of the outer *_one_SNP functions: 0 Add, 1 Dom, 2 Rec (process-wide, like set_mean_mode)."""
    lib().orc_set_gmodel(C.c_int(mode))


def pnorm(x, lower=True):
    return lib().orc_pnorm(float(x), 1 if lower else 0)


def pcauchy(x):
    return lib().orc_pcauchy(float(x))


def pt2(t, df):
    return lib().orc_pt2(float(t), float(df))


def cct2(p1, p2):
    out = C.c_double()
    warn = C.c_int()
    rc = lib().orc_cct2(C.c_double(p1), C.c_double(p2), C.byref(out), C.byref(warn))
    return rc, out.value, warn.value


def bed_decode(bed_rows, n, pitch=None):
    """bed_rows: uint8 array holding M packed rows -> (n, M) double with NaN=missing."""
    bed_rows = np.ascontiguousarray(bed_rows, dtype=np.uint8).ravel()
    pitch = pitch or (n + 3) // 4
    m = bed_rows.size // pitch
    out = np.empty((n, m), dtype=np.float64, order="F")
    for j in range(m):
        lib().orc_bed_decode_row(C.c_void_p(bed_rows.ctypes.data + j * pitch), C.c_long(n),
                                 C.c_void_p(out.ctypes.data + j * n * 8))
    return out


def bed_counts(bed_rows, n, pitch=None):
    bed_rows = np.ascontiguousarray(bed_rows, dtype=np.uint8).ravel()
    pitch = pitch or (n + 3) // 4
    m = bed_rows.size // pitch
    out = np.zeros((m, 4), dtype=np.int64)
    for j in range(m):
        lib().orc_bed_counts(C.c_void_p(bed_rows.ctypes.data + j * pitch), C.c_long(n),
                             C.c_void_p(out.ctypes.data + j * 32))
    return out


def getprob_spa(R, maf, s, lower_tail, mafvec=None):
    R, Rp = _d(R)
    it = C.c_int()
    if mafvec is not None:
        mv, mvp = _d(mafvec)
    else:
        mvp = None
    p = lib().orc_getprob_spa(Rp, mvp, C.c_double(maf), C.c_long(R.size), C.c_double(s),
                              C.c_int(1 if lower_tail else 0), C.byref(it))
    return p, it.value


def spa_root(R, maf, s, mafvec=None, init_t=0.0):
    R, Rp = _d(R)
    mvp = None
    if mafvec is not None:
        mv, mvp = _d(mafvec)
    root = C.c_double(); it = C.c_int(); cv = C.c_int(); h2 = C.c_double()
    lib().orc_spa_root(Rp, mvp, C.c_double(maf), C.c_long(R.size), C.c_double(s), C.c_double(init_t),
                       C.byref(root), C.byref(it), C.byref(cv), C.byref(h2))
    return root.value, it.value, cv.value, h2.value


def spa_homo(g, R, cutoff=2.0, min_maf=1e-5, g_is_int=False):
    g, gp = _d(g); R, Rp = _d(R)
    out = np.empty(7); info = C.c_int(); its = (C.c_int * 2)()
    lib().orc_spa_homo(gp, C.c_int(int(g_is_int)), Rp, C.c_long(g.size), C.c_double(cutoff), C.c_double(min_maf),
                       out.ctypes.data_as(c_dp), C.byref(info), its)
    return out, info.value, (its[0], its[1])


TRAIT = {"binary": 1, "quantitative": 2, "survival": 3, "categorical": 4}
# survival: Y = np.r_[surv.time, event] (2 n doubles); categorical: Y = index of the level in sort(unique(Y)), 0 .. J-1


def clm(X, ycat, J):
    """ordinal::clm(as.factor(y) ~ X), logit link, flexible thresholds: (coef, se, iterations, rc); thresholds first."""
    X, Xp = _f(np.asarray(X, dtype=np.float64))
    yc = np.ascontiguousarray(ycat, dtype=np.int32)
    n, q = X.shape
    coef = np.empty(J - 1 + q); se = np.empty(J - 1 + q); it = C.c_int()
    rc = lib().orc_clm(Xp, yc.ctypes.data_as(C.c_void_p), C.c_long(n), C.c_int(q), C.c_int(J), coef.ctypes.data_as(c_dp),
                       se.ctypes.data_as(c_dp), C.byref(it))
    return coef, se, it.value, rc


def coxph(X, time, event, maxiter=20):
    """coxph(Surv(time, event) ~ X, ties = "efron") restated from coxph.fit / coxfit6.c: (coef, se, iterations, rc)."""
    X, Xp = _f(np.asarray(X, dtype=np.float64)); time, tp = _d(time); event, ep = _d(event)
    n, q = X.shape
    coef = np.empty(q); se = np.empty(q); it = C.c_int()
    rc = lib().orc_coxph(Xp, tp, ep, C.c_long(n), C.c_int(q), C.c_int(maxiter), coef.ctypes.data_as(c_dp),
                         se.ctypes.data_as(c_dp), C.byref(it))
    return coef, se, it.value, rc


def _threads(nthreads):
    return nthreads or os.cpu_count() or 1


def _shards(m, nthreads):
    nthreads = max(1, min(nthreads, m))
    edges = np.linspace(0, m, nthreads + 1).astype(np.int64)
    return [(int(edges[i]), int(edges[i + 1])) for i in range(nthreads) if edges[i + 1] > edges[i]]


def cct_matrix(trait, G, R, E, Y, Cova, epsilon=0.001, min_maf=0.001, g_is_int=False, nthreads=1):
    """SPAGxE_CCT loop body over an (N, M) genotype matrix (NaN = missing)."""
    G, _ = _f(G)
    n, m = G.shape
    R, Rp = _d(R); E, Ep = _d(E); Y, Yp = _d(Y)
    Cova, Cp = _f(np.asarray(Cova, dtype=np.float64).reshape(n, -1))
    p = Cova.shape[1]
    out = np.empty((m, 10), order="F"); route = np.zeros(m, dtype=np.int32); iters = np.zeros((m, 2), dtype=np.int32)

    def run(lo, hi):
        o = np.empty((hi - lo, 10), order="F")
        rc = lib().orc_cct_matrix(C.c_int(TRAIT[trait]), C.c_void_p(G.ctypes.data + lo * n * 8), C.c_int(int(g_is_int)),
                                  C.c_long(n), C.c_long(hi - lo), Rp, Ep, Yp, Cp, C.c_int(p), C.c_double(epsilon),
                                  C.c_double(min_maf), o.ctypes.data_as(c_dp),
                                  C.c_void_p(route.ctypes.data + lo * 4), C.c_void_p(iters.ctypes.data + lo * 8))
        out[lo:hi] = o
        return rc

    sh = _shards(m, _threads(nthreads) if nthreads != 1 else 1)
    if len(sh) == 1:
        rcs = [run(*sh[0])]
    else:
        with ThreadPoolExecutor(len(sh)) as ex:
            rcs = list(ex.map(lambda a: run(*a), sh))
    return out, route, iters, max(rcs)


def cct_bed(trait, bed_rows, n, R, E, Y, Cova, epsilon=0.001, min_maf=0.001, pitch=None, nthreads=1):
    bed_rows = np.ascontiguousarray(bed_rows, dtype=np.uint8).ravel()
    pitch = pitch or (n + 3) // 4
    m = bed_rows.size // pitch
    R, Rp = _d(R); E, Ep = _d(E); Y, Yp = _d(Y)
    Cova, Cp = _f(np.asarray(Cova, dtype=np.float64).reshape(n, -1))
    p = Cova.shape[1]
    out = np.empty((m, 10), order="F"); route = np.zeros(m, dtype=np.int32); iters = np.zeros((m, 2), dtype=np.int32)

    def run(lo, hi):
        o = np.empty((hi - lo, 10), order="F")
        rc = lib().orc_cct_bed(C.c_int(TRAIT[trait]), C.c_void_p(bed_rows.ctypes.data + lo * pitch), C.c_long(pitch),
                               C.c_long(n), C.c_long(hi - lo), Rp, Ep, Yp, Cp, C.c_int(p), C.c_double(epsilon),
                               C.c_double(min_maf), o.ctypes.data_as(c_dp),
                               C.c_void_p(route.ctypes.data + lo * 4), C.c_void_p(iters.ctypes.data + lo * 8))
        out[lo:hi] = o
        return rc

    sh = _shards(m, _threads(nthreads) if nthreads != 1 else 1)
    if len(sh) == 1:
        rcs = [run(*sh[0])]
    else:
        with ThreadPoolExecutor(len(sh)) as ex:
            rcs = list(ex.map(lambda a: run(*a), sh))
    return out, route, iters, max(rcs)


def mix_bed(trait, bed_rows, n, R, E, Y, Cova, PCs, epsilon=0.001, cutoff=2.0, min_maf=0.001,
            pc_pcut=0.05, neg_ratio_cut=0.1, pitch=None, nthreads=1):
    bed_rows = np.ascontiguousarray(bed_rows, dtype=np.uint8).ravel()
    pitch = pitch or (n + 3) // 4
    m = bed_rows.size // pitch
    R, Rp = _d(R); E, Ep = _d(E); Y, Yp = _d(Y)
    Cova, Cp = _f(np.asarray(Cova, dtype=np.float64).reshape(n, -1))
    PCs, Pp = _f(np.asarray(PCs, dtype=np.float64).reshape(n, -1))
    p = Cova.shape[1]; k = PCs.shape[1]
    out = np.empty((m, 12), order="F"); route = np.zeros(m, dtype=np.int32); iters = np.zeros((m, 2), dtype=np.int32)

    def run(lo, hi):
        o = np.empty((hi - lo, 12), order="F")
        rc = lib().orc_mix_bed(C.c_int(TRAIT[trait]), C.c_void_p(bed_rows.ctypes.data + lo * pitch), C.c_long(pitch),
                               C.c_long(n), C.c_long(hi - lo), Rp, Ep, Yp, Cp, C.c_int(p), Pp, C.c_int(k),
                               C.c_double(epsilon), C.c_double(cutoff), C.c_double(min_maf), C.c_double(pc_pcut),
                               C.c_double(neg_ratio_cut), o.ctypes.data_as(c_dp),
                               C.c_void_p(route.ctypes.data + lo * 4), C.c_void_p(iters.ctypes.data + lo * 8))
        out[lo:hi] = o
        return rc

    sh = _shards(m, _threads(nthreads) if nthreads != 1 else 1)
    if len(sh) == 1:
        rcs = [run(*sh[0])]
    else:
        with ThreadPoolExecutor(len(sh)) as ex:
            rcs = list(ex.map(lambda a: run(*a), sh))
    return out, route, iters, max(rcs)


def mix_one_snp(trait, g, R, E, Y, Cova, PCs, epsilon=0.001, cutoff=2.0, min_maf=0.001, pc_pcut=0.05,
                neg_ratio_cut=0.1, g_is_int=False):
    g, gp = _d(g); n = g.size
    R, Rp = _d(R); E, Ep = _d(E); Y, Yp = _d(Y)
    Cova, Cp = _f(np.asarray(Cova, dtype=np.float64).reshape(n, -1))
    PCs, Pp = _f(np.asarray(PCs, dtype=np.float64).reshape(n, -1))
    out = np.empty(12); route = C.c_int(); its = (C.c_int * 2)(); mafest = np.empty(n)
    rc = lib().orc_mix_one_snp(C.c_int(TRAIT[trait]), gp, C.c_int(int(g_is_int)), Rp, Ep, Yp, Cp, C.c_int(Cova.shape[1]),
                               Pp, C.c_int(PCs.shape[1]), C.c_long(n), C.c_double(epsilon), C.c_double(cutoff),
                               C.c_double(min_maf), C.c_double(pc_pcut), C.c_double(neg_ratio_cut),
                               out.ctypes.data_as(c_dp), C.byref(route), its, mafest.ctypes.data_as(c_dp))
    return out, route.value, (its[0], its[1]), mafest, rc


def mixplus_bed(bed_rows, n, R, E, PCs, grm_i, grm_j, grm_v, epsilon=0.001, cutoff=2.0, min_maf=0.001, pc_pcut=0.05,
                neg_ratio_cut=0.1, pitch=None, nthreads=1):
    """SPAGxEmix_CCT_Plus loop (ref R/SPAmixGxECCTPlus_functions_MYZ_2024-09-17.R:100-127 -> :144-427) over packed rows.
    Branch-B rows carry NA in the Wald / CCT columns (the mixed-model Wald test is the caller's, see spagxe_oracle.c)."""
    bed_rows = np.ascontiguousarray(bed_rows, dtype=np.uint8).ravel()
    pitch = pitch or (n + 3) // 4
    m = bed_rows.size // pitch
    R, Rp = _d(R); E, Ep = _d(E)
    PCs, Pp = _f(np.asarray(PCs, dtype=np.float64).reshape(n, -1))
    k = PCs.shape[1]
    gi = np.ascontiguousarray(grm_i, dtype=np.int32); gj = np.ascontiguousarray(grm_j, dtype=np.int32)
    gv, gvp = _d(grm_v)
    out = np.empty((m, 12), order="F"); route = np.zeros(m, dtype=np.int32); iters = np.zeros((m, 2), dtype=np.int32)

    def run(lo, hi):
        o = np.empty((hi - lo, 12), order="F")
        rc = lib().orc_mixplus_bed(C.c_void_p(bed_rows.ctypes.data + lo * pitch), C.c_long(pitch), C.c_long(n),
                                   C.c_long(hi - lo), Rp, Ep, Pp, C.c_int(k), C.c_long(gi.size), gi.ctypes.data_as(c_ip),
                                   gj.ctypes.data_as(c_ip), gvp, C.c_double(epsilon), C.c_double(cutoff),
                                   C.c_double(min_maf), C.c_double(pc_pcut), C.c_double(neg_ratio_cut),
                                   o.ctypes.data_as(c_dp), C.c_void_p(route.ctypes.data + lo * 4),
                                   C.c_void_p(iters.ctypes.data + lo * 8))
        out[lo:hi] = o
        return rc

    sh = _shards(m, _threads(nthreads) if nthreads != 1 else 1)
    if len(sh) == 1:
        rcs = [run(*sh[0])]
    else:
        with ThreadPoolExecutor(len(sh)) as ex:
            rcs = list(ex.map(lambda a: run(*a), sh))
    return out, route, iters, max(rcs)


def plus_one_snp(g, R, E, Rnew, R_GRM_R, R_GRM_RE, Rnew_GRM_Rnew, grm_i, grm_j, grm_v, epsilon=0.001, cutoff=2.0, min_maf=0.001,
                 g_is_int=False):
    g, gp = _d(g); n = g.size
    R, Rp = _d(R); E, Ep = _d(E); Rnew, Rnp = _d(Rnew)
    gi = np.ascontiguousarray(grm_i, dtype=np.int32); gj = np.ascontiguousarray(grm_j, dtype=np.int32)
    gv, gvp = _d(grm_v)
    out = np.empty(8); route = C.c_int(); its = (C.c_int * 2)()
    lib().orc_plus_one_snp(gp, C.c_int(int(g_is_int)), Rp, Ep, Rnp, C.c_double(R_GRM_R), C.c_double(R_GRM_RE),
                           C.c_double(Rnew_GRM_Rnew), C.c_long(gi.size), gi.ctypes.data_as(c_ip), gj.ctypes.data_as(c_ip), gvp,
                           C.c_long(n), C.c_double(epsilon), C.c_double(cutoff), C.c_double(min_maf), out.ctypes.data_as(c_dp),
                           C.byref(route), its)
    return out, route.value, (its[0], its[1])


def mixplus_one_snp(g, R, E, PCs, grm_i, grm_j, grm_v, epsilon=0.001, cutoff=2.0, min_maf=0.001, pc_pcut=0.05,
                    neg_ratio_cut=0.1, g_is_int=False):
    g, gp = _d(g); n = g.size
    R, Rp = _d(R); E, Ep = _d(E)
    PCs, Pp = _f(np.asarray(PCs, dtype=np.float64).reshape(n, -1))
    gi = np.ascontiguousarray(grm_i, dtype=np.int32); gj = np.ascontiguousarray(grm_j, dtype=np.int32)
    gv, gvp = _d(grm_v)
    out = np.empty(12); route = C.c_int(); its = (C.c_int * 2)()
    lib().orc_mixplus_one_snp(gp, C.c_int(int(g_is_int)), Rp, Ep, Pp, C.c_int(PCs.shape[1]), C.c_long(n), C.c_long(gi.size),
                              gi.ctypes.data_as(c_ip), gj.ctypes.data_as(c_ip), gvp, C.c_double(epsilon), C.c_double(cutoff),
                              C.c_double(min_maf), C.c_double(pc_pcut), C.c_double(neg_ratio_cut), out.ctypes.data_as(c_dp),
                              C.byref(route), its)
    return out, route.value, (its[0], its[1])


def localance_matrix(trait, G, H, R, E, Y, Cova, cova_haplo, epsilon=0.001, min_maf=0.0001):
    """SPAGxEmixCCT_localance loop (ref R/SPAGxECCT.R:1443-1476) over (N, M) matrices: G ancestry-specific genotypes,
    H local ancestry counts, cova_haplo a list of (N, M) matrices (Cova.haplo.mtx.list).  OpenMP over SNPs."""
    G, Gp = _f(G); H, Hp = _f(H)
    n, m = G.shape
    R, Rp = _d(R); E, Ep = _d(E); Y, Yp = _d(Y)
    Cova, Cp = _f(np.asarray(Cova, dtype=np.float64).reshape(n, -1))
    chs = [_f(c)[0] for c in cova_haplo]
    arr = (C.c_void_p * max(len(chs), 1))(*[c.ctypes.data for c in chs])
    out = np.empty((m, 10), order="F"); route = np.zeros(m, dtype=np.int32); iters = np.zeros((m, 2), dtype=np.int32)
    rc = lib().orc_localance_matrix(C.c_int(TRAIT[trait]), Gp, Hp, C.c_long(n), C.c_long(m), Rp, Ep, Yp, Cp,
                                    C.c_int(Cova.shape[1]), arr, C.c_int(len(chs)), C.c_double(epsilon), C.c_double(min_maf),
                                    out.ctypes.data_as(c_dp), C.c_void_p(route.ctypes.data), C.c_void_p(iters.ctypes.data))
    return out, route, iters, rc


def plus_nullmodel_scalars(grm_i, grm_j, grm_v, R, E):
    gi = np.ascontiguousarray(grm_i, dtype=np.int32); gj = np.ascontiguousarray(grm_j, dtype=np.int32)
    gv, gvp = _d(grm_v); R, Rp = _d(R); E, Ep = _d(E)
    n = R.size
    Rnew = np.empty(n); a = C.c_double(); b = C.c_double(); c = C.c_double()
    lib().orc_plus_nullmodel_scalars(C.c_long(gi.size), gi.ctypes.data_as(c_ip), gj.ctypes.data_as(c_ip), gvp, Rp, Ep,
                                     C.c_long(n), C.byref(a), C.byref(b), Rnew.ctypes.data_as(c_dp), C.byref(c))
    return a.value, b.value, Rnew, c.value


def plus_bed(bed_rows, n, R, E, Rnew, R_GRM_R, R_GRM_RE, Rnew_GRM_Rnew, grm_i, grm_j, grm_v,
             epsilon=0.001, cutoff=2.0, min_maf=0.001, pitch=None, nthreads=1):
    bed_rows = np.ascontiguousarray(bed_rows, dtype=np.uint8).ravel()
    pitch = pitch or (n + 3) // 4
    m = bed_rows.size // pitch
    gi = np.ascontiguousarray(grm_i, dtype=np.int32); gj = np.ascontiguousarray(grm_j, dtype=np.int32)
    gv, gvp = _d(grm_v); R, Rp = _d(R); E, Ep = _d(E); Rnew, Rnp = _d(Rnew)
    out = np.empty((m, 8), order="F"); route = np.zeros(m, dtype=np.int32); iters = np.zeros((m, 2), dtype=np.int32)

    def run(lo, hi):
        o = np.empty((hi - lo, 8), order="F")
        lib().orc_plus_bed(C.c_void_p(bed_rows.ctypes.data + lo * pitch), C.c_long(pitch), C.c_long(n), C.c_long(hi - lo),
                           Rp, Ep, Rnp, C.c_double(R_GRM_R), C.c_double(R_GRM_RE), C.c_double(Rnew_GRM_Rnew),
                           C.c_long(gi.size), gi.ctypes.data_as(c_ip), gj.ctypes.data_as(c_ip), gvp,
                           C.c_double(epsilon), C.c_double(cutoff), C.c_double(min_maf), o.ctypes.data_as(c_dp),
                           C.c_void_p(route.ctypes.data + lo * 4), C.c_void_p(iters.ctypes.data + lo * 8))
        out[lo:hi] = o
        return 0

    sh = _shards(m, _threads(nthreads) if nthreads != 1 else 1)
    if len(sh) == 1:
        run(*sh[0])
    else:
        with ThreadPoolExecutor(len(sh)) as ex:
            list(ex.map(lambda a: run(*a), sh))
    return out, route, iters, 0


def glm_binomial(X, y):
    X, Xp = _f(X); y, yp = _d(y)
    n, q = X.shape
    coef = np.empty(q); se = np.empty(q); it = C.c_int(); cv = C.c_int(); mu = np.empty(n)
    rc = lib().orc_glm_binomial(Xp, yp, C.c_long(n), C.c_int(q), coef.ctypes.data_as(c_dp), se.ctypes.data_as(c_dp),
                                C.byref(it), C.byref(cv), mu.ctypes.data_as(c_dp))
    return rc, coef, se, it.value, cv.value, mu


def lm(X, y):
    X, Xp = _f(X); y, yp = _d(y)
    n, q = X.shape
    coef = np.empty(q); se = np.empty(q); pv = np.empty(q); fit = np.empty(n)
    rc = lib().orc_lm(Xp, yp, C.c_long(n), C.c_int(q), coef.ctypes.data_as(c_dp), se.ctypes.data_as(c_dp),
                      pv.ctypes.data_as(c_dp), fit.ctypes.data_as(c_dp))
    return rc, coef, se, pv, fit
