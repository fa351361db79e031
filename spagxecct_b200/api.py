"""Host-side mirror of the reference's exported Step-2 functions.

Same names, argument meaning, printed messages, output columns and error behaviour as the R
closures (R is not available in the build image, so the Python mirror is what the parity tests
drive; the R wrappers that bind the very same C ABI are in spagxecct_b200/R/).

    SPAGxE_CCT            R/SPAGxECCT.R:455-530
    SPAGxE_CCT_one_SNP    R/SPAGxECCT.R:543-683
    SPAGxEmix_CCT(_one_SNP)  R/SPAGxECCT.R:916-1002, :1013-1269
    SPAGxE_Plus(_one_SNP)    R/SPAGxEPlus_functions_MYZ.R:159-238, :248-405
    SPAGxEmix_CCT_Plus(_one_SNP)  R/SPAmixGxECCTPlus_functions_MYZ_2024-09-17.R:38-133, :144-427
    check_input_Resid     R/SPAGxECCT.R:1900-1913

R's `.` in argument names becomes `_` (Geno.mtx -> Geno_mtx).  Phen_mtx is any mapping/DataFrame
with the columns the reference reads (`ID`, `Y`).  Geno_mtx is an (N, M) array (NaN / INT_MIN =
NA) or a pandas DataFrame whose index/columns play the role of dimnames.
"""
import datetime
import os

import numpy as np

from . import _lib
from .bedio import BedFile

COLS_CCT = ["MAF", "missing.rate", "p.value.spaGxE", "p.value.spaGxE.Wald", "p.value.spaGxE.CCT.Wald",
            "p.value.normGxE", "p.value.betaG", "Stat.betaG", "Var.betaG", "z.betaG"]
COLS_MIX = COLS_CCT + ["MAF.est.negative.num", "MAC"]
COLS_PLUS = ["MAF", "missing.rate", "p.value.spaGxE.plus", "p.value.normGxE.plus", "p.value.betaG", "Stat.betaG",
             "Var.betaG", "z.betaG"]

_TRAITS = {"binary": _lib.TRAIT_BINARY, "quantitative": _lib.TRAIT_QUANTITATIVE, "survival": _lib.TRAIT_SURVIVAL,
           "categorical": _lib.TRAIT_CATEGORICAL}


def _trait_code(traits):
    if traits not in _TRAITS:
        raise RStop(f'traits="{traits}": expected "survival", "binary", "quantitative" or "categorical"')
    return _TRAITS[traits]


def _wald_y(traits, Phen_mtx):
    """Phen.mtx$Y as the C ABI takes it; categorical: as.factor(Y) (ref R/SPAGxECCT.R:666) = index into sort(unique(Y))."""
    y = _col(Phen_mtx, "Y")
    if _trait_code(traits) == _lib.TRAIT_CATEGORICAL:
        return np.unique(np.asarray(y), return_inverse=True)[1].astype(np.float64)
    return y


def _set_wald(ctx, traits, Phen_mtx, Cova_mtx):
    """Wald data of branch B (ref R/SPAGxECCT.R:622-677): Phen.mtx$Y, or $surv.time / $event for a survival trait."""
    cova = np.asarray(Cova_mtx, dtype=np.float64)
    if _trait_code(traits) == _lib.TRAIT_SURVIVAL:
        ctx.set_wald_survival(_col(Phen_mtx, "surv.time"), _col(Phen_mtx, "event"), cova)
    else:
        ctx.set_wald(_trait_code(traits), _wald_y(traits, Phen_mtx), cova)


class RStop(Exception):
    """The reference would have called stop() with this message."""


class Result:
    """M x k named matrix (what the R functions return)."""

    def __init__(self, values, rownames, colnames, diag=None):
        self.values = values
        self.rownames = list(rownames) if rownames is not None else None
        self.colnames = list(colnames)
        self.diag = diag

    def __getitem__(self, col):
        return self.values[:, self.colnames.index(col)]

    def to_frame(self):
        import pandas as pd
        return pd.DataFrame(self.values, index=self.rownames, columns=self.colnames)


def _col(df, name):
    try:
        v = df[name]
    except Exception:
        return None
    return None if v is None else np.asarray(v)


def check_input_Resid(pIDs, gIDs, R, range_=(-100, 100)):
    """R/SPAGxECCT.R:1900-1913 (same stop() messages)."""
    if pIDs is None and gIDs is None:
        raise RStop("Arguments 'pIDs' and 'gIDs' are required in case of potential errors. For more information, "
                    "please refer to 'Details'.")
    p = [] if pIDs is None else [str(x) for x in pIDs]
    g = [] if gIDs is None else [str(x) for x in gIDs]
    # `any(sort(unique(pIDs)) != sort(unique(gIDs)))` (:1904): with rownames(Geno.mtx) NULL the comparison is
    # against logical(0) and any() is FALSE, so R passes silently; a length mismatch of two non-empty sets recycles
    # with a warning and (practically always) trips the test, which is the behaviour kept here
    if gIDs is not None and pIDs is not None and sorted(set(p)) != sorted(set(g)):
        raise RStop("unique(pIDs) should be the same as unique(gIDs).")
    if len(set(g)) != len(g):
        raise RStop("Argument 'gIDs' should not have a duplicated element.")
    if range_[1] != -1 * range_[0]:
        raise RStop("range[2] should be -1*range[1]")
    if len(R) != len(p):
        raise RStop("Argument 'pIDs' should be of the same length as input data.")
    if gIDs is None or p == g:
        return None
    pos = {s: i for i, s in enumerate(g)}
    return np.array([pos[s] for s in p])


def _split_geno(Geno_mtx):
    """-> (ndarray (N, M), rownames or None, colnames or None)"""
    try:
        import pandas as pd
        if isinstance(Geno_mtx, pd.DataFrame):
            return Geno_mtx.to_numpy(), [str(x) for x in Geno_mtx.index], [str(x) for x in Geno_mtx.columns]
    except ImportError:
        pass
    if isinstance(Geno_mtx, tuple):
        return np.asarray(Geno_mtx[0]), Geno_mtx[1], Geno_mtx[2]
    return np.asarray(Geno_mtx), None, None


def _geno_desc_from_matrix(G):
    if G.dtype.kind in "iu":
        G = np.asfortranarray(G, dtype=np.int32)
        kind = _lib.GENO_I32
    else:
        G = np.asfortranarray(G, dtype=np.float64)
        kind = _lib.GENO_F64
    n, m = G.shape
    return G, _lib.Context.geno_desc(kind, G.ctypes.data, n, m, n)


_CONTROL_KEYS = {"AllMarkers", "IDsToIncludeFile", "IDsToExcludeFile", "RangesToIncludeFile", "RangesToExcludeFile",
                 "AlleleOrder", "ImputeMethod"}


def _select_markers(bf, control):
    """GRAB.ReadGeno's marker selection (`control`): marker ids from IDsToIncludeFile / IDsToExcludeFile (one id per
    line) and chromosome ranges from RangesToIncludeFile / RangesToExcludeFile (CHROM START END per line).  Include
    files are united, then exclude files are subtracted, as in GRAB; file order of the markers is kept."""
    m = bf.m
    has_inc = "IDsToIncludeFile" in control or "RangesToIncludeFile" in control
    has_exc = "IDsToExcludeFile" in control or "RangesToExcludeFile" in control
    if has_inc and has_exc:
        raise RStop("We currently do not support both 'IncludeFile' and 'ExcludeFile' at the same time.")
    if not (has_inc or has_exc):
        if not control.get("AllMarkers", False):
            raise RStop("If 'AllMarkers' is FALSE, at least one of the include / exclude files should be specified.")
        return None
    ids = np.asarray(bf.snp_ids)
    chrom, pos = bf.bim_chrom_pos()
    hit = np.zeros(m, dtype=bool)
    for key in ("IDsToIncludeFile", "IDsToExcludeFile"):
        if key in control:
            want = {l.strip() for l in open(control[key]) if l.strip()}
            hit |= np.isin(ids, list(want))
    for key in ("RangesToIncludeFile", "RangesToExcludeFile"):
        if key in control:
            for l in open(control[key]):
                f = l.split()
                if len(f) >= 3:
                    hit |= (chrom == f[0]) & (pos >= int(f[1])) & (pos <= int(f[2]))
    keep = hit if has_inc else ~hit
    return np.nonzero(keep)[0]


def _read_geno_file(GenoFile, GenoFileIndex, sample_ids, control):
    """The GRAB.ReadGeno step for PLINK input (ref R/SPAGxECCT.R:484-487) without materialising the N x M matrix:
    returns (geno descriptor, objects to keep alive, N, M, snp ids, row ids).  The .bed file is memory-mapped and
    handed to the C ABI as it is (the library streams it through pinned staging buffers in ~1 GiB chunks, so a file larger
    than RAM is fine); the SampleIDs subset / reorder is a device-side gather (spagxe_geno.sample_index); the allele
    order is a flag; only a marker selection copies the selected rows on the host."""
    control = dict(control) if control else {"AllMarkers": True}
    unknown = set(control) - _CONTROL_KEYS
    if unknown:
        raise NotImplementedError(f"control option(s) {sorted(unknown)} are not supported")
    if control.get("ImputeMethod", "none") != "none":
        raise NotImplementedError('control$ImputeMethod other than "none": the Step-2 functions impute themselves (impute.method)')
    is_bgen = str(GenoFile).endswith(".bgen")
    order = control.get("AlleleOrder", "ref-first" if is_bgen else "alt-first")   # GRAB's defaults: PLINK counts A1, BGEN the second allele
    if order not in ("alt-first", "ref-first"):
        raise RStop("control$AlleleOrder should be 'ref-first' or 'alt-first'.")
    if is_bgen:
        return _read_bgen_file(GenoFile, GenoFileIndex, sample_ids, control, order)
    if not str(GenoFile).endswith(".bed"):
        raise NotImplementedError("GenoFile must be a PLINK .bed or a .bgen file")
    bim = fam = None
    if GenoFileIndex is not None:
        bim, fam = GenoFileIndex[0], GenoFileIndex[1]
    bf = BedFile(GenoFile, bim, fam)
    rows = bf.rows
    snp_ids = bf.snp_ids
    sel = _select_markers(bf, control)
    m = bf.m
    if sel is not None:
        rows = np.ascontiguousarray(np.asarray(rows).reshape(bf.m, bf.row_bytes)[sel]).ravel()
        snp_ids = [bf.snp_ids[j] for j in sel]
        m = len(sel)
    ids = bf.sample_ids
    sidx = None
    if sample_ids is not None:
        want = [str(x) for x in sample_ids]
        if want != ids:
            pos = {s: i for i, s in enumerate(ids)}
            missing = [s for s in want if s not in pos]
            if missing:
                raise RStop(f"SampleIDs not found in the .fam file: {missing[:3]}...")
            sidx = np.ascontiguousarray([pos[s] for s in want], dtype=np.int64)
            ids = want
    n = len(ids)
    geno = _lib.Context.geno_desc(_lib.GENO_BED, rows.ctypes.data, n, m, bf.row_bytes, sample_index=sidx,
                                  n_file_samples=bf.n, count_a2=(order == "ref-first"))
    return geno, (rows, sidx, bf), n, m, snp_ids, ids


def _read_bgen_file(GenoFile, GenoFileIndex, sample_ids, control, order):
    """GRAB.ReadGeno for BGEN input (GenoFileIndex = c(bgiFile, sampleFile); the .bgi index is not needed here): the
    probabilities are decoded on the host into the N x M dosage matrix (spagxecct_b200/bgenio.py); hard-call-valued
    dosages then take the 2-bit pipeline, imputed ones the dense dosage kernel of SPAGxE_CCT."""
    from .bgenio import BgenFile
    sample_path = None
    if GenoFileIndex is not None and len(GenoFileIndex) > 1:
        sample_path = GenoFileIndex[1]
    elif os.path.exists(str(GenoFile)[:-5] + ".sample"):
        sample_path = str(GenoFile)[:-5] + ".sample"
    bf = BgenFile(GenoFile, sample_path)
    sel = _select_markers(bf, control)
    D = bf.read_dosages(sel, allele_order=order)
    snp_ids = bf.snp_ids if sel is None else [bf.snp_ids[j] for j in sel]
    ids = bf.sample_ids
    if sample_ids is not None:
        want = [str(x) for x in sample_ids]
        if want != ids:
            pos = {s_: i for i, s_ in enumerate(ids)}
            missing = [s_ for s_ in want if s_ not in pos]
            if missing:
                raise RStop(f"SampleIDs not found in the BGEN sample list: {missing[:3]}...")
            D = np.asfortranarray(D[[pos[s_] for s_ in want], :])
            ids = want
    n, m = D.shape
    geno = _lib.Context.geno_desc(_lib.GENO_F64, D.ctypes.data, n, m, n)
    return geno, (D, bf), n, m, snp_ids, ids


def _now():
    return datetime.datetime.now().strftime('[1] "%Y-%m-%d %H:%M:%S"')


def _prep_common(ctx, R, E):
    R = np.ascontiguousarray(R, dtype=np.float64).ravel()
    E = np.ascontiguousarray(E, dtype=np.float64).ravel()
    ctx.set_vectors(R, E)
    return R, E


class _MultiGpu:
    """A _lib.MultiContext behind the single-GPU Context's method names, for the loops that take `n_gpus` (the SNPs are
    cut into contiguous ranges over the GPUs of the node; R / E / Y / Cova reach them by one NCCL broadcast)."""

    def __init__(self, n_gpus):
        self.mg = _lib.MultiContext(n_gpus)
        self._vec = None
        self._sent = False

    def close(self):
        self.mg.close()

    def set_vectors(self, R, E):
        self._vec = (R, E); self._sent = False

    def _send(self, trait=0, Y=None, cova=None):
        self.mg.set_inputs(self._vec[0], self._vec[1], trait, Y, cova)
        self._sent = True

    def set_wald(self, trait, Y, cova):
        self._send(int(trait), Y, cova)

    def set_wald_survival(self, surv_time, event, cova):
        self._send()
        self.mg.set_wald_survival(surv_time, event, cova)

    def _ready(self):
        if not self._sent:
            self._send()

    def set_pcs(self, pcs):
        self._ready(); self.mg.set_pcs(pcs)

    def set_grm(self, *a):
        self._ready(); self.mg.set_grm(*a)

    def set_sparse_grm(self, *a):
        self._ready(); self.mg.set_sparse_grm(*a)

    def run_mix(self, geno, prm):
        return self.mg.run_mix(geno, prm)

    def run_plus(self, geno, prm):
        return self.mg.run_plus(geno, prm)

    def run_mix_plus(self, geno, prm):
        return self.mg.run_mix_plus(geno, prm)


def _make_ctx(ctx, device, n_gpus):
    if ctx is not None:
        return ctx
    return _MultiGpu(n_gpus) if n_gpus > 1 else _lib.Context(device)


def _params(epsilon, Cutoff, impute_method, missing_cutoff, min_maf, G_model, **kw):
    if impute_method != "fixed":
        raise NotImplementedError('only impute.method = "fixed" exists in the reference')
    # ref R/SPAGxECCT.R:572-574: any other string leaves g untouched, i.e. behaves like "Add"
    g_model = {"Dom": 1, "Rec": 2}.get(G_model, 0)
    return _lib.make_params(epsilon=epsilon, cutoff=Cutoff, min_maf=min_maf, missing_cutoff=missing_cutoff,
                            g_model=g_model, **kw)


def SPAGxE_CCT(traits="survival/binary/quantitative/categorical", GenoFile=None, GenoFileIndex=None,
               control=None, Geno_mtx=None, R=None, E=None, Phen_mtx=None, Cova_mtx=None, epsilon=0.001, Cutoff=2,
               impute_method="fixed", missing_cutoff=0.15, min_maf=0.001, G_model="Add", ctx=None, device=0,
               quiet=False, n_gpus=1):
    """Drop-in for SPAGxE_CCT (R/SPAGxECCT.R:455-530): returns an M x 10 named matrix.
    n_gpus > 1 (an addition to the reference's formals, like `n.gpus` of the R wrapper): the SNPs are cut into contiguous
    ranges over that many GPUs of the node (spagxe_mgpu_*: one NCCL broadcast of R/E/Y/Cova, no other collective)."""
    if _trait_code(traits) == _lib.TRAIT_SURVIVAL and n_gpus > 1:
        raise NotImplementedError('traits="survival" runs on one GPU (spagxe_mgpu_set_inputs carries a single Y vector)')
    if n_gpus > 1 and ctx is None:
        return _spagxe_cct_multi(traits, GenoFile, GenoFileIndex, control, Geno_mtx, R, E, Phen_mtx, Cova_mtx, epsilon, Cutoff,
                                 impute_method, missing_cutoff, min_maf, G_model, quiet, n_gpus)
    own = ctx is None
    ctx = ctx or _lib.Context(device)
    try:
        pIDs = _col(Phen_mtx, "ID")
        keep = None
        if Geno_mtx is None:
            geno, keep, n, m, snp_ids, gIDs = _read_geno_file(GenoFile, GenoFileIndex, pIDs, control)
        else:
            G, gIDs, snp_ids = _split_geno(Geno_mtx)
            keep, geno = _geno_desc_from_matrix(G)
            n, m = G.shape
        check_input_Resid(pIDs, gIDs, R)
        if not quiet:
            print(f'[1] "Sample size is {n}."')
            print(f'[1] "Number of variants is {m}."')
        _prep_common(ctx, R, E)
        _set_wald(ctx, traits, Phen_mtx, Cova_mtx)
        if not quiet:
            print('[1] "Start Analyzing..."')
            print(_now())
        out, diag = ctx.run_cct(geno, _params(epsilon, Cutoff, impute_method, missing_cutoff, min_maf, G_model))
        del keep
        if not quiet:
            print('[1] "Analysis Complete."')
            print(_now())
        return Result(out, snp_ids, COLS_CCT, diag)
    finally:
        if own:
            ctx.close()


def _spagxe_cct_multi(traits, GenoFile, GenoFileIndex, control, Geno_mtx, R, E, Phen_mtx, Cova_mtx, epsilon, Cutoff,
                      impute_method, missing_cutoff, min_maf, G_model, quiet, n_gpus):
    with _lib.MultiContext(n_gpus) as mg:
        pIDs = _col(Phen_mtx, "ID")
        if Geno_mtx is None:
            geno, keep, n, m, snp_ids, gIDs = _read_geno_file(GenoFile, GenoFileIndex, pIDs, control)
        else:
            G, gIDs, snp_ids = _split_geno(Geno_mtx)
            keep, geno = _geno_desc_from_matrix(G)
            n, m = G.shape
        check_input_Resid(pIDs, gIDs, R)
        if not quiet:
            print(f'[1] "Sample size is {n}."')
            print(f'[1] "Number of variants is {m}."')
            print('[1] "Start Analyzing..."')
            print(_now())
        mg.set_inputs(R, E, _trait_code(traits), _wald_y(traits, Phen_mtx), np.asarray(Cova_mtx, dtype=np.float64))
        out, diag = mg.run_cct(geno, _params(epsilon, Cutoff, impute_method, missing_cutoff, min_maf, G_model))
        del keep
        if not quiet:
            print('[1] "Analysis Complete."')
            print(_now())
        return Result(out, snp_ids, COLS_CCT, diag)


def SPAGxE_CCT_one_SNP(traits, g, R, E, Phen_mtx, Cova_mtx, epsilon=0.001, Cutoff=2, impute_method="fixed",
                       missing_cutoff=0.15, min_maf=0.001, G_model="Add", ctx=None, device=0):
    """Drop-in for SPAGxE_CCT_one_SNP (R/SPAGxECCT.R:543-683): numeric(10). No ID checks, as in R."""
    own = ctx is None
    ctx = ctx or _lib.Context(device)
    try:
        g = np.asarray(g)
        G, geno = _geno_desc_from_matrix(g.reshape(-1, 1))
        _prep_common(ctx, R, E)
        _set_wald(ctx, traits, Phen_mtx, Cova_mtx)
        out, _ = ctx.run_cct(geno, _params(epsilon, Cutoff, impute_method, missing_cutoff, min_maf, G_model))
        return out[0].copy()
    finally:
        if own:
            ctx.close()


def SPAGxEmix_CCT(traits="survival/binary/quantitative/categorical", GenoFile=None, GenoFileIndex=None,
                  control=None, Geno_mtx=None, R=None, E=None, Phen_mtx=None, Cova_mtx=None, topPCs=None,
                  epsilon=0.001, Cutoff=2, impute_method="fixed", missing_cutoff=0.15, min_maf=0.001, G_model="Add",
                  topPCs_pvalue_cutoff=0.05, MAF_est_negative_ratio_cutoff=0.1, ctx=None, device=0, quiet=False, n_gpus=1):
    """Drop-in for SPAGxEmix_CCT (R/SPAGxECCT.R:916-1002): M x 12. (No ID check: commented out at :955.)
    n_gpus > 1 (an addition to the reference's formals): contiguous SNP ranges over the GPUs of the node."""
    _trait_code(traits)
    own = ctx is None
    ctx = _make_ctx(ctx, device, n_gpus)
    try:
        if Geno_mtx is None:
            geno, keep, n, m, snp_ids, _ = _read_geno_file(GenoFile, GenoFileIndex, _col(Phen_mtx, "ID"), control)
        else:
            G, _, snp_ids = _split_geno(Geno_mtx)
            keep, geno = _geno_desc_from_matrix(G)
            n, m = G.shape
        if not quiet:
            print(f'[1] "Sample size is {n}."')
            print(f'[1] "Number of variants is {m}."')
        _prep_common(ctx, R, E)
        _set_wald(ctx, traits, Phen_mtx, Cova_mtx)      # all four traits (ref :1187-1262)
        ctx.set_pcs(np.asarray(topPCs, dtype=np.float64))
        if not quiet:
            print('[1] "Start Analyzing..."')
            print(_now())
        prm = _params(epsilon, Cutoff, impute_method, missing_cutoff, min_maf, G_model,
                      pc_pvalue_cutoff=topPCs_pvalue_cutoff, neg_ratio_cutoff=MAF_est_negative_ratio_cutoff)
        out, diag = ctx.run_mix(geno, prm)
        del keep
        if not quiet:
            print('[1] "Analysis Complete."')
            print(_now())
        return Result(out, snp_ids, COLS_MIX, diag)
    finally:
        if own:
            ctx.close()


def SPAGxEmix_CCT_one_SNP(traits, g, R, E, Phen_mtx, Cova_mtx, topPCs, epsilon=0.001, Cutoff=2,
                          impute_method="fixed", missing_cutoff=0.15, min_maf=0.001, G_model="Add",
                          topPCs_pvalue_cutoff=0.05, MAF_est_negative_ratio_cutoff=0.1, ctx=None, device=0):
    """Drop-in for SPAGxEmix_CCT_one_SNP (R/SPAGxECCT.R:1013-1269): numeric(12)."""
    res = SPAGxEmix_CCT(traits, Geno_mtx=np.asarray(g).reshape(-1, 1), R=R, E=E, Phen_mtx=Phen_mtx,
                        Cova_mtx=Cova_mtx, topPCs=topPCs, epsilon=epsilon, Cutoff=Cutoff, impute_method=impute_method,
                        missing_cutoff=missing_cutoff, min_maf=min_maf, G_model=G_model,
                        topPCs_pvalue_cutoff=topPCs_pvalue_cutoff,
                        MAF_est_negative_ratio_cutoff=MAF_est_negative_ratio_cutoff, ctx=ctx, device=device, quiet=True)
    return res.values[0].copy()


COLS_MIX_PLUS = ["MAF", "missing.rate", "p.value.spaGxEmix.plus", "p.value.spaGxEmix.plus.Wald",
                 "p.value.spaGxEmix.plus.CCT.Wald", "p.value.normGxEmix.plus", "p.value.betaG", "Stat.betaG", "Var.betaG",
                 "z.betaG", "MAF.est.negative.num", "MAC"]


def CCT(pvals, weights=None):
    """The reference's CCT() (R/SPAGxECCT.R:2070-2123) for the handful of p-values a caller combines on the host (the Wald
    p-values of SPAGxEmix_CCT_Plus' branch B come from the caller's mixed-model fit); same stop() conditions."""
    import math
    p = np.asarray(pvals, dtype=np.float64)
    if np.isnan(p).any():
        raise RStop("Cannot have NAs in the p-values!")
    if ((p < 0) | (p > 1)).any():
        raise RStop("All p-values must be between 0 and 1!")
    is_zero, is_one = bool((p == 0).any()), bool((p == 1).any())
    if is_zero and is_one:
        raise RStop("Cannot have both 0 and 1 p-values!")
    if is_zero:
        return 0.0
    if is_one:
        p = np.where(p == 1, 0.999, p)     # the reference warns and goes on with 0.999 (:2090-2093)
    w = np.full(p.size, 1.0 / p.size) if weights is None else np.asarray(weights, dtype=np.float64) / np.sum(weights)
    small = p < 1e-16
    # R's sum() accumulates in long double: fsum (exactly rounded) agrees with it for these few terms
    stat = (math.fsum((wi / pi_) / math.pi for wi, pi_ in zip(w[small], p[small]))
            + math.fsum(wi * math.tan((0.5 - pi_) * math.pi) for wi, pi_ in zip(w[~small], p[~small])))
    if stat > 1e15:
        return (1.0 / stat) / math.pi
    # 1 - pcauchy(stat): R's pcauchy evaluates atan(1 / x) / pi for |x| > 1 (upper tail first, then 1 - . twice)
    if abs(stat) > 1:
        y = math.atan(1.0 / stat) / math.pi
        low = (0.5 - y + 0.5) if stat > 0 else -y      # R_D_Cval
    else:
        low = 0.5 + math.atan(stat) / math.pi
    return 1.0 - low


def _mixplus_grm(ResidMat, sparseGRM):
    """the ID handling of SPAGxEmix_CCT_Plus_one_SNP (ref mixPlus :212-229): every subject of ResidMat needs GRM
    information, GRM rows with an ID outside ResidMat are dropped, ID1 / ID2 are matched to their position in SubjID."""
    if "i" in sparseGRM:
        return (np.asarray(sparseGRM["i"], dtype=np.int32), np.asarray(sparseGRM["j"], dtype=np.int32),
                np.asarray(sparseGRM["Value"], dtype=np.float64))
    ids = [str(x) for x in _col(ResidMat, "SubjID")]
    id1 = [str(x) for x in sparseGRM["ID1"]]; id2 = [str(x) for x in sparseGRM["ID2"]]
    in_grm = set(id1) | set(id2)
    if any(x not in in_grm for x in ids):
        raise RStop("At least one subject in residual matrix does not have GRM information.")
    pos = {}
    for k, x in enumerate(ids):
        pos.setdefault(x, k)      # match() returns the first position
    val = np.asarray(sparseGRM["Value"], dtype=np.float64)
    keep = [e for e in range(len(id1)) if id1[e] in pos and id2[e] in pos]
    gi = np.array([pos[id1[e]] for e in keep], dtype=np.int32)
    gj = np.array([pos[id2[e]] for e in keep], dtype=np.int32)
    return gi, gj, val[keep]


def SPAGxEmix_CCT_Plus(traits="binary/quantitative", Geno_mtx=None, ResidMat=None, E=None, Phen_mtx=None, Cova_mtx=None,
                       topPCs=None, sparseGRM=None, epsilon=0.001, Cutoff=2, impute_method="fixed", missing_cutoff=0.15,
                       min_maf=0.001, G_model="Add", topPCs_pvalue_cutoff=0.05, MAF_est_negative_ratio_cutoff=0.1,
                       wald=None, ctx=None, device=0, quiet=False, n_gpus=1):
    """Drop-in for SPAGxEmix_CCT_Plus (R/SPAmixGxECCTPlus_functions_MYZ_2024-09-17.R:38-133): M x 12.  Like the reference it
    takes an R matrix only (no GenoFile) and prints the SNP counter.  R = ResidMat$Resid (the reference's one-SNP function
    reads `R` as a free variable, which its scripts define that way).

    Branch B (p.value.betaG <= epsilon) needs the Wald p-value of `E:g` from lme4::glmer / lmer(Y ~ Cova + E + g + g*E +
    (1 | IID)) (:384, :404), a mixed model that is fitted by the caller, not on the device: `wald(j, g)` receives the
    column index and the imputed / flipped genotype vector of every such SNP and returns that p-value (NaN -> pval.spaGxE,
    :389, :409); the result is combined by CCT() here.  With wald=None a call that meets a branch-B SNP raises
    NotImplementedError; wald="skip" leaves NA in the Wald / CCT columns of those rows."""
    if traits not in ("binary", "quantitative"):
        raise RStop('traits must be "binary" or "quantitative"')
    own = ctx is None
    ctx = _make_ctx(ctx, device, n_gpus)
    try:
        G, _, snp_ids = _split_geno(Geno_mtx)
        keep, geno = _geno_desc_from_matrix(G)
        n, m = G.shape
        if not quiet:
            print(f'[1] "Sample size is {n}."')
            print(f'[1] "Number of variants is {m}."')
        R = np.asarray(_col(ResidMat, "Resid"), dtype=np.float64)
        _prep_common(ctx, R, E)
        ctx.set_pcs(np.asarray(topPCs, dtype=np.float64))
        ctx.set_sparse_grm(*_mixplus_grm(ResidMat, sparseGRM))
        if not quiet:
            print('[1] "Start Analyzing..."')
            print(_now())
            for i in range(1, m + 1):
                print(f"[1] {i}")
        prm = _params(epsilon, Cutoff, impute_method, missing_cutoff, min_maf, G_model,
                      pc_pvalue_cutoff=topPCs_pvalue_cutoff, neg_ratio_cutoff=MAF_est_negative_ratio_cutoff)
        out, diag = ctx.run_mix_plus(geno, prm)
        del keep
        bb = np.flatnonzero(out[:, 6] <= epsilon)      # NA rows compare False
        if bb.size and wald is None:
            raise NotImplementedError(
                f"SPAGxEmix_CCT_Plus: {bb.size} SNP(s) take the genotype-adjusted branch, whose Wald test is a glmer / lmer "
                "mixed model fitted by the caller: pass wald=callable (or wald='skip' for NA in those columns)")
        if bb.size and callable(wald):
            Gd = np.asarray(G, dtype=np.float64)
            for j in bb:
                g = Gd[:, j].copy()
                if G.dtype.kind == "i":
                    g[G[:, j] == np.iinfo(np.int32).min] = np.nan
                maf = np.nanmean(g) / 2
                g[np.isnan(g)] = 2 * maf
                if maf > 0.5:
                    g = 2 - g
                if G_model == "Dom":
                    g = np.where(g >= 1, 1.0, 0.0)
                elif G_model == "Rec":
                    g = np.where(g <= 1, 0.0, 1.0)
                pw = float(wald(int(j), g))
                if np.isnan(pw):
                    pw = out[j, 2]
                out[j, 3] = pw
                out[j, 4] = CCT([out[j, 2], pw])
        if not quiet:
            print('[1] "Analysis Complete."')
            print(_now())
        return Result(out, snp_ids, COLS_MIX_PLUS, diag)
    finally:
        if own:
            ctx.close()


def SPAGxEmix_CCT_Plus_one_SNP(traits, g, ResidMat, E, Phen_mtx, Cova_mtx, topPCs, sparseGRM, epsilon=0.001, Cutoff=2,
                               impute_method="fixed", missing_cutoff=0.15, min_maf=0.001, G_model="Add",
                               topPCs_pvalue_cutoff=0.05, MAF_est_negative_ratio_cutoff=0.1, wald=None, ctx=None, device=0):
    """Drop-in for SPAGxEmix_CCT_Plus_one_SNP (R/SPAmixGxECCTPlus_functions_MYZ_2024-09-17.R:144-427): numeric(12)."""
    res = SPAGxEmix_CCT_Plus(traits, Geno_mtx=np.asarray(g).reshape(-1, 1), ResidMat=ResidMat, E=E, Phen_mtx=Phen_mtx,
                             Cova_mtx=Cova_mtx, topPCs=topPCs, sparseGRM=sparseGRM, epsilon=epsilon, Cutoff=Cutoff,
                             impute_method=impute_method, missing_cutoff=missing_cutoff, min_maf=min_maf, G_model=G_model,
                             topPCs_pvalue_cutoff=topPCs_pvalue_cutoff,
                             MAF_est_negative_ratio_cutoff=MAF_est_negative_ratio_cutoff, wald=wald, ctx=ctx, device=device,
                             quiet=True)
    return res.values[0].copy()


def SPAGxE_Plus(GenoFile=None, GenoFileIndex=None, control=None, Geno_mtx=None, E=None, Phen_mtx=None,
                Cova_mtx=None, obj_SPAGxE_Plus_Nullmodel=None, epsilon=0.001, Cutoff=2, impute_method="fixed",
                missing_cutoff=0.15, min_maf=0.001, G_model="Add", ctx=None, device=0, quiet=False, n_gpus=1):
    """Drop-in for SPAGxE_Plus (R/SPAGxEPlus_functions_MYZ.R:159-238): M x 8.
    obj_SPAGxE_Plus_Nullmodel: mapping with R, R_GRM_R, R_GRM_RE, R.new, R.new_GRM_R.new and sparseGRM
    (ID1/ID2 as 0-based sample indices `i`,`j` or as id strings matched against ResidMat$SubjID, Value)."""
    own = ctx is None
    ctx = _make_ctx(ctx, device, n_gpus)
    try:
        obj = obj_SPAGxE_Plus_Nullmodel
        if Geno_mtx is None:
            geno, keep, n, m, snp_ids, _ = _read_geno_file(GenoFile, GenoFileIndex, _col(Phen_mtx, "ID"), control)
        else:
            G, _, snp_ids = _split_geno(Geno_mtx)
            keep, geno = _geno_desc_from_matrix(G)
            n, m = G.shape
        if not quiet:
            print(f'[1] "Sample size is {n}."')
            print(f'[1] "Number of variants is {m}."')
        _prep_common(ctx, obj["R"], E)
        gi, gj, gv = _grm_triplets(obj)
        ctx.set_grm(gi, gj, gv, obj["R_GRM_R"], obj["R_GRM_RE"], obj["R.new_GRM_R.new"], obj["R.new"])
        if not quiet:
            print('[1] "Start Analyzing..."')
            print(_now())
        out, diag = ctx.run_plus(geno, _params(epsilon, Cutoff, impute_method, missing_cutoff, min_maf, G_model))
        del keep
        if not quiet:
            print('[1] "Analysis Complete."')
            print(_now())
        return Result(out, snp_ids, COLS_PLUS, diag)
    finally:
        if own:
            ctx.close()


def SPAGxE_Plus_one_SNP(g, E, Phen_mtx, Cova_mtx, obj_SPAGxE_Plus_Nullmodel, epsilon=0.001, Cutoff=2,
                        impute_method="fixed", missing_cutoff=0.15, min_maf=0.001, G_model="Add", ctx=None, device=0):
    """Drop-in for SPAGxE_Plus_one_SNP (R/SPAGxEPlus_functions_MYZ.R:248-405): numeric(8)."""
    res = SPAGxE_Plus(Geno_mtx=np.asarray(g).reshape(-1, 1), E=E, Phen_mtx=Phen_mtx, Cova_mtx=Cova_mtx,
                      obj_SPAGxE_Plus_Nullmodel=obj_SPAGxE_Plus_Nullmodel, epsilon=epsilon, Cutoff=Cutoff,
                      impute_method=impute_method, missing_cutoff=missing_cutoff, min_maf=min_maf, G_model=G_model,
                      ctx=ctx, device=device, quiet=True)
    return res.values[0].copy()


COLS_LOCAL = ["MAF", "missing.rate", "p.value.spaGxE.index.ance", "p.value.spaGxE.Wald.index.ance",
              "p.value.spaGxE.CCT.Wald.index.ance", "p.value.normGxE.index.ance", "p.value.betaG.index.ance",
              "Stat.betaG.index.ance", "Var.betaG.index.ance", "z.betaG.index.ance"]


def SPAGxEmixCCT_localance(traits="survival/binary/quantitative/categorical", Geno_mtx=None, haplo_mtx=None, R=None, Cutoff=2,
                           E=None, Phen_mtx=None, Cova_mtx=None, Cova_haplo_mtx_list=None, epsilon=0.001,
                           impute_method="fixed", missing_cutoff=0.15, min_maf=0.0001, G_model="Add", ctx=None, device=0,
                           quiet=False):
    """Drop-in for SPAGxEmixCCT_localance (R/SPAGxECCT.R:1401-1481): M x 10 (the reference returns them as a data
    frame behind an rsID column: Result.rownames holds the marker names).  Cova_haplo_mtx_list: mapping or sequence
    of (N, M) matrices.  Binary and quantitative traits; Cutoff is accepted and not used, like the reference."""
    if _trait_code(traits) not in (_lib.TRAIT_BINARY, _lib.TRAIT_QUANTITATIVE):
        raise NotImplementedError('SPAGxEmixCCT_localance: the survival and categorical Wald tests are implemented for SPAGxE_CCT only')
    if impute_method != "fixed":
        raise NotImplementedError('only impute.method = "fixed" exists in the reference')
    own = ctx is None
    ctx = ctx or _lib.Context(device)
    try:
        G, _, snp_ids = _split_geno(Geno_mtx)
        H, _, _ = _split_geno(haplo_mtx)
        chl = Cova_haplo_mtx_list or []
        chs = [_split_geno(c)[0] for c in (chl.values() if hasattr(chl, "values") else chl)]
        n, m = np.shape(G)
        if not quiet:
            print(f'[1] "Sample size is {n}."')
            print(f'[1] "Number of variants is {m}."')
        _prep_common(ctx, R, E)
        ctx.set_wald(_trait_code(traits), _wald_y(traits, Phen_mtx), np.asarray(Cova_mtx, dtype=np.float64))
        if not quiet:
            print('[1] "Start Analyzing..."')
            print(_now())
        out, diag = ctx.run_localance(G, H, chs, _params(epsilon, Cutoff, impute_method, missing_cutoff, min_maf, G_model))
        if not quiet:
            print('[1] "Analysis Complete."')
            print(_now())
        return Result(out, snp_ids, COLS_LOCAL, diag)
    finally:
        if own:
            ctx.close()


def SPAGxEmixCCT_localance_one_SNP(traits, g, haplo_numVec, R, E, Phen_mtx, Cova_mtx, Cova_haplo_mtx=None, epsilon=0.001,
                                   Cutoff=2, impute_method="fixed", missing_cutoff=0.15, min_maf=0.0001, G_model="Add",
                                   ctx=None, device=0):
    """Drop-in for SPAGxEmixCCT_localance_one_SNP (R/SPAGxECCT.R:1496-1647): numeric(10).  Cova_haplo_mtx: (N, L)."""
    ch = [] if Cova_haplo_mtx is None else [np.asarray(Cova_haplo_mtx, dtype=np.float64).reshape(len(g), -1)[:, [l]]
                                            for l in range(np.asarray(Cova_haplo_mtx).reshape(len(g), -1).shape[1])]
    res = SPAGxEmixCCT_localance(traits, Geno_mtx=np.asarray(g, dtype=np.float64).reshape(-1, 1),
                                 haplo_mtx=np.asarray(haplo_numVec, dtype=np.float64).reshape(-1, 1), R=R, Cutoff=Cutoff, E=E,
                                 Phen_mtx=Phen_mtx, Cova_mtx=Cova_mtx, Cova_haplo_mtx_list=ch, epsilon=epsilon,
                                 impute_method=impute_method, missing_cutoff=missing_cutoff, min_maf=min_maf, G_model=G_model,
                                 ctx=ctx, device=device, quiet=True)
    return res.values[0].copy()


def _grm_triplets(obj):
    grm = obj["sparseGRM"]
    if "i" in grm:
        return np.asarray(grm["i"], dtype=np.int32), np.asarray(grm["j"], dtype=np.int32), np.asarray(grm["Value"])
    ids = [str(s) for s in obj["ResidMat"]["SubjID"]]
    pos = {s: k for k, s in enumerate(ids)}
    gi = np.array([pos[str(s)] for s in grm["ID1"]], dtype=np.int32)
    gj = np.array([pos[str(s)] for s in grm["ID2"]], dtype=np.int32)
    return gi, gj, np.asarray(grm["Value"], dtype=np.float64)
