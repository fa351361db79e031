"""ctypes binding of libspagxe_cuda.so (include/spagxe.h).

There is no CPU fallback: importing the package works anywhere (so host-side logic can be
tested), but every compute entry point raises if the CUDA library is missing or no sm_100
device is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspagxe_cuda.so")

OK, ERR_CUDA, ERR_ARG, ERR_STATE, ERR_UNSUPPORTED, ERR_CCT_STOP, ERR_SINGULAR = range(7)
TRAIT_BINARY, TRAIT_QUANTITATIVE, TRAIT_SURVIVAL, TRAIT_CATEGORICAL = 1, 2, 3, 4
GENO_BED, GENO_F64, GENO_I32 = 0, 1, 2
LOC_HOST, LOC_DEVICE = 0, 1
GENO_COUNT_A2 = 1

EXPORTS = [
    "spagxe_ctx_create", "spagxe_ctx_destroy", "spagxe_last_error", "spagxe_ctx_set_stream", "spagxe_version",
    "spagxe_set_vectors", "spagxe_set_vectors_device", "spagxe_set_wald", "spagxe_set_pcs", "spagxe_set_grm",
    "spagxe_run_cct", "spagxe_run_mix", "spagxe_run_plus", "spagxe_run_localance", "spagxe_bed_counts", "spagxe_bed_decode",
    "spagxe_measure_fp64_peak", "spagxe_score_stats", "spagxe_ctx_set_option", "spagxe_plus_nullmodel",
    "spagxe_set_wald_device", "spagxe_set_wald_survival", "spagxe_mgpu_create", "spagxe_mgpu_destroy", "spagxe_mgpu_last_error", "spagxe_mgpu_n_gpus",
    "spagxe_mgpu_set_inputs", "spagxe_mgpu_run_cct", "spagxe_mgpu_run_mix", "spagxe_mgpu_run_plus", "spagxe_mgpu_set_wald_survival",
    "spagxe_mgpu_set_pcs", "spagxe_mgpu_set_grm", "spagxe_set_sparse_grm", "spagxe_run_mix_plus",
    "spagxe_mgpu_set_sparse_grm", "spagxe_mgpu_run_mix_plus",
]
# test / tuning hooks (include/spagxe.h SPAGXE_OPT_*): not part of the reference-facing surface
OPT_CHUNK_SNPS, OPT_EXACT_SPA, OPT_MIX_PER_SNP, OPT_SCORE_IMPL, OPT_MIX_CLUSTER = 1, 2, 3, 4, 5


class Geno(C.Structure):
    _fields_ = [("kind", C.c_int32), ("location", C.c_int32), ("data", C.c_void_p), ("n_samples", C.c_int64),
                ("n_snps", C.c_int64), ("pitch", C.c_int64), ("sample_index", C.c_void_p), ("n_file_samples", C.c_int64),
                ("flags", C.c_int32), ("reserved", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("epsilon", C.c_double), ("cutoff", C.c_double), ("min_maf", C.c_double),
                ("missing_cutoff", C.c_double), ("pc_pvalue_cutoff", C.c_double), ("neg_ratio_cutoff", C.c_double),
                ("g_model", C.c_int32), ("impute_method", C.c_int32), ("out_location", C.c_int32),
                ("reserved", C.c_int32), ("out_ld", C.c_int64)]


class Diag(C.Structure):
    _fields_ = [(k, C.c_int64) for k in
                ["n_snps", "n_filtered", "n_branch_a", "n_branch_b", "n_spa", "n_wald_fail", "n_cct_stop",
                 "n_singular", "n_mix_logistic", "newton_iters", "kernel_launches", "h2d_bytes", "d2h_bytes"]] + \
               [("ms_total", C.c_double), ("ms_score", C.c_double), ("ms_tests", C.c_double),
                ("score_bytes", C.c_int64), ("ms_spa", C.c_double), ("ms_branch_b", C.c_double),
                ("spa_nodes", C.c_int64), ("bb_irls_sample_passes", C.c_int64), ("bb_tail_sample_passes", C.c_int64),
                ("bb_prep_sample_passes", C.c_int64), ("bb_iterations", C.c_int64), ("cox_evals", C.c_int64), ("clm_evals", C.c_int64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SpagxeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[spagxe status {code}] {msg}")
        self.code = code


_lib = None


def load():
    """Load the CUDA library or fail loudly (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). spagxecct_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.spagxe_last_error.restype = C.c_char_p
    L.spagxe_last_error.argtypes = [C.c_void_p]
    L.spagxe_version.restype = C.c_char_p
    L.spagxe_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.spagxe_ctx_destroy.argtypes = [C.c_void_p]
    L.spagxe_ctx_destroy.restype = None
    L.spagxe_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.spagxe_ctx_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int64]
    L.spagxe_set_vectors.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.spagxe_set_vectors_device.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.spagxe_set_wald.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.spagxe_set_wald_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.spagxe_set_wald_survival.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.spagxe_mgpu_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    L.spagxe_mgpu_destroy.argtypes = [C.c_void_p]
    L.spagxe_mgpu_destroy.restype = None
    L.spagxe_mgpu_last_error.argtypes = [C.c_void_p]
    L.spagxe_mgpu_last_error.restype = C.c_char_p
    L.spagxe_mgpu_n_gpus.argtypes = [C.c_void_p]
    L.spagxe_mgpu_set_inputs.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.spagxe_mgpu_set_sparse_grm.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    for f in ("spagxe_mgpu_run_cct", "spagxe_mgpu_run_mix", "spagxe_mgpu_run_plus", "spagxe_mgpu_run_mix_plus"):
        getattr(L, f).argtypes = [C.c_void_p, C.POINTER(Geno), C.c_int, C.POINTER(Params), C.c_void_p, C.POINTER(Diag)]
    L.spagxe_mgpu_set_wald_survival.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.spagxe_mgpu_set_pcs.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.spagxe_mgpu_set_grm.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                      C.c_double, C.c_void_p]
    L.spagxe_set_pcs.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.spagxe_set_grm.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                 C.c_double, C.c_void_p]
    L.spagxe_plus_nullmodel.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
    for f in ("spagxe_run_cct", "spagxe_run_mix", "spagxe_run_plus", "spagxe_run_mix_plus"):
        getattr(L, f).argtypes = [C.c_void_p, C.POINTER(Geno), C.POINTER(Params), C.c_void_p, C.POINTER(Diag)]
    L.spagxe_set_sparse_grm.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.spagxe_run_localance.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64,
                                       C.POINTER(Params), C.c_void_p, C.POINTER(Diag)]
    L.spagxe_bed_counts.argtypes = [C.c_void_p, C.POINTER(Geno), C.c_void_p]
    L.spagxe_bed_decode.argtypes = [C.c_void_p, C.POINTER(Geno), C.c_void_p]
    L.spagxe_measure_fp64_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.spagxe_score_stats.argtypes = [C.c_void_p, C.POINTER(Geno), C.c_int] + [C.c_void_p] * 6
    _lib = L
    return L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """One GPU worth of Step-2 state (device copies of R, E, Wald data, PCs, GRM)."""

    def __init__(self, device=0):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.spagxe_ctx_create(int(device), C.byref(h))
        if rc != OK:
            raise SpagxeError(rc, self.L.spagxe_last_error(None).decode())
        self.h = h
        self.n = 0
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.spagxe_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc):
        if rc != OK:
            raise SpagxeError(rc, self.L.spagxe_last_error(self.h).decode())

    def set_stream(self, cuda_stream_ptr):
        self._chk(self.L.spagxe_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_option(self, option, value):
        """Test / tuning hook (never needed by a caller of the R-mirroring functions)."""
        self._chk(self.L.spagxe_ctx_set_option(self.h, int(option), int(value)))

    def set_vectors(self, R, E):
        R = _f64(R); E = _f64(E)
        if R.shape != E.shape or R.ndim != 1:
            raise ValueError("R and E must be vectors of the same length")
        self._chk(self.L.spagxe_set_vectors(self.h, R.size, R.ctypes.data, E.ctypes.data))
        self.n = R.size

    def set_vectors_device(self, n, dR_ptr, dE_ptr):
        self._chk(self.L.spagxe_set_vectors_device(self.h, int(n), C.c_void_p(dR_ptr), C.c_void_p(dE_ptr)))
        self.n = int(n)

    def set_wald(self, trait, Y, cova):
        Y = _f64(Y)
        cova = np.asfortranarray(np.asarray(cova, dtype=np.float64).reshape(Y.size, -1))
        self._chk(self.L.spagxe_set_wald(self.h, int(trait), Y.ctypes.data, cova.ctypes.data, cova.shape[1]))

    def set_wald_survival(self, surv_time, event, cova):
        """traits = "survival": Phen.mtx$surv.time, Phen.mtx$event, as.matrix(Cova.mtx) (ref R/SPAGxECCT.R:624)."""
        t = _f64(surv_time); ev = _f64(event)
        cova = np.asfortranarray(np.asarray(cova, dtype=np.float64).reshape(t.size, -1))
        self._chk(self.L.spagxe_set_wald_survival(self.h, t.ctypes.data, ev.ctypes.data, cova.ctypes.data, cova.shape[1]))

    def set_wald_device(self, trait, dY_ptr, dcova_ptr, p):
        self._chk(self.L.spagxe_set_wald_device(self.h, int(trait), C.c_void_p(dY_ptr), C.c_void_p(dcova_ptr), int(p)))

    def set_pcs(self, pcs):
        pcs = np.asfortranarray(np.asarray(pcs, dtype=np.float64).reshape(self.n, -1))
        self._chk(self.L.spagxe_set_pcs(self.h, pcs.ctypes.data, pcs.shape[1]))

    def set_grm(self, gi, gj, gv, R_GRM_R, R_GRM_RE, Rnew_GRM_Rnew, Rnew):
        gi = np.ascontiguousarray(gi, dtype=np.int32); gj = np.ascontiguousarray(gj, dtype=np.int32)
        gv = _f64(gv); Rnew = _f64(Rnew)
        self._chk(self.L.spagxe_set_grm(self.h, gi.size, gi.ctypes.data, gj.ctypes.data, gv.ctypes.data,
                                        float(R_GRM_R), float(R_GRM_RE), float(Rnew_GRM_Rnew), Rnew.ctypes.data))

    def set_sparse_grm(self, gi, gj, gv):
        """sparseGRM of SPAGxEmix_CCT_Plus as 0-based COO triplets (no Step-1 scalars)."""
        gi = np.ascontiguousarray(gi, dtype=np.int32); gj = np.ascontiguousarray(gj, dtype=np.int32)
        gv = _f64(gv)
        self._chk(self.L.spagxe_set_sparse_grm(self.h, gi.size, gi.ctypes.data, gj.ctypes.data, gv.ctypes.data))

    def plus_nullmodel(self, gi, gj, gv, R, E):
        """(R_GRM_R, R_GRM_RE, R.new, R.new_GRM_R.new) of SPAGxE_Plus_Nullmodel, computed on the device."""
        gi = np.ascontiguousarray(gi, dtype=np.int32); gj = np.ascontiguousarray(gj, dtype=np.int32)
        gv = _f64(gv); R = _f64(R); E = _f64(E)
        out3 = np.empty(3); Rnew = np.empty(R.size)
        self._chk(self.L.spagxe_plus_nullmodel(self.h, R.size, R.ctypes.data, E.ctypes.data, gi.size, gi.ctypes.data,
                                               gj.ctypes.data, gv.ctypes.data, out3.ctypes.data, Rnew.ctypes.data))
        return float(out3[0]), float(out3[1]), Rnew, float(out3[2])

    @staticmethod
    def geno_desc(kind, data_ptr, n, m, pitch, location=LOC_HOST, sample_index=None, n_file_samples=0, count_a2=False):
        """sample_index: int64 array (kept alive by the caller): position inside a file row of analysis sample i."""
        sidx = None
        if sample_index is not None:
            assert sample_index.dtype == np.int64 and sample_index.flags.c_contiguous and sample_index.size == n
            sidx = C.c_void_p(sample_index.ctypes.data)
        return Geno(kind, location, C.c_void_p(data_ptr), int(n), int(m), int(pitch), sidx, int(n_file_samples),
                    GENO_COUNT_A2 if count_a2 else 0, 0)

    def _run(self, fn, ncol, geno, params, out_ptr=None):
        prm = params if isinstance(params, Params) else make_params(**(params or {}))
        diag = Diag()
        if out_ptr is None:
            out = np.empty((geno.n_snps, ncol), dtype=np.float64, order="F")
            rc = fn(self.h, C.byref(geno), C.byref(prm), C.c_void_p(out.ctypes.data), C.byref(diag))
        else:
            out = None
            rc = fn(self.h, C.byref(geno), C.byref(prm), C.c_void_p(out_ptr), C.byref(diag))
        self.last_diag = diag.asdict()
        self.last_out = out
        self._chk(rc)
        return out, self.last_diag

    def run_cct(self, geno, params=None, out_ptr=None):
        return self._run(self.L.spagxe_run_cct, 10, geno, params, out_ptr)

    def run_mix(self, geno, params=None, out_ptr=None):
        return self._run(self.L.spagxe_run_mix, 12, geno, params, out_ptr)

    def run_plus(self, geno, params=None, out_ptr=None):
        return self._run(self.L.spagxe_run_plus, 8, geno, params, out_ptr)

    def run_mix_plus(self, geno, params=None, out_ptr=None):
        return self._run(self.L.spagxe_run_mix_plus, 12, geno, params, out_ptr)

    def run_localance(self, G, H, cova_haplo=(), params=None):
        """SPAGxEmixCCT_localance loop (ref R/SPAGxECCT.R:1443-1476): G, H and every matrix of cova_haplo are (N, M)."""
        G = np.asfortranarray(G, dtype=np.float64); H = np.asfortranarray(H, dtype=np.float64)
        chs = [np.asfortranarray(c, dtype=np.float64) for c in cova_haplo]
        n, m = G.shape
        assert H.shape == (n, m) and all(c.shape == (n, m) for c in chs)
        arr = (C.c_void_p * max(len(chs), 1))(*[c.ctypes.data for c in chs])
        prm = params if isinstance(params, Params) else make_params(**(params or {}))
        diag = Diag()
        out = np.empty((m, 10), dtype=np.float64, order="F")
        rc = self.L.spagxe_run_localance(self.h, n, m, G.ctypes.data, H.ctypes.data, len(chs), arr, n, C.byref(prm),
                                         out.ctypes.data, C.byref(diag))
        self.last_diag = diag.asdict()
        self._chk(rc)
        return out, self.last_diag

    def bed_counts(self, geno):
        counts = np.zeros((geno.n_snps, 4), dtype=np.int64)
        self._chk(self.L.spagxe_bed_counts(self.h, C.byref(geno), counts.ctypes.data))
        return counts

    def bed_decode(self, geno):
        out = np.empty((geno.n_samples, geno.n_snps), dtype=np.float64, order="F")
        self._chk(self.L.spagxe_bed_decode(self.h, C.byref(geno), out.ctypes.data))
        return out

    def score_stats(self, geno, impl=0):
        m = geno.n_snps
        d = [np.empty(m) for _ in range(4)]
        a = np.empty(m, dtype=np.int32); nm = np.empty(m, dtype=np.int32)
        self._chk(self.L.spagxe_score_stats(self.h, C.byref(geno), int(impl), *[x.ctypes.data for x in d],
                                            a.ctypes.data, nm.ctypes.data))
        return dict(D0=d[0], D1=d[1], T0=d[2], T1=d[3], asum=a, nmiss=nm)

    def fp64_peak_tflops(self):
        v = C.c_double()
        self._chk(self.L.spagxe_measure_fp64_peak(self.h, C.byref(v)))
        return v.value


class MultiContext:
    """SPAGxE_CCT on several GPUs of one node through the C ABI (spagxe_mgpu_*): one process, one host thread + stream
    per GPU, ONE ncclBroadcast of [R | E | Y | Cova], contiguous SNP ranges, results copied into disjoint row blocks."""

    def __init__(self, n_gpus, devices=None):
        self.L = load()
        h = C.c_void_p()
        dev = None
        if devices is not None:
            dev = (C.c_int * n_gpus)(*[int(d) for d in devices])
        rc = self.L.spagxe_mgpu_create(int(n_gpus), dev, C.byref(h))
        if rc != OK:
            raise SpagxeError(rc, self.L.spagxe_mgpu_last_error(None).decode())
        self.h = h
        self.n_gpus = int(n_gpus)

    def close(self):
        if getattr(self, "h", None):
            self.L.spagxe_mgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc):
        if rc != OK:
            raise SpagxeError(rc, self.L.spagxe_mgpu_last_error(self.h).decode())

    def set_inputs(self, R, E, trait=0, Y=None, cova=None):
        """trait 0: no Wald data (SPAGxE_Plus, or a survival trait followed by set_wald_survival)."""
        R = _f64(R); E = _f64(E)
        self.n = R.size
        if not trait:
            self._chk(self.L.spagxe_mgpu_set_inputs(self.h, R.size, R.ctypes.data, E.ctypes.data, 0, None, None, 0))
            return
        Y = _f64(Y)
        cova = np.asfortranarray(np.asarray(cova, dtype=np.float64).reshape(Y.size, -1))
        self._chk(self.L.spagxe_mgpu_set_inputs(self.h, R.size, R.ctypes.data, E.ctypes.data, int(trait), Y.ctypes.data,
                                                cova.ctypes.data, cova.shape[1]))

    def set_wald_survival(self, surv_time, event, cova):
        t = _f64(surv_time); ev = _f64(event)
        cova = np.asfortranarray(np.asarray(cova, dtype=np.float64).reshape(t.size, -1))
        self._chk(self.L.spagxe_mgpu_set_wald_survival(self.h, t.ctypes.data, ev.ctypes.data, cova.ctypes.data, cova.shape[1]))

    def set_pcs(self, pcs):
        pcs = np.asfortranarray(np.asarray(pcs, dtype=np.float64).reshape(self.n, -1))
        self._chk(self.L.spagxe_mgpu_set_pcs(self.h, pcs.ctypes.data, pcs.shape[1]))

    def set_grm(self, gi, gj, gv, R_GRM_R, R_GRM_RE, Rnew_GRM_Rnew, Rnew):
        gi = np.ascontiguousarray(gi, dtype=np.int32); gj = np.ascontiguousarray(gj, dtype=np.int32)
        gv = _f64(gv); Rnew = _f64(Rnew)
        self._chk(self.L.spagxe_mgpu_set_grm(self.h, gi.size, gi.ctypes.data, gj.ctypes.data, gv.ctypes.data,
                                             float(R_GRM_R), float(R_GRM_RE), float(Rnew_GRM_Rnew), Rnew.ctypes.data))

    def set_sparse_grm(self, gi, gj, gv):
        gi = np.ascontiguousarray(gi, dtype=np.int32); gj = np.ascontiguousarray(gj, dtype=np.int32)
        gv = _f64(gv)
        self._chk(self.L.spagxe_mgpu_set_sparse_grm(self.h, gi.size, gi.ctypes.data, gj.ctypes.data, gv.ctypes.data))

    def run_mix_plus(self, geno, params=None, out_ptr=None):
        return self.run_cct(geno, params, out_ptr, _fn="spagxe_mgpu_run_mix_plus", _ncol=12)

    def run_mix(self, geno, params=None, out_ptr=None):
        return self.run_cct(geno, params, out_ptr, _fn="spagxe_mgpu_run_mix", _ncol=12)

    def run_plus(self, geno, params=None, out_ptr=None):
        return self.run_cct(geno, params, out_ptr, _fn="spagxe_mgpu_run_plus", _ncol=8)

    def run_cct(self, geno, params=None, out_ptr=None, _fn="spagxe_mgpu_run_cct", _ncol=10):
        """geno: one host Geno (cut into SNP ranges) or a list of n_gpus Geno shards (host or that GPU's memory)."""
        prm = params if isinstance(params, Params) else make_params(**(params or {}))
        diag = Diag()
        if isinstance(geno, (list, tuple)):
            arr = (Geno * len(geno))(*geno)
            n_sh, m = len(geno), sum(g.n_snps for g in geno)
            gp = C.cast(arr, C.POINTER(Geno))
        else:
            n_sh, m = 1, geno.n_snps
            gp = C.byref(geno)
        out = None
        if out_ptr is None:
            out = np.empty((m, _ncol), dtype=np.float64, order="F")
            out_ptr = out.ctypes.data
        rc = getattr(self.L, _fn)(self.h, gp, n_sh, C.byref(prm), C.c_void_p(out_ptr), C.byref(diag))
        self.last_diag = diag.asdict()
        self._chk(rc)
        return out, self.last_diag


def make_params(epsilon=0.001, cutoff=2.0, min_maf=0.001, missing_cutoff=0.15, pc_pvalue_cutoff=0.05,
                neg_ratio_cutoff=0.1, g_model=0, impute_method=0, out_location=LOC_HOST):
    return Params(epsilon, cutoff, min_maf, missing_cutoff, pc_pvalue_cutoff, neg_ratio_cutoff, g_model,
                  impute_method, out_location, 0, 0)
