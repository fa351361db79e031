"""Bridge to the real reference: when an `Rscript` is on PATH and the reference sources are reachable
(/root/reference or $SPAGXE_REFERENCE_DIR), tools/run_reference.R source()s the UNMODIFIED R files and runs the
exported Step-2 functions on inputs written here.  Absent either, callers skip (R is not in the build image; the GPU box
has no /root/reference), and the oracle stays "parity unpinned" for numeric outputs."""
import os
import shutil
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_dir():
    for d in (os.environ.get("SPAGXE_REFERENCE_DIR"), "/root/reference"):
        if d and os.path.isdir(os.path.join(d, "R")):
            return d
    return None


def available():
    return shutil.which("Rscript") is not None and reference_dir() is not None


def run(mode, trait, G, R, E, Y, Cova, PCs=None, grm=None, scal=None, Rnew=None, epsilon=0.001, cutoff=2.0,
        min_maf=0.001, timeout=3600):
    """-> (M x ncol matrix from the reference R code, seconds spent inside the R function)."""
    G = np.asfortranarray(G, dtype=np.float64)
    n, m = G.shape
    Cova = np.asfortranarray(np.asarray(Cova, dtype=np.float64).reshape(n, -1))
    p = Cova.shape[1]
    k = 0 if PCs is None else np.asarray(PCs).reshape(n, -1).shape[1]
    ncol = {"cct": 10, "mix": 12, "plus": 8}[mode]
    with tempfile.TemporaryDirectory() as d:
        def w(name, a):
            np.asfortranarray(a, dtype=np.float64).ravel(order="F").astype("<f8").tofile(os.path.join(d, name))
        w("geno.bin", G); w("R.bin", R); w("E.bin", E); w("Y.bin", Y)
        if p:
            w("cova.bin", Cova)
        meta = [mode, trait, n, m, p, k, repr(float(epsilon)), repr(float(cutoff)), repr(float(min_maf))]
        if mode == "mix":
            w("pcs.bin", np.asarray(PCs).reshape(n, -1))
        if mode == "plus":
            gi, gj, gv = grm
            w("grm_i.bin", np.asarray(gi, dtype=np.float64) + 1); w("grm_j.bin", np.asarray(gj, dtype=np.float64) + 1)
            w("grm_v.bin", gv); w("scal.bin", np.asarray(scal, dtype=np.float64)); w("rnew.bin", Rnew)
            meta.append(len(gv))
        open(os.path.join(d, "meta.txt"), "w").write(" ".join(str(x) for x in meta) + "\n")
        outf, timf = os.path.join(d, "out.bin"), os.path.join(d, "time.txt")
        cp = subprocess.run(["Rscript", os.path.join(ROOT, "tools", "run_reference.R"), reference_dir(), d, outf, timf],
                            capture_output=True, text=True, timeout=timeout)
        if cp.returncode != 0:
            raise RuntimeError(f"Rscript failed: {cp.stderr[-2000:]}")
        out = np.fromfile(outf, dtype="<f8").reshape(ncol, m).T.copy()
        secs = float(open(timf).read().split()[2])
    return out, secs
