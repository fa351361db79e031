#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's bundled fixtures.

Run in the build container only (needs /root/reference).  The GPU box has no
/root/reference, so everything the tests need is committed under tests/golden/.

What is pinned by the reference's own files:
  * PLINK 2-bit decode: inst/GenoMat_*.raw (plink --recode A text) vs inst/GenoMat_*.bed
  * allele convention of GRAB.ReadGeno: data/GenoMat.SPAGxEPlus.rda[, 1:100] vs the .bed
Everything else stored here is *input* data (phenotypes, PCs, sparse GRM).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from rdata import RObj, load_rdata  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def read_raw(path):
    """plink --recode A: header + one row per sample, dosage columns from col 6."""
    with open(path) as f:
        hdr = f.readline().split()
        snps = hdr[6:]
        rows = []
        iids = []
        for line in f:
            t = line.split()
            iids.append(t[1])
            rows.append([-1 if x == "NA" else int(x) for x in t[6:]])
    return np.asarray(rows, dtype=np.int8), snps, iids


def main():
    os.makedirs(OUT, exist_ok=True)
    # ---- C1: data/Geno.mtx.rda + data/Pheno.mtx.rda
    geno = load_rdata(f"{REF}/data/Geno.mtx.rda")["Geno.mtx"]
    phen = load_rdata(f"{REF}/data/Pheno.mtx.rda")["Pheno.mtx"]
    assert isinstance(geno, RObj) and geno.value.dtype == np.int32
    np.savez_compressed(
        f"{OUT}/c1.npz",
        geno=geno.value.astype(np.int8),
        snp_names=np.array(geno.attrs["dimnames"][1]),
        iid_is_sequential=np.array(phen["IID"] == [f"IID-{i + 1}" for i in range(10000)]),
        **{k: np.asarray(phen[k], dtype=np.float64) for k in ["y", "Cov1", "Cov2", "E", "PC1", "PC2", "PC3", "PC4", "Y"]},
    )
    # ---- C2 and friends: the three .bed filesets with their .raw golden decode
    for name in ["SPAGxE", "SPAGxEmix", "SPAGxE_Plus"]:
        bed = np.fromfile(f"{REF}/inst/GenoMat_{name}.bed", dtype=np.uint8)
        raw, snps, iids = read_raw(f"{REF}/inst/GenoMat_{name}.raw")
        fam = [l.split()[1] for l in open(f"{REF}/inst/GenoMat_{name}.fam")]
        bim = [l.split() for l in open(f"{REF}/inst/GenoMat_{name}.bim")]
        assert fam == iids
        np.savez_compressed(
            f"{OUT}/bed_{name}.npz", bed=bed, raw=raw, snp_ids=np.array([b[1] for b in bim]),
            a1=np.array([b[4] for b in bim]), raw_hdr=np.array(snps),
            fam_first=np.array(fam[:4]), n=np.int64(len(fam)))
    # ---- Plus: phenotype + sparse GRM; check rda[,1:100] == bed decode (A1 count)
    ph = load_rdata(f"{REF}/data/Phen.mtx.SPAGxEPlus.binary.rda")["Phen.mtx.SPAGxEPlus.binary"]
    grm = load_rdata(f"{REF}/data/sparseGRM.SPAGxEPlus.rda")["sparseGRM.SPAGxEPlus"]
    gm = load_rdata(f"{REF}/data/GenoMat.SPAGxEPlus.rda")
    gm = list(gm.values())[0]
    raw_plus = np.load(f"{OUT}/bed_SPAGxE_Plus.npz")["raw"]
    first100_equal = bool(np.array_equal(gm.value[:, :100].astype(np.int8), raw_plus))
    idx = {s: i for i, s in enumerate(ph["IID"])}
    fam_plus = [l.split()[1] for l in open(f"{REF}/inst/GenoMat_SPAGxE_Plus.fam")]
    np.savez_compressed(
        f"{OUT}/plus.npz",
        Y=np.asarray(ph["Y"], dtype=np.float64), Cov1=np.asarray(ph["Cov1"], dtype=np.float64),
        Cov2=np.asarray(ph["Cov2"], dtype=np.float64), E=np.asarray(ph["E"], dtype=np.float64),
        grm_i=np.array([idx[s] for s in grm["ID1"]], dtype=np.int32),
        grm_j=np.array([idx[s] for s in grm["ID2"]], dtype=np.int32),
        grm_v=np.asarray(grm["Value"], dtype=np.float64),
        rda_first100_equals_bed_a1_count=np.array(first100_equal),
        fam_equals_pheno_iid=np.array(fam_plus == list(ph["IID"])),
        rda_maf_range=np.array([gm.value.mean(0).min() / 2, gm.value.mean(0).max() / 2]),
    )
    # ---- mix recipe inputs
    a6 = load_rdata(f"{REF}/simulation_studies/SPAGxEmixCCT/data/a.population.structure.6.RData")["a6"]
    pcs = load_rdata(f"{REF}/simulation_studies/SPAGxEmixCCT/data/top10PCs_MAFsumover0.1.RData")["top10PCs"]
    np.savez_compressed(f"{OUT}/mix.npz", a6=a6.value, top10PCs=pcs.value)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
    print("rda[,1:100] == bed A1-count decode:", first100_equal)


if __name__ == "__main__":
    main()
