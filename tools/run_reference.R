#!/usr/bin/env Rscript
# run_reference.R -- run the UNMODIFIED reference R functions on inputs written by the Python harness.
#
#   Rscript tools/run_reference.R <reference_dir> <input_dir> <output_file> [timing_file]
#
# <reference_dir>/R/*.R are source()d as they are (nothing from the reference is copied into this repo).
# <input_dir> holds raw little-endian doubles written by tests/test_vs_rscript.py / bench.py:
#   meta.txt   one line: mode trait n m p k epsilon cutoff min_maf      (mode = cct | mix | plus)
#   geno.bin   n x m column-major doubles, NaN = NA      R.bin E.bin Y.bin  length n      cova.bin  n x p
#   pcs.bin    n x k (mix)      grm_i.bin grm_j.bin grm_v.bin (plus; 1-based ids as doubles)  scal.bin (3)  rnew.bin (n)
# Output: the M x ncol result matrix as raw doubles (column-major), NA kept as NA_real_.
args <- commandArgs(trailingOnly = TRUE)
refdir <- args[1]; indir <- args[2]; outfile <- args[3]
timing <- if (length(args) >= 4) args[4] else NA

suppressWarnings(suppressMessages({
  ok_surv <- requireNamespace("survival", quietly = TRUE)
  if (ok_surv) library(survival)
  if (requireNamespace("dplyr", quietly = TRUE)) library(dplyr)
}))
for (f in list.files(file.path(refdir, "R"), pattern = "\\.R$", full.names = TRUE)) source(f)

meta <- strsplit(readLines(file.path(indir, "meta.txt"))[1], " +")[[1]]
mode <- meta[1]; trait <- meta[2]
n <- as.integer(meta[3]); m <- as.integer(meta[4]); p <- as.integer(meta[5]); k <- as.integer(meta[6])
epsilon <- as.numeric(meta[7]); cutoff <- as.numeric(meta[8]); min.maf <- as.numeric(meta[9])
rd <- function(name, len) readBin(file.path(indir, name), what = "double", n = len, size = 8, endian = "little")

G <- matrix(rd("geno.bin", n * m), n, m)
G[is.nan(G)] <- NA
ids <- paste0("IID-", seq_len(n))
rownames(G) <- ids; colnames(G) <- paste0("SNP-", seq_len(m))
R <- rd("R.bin", n); E <- rd("E.bin", n); Y <- rd("Y.bin", n)
Cova <- if (p > 0) matrix(rd("cova.bin", n * p), n, p) else matrix(0, n, 0)
if (p > 0) colnames(Cova) <- paste0("Cov", seq_len(p))
Phen <- data.frame(ID = ids, IID = ids, Y = Y)

t0 <- proc.time()[["elapsed"]]
if (mode == "cct") {
  out <- SPAGxE_CCT(traits = trait, Geno.mtx = G, R = R, E = E, Phen.mtx = Phen, Cova.mtx = Cova,
                    epsilon = epsilon, Cutoff = cutoff, min.maf = min.maf)
} else if (mode == "mix") {
  PCs <- matrix(rd("pcs.bin", n * k), n, k)
  out <- SPAGxEmix_CCT(traits = trait, Geno.mtx = G, R = R, E = E, Phen.mtx = Phen, Cova.mtx = Cova, topPCs = PCs,
                       epsilon = epsilon, Cutoff = cutoff, min.maf = min.maf)
} else if (mode == "plus") {
  nnz <- as.integer(meta[10])
  grm <- data.frame(ID1 = ids[as.integer(rd("grm_i.bin", nnz))], ID2 = ids[as.integer(rd("grm_j.bin", nnz))],
                    Value = rd("grm_v.bin", nnz), stringsAsFactors = FALSE)
  sc <- rd("scal.bin", 3)
  obj <- list(ResidMat = data.frame(SubjID = ids, Resid = R, RE = R * E), R = R, R_GRM_R = sc[1], R_GRM_RE = sc[2],
              R.new = rd("rnew.bin", n), R.new_GRM_R.new = sc[3], sparseGRM = grm)
  out <- SPAGxE_Plus(Geno.mtx = G, E = E, Phen.mtx = Phen, Cova.mtx = Cova, obj.SPAGxE_Plus_Nullmodel = obj,
                     epsilon = epsilon, Cutoff = cutoff, min.maf = min.maf)
} else stop("unknown mode")
dt <- proc.time()[["elapsed"]] - t0

writeBin(as.double(out), outfile, size = 8, endian = "little")
if (!is.na(timing)) writeLines(sprintf("%d %d %.6f", nrow(out), ncol(out), dt), timing)
