# SPAGxECCT_cuda.R -- drop-in replacements of the reference's exported Step-2 functions.
# Same formals, printed lines, column names and stop() behaviour as R/SPAGxECCT.R:455-530 (SPAGxE_CCT),
# :916-1002 (SPAGxEmix_CCT) and R/SPAGxEPlus_functions_MYZ.R:159-238 (SPAGxE_Plus); the serial
# `for(i in 1:n.Geno)` loop is replaced by one .Call into libspagxe_cuda.so.  Step 1
# (SPA_G_Get_Resid, SPAGxE_Plus_Nullmodel) is untouched and stays in R.
# dyn.load("spagxe_rshim.so") once (it links libspagxe_cuda.so).

.spagxe_trait <- function(traits) {
  if (traits == "binary") return(1L)
  if (traits == "quantitative") return(2L)
  stop('the CUDA path implements traits = "binary" and "quantitative" (survival / categorical Wald tests: use the reference)')
}

.spagxe_geno <- function(GenoFile, GenoFileIndex, control, Geno.mtx, SampleIDs) {
  if (!is.null(Geno.mtx)) return(list(g = Geno.mtx, n = nrow(Geno.mtx), m = ncol(Geno.mtx), snp = colnames(Geno.mtx), ids = rownames(Geno.mtx)))
  bim <- if (is.null(GenoFileIndex)) sub("\\.bed$", ".bim", GenoFile) else GenoFileIndex[1]
  fam <- if (is.null(GenoFileIndex)) sub("\\.bed$", ".fam", GenoFile) else GenoFileIndex[2]
  ids <- read.table(fam, stringsAsFactors = FALSE)[[2]]
  snp <- read.table(bim, stringsAsFactors = FALSE)[[2]]
  if (!identical(as.character(SampleIDs), as.character(ids)))
    stop("the CUDA .bed path needs Phen.mtx$ID in .fam order (or pass Geno.mtx); sample reordering is done by the Python host today")
  raw <- readBin(GenoFile, "raw", file.size(GenoFile))
  if (!identical(as.integer(raw[1:3]), c(0x6cL, 0x1bL, 0x01L))) stop("not a variant-major PLINK .bed file")
  list(g = raw[-(1:3)], n = length(ids), m = length(snp), snp = snp, ids = ids)
}

SPAGxE_CCT <- function(traits = "survival/binary/quantitative/categorical", GenoFile = NULL, GenoFileIndex = NULL,
                       control = list(AllMarkers = TRUE), Geno.mtx = NULL, R, E, Phen.mtx, Cova.mtx,
                       epsilon = 0.001, Cutoff = 2, impute.method = "fixed", missing.cutoff = 0.15,
                       min.maf = 0.001, G.model = "Add") {
  if (G.model != "Add") stop('the CUDA path implements G.model = "Add"')
  gi <- .spagxe_geno(GenoFile, GenoFileIndex, control, Geno.mtx, Phen.mtx$ID)
  check_input_Resid(pIDs = Phen.mtx$ID, gIDs = gi$ids, R = R)          # reference helper, unchanged (:1900-1913)
  print(paste0("Sample size is ", gi$n, "."))
  print(paste0("Number of variants is ", gi$m, "."))
  ctx <- .Call("spagxe_r_create", 0L)
  .Call("spagxe_r_set_vectors", ctx, as.double(R), as.double(E))
  .Call("spagxe_r_set_wald", ctx, .spagxe_trait(traits), as.double(Phen.mtx$Y), as.matrix(Cova.mtx) + 0)
  print("Start Analyzing...")
  print(Sys.time())
  output <- .Call("spagxe_r_run", ctx, 0L, gi$g, gi$n, gi$m, c(epsilon, Cutoff, min.maf, missing.cutoff, 0.05, 0.1))
  colnames(output) <- c("MAF", "missing.rate", "p.value.spaGxE", "p.value.spaGxE.Wald", "p.value.spaGxE.CCT.Wald",
                        "p.value.normGxE", "p.value.betaG", "Stat.betaG", "Var.betaG", "z.betaG")
  rownames(output) <- gi$snp
  print("Analysis Complete.")
  print(Sys.time())
  output
}

SPAGxE_CCT_one_SNP <- function(traits = "survival/binary/quantitative/categorical", g, R, E, Phen.mtx, Cova.mtx,
                               epsilon = 0.001, Cutoff = 2, impute.method = "fixed", missing.cutoff = 0.15,
                               min.maf = 0.001, G.model = "Add") {
  ctx <- .Call("spagxe_r_create", 0L)
  .Call("spagxe_r_set_vectors", ctx, as.double(R), as.double(E))
  .Call("spagxe_r_set_wald", ctx, .spagxe_trait(traits), as.double(Phen.mtx$Y), as.matrix(Cova.mtx) + 0)
  drop(.Call("spagxe_r_run", ctx, 0L, matrix(g, ncol = 1), length(g), 1L, c(epsilon, Cutoff, min.maf, missing.cutoff, 0.05, 0.1)))
}

SPAGxEmix_CCT <- function(traits = "survival/binary/quantitative/categorical", GenoFile = NULL, GenoFileIndex = NULL,
                          control = list(AllMarkers = TRUE), Geno.mtx = NULL, R, E, Phen.mtx, Cova.mtx, topPCs,
                          epsilon = 0.001, Cutoff = 2, impute.method = "fixed", missing.cutoff = 0.15, min.maf = 0.001,
                          G.model = "Add", topPCs.pvalue.cutoff = 0.05, MAF.est.negative.ratio.cutoff = 0.1) {
  gi <- .spagxe_geno(GenoFile, GenoFileIndex, control, Geno.mtx, Phen.mtx$ID)
  print(paste0("Sample size is ", gi$n, "."))
  print(paste0("Number of variants is ", gi$m, "."))
  ctx <- .Call("spagxe_r_create", 0L)
  .Call("spagxe_r_set_vectors", ctx, as.double(R), as.double(E))
  .Call("spagxe_r_set_wald", ctx, .spagxe_trait(traits), as.double(Phen.mtx$Y), as.matrix(Cova.mtx) + 0)
  .Call("spagxe_r_set_pcs", ctx, as.matrix(topPCs) + 0)
  print("Start Analyzing...")
  print(Sys.time())
  output <- .Call("spagxe_r_run", ctx, 1L, gi$g, gi$n, gi$m,
                  c(epsilon, Cutoff, min.maf, missing.cutoff, topPCs.pvalue.cutoff, MAF.est.negative.ratio.cutoff))
  colnames(output) <- c("MAF", "missing.rate", "p.value.spaGxE", "p.value.spaGxE.Wald", "p.value.spaGxE.CCT.Wald",
                        "p.value.normGxE", "p.value.betaG", "Stat.betaG", "Var.betaG", "z.betaG",
                        "MAF.est.negative.num", "MAC")
  rownames(output) <- gi$snp
  print("Analysis Complete.")
  print(Sys.time())
  output
}

SPAGxE_Plus <- function(GenoFile = NULL, GenoFileIndex = NULL, control = list(AllMarkers = TRUE), Geno.mtx = NULL, E,
                        Phen.mtx, Cova.mtx, obj.SPAGxE_Plus_Nullmodel, epsilon = 0.001, Cutoff = 2,
                        impute.method = "fixed", missing.cutoff = 0.15, min.maf = 0.001, G.model = "Add") {
  obj <- obj.SPAGxE_Plus_Nullmodel
  gi <- .spagxe_geno(GenoFile, GenoFileIndex, control, Geno.mtx, Phen.mtx$ID)
  print(paste0("Sample size is ", gi$n, "."))
  print(paste0("Number of variants is ", gi$m, "."))
  ids <- obj$ResidMat$SubjID
  ctx <- .Call("spagxe_r_create", 0L)
  .Call("spagxe_r_set_vectors", ctx, as.double(obj$R), as.double(E))
  .Call("spagxe_r_set_grm", ctx, match(obj$sparseGRM$ID1, ids) - 1L, match(obj$sparseGRM$ID2, ids) - 1L,
        as.double(obj$sparseGRM$Value), c(obj$R_GRM_R, obj$R_GRM_RE, obj$R.new_GRM_R.new), as.double(obj$R.new))
  print("Start Analyzing...")
  print(Sys.time())
  output <- .Call("spagxe_r_run", ctx, 2L, gi$g, gi$n, gi$m, c(epsilon, Cutoff, min.maf, missing.cutoff, 0.05, 0.1))
  colnames(output) <- c("MAF", "missing.rate", "p.value.spaGxE.plus", "p.value.normGxE.plus", "p.value.betaG",
                        "Stat.betaG", "Var.betaG", "z.betaG")
  rownames(output) <- gi$snp
  print("Analysis Complete.")
  print(Sys.time())
  output
}
