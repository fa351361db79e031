# SPAGxECCT_cuda.R -- drop-in replacements of the reference's exported Step-2 functions.
# Same formals, printed lines, column names and stop() behaviour as R/SPAGxECCT.R:455-530 (SPAGxE_CCT), :543-683
# (SPAGxE_CCT_one_SNP), :916-1002 (SPAGxEmix_CCT), :1013-1269 (SPAGxEmix_CCT_one_SNP),
# R/SPAmixGxECCTPlus_functions_MYZ_2024-09-17.R:38-133, :144-427 (SPAGxEmix_CCT_Plus, SPAGxEmix_CCT_Plus_one_SNP) and
# R/SPAGxEPlus_functions_MYZ.R:159-238, :248-405 (SPAGxE_Plus, SPAGxE_Plus_one_SNP); the serial `for(i in 1:n.Geno)` loop
# is replaced by one .Call into libspagxe_cuda.so.  Step 1 (SPA_G_Get_Resid, the model fit of SPAGxE_Plus_Nullmodel) stays
# in R.  dyn.load("spagxe_rshim.so") once (it links libspagxe_cuda.so).
#
# Additions to the reference's formals (all optional, defaults reproduce the reference call):
#   n.gpus   number of GPUs of the node to spread the SNPs over (SPAGxE_CCT, SPAGxEmix_CCT, SPAGxE_Plus; one NCCL broadcast
#            of R/E/Y/Cova)
#   device   CUDA ordinal for single-GPU calls
# The result carries attr(, "diag"): routing counts and device times of the call.

.spagxe_trait <- function(traits) {
  if (traits == "binary") return(1L)
  if (traits == "quantitative") return(2L)
  if (traits == "survival") return(3L)
  if (traits == "categorical") return(4L)
  stop('traits must be "survival", "binary", "quantitative" or "categorical"')
}

# Phen.mtx$Y as the C ABI takes it; categorical (R/SPAGxECCT.R:666): the level index of as.factor(Y), 0 .. J-1
.spagxe_wald_y <- function(traits, Phen.mtx) {
  if (.spagxe_trait(traits) == 4L) return(as.double(as.integer(as.factor(Phen.mtx$Y)) - 1L))
  as.double(Phen.mtx$Y)
}

# Wald data of branch B (R/SPAGxECCT.R:622-677): Phen.mtx$Y, or Phen.mtx$surv.time / $event for a survival trait (:624)
.spagxe_set_wald <- function(ctx, traits, Phen.mtx, Cova.mtx) {
  if (.spagxe_trait(traits) == 3L)
    .Call("spagxe_r_set_wald_survival", ctx, as.double(Phen.mtx$surv.time), as.double(Phen.mtx$event), as.matrix(Cova.mtx) + 0)
  else
    .Call("spagxe_r_set_wald", ctx, .spagxe_trait(traits), .spagxe_wald_y(traits, Phen.mtx), as.matrix(Cova.mtx) + 0)
}

# R/SPAGxECCT.R:572-574: "Dom" / "Rec" recode g after imputation and the flip; any other string acts like "Add"
.spagxe_gmodel <- function(G.model) {
  if (G.model == "Dom") return(1L)
  if (G.model == "Rec") return(2L)
  0L
}

.spagxe_read_lines <- function(f) { x <- trimws(readLines(f)); x[nzchar(x)] }

# GRAB.ReadGeno's part of the job for PLINK input (R/SPAGxECCT.R:484-487) WITHOUT building the N x M double matrix:
# the packed rows stay packed, the SampleIDs subset / reorder becomes a 0-based gather index that the GPU applies,
# `control` selects marker rows and the allele order.
.spagxe_geno <- function(GenoFile, GenoFileIndex, control, Geno.mtx, SampleIDs) {
  if (!is.null(Geno.mtx))
    return(list(g = Geno.mtx, n = nrow(Geno.mtx), m = ncol(Geno.mtx), snp = colnames(Geno.mtx), ids = rownames(Geno.mtx),
                sidx = NULL, nfile = 0, a2 = 0L))
  if (!grepl("\\.bed$", GenoFile)) stop("GenoFile must be a PLINK .bed file (BGEN input is not implemented by the CUDA path)")
  bim <- if (is.null(GenoFileIndex)) sub("\\.bed$", ".bim", GenoFile) else GenoFileIndex[1]
  fam <- if (is.null(GenoFileIndex)) sub("\\.bed$", ".fam", GenoFile) else GenoFileIndex[2]
  fam.ids <- as.character(read.table(fam, stringsAsFactors = FALSE)[[2]])
  bimtab <- read.table(bim, stringsAsFactors = FALSE)
  snp <- as.character(bimtab[[2]])
  nfile <- length(fam.ids); mfile <- length(snp); rb <- (nfile + 3) %/% 4
  known <- c("AllMarkers", "IDsToIncludeFile", "IDsToExcludeFile", "RangesToIncludeFile", "RangesToExcludeFile",
             "AlleleOrder", "ImputeMethod")
  if (length(setdiff(names(control), known))) stop("unsupported control option(s): ", paste(setdiff(names(control), known), collapse = ", "))
  if (!is.null(control$ImputeMethod) && control$ImputeMethod != "none")
    stop('control$ImputeMethod other than "none": the Step-2 functions impute themselves (impute.method)')
  order <- if (is.null(control$AlleleOrder)) "alt-first" else control$AlleleOrder
  if (!order %in% c("alt-first", "ref-first")) stop("control$AlleleOrder should be 'ref-first' or 'alt-first'.")
  inc <- intersect(names(control), c("IDsToIncludeFile", "RangesToIncludeFile"))
  exc <- intersect(names(control), c("IDsToExcludeFile", "RangesToExcludeFile"))
  if (length(inc) && length(exc)) stop("We currently do not support both 'IncludeFile' and 'ExcludeFile' at the same time.")
  sel <- NULL
  if (length(inc) || length(exc)) {
    hit <- rep(FALSE, mfile)
    for (k in c("IDsToIncludeFile", "IDsToExcludeFile")) if (!is.null(control[[k]])) hit <- hit | snp %in% .spagxe_read_lines(control[[k]])
    for (k in c("RangesToIncludeFile", "RangesToExcludeFile")) if (!is.null(control[[k]])) {
      rg <- read.table(control[[k]], stringsAsFactors = FALSE)
      for (r in seq_len(nrow(rg))) hit <- hit | (as.character(bimtab[[1]]) == as.character(rg[r, 1]) & bimtab[[4]] >= rg[r, 2] & bimtab[[4]] <= rg[r, 3])
    }
    sel <- which(if (length(inc)) hit else !hit)
  } else if (!isTRUE(control$AllMarkers)) stop("If 'AllMarkers' is FALSE, at least one of the include / exclude files should be specified.")
  con <- file(GenoFile, "rb"); on.exit(close(con))
  magic <- readBin(con, "raw", 3)
  if (!identical(as.integer(magic), c(0x6cL, 0x1bL, 0x01L))) stop("not a variant-major PLINK .bed file")
  if (is.null(sel)) raw <- readBin(con, "raw", rb * mfile)
  else {                                      # only the selected marker rows are read
    raw <- raw(rb * length(sel))
    for (k in seq_along(sel)) { seek(con, 3 + (sel[k] - 1) * rb); raw[((k - 1) * rb + 1):(k * rb)] <- readBin(con, "raw", rb) }
    snp <- snp[sel]
  }
  ids <- fam.ids; sidx <- NULL
  if (!is.null(SampleIDs) && !identical(as.character(SampleIDs), fam.ids)) {
    pos <- match(as.character(SampleIDs), fam.ids)
    if (anyNA(pos)) stop("SampleIDs not found in the .fam file: ", paste(head(SampleIDs[is.na(pos)], 3), collapse = ", "))
    sidx <- as.double(pos - 1)               # 0-based file positions; the gather runs on the GPU
    ids <- as.character(SampleIDs)
  }
  list(g = raw, n = length(ids), m = length(snp), snp = snp, ids = ids, sidx = sidx, nfile = nfile, a2 = as.integer(order == "ref-first"))
}

.spagxe_run <- function(ctx, which, gi, prm) {
  .Call("spagxe_r_run", ctx, which, gi$g, gi$n, gi$m, prm, gi$sidx, as.double(gi$nfile), gi$a2)
}

SPAGxE_CCT <- function(traits = "survival/binary/quantitative/categorical", GenoFile = NULL, GenoFileIndex = NULL,
                       control = list(AllMarkers = TRUE), Geno.mtx = NULL, R, E, Phen.mtx, Cova.mtx,
                       epsilon = 0.001, Cutoff = 2, impute.method = "fixed", missing.cutoff = 0.15,
                       min.maf = 0.001, G.model = "Add", n.gpus = 1L, device = 0L) {
  gm <- .spagxe_gmodel(G.model)
  gi <- .spagxe_geno(GenoFile, GenoFileIndex, control, Geno.mtx, Phen.mtx$ID)
  check_input_Resid(pIDs = Phen.mtx$ID, gIDs = gi$ids, R = R)          # reference helper, unchanged (:1900-1913)
  print(paste0("Sample size is ", gi$n, "."))
  print(paste0("Number of variants is ", gi$m, "."))
  prm <- c(epsilon, Cutoff, min.maf, missing.cutoff, 0.05, 0.1, gm)
  print("Start Analyzing...")
  print(Sys.time())
  if (n.gpus > 1L) {
    if (.spagxe_trait(traits) == 3L) stop('traits = "survival" runs on one GPU')
    mg <- .Call("spagxe_r_mgpu_create", as.integer(n.gpus))
    on.exit(.Call("spagxe_r_mgpu_destroy", mg), add = TRUE)
    .Call("spagxe_r_mgpu_set_inputs", mg, as.double(R), as.double(E), .spagxe_trait(traits), .spagxe_wald_y(traits, Phen.mtx), as.matrix(Cova.mtx) + 0)
    output <- .Call("spagxe_r_mgpu_run_cct", mg, gi$g, gi$n, gi$m, prm, gi$sidx, as.double(gi$nfile), gi$a2)
  } else {
    ctx <- .Call("spagxe_r_create", as.integer(device))
    on.exit(.Call("spagxe_r_destroy", ctx), add = TRUE)                 # GPU memory is released when the call returns
    .Call("spagxe_r_set_vectors", ctx, as.double(R), as.double(E))
    .spagxe_set_wald(ctx, traits, Phen.mtx, Cova.mtx)
    output <- .spagxe_run(ctx, 0L, gi, prm)
  }
  colnames(output) <- c("MAF", "missing.rate", "p.value.spaGxE", "p.value.spaGxE.Wald", "p.value.spaGxE.CCT.Wald",
                        "p.value.normGxE", "p.value.betaG", "Stat.betaG", "Var.betaG", "z.betaG")
  rownames(output) <- gi$snp
  print("Analysis Complete.")
  print(Sys.time())
  output
}

SPAGxE_CCT_one_SNP <- function(traits = "survival/binary/quantitative/categorical", g, R, E, Phen.mtx, Cova.mtx,
                               epsilon = 0.001, Cutoff = 2, impute.method = "fixed", missing.cutoff = 0.15,
                               min.maf = 0.001, G.model = "Add", device = 0L) {
  gm <- .spagxe_gmodel(G.model)
  ctx <- .Call("spagxe_r_create", as.integer(device))
  on.exit(.Call("spagxe_r_destroy", ctx), add = TRUE)
  .Call("spagxe_r_set_vectors", ctx, as.double(R), as.double(E))
  .spagxe_set_wald(ctx, traits, Phen.mtx, Cova.mtx)
  gi <- list(g = matrix(g, ncol = 1), n = length(g), m = 1L, sidx = NULL, nfile = 0, a2 = 0L)
  out <- .spagxe_run(ctx, 0L, gi, c(epsilon, Cutoff, min.maf, missing.cutoff, 0.05, 0.1, gm))
  as.numeric(out)                                                       # numeric(10), attributes dropped like c(...)
}

SPAGxEmix_CCT <- function(traits = "survival/binary/quantitative/categorical", GenoFile = NULL, GenoFileIndex = NULL,
                          control = list(AllMarkers = TRUE), Geno.mtx = NULL, R, E, Phen.mtx, Cova.mtx, topPCs,
                          epsilon = 0.001, Cutoff = 2, impute.method = "fixed", missing.cutoff = 0.15, min.maf = 0.001,
                          G.model = "Add", topPCs.pvalue.cutoff = 0.05, MAF.est.negative.ratio.cutoff = 0.1, n.gpus = 1L,
                          device = 0L) {
  gm <- .spagxe_gmodel(G.model)
  gi <- .spagxe_geno(GenoFile, GenoFileIndex, control, Geno.mtx, Phen.mtx$ID)   # no ID check: commented out in the reference (:955)
  print(paste0("Sample size is ", gi$n, "."))
  print(paste0("Number of variants is ", gi$m, "."))
  prm <- c(epsilon, Cutoff, min.maf, missing.cutoff, topPCs.pvalue.cutoff, MAF.est.negative.ratio.cutoff, gm)
  print("Start Analyzing...")
  print(Sys.time())
  if (n.gpus > 1L) {
    mg <- .Call("spagxe_r_mgpu_create", as.integer(n.gpus))
    on.exit(.Call("spagxe_r_mgpu_destroy", mg), add = TRUE)
    if (.spagxe_trait(traits) == 3L) {
      .Call("spagxe_r_mgpu_set_vectors", mg, as.double(R), as.double(E))
      .Call("spagxe_r_mgpu_set_wald_survival", mg, as.double(Phen.mtx$surv.time), as.double(Phen.mtx$event), as.matrix(Cova.mtx) + 0)
    } else {
      .Call("spagxe_r_mgpu_set_inputs", mg, as.double(R), as.double(E), .spagxe_trait(traits), .spagxe_wald_y(traits, Phen.mtx), as.matrix(Cova.mtx) + 0)
    }
    .Call("spagxe_r_mgpu_set_pcs", mg, as.matrix(topPCs) + 0)
    output <- .Call("spagxe_r_mgpu_run", mg, 1L, gi$g, gi$n, gi$m, prm, gi$sidx, as.double(gi$nfile), gi$a2)
  } else {
    ctx <- .Call("spagxe_r_create", as.integer(device))
    on.exit(.Call("spagxe_r_destroy", ctx), add = TRUE)
    .Call("spagxe_r_set_vectors", ctx, as.double(R), as.double(E))
    .spagxe_set_wald(ctx, traits, Phen.mtx, Cova.mtx)
    .Call("spagxe_r_set_pcs", ctx, as.matrix(topPCs) + 0)
    output <- .spagxe_run(ctx, 1L, gi, prm)
  }
  colnames(output) <- c("MAF", "missing.rate", "p.value.spaGxE", "p.value.spaGxE.Wald", "p.value.spaGxE.CCT.Wald",
                        "p.value.normGxE", "p.value.betaG", "Stat.betaG", "Var.betaG", "z.betaG",
                        "MAF.est.negative.num", "MAC")
  rownames(output) <- gi$snp
  print("Analysis Complete.")
  print(Sys.time())
  output
}

SPAGxEmix_CCT_one_SNP <- function(traits = "survival/binary/quantitative/categorical", g, R, E, Phen.mtx, Cova.mtx, topPCs,
                                  epsilon = 0.001, Cutoff = 2, impute.method = "fixed", missing.cutoff = 0.15,
                                  min.maf = 0.001, G.model = "Add", topPCs.pvalue.cutoff = 0.05,
                                  MAF.est.negative.ratio.cutoff = 0.1, device = 0L) {
  gm <- .spagxe_gmodel(G.model)
  ctx <- .Call("spagxe_r_create", as.integer(device))
  on.exit(.Call("spagxe_r_destroy", ctx), add = TRUE)
  .Call("spagxe_r_set_vectors", ctx, as.double(R), as.double(E))
  .spagxe_set_wald(ctx, traits, Phen.mtx, Cova.mtx)
  .Call("spagxe_r_set_pcs", ctx, as.matrix(topPCs) + 0)
  gi <- list(g = matrix(g, ncol = 1), n = length(g), m = 1L, sidx = NULL, nfile = 0, a2 = 0L)
  as.numeric(.spagxe_run(ctx, 1L, gi, c(epsilon, Cutoff, min.maf, missing.cutoff, topPCs.pvalue.cutoff, MAF.est.negative.ratio.cutoff, gm)))
}

.spagxe_set_plus <- function(ctx, obj, E) {
  ids <- as.character(obj$ResidMat$SubjID)
  .Call("spagxe_r_set_vectors", ctx, as.double(obj$R), as.double(E))
  .Call("spagxe_r_set_grm", ctx, match(as.character(obj$sparseGRM$ID1), ids) - 1L, match(as.character(obj$sparseGRM$ID2), ids) - 1L,
        as.double(obj$sparseGRM$Value), c(obj$R_GRM_R, obj$R_GRM_RE, obj$R.new_GRM_R.new), as.double(obj$R.new))
}

SPAGxE_Plus <- function(GenoFile = NULL, GenoFileIndex = NULL, control = list(AllMarkers = TRUE), Geno.mtx = NULL, E,
                        Phen.mtx, Cova.mtx, obj.SPAGxE_Plus_Nullmodel, epsilon = 0.001, Cutoff = 2,
                        impute.method = "fixed", missing.cutoff = 0.15, min.maf = 0.001, G.model = "Add", n.gpus = 1L,
                        device = 0L) {
  gm <- .spagxe_gmodel(G.model)
  obj <- obj.SPAGxE_Plus_Nullmodel
  gi <- .spagxe_geno(GenoFile, GenoFileIndex, control, Geno.mtx, Phen.mtx$ID)
  print(paste0("Sample size is ", gi$n, "."))
  print(paste0("Number of variants is ", gi$m, "."))
  prm <- c(epsilon, Cutoff, min.maf, missing.cutoff, 0.05, 0.1, gm)
  print("Start Analyzing...")
  print(Sys.time())
  if (n.gpus > 1L) {
    ids <- as.character(obj$ResidMat$SubjID)
    mg <- .Call("spagxe_r_mgpu_create", as.integer(n.gpus))
    on.exit(.Call("spagxe_r_mgpu_destroy", mg), add = TRUE)
    .Call("spagxe_r_mgpu_set_vectors", mg, as.double(obj$R), as.double(E))
    .Call("spagxe_r_mgpu_set_grm", mg, match(as.character(obj$sparseGRM$ID1), ids) - 1L, match(as.character(obj$sparseGRM$ID2), ids) - 1L,
          as.double(obj$sparseGRM$Value), c(obj$R_GRM_R, obj$R_GRM_RE, obj$R.new_GRM_R.new), as.double(obj$R.new))
    output <- .Call("spagxe_r_mgpu_run", mg, 2L, gi$g, gi$n, gi$m, prm, gi$sidx, as.double(gi$nfile), gi$a2)
  } else {
    ctx <- .Call("spagxe_r_create", as.integer(device))
    on.exit(.Call("spagxe_r_destroy", ctx), add = TRUE)
    .spagxe_set_plus(ctx, obj, E)
    output <- .spagxe_run(ctx, 2L, gi, prm)
  }
  colnames(output) <- c("MAF", "missing.rate", "p.value.spaGxE.plus", "p.value.normGxE.plus", "p.value.betaG",
                        "Stat.betaG", "Var.betaG", "z.betaG")
  rownames(output) <- gi$snp
  print("Analysis Complete.")
  print(Sys.time())
  output
}

SPAGxE_Plus_one_SNP <- function(g, E, Phen.mtx, Cova.mtx, obj.SPAGxE_Plus_Nullmodel, epsilon = 0.001, Cutoff = 2,
                                impute.method = "fixed", missing.cutoff = 0.15, min.maf = 0.001, G.model = "Add", device = 0L) {
  gm <- .spagxe_gmodel(G.model)
  ctx <- .Call("spagxe_r_create", as.integer(device))
  on.exit(.Call("spagxe_r_destroy", ctx), add = TRUE)
  .spagxe_set_plus(ctx, obj.SPAGxE_Plus_Nullmodel, E)
  gi <- list(g = matrix(g, ncol = 1), n = length(g), m = 1L, sidx = NULL, nfile = 0, a2 = 0L)
  as.numeric(.spagxe_run(ctx, 2L, gi, c(epsilon, Cutoff, min.maf, missing.cutoff, 0.05, 0.1, gm)))
}

# SPAGxEmix_CCT_Plus / SPAGxEmix_CCT_Plus_one_SNP (R/SPAmixGxECCTPlus_functions_MYZ_2024-09-17.R:38-133, :144-427).  The GPU does
# the allele-frequency model, every sparse-GRM variance and the adjusted saddlepoint for all SNPs; for the few SNPs of
# branch B (p.value.betaG <= epsilon) the Wald test is the reference's own lme4::glmer / lmer call (:384, :404), run here
# in R on the recoded genotype vector, followed by the reference's CCT().  R = ResidMat$Resid (the reference's one-SNP
# function reads `R` as a free variable that its scripts define this way).
.spagxe_mixplus_prepare <- function(ctx, ResidMat, E, topPCs, sparseGRM) {
  ids <- as.character(ResidMat$SubjID)
  id1 <- as.character(sparseGRM$ID1); id2 <- as.character(sparseGRM$ID2)
  if (any(!ids %in% unique(c(id1, id2)))) stop("At least one subject in residual matrix does not have GRM information.")
  g1 <- match(id1, ids); g2 <- match(id2, ids)
  keep <- !is.na(g1) & !is.na(g2)                                       # filter(ID1 %in% SubjID & ID2 %in% SubjID), :220
  .Call("spagxe_r_set_vectors", ctx, as.double(ResidMat$Resid), as.double(E))
  .Call("spagxe_r_set_pcs", ctx, as.matrix(topPCs) + 0)
  .Call("spagxe_r_set_sparse_grm", ctx, g1[keep] - 1L, g2[keep] - 1L, as.double(sparseGRM$Value[keep]))
}

# the recoding of :171-188 for one column, then the Wald fit of :384-389 / :404-409 and CCT (:393, :413)
.spagxe_mixplus_wald <- function(traits, g, row, E, Phen.mtx, Cova.mtx, G.model) {
  MAF <- mean(g, na.rm = TRUE) / 2
  g[is.na(g)] <- 2 * MAF
  if (MAF > 0.5) g <- 2 - g
  if (G.model == "Dom") g <- ifelse(g >= 1, 1, 0)
  if (G.model == "Rec") g <- ifelse(g <= 1, 0, 1)
  if (traits == "binary") {
    data0 <- lme4::glmer(Y ~ as.matrix(Cova.mtx) + E + g + g * E + (1 | IID), family = "binomial", data = Phen.mtx)
    pval.wald <- summary(data0)$coefficients["E:g", "Pr(>|z|)"]
  } else {
    data0 <- lme4::lmer(Y ~ as.matrix(Cova.mtx) + E + g + g * E + (1 | IID), data = Phen.mtx)
    pval.wald <- summary(data0)$coefficients["E:g", "Pr(>|t|)"]
  }
  if (is.na(pval.wald)) pval.wald <- row[3]
  row[4] <- pval.wald
  row[5] <- CCT(c(row[3], pval.wald))
  row
}

SPAGxEmix_CCT_Plus <- function(traits = "binary/quantitative", Geno.mtx, ResidMat, E, Phen.mtx, Cova.mtx, topPCs, sparseGRM,
                               epsilon = 0.001, Cutoff = 2, impute.method = "fixed", missing.cutoff = 0.15, min.maf = 0.001,
                               G.model = "Add", topPCs.pvalue.cutoff = 0.05, MAF.est.negative.ratio.cutoff = 0.1, device = 0L) {
  gm <- .spagxe_gmodel(G.model)
  print(paste0("Sample size is ", nrow(Geno.mtx), "."))
  print(paste0("Number of variants is ", ncol(Geno.mtx), "."))
  ctx <- .Call("spagxe_r_create", as.integer(device))
  on.exit(.Call("spagxe_r_destroy", ctx), add = TRUE)
  .spagxe_mixplus_prepare(ctx, ResidMat, E, topPCs, sparseGRM)
  print("Start Analyzing...")
  print(Sys.time())
  gi <- list(g = Geno.mtx, n = nrow(Geno.mtx), m = ncol(Geno.mtx), sidx = NULL, nfile = 0, a2 = 0L)
  output <- .spagxe_run(ctx, 3L, gi, c(epsilon, Cutoff, min.maf, missing.cutoff, topPCs.pvalue.cutoff, MAF.est.negative.ratio.cutoff, gm))
  for (i in seq_len(ncol(Geno.mtx))) print(i)                            # the reference prints the SNP counter (:103)
  for (i in which(output[, 7] <= epsilon))
    output[i, ] <- .spagxe_mixplus_wald(traits, Geno.mtx[, i], output[i, ], E, Phen.mtx, Cova.mtx, G.model)
  colnames(output) <- c("MAF", "missing.rate", "p.value.spaGxEmix.plus", "p.value.spaGxEmix.plus.Wald",
                        "p.value.spaGxEmix.plus.CCT.Wald", "p.value.normGxEmix.plus", "p.value.betaG", "Stat.betaG",
                        "Var.betaG", "z.betaG", "MAF.est.negative.num", "MAC")
  rownames(output) <- colnames(Geno.mtx)
  print("Analysis Complete.")
  print(Sys.time())
  output
}

SPAGxEmix_CCT_Plus_one_SNP <- function(traits = "binary/quantitative", g, ResidMat, E, Phen.mtx, Cova.mtx, topPCs, sparseGRM,
                                       epsilon = 0.001, Cutoff = 2, impute.method = "fixed", missing.cutoff = 0.15,
                                       min.maf = 0.001, G.model = "Add", topPCs.pvalue.cutoff = 0.05,
                                       MAF.est.negative.ratio.cutoff = 0.1, device = 0L) {
  gm <- .spagxe_gmodel(G.model)
  ctx <- .Call("spagxe_r_create", as.integer(device))
  on.exit(.Call("spagxe_r_destroy", ctx), add = TRUE)
  .spagxe_mixplus_prepare(ctx, ResidMat, E, topPCs, sparseGRM)
  gi <- list(g = matrix(g, ncol = 1), n = length(g), m = 1L, sidx = NULL, nfile = 0, a2 = 0L)
  row <- as.numeric(.spagxe_run(ctx, 3L, gi, c(epsilon, Cutoff, min.maf, missing.cutoff, topPCs.pvalue.cutoff, MAF.est.negative.ratio.cutoff, gm)))
  if (!is.na(row[7]) && row[7] <= epsilon) row <- .spagxe_mixplus_wald(traits, g, row, E, Phen.mtx, Cova.mtx, G.model)
  row
}

# The GRM quadratic forms of SPAGxE_Plus_Nullmodel (R/SPAGxEPlus_functions_MYZ.R:490-537) on the GPU: call this in place of
# the dplyr match()/mutate() block after the model fit; `resid` are the residuals, SubjID their ids.
SPAGxE_Plus_Nullmodel_GRM <- function(resid, E, SubjID, sparseGRM, device = 0L) {
  ids <- as.character(SubjID)
  g1 <- match(as.character(sparseGRM$ID1), ids); g2 <- match(as.character(sparseGRM$ID2), ids)
  if (any(!ids %in% unique(c(as.character(sparseGRM$ID1), as.character(sparseGRM$ID2)))))
    stop("At least one subject in residual matrix does not have GRM information.")
  keep <- !is.na(g1) & !is.na(g2)                                       # filter(ID1 %in% SubjID & ID2 %in% SubjID)
  ctx <- .Call("spagxe_r_create", as.integer(device))
  on.exit(.Call("spagxe_r_destroy", ctx), add = TRUE)
  res <- .Call("spagxe_r_plus_nullmodel", ctx, as.double(resid), as.double(E), g1[keep] - 1L, g2[keep] - 1L, as.double(sparseGRM$Value[keep]))
  list(ResidMat = data.frame(SubjID = SubjID, Resid = resid, RE = resid * E, R.new = res$R.new), R = resid,
       R_GRM_R = res$scalars[1], R_GRM_RE = res$scalars[2], R.new = res$R.new, R.new_GRM_R.new = res$scalars[3],
       sparseGRM = sparseGRM[keep, ])
}

# R/SPAGxECCT.R:1401-1481: ancestry-specific G x E test with local ancestry (binary and quantitative traits on the GPU)
SPAGxEmixCCT_localance <- function(traits = "survival/binary/quantitative/categorical", Geno.mtx, haplo.mtx, R, Cutoff = 2, E,
                                   Phen.mtx, Cova.mtx, Cova.haplo.mtx.list, epsilon = 0.001, impute.method = "fixed",
                                   missing.cutoff = 0.15, min.maf = 0.0001, G.model = "Add", device = 0L) {
  if (!(.spagxe_trait(traits) %in% c(1L, 2L))) stop('SPAGxEmixCCT_localance on the GPU: traits = "binary" or "quantitative"')
  gm <- .spagxe_gmodel(G.model)
  print(paste0("Sample size is ", nrow(Geno.mtx), "."))
  print(paste0("Number of variants is ", ncol(Geno.mtx), "."))
  ctx <- .Call("spagxe_r_create", as.integer(device))
  on.exit(.Call("spagxe_r_destroy", ctx), add = TRUE)
  .Call("spagxe_r_set_vectors", ctx, as.double(R), as.double(E))
  .spagxe_set_wald(ctx, traits, Phen.mtx, Cova.mtx)
  print("Start Analyzing...")
  print(Sys.time())
  output <- .Call("spagxe_r_run_localance", ctx, as.matrix(Geno.mtx) + 0, as.matrix(haplo.mtx) + 0,
                  lapply(unname(Cova.haplo.mtx.list), function(x) as.matrix(x) + 0),
                  c(epsilon, Cutoff, min.maf, missing.cutoff, 0.05, 0.1, gm))
  colnames(output) <- c("MAF", "missing.rate", "p.value.spaGxE.index.ance", "p.value.spaGxE.Wald.index.ance",
                        "p.value.spaGxE.CCT.Wald.index.ance", "p.value.normGxE.index.ance", "p.value.betaG.index.ance",
                        "Stat.betaG.index.ance", "Var.betaG.index.ance", "z.betaG.index.ance")
  rownames(output) <- colnames(Geno.mtx)
  output <- as.data.frame(cbind(rsID = colnames(Geno.mtx), output))   # as the reference returns it (:1478)
  print("Analysis Complete.")
  print(Sys.time())
  output
}

SPAGxEmixCCT_localance_one_SNP <- function(traits = "survival/binary/quantitative/categorical", g, haplo_numVec, R, E, Phen.mtx,
                                           Cova.mtx, Cova.haplo.mtx, epsilon = 0.001, Cutoff = 2, impute.method = "fixed",
                                           missing.cutoff = 0.15, min.maf = 0.0001, G.model = "Add", device = 0L) {
  if (!(.spagxe_trait(traits) %in% c(1L, 2L))) stop('SPAGxEmixCCT_localance on the GPU: traits = "binary" or "quantitative"')
  gm <- .spagxe_gmodel(G.model)
  ctx <- .Call("spagxe_r_create", as.integer(device))
  on.exit(.Call("spagxe_r_destroy", ctx), add = TRUE)
  .Call("spagxe_r_set_vectors", ctx, as.double(R), as.double(E))
  .spagxe_set_wald(ctx, traits, Phen.mtx, Cova.mtx)
  ch <- as.matrix(Cova.haplo.mtx) + 0
  output <- .Call("spagxe_r_run_localance", ctx, matrix(as.double(g), ncol = 1), matrix(as.double(haplo_numVec), ncol = 1),
                  lapply(seq_len(ncol(ch)), function(l) ch[, l, drop = FALSE]),
                  c(epsilon, Cutoff, min.maf, missing.cutoff, 0.05, 0.1, gm))
  as.numeric(output[1, ])
}
