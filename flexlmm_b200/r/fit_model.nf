// Drop-in replacement for modules/local/r/fit_model.nf of birneylab/flexlmm: same process name, same inputs, same
// outputs.  The R script keeps every file-format concern of the reference (readRDS, pvar/psam parsing, sprintf TSV
// lines, versions.yml); the per-SNP loop (fit_model.nf:120-171) and the permutation prologue (:96-109) are one .Call
// into libflexlmm_cuda.so.  `perm_seed` may be a single seed (reference behaviour) or a list of seeds: all of them are
// then tested in one task and the output holds one line per seed, which CAT_PERM concatenates unchanged
// (get_null_p_dist.nf:24-25).
process FIT_MODEL {
    tag "$meta.id"
    label 'process_gpu'

    // the pgenlibr image + libflexlmm_cuda.so + flexlmm_rshim.so (see INTEGRATION.md)
    container "${params.flexlmm_cuda_container}"

    input:
    tuple val(meta), path(mm), path(var_idx), path(null_model_fit), val(perm_seed)
    tuple val(meta2), path(pgen), path(pvar), path(psam)
    path model
    tuple val(meta3), path(model_frame)
    val use_dosage

    output:
    tuple val(meta), path("*.${perm_seed ? 'perm.tsv.gz' : 'tsv.gwas.gz'}") , emit: out
    path "versions.yml" , emit: versions

    when:
    task.ext.when == null || task.ext.when

    script:
    def prefix     = task.ext.prefix ?: "${meta.id}"
    def do_permute = perm_seed ? "TRUE" : "FALSE"
    def seeds      = perm_seed ? "c(${([perm_seed].flatten()).join(',')})" : "integer(0)"
    def dosage     = use_dosage ? "TRUE" : "FALSE"
    """
    #!/usr/bin/env Rscript
    dyn.load(Sys.getenv("FLEXLMM_RSHIM", "flexlmm_rshim.so"))

    l.mm <- readRDS("${mm}")
    var_idx <- readRDS("${var_idx}")
    l.null <- readRDS("${null_model_fit}")
    pvar_table <- read.table(text = sub("^#", "", readLines("${pvar}")), header = TRUE)[var_idx,]
    pvar <- pgenlibr::NewPvar("${pvar}")      # only for GetVariantId below
    psam <- read.table("${psam}", sep = "\\t", header = TRUE, comment.char = "", check.names = FALSE)
    colnames(psam)[colnames(psam) == "#IID"] <- "IID"
    colnames(psam)[colnames(psam) == "#FID"] <- "FID"

    model <- readRDS("${model}")
    model_frame_raw <- readRDS("${model_frame}")
    y.mm <- l.mm[["y.mm"]]; X.mm.null <- l.mm[["X.mm"]]; L <- l.mm[["L"]]
    ll.null <- l.null[["ll"]]; y.mm.pred <- l.null[["y.mm.pred"]]; e.p <- l.null[["e.p"]]; U1 <- l.null[["U1"]]

    samples <- names(y.mm)
    pgen_order <- match(samples, psam[["IID"]])
    psam <- psam[pgen_order,]
    model_frame_raw <- model_frame_raw[match(samples, rownames(model_frame_raw)),]
    stopifnot(all(names(y.mm) == rownames(X.mm.null)), all(names(y.mm) == rownames(L)), all(names(y.mm) == colnames(L)),
              all(names(y.mm) == psam[["IID"]]), all(names(y.mm) == names(y.mm.pred)), all(names(y.mm) == rownames(model_frame_raw)))
    stopifnot((sum(is.na(L)) + sum(is.na(X.mm.null)) + sum(is.na(y.mm)) + sum(is.na(psam[["IID"]])) +
               sum(is.na(y.mm.pred)) + sum(is.na(model_frame_raw))) == 0)

    # the real model as a genotype lookup: model.matrix with x == 0, 1, 2 on every row (reference: :124-127 per SNP)
    design_at <- function(g) {
        mf <- model_frame_raw; mf[["x"]] <- rep(g, nrow(mf))
        model.matrix(model[-2], data = model.frame(model[-2], data = mf))
    }
    M0 <- design_at(0); M1 <- design_at(1); M2 <- design_at(2)
    stopifnot(all(colnames(M0) == colnames(M1)), all(colnames(M0) == colnames(M2)), nrow(M0) == length(y.mm))
    # two more probes tell the engine how every SNP column continues between the hard-call values (needed when the .pgen carries
    # dosages and use_dosage is true, :26: x is then pgenlibr::Read's dosage and the columns are evaluated on it)
    M05 <- design_at(0.5); M15 <- design_at(1.5)
    # every visible GPU of the box works on this task's SNPs (FLEXLMM_DEVICES="0,1,..." narrows it)
    devices <- as.integer(strsplit(Sys.getenv("FLEXLMM_DEVICES", "0"), ",")[[1]])

    seeds <- as.integer(${seeds})
    res <- .Call("flexlmm_fit", normalizePath("${pgen}"), as.integer(var_idx), as.integer(pgen_order),
                 L, as.numeric(y.mm), X.mm.null, as.numeric(ll.null), as.integer(attributes(ll.null)[["df"]]),
                 M0, M1, M2, M05, M15, as.numeric(y.mm.pred), U1, as.numeric(e.p), seeds, devices, ${dosage})
    nvars <- length(var_idx)

    if (${do_permute}) {
        outname <- "${prefix}.perm.tsv.gz"
        out_con <- gzfile(outname, "w")
        stopifnot(all(res[["min_pval"]] >= 0 & res[["min_pval"]] <= 1))
        stopifnot(res[["nvars_tested"]] == nvars)
        for (j in seq_along(seeds)) writeLines(sprintf("%s\\t${meta.chr}\\t%s\\t%s", seeds[j], res[["min_pval"]][j], nvars), out_con)
        close(out_con)
    } else {
        outname <- "${prefix}.tsv.gwas.gz"
        out_con <- gzfile(outname, "w")
        writeLines("chr\\tpos\\tid\\tref\\talt\\tlrt_chisq\\tlrt_df\\tpval\\tbeta", out_con)
        for (i in seq_along(var_idx)) {
            beta <- paste(colnames(M0), res[["coef"]][, i], sep = "~", collapse = ",")
            writeLines(sprintf("%s\\t%s\\t%s\\t%s\\t%s\\t%s\\t%s\\t%s\\t%s", pvar_table[i, "CHROM"], pvar_table[i, "POS"],
                               pgenlibr::GetVariantId(pvar, var_idx[[i]]), pvar_table[i, "REF"], pvar_table[i, "ALT"],
                               res[["lrt_chisq"]][i], res[["lrt_df"]][i], res[["pval"]][i], beta), out_con)
        }
        close(out_con)
        gwas <- read.table(outname, header = TRUE, sep = "\\t")
        stopifnot(nrow(gwas) == nvars, ncol(gwas) == 9)
    }

    ver_r <- strsplit(as.character(R.version["version.string"]), " ")[[1]][3]
    writeLines(c("\\"${task.process}\\":", sprintf("    r-base: %s", ver_r),
                 sprintf("    pgenlibr: %s", utils::packageVersion("pgenlibr")), "    flexlmm_cuda: 0.1.0"), "versions.yml")
    """

    stub:
    def prefix      = task.ext.prefix ?: "${meta.id}"
    """
    touch ${prefix}.${perm_seed ? '.perm.tsv.gz' : '.tsv.gwas.gz'}
    touch versions.yml
    """
}
