#!/usr/bin/env Rscript
# make_r_fixture.R — THE PARITY PIN.  Run by anyone who has R (>= 3.6; the reference's processes pin r-base 4.2.1):
#
#     Rscript tests/golden/make_r_fixture.R            # from the repo root; base R only, no packages
#
# For every committed input bundle tests/golden/r_pin/<case>.in.txt it runs the reference's own statements — the literal
# loop of /root/reference/modules/local/r/fit_model.nf:96-143 (set.seed / sample / U1 %*% . / .lm.fit / stats:::logLik.lm /
# model.frame / model.matrix / forwardsolve / pchisq / min(..., na.rm = TRUE)), with model_frame_raw built the way
# GET_DESIGN_MATRIX builds it (get_design_matrix.nf:110-116) — and writes tests/golden/r_pin/<case>.out.txt.
# tests/test_r_pin.py then compares the CPU oracle AND the CUDA path with those files (the tests skip, with this reason,
# while the .out.txt files are absent).  Until a maintainer runs this script the oracle's parity stays "unpinned".
#
# Bundle format (written by tests/golden/make_golden.py, read back by tests/test_r_pin.py): records of two lines,
#   "@name nrow ncol" then nrow*ncol numbers, column-major, one line;   "$name text" for strings.
# Genotypes come from the bundle as hard-call counts (what pgenlibr::ReadHardcalls puts in buf, fit_model.nf:123): the
# .pgen decoding itself is pinned separately by the reference's own 77-byte in.pgen (tests/test_oracle.py).

args <- commandArgs(trailingOnly = FALSE)
self <- sub("^--file=", "", args[grep("^--file=", args)])
here <- if (length(self)) dirname(normalizePath(self[[1]])) else file.path(getwd(), "tests", "golden")
pin_dir <- file.path(here, "r_pin")

read_bundle <- function(path) {
    lines <- readLines(path)
    out <- list()
    i <- 1
    while (i <= length(lines)) {
        ln <- lines[[i]]
        if (startsWith(ln, "$")) {
            sp <- regexpr(" ", ln)
            out[[substr(ln, 2, sp - 1)]] <- substr(ln, sp + 1, nchar(ln))
            i <- i + 1
        } else if (startsWith(ln, "@")) {
            h <- strsplit(substr(ln, 2, nchar(ln)), " ")[[1]]
            nr <- as.integer(h[[2]]); nc <- as.integer(h[[3]])
            vals <- if (nr * nc > 0) as.numeric(strsplit(lines[[i + 1]], " ")[[1]]) else numeric(0)
            stopifnot(length(vals) == nr * nc)
            out[[h[[1]]]] <- matrix(vals, nrow = nr, ncol = nc)
            i <- i + 2
        } else {
            i <- i + 1
        }
    }
    out
}

num <- function(v) paste(ifelse(is.na(v), "NA", sprintf("%.17g", v)), collapse = " ")
put <- function(con, name, m) {
    m <- as.matrix(m)
    writeLines(sprintf("@%s %d %d", name, nrow(m), ncol(m)), con)
    writeLines(num(as.numeric(m)), con)
}

run_case <- function(path) {
    b <- read_bundle(path)
    n <- nrow(b[["L"]])
    samples <- sprintf("s%05d", seq_len(n))
    L <- b[["L"]]; dimnames(L) <- list(samples, samples)
    y.mm.orig <- setNames(drop(b[["y_mm"]]), samples)
    X.mm.null <- b[["X_null"]]; rownames(X.mm.null) <- samples
    ll.null.given <- drop(b[["ll_null"]]); df.null.given <- as.integer(drop(b[["df_null"]]))
    y.mm.pred <- setNames(drop(b[["y_pred"]]), samples)
    U1 <- b[["U1"]]
    e.p <- b[["e_p"]]                       # m x 1 matrix, as crossprod() leaves it (fit_null_model.nf:65)
    geno <- b[["geno"]]                     # n x M hard-call counts in model order = buf[pgen_order]
    seeds <- as.integer(drop(b[["seeds"]]))
    model <- as.formula(b[["formula"]])

    # ---- model_frame_raw as GET_DESIGN_MATRIX saves it (get_design_matrix.nf:110-116)
    df <- data.frame(IID = samples, stringsAsFactors = FALSE)
    for (nm in names(b)) {
        if (startsWith(nm, "numeric.")) df[[sub("^numeric\\.", "", nm)]] <- drop(b[[nm]])
        if (startsWith(nm, "factorcodes.")) {
            v <- sub("^factorcodes\\.", "", nm)
            lev <- strsplit(b[[paste0("factorlevels.", v)]], " ")[[1]]
            df[[v]] <- factor(lev[as.integer(drop(b[[nm]])) + 1], levels = lev)
        }
    }
    rownames(df) <- df[["IID"]]
    df[["x"]] <- -200
    model_frame_raw <- model.frame(model[-2], data = df)

    out <- file(sub("\\.in\\.txt$", ".out.txt", path), "w")
    writeLines(sprintf("$r_version %s", as.character(R.version[["version.string"]])), out)
    writeLines(sprintf("$case %s", b[["case"]]), out)

    one_task <- function(perm_seed) {
        do_permute <- !is.null(perm_seed)
        y.mm <- y.mm.orig
        ll.null <- ll.null.given
        attr(ll.null, "df") <- df.null.given
        idx <- NULL
        if (do_permute) {
            # fit_model.nf:96-104
            set.seed(perm_seed)
            idx <- sample.int(length(e.p))          # what sample(e.p) draws: sample(x) is x[sample.int(length(x))]
            set.seed(perm_seed)
            y.mm <- y.mm.pred + drop(U1 %*% sample(e.p))
            fit <- .lm.fit(x = X.mm.null, y = y.mm)
            ll.null <- stats:::logLik.lm(fit)
            min_pval <- 2
            nvars_tested <- 0
        }
        M <- ncol(geno)
        chisq <- numeric(M); lrt_dfs <- integer(M); pvals <- numeric(M); ranks <- integer(M)
        pivots <- NULL; coefs <- NULL; cn <- NULL
        for (i in seq_len(M)) {
            # fit_model.nf:123-139
            model_frame_raw[["x"]] <- geno[, i]
            model_frame <- model.frame(model[-2], data = model_frame_raw)
            X <- model.matrix(model[-2], data = model_frame)
            X.mm <- forwardsolve(L, X)
            colnames(X.mm) <- colnames(X)
            fit <- .lm.fit(x = X.mm, y = y.mm)
            ll <- stats:::logLik.lm(fit)
            lrt_df <- attributes(ll)[["df"]] - attributes(ll.null)[["df"]]
            lrt_chisq <- 2 * as.numeric(ll - ll.null)
            pval <- (
                if (lrt_df > 0) pchisq(lrt_chisq, df = lrt_df, lower.tail = FALSE) else NA
            )
            if (do_permute) {
                nvars_tested <- nvars_tested + 1
                min_pval <- min(pval, min_pval, na.rm = TRUE)
            }
            chisq[[i]] <- lrt_chisq; lrt_dfs[[i]] <- as.integer(lrt_df); pvals[[i]] <- pval; ranks[[i]] <- fit[["rank"]]
            if (is.null(pivots)) { pivots <- matrix(0L, ncol(X.mm), M); coefs <- matrix(0, ncol(X.mm), M); cn <- colnames(X.mm) }
            pivots[, i] <- fit[["pivot"]]; coefs[, i] <- coef(fit)
        }
        list(chisq = chisq, df = lrt_dfs, pval = pvals, rank = ranks, pivot = pivots, coef = coefs, colnames = cn, idx = idx,
             min_pval = if (do_permute) min_pval else NA, nvars = if (do_permute) nvars_tested else M,
             ll_null = as.numeric(ll.null), df_null = attributes(ll.null)[["df"]])
    }

    o <- one_task(NULL)
    writeLines(sprintf("$colnames %s", paste(o[["colnames"]], collapse = "\t")), out)
    put(out, "lrt_chisq", o[["chisq"]]); put(out, "lrt_df", o[["df"]]); put(out, "pval", o[["pval"]]); put(out, "rank", o[["rank"]])
    put(out, "pivot", o[["pivot"]]); put(out, "coef", o[["coef"]])
    # the null fit the way FIT_NULL_MODEL produces ll (fit_null_model.nf:33-34), to pin the given ll.null as well
    nfit <- lm.fit(x = X.mm.null, y = y.mm.orig)
    put(out, "ll_null_r", as.numeric(stats:::logLik.lm(nfit))); put(out, "df_null_r", nfit[["rank"]] + 1)
    for (sd in seeds) {
        p <- one_task(sd)
        put(out, sprintf("perm%d.sample_idx", sd), p[["idx"]])
        put(out, sprintf("perm%d.min_pval", sd), p[["min_pval"]])
        put(out, sprintf("perm%d.nvars", sd), p[["nvars"]])
        put(out, sprintf("perm%d.ll_null", sd), p[["ll_null"]])
        put(out, sprintf("perm%d.pval", sd), p[["pval"]])
    }
    close(out)
    cat("wrote", sub("\\.in\\.txt$", ".out.txt", path), "\n")
}

# ---- known answers of the un-vendored arithmetic (SURVEY.md Appendix B), from R itself
kat <- file(file.path(pin_dir, "kat.out.txt"), "w")
writeLines(sprintf("$r_version %s", as.character(R.version[["version.string"]])), kat)
for (sd in c(1L, 42L, 123L)) { set.seed(sd); put(kat, sprintf("sample10.seed%d", sd), sample(10)) }
for (m in c(9L, 298L, 4998L, 70001L)) { set.seed(7L); put(kat, sprintf("sample_int%d.seed7", m), sample.int(m)) }
set.seed(1L); put(kat, "runif3.seed1", runif(3))
q <- c(1e-300, 1e-6, 0.5, 3.841458820694124, 10, 50, 200, 800, 1400, 1500)
for (d in 1:6) put(kat, sprintf("pchisq_upper.df%d", d), pchisq(q, df = d, lower.tail = FALSE))
put(kat, "pchisq_q", q)
put(kat, "quantile7", quantile(c(0.3, 0.01, 0.2, 0.05, 0.5, 0.04, 0.09), 0.05))
close(kat)

for (f in sort(list.files(pin_dir, pattern = "\\.in\\.txt$", full.names = TRUE))) run_case(f)
