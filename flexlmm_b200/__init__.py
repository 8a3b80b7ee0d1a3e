"""flexlmm_b200 — B200-native per-SNP likelihood-ratio-test engine (the FIT_MODEL hot path of birneylab/flexlmm).

The product is the C-ABI library (include/flexlmm_cuda.h, flexlmm_b200/csrc); this package is the host-side
mirror of the reference's process interface used by tests, bench.py and Python callers.
"""
from ._capi import FlxError, load, lib_path  # noqa: F401
from .fit_model import fit_null_model, FitModel, FitModelMulti, fit_model_task, fit_model_files, perm_indices, minp_threshold, write_gwas_tsv, write_perm_tsv  # noqa: F401

from .pfiles import read_pvar, read_psam, pgen_order  # noqa: F401
from .loco import LocoRunner, chromosomes_of_rank  # noqa: F401
from .postprocess import cat_perm, read_perm_table, summarise_by_chr, write_genome_wide_min_p, thresholds  # noqa: F401

__version__ = "0.1.0"
