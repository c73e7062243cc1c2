"""Host-side mirror of the two per-genotype passes the CLI applies to call()'s output
(CLI/BiallelicGenotyper.scala:279-282) over the C ABI (avo_filter_genotypes):

    RewriteHets(genotypes, args)                                  -> CORE/util/RewriteHets.scala:60-87
    HardFilterGenotypes(genotypes, args, filterRefGenotypes, emitAllGenotypes)
                                                                  -> CORE/util/HardFilterGenotypes.scala:176-247

Both run in one kernel over the result columns."""
from __future__ import annotations

import ctypes as C
import dataclasses
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib as L
from .batch import _device_empty, _ptr
from .records import ALT, OTHER_ALT, REF, Genotype


@dataclass
class FilterArgs:
    """RewriteHetsArgs + HardFilterGenotypesArgs with the CLI's defaults
    (CLI/BiallelicGenotyper.scala:113-201).  A hard filter is applied iff its value is > 0."""
    minQuality: int = 30
    minHetSnpQualityByDepth: float = 2.0
    minHomSnpQualityByDepth: float = 1.0
    minHetIndelQualityByDepth: float = 2.0
    minHomIndelQualityByDepth: float = 1.0
    maxSnpPhredStrandBias: float = -1.0
    maxIndelPhredStrandBias: float = -1.0
    minSnpRMSMappingQuality: float = 30.0
    minIndelRMSMappingQuality: float = -1.0
    minSnpDepth: int = 10
    maxSnpDepth: int = 200
    minIndelDepth: int = 10
    maxIndelDepth: int = 200
    minHetSnpAltAllelicFraction: float = 0.333
    maxHetSnpAltAllelicFraction: float = 0.666
    minHomSnpAltAllelicFraction: float = 0.666
    minHetIndelAltAllelicFraction: float = 0.333
    maxHetIndelAltAllelicFraction: float = 0.666
    minHomIndelAltAllelicFraction: float = 0.666
    disableHetSnpRewriting: bool = False
    disableHetIndelRewriting: bool = False

    def as_struct(self, filter_ref_genotypes=True, emit_all_genotypes=False, rewrite=True):
        p = L.FilterParamsStruct()
        p.min_quality = self.minQuality
        p.filter_ref_genotypes, p.emit_all_genotypes = int(filter_ref_genotypes), int(emit_all_genotypes)
        p.disable_het_snp_rewriting = int(self.disableHetSnpRewriting or not rewrite)
        p.disable_het_indel_rewriting = int(self.disableHetIndelRewriting or not rewrite)
        p.min_snp_depth, p.max_snp_depth = self.minSnpDepth, self.maxSnpDepth
        p.min_indel_depth, p.max_indel_depth = self.minIndelDepth, self.maxIndelDepth
        p.min_het_snp_quality_by_depth, p.min_hom_snp_quality_by_depth = self.minHetSnpQualityByDepth, self.minHomSnpQualityByDepth
        p.min_het_indel_quality_by_depth, p.min_hom_indel_quality_by_depth = self.minHetIndelQualityByDepth, self.minHomIndelQualityByDepth
        p.max_snp_phred_strand_bias, p.max_indel_phred_strand_bias = self.maxSnpPhredStrandBias, self.maxIndelPhredStrandBias
        p.min_snp_rms_mapping_quality, p.min_indel_rms_mapping_quality = self.minSnpRMSMappingQuality, self.minIndelRMSMappingQuality
        p.min_het_snp_alt_allelic_fraction = self.minHetSnpAltAllelicFraction
        p.max_het_snp_alt_allelic_fraction = self.maxHetSnpAltAllelicFraction
        p.min_hom_snp_alt_allelic_fraction = self.minHomSnpAltAllelicFraction
        p.min_het_indel_alt_allelic_fraction = self.minHetIndelAltAllelicFraction
        p.max_het_indel_alt_allelic_fraction = self.maxHetIndelAltAllelicFraction
        p.min_hom_indel_alt_allelic_fraction = self.minHomIndelAltAllelicFraction
        return p


NEEDED = ("emitted", "n_alt", "n_ref", "n_other", "gq", "total_cov", "allele_cov", "fisher", "rms_mapq")


def filter_columns(ctx, cols: Dict[str, object], ref_len, alt_len, args: FilterArgs, filterRefGenotypes=True,
                   emitAllGenotypes=False, rewrite=True) -> Dict[str, object]:
    """avo_filter_genotypes over result columns (numpy or torch CUDA): `cols` needs the NEEDED
    columns; `alt_len` < 0 (or None) marks rows without alternate allele.  Returns the
    FILTER_OUT_COLUMNS."""
    n = len(ref_len)
    dev = not isinstance(cols["gq"], np.ndarray)
    rows = L.SiteResultsStruct()
    rows.n_sites, rows.n_states = n, 0
    rows.mem = L.MEM_DEVICE if dev else L.MEM_HOST
    keep = []
    for k in NEEDED:
        a = cols[k]
        if not dev:
            dt = np.uint8 if k == "emitted" else np.float32 if k in ("fisher", "rms_mapq") else np.int32
            a = np.ascontiguousarray(a, dt)
            keep.append(a)
        setattr(rows, k, _ptr(a))
    if not dev:
        ref_len = np.ascontiguousarray(ref_len, np.int32)
        alt_len = None if alt_len is None else np.ascontiguousarray(alt_len, np.int32)
    out = L.FilterOutStruct()
    out.n, out.mem = n, rows.mem
    res = {}
    for name, dt in L.FILTER_OUT_COLUMNS:
        a = _device_empty(n, np.dtype(dt).type, "cuda:%d" % ctx.device) if dev else np.zeros(max(n, 1), dt)
        if dev and dt == "uint32":
            a = _device_empty(n, np.int32, "cuda:%d" % ctx.device)
        res[name] = a
        setattr(out, name, _ptr(a))
    p = args.as_struct(filterRefGenotypes, emitAllGenotypes, rewrite)
    ctx._check(ctx.lib.avo_filter_genotypes(ctx.h, C.byref(rows), _ptr(ref_len), _ptr(alt_len), C.byref(p),
                                            C.byref(out)))
    return {k: v[:n] for k, v in res.items()}


def _apply(genotypes: Sequence[Genotype], args: FilterArgs, filterRef, emitAll, rewrite, hard, ctx) -> List[Genotype]:
    from .genotyping import default_context
    ctx = ctx or default_context()
    n = len(genotypes)
    if n == 0:
        return []
    cnt = lambda g, a: sum(1 for x in g.alleles if x == a)
    cols = dict(emitted=np.ones(n, np.uint8),
                n_alt=np.array([cnt(g, ALT) for g in genotypes], np.int32),
                n_ref=np.array([cnt(g, REF) for g in genotypes], np.int32),
                n_other=np.array([cnt(g, OTHER_ALT) for g in genotypes], np.int32),
                gq=np.array([0 if g.genotypeQuality is None else g.genotypeQuality for g in genotypes], np.int32),
                total_cov=np.array([g.readDepth for g in genotypes], np.int32),
                allele_cov=np.array([g.alternateReadDepth for g in genotypes], np.int32),
                fisher=np.array([g.fisherStrandBiasPValue for g in genotypes], np.float32),
                rms_mapq=np.array([g.rmsMapQ for g in genotypes], np.float32))
    ref_len = np.array([len(g.variant.referenceAllele) for g in genotypes], np.int32)
    alt_len = np.array([-1 if g.variant.alternateAllele is None else len(g.variant.alternateAllele)
                        for g in genotypes], np.int32)
    a = args if hard else dataclasses.replace(
        args, minQuality=-(1 << 30), minHetSnpQualityByDepth=-1, minHomSnpQualityByDepth=-1,
        minHetIndelQualityByDepth=-1, minHomIndelQualityByDepth=-1, maxSnpPhredStrandBias=-1,
        maxIndelPhredStrandBias=-1, minSnpRMSMappingQuality=-1, minIndelRMSMappingQuality=-1, minSnpDepth=-1,
        maxSnpDepth=-1, minIndelDepth=-1, maxIndelDepth=-1, minHetSnpAltAllelicFraction=-1,
        minHomSnpAltAllelicFraction=-1, minHetIndelAltAllelicFraction=-1, minHomIndelAltAllelicFraction=-1)
    if not hard:
        # RewriteHets alone: the rewrite thresholds are the max-AF values, which must stay; they are
        # also hard filters, so emission of everything and no filter bookkeeping is requested instead
        emitAll, filterRef = True, False
    out = filter_columns(ctx, cols, ref_len, alt_len, a, filterRef, emitAll, rewrite)
    res = []
    for i, g in enumerate(genotypes):
        if not out["emit"][i]:
            continue
        new = g
        if out["rewritten"][i]:                                           # RewriteHets.scala:134-143
            new = dataclasses.replace(new, alleles=[ALT] * len(g.alleles), genotypeQuality=None)
        if hard and g.variant.alternateAllele is not None:               # HardFilterGenotypes.scala:590-617
            failed = [L.FILTER_NAMES[b] for b in range(18) if (int(out["filters_failed"][i]) >> b) & 1]
            applied = bool(out["filters_applied"][i])
            new = dataclasses.replace(new, filtersApplied=True if applied else None,
                                      filtersPassed=(not failed) if applied else None,
                                      filtersFailed=failed if failed else None)
        res.append(new)
    return res


def RewriteHets(genotypes: Sequence[Genotype], args: FilterArgs, ctx=None) -> List[Genotype]:
    """RewriteHets.apply (RewriteHets.scala:65-87)."""
    return _apply(genotypes, args, False, True, True, False, ctx)


def HardFilterGenotypes(genotypes: Sequence[Genotype], args: FilterArgs, filterRefGenotypes: bool = True,
                        emitAllGenotypes: bool = False, ctx=None) -> List[Genotype]:
    """HardFilterGenotypes.apply (HardFilterGenotypes.scala:176-247); no rewriting."""
    return _apply(genotypes, args, filterRefGenotypes, emitAllGenotypes, False, True, ctx)


def rewrite_and_filter(genotypes: Sequence[Genotype], args: FilterArgs, filterRefGenotypes: bool = True,
                       emitAllGenotypes: bool = False, ctx=None) -> List[Genotype]:
    """HardFilterGenotypes(RewriteHets(genotypes, args), args, ...) as the CLI composes them
    (CLI/BiallelicGenotyper.scala:279-282), one kernel launch."""
    return _apply(genotypes, args, filterRefGenotypes, emitAllGenotypes, True, True, ctx)
