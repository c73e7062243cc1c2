"""Parquet output of the called genotypes: the `Genotype` records that
`BiallelicGenotyper.observationToGenotype` builds (CORE/genotyping/BiallelicGenotyper.scala:731-747) and the
CLI saves with `saveAsParquet` (CLI/BiallelicGenotyper.scala:285), written COLUMN BY COLUMN from the SoA result
columns of the C ABI -- no per-record objects, no arithmetic on the host (pyarrow builds the list columns from
offsets).  Field names and nesting follow bdg-formats' Genotype / Variant / VariantCallingAnnotations Avro
schema (0.11.x, the version ADAM 0.24 pins), restricted to the fields this path sets; ADAM's own
`loadParquetGenotypes` reads the file by column name.  Not pinned by a reference fixture (the reference's tests
write no Parquet); tests/test_parquet_cpu.py pins the round trip and the record-level equivalence."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

from .batch import NIB, SiteTable

ALLELE_NAMES = np.array(["ALT", "REF", "OTHER_ALT"])     # GenotypeAllele symbols, in the order BG:691-697 emits them


def _allele_strings(sites: SiteTable, rows: np.ndarray):
    """(referenceAllele, alternateAllele) string arrays of the given rows, from the 4-bit packed table."""
    t = sites.to_host()
    lut = np.frombuffer(NIB.encode(), np.uint8)
    a4 = np.asarray(t.alleles4, np.uint8)
    nib = np.empty(a4.size * 2, np.uint8)
    nib[0::2], nib[1::2] = a4 >> 4, a4 & 15
    chars = lut[nib]
    off = np.asarray(t.allele_off, np.int64)[rows]
    rl = np.asarray(t.ref_len, np.int64)[rows]
    al = np.asarray(t.alt_len, np.int64)[rows]
    ref = [chars[o:o + r].tobytes().decode() for o, r in zip(off, rl)]
    alt = [chars[o + r:o + r + a].tobytes().decode() for o, r, a in zip(off, rl, al)]
    return ref, alt


def genotype_table(sites: SiteTable, cols: Dict[str, np.ndarray], contig_name: str, sample_id: Optional[str]):
    """A pyarrow Table of the emitted genotypes (rows with cols['emitted'] != 0), in table order."""
    import pyarrow as pa
    cols = {k: np.asarray(v) for k, v in cols.items()}
    rows = np.nonzero(cols["emitted"][:sites.n])[0]
    n = len(rows)
    t = sites.to_host()
    start = np.asarray(t.start, np.int64)[rows]
    end = start + np.asarray(t.ref_len, np.int64)[rows]
    ref, alt = _allele_strings(t, rows)
    contig = pa.array([contig_name] * n, pa.string())
    variant = pa.StructArray.from_arrays(
        [contig, pa.array(start), pa.array(end), pa.array(ref, pa.string()), pa.array(alt, pa.string())],
        ["contigName", "start", "end", "referenceAllele", "alternateAllele"])
    annotations = pa.StructArray.from_arrays(
        [pa.array(cols["fisher"][rows].astype(np.float32)), pa.array(cols["rms_mapq"][rows].astype(np.float32))],
        ["fisherStrandBiasPValue", "rmsMapQ"])
    # alleles: ALT x nAlt ++ REF x nRef ++ OTHER_ALT x nOther (BG:691-697), as one flat array + offsets
    counts = np.stack([cols["n_alt"][rows], cols["n_ref"][rows], cols["n_other"][rows]], 1).astype(np.int64)
    per_row = counts.sum(1)
    a_off = np.zeros(n + 1, np.int32)
    np.cumsum(per_row, out=a_off[1:])
    flat = np.repeat(np.tile(np.arange(3), n), counts.ravel())
    alleles = pa.ListArray.from_arrays(pa.array(a_off), pa.array(ALLELE_NAMES[flat].tolist(), pa.string()))
    # likelihood arrays hold copyNumber + 1 entries (BG:713-729)
    k = (cols["copy_number"][rows] + 1).astype(np.int64)
    l_off = np.zeros(n + 1, np.int32)
    np.cumsum(k, out=l_off[1:])
    ns = cols["gl"].shape[1] if cols["gl"].ndim == 2 else 1
    mask = np.arange(ns)[None, :] < k[:, None]
    gl = pa.ListArray.from_arrays(pa.array(l_off), pa.array(cols["gl"].reshape(-1, ns)[rows][mask]))
    nrl = pa.ListArray.from_arrays(pa.array(l_off), pa.array(cols["nonref_gl"].reshape(-1, ns)[rows][mask]))
    sb = pa.FixedSizeListArray.from_arrays(pa.array(cols["sb"].reshape(-1, 4)[rows].ravel().astype(np.int32)), 4)
    sb = sb.cast(pa.list_(pa.int32()))
    return pa.table({
        "variant": variant, "contigName": contig, "start": pa.array(start), "end": pa.array(end),
        "variantCallingAnnotations": annotations,
        "sampleId": pa.array([sample_id] * n, pa.string()),
        "alleles": alleles,
        "genotypeQuality": pa.array(cols["gq"][rows].astype(np.int32)),
        "readDepth": pa.array(cols["total_cov"][rows].astype(np.int32)),
        "referenceReadDepth": pa.array(cols["other_cov"][rows].astype(np.int32)),
        "alternateReadDepth": pa.array(cols["allele_cov"][rows].astype(np.int32)),
        "strandBiasComponents": sb,
        "genotypeLikelihoods": gl,
        "nonReferenceLikelihoods": nrl,
    })


def write_genotypes_parquet(path, sites: SiteTable, cols, contig_name: str, sample_id: Optional[str], **kw):
    """saveAsParquet of the genotypes of one sample / contig (CLI/BiallelicGenotyper.scala:285)."""
    import pyarrow.parquet as pq
    pq.write_table(genotype_table(sites, cols, contig_name, sample_id), path, **kw)


def read_genotypes_parquet(path):
    """-> list of records.Genotype (the mirror's record type), for round-trip checks and small outputs."""
    import pyarrow.parquet as pq
    from .records import Genotype, Variant
    out = []
    for r in pq.read_table(path).to_pylist():
        v = r["variant"]
        var = Variant(v["contigName"], v["start"], v["end"], v["referenceAllele"], v["alternateAllele"])
        out.append(Genotype(variant=var, contigName=r["contigName"], start=r["start"], end=r["end"], sampleId=r["sampleId"],
                            alleles=list(r["alleles"]), genotypeQuality=r["genotypeQuality"],
                            genotypeLikelihoods=list(r["genotypeLikelihoods"]),
                            nonReferenceLikelihoods=list(r["nonReferenceLikelihoods"]),
                            strandBiasComponents=list(r["strandBiasComponents"]), readDepth=r["readDepth"],
                            referenceReadDepth=r["referenceReadDepth"], alternateReadDepth=r["alternateReadDepth"],
                            rmsMapQ=r["variantCallingAnnotations"]["rmsMapQ"],
                            fisherStrandBiasPValue=r["variantCallingAnnotations"]["fisherStrandBiasPValue"]))
    return out
