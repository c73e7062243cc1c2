"""VCF in and out for the path's callers: what the reference does through ADAM at
`sc.loadGenotypes` / `sc.loadVariants` (CLI/Jointer.scala:108-124, CLI/BiallelicGenotyper.scala:267,
CLI/TrioGenotyper.scala:230) and `saveAsVcf` (CLI/Jointer.scala:137-145, TrioCallerSuite.scala:239-251).

ADAM (adam-core 0.24.0, `VariantContextConverter`) and htsjdk are not under /root/reference; what
this module follows is pinned by the reference's own fixtures, committed (gzip-compressed) under tests/golden/vcf/:

* `trio.1_837214.vcf` -> TrioCaller -> `trio.1_837214.phased.vcf` compared byte for byte by the
  reference (TrioCallerSuite.scala:239-251): header order, FORMAT key order, Java float text, GT
  allele order and phasing, PL / AD / SB / FT text;
* `gvcf_multiallelic.g.vcf` -> loadGenotypes -> extractVariants (SquareOffReferenceModelSuite.scala:51-76):
  multi-allelic records split per alternate allele (the other alternates become OTHER_ALT), the
  `<NON_REF>` allele carried as nonReferenceLikelihoods, `END=` blocks as reference-model genotypes;
* `random.vcf` + HardFilterGenotypes' header lines (HardFilterGenotypesSuite.scala:368-375,
  util/HardFilterGenotypes.scala:186-240).
Anything those files do not exercise (OTHER_ALT in GT text, ploidy other than two in PL splitting,
float text beyond Float.toString's plain range) follows the VCF 4.2 text and is marked below.
"""
from __future__ import annotations

import gzip
import math
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .filters import FilterArgs
from .joint import VariantContext
from .records import ALT, NO_CALL, OTHER_ALT, REF, Genotype, Variant

NON_REF = "<NON_REF>"
LN10_DIV_10 = math.log(10.0) / 10.0

# org.bdgenomics.adam.converters.DefaultHeaderLines.allHeaderLines as htsjdk writes them: the FORMAT and
# INFO lines of every VCF the reference saves (tests/golden/vcf/trio.1_837214.phased.vcf, lines 15-41).
DEFAULT_FORMAT_LINES = [
    '##FORMAT=<ID=AD,Number=R,Type=Integer,Description="Allelic depths for the ref and alt alleles in the order listed">',
    '##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Approximate read depth (reads with MQ=255 or with bad mates are filtered)">',
    '##FORMAT=<ID=FS,Number=1,Type=Float,Description="Phred-scaled p-value using Fisher\'s exact test to detect strand bias">',
    '##FORMAT=<ID=FT,Number=.,Type=String,Description="Genotype-level filter">',
    '##FORMAT=<ID=GQ,Number=1,Type=Integer,Description="Genotype Quality">',
    '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
    '##FORMAT=<ID=MIN_DP,Number=1,Type=Integer,Description="Minimum DP observed within the gVCF block">',
    '##FORMAT=<ID=MQ,Number=1,Type=Float,Description="Root mean square (RMS) mapping quality">',
    '##FORMAT=<ID=MQ0,Number=1,Type=Float,Description="Total number of reads with mapping quality=0">',
    '##FORMAT=<ID=PL,Number=G,Type=Integer,Description="Normalized, Phred-scaled likelihoods for genotypes as defined in the VCF specification">',
    '##FORMAT=<ID=PQ,Number=1,Type=Float,Description="Read-backed phasing quality">',
    '##FORMAT=<ID=PS,Number=1,Type=Integer,Description="Phase set ID">',
    '##FORMAT=<ID=SB,Number=4,Type=Integer,Description="Per-sample component statistics which comprise the Fisher\'s Exact Test to detect strand bias.">',
]
DEFAULT_INFO_LINES = [
    '##INFO=<ID=1000G,Number=0,Type=Flag,Description="Membership in 1000 Genomes">',
    '##INFO=<ID=AA,Number=1,Type=String,Description="Ancestral allele">',
    '##INFO=<ID=AC,Number=A,Type=Integer,Description="Allele count">',
    '##INFO=<ID=AD,Number=R,Type=Integer,Description="Total read depths for each allele">',
    '##INFO=<ID=ADF,Number=R,Type=Integer,Description="Read depths for each allele on the forward strand">',
    '##INFO=<ID=ADR,Number=R,Type=Integer,Description="Read depths for each allele on the reverse strand">',
    '##INFO=<ID=AF,Number=A,Type=Float,Description="Allele frequency for each allele">',
    "##INFO=<ID=ANN,Number=.,Type=String,Description=\"Functional annotations: 'Allele | Annotation | Annotation_Impact | "
    "Gene_Name | Gene_ID | Feature_Type | Feature_ID | Transcript_BioType | Rank | HGVS.c | HGVS.p | cDNA.pos / cDNA.length | "
    "CDS.pos / CDS.length | AA.pos / AA.length | Distance | ERRORS / WARNINGS / INFO'\">",
    '##INFO=<ID=CIGAR,Number=A,Type=String,Description="Cigar string describing how to align alternate alleles to the reference allele">',
    '##INFO=<ID=DB,Number=0,Type=Flag,Description="Membership in dbSNP">',
    '##INFO=<ID=H2,Number=0,Type=Flag,Description="Membership in HapMap2">',
    '##INFO=<ID=H3,Number=0,Type=Flag,Description="Membership in HapMap3">',
    '##INFO=<ID=SOMATIC,Number=0,Type=Flag,Description="Somatic event">',
    '##INFO=<ID=VALIDATED,Number=0,Type=Flag,Description="Validated by follow-up experiment">',
]


def filter_header_lines(args: FilterArgs) -> List[str]:
    """The FILTER lines HardFilterGenotypes adds to the header (util/HardFilterGenotypes.scala:186-240):
    one per hard filter whose threshold is > 0, thresholds printed with Java's %f / %d."""
    spec = [
        ("HETSNPQD", "Quality by depth was below %f for a heterozygous SNP.", args.minHetSnpQualityByDepth),
        ("HOMSNPQD", "Quality by depth was below %f for a homozygous SNP.", args.minHomSnpQualityByDepth),
        ("HETINDELQD", "Quality by depth was below %f for a heterozygous INDEL.", args.minHetIndelQualityByDepth),
        ("HOMINDELQD", "Quality by depth was below %f for a homozygous INDEL.", args.minHomIndelQualityByDepth),
        ("SNPFS", "Phred Fisher scored strand bias was above %f for a SNP.", args.maxSnpPhredStrandBias),
        ("INDELFS", "Phred Fisher scored strand bias was above %f for a INDEL.", args.maxIndelPhredStrandBias),
        ("SNPMQ", "RMS mapping quality was below %f for a SNP.", args.minSnpRMSMappingQuality),
        ("INDELMQ", "RMS mapping quality was below %f for a INDEL.", args.minIndelRMSMappingQuality),
        ("SNPMINDP", "Read depth was below %d for a SNP.", args.minSnpDepth),
        ("INDELMINDP", "Read depth was below %d for a INDEL.", args.minIndelDepth),
        ("SNPMAXDP", "Read depth was above %d for a SNP.", args.maxSnpDepth),
        ("INDELMAXDP", "Read depth was above %d for a INDEL.", args.maxIndelDepth),
        ("HETSNPMINAF", "Allelic fraction was below %f for a het SNP.", args.minHetSnpAltAllelicFraction),
        ("HETSNPMAXAF", "Allelic fraction was above %f for a het SNP.", args.maxHetSnpAltAllelicFraction),
        ("HOMSNPMINAF", "Allelic fraction was below %f for a hom SNP.", args.minHomSnpAltAllelicFraction),
        ("HETINDELMINAF", "Allelic fraction was below %f for a het INDEL.", args.minHetIndelAltAllelicFraction),
        ("HETINDELMAXAF", "Allelic fraction was above %f for a het INDEL.", args.maxHetIndelAltAllelicFraction),
        ("HOMINDELMINAF", "Allelic fraction was below %f for a hom INDEL.", args.minHomIndelAltAllelicFraction),
    ]
    out = []
    for fid, text, value in spec:
        if value > 0:
            v = float(np.float32(value)) if "%f" in text else int(value)     # the arguments are Java floats / ints
            out.append('##FILTER=<ID=%s,Description="%s">' % (fid, text % v))
    return out


@dataclass
class VcfHeader:
    """Meta lines (without the line end) and sample names, in file order."""
    lines: List[str] = field(default_factory=list)
    samples: List[str] = field(default_factory=list)

    @property
    def contigs(self) -> List[str]:
        out = []
        for ln in self.lines:
            if ln.startswith("##contig=<"):
                for part in ln[len("##contig=<"):-1].split(","):
                    if part.startswith("ID="):
                        out.append(part[3:])
        return out

    def with_lines(self, extra: Iterable[str]) -> "VcfHeader":
        have = set(self.lines)
        return VcfHeader(self.lines + [ln for ln in extra if ln not in have], list(self.samples))

    def sorted_lines(self) -> List[str]:
        """htsjdk's order (VCFHeader keeps its meta data in a sorted set): the file format first, the
        other lines by their text, contig lines last in dictionary order -- the order of both golden files."""
        first = [ln for ln in self.lines if ln.startswith("##fileformat=")]
        contigs = [ln for ln in self.lines if ln.startswith("##contig=")]
        rest = sorted(set(ln for ln in self.lines if not ln.startswith("##fileformat=") and not ln.startswith("##contig=")))
        return (first or ["##fileformat=VCFv4.2"]) + rest + contigs


def default_header(samples: Sequence[str], contigs: Sequence[Tuple[str, int]] = (), filters: Optional[FilterArgs] = None
                   ) -> VcfHeader:
    """The header of a GenotypeRDD built by BiallelicGenotyper.call (DefaultHeaderLines.allHeaderLines,
    BG:127-134) after HardFilterGenotypes, for the given sequence dictionary."""
    lines = ["##fileformat=VCFv4.2"] + (filter_header_lines(filters) if filters else []) + DEFAULT_FORMAT_LINES + \
        DEFAULT_INFO_LINES + ["##contig=<ID=%s,length=%d>" % (n, ln) for n, ln in contigs]
    return VcfHeader(lines, list(samples))


# ---- text of numbers -------------------------------------------------------------------------------
def java_float_to_string(x: float) -> str:
    """java.lang.Float.toString: the shortest decimal that identifies the float, at least one digit
    after the point, plain between 1e-3 and 1e7 and computerised scientific notation outside."""
    f = np.float32(x)
    if np.isnan(f):
        return "NaN"
    if np.isinf(f):
        return "Infinity" if f > 0 else "-Infinity"
    if f == 0:
        return "-0.0" if np.signbit(f) else "0.0"
    a = abs(float(f))
    if 1e-3 <= a < 1e7:
        return np.format_float_positional(f, unique=True, trim="0")
    s = np.format_float_scientific(f, unique=True, trim="0", exp_digits=1)      # d.ddde[+-]x
    mant, exp = s.split("e")
    return "%sE%d" % (mant, int(exp))


def _pl_of(log_likelihoods: Sequence[float]) -> List[int]:
    # PhredUtils.logProbabilityToPhred on every likelihood (natural log -> phred, rounded)
    return [int(round(-ll / LN10_DIV_10)) for ll in log_likelihoods]


# ---- reading ---------------------------------------------------------------------------------------
@dataclass
class VcfRecord:
    contig: str
    pos: int                      # 1-based, as in the file
    id: str
    ref: str
    alts: List[str]
    qual: Optional[float]
    filters: str
    info: Dict[str, Optional[str]]
    format_keys: List[str]
    samples: List[Dict[str, str]]  # per sample: FORMAT key -> text (missing values left out)


def read_vcf(path) -> Tuple[VcfHeader, List[VcfRecord]]:
    """Plain or gzip / bgzip compressed (.gz) VCF text."""
    header, records = VcfHeader(), []
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rt") as fh:
        for raw in fh:
            line = raw.rstrip("\n").rstrip("\r")
            if not line:
                continue
            if line.startswith("##"):
                header.lines.append(line)
                continue
            if line.startswith("#"):
                header.samples = line.split("\t")[9:]
                continue
            f = line.split("\t")
            info: Dict[str, Optional[str]] = {}
            if len(f) > 7 and f[7] != ".":
                for kv in f[7].split(";"):
                    k, _, v = kv.partition("=")
                    info[k] = v if _ else None
            keys = f[8].split(":") if len(f) > 8 else []
            samples = []
            for text in f[9:]:
                vals = text.split(":")
                samples.append({k: v for k, v in zip(keys, vals) if v != "." and v != ""})
            records.append(VcfRecord(f[0], int(f[1]), f[2], f[3], [] if f[4] == "." else f[4].split(","),
                                     None if f[5] == "." else float(f[5]), f[6], info, keys, samples))
    return header, records


def _ints(text: Optional[str]) -> Optional[List[int]]:
    if text is None:
        return None
    return [int(t) if t != "." else -1 for t in text.split(",")]


def _gl_index(j: int, k: int) -> int:
    """Position of the unordered diploid genotype j/k (j <= k) in PL / GL order (VCF 4.2, 1.4.2)."""
    return k * (k + 1) // 2 + j


def _genotype_of(rec: VcfRecord, sample: str, fields: Dict[str, str], alt_index: Optional[int],
                 non_ref_index: Optional[int], variant: Variant) -> Genotype:
    """One sample's call at one (ref, alt) pair of a record: VariantContextConverter's genotype
    extraction.  alt_index: index of the alternate allele in the record's allele list (None for a
    reference-model record).  Alleles of the GT that are neither REF nor this alternate become
    OTHER_ALT (the split of a multi-allelic record)."""
    gt = fields.get("GT", "")
    phased = "|" in gt
    alleles = []
    for tok in gt.replace("|", "/").split("/") if gt else []:
        if tok == ".":
            alleles.append(NO_CALL)
        else:
            a = int(tok)
            alleles.append(REF if a == 0 else ALT if a == alt_index else OTHER_ALT)
    ad, pl, sb = _ints(fields.get("AD")), _ints(fields.get("PL")), _ints(fields.get("SB"))

    def punch(allele: Optional[int]) -> List[float]:
        # the (ref, allele) corner of a diploid PL vector, as log likelihoods
        if pl is None or allele is None:
            return []
        idx = [_gl_index(0, 0), _gl_index(0, allele), _gl_index(allele, allele)]
        if max(idx) >= len(pl):
            return []
        return [-p * LN10_DIV_10 + 0.0 for p in (pl[i] for i in idx)]

    if alt_index is None:                    # reference model: the likelihoods are those against <NON_REF>
        gl, nonref = [], punch(non_ref_index) if non_ref_index is not None else punch(1)
    else:
        gl, nonref = punch(alt_index), punch(non_ref_index)
    ft = fields.get("FT")
    g = Genotype(
        variant=variant, contigName=variant.contigName, start=variant.start, end=variant.end, sampleId=sample,
        alleles=alleles, genotypeQuality=int(fields["GQ"]) if "GQ" in fields else None,
        genotypeLikelihoods=gl, nonReferenceLikelihoods=nonref, strandBiasComponents=sb or [],
        readDepth=int(fields["DP"]) if "DP" in fields else None,
        referenceReadDepth=ad[0] if ad else None,
        alternateReadDepth=ad[alt_index] if ad and alt_index is not None and alt_index < len(ad) else None,
        rmsMapQ=float(np.float32(fields["MQ"])) if "MQ" in fields else float("nan"),
        fisherStrandBiasPValue=float(np.float32(fields["FS"])) if "FS" in fields else float("nan"),
        phased=phased)
    if ft is not None:
        g.filtersApplied = True
        g.filtersPassed = ft == "PASS"
        g.filtersFailed = [] if ft == "PASS" else ft.split(";")
    if "MIN_DP" in fields:
        g.minReadDepth = int(fields["MIN_DP"])
    return g


def _split(rec: VcfRecord):
    """(variant, alt index, <NON_REF> index) per biallelic record ADAM makes of a VCF line: one per
    alternate allele other than <NON_REF>; a line whose only alternate is <NON_REF> (or that has none)
    is one reference-model record, covering [POS - 1, END) when INFO carries END."""
    start = rec.pos - 1
    non_ref = rec.alts.index(NON_REF) + 1 if NON_REF in rec.alts else None
    real = [(i + 1, a) for i, a in enumerate(rec.alts) if a != NON_REF]
    if not real:
        end = int(rec.info["END"]) if rec.info.get("END") else start + len(rec.ref)
        yield Variant(rec.contig, start, end, rec.ref, None), None, non_ref
        return
    for idx, alt in real:
        yield Variant(rec.contig, start, start + len(rec.ref), rec.ref, alt), idx, non_ref


def load_genotypes(path) -> Tuple[List[Genotype], VcfHeader]:
    """sc.loadGenotypes for a VCF / gVCF file: every (record, alternate allele, sample) as a Genotype."""
    header, records = read_vcf(path)
    out = []
    for rec in records:
        for variant, alt_index, non_ref in _split(rec):
            for sample, fields in zip(header.samples, rec.samples):
                out.append(_genotype_of(rec, sample, fields, alt_index, non_ref, variant))
    return out, header


def load_variants(path) -> List[Variant]:
    """sc.loadVariants: the sites of a VCF, multi-allelic lines split, reference-model lines dropped."""
    _, records = read_vcf(path)
    return [v for rec in records for v, alt_index, _ in _split(rec) if alt_index is not None]


def merge_discovered(paths: Sequence[str]) -> List[Variant]:
    """CLI/MergeDiscovered.scala:52-62: the variants of several files with duplicates (same contig, start,
    end, alleles) dropped -- the union of per-sample discovery every sample is then called against."""
    seen, out = set(), []
    for path in paths:
        for v in load_variants(path):
            if v.as_tuple() not in seen:
                seen.add(v.as_tuple())
                out.append(v)
    return out


def reference_model_blocks(genotypes: Sequence[Genotype], sample: str) -> Tuple[np.ndarray, np.ndarray, List[Genotype]]:
    """The reference-model genotypes of one sample as (start, length, rows) sorted by start: what
    avo_square_off takes as ref_model_start / ref_model_len for a banded gVCF."""
    rows = sorted((g for g in genotypes if g.sampleId == sample and g.variant.alternateAllele is None),
                  key=lambda g: (g.start, g.end))
    return (np.asarray([g.start for g in rows], np.int32), np.asarray([g.end - g.start for g in rows], np.int32), rows)


def _row_columns(rows: Sequence[Genotype], n_states: int) -> Dict[str, np.ndarray]:
    n = len(rows)
    cols = {"emitted": np.ones(max(n, 1), np.uint8), "n_alt": np.zeros(max(n, 1), np.int32),
            "n_ref": np.zeros(max(n, 1), np.int32), "n_other": np.zeros(max(n, 1), np.int32),
            "gl": np.zeros((max(n, 1), n_states)), "nonref_gl": np.zeros((max(n, 1), n_states))}
    for k, g in enumerate(rows):
        cols["n_alt"][k], cols["n_ref"][k] = g.alleles.count(ALT), g.alleles.count(REF)
        cols["n_other"][k] = g.alleles.count(OTHER_ALT)
        for name, src in (("gl", g.genotypeLikelihoods), ("nonref_gl", g.nonReferenceLikelihoods)):
            cols[name][k, :min(len(src), n_states)] = src[:n_states]
    return cols


def sample_tables(genotypes: Sequence[Genotype], sample: str, n_states: int = 3):
    """One sample's genotypes of ONE contig, as read from a gVCF, in the form avo_square_off takes
    (squareoff.square_off_packed): (variant table, variant rows in table order, their columns,
    reference-model starts, lengths, rows, their columns)."""
    from .batch import sites_from_variants
    mine = [g for g in genotypes if g.sampleId == sample]
    if len({g.contigName for g in mine}) > 1:
        raise ValueError("one contig per call")
    var_rows = [g for g in mine if g.variant.alternateAllele is not None]
    table = sites_from_variants([g.variant for g in var_rows])
    table.contig = None
    var_rows = [var_rows[i] for i in table.order]
    start, length, ref_rows = reference_model_blocks(mine, sample)
    return table, var_rows, _row_columns(var_rows, n_states), start, length, ref_rows, _row_columns(ref_rows, n_states)


# ---- writing ---------------------------------------------------------------------------------------
_ALLELE_INDEX = {REF: "0", ALT: "1", OTHER_ALT: "2", NO_CALL: "."}    # OTHER_ALT -> 2: not exercised by the fixtures


def _sample_fields(g: Optional[Genotype]) -> Dict[str, str]:
    if g is None:
        return {"GT": "./."}
    d: Dict[str, str] = {}
    d["GT"] = ("|" if g.phased else "/").join(_ALLELE_INDEX[a] for a in g.alleles) if g.alleles else "./."
    if g.referenceReadDepth is not None and g.alternateReadDepth is not None:
        d["AD"] = "%d,%d" % (g.referenceReadDepth, g.alternateReadDepth)
    if g.readDepth is not None:
        d["DP"] = str(int(g.readDepth))
    if g.minReadDepth is not None:
        d["MIN_DP"] = str(int(g.minReadDepth))
    if g.fisherStrandBiasPValue is not None and not math.isnan(g.fisherStrandBiasPValue):
        d["FS"] = java_float_to_string(g.fisherStrandBiasPValue)
    if g.rmsMapQ is not None and not math.isnan(g.rmsMapQ):
        d["MQ"] = java_float_to_string(g.rmsMapQ)
    if g.filtersApplied:
        d["FT"] = "PASS" if g.filtersPassed else ";".join(g.filtersFailed or [])
    if g.genotypeQuality is not None:
        d["GQ"] = str(int(g.genotypeQuality))
    if g.genotypeLikelihoods:
        d["PL"] = ",".join(str(p) for p in _pl_of(g.genotypeLikelihoods))
    if g.strandBiasComponents:
        d["SB"] = ",".join(str(int(x)) for x in g.strandBiasComponents)
    return d


def format_record(vc: VariantContext, samples: Sequence[str]) -> str:
    """One VCF line for a site and its genotypes (htsjdk's VCFEncoder: GT first, the other FORMAT keys
    in alphabetical order, missing values as '.', trailing missing values dropped)."""
    v = vc.variant
    by_sample = {}
    for g in vc.genotypes:
        by_sample.setdefault(g.sampleId, g)
    per = [_sample_fields(by_sample.get(s)) for s in samples]
    keys = sorted({k for d in per for k in d if k != "GT"})
    keys = (["GT"] if any("GT" in d for d in per) else []) + keys
    cols = []
    for d in per:
        vals = [d.get(k, ".") for k in keys]
        while len(vals) > 1 and vals[-1] == ".":
            vals.pop()
        cols.append(":".join(vals))
    qual = "."
    if vc.quality is not None and not math.isnan(vc.quality):
        qual = "%.2f" % vc.quality            # htsjdk VCFEncoder.formatQualValue: two decimals, ".00" dropped
        if qual.endswith(".00"):
            qual = qual[:-3]
    alt = v.alternateAllele if v.alternateAllele is not None else NON_REF
    head = [v.contigName, str(v.start + 1), ".", v.referenceAllele, alt, qual, ".", "."]
    return "\t".join(head + ([":".join(keys)] + cols if samples else []))


def group_by_variant(genotypes: Sequence[Genotype]) -> List[VariantContext]:
    """GenotypeRDD.toVariantContexts: genotypes grouped by their variant (order of first appearance)."""
    groups: Dict[tuple, VariantContext] = {}
    for g in genotypes:
        key = g.variant.as_tuple()
        if key not in groups:
            groups[key] = VariantContext(g.variant, float("nan"), [], float("nan"))
        groups[key].genotypes.append(g)
    return list(groups.values())


def write_vcf(path, header: VcfHeader, contexts: Sequence[VariantContext]) -> None:
    """saveAsVcf: the header in htsjdk's order, then one line per site sorted by (contig, start)."""
    rank = {c: i for i, c in enumerate(header.contigs)}
    ordered = sorted(contexts, key=lambda vc: (rank.get(vc.variant.contigName, len(rank)), vc.variant.contigName,
                                               vc.variant.start, vc.variant.end))
    with open(path, "wt") as fh:
        for ln in header.sorted_lines():
            fh.write(ln + "\n")
        fh.write("\t".join(["#CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO"] +
                           (["FORMAT"] + list(header.samples) if header.samples else [])) + "\n")
        for vc in ordered:
            fh.write(format_record(vc, header.samples) + "\n")
