"""VCF in / out (avocado_b200/vcf.py) against the reference's own VCF fixtures (tests/golden/vcf/*.vcf.gz, made
by tests/golden/make_golden.py): the byte-for-byte expectation of TrioCallerSuite.scala:239-251, the
loadGenotypes -> extractVariants expectations of SquareOffReferenceModelSuite.scala:51-76 and the header
lines of HardFilterGenotypesSuite.scala:368-375.  The trio decision itself comes from the oracle here
(the GPU path is tests/test_gpu_trio.py)."""
import dataclasses
import os

import numpy as np
import pytest

from avocado_b200 import vcf as V
from avocado_b200.filters import FilterArgs
from avocado_b200.records import ALT, NO_CALL, OTHER_ALT, REF

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vcf")
CODE = {REF: 0, ALT: 1, OTHER_ALT: 2, NO_CALL: 3}


def _bytes(path):
    import gzip
    with (gzip.open if str(path).endswith(".gz") else open)(path, "rb") as fh:
        return fh.read()


def test_round_trip_reproduces_the_input_file(tmp_path):
    gts, header = V.load_genotypes(os.path.join(GOLD, "trio.1_837214.vcf.gz"))
    assert header.samples == ["NA12891", "NA12892", "NA12878"] and len(gts) == 3
    g = gts[1]
    assert (g.contigName, g.start, g.end, g.variant.referenceAllele, g.variant.alternateAllele) == ("chr1", 837213, 837214, "G", "C")
    assert g.alleles == [ALT, REF] and not g.phased and g.genotypeQuality == 99
    assert (g.referenceReadDepth, g.alternateReadDepth, g.readDepth) == (26, 27, 53)
    assert g.strandBiasComponents == [13, 13, 13, 14] and g.filtersApplied and g.filtersPassed
    assert np.float32(g.fisherStrandBiasPValue) == np.float32(6.689699) and np.float32(g.rmsMapQ) == np.float32(59.564617)
    assert np.allclose(g.genotypeLikelihoods, [0.0, -25.6 * np.log(10.0), 0.0])
    assert gts[0].filtersFailed == ["HETSNPMINAF"] and not gts[0].filtersPassed
    out = tmp_path / "same.vcf"
    V.write_vcf(out, header, V.group_by_variant(gts))
    assert _bytes(out) == _bytes(os.path.join(GOLD, "trio.1_837214.vcf.gz")) + b"\n"      # the input lacks the final newline


def test_trio_call_end_to_end_matches_the_reference_file(oracle, tmp_path):
    """TrioCallerSuite.scala:239-251: loadGenotypes -> TrioCaller -> saveAsVcf == trio.1_837214.phased.vcf."""
    gts, header = V.load_genotypes(os.path.join(GOLD, "trio.1_837214.vcf.gz"))
    by = {g.sampleId: g for g in gts}
    p1, p2, ch = by["NA12891"], by["NA12892"], by["NA12878"]
    r = oracle.trio_process(*[[CODE[a] for a in g.alleles] for g in (p1, p2, ch)])
    assert r["kept"] and r["childAction"] == 4           # phased [REF, ALT]: the ALT comes from the second parent
    names = {v: k for k, v in CODE.items()}
    child = dataclasses.replace(ch, phased=True, alleles=[names[a] for a in r["childAlleles"]])
    out = tmp_path / "trio.vcf"
    V.write_vcf(out, header, V.group_by_variant([p1, p2, child]))
    assert _bytes(out) == _bytes(os.path.join(GOLD, "trio.1_837214.phased.vcf.gz"))


def test_default_header_lines_are_those_of_the_reference_files():
    ref_lines = set(V.read_vcf(os.path.join(GOLD, "trio.1_837214.phased.vcf.gz"))[0].lines)
    for ln in V.DEFAULT_FORMAT_LINES + V.DEFAULT_INFO_LINES:
        assert ln in ref_lines, ln
    assert len([ln for ln in ref_lines if ln.startswith("##FORMAT") or ln.startswith("##INFO")]) == \
        len(V.DEFAULT_FORMAT_LINES) + len(V.DEFAULT_INFO_LINES)
    # the FILTER lines of that file are HardFilterGenotypes' text for the thresholds it was made with
    made_with = FilterArgs(minHetIndelQualityByDepth=-1, minHomIndelQualityByDepth=-1,
                           minHetSnpAltAllelicFraction=0.25, maxHetSnpAltAllelicFraction=0.75,
                           minHetIndelAltAllelicFraction=0.25, maxHetIndelAltAllelicFraction=0.75)
    assert set(V.filter_header_lines(made_with)) == {ln for ln in ref_lines if ln.startswith("##FILTER")}
    h = V.default_header(["s"], [("chr1", 249250621)], made_with)
    assert h.sorted_lines()[:2] == ["##fileformat=VCFv4.2",
                                    '##FILTER=<ID=HETINDELMAXAF,Description="Allelic fraction was above 0.750000 for a het INDEL.">']
    assert h.sorted_lines()[-1] == "##contig=<ID=chr1,length=249250621>"


def test_adding_filters_adds_eighteen_header_lines():
    """HardFilterGenotypesSuite.scala:368-375 ("test adding filters": every filter enabled -> + 18 lines)."""
    every = FilterArgs(maxSnpPhredStrandBias=20.0, maxIndelPhredStrandBias=20.0, minIndelRMSMappingQuality=30.0)
    _, header = V.load_genotypes(os.path.join(GOLD, "random.vcf.gz"))
    assert len(header.with_lines(V.filter_header_lines(every)).lines) == len(header.lines) + 18
    assert len(V.filter_header_lines(FilterArgs())) == 15          # the CLI defaults leave three filters off


def test_gvcf_extract_variants_matches_the_reference_suite():
    """SquareOffReferenceModelSuite.scala:51-76 on gvcf_multiallelic.g.vcf."""
    from avocado_b200.squareoff import extract_variants
    gts, header = V.load_genotypes(os.path.join(GOLD, "gvcf_multiallelic.g.vcf.gz"))
    assert header.samples == ["NA12878i"]
    found = extract_variants([g.variant for g in gts], [ALT in g.alleles for g in gts])
    assert len(found) == 3 and all(v.contigName == "chr22" for v in found)
    s602 = [v for v in found if v.start == 16157602]
    assert [(v.referenceAllele, v.alternateAllele, v.end) for v in s602] == [("G", "C", 16157603)]
    s095 = [v for v in found if v.start == 18030095]
    assert len(s095) == 2 and all(v.alternateAllele == "T" for v in s095)
    assert [v.end for v in s095 if v.referenceAllele == "TAA"] == [18030098]
    assert [v.end for v in s095 if v.referenceAllele == "TA"] == [18030097]
    # the multi-allelic line: one genotype per real alternate, the alternates not called here are OTHER_ALT
    multi = [g for g in gts if g.start == 18030095]
    assert [g.variant.alternateAllele for g in multi] == ["T", "TA", "TAA"]
    assert [g.alleles for g in multi] == [[OTHER_ALT, OTHER_ALT], [ALT, OTHER_ALT], [OTHER_ALT, ALT]]
    # reference-model blocks: [POS - 1, END), likelihoods against <NON_REF>
    start, length, rows = V.reference_model_blocks(gts, "NA12878i")
    assert (int(start[0]), int(length[0])) == (16157520, 16157602 - 16157520)
    assert rows[0].alleles == [REF, REF] and rows[0].minReadDepth == 2 and rows[0].readDepth == 4
    assert np.allclose(rows[0].nonReferenceLikelihoods, [0.0, 0.0, -4.5 * np.log(10.0)])
    assert np.all(start[1:] >= (start + length)[:-1])              # blocks do not overlap
    # G -> C,<NON_REF> with PL 41,3,0,41,3,41: the (ref, C) corner and the (ref, <NON_REF>) corner
    snv = [g for g in gts if g.start == 16157602][0]
    assert np.allclose(snv.genotypeLikelihoods, -np.array([41, 3, 0]) * np.log(10.0) / 10)
    assert np.allclose(snv.nonReferenceLikelihoods, -np.array([41, 41, 41]) * np.log(10.0) / 10)
    assert (snv.referenceReadDepth, snv.alternateReadDepth) == (0, 1)
    assert V.load_variants(os.path.join(GOLD, "gvcf_multiallelic.g.vcf.gz"))[0].as_tuple() == ("chr22", 16157602, 16157603, "G", "C")


@pytest.mark.parametrize("value,text", [(6.689699, "6.689699"), (-0.0, "-0.0"), (60.0, "60.0"), (59.357906, "59.357906"),
                                        (6.0934744, "6.0934744"), (0.001, "0.001"), (1.0e-4, "1.0E-4"),
                                        (1.0e7, "1.0E7"), (1234567.0, "1234567.0"), (3.4028235e38, "3.4028235E38"),
                                        (float("nan"), "NaN")])
def test_java_float_text(value, text):
    assert V.java_float_to_string(value) == text


def test_site_quality_text(tmp_path):
    from avocado_b200.joint import VariantContext
    gts, header = V.load_genotypes(os.path.join(GOLD, "trio.1_837214.vcf.gz"))
    vc = V.group_by_variant(gts)[0]
    for q, text in ((16.07, "16.07"), (564.734, "564.73"), (30.0, "30"), (float("nan"), ".")):
        line = V.format_record(VariantContext(vc.variant, q, vc.genotypes, float("nan")), header.samples)
        assert line.split("\t")[5] == text


def test_merge_variants_discovered_from_two_samples():
    """MergeDiscoveredSuite.scala:25-33: two site lists sharing one variant -> 3 distinct variants."""
    merged = V.merge_discovered([os.path.join(GOLD, "NA12878.1_907170_A2C.vcf.gz"), os.path.join(GOLD, "NA12878.1_907170_AG2A.vcf.gz")])
    assert sorted(v.as_tuple() for v in merged) == [("1", 907167, 907168, "A", "T"), ("1", 907169, 907170, "A", "C"),
                                                    ("1", 907169, 907171, "AG", "A")]


def test_reader_handles_ragged_records(tmp_path):
    """Lines the fixtures do not contain: sites only, haploid and missing GT, phased input, dropped trailing
    fields, a record without FORMAT values for one sample."""
    text = "\n".join([
        "##fileformat=VCFv4.2", "##contig=<ID=c2,length=100>", "##contig=<ID=c10,length=100>",
        "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts1\ts2",
        "c10\t5\t.\tA\tT\t12.5\t.\t.\tGT:DP:GQ\t1\t./.",
        "c2\t7\trs1\tAC\tA,<NON_REF>\t.\tPASS\tDP=3;DB\tGT:AD:PL\t0|1:3,4,0:10,0,20,30,40,50\t.",
        "c2\t9\t.\tG\t<NON_REF>\t.\t.\tEND=20\tGT:DP\t0/0:8\t0/0",
    ]) + "\n"
    path = tmp_path / "ragged.vcf"
    path.write_text(text)
    gts, header = V.load_genotypes(path)
    assert header.contigs == ["c2", "c10"] and header.samples == ["s1", "s2"] and len(gts) == 6
    g = {(x.contigName, x.start, x.sampleId): x for x in gts}
    assert g[("c10", 4, "s1")].alleles == [ALT] and g[("c10", 4, "s2")].alleles == [NO_CALL, NO_CALL]
    d = g[("c2", 6, "s1")]
    assert d.phased and d.alleles == [REF, ALT] and (d.referenceReadDepth, d.alternateReadDepth) == (3, 4)
    assert np.allclose(d.genotypeLikelihoods, -np.array([10, 0, 20]) * np.log(10) / 10)
    assert np.allclose(d.nonReferenceLikelihoods, -np.array([10, 30, 50]) * np.log(10) / 10)
    assert g[("c2", 6, "s2")].alleles == [] and g[("c2", 6, "s2")].readDepth is None
    b = g[("c2", 8, "s1")]
    assert (b.start, b.end, b.variant.alternateAllele, b.readDepth) == (8, 20, None, 8)
    assert [v.as_tuple() for v in V.load_variants(path)] == [("c10", 4, 5, "A", "T"), ("c2", 6, 8, "AC", "A")]
    # written back: contigs in header order (c2 before c10), missing genotypes as ./.
    out = tmp_path / "back.vcf"
    V.write_vcf(out, header, V.group_by_variant(gts))
    lines = [ln for ln in out.read_text().splitlines() if not ln.startswith("#")]
    assert [ln.split("\t")[0] for ln in lines] == ["c2", "c2", "c10"]
    assert lines[0].split("\t")[8:] == ["GT:AD:PL", "0|1:3,4:10,0,20", "./."]
    assert lines[2].split("\t")[8:] == ["GT", "1", "./."]
