"""Oracle restatement of SquareOffReferenceModel against the reference suite
(SquareOffReferenceModelSuite.scala:31-210) and the Python mirror of its host-side part."""
from avocado_b200.records import Variant
from avocado_b200.squareoff import extract_variants, trim_right


def test_trim_right(oracle):                                            # :33-51
    for ref, alt, want in (("A", "T", 0), ("ACG", "TAG", 1), ("TAA", "TCAA", 2), ("TCCAGT", "TG", 0),
                           ("AA", "A", 0), ("", "A", 0), ("GAA", "AA", 1)):
        assert oracle.trim_right(ref, alt) == want == trim_right(ref, alt), (ref, alt)


def test_extract_variants_trims_and_dedupes():                          # :53-81 (gvcf_multiallelic.g.vcf shape)
    vs = [Variant("chr22", 16157602, 16157603, "G", "C"), Variant("chr22", 16157602, 16157603, "G", "C"),
          Variant("chr22", 18030095, 18030098, "TAA", "T"), Variant("chr22", 18030095, 18030097, "TA", "T"),
          Variant("chr22", 100, 101, "A", "T"), Variant("chr22", 200, 203, "TAA", "TCAA"),
          Variant("chr22", 300, 301, "N", None)]
    out = extract_variants(vs, [True, True, True, True, False, True, True])
    assert [v.as_tuple() for v in out] == [("chr22", 16157602, 16157603, "G", "C"), ("chr22", 18030095, 18030098, "TAA", "T"),
                                           ("chr22", 18030095, 18030097, "TA", "T"), ("chr22", 200, 201, "T", "TC")]


def test_has_scored_variant_and_excise(oracle):                         # :83-160
    v = ("ctg", 1, 2, "A", "T")
    assert oracle.square_off_sample(v, [("ctg", 1, 2, "A", "T")]) == (0, False)
    assert oracle.square_off_sample(v, [("ctg", 1, 10, "A", None)]) == (0, True)        # not scored, but overlapping
    assert oracle.square_off_sample(("ctg", 100, 101, "A", "G"), [("ctg", 90, 110, "N", None)]) == (0, True)
    assert oracle.square_off_sample(("ctg", 100, 101, "A", "G"), [("ctg", 110, 120, "N", None)]) == (-1, False)
    assert oracle.square_off_sample(("ctg", 100, 101, "A", "G"), [("other", 90, 110, "N", None)]) == (-1, False)
    # canonical order of the excision: lowest (start, end), an alternate-carrying genotype first
    gts = [("ctg", 100, 101, "N", None), ("ctg", 98, 103, "ACGTA", "A"), ("ctg", 100, 101, "A", "C")]
    assert oracle.square_off_sample(("ctg", 100, 101, "A", "G"), gts) == (1, True)
    assert oracle.square_off_sample(("ctg", 100, 101, "A", "G"), gts[::2]) == (1, True)
    assert oracle.square_off_sample(("ctg", 100, 101, "A", "C"), gts) == (2, False)
