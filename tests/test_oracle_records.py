"""The reference's record-level suites, on the oracle and on the host-side mirrors:
CopyNumberMapSuite.scala:27-119, VariantSummarySuite.scala:30-141, ObservationSuite.scala:24-120,
DiscoveredVariantSuite.scala:24-33."""
import pytest

from avocado_b200.genotyping import CopyNumberMap
from avocado_b200.joint import VariantSummary, calculate_annotations
from avocado_b200.records import Genotype, Variant

MIX = [("chr1", 100, 201, "DIP"), ("chr1", 1000, 2000, "DUP"), ("chr1", 2000, 3000, "DEL"), ("chr2", 2000, 3000, "DEL")]


def test_copy_number_map_empty_and_diploid_only(oracle):                          # :27-53
    e = oracle.copy_number_map(2, [])
    assert (e["basePloidy"], e["minPloidy"], e["maxPloidy"], e["variantsByContig"]) == (2, 2, 2, {})
    d = oracle.copy_number_map(2, [("chr1", 100, 201, "DIP")], ("chr1", 100, 201))
    assert (d["minPloidy"], d["maxPloidy"], d["variantsByContig"], d["overlapping"]) == (2, 2, {}, [])
    m = CopyNumberMap.empty(2)
    assert (m.basePloidy, m.minPloidy, m.maxPloidy, m.variantsByContig) == (2, 2, 2, {})
    m = CopyNumberMap.from_features(2, [("chr1", 100, 201, "DIP")])
    assert (m.minPloidy, m.maxPloidy, m.variantsByContig, m.overlappingVariants("chr1", 100, 201)) == (2, 2, {}, [])


def test_copy_number_map_with_a_mix_of_features(oracle):                          # :55-118
    m = CopyNumberMap.from_features(2, MIX)
    queries = {("chr1", 100, 150): [], ("chr1", 100, 2000): [(1000, 2000, 3)],
               ("chr1", 100, 2001): [(1000, 2000, 3), (2000, 3000, 1)], ("chr1", 2000, 2001): [(2000, 3000, 1)],
               ("chr2", 100, 2001): [(2000, 3000, 1)]}
    for q, want in queries.items():
        o = oracle.copy_number_map(2, MIX, q)
        assert (o["basePloidy"], o["minPloidy"], o["maxPloidy"]) == (2, 1, 3)
        assert o["variantsByContig"] == {"chr1": [(1000, 2000, 3), (2000, 3000, 1)], "chr2": [(2000, 3000, 1)]}
        assert o["overlapping"] == want
        assert m.overlappingVariants(*q) == want
    assert (m.basePloidy, m.minPloidy, m.maxPloidy) == (2, 1, 3)
    assert m.variantsByContig == {"chr1": [(1000, 2000, 3), (2000, 3000, 1)], "chr2": [(2000, 3000, 1)]}


def _gt(read_depth=None, ref_depth=None, sb=()):
    v = Variant("ctg", 1, 2, "A", "T")
    return Genotype(variant=v, contigName="ctg", start=1, end=2, sampleId="s", alleles=[], genotypeQuality=None,
                    genotypeLikelihoods=[], nonReferenceLikelihoods=[], strandBiasComponents=list(sb), readDepth=read_depth,
                    referenceReadDepth=ref_depth, alternateReadDepth=None, rmsMapQ=float("nan"),
                    fisherStrandBiasPValue=float("nan"))


def test_variant_summary_from_a_genotype(oracle):                                 # :30-69
    want = {"readDepth": 10, "referenceReadDepth": 5, "forwardReadDepth": None, "reverseReadDepth": None,
            "forwardReferenceReadDepth": None, "reverseReferenceReadDepth": None}
    assert oracle.variant_summary([(10, 5, [])]) == want
    assert VariantSummary.of(_gt(10, 5)) == VariantSummary(10, 5)
    want.update(forwardReadDepth=6, reverseReadDepth=4, forwardReferenceReadDepth=2, reverseReferenceReadDepth=3)
    assert oracle.variant_summary([(10, 5, [2, 3, 4, 1])]) == want
    assert VariantSummary.of(_gt(10, 5, [2, 3, 4, 1])) == VariantSummary(10, 5, 6, 4, 2, 3)
    with pytest.raises(ValueError):                                               # IllegalArgumentException, :61-69
        oracle.variant_summary([(None, None, [0, 1, 2])])
    with pytest.raises(ValueError):
        VariantSummary.of(_gt(sb=[0, 1, 2]))


def test_variant_summary_merge_and_annotation(oracle):                            # :71-140
    full = oracle.variant_summary([(10, 5, 6, 4, 3, 2), (12, 7, 6, 4, 6, 2)])
    assert [full[k] for k in ("readDepth", "referenceReadDepth", "forwardReadDepth", "reverseReadDepth",
                              "forwardReferenceReadDepth", "reverseReferenceReadDepth")] == [22, 12, 12, 8, 9, 4]
    va = VariantSummary(10, 5, 6, 4, 3, 2).merge(VariantSummary(12, 7, 6, 4, 6, 2)).to_annotation()
    assert va == {"readDepth": 22, "referenceReadDepth": 12, "forwardReadDepth": 12, "reverseReadDepth": 8,
                  "referenceForwardReadDepth": 9, "referenceReverseReadDepth": 4}
    part = oracle.variant_summary([(0, 1, 1, None, 3, None), (0, None, 1, None, None, None)])
    assert [part[k] for k in ("readDepth", "referenceReadDepth", "forwardReadDepth", "reverseReadDepth",
                              "forwardReferenceReadDepth", "reverseReferenceReadDepth")] == [0, 1, 2, None, 3, None]
    assert VariantSummary(0, 1, 1, None, 3, None).merge(VariantSummary(0, None, 1)) == VariantSummary(0, 1, 2, None, 3, None)
    old = {"readDepth": 6, "referenceReadDepth": 2, "forwardReadDepth": 3, "referenceForwardReadDepth": 2,
           "reverseReadDepth": 1, "referenceReverseReadDepth": 0}
    assert VariantSummary().to_annotation(old) == old                              # "should carry old fields"
    # calculateAnnotations over a site's genotypes = the oracle's reduce over the same genotypes
    gts = [_gt(10, 5, [2, 3, 4, 1]), _gt(7, 1), _gt(12, 7, [3, 4, 2, 3])]
    o = oracle.variant_summary([(g.readDepth, g.referenceReadDepth, g.strandBiasComponents) for g in gts])
    assert calculate_annotations(gts) == {"readDepth": o["readDepth"], "referenceReadDepth": o["referenceReadDepth"],
                                          "forwardReadDepth": o["forwardReadDepth"], "reverseReadDepth": o["reverseReadDepth"],
                                          "referenceForwardReadDepth": o["forwardReferenceReadDepth"],
                                          "referenceReverseReadDepth": o["reverseReferenceReadDepth"]}


Z2, Z3 = [0.0, 0.0], [0.0, 0.0, 0.0]


@pytest.mark.parametrize("args,kw", [
    ((0, 0, 0.0, [], [], [], [], 1, 0), {}),                                       # :24-28 empty likelihoods
    ((0, 0, 0.0, [0.0], [0.0], [0.0], [0.0], 1, 0), {}),                           # :30-34 1-length likelihoods
    ((0, 0, 0.0, Z3, Z3, [0.0, 0.0], [0.0, 0.0], 1, 0), {}),                       # :36-42 mismatching lengths
    ((-1, 0, 0.0, Z2, Z2, Z2, Z2, 1, 0), {}), ((0, -1, 0.0, Z2, Z2, Z2, Z2, 1, 0), {}),      # :46-53
    ((2, 0, 0.0, Z2, Z2, Z2, Z2, 1, 0), {}), ((0, 3, 0.0, Z2, Z2, Z2, Z2, 1, 2), {}),        # :55-62
    ((0, 0, -1.0, Z2, Z2, Z2, Z2, 1, 0), {}),                                      # :64-68
    ((0, 0, 0.0, Z2, Z2, Z2, Z2, 0, 0), {"totalCoverage": 0}),                     # :70-82
    ((0, 0, 0.0, Z2, Z2, Z2, Z2, -2, 4), {}), ((0, 0, 0.0, Z2, Z2, Z2, Z2, 2, -1), {}),
])
def test_observation_invariants(oracle, args, kw):
    with pytest.raises(ValueError):
        oracle.observation(*args, **kw)


def test_observation_invert_and_null_out(oracle):                                 # :84-120
    base = (0, 1, 0.0, [5.0, 6.0], [1.0, 2.0], [3.0, 4.0], [7.0, 8.0], 1, 2)
    inv = oracle.observation(*base, totalCoverage=4, op="invert")
    assert (inv["alleleForwardStrand"], inv["otherForwardStrand"], inv["squareMapQ"], inv["copyNumber"]) == (1, 0, 0.0, 1)
    assert (inv["alleleLogLikelihoods"], inv["otherLogLikelihoods"]) == ([3.0, 4.0], [1.0, 2.0])
    assert (inv["alleleCoverage"], inv["otherCoverage"], inv["totalCoverage"], inv["isRef"]) == (2, 1, 4, False)
    nul = oracle.observation(*base, totalCoverage=4, op="nullOut")
    assert (nul["alleleForwardStrand"], nul["otherForwardStrand"], nul["squareMapQ"], nul["copyNumber"]) == (0, 0, 0.0, 1)
    assert (nul["alleleLogLikelihoods"], nul["otherLogLikelihoods"]) == ([0.0, 0.0], [1.0, 2.0])
    assert (nul["alleleCoverage"], nul["otherCoverage"], nul["totalCoverage"], nul["isRef"]) == (0, 0, 4, False)


def test_discovered_variant_round_trip():                                         # DiscoveredVariantSuite.scala:24-33
    from avocado_b200.batch import sites_from_variants
    v = Variant("ctg", 100, 101, "A", "T")
    assert sites_from_variants([v]).to_variants(["ctg"])[0].as_tuple() == v.as_tuple()
