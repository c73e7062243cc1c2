"""GPU parity of TrioCaller (avo_trio_call) against the oracle on random trios (bit-exact: integer
logic) and the reference suite's cases through the record-level mirror (TrioCallerSuite.scala:109-237)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from avocado_b200.genotyping import Context
    c = Context(0)
    yield c
    c.close()


def test_random_trios_match_oracle(oracle, ctx):
    import torch
    from avocado_b200.trio import trio_call_packed
    rng = np.random.default_rng(8)
    n = 6000
    present = (rng.random((3, n)) > 0.1).astype(np.uint8)
    copies = rng.choice([1, 2, 2, 2, 2, 3], (3, n))
    na, nr, no, nn = (np.zeros((3, n), np.int32) for _ in range(4))
    for s in range(3):
        for i in range(n):
            al = rng.choice([0, 1, 2, 3], int(copies[s, i]), p=[0.45, 0.45, 0.05, 0.05])
            na[s, i], nr[s, i], no[s, i], nn[s, i] = (al == 1).sum(), (al == 0).sum(), (al == 2).sum(), (al == 3).sum()
    out = trio_call_packed(ctx, n, na.ravel(), nr.ravel(), no.ravel(), present.ravel(), nn.ravel())
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a.ravel())).cuda()
    out_d = trio_call_packed(ctx, n, d(na), d(nr), d(no), d(present), d(nn))
    for k in out:
        assert np.array_equal(out[k], out_d[k].cpu().numpy()), k
    seen = set()
    for i in range(n):
        gts = [([1] * int(na[s, i]) + [0] * int(nr[s, i]) + [2] * int(no[s, i]) + [3] * int(nn[s, i])) if present[s, i] else None
               for s in range(3)]
        w = oracle.trio_process(*gts)
        assert bool(out["kept"][i]) == w["kept"], (i, gts, w)
        if w["kept"]:
            assert int(out["child_action"][i]) == w["childAction"], (i, gts, w)
            assert [bool(out["filled"][i] >> s & 1) for s in range(3)] == w["filled"]
            seen.add(w["childAction"])
    assert seen == {0, 1, 2, 3, 4}


def test_reference_suite_cases_through_the_mirror(ctx):
    from avocado_b200.records import Genotype, Variant
    from avocado_b200.trio import TrioCaller
    v = Variant("chr", 100, 101, "A", "T")

    def gt(sample, alleles):
        return Genotype(variant=v, contigName="chr", start=100, end=101, sampleId=sample, alleles=alleles,
                        genotypeQuality=50, genotypeLikelihoods=[], nonReferenceLikelihoods=[], strandBiasComponents=[],
                        readDepth=0, referenceReadDepth=0, alternateReadDepth=0, rmsMapQ=0.0, fisherStrandBiasPValue=0.0)
    run = lambda gts: TrioCaller.apply(gts, "parent1", "parent2", "proband", ctx=ctx)
    vc, = run([gt("proband", ["REF", "ALT"])])                                            # :109-131
    assert [g.sampleId for g in vc.genotypes] == ["parent1", "parent2", "proband"]
    assert vc.genotypes[0].alleles == ["NO_CALL"] * 2 and vc.genotypes[2].alleles == ["REF", "ALT"]
    vc, = run([gt("proband", ["ALT", "REF"]), gt("parent1", ["REF", "REF"]), gt("parent2", ["REF", "ALT"])])   # :155-183
    assert vc.genotypes[2].alleles == ["REF", "ALT"] and vc.genotypes[2].phased
    vc, = run([gt("proband", ["ALT", "REF"]), gt("parent1", ["REF", "ALT"]), gt("parent2", ["REF", "ALT"])])   # :185-212
    assert vc.genotypes[2].alleles == ["ALT", "REF"] and not vc.genotypes[2].phased
    vc, = run([gt("proband", ["ALT", "ALT"]), gt("parent1", ["REF", "REF"]), gt("parent2", ["REF", "ALT"])])   # :214-237
    assert vc.genotypes[2].alleles == ["NO_CALL"] * 2 and vc.genotypes[2].genotypeQuality is None
    vc, = run([gt("proband", ["REF", "ALT"]), gt("parent1", ["REF"]), gt("parent2", ["REF", "ALT"])])          # :133-153
    assert vc.genotypes[2].alleles == ["REF", "ALT"] and not vc.genotypes[2].phased
    assert run([gt("proband", ["REF", "REF"]), gt("parent1", ["REF", "REF"])]) == []                           # :84-98


def test_trio_vcf_end_to_end_matches_the_reference_file(ctx, tmp_path):
    """TrioCallerSuite.scala:239-251 on the GPU path: loadGenotypes(trio.1_837214.vcf) -> TrioCaller
    (avo_trio_call) -> saveAsVcf, compared byte for byte with the reference's trio.1_837214.phased.vcf."""
    import os
    from avocado_b200 import vcf as V
    from avocado_b200.trio import TrioCaller
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vcf")
    import gzip
    gts, header = V.load_genotypes(os.path.join(gold, "trio.1_837214.vcf.gz"))
    contexts = TrioCaller.apply(gts, "NA12891", "NA12892", "NA12878", ctx=ctx)
    out = tmp_path / "trio.vcf"
    V.write_vcf(out, header, contexts)
    with open(out, "rb") as a, gzip.open(os.path.join(gold, "trio.1_837214.phased.vcf.gz"), "rb") as b:
        assert a.read() == b.read()
