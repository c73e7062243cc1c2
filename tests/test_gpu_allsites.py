"""GPU parity of scoreAllSites (gVCF mode, avo_call_all_sites) against the oracle: variant rows and
reference-model rows, same bar as test_gpu_parity (integers bit-exact, floating point bit-identical
to the oracle's read-order sums and within 1e-6 relative)."""
import numpy as np
import pytest

from helpers import compare_call, load_fixture, oracle_call, oracle_discover, oracle_reads_from_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from avocado_b200.genotyping import Context
    c = Context(0)
    yield c
    c.close()


def all_rows(keys_v, cols, rstart, rcols, contig):
    keys = list(keys_v) + [(contig, int(s), "N", None) for s in rstart]
    merged = {}
    for k in cols:
        merged[k] = np.concatenate([np.asarray(cols[k]), np.asarray(rcols[k])])
    return keys, merged


@pytest.mark.parametrize("seed,n_reads,contig_len,phred,min_obs", [
    (1, 6000, 30000, 25, 5), (6, 4000, 9000, 0, 2), (7, 15000, 100000, 25, 5)])
def test_synthetic_all_sites_matches_oracle(oracle, ctx, seed, n_reads, contig_len, phred, min_obs):
    from avocado_b200 import batch as B
    p = B.synth_params(seed=seed, n_reads_total=n_reads, contig_len=contig_len, first_pos=777 * seed)
    host = B.synth_host(p)
    rs = oracle_reads_from_batch(oracle, host)
    want_v = oracle.discover(rs, phred, min_obs)
    sites = ctx.discover_packed(host, phred=phred, min_obs=min_obs)
    got = [(v.start, v.end, v.referenceAllele, v.alternateAllele) for v in sites.to_variants(["ctg"])]
    want = oracle.call(rs, [(c, s, e, r, a) for c, s, e, r, a, _ in want_v], 2, [], True, 93, 93, 1)
    for batch in (host, host.to_device()):
        cols, rstart, rcols = ctx.call_all_sites_packed(batch, sites)
        assert np.all(np.diff(rstart) > 0)                       # ascending, one row per position
        keys, merged = all_rows([("ctg", s, r, a) for (s, e, r, a) in got], cols, rstart, rcols, "ctg")
        st = compare_call(keys, merged, want)
        assert st["sites"] > sites.n * 20
    # the variant rows are the ones variant-only mode produces
    plain = ctx.call_packed(host, sites)
    for k in plain:
        assert np.array_equal(np.asarray(plain[k]), np.asarray(cols[k]), equal_nan=True), k
    # two-phase capacity
    from avocado_b200 import _lib as L
    c2, r2, rc2 = ctx.call_all_sites_packed(host, sites, capacity=10)
    assert np.array_equal(r2, rstart)


GVCF = [("NA12878_snp_A2G_chr20_225058", None, 6, (225000, 225120, 120, 119)),   # BiallelicGenotyperSuite.scala:412-443
        ("NA12878.chr1.905130", 30, 4, (905100, 905200, 98, 95))]                # :746-778


@pytest.mark.parametrize("name,phred,min_obs,expect", GVCF)
def test_reference_gvcf_expectations_and_oracle_parity(oracle, ctx, name, phred, min_obs, expect):
    from avocado_b200.genotyping import BiallelicGenotyper
    recs = [r for r in load_fixture(name)[0] if r.mapq is not None and r.mapq > 0 and r.readMapped]
    recs = sorted(recs, key=lambda r: (r.contigName, r.start))
    gts = BiallelicGenotyper.discoverAndCall(recs, scoreAllSites=True, optPhredThreshold=phred,
                                             optMinObservations=min_obs, samples=["s"], ctx=ctx)
    lo, hi, n_sites, n_ref_model = expect
    window = [g for g in gts if lo <= g.start < hi]
    assert len(window) == n_sites
    assert sum(1 for g in window if g.variant.alternateAllele is None) == n_ref_model
    assert sum(1 for g in window if g.alleles.count("REF") == 2) == n_ref_model
    # every genotype against the oracle
    want_v = oracle_discover(oracle, recs, phred, min_obs)
    want = oracle_call(oracle, recs, want_v, score_all=True)
    okey = {(c, int(s), r, a): j for j, (c, s, r, a) in enumerate(
        zip(want["contigName"], want["start"], want["referenceAllele"], want["alternateAllele"]))}
    assert {(g.contigName, g.start, g.variant.referenceAllele, g.variant.alternateAllele) for g in gts} == set(okey)
    for g in gts:
        j = okey[(g.contigName, g.start, g.variant.referenceAllele, g.variant.alternateAllele)]
        assert g.alleles == ["ALT"] * int(want["nAlt"][j]) + ["REF"] * int(want["nRef"][j]) + \
            ["OTHER_ALT"] * int(want["nOther"][j])
        assert g.genotypeQuality == want["genotypeQuality"][j]
        assert (g.readDepth, g.referenceReadDepth, g.alternateReadDepth) == \
            (want["totalCoverage"][j], want["otherCoverage"][j], want["alleleCoverage"][j])
        assert g.strandBiasComponents == list(want["strandBiasComponents"][j])
        cn = int(want["copyNumber"][j])
        assert g.genotypeLikelihoods == list(want["genotypeLikelihoods"][j][:cn + 1])
        assert g.nonReferenceLikelihoods == list(want["nonReferenceLikelihoods"][j][:cn + 1])
