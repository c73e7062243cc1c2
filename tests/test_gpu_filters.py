"""GPU parity of RewriteHets + HardFilterGenotypes (avo_filter_genotypes) against the oracle
(every field bit-exact: the pass is integer / float32 comparisons), the reference suites' cases
through the record-level mirror, and the reference's one end-to-end test that uses the filters
(BiallelicGenotyperSuite.scala:445-477)."""
import dataclasses

import numpy as np
import pytest

from helpers import load_fixture
from test_oracle_filters import GTS, TEST_ARGS

pytestmark = pytest.mark.gpu
CODE = {"REF": 0, "ALT": 1, "OTHER_ALT": 2}


@pytest.fixture(scope="module")
def ctx():
    from avocado_b200.genotyping import Context
    c = Context(0)
    yield c
    c.close()


def to_oracle_args(a):
    return {f.name: getattr(a, f.name) for f in dataclasses.fields(a)}


def rand_rows(rng, n):
    copies = rng.choice([1, 2, 2, 2, 3], n)
    n_alt = np.array([rng.integers(0, c + 1) for c in copies], np.int32)
    n_other = np.array([rng.integers(0, c - a + 1) if rng.random() < 0.2 else 0 for c, a in zip(copies, n_alt)], np.int32)
    n_ref = (copies - n_alt - n_other).astype(np.int32)
    dp = rng.integers(0, 260, n).astype(np.int32)
    adp = np.minimum(dp, rng.integers(0, 260, n)).astype(np.int32)
    cols = dict(emitted=(rng.random(n) > 0.05).astype(np.uint8), n_alt=n_alt, n_ref=n_ref, n_other=n_other,
                gq=rng.integers(0, 120, n).astype(np.int32), total_cov=dp, allele_cov=adp,
                fisher=rng.choice([0.0, 10.0, 59.99, 60.0, 61.0, 250.0, np.nan, np.inf], n).astype(np.float32),
                rms_mapq=rng.choice([0.0, 25.0, 29.99, 30.0, 40.0, 60.0, np.nan, np.inf], n).astype(np.float32))
    ref_len = rng.choice([1, 1, 1, 2, 5], n).astype(np.int32)
    alt_len = rng.choice([-1, 1, 1, 1, 3], n).astype(np.int32)
    return cols, ref_len, alt_len


@pytest.mark.parametrize("seed,filter_ref,emit_all,kw", [
    (1, True, False, {}), (2, False, False, {}), (3, True, True, {}),
    (4, True, False, dict(TEST_ARGS)), (5, True, False, dict(disableHetSnpRewriting=True, minQuality=5)),
    (6, False, False, dict(maxSnpPhredStrandBias=60.0, maxIndelPhredStrandBias=200.0, minIndelRMSMappingQuality=30.0,
                           disableHetIndelRewriting=True))])
def test_filter_rows_match_oracle(oracle, ctx, seed, filter_ref, emit_all, kw):
    import torch
    from avocado_b200 import _lib as L
    from avocado_b200.filters import FilterArgs, filter_columns
    rng = np.random.default_rng(seed)
    n = 3000
    cols, ref_len, alt_len = rand_rows(rng, n)
    args = FilterArgs(**kw)
    out = filter_columns(ctx, cols, ref_len, alt_len, args, filter_ref, emit_all)
    dcols = {k: torch.from_numpy(v).cuda() for k, v in cols.items()}
    out_d = filter_columns(ctx, dcols, torch.from_numpy(ref_len).cuda(), torch.from_numpy(alt_len).cuda(), args,
                           filter_ref, emit_all)
    for k in out:
        assert np.array_equal(np.asarray(out[k]).view(np.uint8), out_d[k].cpu().numpy().view(np.uint8)), k
    oargs = to_oracle_args(args)
    n_emit = n_rew = n_fail = 0
    for i in range(n):
        if not cols["emitted"][i]:
            assert not out["emit"][i]
            continue
        g = dict(ref="A" * int(ref_len[i]), alt=None if alt_len[i] < 0 else "C" * int(alt_len[i]),
                 alleles=[1] * int(cols["n_alt"][i]) + [0] * int(cols["n_ref"][i]) + [2] * int(cols["n_other"][i]),
                 gq=int(cols["gq"][i]), dp=int(cols["total_cov"][i]), adp=int(cols["allele_cov"][i]),
                 fs=float(cols["fisher"][i]), rms=float(cols["rms_mapq"][i]))
        w = oracle.rewrite_and_filter(g, oargs, filter_ref, emit_all, True)
        assert bool(out["emit"][i]) == w["emitted"], (i, g, w)
        assert bool(out["rewritten"][i]) == w["rewritten"], (i, g, w)
        assert bool(out["gq_valid"][i]) == (w["gq"] is not None)
        assert int(out["n_alt"][i]) == sum(1 for a in w["alleles"] if a == 1)
        if w["emitted"] and g["alt"] is not None:
            failed = [L.FILTER_NAMES[b] for b in range(18) if (int(out["filters_failed"][i]) >> b) & 1]
            assert failed == w["filtersFailed"], (i, g, failed, w)
            assert bool(out["filters_applied"][i]) == w["filtersApplied"]
            n_fail += bool(failed)
        n_emit += w["emitted"]; n_rew += w["rewritten"]
    assert n_emit > 100 and n_fail > 20
    if not kw.get("disableHetSnpRewriting"):
        assert n_rew > 20


def _gt(d):
    from avocado_b200.records import Genotype, Variant
    names = {0: "REF", 1: "ALT", 2: "OTHER_ALT"}
    return Genotype(variant=Variant("1", 10, 10 + len(d["ref"]), d["ref"], d["alt"]), contigName="1", start=10,
                    end=10 + len(d["ref"]), sampleId="s", alleles=[names[a] for a in d["alleles"]],
                    genotypeQuality=d.get("gq"), genotypeLikelihoods=[], nonReferenceLikelihoods=[],
                    strandBiasComponents=[0] * 4, readDepth=d.get("dp", 0), referenceReadDepth=0,
                    alternateReadDepth=d.get("adp", 0), rmsMapQ=d.get("rms", 60.0),
                    fisherStrandBiasPValue=d.get("fs", 0.0))


def test_reference_suite_cases_through_the_mirror(ctx):
    """RewriteHetsSuite.scala:184-205 and HardFilterGenotypesSuite.scala:306-383."""
    from avocado_b200.filters import FilterArgs, HardFilterGenotypes, RewriteHets
    gts = [_gt(g) for g in GTS.values()]
    out = RewriteHets(gts, FilterArgs(maxHetSnpAltAllelicFraction=0.75, maxHetIndelAltAllelicFraction=0.65), ctx=ctx)
    touched = [g for g in out if g.genotypeQuality is None]
    assert len(out) == 8 and len(touched) == 2 and all(set(g.alleles) == {"ALT"} for g in touched)
    untouched = [g for g in out if g.genotypeQuality is not None]
    assert all(g in gts for g in untouched)
    off = RewriteHets(gts, FilterArgs(disableHetSnpRewriting=True, disableHetIndelRewriting=True), ctx=ctx)
    assert off == gts
    args = FilterArgs(**dict(TEST_ARGS, minQuality=20))
    snp = _gt(dict(ref="A", alt="T", alleles=[1, 0], gq=30, adp=15, dp=20, rms=35.0, fs=70.0))
    r, = HardFilterGenotypes([snp], args, ctx=ctx)
    assert r.filtersApplied and not r.filtersPassed
    assert set(r.filtersFailed) == {"SNPMQ", "HETSNPQD", "SNPFS", "HETSNPMAXAF"}
    indel = _gt(dict(ref="A", alt="AT", alleles=[1, 0], gq=30, adp=10, dp=20, rms=35.0, fs=70.0))
    r, = HardFilterGenotypes([indel], args, ctx=ctx)
    assert r.filtersPassed and r.filtersFailed is None
    ref = _gt(dict(ref="A", alt="T", alleles=[0, 0], gq=99))
    assert HardFilterGenotypes([ref], args, True, ctx=ctx) == []
    assert len(HardFilterGenotypes([ref], args, False, ctx=ctx)) == 1


def test_discover_and_call_short_indel_with_hard_filters(ctx):
    """BiallelicGenotyperSuite.scala:445-477: after HardFilterGenotypes exactly one passing
    genotype with alternate depth remains, the het insertion A -> AGTTTT at chr1:832735."""
    from avocado_b200.filters import FilterArgs, HardFilterGenotypes
    from avocado_b200.genotyping import BiallelicGenotyper
    recs = [r for r in load_fixture("NA12878.chr1.832736")[0] if r.mapq is not None and r.mapq > 0]
    gts = BiallelicGenotyper.discoverAndCall(recs, samples=["s"], ctx=ctx)
    args = FilterArgs(**dict(TEST_ARGS, minHetIndelAltAllelicFraction=-1.0, maxHetIndelAltAllelicFraction=-1.0,
                             minHomIndelAltAllelicFraction=-1.0))
    kept = [g for g in HardFilterGenotypes(gts, args, ctx=ctx) if g.filtersPassed and g.alternateReadDepth != 0]
    assert len(kept) == 1
    g = kept[0]
    assert (g.start, g.end, g.variant.referenceAllele, g.variant.alternateAllele) == (832735, 832736, "A", "AGTTTT")
    assert g.alleles.count("REF") == 1 and g.alleles.count("ALT") == 1
