"""GPU parity of the joint caller (avo_joint) against the oracle's annotateSite
(JointAnnotatorCaller.scala:74-281): integers (re-called alleles, GQ, flags) bit-exact, floating
point (MAF, priors, posteriors, quality) within 1e-6 relative (helpers.REL_TOL)."""
import math

import numpy as np
import pytest

from helpers import REL_TOL

pytestmark = pytest.mark.gpu
A = {"REF": 0, "ALT": 1, "OTHER_ALT": 2, "NO_CALL": 3}


@pytest.fixture(scope="module")
def ctx():
    from avocado_b200.genotyping import Context
    c = Context(0)
    yield c
    c.close()


def close(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    return bool(np.all(same | (rel <= REL_TOL)))


def random_cohort(rng, S, N, ploidies=(2,), p_missing=0.1, p_other=0.05, p_nocall=0.03):
    ns = max(ploidies) + 1
    cols = {k: np.zeros((S, N), np.int32) for k in ("n_alt", "n_ref", "n_other", "n_nocall")}
    present = (rng.random((S, N)) >= p_missing).astype(np.uint8)
    gl = np.zeros((S, N, ns))
    maf = rng.beta(0.5, 5.0, N)
    maf[rng.random(N) < 0.15] = 0.0          # reference-only sites are discarded
    maf[rng.random(N) < 0.05] = 1.0          # undefined priors
    for s in range(S):
        cn = rng.choice(ploidies, N)
        for i in range(N):
            k = int(cn[i])
            al = rng.random(k) < maf[i]
            n_alt = int(al.sum())
            n_oth = n_no = 0
            r = rng.random()
            if r < p_other:
                n_oth = 1
            elif r < p_other + p_nocall:
                n_no = 1
            n_alt = min(n_alt, k - n_oth - n_no)
            cols["n_alt"][s, i], cols["n_other"][s, i], cols["n_nocall"][s, i] = n_alt, n_oth, n_no
            cols["n_ref"][s, i] = k - n_alt - n_oth - n_no
            gl[s, i, :k + 1] = -np.abs(rng.normal(0, 8, k + 1))
            gl[s, i, n_alt] = -abs(rng.normal(0, 0.2))
    return cols, present, gl, ns


def oracle_sites(oracle, cols, present, gl):
    S, N = present.shape
    want = []
    for i in range(N):
        gts, idx = [], []
        for s in range(S):
            if not present[s, i]:
                continue
            alleles = [1] * int(cols["n_alt"][s, i]) + [0] * int(cols["n_ref"][s, i]) + \
                      [2] * int(cols["n_other"][s, i]) + [3] * int(cols["n_nocall"][s, i])
            gts.append((alleles, gl[s, i, :len(alleles) + 1].tolist()))
            idx.append(s)
        want.append((idx, oracle.annotate_site(gts) if gts else None))
    return want


def compare(out, want, cols, present):
    from avocado_b200 import _lib as L
    n_checked = 0
    for i, (idx, site) in enumerate(want):
        if site is None:       # no genotypes at all: 0/0 alleles -> NaN MAF; nothing to emit per sample
            continue
        assert bool(out["site_emitted"][i]) == site["emitted"], i
        assert close(out["maf"][i], site["maf"])
        if not site["emitted"]:
            continue
        assert close(out["quality"][i], site["quality"]), (i, out["quality"][i], site["quality"])
        for s, g in zip(idx, site["genotypes"]):
            fl = int(out["flags"][s, i])
            assert bool(fl & L.JOINT_POSTERIORS) == (g["posteriors"] is not None)
            assert bool(fl & L.JOINT_PRIORS) == (g["priors"] is not None)
            assert bool(fl & L.JOINT_RECALLED) == (g["genotypeQuality"] is not None)
            if g["posteriors"] is not None:
                k = len(g["posteriors"])
                assert close(out["posteriors"][s, i, :k], g["posteriors"])
            if g["priors"] is not None:
                assert close(out["priors"][s, i, :len(g["priors"])], g["priors"])
            if g["genotypeQuality"] is not None:
                assert int(out["gq"][s, i]) == g["genotypeQuality"]
                assert [1] * int(out["n_alt"][s, i]) + [0] * int(out["n_ref"][s, i]) == g["alleles"]
            n_checked += 1
    return n_checked


@pytest.mark.parametrize("seed,S,N,ploidies", [(1, 3, 400, (2,)), (2, 17, 300, (2,)), (3, 8, 300, (1, 2, 3)),
                                               (4, 100, 150, (2,))])
def test_joint_matches_oracle(oracle, ctx, seed, S, N, ploidies):
    from avocado_b200.joint import joint_call
    rng = np.random.default_rng(seed)
    cols, present, gl, ns = random_cohort(rng, S, N, ploidies)
    out = joint_call(ctx, N, S, ns, cols["n_alt"].ravel(), cols["n_ref"].ravel(), cols["n_other"].ravel(),
                     gl.ravel(), present.ravel(), cols["n_nocall"].ravel())
    want = oracle_sites(oracle, cols, present, gl)
    assert compare(out, want, cols, present) > 100
    # device-resident input gives the same bytes
    import torch
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out_d = joint_call(ctx, N, S, ns, d(cols["n_alt"].ravel()), d(cols["n_ref"].ravel()),
                       d(cols["n_other"].ravel()), d(gl.ravel()), d(present.ravel()), d(cols["n_nocall"].ravel()))
    for k in out:
        assert np.array_equal(np.asarray(out[k]), out_d[k].cpu().numpy(), equal_nan=True), k


def test_joint_sample_shards_reproduce_the_cohort(ctx):
    """Sharding the samples (what each rank of the multi-GPU job holds) and feeding the summed
    counts back gives the cohort's per-sample results exactly and its quality to rounding."""
    from avocado_b200.joint import joint_call, joint_counts
    rng = np.random.default_rng(7)
    S, N = 12, 500
    cols, present, gl, ns = random_cohort(rng, S, N)
    args = lambda lo, hi: (N, hi - lo, ns, cols["n_alt"][lo:hi].ravel(), cols["n_ref"][lo:hi].ravel(),
                           cols["n_other"][lo:hi].ravel(), gl[lo:hi].ravel(), present[lo:hi].ravel(),
                           cols["n_nocall"][lo:hi].ravel())
    whole = joint_call(ctx, *args(0, S))
    shards = [(0, 5), (5, 12)]
    counts = [joint_counts(ctx, *args(lo, hi)) for lo, hi in shards]
    alt = sum(c[0].astype(np.int64) for c in counts).astype(np.int32)
    called = sum(c[1].astype(np.int64) for c in counts).astype(np.int32)
    parts = [joint_call(ctx, *args(lo, hi), totals=(alt, called)) for lo, hi in shards]
    for k in ("flags", "posteriors", "priors", "n_alt", "n_ref", "gq"):
        assert np.array_equal(np.concatenate([p[k] for p in parts]), whole[k], equal_nan=True), k
    q = (-10.0 / math.log(10.0)) * sum(p["post0_sum"] for p in parts)
    assert close(q, whole["quality"])
    assert np.array_equal(parts[0]["site_emitted"], whole["site_emitted"])


def test_joint_record_api_follows_the_reference_suite(ctx):
    """JointAnnotatorCallerSuite.scala:158-210: priors at MAF 0.1 and the re-call to hom-ref."""
    from avocado_b200.joint import JointAnnotatorCaller
    from avocado_b200.records import Genotype, Variant
    v = Variant("1", 100, 101, "A", "T")

    def gt(sample, alleles, gl):
        return Genotype(variant=v, contigName="1", start=100, end=101, sampleId=sample, alleles=alleles,
                        genotypeQuality=0, genotypeLikelihoods=gl, nonReferenceLikelihoods=[0.0] * len(gl),
                        strandBiasComponents=[0] * 4, readDepth=0, referenceReadDepth=0,
                        alternateReadDepth=0, rmsMapQ=0.0, fisherStrandBiasPValue=0.0)
    gts = [gt("s0", ["REF", "ALT"], [math.log(0.25), math.log(0.5), math.log(0.25)])] + \
          [gt("s%d" % k, ["REF", "REF"], [0.0, -10.0, -20.0]) for k in range(1, 5)]
    sites = JointAnnotatorCaller.apply(gts, ctx=ctx)
    assert len(sites) == 1 and abs(sites[0].minorAlleleFrequency - 0.1) < 1e-12
    g0 = [g for g in sites[0].genotypes if g.sampleId == "s0"][0]
    want_prior = [2 * math.log(0.9), math.log(2 * 0.9 * 0.1), 2 * math.log(0.1)]
    assert close(g0.genotypePriors, want_prior)
    assert g0.nonReferenceLikelihoods == [] and len(g0.genotypePosteriors) == 3
    assert abs(sum(math.exp(x) for x in g0.genotypePosteriors) - 1.0) < 1e-6
    assert JointAnnotatorCaller.apply([gt("s0", ["REF", "REF"], [0.0, -1.0, -2.0])], ctx=ctx) == []


def test_site_depth_roll_up_matches_oracle(oracle, ctx):
    """avo_joint_depths = JointAnnotatorCaller.calculateAnnotations (:85-90, :141-147; VariantSummary.scala:39-146)
    at every site of a random (samples x sites) matrix with unset fields and absent genotypes, host and
    device memory; then the record-level mirror against the host restatement."""
    import torch
    from avocado_b200 import _lib as L
    from avocado_b200.joint import joint_depths
    rng = np.random.default_rng(12)
    S, N = 7, 3000
    present = (rng.random((S, N)) > 0.25).astype(np.uint8)
    present[:, :40] = 0
    present[0, 20:40] = 1                                   # sites with no and with exactly one genotype
    dp = np.where(rng.random((S, N)) > 0.2, rng.integers(0, 300, (S, N)), -1).astype(np.int32)
    refdp = np.where(rng.random((S, N)) > 0.3, rng.integers(0, 200, (S, N)), -1).astype(np.int32)
    sb = rng.integers(0, 80, (S, N, 4)).astype(np.int32)
    sb[rng.random((S, N)) > 0.6] = -1
    out = joint_depths(ctx, N, S, dp.ravel(), refdp.ravel(), sb.ravel(), present.ravel())
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a.ravel())).cuda()
    out_d = joint_depths(ctx, N, S, d(dp), d(refdp), d(sb), d(present))
    for k in out:
        assert np.array_equal(np.asarray(out[k]), out_d[k].cpu().numpy()), k
    names = dict(zip(L.SITE_DEPTH_COLUMNS, ("readDepth", "referenceReadDepth", "forwardReadDepth", "reverseReadDepth",
                                            "forwardReferenceReadDepth", "reverseReferenceReadDepth")))
    opt = lambda v: None if v < 0 else int(v)
    for i in range(N):
        idx = [s for s in range(S) if present[s, i]]
        assert bool(out["annotated"][i]) == (len(idx) > 1)
        if not idx:
            assert all(out[c][i] == -1 for c in L.SITE_DEPTH_COLUMNS)
            continue
        want = oracle.variant_summary([(opt(dp[s, i]), opt(refdp[s, i]), [] if sb[s, i, 0] < 0 else sb[s, i].tolist())
                                       for s in idx])
        for c, name in names.items():
            assert opt(out[c][i]) == want[name], (i, c)
    # all genotypes present, no optional column at all
    bare = joint_depths(ctx, N, S, dp.ravel())
    assert np.all(bare["annotated"] == 1) and np.all(bare["ref_read_depth"] == -1) and np.all(bare["fwd_read_depth"] == -1)
    assert np.array_equal(bare["read_depth"], np.where((dp >= 0).any(0), np.where(dp >= 0, dp, 0).sum(0), -1))


def test_joint_caller_attaches_the_depth_annotation(ctx):
    from avocado_b200.joint import JointAnnotatorCaller, calculate_annotations
    from avocado_b200.records import ALT, REF, Genotype, Variant
    ln = np.log
    def gt(sample, pos, alleles, dp, refdp, sb):
        v = Variant("ctg", pos, pos + 1, "A", "T")
        return Genotype(variant=v, contigName="ctg", start=pos, end=pos + 1, sampleId=sample, alleles=alleles,
                        genotypeQuality=30, genotypeLikelihoods=[ln(0.1), ln(0.8), ln(0.1)], nonReferenceLikelihoods=[],
                        strandBiasComponents=sb, readDepth=dp, referenceReadDepth=refdp, alternateReadDepth=None,
                        rmsMapQ=60.0, fisherStrandBiasPValue=0.0)
    gts = [gt("a", 10, [REF, ALT], 10, 5, [2, 3, 4, 1]), gt("b", 10, [REF, ALT], 12, None, []),
           gt("c", 10, [ALT, ALT], None, 7, [3, 4, 2, 3]), gt("a", 20, [REF, ALT], 9, 4, [1, 1, 1, 1])]
    sites = {vc.variant.start: vc for vc in JointAnnotatorCaller.apply(gts, ctx=ctx)}
    assert sites[10].annotation == calculate_annotations(gts[:3]) == \
        {"readDepth": 22, "referenceReadDepth": 12, "forwardReadDepth": 11, "reverseReadDepth": 11,
         "referenceForwardReadDepth": 5, "referenceReverseReadDepth": 7}
    assert sites[20].annotation is None                     # one genotype at the site: no roll-up (:85-90)
    with pytest.raises(ValueError):
        JointAnnotatorCaller.apply([gt("a", 10, [REF, ALT], 1, 1, [0, 1, 2]), gt("b", 10, [REF, ALT], 1, 1, [])], ctx=ctx)
