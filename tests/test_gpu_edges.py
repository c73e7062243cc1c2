"""GPU edge cases and full-size properties of the hot path.

Edge cases the reference's tests touch (copy-number variants over site boundaries
BiallelicGenotyperSuite.scala:145-213, reads the reference skips, empty inputs) and, at the
BASELINE.json configs[1] size, properties that do not need the oracle: determinism, invariance to
region sharding, and internal consistency of every emitted row (recomputed with torch in fp64)."""
import math

import numpy as np
import pytest

from helpers import compare_call, oracle_reads_from_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from avocado_b200.genotyping import Context
    c = Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("ploidy,cnvs", [
    (2, [("ctg", 3000, 9000, 3), ("ctg", 15000, 15500, 1)]),       # DUP and DEL inside the batch
    (2, [("ctg", 0, 100000, 3)]),                                     # everything duplicated
    (1, []), (3, []), (4, [("ctg", 10000, 20000, 2)]),
    (2, [("other", 0, 100000, 5), ("ctg", 7000, 7200, 1)])])          # a CNV on another contig widens the ploidy range
def test_copy_number_and_ploidy(oracle, ctx, ploidy, cnvs):
    from avocado_b200 import batch as B
    p = B.synth_params(seed=31, n_reads_total=8000, contig_len=30000, first_pos=1000)
    host = B.synth_host(p)
    rs = oracle_reads_from_batch(oracle, host)
    want_v = oracle.discover(rs, 25, 5)
    sites = ctx.discover_packed(host, phred=25, min_obs=5)
    want = oracle.call(rs, [(c, s, e, r, a) for c, s, e, r, a, _ in want_v], ploidy, list(cnvs), False, 93, 93, 1)
    cols = ctx.call_packed(host, sites, ploidy=ploidy, cnvs=cnvs, contig_rank={"ctg": 0, "other": 1})
    keys = [("ctg", v.start, v.referenceAllele, v.alternateAllele) for v in sites.to_variants(["ctg"])]
    st = compare_call(keys, cols, want)
    assert st["sites"] > 10
    if len(cnvs) == 2 and cnvs[0][0] == "ctg":
        assert len(set(np.asarray(cols["copy_number"])[np.asarray(cols["emitted"]).astype(bool)].tolist())) > 1
    # gVCF rows carry the copy number of their position too
    want_all = oracle.call(rs, [(c, s, e, r, a) for c, s, e, r, a, _ in want_v], ploidy, list(cnvs), True, 93, 93, 1)
    c2, rstart, rcols = ctx.call_all_sites_packed(host, sites, ploidy=ploidy, cnvs=cnvs,
                                                  contig_rank={"ctg": 0, "other": 1})
    keys2 = keys + [("ctg", int(s), "N", None) for s in rstart]
    merged = {k: np.concatenate([np.asarray(c2[k]), np.asarray(rcols[k])]) for k in c2}
    compare_call(keys2, merged, want_all)


def test_quality_and_mapq_caps(oracle, ctx):
    """call(maxQuality, maxMapQ) (BG:95-96, 451-452): clamping through least()."""
    from avocado_b200 import batch as B
    p = B.synth_params(seed=5, n_reads_total=5000, contig_len=20000)
    host = B.synth_host(p)
    rs = oracle_reads_from_batch(oracle, host)
    want_v = oracle.discover(rs, 25, 5)
    sites = ctx.discover_packed(host, phred=25, min_obs=5)
    keys = [("ctg", v.start, v.referenceAllele, v.alternateAllele) for v in sites.to_variants(["ctg"])]
    for mq, mm in ((30, 50), (93, 20), (10, 10)):
        want = oracle.call(rs, [(c, s, e, r, a) for c, s, e, r, a, _ in want_v], 2, [], False, mq, mm, 1)
        compare_call(keys, ctx.call_packed(host, sites, max_quality=mq, max_mapq=mm), want)


def test_reads_the_reference_skips_and_empty_inputs(oracle, ctx):
    from avocado_b200 import _lib as L, batch as B
    from avocado_b200.genotyping import BiallelicGenotyper, DiscoverVariants
    from avocado_b200.records import AlignmentRecord, Variant

    def read(start, seq, cigar, md, **kw):
        d = dict(readMapped=True, contigName="1", start=start, end=start + len(seq), sequence=seq,
                 qual="I" * len(seq), cigar=cigar, mismatchingPositions=md, mapq=40)
        d.update(kw)
        return AlignmentRecord(**d)
    good = [read(100, "ACGTACGTAC", "10M", "4A5") for _ in range(3)]
    odd = [read(100, "ACGTACGTAC", "10M", None),                     # no MD: no operators
           read(100, "ACGTACGTAC", "*", "10"),                       # no CIGAR
           read(100, "ACGTACGTAC", "4M2N4M", "8"),                   # unsupported CIGAR operator
           read(100, "ACGTACGTAC", "10M", "4A5", mapq=None),         # observeRead throws: skipped by call
           read(100, "ACGTA", "10M", "4A5"),                         # sequence shorter than the CIGAR
           AlignmentRecord(readMapped=False, sequence="ACGT", qual="IIII")]
    recs = good + odd
    # (a sequence shorter than its CIGAR crashes the reference's discovery job; the packer drops
    # such a read instead, so the oracle is not asked about it there)
    no_short = [r for r in recs if r.sequence != "ACGTA"]
    want_v = oracle.discover(oracle.reads_from_dicts([r.as_dict() for r in no_short]), 0, None)
    got_v = DiscoverVariants(recs, ctx=ctx)
    assert sorted(v.as_tuple() for v in got_v) == sorted((c, s, e, r, a) for c, s, e, r, a, _ in want_v)
    assert got_v == DiscoverVariants(no_short, ctx=ctx)
    gts = BiallelicGenotyper.call(recs, [Variant("1", 104, 105, "A", "A")], samples=["s"], ctx=ctx)
    want = oracle.call(oracle.reads_from_dicts([r.as_dict() for r in recs]), [("1", 104, 105, "A", "A")],
                       2, [], False, 93, 93, 1)
    assert len(gts) == len(want["start"]) == 1 and gts[0].readDepth == want["totalCoverage"][0]
    # no mapped read at all / an empty packed batch
    assert DiscoverVariants([odd[-1]], ctx=ctx) == []
    empty = B.pack_reads([])
    assert empty.n_reads == 0 and ctx.discover_packed(empty).n == 0
    table = B.sites_from_variants([Variant("1", 104, 105, "A", "C")], {"1": 0})
    cols = ctx.call_packed(empty, table)
    assert not cols["emitted"].any()
    # parameters outside the supported domain are errors, not crashes
    for kw in (dict(ploidy=0), dict(ploidy=L.MAX_PLOIDY + 1), dict(max_quality=0), dict(max_mapq=200)):
        with pytest.raises(L.AvocadoError):
            ctx.call_packed(B.pack_reads(good), table, **kw)
    unsorted = B.sites_from_variants([Variant("1", 104, 105, "A", "C"), Variant("1", 102, 103, "G", "C")], {"1": 0})
    unsorted.start = np.ascontiguousarray(unsorted.start[::-1])
    with pytest.raises(L.AvocadoError) as e:
        ctx.call_packed(B.pack_reads(good), unsorted)
    assert e.value.status == L.AVO_E_UNSORTED


def test_full_size_properties(ctx):
    """BASELINE.json configs[1] (25.6 M reads, 64 Mb): the oracle cannot run at this size, so the
    check is through properties: bit-identical repeat, invariance of a region's rows to computing
    only that region (the sharding property of SURVEY.md 8e), and every row recomputed from its
    own aggregates in torch fp64 (blend / arg-max / quality, strand components, RMS mapQ)."""
    import torch
    from avocado_b200 import batch as B
    n, clen = 25_600_000, 64_000_000
    p = B.synth_params(seed=0xA70CAD0 + 1, contig_len=clen, n_reads_total=n)
    dev = B.synth_device(p)
    s1, c1 = ctx.discover_and_call_packed(dev, phred=25, min_obs=5, capacity=1 << 18, device_out=True, copy=False)
    n_sites = s1.n
    assert 50_000 < n_sites < 100_000
    snap = {k: v.clone() for k, v in c1.items()}
    start1 = s1.start[:n_sites].clone()
    s2, c2 = ctx.discover_and_call_packed(dev, phred=25, min_obs=5, capacity=1 << 18, device_out=True, copy=False)
    assert s2.n == n_sites and torch.equal(s2.start[:n_sites], start1)
    bits = lambda t: t.contiguous().view(-1).view(torch.uint8)
    for k, v in c2.items():
        assert torch.equal(bits(snap[k]), bits(v)), k                        # bit-identical repeat
    em = snap["emitted"].bool()
    assert int(em.sum()) == n_sites                                      # every discovered site is observed
    ns = 3
    al, rf, ot = (snap[k].view(n_sites, ns) for k in ("allele_ll", "ref_ll", "other_ll"))
    gl = al + rf.flip(1)
    assert torch.equal(gl, snap["gl"].view(n_sites, ns))                 # BG:728, bit-exact
    k10 = 10.0 / math.log(10.0)

    def qual_state(x):
        top2 = torch.topk(x, 2, dim=1).values
        return k10 * (top2[:, 0] - top2[:, 1]), torch.argmax(x, dim=1)
    q_ar, s_ar = qual_state(gl)
    q_ao, _ = qual_state(al + ot.flip(1))
    q_or, _ = qual_state(ot + rf.flip(1))
    qual = torch.maximum(torch.maximum(q_ar, q_ao), q_or)
    assert torch.equal(qual, snap["qual"])                               # BG:598-615, 665
    assert torch.equal(qual.to(torch.int32), snap["gq"])                 # toInt truncation
    pick = (q_ar >= q_ao) & (q_ar >= q_or)
    assert torch.equal(snap["n_alt"][pick], s_ar[pick].to(torch.int32))
    assert torch.equal(snap["n_alt"] + snap["n_ref"] + snap["n_other"], snap["copy_number"])
    sb = snap["sb"].view(n_sites, 4)
    assert torch.equal(sb[:, 0] + sb[:, 1], snap["other_cov"]) and torch.equal(sb[:, 2] + sb[:, 3], snap["allele_cov"])
    assert bool((snap["allele_cov"] + snap["other_cov"] <= snap["total_cov"]).all())
    rms = torch.sqrt(snap["square_mapq"] / (snap["allele_cov"] + snap["other_cov"]).double()).float()
    assert torch.equal(rms, snap["rms_mapq"])
    assert 55 < float(snap["total_cov"].double().mean()) < 62           # 60x coverage, minus mapQ-0 / Q0 drops

    # a region computed on its own (reads overlapping it, with halo) gives the same rows
    host_meta = dev.meta.view(torch.int32).view(-1, 8)
    st = host_meta[:, 0].contiguous()
    lo_pos, hi_pos = 20_000_000, 21_000_000
    r_lo = int(torch.searchsorted(st, torch.tensor(lo_pos - 400, device=st.device, dtype=st.dtype)))
    r_hi = int(torch.searchsorted(st, torch.tensor(hi_pos + 400, device=st.device, dtype=st.dtype)))
    sub = B.synth_device(p, first=r_lo, n=r_hi - r_lo)                   # the same reads, generated as their own batch
    s3, c3 = ctx.discover_and_call_packed(sub, phred=25, min_obs=5, capacity=1 << 18, device_out=True, copy=False)
    m1 = (start1 >= lo_pos) & (start1 < hi_pos)
    m3 = (s3.start[:s3.n] >= lo_pos) & (s3.start[:s3.n] < hi_pos)
    assert int(m1.sum()) == int(m3.sum()) > 500
    assert torch.equal(start1[m1], s3.start[:s3.n][m3])
    for k in ("total_cov", "allele_cov", "gq", "qual", "allele_ll", "ref_ll", "fisher"):
        assert torch.equal(bits(snap[k][m1]), bits(c3[k][m3])), k


def test_forest_quirk_and_multi_contig_table(oracle, ctx):
    """Forest.get returns only the contiguous run of overlapping entries around the first probe hit
    (TreeRegionJoin.scala:74-119), over the GLOBAL sorted table of all contigs.  Long deletions that
    reach across shorter non-overlapping variants make the result depend on the probe order; the
    oracle replays it literally, and so must the kernels (fast path + replay)."""
    from avocado_b200.genotyping import BiallelicGenotyper
    from avocado_b200.records import AlignmentRecord, Variant
    rng = np.random.default_rng(17)
    bases = "ACGT"
    ref = {c: "".join(bases[i] for i in rng.integers(0, 4, 1400)) for c in ("chrA", "chrB", "chrC")}

    def read(contig, start, length, name, edits=()):
        seq = list(ref[contig][start:start + length])
        md, run = "", 0
        for k in range(length):
            if k in edits:
                b = bases[(bases.index(seq[k]) + 1) % 4]
                md += "%d%s" % (run, seq[k]); run = 0
                seq[k] = b
            else:
                run += 1
        md += str(run)
        return AlignmentRecord(readMapped=True, contigName=contig, start=start, end=start + length,
                               sequence="".join(seq), qual="".join(chr(33 + int(q)) for q in rng.integers(20, 41, length)),
                               cigar="%dM" % length, mismatchingPositions=md, mapq=int(rng.integers(20, 61)),
                               readName=name, readNegativeStrand=bool(rng.integers(0, 2)))
    recs = []
    for c in ref:
        for i in range(400):
            s = int(rng.integers(0, 1300))
            ln = int(rng.integers(40, 100))
            ed = set(int(x) for x in rng.integers(0, ln, 2)) if i % 3 == 0 else ()
            recs.append(read(c, s, min(ln, 1400 - s), "%s_%d" % (c, i), ed))
    variants = []
    for c in ref:
        pos = sorted(set(int(x) for x in rng.integers(5, 1350, 60)))
        for k, p in enumerate(pos):
            r = ref[c]
            if k % 7 == 0:          # long deletion reaching across the next few short variants
                ln = int(rng.integers(15, 45))
                variants.append(Variant(c, p, p + ln + 1, r[p:p + ln + 1], r[p]))
            elif k % 5 == 0:        # insertion
                variants.append(Variant(c, p, p + 1, r[p], r[p] + "GT"))
            else:                   # SNV
                variants.append(Variant(c, p, p + 1, r[p], bases[(bases.index(r[p]) + 1) % 4]))
    variants = list({v.as_tuple(): v for v in variants}.values())
    recs.sort(key=lambda r: (r.contigName, r.start))
    want = oracle.call(oracle.reads_from_dicts([r.as_dict() for r in recs]), [v.as_tuple() for v in variants],
                       2, [], False, 93, 93, 1)
    gts = BiallelicGenotyper.call(recs, variants, samples=["s"], ctx=ctx)
    okey = {(c, int(s), r, a): j for j, (c, s, r, a) in enumerate(
        zip(want["contigName"], want["start"], want["referenceAllele"], want["alternateAllele"]))}
    assert {(g.contigName, g.start, g.variant.referenceAllele, g.variant.alternateAllele) for g in gts} == set(okey)
    assert len(gts) > 100
    for g in gts:
        j = okey[(g.contigName, g.start, g.variant.referenceAllele, g.variant.alternateAllele)]
        assert (g.readDepth, g.referenceReadDepth, g.alternateReadDepth) == \
            (want["totalCoverage"][j], want["otherCoverage"][j], want["alleleCoverage"][j]), g.variant
        assert g.genotypeQuality == want["genotypeQuality"][j]
        assert g.alleles == ["ALT"] * int(want["nAlt"][j]) + ["REF"] * int(want["nRef"][j]) + \
            ["OTHER_ALT"] * int(want["nOther"][j])
        cn = int(want["copyNumber"][j])
        assert g.genotypeLikelihoods == list(want["genotypeLikelihoods"][j][:cn + 1])
    # the quirk is really exercised: for some reads Forest.get misses variants they overlap
    regions = sorted((v.contigName, v.start, v.end) for v in variants)
    quirky = 0
    for r in recs:
        got = set(oracle.forest_get(regions, (r.contigName, r.start, r.end)))
        full = {i for i, (c, s, e) in enumerate(regions) if c == r.contigName and s < r.end and e > r.start}
        quirky += got != full
    assert quirky > 20
    # gVCF rows over the same multi-contig table
    want_all = oracle.call(oracle.reads_from_dicts([r.as_dict() for r in recs]), [v.as_tuple() for v in variants],
                           2, [], True, 93, 93, 1)
    gts_all = BiallelicGenotyper.call(recs, variants, scoreAllSites=True, samples=["s"], ctx=ctx)
    okey = {(c, int(s), r, a): j for j, (c, s, r, a) in enumerate(
        zip(want_all["contigName"], want_all["start"], want_all["referenceAllele"], want_all["alternateAllele"]))}
    assert {(g.contigName, g.start, g.variant.referenceAllele, g.variant.alternateAllele) for g in gts_all} == set(okey)
    for g in gts_all:
        j = okey[(g.contigName, g.start, g.variant.referenceAllele, g.variant.alternateAllele)]
        assert g.readDepth == want_all["totalCoverage"][j] and g.genotypeQuality == want_all["genotypeQuality"][j]
        cn = int(want_all["copyNumber"][j])
        assert g.genotypeLikelihoods == list(want_all["genotypeLikelihoods"][j][:cn + 1])


def test_one_very_long_read_span(oracle, ctx):
    """A read with a huge deletion makes the look-back (max read span) cover the whole batch: the
    per-site buckets then hold a slot for every read.  Results must not change."""
    from avocado_b200.genotyping import BiallelicGenotyper
    from avocado_b200.records import AlignmentRecord
    rng = np.random.default_rng(23)
    bases = "ACGT"
    ref = "".join(bases[i] for i in rng.integers(0, 4, 61000))
    recs = []
    for i in range(900):
        s = int(rng.integers(0, 60000 - 120)) if i % 3 else int(rng.integers(100, 400))
        ln = int(rng.integers(60, 110))
        seq = list(ref[s:s + ln])
        md, run = "", 0
        for k in range(ln):
            if (s + k) % 97 == 0:                                  # a shared SNV every 97 bases
                md += "%d%s" % (run, seq[k]); run = 0
                seq[k] = bases[(bases.index(seq[k]) + 1) % 4]
            else:
                run += 1
        md += str(run)
        recs.append(AlignmentRecord(readMapped=True, contigName="1", start=s, end=s + ln, sequence="".join(seq),
                                    qual="".join(chr(33 + int(q)) for q in rng.integers(25, 41, ln)),
                                    cigar="%dM" % ln, mismatchingPositions=md, mapq=50, readName="r%d" % i))
    # the long one: 50 bases, a 50 kb deletion, 50 bases
    s = 200
    seq = ref[s:s + 50] + ref[s + 50 + 50000:s + 100 + 50000]
    recs.append(AlignmentRecord(readMapped=True, contigName="1", start=s, end=s + 100 + 50000, sequence=seq,
                                qual="I" * 100, cigar="50M50000D50M",
                                mismatchingPositions="50^%s50" % ref[s + 50:s + 50 + 50000], mapq=50, readName="long"))
    recs.sort(key=lambda r: r.start)
    want_v = oracle.discover(oracle.reads_from_dicts([r.as_dict() for r in recs]), 20, 2)
    want = oracle.call(oracle.reads_from_dicts([r.as_dict() for r in recs]),
                       [(c, s_, e, r, a) for c, s_, e, r, a, _ in want_v], 2, [], False, 93, 93, 1)
    gts = BiallelicGenotyper.discoverAndCall(recs, optPhredThreshold=20, optMinObservations=2, samples=["s"], ctx=ctx)
    okey = {(int(s_), r, a): j for j, (s_, r, a) in enumerate(zip(want["start"], want["referenceAllele"],
                                                                   want["alternateAllele"]))}
    assert {(g.start, g.variant.referenceAllele, g.variant.alternateAllele) for g in gts} == set(okey)
    assert len(gts) > 20
    for g in gts:
        j = okey[(g.start, g.variant.referenceAllele, g.variant.alternateAllele)]
        assert (g.readDepth, g.alternateReadDepth, g.genotypeQuality) == \
            (want["totalCoverage"][j], want["alleleCoverage"][j], want["genotypeQuality"][j])
        assert g.genotypeLikelihoods == list(want["genotypeLikelihoods"][j][:3])


def test_deletion_longer_than_65535_bases_keeps_its_allele(oracle, ctx):
    """Operators are 28-bit; a 70 kb deletion seen by enough reads is ONE candidate whose reference
    allele is 70 001 bases long (the indel list used to keep both lengths in 16 bits and
    such a site came out as ref_len 4465 with the wrong alleles)."""
    from avocado_b200.genotyping import BiallelicGenotyper, DiscoverVariants
    from avocado_b200.records import AlignmentRecord
    rng = np.random.default_rng(5)
    D = 70000
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, D + 400))
    recs = []
    for i in range(4):                       # > minObservations = 2
        s = 100 + 3 * i
        a = 140 - s
        seq = ref[s:s + a] + ref[s + a + D:s + a + D + 40]
        recs.append(AlignmentRecord(readMapped=True, contigName="1", start=s, end=s + a + D + 40, sequence=seq,
                                    qual="I" * (a + 40), cigar="%dM%dD40M" % (a, D),
                                    mismatchingPositions="%d^%s40" % (a, ref[s + a:s + a + D]), mapq=50, readName="d%d" % i))
    for i in range(4):                       # ten bases on, one base shorter: another candidate
        s = 90 + i
        a = 150 - s
        seq = ref[s:s + a] + ref[s + a + D - 1:s + a + D - 1 + 30]
        recs.append(AlignmentRecord(readMapped=True, contigName="1", start=s, end=s + a + D - 1 + 30, sequence=seq,
                                    qual="I" * (a + 30), cigar="%dM%dD30M" % (a, D - 1),
                                    mismatchingPositions="%d^%s30" % (a, ref[s + a:s + a + D - 1]), mapq=50,
                                    readName="e%d" % i))
    recs.sort(key=lambda r: r.start)
    rs = oracle.reads_from_dicts([r.as_dict() for r in recs])
    want_v = oracle.discover(rs, 20, 2)
    assert sorted(len(r) for _, _, _, r, _, _ in want_v) == [D, D + 1]
    got_v = DiscoverVariants(recs, 20, 2, ctx=ctx)
    assert sorted(v.as_tuple() for v in got_v) == sorted((c, s, e, r, a) for c, s, e, r, a, _ in want_v)
    want = oracle.call(rs, [(c, s, e, r, a) for c, s, e, r, a, _ in want_v], 2, [], False, 93, 93, 1)
    gts = BiallelicGenotyper.discoverAndCall(recs, optPhredThreshold=20, optMinObservations=2, samples=["s"], ctx=ctx)
    okey = {(int(s_), r, a): j for j, (s_, r, a) in enumerate(zip(want["start"], want["referenceAllele"],
                                                                   want["alternateAllele"]))}
    assert {(g.start, g.variant.referenceAllele, g.variant.alternateAllele) for g in gts} == set(okey)
    for g in gts:
        j = okey[(g.start, g.variant.referenceAllele, g.variant.alternateAllele)]
        assert (g.readDepth, g.alternateReadDepth, g.genotypeQuality) == \
            (want["totalCoverage"][j], want["alleleCoverage"][j], want["genotypeQuality"][j])
        assert g.alleles == ["ALT"] * int(want["nAlt"][j]) + ["REF"] * int(want["nRef"][j]) + ["OTHER_ALT"] * int(want["nOther"][j])


def test_bucket_capacity_guess_is_checked():
    """The call stage runs with the previous call's bucket count as capacity; a larger batch than the
    one before overflows that guess and the stage is rerun: rows equal a fresh context's."""
    from avocado_b200 import batch as B
    from avocado_b200.genotyping import Context
    small = B.synth_host(B.synth_params(seed=3, n_reads_total=2000, contig_len=8000))
    big = B.synth_host(B.synth_params(seed=4, n_reads_total=60000, contig_len=150000))
    a, b = Context(0), Context(0)
    try:
        a.discover_and_call_packed(small, phred=25, min_obs=5)
        got = [a.discover_and_call_packed(big, phred=25, min_obs=5) for _ in range(2)]   # rerun, then a good guess
        a.discover_and_call_packed(small, phred=25, min_obs=5)                           # guess far too large
        got.append(a.discover_and_call_packed(big, phred=25, min_obs=5))
        want_s, want_c = b.discover_and_call_packed(big, phred=25, min_obs=5)
    finally:
        a.close(); b.close()
    assert want_s.n > 100
    for s, c in got:
        assert np.array_equal(s.start, want_s.start)
        for k in want_c:
            assert np.array_equal(np.asarray(c[k]), np.asarray(want_c[k]), equal_nan=True), k


@pytest.mark.parametrize("seed,min_obs", [(1, 1), (2, 1), (3, 0)])
def test_messy_reads_against_the_oracle(oracle, ctx, seed, min_obs):
    """Short reads with every operator mix the packer accepts (insertions at read ends, deletions next
    to insertions, clips, X runs, N / IUPAC bases, mapq 0 and 255, reads whose strings are too short):
    dense, overlapping candidate sites; discovery, variant calls and gVCF rows equal the oracle's."""
    from avocado_b200 import batch as B
    from helpers import fuzz_records
    recs = [r for r in fuzz_records(100 + seed, 5000) if r.readMapped]
    batch = B.pack_reads(recs, sort=False)
    assert (batch.meta["n_ops"] == 0).any() and (batch.meta["flags"] & 2).any()
    rs = oracle_reads_from_batch(oracle, batch)
    want_v = oracle.discover(rs, 10, min_obs)
    sites = ctx.discover_packed(batch, phred=10, min_obs=min_obs)
    got = [(v.start, v.referenceAllele, v.alternateAllele) for v in sites.to_variants(["ctg"])]
    assert sorted(got) == sorted((s, r, a) for _, s, _, r, a, _ in want_v)
    assert len(got) > 50
    variants = [(c, s, e, r, a) for c, s, e, r, a, _ in want_v]
    want = oracle.call(rs, variants, 2, [], False, 93, 93, 1)
    cols = ctx.call_packed(batch, sites)
    keys = [("ctg", s, r, a) for s, r, a in got]
    st = compare_call(keys, cols, want)
    assert st["sites"] > 30
    want_all = oracle.call(rs, variants, 2, [], True, 93, 93, 1)
    c2, rstart, rcols = ctx.call_all_sites_packed(batch, sites)
    keys2 = keys + [("ctg", int(s), "N", None) for s in rstart]
    merged = {k: np.concatenate([np.asarray(c2[k]), np.asarray(rcols[k])]) for k in c2}
    compare_call(keys2, merged, want_all)
    # another ploidy, copy-number variants across the batch, tighter quality caps
    cnvs = [("ctg", 2000, 6000, 4), ("ctg", 9000, 9100, 1), ("ctg", 12000, 20000, 2)]
    want3 = oracle.call(rs, variants, 3, list(cnvs), True, 40, 50, 1)
    c3, rstart3, rcols3 = ctx.call_all_sites_packed(batch, sites, ploidy=3, cnvs=cnvs, contig_rank={"ctg": 0},
                                                    max_quality=40, max_mapq=50)
    keys3 = keys + [("ctg", int(s), "N", None) for s in rstart3]
    merged3 = {k: np.concatenate([np.asarray(c3[k]), np.asarray(rcols3[k])]) for k in c3}
    compare_call(keys3, merged3, want3)
