"""GPU parity of avo_square_off against the oracle's squareOffSite, inside the documented cohort
workflow (docs/workflows/joint.rst): per-sample scoreAllSites calling -> extractVariants ->
squareOffSite -> JointAnnotatorCaller, every stage on the GPU."""
import numpy as np
import pytest

from helpers import oracle_reads_from_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from avocado_b200.genotyping import Context
    c = Context(0)
    yield c
    c.close()


def test_gvcf_cohort_square_off_and_joint(oracle, ctx):
    import torch
    from avocado_b200 import _lib as L, batch as B
    from avocado_b200.joint import joint_call
    from avocado_b200.records import Variant
    from avocado_b200.squareoff import extract_variants, square_off_packed
    S, contig_len, ns = 4, 40000, 3
    samples = []
    for s in range(S):
        p = B.synth_params(seed=91, sample=s + 1, n_reads_total=contig_len * 12 // 150, contig_len=contig_len, first_pos=3000)
        b = B.synth_host(p)
        sites = ctx.discover_packed(b, phred=25, min_obs=2)
        cols, rstart, rcols = ctx.call_all_sites_packed(b, sites)
        samples.append((sites, cols, rstart, rcols))
    # extractVariants over all samples' genotypes
    allv, called = [], []
    for sites, cols, _, _ in samples:
        vs = sites.to_variants(["ctg"])
        allv += vs
        called += [bool(cols["emitted"][i]) and int(cols["n_alt"][i]) > 0 for i in range(sites.n)]
    cohort_v = extract_variants(allv, called)
    assert len(cohort_v) > 30
    table = B.sites_from_variants(cohort_v, {"ctg": 0})
    table.contig = None
    cohort_sorted = [cohort_v[i] for i in table.order]
    N = table.n
    present = np.zeros((S, N), np.uint8)
    mat = {k: np.zeros((S, N), np.int32) for k in ("n_alt", "n_ref", "n_other")}
    gl = np.zeros((S, N, ns))
    kinds = set()
    for s, (sites, cols, rstart, rcols) in enumerate(samples):
        out = square_off_packed(ctx, table, sites, cols, rstart, rcols, ns)
        # the same tables resident on the device give the same bytes
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        dt = B.SiteTable(table.n, table.n_allele_nibbles, d(table.start), d(table.ref_len), d(table.alt_len),
                         d(table.allele_off.view(np.int32)), d(table.alleles4))
        ds = B.SiteTable(sites.n, sites.n_allele_nibbles, d(sites.start), d(sites.ref_len), d(sites.alt_len),
                         d(sites.allele_off.view(np.int32)), d(sites.alleles4))
        out_d = square_off_packed(ctx, dt, ds, {k: d(v) for k, v in cols.items()}, d(rstart),
                                  {k: d(v) for k, v in rcols.items()}, ns)
        for k in out:
            assert np.array_equal(np.asarray(out[k]), out_d[k].cpu().numpy(), equal_nan=True), k
        # oracle: squareOffSite for this sample at every cohort site, over the genotypes near the site
        sv = sites.to_variants(["ctg"])
        v_em = np.asarray(cols["emitted"]).astype(bool)
        r_em = np.asarray(rcols["emitted"]).astype(bool)
        for j, v in enumerate(cohort_sorted):
            gts, handle = [], []
            for k, x in enumerate(sv):
                if v_em[k] and x.start < v.end + 2 and x.end > v.start - 2:
                    gts.append(("ctg", x.start, x.end, x.referenceAllele, x.alternateAllele)); handle.append(("v", k))
            lo, hi = np.searchsorted(rstart, v.start - 2), np.searchsorted(rstart, v.end + 2)
            for k in range(lo, hi):
                if r_em[k]:
                    gts.append(("ctg", int(rstart[k]), int(rstart[k]) + 1, "N", None)); handle.append(("r", k))
            idx, excised = oracle.square_off_sample(("ctg", v.start, v.end, v.referenceAllele, v.alternateAllele), gts)
            kind = int(out["kind"][j])
            kinds.add(kind)
            if idx < 0:
                assert kind == L.SQ_ABSENT
                continue
            tab, k = handle[idx]
            want_kind = L.SQ_SCORED if not excised else (L.SQ_EXCISED_VARIANT if tab == "v" else L.SQ_EXCISED_REF_MODEL)
            assert (kind, int(out["src"][j])) == (want_kind, k), (s, j, v, gts)
            src = cols if tab == "v" else rcols
            want_gl = src["gl"][k] if not excised else src["nonref_gl"][k]
            assert np.array_equal(out["gl"][j], want_gl)
            assert (int(out["n_alt"][j]), int(out["n_ref"][j]), int(out["n_other"][j])) == \
                (int(src["n_alt"][k]), int(src["n_ref"][k]), int(src["n_other"][k]))
        present[s] = np.asarray(out["kind"]) != L.SQ_ABSENT
        for k in mat:
            mat[k][s] = out[k]
        gl[s] = out["gl"]
    assert {L.SQ_SCORED, L.SQ_EXCISED_REF_MODEL} <= kinds
    # the squared-off matrix feeds the joint caller
    jo = joint_call(ctx, N, S, ns, mat["n_alt"].ravel(), mat["n_ref"].ravel(), mat["n_other"].ravel(), gl.ravel(),
                    present.ravel())
    n_checked = 0
    for i in range(N):
        idx = [s for s in range(S) if present[s, i]]
        gts = [([1] * int(mat["n_alt"][s, i]) + [0] * int(mat["n_ref"][s, i]) + [2] * int(mat["n_other"][s, i]),
                gl[s, i, :int(mat["n_alt"][s, i] + mat["n_ref"][s, i] + mat["n_other"][s, i]) + 1].tolist()) for s in idx]
        site = oracle.annotate_site(gts)
        assert bool(jo["site_emitted"][i]) == site["emitted"]
        if site["emitted"]:
            assert abs(jo["quality"][i] - site["quality"]) <= 1e-6 * max(1.0, abs(site["quality"]))
            n_checked += 1
    assert n_checked > 20


def test_banded_reference_model_blocks(oracle, ctx):
    """Reference-model rows as BLOCKS (a banded gVCF, SquareOffReferenceModel.scala:224-232 takes any
    genotype whose region overlaps the site): the per-position rows of one sample are merged into
    blocks of 1-40 positions and squared off against another sample's variants."""
    import torch
    from avocado_b200 import _lib as L, batch as B
    from avocado_b200.squareoff import square_off_packed
    contig_len, ns = 30000, 3
    tabs = []
    for s in range(2):
        p = B.synth_params(seed=57, sample=s + 1, n_reads_total=contig_len * 10 // 150, contig_len=contig_len, first_pos=2000)
        b = B.synth_host(p)
        sites = ctx.discover_packed(b, phred=25, min_obs=2)
        tabs.append((sites,) + tuple(ctx.call_all_sites_packed(b, sites)))
    sites, cols, rstart, rcols = tabs[0]
    rstart = np.asarray(rstart)
    r_em = np.asarray(rcols["emitted"]).astype(bool)
    # blocks: runs of consecutive emitted positions cut at random lengths; the first row stands for the block
    rng = np.random.default_rng(5)
    first, length = [], []
    k = 0
    while k < len(rstart):
        if not r_em[k]:
            k += 1
            continue
        want, n = int(rng.integers(1, 41)), 1
        while n < want and k + n < len(rstart) and r_em[k + n] and rstart[k + n] == rstart[k] + n:
            n += 1
        first.append(k); length.append(n)
        k += n
    first, length = np.asarray(first), np.asarray(length, np.int32)
    assert length.max() > 20 and len(first) > 200
    bstart = rstart[first].astype(np.int32)
    bcols = {name: np.ascontiguousarray(np.asarray(a)[first]) for name, a in rcols.items()}
    # cohort: the other sample's variants (most of them shared) + made-up SNVs and deletions at random
    # positions, which only reference-model blocks (or nothing) can represent
    from avocado_b200.records import Variant
    cohort_v = {v.as_tuple(): v for v in tabs[1][0].to_variants(["ctg"])}
    for pos in rng.integers(2000, contig_len - 100, 200):
        n = int(rng.integers(0, 31)) if pos % 4 == 0 else 0
        v = Variant("ctg", int(pos), int(pos) + n + 1, "A" * (n + 1), "C" if n == 0 else "A")
        cohort_v.setdefault(v.as_tuple(), v)
    cohort_v = list(cohort_v.values())
    table = B.sites_from_variants(cohort_v, {"ctg": 0})
    table.contig = None
    cohort = [cohort_v[i] for i in table.order]
    out = square_off_packed(ctx, table, sites, cols, bstart, bcols, ns, ref_len=length)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    dt = B.SiteTable(table.n, table.n_allele_nibbles, d(table.start), d(table.ref_len), d(table.alt_len),
                     d(table.allele_off.view(np.int32)), d(table.alleles4))
    ds = B.SiteTable(sites.n, sites.n_allele_nibbles, d(sites.start), d(sites.ref_len), d(sites.alt_len),
                     d(sites.allele_off.view(np.int32)), d(sites.alleles4))
    out_d = square_off_packed(ctx, dt, ds, {k: d(v) for k, v in cols.items()}, d(bstart),
                              {k: d(v) for k, v in bcols.items()}, ns, ref_len=d(length), ref_max_len=int(length.max()))
    for k in out:
        assert np.array_equal(np.asarray(out[k]), out_d[k].cpu().numpy(), equal_nan=True), k
    sv = sites.to_variants(["ctg"])
    v_em = np.asarray(cols["emitted"]).astype(bool)
    bend = bstart + length
    kinds, reached_from_before = set(), 0
    for j, v in enumerate(cohort):
        gts, handle = [], []
        for k, x in enumerate(sv):
            if v_em[k] and x.start < v.end + 2 and x.end > v.start - 2:
                gts.append(("ctg", x.start, x.end, x.referenceAllele, x.alternateAllele)); handle.append(("v", k))
        for k in np.nonzero((bstart < v.end + 2) & (bend > v.start - 2))[0]:
            gts.append(("ctg", int(bstart[k]), int(bend[k]), "N", None)); handle.append(("r", int(k)))
        idx, excised = oracle.square_off_sample(("ctg", v.start, v.end, v.referenceAllele, v.alternateAllele), gts)
        kind = int(out["kind"][j])
        kinds.add(kind)
        if idx < 0:
            assert kind == L.SQ_ABSENT
            continue
        tab, k = handle[idx]
        want_kind = L.SQ_SCORED if not excised else (L.SQ_EXCISED_VARIANT if tab == "v" else L.SQ_EXCISED_REF_MODEL)
        assert (kind, int(out["src"][j])) == (want_kind, k), (j, v, gts)
        src = cols if tab == "v" else bcols
        assert np.array_equal(out["gl"][j], src["gl"][k] if not excised else src["nonref_gl"][k])
        assert (int(out["n_alt"][j]), int(out["n_ref"][j]), int(out["n_other"][j])) == \
            (int(src["n_alt"][k]), int(src["n_ref"][k]), int(src["n_other"][k]))
        reached_from_before += tab == "r" and bstart[k] < v.start
    assert {L.SQ_ABSENT, L.SQ_SCORED, L.SQ_EXCISED_REF_MODEL} <= kinds and reached_from_before > 20
    # a block length below 1 is refused
    bad = length.copy(); bad[3] = 0
    with pytest.raises(Exception):
        square_off_packed(ctx, table, sites, cols, bstart, bcols, ns, ref_len=bad)


def test_reference_gvcf_fixture_squares_off_like_the_oracle(oracle, ctx):
    """The reference's banded gVCF (gvcf_multiallelic.g.vcf, SquareOffReferenceModelSuite.scala:51-76):
    loadGenotypes -> extractVariants -> avo_square_off with reference-model blocks, per contig, against
    the oracle's squareOffSite; plus sites made up inside and between its blocks."""
    import os
    from avocado_b200 import _lib as L, batch as B, vcf as V
    from avocado_b200.records import ALT, Variant
    from avocado_b200.squareoff import extract_variants, square_off_packed
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vcf", "gvcf_multiallelic.g.vcf.gz")
    gts, header = V.load_genotypes(path)
    found = extract_variants([g.variant for g in gts], [ALT in g.alleles for g in gts])
    assert len(found) == 3
    kinds = set()
    for contig in sorted({g.contigName for g in gts}):
        mine = [g for g in gts if g.contigName == contig]
        table, vrows, vcols, rstart, rlen, rrows, rcols = V.sample_tables(mine, "NA12878i")
        sites = {v.as_tuple(): v for v in found if v.contigName == contig}
        for g in mine:                                   # made-up SNVs at block starts, ends and just past them
            for pos in (g.start, g.end - 1, g.end):
                v = Variant(contig, pos, pos + 1, "A", "C")
                sites.setdefault(v.as_tuple(), v)
        cohort_v = list(sites.values())
        cohort = B.sites_from_variants(cohort_v)
        cohort.contig = None
        out = square_off_packed(ctx, cohort, table, vcols, rstart, rcols, 3, ref_len=rlen)
        all_rows = [("v", k, g) for k, g in enumerate(vrows)] + [("r", k, g) for k, g in enumerate(rrows)]
        olist = [(contig, g.start, g.end, g.variant.referenceAllele, g.variant.alternateAllele) for _, _, g in all_rows]
        for j, i in enumerate(cohort.order):
            v = cohort_v[i]
            idx, excised = oracle.square_off_sample(v.as_tuple(), olist)
            kind = int(out["kind"][j])
            kinds.add(kind)
            if idx < 0:
                assert kind == L.SQ_ABSENT, v
                continue
            tab, k, g = all_rows[idx]
            want = L.SQ_SCORED if not excised else (L.SQ_EXCISED_VARIANT if tab == "v" else L.SQ_EXCISED_REF_MODEL)
            assert (kind, int(out["src"][j])) == (want, k), (v, olist[idx])
            src = vcols if tab == "v" else rcols
            assert np.array_equal(out["gl"][j], src["gl"][k] if not excised else src["nonref_gl"][k])
    assert {L.SQ_ABSENT, L.SQ_EXCISED_VARIANT, L.SQ_EXCISED_REF_MODEL} <= kinds
