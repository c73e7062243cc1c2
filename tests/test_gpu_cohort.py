"""Cohort pipeline parity at small size (the shape of BASELINE.json configs[2] / [3]: several
samples of one region, joint calling): per-sample DiscoverVariants, union of the sites, per-sample
BiallelicGenotyper.call against the union (the squared-off matrix), JointAnnotatorCaller -- every
stage on the GPU through the C ABI, against the oracle running the same stages."""
import numpy as np
import pytest

from helpers import REL_TOL, compare_call, oracle_reads_from_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from avocado_b200.genotyping import Context
    c = Context(0)
    yield c
    c.close()


def close(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    return bool(np.all(same | (np.abs(a - b) / np.maximum(np.abs(b), 1e-300) <= REL_TOL)))


@pytest.mark.parametrize("n_samples,coverage", [(3, 30), (8, 8)])
def test_cohort_discover_call_joint_matches_oracle(oracle, ctx, n_samples, coverage):
    from avocado_b200 import _lib as L, batch as B
    from avocado_b200.joint import joint_call
    from avocado_b200.records import Variant
    contig_len = 60000
    n_reads = contig_len * coverage // 150
    batches, sets, union = [], [], {}
    for s in range(n_samples):
        p = B.synth_params(seed=77, sample=s + 1, n_reads_total=n_reads, contig_len=contig_len, first_pos=5000)
        b = B.synth_host(p)
        rs = oracle_reads_from_batch(oracle, b)
        want_v = oracle.discover(rs, 25, 2)
        got = ctx.discover_packed(b, phred=25, min_obs=2)
        gv = [(v.start, v.end, v.referenceAllele, v.alternateAllele) for v in got.to_variants(["ctg"])]
        assert sorted(gv) == sorted((s_, e, r, a) for _, s_, e, r, a, _ in want_v)
        for v in gv:
            union[v] = Variant("ctg", *v)
        batches.append(b); sets.append(rs)
    # samples of one cohort share variant sites (the generator's truth set) but not genotypes
    table = B.sites_from_variants(list(union.values()), {"ctg": 0})
    variants = [list(union.values())[i] for i in table.order]
    keys = [("ctg", v.start, v.referenceAllele, v.alternateAllele) for v in variants]
    N, ns = table.n, 3
    assert N > 40
    cols_g = {k: np.zeros((n_samples, N), np.int32) for k in ("n_alt", "n_ref", "n_other")}
    gl_g = np.zeros((n_samples, N, ns))
    present = np.zeros((n_samples, N), np.uint8)
    want_rows = []
    for s in range(n_samples):
        cols = ctx.call_packed(batches[s], table)
        want = oracle.call(sets[s], [v.as_tuple() for v in variants], 2, [], False, 93, 93, 1)
        compare_call(keys, cols, want)                                   # per-sample rows: bit-identical
        present[s] = cols["emitted"]
        for k in cols_g:
            cols_g[k][s] = cols[k]
        gl_g[s] = cols["gl"]
        want_rows.append(want)
    # samples differ in their genotypes at the shared sites
    assert len({tuple(cols_g["n_alt"][:, i]) for i in range(N)}) > 3
    out = joint_call(ctx, N, n_samples, ns, cols_g["n_alt"].ravel(), cols_g["n_ref"].ravel(),
                     cols_g["n_other"].ravel(), gl_g.ravel(), present.ravel())
    n_joint = 0
    for i in range(N):
        idx = [s for s in range(n_samples) if present[s, i]]
        if not idx:
            continue
        gts = []
        for s in idx:
            al = [1] * int(cols_g["n_alt"][s, i]) + [0] * int(cols_g["n_ref"][s, i]) + [2] * int(cols_g["n_other"][s, i])
            gts.append((al, gl_g[s, i, :len(al) + 1].tolist()))
        site = oracle.annotate_site(gts)
        assert bool(out["site_emitted"][i]) == site["emitted"]
        if not site["emitted"]:
            continue
        assert close(out["maf"][i], site["maf"]) and close(out["quality"][i], site["quality"])
        for s, g in zip(idx, site["genotypes"]):
            fl = int(out["flags"][s, i])
            assert bool(fl & L.JOINT_RECALLED) == (g["genotypeQuality"] is not None)
            if g["posteriors"] is not None:
                assert close(out["posteriors"][s, i, :len(g["posteriors"])], g["posteriors"])
            if g["genotypeQuality"] is not None:
                assert int(out["gq"][s, i]) == g["genotypeQuality"]
                assert [1] * int(out["n_alt"][s, i]) + [0] * int(out["n_ref"][s, i]) == g["alleles"]
        n_joint += 1
    assert n_joint > 20


def test_trio_discovers_once_over_the_union_of_the_reads(oracle, ctx):
    """CLI/TrioGenotyper.scala:216-273 (BASELINE.json configs[2]): the three samples' reads are
    merged, DiscoverVariants runs ONCE over the union (so `count > minObservations` sees the
    candidates of all three samples), every sample is called against that table, and TrioCaller
    post-processes the sites.  Every stage on the GPU through the C ABI, against the oracle."""
    from avocado_b200 import batch as B
    from avocado_b200.trio import trio_call_packed
    contig_len, coverage = 80000, 10       # low enough that single samples miss sites the union finds
    n_reads = contig_len * coverage // 150
    batches = [B.synth_host(B.synth_params(seed=91, sample=s + 1, n_reads_total=n_reads, contig_len=contig_len,
                                           first_pos=3000)) for s in range(3)]
    union = B.merge_batches(batches)
    assert union.n_reads == 3 * n_reads and np.all(np.diff(union.meta["start"].astype(np.int64)) >= 0)
    rs_union = oracle_reads_from_batch(oracle, union)
    want_v = oracle.discover(rs_union, 25, 5)
    table = ctx.discover_packed(union, phred=25, min_obs=5)
    got = [(v.start, v.end, v.referenceAllele, v.alternateAllele) for v in table.to_variants(["ctg"])]
    assert sorted(zip(got, table.count.tolist())) == sorted(((s, e, r, a), n) for _, s, e, r, a, n in want_v)
    # the union finds sites no single sample would: per-sample discovery is a different candidate set
    per_sample = set()
    for b in batches:
        t = ctx.discover_packed(b, phred=25, min_obs=5)
        per_sample |= {(v.start, v.referenceAllele, v.alternateAllele) for v in t.to_variants(["ctg"])}
    assert {(s, r, a) for (s, e, r, a) in got} != per_sample
    # the device merge gives the same union batch
    import torch
    union_d = B.merge_batches([b.to_device("cuda:%d" % ctx.device) for b in batches])
    table_d = ctx.discover_packed(union_d, phred=25, min_obs=5)
    assert np.array_equal(table_d.start, table.start) and np.array_equal(table_d.count, table.count)
    keys = [("ctg", s, r, a) for (s, e, r, a) in got]
    variants = [("ctg", s, e, r, a) for (s, e, r, a) in got]
    N = table.n
    cnt = {k: np.zeros((3, N), np.int32) for k in ("n_alt", "n_ref", "n_other")}
    present = np.zeros((3, N), np.uint8)
    for s, b in enumerate(batches):
        cols = ctx.call_packed(b, table)
        want = oracle.call(oracle_reads_from_batch(oracle, b), variants, 2, [], False, 93, 93, 1)
        compare_call(keys, cols, want)                     # bit-identical per-sample rows
        present[s] = cols["emitted"]
        for k in cnt:
            cnt[k][s] = cols[k]
    out = trio_call_packed(ctx, N, cnt["n_alt"].ravel(), cnt["n_ref"].ravel(), cnt["n_other"].ravel(), present.ravel())
    kept = 0
    for i in range(N):
        gts = [([1] * int(cnt["n_alt"][s, i]) + [0] * int(cnt["n_ref"][s, i]) + [2] * int(cnt["n_other"][s, i]))
               if present[s, i] else None for s in range(3)]
        w = oracle.trio_process(*gts)
        assert bool(out["kept"][i]) == w["kept"]
        if w["kept"]:
            assert int(out["child_action"][i]) == w["childAction"]
            assert [bool(out["filled"][i] >> s & 1) for s in range(3)] == w["filled"]
            kept += 1
    assert kept > 20
