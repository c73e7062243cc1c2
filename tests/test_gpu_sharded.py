"""Region-sharded discoverAndCall (the C ABI's avo_shard_discover / avo_shard_call / avo_shard_finish, SURVEY.md
8e) on ONE GPU: several ranks are emulated by several contexts whose phases are interleaved by hand and
whose all-gathers are torch.cat -- exactly the bytes NCCL would deliver.  Every rank must end with the
table and the rows a single GPU computes for the whole batch, bit for bit.  (tests/multi_gpu_check.py runs
the same pipeline over NCCL on real ranks.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run_emulated(whole, world, phred, min_obs, site_cap, allele_cap, global_cap=None, max_site_len=64, rows_to="all"):
    import torch
    from avocado_b200.genotyping import Context
    from avocado_b200.parallel import ShardedPipeline, plan_region_shards
    shards = plan_region_shards(whole.meta["start"], whole.meta["end"], world, max_site_len=max_site_len)
    ctxs = [Context(0) for _ in range(world)]
    stream = torch.cuda.Stream("cuda:0")          # the emulated ranks share one stream (and one GPU)
    try:
        pipes, batches = [], []
        for sh in shards:
            lo = max(sh.lo, -(1 << 31) + 1)
            hi = min(sh.hi, (1 << 31) - 1)
            pipes.append(ShardedPipeline(ctxs[sh.rank], sh.rank, world, lo, hi, site_cap=site_cap, allele_cap=allele_cap,
                                         global_cap=global_cap, gather=lambda out, inp: None, stream=stream,
                                         rows_to="direct" if rows_to == "direct" else "all",
                                         direct_root=(pipes[0] if pipes else True) if rows_to == "direct" else None))
            batches.append(whole.slice(sh.read_lo, sh.read_hi).to_device("cuda:0"))
        tries = 0
        while True:
            tries += 1
            assert tries < 8
            with torch.cuda.stream(stream):
                for p, b in zip(pipes, batches):
                    p.phase_discover(b, phred, min_obs)
                allt = torch.cat([p.t_send for p in pipes])
                for p in pipes:
                    p.t_all.copy_(allt)
                    p.phase_call()
                allr = torch.cat([p.r_send for p in pipes])
                got = []
                for p in pipes:
                    if rows_to == "direct":         # only the 32-byte headers are gathered; the rows are already in place
                        p.rh_all.copy_(torch.cat([q.r_send[:32] for q in pipes]))
                        p._place_headers()
                    elif rows_to == "all" or p.plan.rank == 0:
                        p.r_all.copy_(allr)
                    else:                       # gather-to-root: the others see only the 32-byte headers
                        p.r_all.fill_(0xEE)
                        p.rh_all.copy_(torch.cat([q.r_send[:32] for q in pipes]))
                        p._place_own_rows()
                    got.append(p.phase_finish())
            assert len({g is None for g in got}) == 1, "every rank takes the same decision"
            if got[0] is not None:
                break
        out = []
        torch.cuda.synchronize()
        for table, cols in got:
            out.append((table.to_host(), {k: v.cpu().numpy() for k, v in cols.items()}))
        return out, tries, shards
    finally:
        if rows_to == "direct":
            for p in reversed(pipes):
                p.close()
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("world,site_cap,allele_cap", [(2, 1 << 12, 1 << 14), (3, 1 << 12, 1 << 14), (4, 64, 64), (8, 1 << 10, 1 << 12)])
def test_emulated_ranks_reproduce_the_single_gpu_rows(world, site_cap, allele_cap):
    from avocado_b200 import batch as B
    from avocado_b200.genotyping import Context
    p = B.synth_params(seed=33, n_reads_total=80000, contig_len=400000, first_pos=10000)
    whole = B.synth_host(p)
    ctx = Context(0)
    try:
        t1, c1 = ctx.discover_and_call_packed(whole, phred=25, min_obs=5)
    finally:
        ctx.close()
    assert t1.n > 300
    got, tries, shards = run_emulated(whole, world, 25, 5, site_cap, allele_cap)
    if site_cap == 64:
        assert tries > 1          # the blocks were too small: every rank asked for more and ran again
    # every rank owns something, and the halos hold reads of the neighbours
    assert all(sh.read_hi > sh.read_lo for sh in shards) and sum(sh.read_hi - sh.read_lo for sh in shards) > whole.n_reads
    for table, cols in got:
        assert table.n == t1.n
        for k in ("start", "ref_len", "alt_len", "count", "allele_off"):
            assert np.array_equal(np.asarray(getattr(table, k)), np.asarray(getattr(t1, k))), k
        assert np.array_equal(table.alleles4[:(table.n_allele_nibbles + 1) // 2], t1.alleles4[:(t1.n_allele_nibbles + 1) // 2])
        for k, v in c1.items():
            assert np.array_equal(cols[k], np.asarray(v), equal_nan=True), "column %s differs" % k


def test_rows_gathered_to_the_root_only():
    """rows_to="root": rank 0 ends with every row; the other ranks with the global table and their OWN rows
    (the blocks of the others reach them as headers with a row count of 0)."""
    from avocado_b200 import batch as B
    from avocado_b200.genotyping import Context
    p = B.synth_params(seed=34, n_reads_total=60000, contig_len=300000, first_pos=10000)
    whole = B.synth_host(p)
    ctx = Context(0)
    try:
        t1, c1 = ctx.discover_and_call_packed(whole, phred=25, min_obs=5)
    finally:
        ctx.close()
    got, tries, shards = run_emulated(whole, 4, 25, 5, 64, 64, rows_to="root")
    assert tries > 1
    start = np.asarray(t1.start)
    for sh, (table, cols) in zip(shards, got):
        assert table.n == t1.n and np.array_equal(np.asarray(table.start), start)
        own = np.ones(t1.n, bool) if sh.rank == 0 else (start >= sh.lo) & (start < sh.hi)
        assert own.sum() > 20
        for k, v in c1.items():
            v = np.asarray(v)
            per = v.size // t1.n
            m = np.repeat(own, per) if per > 1 else own
            assert np.array_equal(cols[k].ravel()[m], v.ravel()[m], equal_nan=True), "column %s differs on rank %d" % (k, sh.rank)


def test_rows_stored_directly_into_the_root_columns():
    """rows_to="direct" (avo_shard_call_direct): every rank's reduction writes its owned rows into rank 0's
    result columns; rank 0 ends with every row, bit for bit, with nothing gathered but headers."""
    from avocado_b200 import batch as B
    from avocado_b200.genotyping import Context
    p = B.synth_params(seed=35, n_reads_total=60000, contig_len=300000, first_pos=10000)
    whole = B.synth_host(p)
    ctx = Context(0)
    try:
        t1, c1 = ctx.discover_and_call_packed(whole, phred=25, min_obs=5)
    finally:
        ctx.close()
    for world, cap in ((4, 64), (3, 1 << 12)):
        got, tries, shards = run_emulated(whole, world, 25, 5, cap, cap, rows_to="direct")
        assert (tries > 1) == (cap == 64)
        for rank, (table, cols) in enumerate(got):
            assert table.n == t1.n and np.array_equal(np.asarray(table.start), np.asarray(t1.start))
            if rank:
                assert cols == {}
                continue
            for k, v in c1.items():
                assert np.array_equal(cols[k], np.asarray(v), equal_nan=True), "column %s differs" % k


def test_sharded_forest_quirk_across_the_boundary():
    """A long deletion just below a shard boundary reaches into the next rank's interval: that rank's reads
    must see it through the GLOBAL table's prefix-max end (Forest fast path vs literal replay), not through
    its own halo's incomplete discovery."""
    from avocado_b200 import batch as B
    from avocado_b200.genotyping import Context
    from avocado_b200.records import AlignmentRecord
    rng = np.random.default_rng(3)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, 6000))
    recs = []
    for i in range(1200):
        s = int(rng.integers(0, 5800))
        ln = 100
        seq = list(ref[s:s + ln])
        md, run = "", 0
        for k in range(ln):
            if (s + k) % 53 == 0:
                md += "%d%s" % (run, seq[k]); run = 0
                seq[k] = "ACGT"[("ACGT".index(seq[k]) + 1) % 4]
            else:
                run += 1
        md += str(run)
        recs.append(AlignmentRecord(readMapped=True, contigName="1", start=s, end=s + ln, sequence="".join(seq),
                                    qual="I" * ln, cigar="%dM" % ln, mismatchingPositions=md, mapq=50, readName="r%d" % i))
    for i in range(8):      # a 300-base deletion at 2900..3200 seen by eight reads that start before it
        s = 2860 + i
        a = 2900 - s
        seq = ref[s:s + a] + ref[s + a + 300:s + a + 300 + 60]
        recs.append(AlignmentRecord(readMapped=True, contigName="1", start=s, end=s + a + 300 + 60, sequence=seq,
                                    qual="I" * (a + 60), cigar="%dM300D60M" % a,
                                    mismatchingPositions="%d^%s60" % (a, ref[s + a:s + a + 300]), mapq=50, readName="d%d" % i))
    whole = B.pack_reads(recs)
    ctx = Context(0)
    try:
        t1, c1 = ctx.discover_and_call_packed(whole, phred=20, min_obs=3)
    finally:
        ctx.close()
    assert 300 + 1 in set(np.asarray(t1.ref_len).tolist())
    got, _, shards = run_emulated(whole, 4, 20, 3, 1 << 10, 1 << 14, max_site_len=512)
    for table, cols in got:
        assert table.n == t1.n and np.array_equal(np.asarray(table.start), np.asarray(t1.start))
        for k, v in c1.items():
            assert np.array_equal(cols[k], np.asarray(v), equal_nan=True), "column %s differs" % k
