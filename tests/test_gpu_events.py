"""Discovery from candidate events (avo_derive_events + k_discover_events): a device batch that
carries its events -- every mismatching base, every I / D operator after the first aligned one, derived once
per batch by the packer -- must give the table, the counts and every result column the read-walking kernel
(k_discover_tiles) and the oracle give, bit for bit; and the events a GPU derives must be the events the host
loop derives."""
import numpy as np
import pytest

from helpers import compare_call, fuzz_records, oracle_reads_from_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from avocado_b200.genotyping import Context
    c = Context(0)
    yield c
    c.close()


def _tables_equal(a, b):
    assert a.n == b.n
    for k in ("start", "ref_len", "alt_len", "count", "allele_off"):
        assert np.array_equal(np.asarray(getattr(a, k)), np.asarray(getattr(b, k))), k
    assert np.array_equal(a.alleles4[:(a.n_allele_nibbles + 1) // 2], b.alleles4[:(b.n_allele_nibbles + 1) // 2])


def _events_equal(dev_batch, host_batch):
    d, h = dev_batch.events, host_batch.events
    assert (d["n_snv"], d["n_indel"], tuple(d["stats"])) == (h["n_snv"], h["n_indel"], tuple(h["stats"]))
    for k, n in (("snv_pos", h["n_snv"]), ("snv_pq", h["n_snv"]), ("indel_read", h["n_indel"]), ("indel_op", h["n_indel"]),
                 ("snv_first", len(h["snv_first"]))):
        got = d[k].cpu().numpy()[:n]
        want = h[k][:n]
        assert np.array_equal(got.view(want.dtype) if got.dtype != want.dtype else got, want), k


def _both_ways(ctx, host, phred, min_obs):
    """discoverAndCall on the device batch without and with events -> (table, cols) twice."""
    plain = host.to_device("cuda:0")
    t0, c0 = ctx.discover_and_call_packed(plain, phred=phred, min_obs=min_obs)
    with_ev = ctx.derive_events(host.to_device("cuda:0"))
    assert with_ev.events is not None
    t1, c1 = ctx.discover_and_call_packed(with_ev, phred=phred, min_obs=min_obs)
    _tables_equal(t0, t1)
    for k, v in c0.items():
        assert np.array_equal(np.asarray(v), np.asarray(c1[k]), equal_nan=True), k
    return with_ev, t1, c1


@pytest.mark.parametrize("phred,min_obs", [(25, 5), (0, None), (10, 2)])
def test_synthetic_batch_counts_its_events(ctx, phred, min_obs):
    from avocado_b200 import batch as B
    from avocado_b200.genotyping import derive_events
    p = B.synth_params(seed=41, n_reads_total=120000, contig_len=400000, first_pos=5000, indel_rate=1e-3, softclip_frac=0.1)
    host = B.synth_host(p)
    with_ev, t, _ = _both_ways(ctx, host, phred, min_obs)
    assert t.n > 300 and with_ev.events["n_indel"] > 100
    _events_equal(with_ev, derive_events(host))
    # avo_discover alone takes the same route
    _tables_equal(ctx.discover_packed(with_ev, phred=phred, min_obs=min_obs), t)


@pytest.mark.parametrize("seed", [0, 1])
def test_messy_reads_count_their_events(oracle, ctx, seed):
    """Every operator mix the packer accepts (insertions at read ends, deletions next to insertions, clips, X
    runs longer than the four head qualities, N / IUPAC bases -> the exact path): events == host loop, tables ==
    the walking kernel == the oracle."""
    from avocado_b200 import batch as B
    from avocado_b200.genotyping import derive_events
    recs = [r for r in fuzz_records(300 + seed, 5000) if r.readMapped]
    host = B.pack_reads(recs, sort=False)
    with_ev, t, cols = _both_ways(ctx, host, 10, None)
    _events_equal(with_ev, derive_events(host))
    rs = oracle_reads_from_batch(oracle, host)
    want_v = oracle.discover(rs, 10, None)
    got = [(v.start, v.referenceAllele, v.alternateAllele) for v in t.to_variants(["ctg"])]
    assert sorted(got) == sorted((s, r, a) for _, s, _, r, a, _ in want_v) and len(got) > 50
    want = oracle.call(rs, [(c, s, e, r, a) for c, s, e, r, a, _ in want_v], 2, [], False, 93, 93, 1)
    compare_call([("ctg", s, r, a) for s, r, a in got], cols, want)


def test_exotic_bases_and_deep_piles_take_the_exact_path(ctx):
    """Reads that disagree about a position's reference base, non-ACGT candidates, and a pile of more than
    32768 candidates in one tile: the exact histogram path of k_discover_events."""
    from avocado_b200 import batch as B
    from avocado_b200.records import AlignmentRecord

    def read(start, seq, md, name):
        return AlignmentRecord(readMapped=True, contigName="1", start=start, end=start + len(seq), sequence=seq,
                               qual="I" * len(seq), cigar="%dM" % len(seq), mismatchingPositions=md, mapq=40, readName=name)
    recs = []
    for i in range(6):
        recs.append(read(1000, "ACGTNACGCACGT", "4A3T4", "n%d" % i))
    for i in range(5):
        recs.append(read(1000, "ACGTNACGCACGT", "4C3G4", "m%d" % i))
    for i in range(7):
        recs.append(read(1002, "GTTACGTACG", "2A7", "p%d" % i))
    for i in range(4):
        recs.append(read(1003, "TTACGTAC", "3N4", "q%d" % i))
    _, t, _ = _both_ways(ctx, B.pack_reads(recs), 0, 3)
    assert t.n >= 4
    deep = [read(5000 + (i % 7), "ACGTACGTAC", "3T2A3", "d%d" % i) for i in range(20000)]      # 40000 candidates in one tile
    _, t2, _ = _both_ways(ctx, B.pack_reads(deep), 0, 100)
    assert t2.n >= 2 and int(np.asarray(t2.count).max()) > 2000


def test_unsorted_batch_is_refused_when_deriving(ctx):
    from avocado_b200 import _lib as L
    from avocado_b200 import batch as B
    p = B.synth_params(seed=3, n_reads_total=2000, contig_len=20000)
    host = B.synth_host(p)
    host.meta["start"][1000] -= 500
    with pytest.raises(L.AvocadoError) as e:
        ctx.derive_events(host.to_device("cuda:0"))
    assert e.value.status == L.AVO_E_UNSORTED


def test_sharded_ranks_count_their_events(ctx):
    """The region-sharded pipeline over batches that carry events (emulated ranks, rows stored directly)."""
    import torch
    from avocado_b200 import batch as B
    from avocado_b200.genotyping import Context
    from avocado_b200.parallel import ShardedPipeline, plan_region_shards
    p = B.synth_params(seed=36, n_reads_total=60000, contig_len=300000, first_pos=10000)
    whole = B.synth_host(p)
    t1, c1 = ctx.discover_and_call_packed(whole, phred=25, min_obs=5)
    world = 3
    shards = plan_region_shards(whole.meta["start"], whole.meta["end"], world, max_site_len=64)
    ctxs = [Context(0) for _ in range(world)]
    stream = torch.cuda.Stream("cuda:0")
    pipes, batches = [], []
    try:
        for sh in shards:
            pipes.append(ShardedPipeline(ctxs[sh.rank], sh.rank, world, max(sh.lo, -(1 << 31) + 1), min(sh.hi, (1 << 31) - 1),
                                         site_cap=1 << 12, allele_cap=1 << 14, gather=lambda out, inp: None, stream=stream,
                                         rows_to="direct", direct_root=(pipes[0] if pipes else True)))
            batches.append(ctxs[sh.rank].derive_events(whole.slice(sh.read_lo, sh.read_hi).to_device("cuda:0")))
        for _ in range(4):
            with torch.cuda.stream(stream):
                for q, b in zip(pipes, batches):
                    q.phase_discover(b, 25, 5)
                allt = torch.cat([q.t_send for q in pipes])
                for q in pipes:
                    q.t_all.copy_(allt)
                    q.phase_call()
                got = []
                for q in pipes:
                    q.rh_all.copy_(torch.cat([x.r_send[:32] for x in pipes]))
                    q._place_headers()
                    got.append(q.phase_finish())
            if got[0] is not None:
                break
        assert got[0] is not None
        torch.cuda.synchronize()
        table, cols = got[0]
        _tables_equal(table.to_host(), t1)
        for k, v in c1.items():
            assert np.array_equal(cols[k].cpu().numpy(), np.asarray(v), equal_nan=True), k
    finally:
        for q in reversed(pipes):
            q.close()
        for c in ctxs:
            c.close()
