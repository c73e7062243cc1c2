"""GPU parity tests: the CUDA path, through the C ABI, against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): bit-exact observation counts, allele assignments and called
genotypes; floating-point likelihoods and qualities within 1e-6 relative (helpers.REL_TOL).  The
product additionally sums in the oracle's canonical read order, so the tests also require the
floating-point columns to be bit-identical.
"""
import numpy as np
import pytest

from helpers import (compare_call, load_fixture, oracle_call, oracle_discover, oracle_reads_from_batch)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from avocado_b200.genotyping import Context
    c = Context(0)
    yield c
    c.close()


def mapq_gt(recs, thr):
    return [r for r in recs if r.mapq is not None and r.mapq > thr]


E2E = [  # fixture, mapq filter, phred, min_obs   (TEST/genotyping/BiallelicGenotyperSuite.scala:383-911)
    ("NA12878_snp_A2G_chr20_225058", 0, None, None),
    ("NA12878.chr1.832736", 0, None, None),
    ("NA12878.chr1.839395", 0, None, None),
    ("NA12878.chr1.567239", 0, None, None),
    ("NA12878.chr1.875159", 0, None, None),
    ("NA12878.1_1777263", 0, None, 3),
    ("NA12878.1_1067596", None, None, None),
    ("NA12878.chr1.877715", 0, None, None),
    ("NA12878.chr1.886049", 0, None, None),
    ("NA12878.chr1.889159", 0, 30, None),
    ("NA12878.chr1.866511", 0, 30, None),
    ("NA12878.chr1.905130", 0, 30, None),
    ("NA12878.chr1.905130", 0, 30, 4),
    ("NA12878.chr1.907170", 0, 30, None),
    ("NA12878.chr1.240898", 10, 25, None),
    ("NA12878.1_4120185", None, 18, 3),
    ("NA12878.1_5274547", None, 18, 3),
    ("NA12878.chr1.104160", 0, None, None),
    ("NA12878.chr1.875635", 0, 25, 5),
]


@pytest.mark.parametrize("name,mq,phred,min_obs", E2E)
def test_fixture_discover_and_call_matches_oracle(oracle, ctx, name, mq, phred, min_obs):
    from avocado_b200.genotyping import BiallelicGenotyper, DiscoverVariants
    recs = load_fixture(name)[0]
    if mq is not None:
        recs = mapq_gt(recs, mq)
    # the product's canonical summation order is the batch order = reads stably sorted by start;
    # hand the oracle the same order (the reference's own order is Spark's partition order)
    recs = sorted([r for r in recs if r.readMapped], key=lambda r: (r.contigName, r.start))
    want_v = oracle_discover(oracle, recs, phred, min_obs)
    got_v = DiscoverVariants(recs, phred, min_obs, ctx=ctx)
    assert sorted(v.as_tuple() for v in got_v) == sorted(want_v)
    if not want_v:
        return
    want = oracle_call(oracle, recs, want_v)
    gts = BiallelicGenotyper.discoverAndCall(recs, optPhredThreshold=phred, optMinObservations=min_obs,
                                             samples=["s"], ctx=ctx)
    # rebuild columns from the record-level API to compare with the oracle
    keys = [(g.contigName, g.start, g.variant.referenceAllele, g.variant.alternateAllele) for g in gts]
    okeys = list(zip(want["contigName"], [int(s) for s in want["start"]], want["referenceAllele"],
                     want["alternateAllele"]))
    assert sorted(keys) == sorted(okeys)
    oi = {k: j for j, k in enumerate(okeys)}
    for g in gts:
        j = oi[(g.contigName, g.start, g.variant.referenceAllele, g.variant.alternateAllele)]
        assert g.alleles == ["ALT"] * int(want["nAlt"][j]) + ["REF"] * int(want["nRef"][j]) + \
            ["OTHER_ALT"] * int(want["nOther"][j])
        assert g.genotypeQuality == want["genotypeQuality"][j]
        assert g.readDepth == want["totalCoverage"][j]
        assert g.referenceReadDepth == want["otherCoverage"][j]
        assert g.alternateReadDepth == want["alleleCoverage"][j]
        assert g.strandBiasComponents == list(want["strandBiasComponents"][j])
        cn = int(want["copyNumber"][j])
        assert g.genotypeLikelihoods == list(want["genotypeLikelihoods"][j][:cn + 1])
        assert g.nonReferenceLikelihoods == list(want["nonReferenceLikelihoods"][j][:cn + 1])
        assert np.float32(g.rmsMapQ) == want["rmsMapQ"][j] or (np.isnan(g.rmsMapQ) and np.isnan(want["rmsMapQ"][j]))
        assert np.float32(g.fisherStrandBiasPValue) == want["fisherStrandBiasPValue"][j]


def test_reference_e2e_expectations_on_gpu(ctx):
    """A few of the reference's own assertions evaluated directly on the GPU output
    (BiallelicGenotyperSuite.scala:383-410, 600-619, 828-863, 890-911)."""
    from avocado_b200.genotyping import BiallelicGenotyper
    from avocado_b200.records import AlignmentRecord
    recs = mapq_gt(load_fixture("NA12878_snp_A2G_chr20_225058")[0], 0)
    gts = [g for g in BiallelicGenotyper.discoverAndCall(recs, samples=["s"], ctx=ctx)
           if g.readDepth > 10 and not all(a == "REF" for a in g.alleles) and g.genotypeQuality > 30]
    assert len(gts) == 1
    g = gts[0]
    assert (g.start, g.end, g.variant.referenceAllele, g.variant.alternateAllele) == (225057, 225058, "A", "G")
    assert sorted(g.alleles) == ["ALT", "REF"]

    recs = mapq_gt(load_fixture("NA12878.1_1777263")[0], 0)
    gts = BiallelicGenotyper.discoverAndCall(recs, optMinObservations=3, samples=["s"], ctx=ctx)
    assert len(gts) == 1 and gts[0].start == 1777262 and gts[0].alleles == ["ALT", "ALT"]
    assert gts[0].variant.referenceAllele == "TACACACACACACACACACACACACACACAC"

    def make_read(allele):
        return AlignmentRecord(contigName="ctg", start=10, end=15, sequence="AC%sTG" % allele, cigar="5M",
                               mismatchingPositions="2T2", qual="8383838383", mapq=50, readMapped=True)
    gts = BiallelicGenotyper.discoverAndCall([make_read("A")] * 4 + [make_read("C")] * 4, samples=["s"], ctx=ctx)
    assert len(gts) == 2 and {g.variant.alternateAllele for g in gts} == {"A", "C"}
    assert all(sorted(g.alleles) == ["ALT", "OTHER_ALT"] and g.genotypeQuality == 67 for g in gts)

    recs = load_fixture("NA12878.1_5274547")[0]
    gts = BiallelicGenotyper.discoverAndCall(recs, optPhredThreshold=18, optMinObservations=3,
                                             samples=["s"], ctx=ctx)
    assert len(gts) == 2 and all(g.start == 5274546 and g.variant.alternateAllele == "T" for g in gts)
    assert sorted(g.variant.referenceAllele for g in gts) == ["TTA", "TTATA"]
    assert all(sorted(g.alleles) == ["ALT", "OTHER_ALT"] for g in gts)


@pytest.mark.parametrize("seed,n_reads,contig_len,phred,min_obs", [
    (1, 6000, 30000, 25, 5), (2, 20000, 100000, 25, 5), (3, 3000, 10000, 0, None),
    (4, 12000, 40000, 30, 2), (5, 40000, 100000, 25, 5)])
def test_synthetic_batch_matches_oracle(oracle, ctx, seed, n_reads, contig_len, phred, min_obs):
    from avocado_b200 import batch as B
    p = B.synth_params(seed=seed, n_reads_total=n_reads, contig_len=contig_len, first_pos=1000 * seed)
    host = B.synth_host(p)
    dev = B.synth_device(p)
    dh = dev.to_host()
    used = {"payload": host.n_blocks * 16, "ops": host.n_ops, "refb": host.n_refb}
    for name, _ in B.READ_COLUMNS:      # host and device generators agree byte for byte
        n = used.get(name, len(getattr(host, name)))
        assert np.array_equal(getattr(host, name)[:n], getattr(dh, name)[:n]), name
    rs = oracle_reads_from_batch(oracle, host)
    want_v = oracle.discover(rs, phred, min_obs)
    # discovery: host-resident input and device-resident input give the same table
    sites = ctx.discover_packed(host, phred=phred, min_obs=min_obs)
    sites_d = ctx.discover_packed(dev, phred=phred, min_obs=min_obs)
    got = [(v.start, v.end, v.referenceAllele, v.alternateAllele) for v in sites.to_variants(["ctg"])]
    assert got == [(v.start, v.end, v.referenceAllele, v.alternateAllele) for v in sites_d.to_variants(["ctg"])]
    assert sorted(got) == sorted((s, e, r, a) for _, s, e, r, a, _ in want_v)
    assert sorted(zip(got, sites.count.tolist())) == sorted(((s, e, r, a), n) for _, s, e, r, a, n in want_v)
    # call against those sites
    want = oracle.call(rs, [(c, s, e, r, a) for c, s, e, r, a, _ in want_v], 2, [], False, 93, 93, 1)
    cols = ctx.call_packed(dev, sites)
    keys = [("ctg", s, r, a) for (s, e, r, a) in got]
    st = compare_call(keys, cols, want)
    assert st["sites"] > 0
    # fused entry point gives the same rows (with and without the packer's bounds)
    host.stats = None if seed % 2 else host.stats
    sites2, cols2 = ctx.discover_and_call_packed(host, phred=phred, min_obs=min_obs)
    assert np.array_equal(sites2.start, sites.start)
    for k in cols:
        assert np.array_equal(np.asarray(cols[k]), np.asarray(cols2[k]), equal_nan=True), k


def test_baseline_configs0_at_its_own_size(oracle, ctx):
    """BASELINE.json configs[0] in full -- single-sample BiallelicGenotyper over a synthetic 1 Mb
    region at 30x, 150 bp reads = 2.0e5 reads -- against the single-thread oracle: the discovered
    table incl. counts, and every result column bit-identical."""
    from avocado_b200 import batch as B
    n_reads, contig_len = 200_000, 1_000_000
    p = B.synth_params(seed=0xA70CAD0, n_reads_total=n_reads, contig_len=contig_len)
    dev = B.synth_device(p)
    host = dev.to_host()
    rs = oracle_reads_from_batch(oracle, host)
    want_v = oracle.discover(rs, 25, 5)
    assert len(want_v) > 800                       # about 1.1 sites / kb
    sites, cols = ctx.discover_and_call_packed(dev, ploidy=2, phred=25, min_obs=5, capacity=1 << 14)
    got = [(v.start, v.end, v.referenceAllele, v.alternateAllele) for v in sites.to_variants(["ctg"])]
    assert sorted(zip(got, sites.count.tolist())) == sorted(((s, e, r, a), n) for _, s, e, r, a, n in want_v)
    want = oracle.call(rs, [(c, s, e, r, a) for c, s, e, r, a, _ in want_v], 2, [], False, 93, 93, 1)
    st = compare_call([("ctg", s, r, a) for (s, e, r, a) in got], cols, want, exact_fp=True)
    assert st["sites"] > 800 and st["fp_bit_identical"]
    # the same batch from pinned-style host columns (compact wire form when it fits) gives the same rows
    host.with_wire6() or host.with_wire()
    sites_h, cols_h = ctx.discover_and_call_packed(host, ploidy=2, phred=25, min_obs=5, capacity=1 << 14)
    assert np.array_equal(sites_h.start, sites.start) and np.array_equal(sites_h.count, sites.count)
    for k in cols:
        assert np.array_equal(np.asarray(cols[k]), np.asarray(cols_h[k]), equal_nan=True), k


def test_empty_sites_and_unsorted_reads_are_errors(ctx):
    from avocado_b200 import _lib as L, batch as B
    from avocado_b200.genotyping import BiallelicGenotyper
    from avocado_b200.records import AlignmentRecord
    r = AlignmentRecord(readMapped=True, contigName="1", start=10, end=18, sequence="ACACATGA",
                        qual="!!!!!!!!", cigar="8M", mismatchingPositions="8", mapq=30)
    with pytest.raises(L.AvocadoError) as e:       # TreeRegionJoin.scala:77 throws on an empty forest
        BiallelicGenotyper.call([r], [], samples=["s"], ctx=ctx)
    assert e.value.status == L.AVO_E_EMPTY_SITES
    p = B.synth_params(n_reads_total=1000, contig_len=5000)
    b = B.synth_host(p)
    good = b.stats
    b.stats = (good[0], good[1] - 1, good[2])       # the caller's bounds are checked on the device
    with pytest.raises(L.AvocadoError) as e:
        ctx.discover_packed(b)
    assert e.value.status == L.AVO_E_INVALID
    b.stats = (good[0], good[1], good[2] - 1)
    with pytest.raises(L.AvocadoError) as e:
        ctx.discover_packed(b)
    assert e.value.status == L.AVO_E_INVALID
    b.stats = (good[0] - 100, good[1] + 100, good[2] + 100)   # loose bounds are fine
    want = ctx.discover_packed(b)
    b.stats = None                                   # no bounds: the library computes them
    got = ctx.discover_packed(b)
    assert np.array_equal(want.start, got.start) and np.array_equal(want.count, got.count)
    b.meta["start"] = np.ascontiguousarray(b.meta["start"][::-1])
    for stats in (None, good):
        b.stats = stats
        with pytest.raises(L.AvocadoError) as e:
            ctx.discover_packed(b)
        assert e.value.status == L.AVO_E_UNSORTED
    b.meta["start"] = np.ascontiguousarray(b.meta["start"][::-1])
    b.meta["start"][500], b.meta["start"][501] = b.meta["start"][501] + 1, b.meta["start"][500]
    b.stats = good                                   # one swapped pair in the middle
    with pytest.raises(L.AvocadoError) as e:
        ctx.discover_packed(b)
    assert e.value.status == L.AVO_E_UNSORTED
    with pytest.raises(ValueError):                 # BG:102-105 exactly one sample
        BiallelicGenotyper.call([r], [], samples=["a", "b"], ctx=ctx)


def test_discovery_exact_path_for_exotic_bases(oracle, ctx):
    """Candidates with non-ACGT bases, or reads that disagree about the reference base of a
    position, take the exact (ref, alt) histogram path of k_discover_tiles; the result must still
    be the reference's groupBy(contig, start, ref, alt).count (DiscoverVariants.scala:87-97)."""
    from avocado_b200.genotyping import DiscoverVariants
    from avocado_b200.records import AlignmentRecord

    def read(start, seq, md, name):
        return AlignmentRecord(readMapped=True, contigName="1", start=start, end=start + len(seq),
                               sequence=seq, qual="I" * len(seq), cigar="%dM" % len(seq),
                               mismatchingPositions=md, mapq=40, readName=name)
    recs = []
    for i in range(6):      # alt N over ref A at 1004; alt C over ref A at 1008
        recs.append(read(1000, "ACGTNACGCACGT", "4A3T4", "n%d" % i))
    for i in range(5):      # the same positions, but these reads claim other reference bases
        recs.append(read(1000, "ACGTNACGCACGT", "4C3G4", "m%d" % i))
    for i in range(7):      # plain ACGT candidates in the same tile
        recs.append(read(1002, "GTTACGTACG", "2A7", "p%d" % i))
    for i in range(4):      # reference base N
        recs.append(read(1003, "TTACGTAC", "3N4", "q%d" % i))
    for phred, min_obs in ((None, None), (0, 3), (10, 5)):
        want = oracle_discover(oracle, recs, phred, min_obs)
        got = DiscoverVariants(recs, phred, min_obs, ctx=ctx)
        assert sorted(v.as_tuple() for v in got) == sorted(want), (phred, min_obs)
    assert len(oracle_discover(oracle, recs, 0, 3)) >= 4


def test_pinned_host_batch_is_read_in_place(ctx):
    """A page-locked host batch keeps the payload on the host (the kernels gather single
    sectors over PCIe); results are identical to the copied path and to device-resident input."""
    import torch
    from avocado_b200 import batch as B
    p = B.synth_params(seed=9, n_reads_total=30000, contig_len=80000)
    host = B.synth_host(p)
    pinned = B.ReadBatch(host.n_reads, host.n_blocks, host.n_ops, host.n_refb, host.contig_id, stats=host.stats)
    keep = []
    for name, dt in B.READ_COLUMNS:
        a = getattr(host, name)
        t = torch.from_numpy(a.view(np.uint8) if a.dtype == B.META_DTYPE else B._torch_view(a)).pin_memory()
        keep.append(t)
        v = t.numpy()
        setattr(pinned, name, v.view(dt) if v.dtype != np.dtype(dt) else v)
    outs = []
    for batch, payload, in_place in ((pinned, 0, 1), (pinned, 1, 0), (host, 0, 0), (pinned, 2, 2), (host, 2, 0)):
        ctx.host_payload = payload
        try:
            sites, cols = ctx.discover_and_call_packed(batch, phred=25, min_obs=5)
        finally:
            ctx.host_payload = 0
        t = ctx.timing()
        assert t["payload_in_place"] == in_place
        assert (t["gathered_blocks"] > 1000) == (in_place == 2)
        outs.append((sites, cols))
    for sites, cols in outs[1:]:
        assert np.array_equal(sites.start, outs[0][0].start)
        for k in cols:
            assert np.array_equal(np.asarray(cols[k]), np.asarray(outs[0][1][k]), equal_nan=True), k
    assert outs[0][0].n > 10


def test_two_gpu_region_and_sample_sharding():
    """Region-sharded calling and sample-sharded joint calling over NCCL (tests/multi_gpu_check.py);
    needs two GPUs on the box."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(root, "tests", "multi_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "multi_gpu_check ok" in r.stdout


@pytest.mark.parametrize("form", ["wire16", "wire6"])
def test_wire_form_gives_the_same_results(ctx, form):
    """Host batches in the compact wire forms (16-byte records, or 6-byte records with start
    differences and sizes implied by the operators; 16-bit operators; offsets rebuilt on the device)
    produce exactly the rows of the full form, for discovery, call and gVCF rows."""
    from avocado_b200 import batch as B
    p = B.synth_params(seed=12, n_reads_total=25000, contig_len=70000, first_pos=123, softclip_frac=0.1)
    full = B.synth_host(p)
    wire = B.synth_host(p)
    assert wire.with_wire() if form == "wire16" else wire.with_wire6()
    wire.meta = np.zeros(1, B.META_DTYPE)[:0]      # prove that only the wire columns are used
    wire.ops = np.zeros(1, np.uint32)
    # (as_struct passes whatever is there; the library must not touch meta / ops)
    s_full, c_full = ctx.discover_and_call_packed(full, phred=25, min_obs=5)
    s_wire, c_wire = ctx.discover_and_call_packed(wire, phred=25, min_obs=5)
    assert s_full.n == s_wire.n > 20 and np.array_equal(s_full.start, s_wire.start)
    assert np.array_equal(s_full.alleles4, s_wire.alleles4) and np.array_equal(s_full.count, s_wire.count)
    for k in c_full:
        assert np.array_equal(np.asarray(c_full[k]), np.asarray(c_wire[k]), equal_nan=True), k
    a = ctx.call_all_sites_packed(full, s_full)
    b = ctx.call_all_sites_packed(wire, s_full)
    assert np.array_equal(a[1], b[1])
    for k in a[2]:
        assert np.array_equal(np.asarray(a[2][k]), np.asarray(b[2][k]), equal_nan=True), k


def test_wire6_exceptions_reach_the_device(ctx):
    """Reads whose end / l_seq do not follow from their operators (failed extractions, strings longer
    than the CIGAR) travel as exceptions of the 6-byte form; discovery sees the same batch."""
    from avocado_b200 import batch as B
    from helpers import fuzz_records, short_operators
    recs = short_operators([r for r in fuzz_records(9, 30000) if r.readMapped])
    full = B.pack_reads(recs, sort=False)
    wire = B.pack_reads(recs, sort=False)
    assert wire.with_wire6(max_exceptions=wire.n_reads) and len(wire.wire6_exc) > 100
    wire.meta = np.zeros(1, B.META_DTYPE)[:0]
    wire.ops = np.zeros(1, np.uint32)
    s_full = ctx.discover_packed(full, phred=10, min_obs=1)
    s_wire = ctx.discover_packed(wire, phred=10, min_obs=1)
    assert s_full.n == s_wire.n > 50
    for k in ("start", "ref_len", "alt_len", "alleles4", "count"):
        assert np.array_equal(getattr(s_full, k), getattr(s_wire, k)), k
    c_full = ctx.call_packed(full, s_full)
    c_wire = ctx.call_packed(wire, s_full)
    for k in c_full:
        assert np.array_equal(np.asarray(c_full[k]), np.asarray(c_wire[k]), equal_nan=True), k


def _pinned_copy(B, host):
    import torch
    pinned = B.ReadBatch(host.n_reads, host.n_blocks, host.n_ops, host.n_refb, host.contig_id, stats=host.stats)
    keep = []
    for name, dt in B.READ_COLUMNS:
        a = getattr(host, name)
        t = torch.from_numpy(a.view(np.uint8) if a.dtype == B.META_DTYPE else B._torch_view(a)).pin_memory()
        keep.append(t)
        v = t.numpy()
        setattr(pinned, name, v.view(dt) if v.dtype != np.dtype(dt) else v)
    pinned._keep = keep
    return pinned


@pytest.mark.parametrize("threads", [1, 3])
def test_host_gathered_payload_gives_the_same_rows(ctx, threads):
    """host_payload = 2: the call stage works on payload blocks gathered by host threads (the device
    lists them); rows are bit-identical to the in-place path on an indel- and clip-rich batch, for
    avo_call against a foreign site table too, and on the reference's indel fixtures."""
    from avocado_b200 import batch as B
    from helpers import load_fixture
    p = B.synth_params(seed=21, n_reads_total=40000, contig_len=60000, indel_rate=2e-3, snp_rate=3e-3,
                       softclip_frac=0.3, triallelic_frac=0.05)
    batches = [_pinned_copy(B, B.synth_host(p))]
    for name in ("NA12878.chr1.832736", "NA12878.chr1.875159", "NA12878.1_1777263", "NA12878.chr1.905130"):
        try:
            recs = [r for r in load_fixture(name)[0] if r.readMapped]
        except FileNotFoundError:
            continue
        batches.append(_pinned_copy(B, B.pack_reads(recs)))
    from helpers import fuzz_records
    batches.append(_pinned_copy(B, B.pack_reads([r for r in fuzz_records(55, 6000) if r.readMapped], sort=False)))
    assert len(batches) >= 3
    for batch in batches:
        rows = []
        for mode in (0, 2):
            ctx.host_payload, ctx.gather_threads = mode, threads
            try:
                sites = ctx.discover_packed(batch, phred=10, min_obs=1)
                cols = ctx.call_packed(batch, sites)
                assert ctx.timing()["payload_in_place"] == (2 if mode else 1)
            finally:
                ctx.host_payload, ctx.gather_threads = 0, 0
            rows.append((sites, cols))
        assert rows[0][0].n == rows[1][0].n > 0
        for k in rows[0][1]:
            assert np.array_equal(np.asarray(rows[0][1][k]), np.asarray(rows[1][1][k]), equal_nan=True), k


def test_c_example_runs_through_the_c_abi(tmp_path):
    """examples/discover_and_call.c: the path from plain C, linked against the shared library."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "dac")
    lib = os.path.join(root, "avocado_b200")
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-I", os.path.join(root, "include"),
                           os.path.join(root, "examples", "discover_and_call.c"), "-o", exe, "-L", lib,
                           "-lavocado_b200", "-Wl,-rpath," + lib])
    r = subprocess.run([exe, "30000"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "30000 reads ->" in r.stdout and "GQ" in r.stdout


def test_called_genotypes_survive_a_vcf_round_trip(ctx, tmp_path):
    """discoverAndCall -> RewriteHets + HardFilterGenotypes -> saveAsVcf text -> loadGenotypes: every field
    the VCF carries comes back as it was written (float32 text is Java's shortest form, PL is rounded phred)."""
    from avocado_b200 import vcf as V
    from avocado_b200.filters import FilterArgs, rewrite_and_filter
    from avocado_b200.genotyping import BiallelicGenotyper
    recs = load_fixture("NA12878.1_5274547")[0]          # two overlapping deletions: ALT + OTHER_ALT at both sites
    gts = BiallelicGenotyper.discoverAndCall(recs, optPhredThreshold=18, optMinObservations=3, samples=["NA12878"], ctx=ctx)
    gts = rewrite_and_filter(gts, FilterArgs(), ctx=ctx)
    assert len(gts) == 2
    header = V.default_header(["NA12878"], [(gts[0].contigName, 249250621)], FilterArgs())
    path = tmp_path / "calls.vcf"
    V.write_vcf(path, header, V.group_by_variant(gts))
    back, h2 = V.load_genotypes(path)
    assert h2.samples == ["NA12878"] and h2.lines == header.sorted_lines()
    key = lambda g: g.variant.as_tuple()
    back = {key(g): g for g in back}
    assert len(back) == len(gts)
    for g in gts:
        b = back[key(g)]
        assert b.alleles == g.alleles and b.genotypeQuality == g.genotypeQuality and b.phased == g.phased
        assert (b.readDepth, b.referenceReadDepth, b.alternateReadDepth) == (g.readDepth, g.referenceReadDepth, g.alternateReadDepth)
        assert b.strandBiasComponents == list(g.strandBiasComponents)
        assert np.float32(b.rmsMapQ) == np.float32(g.rmsMapQ)
        assert np.float32(b.fisherStrandBiasPValue) == np.float32(g.fisherStrandBiasPValue)
        assert (b.filtersApplied, b.filtersPassed, b.filtersFailed) == (g.filtersApplied, g.filtersPassed, list(g.filtersFailed or []))
        assert V._pl_of(b.genotypeLikelihoods) == V._pl_of(g.genotypeLikelihoods)
