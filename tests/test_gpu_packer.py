"""The packer on the device (avo_pack_reads_device) against the host packer (avo_pack_reads, itself
pinned to the oracle's extractAlignmentOperators in test_packer_cpu.py): identical bytes on the
reference's SAM fixtures, on fuzzed CIGAR / MD strings, and on synthetic text at scale; and the
device-packed batch drives discoverAndCall to the oracle's genotypes."""
import numpy as np
import pytest

from helpers import compare_call, fuzz_records, load_fixture, oracle_reads_from_batch

pytestmark = pytest.mark.gpu

FIXTURES = ["NA12878.chr1.832736", "NA12878.chr1.875159", "NA12878.chr1.1777263", "NA12878_snp_A2G_chr20_225058"]


@pytest.fixture(scope="module")
def ctx():
    from avocado_b200.genotyping import Context
    c = Context(0)
    yield c
    c.close()


def _same(dev_batch, host_batch):
    d = dev_batch.to_host()
    h = host_batch
    assert (d.n_reads, d.n_blocks, d.n_ops, d.n_refb) == (h.n_reads, h.n_blocks, h.n_ops, h.n_refb)
    dm = np.asarray(d.meta).view(h.meta.dtype)[:h.n_reads]
    for f in h.meta.dtype.names:
        assert np.array_equal(dm[f], h.meta[f]), f
    assert np.array_equal(np.asarray(d.payload)[:h.n_blocks * 16], h.payload[:h.n_blocks * 16])
    assert np.array_equal(np.asarray(d.ops)[:h.n_ops], h.ops[:h.n_ops])
    assert np.array_equal(np.asarray(d.refb)[:h.n_refb], h.refb[:h.n_refb])


def _np_cols(cols):
    return {k: (np.frombuffer(bytes(v) + b"\0", np.uint8) if isinstance(v, (bytes, bytearray)) else v)
            for k, v in cols.items()}


def _available():
    import os
    from helpers import GOLDEN
    return [f[:-8] for f in sorted(os.listdir(os.path.join(GOLDEN, "sam"))) if f.endswith(".json.gz")]


@pytest.mark.parametrize("name", _available())
def test_reference_fixtures_pack_identically(ctx, name):
    from avocado_b200 import batch as B
    recs = [r for r in load_fixture(name)[0] if r.readMapped and r.start is not None]
    cols = _np_cols(B.text_columns(recs))
    host = B.pack_text_columns(cols)
    dev = ctx.pack_text_device(cols, source_index=True)
    _same(dev, host)
    assert np.array_equal(dev.source_index.cpu().numpy(), host.source_index)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fuzzed_cigar_and_md_pack_identically(ctx, seed):
    from avocado_b200 import batch as B
    recs = fuzz_records(seed, 20000)
    cols = _np_cols(B.text_columns(recs))
    host = B.pack_text_columns(cols)
    assert 0 < (host.meta["n_ops"] == 0).sum() < host.n_reads
    assert (host.meta["flags"] & 2).any()
    _same(ctx.pack_text_device(cols), host)
    keep = (np.arange(len(recs)) % 5 != 2).astype(np.uint8)
    _same(ctx.pack_text_device(cols, keep=keep), B.pack_text_columns(cols, keep=keep))


def test_empty_and_all_filtered(ctx):
    from avocado_b200 import batch as B
    recs = fuzz_records(4, 50)
    cols = _np_cols(B.text_columns(recs))
    b = ctx.pack_text_device(cols, keep=np.zeros(len(recs), np.uint8))
    assert (b.n_reads, b.n_blocks, b.n_ops, b.n_refb) == (0, 0, 0, 0)
    cols0 = _np_cols(B.text_columns([]))
    b = ctx.pack_text_device(cols0)
    assert b.n_reads == 0


def test_synthetic_text_packs_to_the_generated_batch(ctx):
    """text(batch) -> device packer == batch at 400k reads, from host and from device-resident text."""
    import torch
    from avocado_b200 import batch as B
    n = 400000
    p = B.synth_params(contig_len=n * 150 // 50, n_reads_total=n, softclip_frac=0.1, indel_rate=3e-4)
    host = B.synth_host(p)
    cols = B.text_columns_of(host)
    _same(ctx.pack_text_device(cols), host)
    dcols = {k: torch.from_numpy(v).cuda() for k, v in cols.items()}
    _same(ctx.pack_text_device(dcols), host)
    # 32-bit positions / offsets and 8-bit mapq (avo_text_reads.narrow), host and device resident
    slim = B.narrow_text_columns(cols)
    _same(ctx.pack_text_device(slim), host)
    dslim = {k: (torch.from_numpy(v.view(np.int32) if v.dtype == np.uint32 else v).cuda() if isinstance(v, np.ndarray) else v)
             for k, v in slim.items()}
    _same(ctx.pack_text_device(dslim), host)


def test_device_packed_batch_genotypes_like_the_oracle(oracle, ctx):
    from avocado_b200 import batch as B
    p = B.synth_params(seed=77, n_reads_total=12000, contig_len=40000, first_pos=500)
    host = B.synth_host(p)
    dev = ctx.pack_text_device(B.text_columns_of(host))
    rs = oracle_reads_from_batch(oracle, host)
    want_v = oracle.discover(rs, 25, 5)
    want = oracle.call(rs, [(c, s, e, r, a) for c, s, e, r, a, _ in want_v], 2, [], False, 93, 93, 1)
    sites, cols = ctx.discover_and_call_packed(dev, ploidy=2, phred=25, min_obs=5)
    keys = [("ctg", v.start, v.referenceAllele, v.alternateAllele) for v in sites.to_variants(["ctg"])]
    assert len(keys) == len(want_v)
    st = compare_call(keys, cols, want)
    assert st["sites"] > 10


def test_device_packer_evaluates_the_read_predicates(ctx):
    """PrefilterReads' read predicates inside k_pack_count / k_pack_records: the device
    packer given avo_prefilter keeps exactly the reads the per-record predicates keep
    (PrefilterReads.scala:142-209) and packs them byte for byte like the host packer."""
    import itertools
    import random
    from avocado_b200 import batch as B
    from avocado_b200.prefilter import PrefilterReadsArgs, prefilter_mask, prefilter_struct
    from avocado_b200.records import AlignmentRecord
    rng = random.Random(4)
    recs = []
    for i in range(3000):
        recs.append(AlignmentRecord(readMapped=rng.random() > 0.1, contigName="chr20", start=100 + i, end=110 + i, cigar="4M1X5M",
                                    mismatchingPositions="4A5", sequence="ACGTCCGTAC", qual="IIIIIIIIII",
                                    mapq=None if rng.random() < 0.1 else rng.choice([0, 5, 10, 11, 30, 60]),
                                    readNegativeStrand=rng.random() < 0.5, readName="r%d" % i,
                                    primaryAlignment=rng.random() > 0.2, duplicateRead=rng.random() < 0.2))
    cols = B.text_columns(recs)
    dcols = {k: (np.frombuffer(v + b"\0", np.uint8).copy() if isinstance(v, (bytes, bytearray)) else v) for k, v in cols.items()}
    for dup, nonprim, minq, contig in itertools.product([False, True], [False, True], [0, 10], ["chr20", "GL000"]):
        args = PrefilterReadsArgs(keepDuplicates=dup, keepNonPrimary=nonprim, minMappingQuality=minq)
        f = prefilter_struct(args, contig)
        want = [i for i, k in enumerate(prefilter_mask(recs, args)) if k and f.contig_kept]
        host = B.pack_text_columns(cols, threads=2, prefilter=f)
        dev = ctx.pack_text_device(dcols, prefilter=f, source_index=True)
        assert dev.source_index.cpu().numpy().tolist() == want == host.source_index.tolist()
        dh = dev.to_host()
        assert np.array_equal(np.asarray(dh.meta).view(np.uint8)[:host.n_reads * 32], host.meta.view(np.uint8))
        assert np.array_equal(np.asarray(dh.ops)[:host.n_ops], host.ops[:host.n_ops])
