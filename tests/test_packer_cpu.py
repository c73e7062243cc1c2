"""CPU tests of the host packer and of the C-ABI library surface (no GPU, no compute calls).

The packer is where ObservationOperator.extractAlignmentOperators (ObservationOperator.scala:42-171)
happens in this build; the oracle's own restatement of that function is the checker.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import load_fixture, oracle_reads_from_batch

FIXTURES = ["NA12878_snp_A2G_chr20_225058", "NA12878.chr1.832736", "NA12878.chr1.567239",
            "NA12878.1_1777263", "NA12878.chr1.905130", "NA12878.1_5274547", "NA12878.chr1.104160"]


def test_library_exports_every_declared_symbol():
    from avocado_b200 import _lib as L
    lib = L.load()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    declared = set()
    for h in ("avocado_b200.h", "avocado_b200_synth.h"):
        src = open(os.path.join(root, "include", h)).read()
        declared |= set(re.findall(r"^\s*(?:int|int64_t|void|const char\*)\s+(avo_\w+)\s*\(", src, re.M))
    assert declared, "no declarations found"
    assert declared == set(L.EXPORTED_SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert C.sizeof(L.ReadBatchStruct) == 4 * 8 + 2 * 4 + 4 * 4 + 6 * 8 + 5 * 8 + 2 * 4 + 8


@pytest.mark.parametrize("name", FIXTURES)
def test_packed_reads_round_trip_through_the_oracle(oracle, name):
    """pack -> unpack gives reads whose operators, sequence, qualities, strand and mapq are the
    ones the oracle extracts from the original text."""
    from avocado_b200 import batch as B
    recs = [r for r in load_fixture(name)[0] if r.readMapped]
    b = B.pack_reads(recs, sort=False)
    assert b.n_reads == len(recs)
    rs = oracle_reads_from_batch(oracle, b, contig="c")
    for i, r in enumerate(recs):
        d = oracle.read_as_dict(rs, i)
        want_ops = oracle.extract_alignment_operators(r.as_dict())
        m = b.meta[i]
        if not want_ops:
            assert m["n_ops"] == 0
            continue
        got_ops = oracle.extract_alignment_operators(d)
        assert got_ops == want_ops, (i, r.cigar, r.mismatchingPositions)
        assert d["sequence"] == r.sequence.upper() and d["qual"] == r.qual
        assert (d["start"], d["end"]) == (r.start, r.end)
        assert d["readNegativeStrand"] == bool(r.readNegativeStrand)
        assert d["mapq"] == r.mapq
        q = [ord(c) - 33 for c in r.qual[:4]] + [0] * 4
        assert int(m["qhead"]) == q[0] | q[1] << 8 | q[2] << 16 | q[3] << 24
        assert int(m["l_seq"]) == len(r.sequence)


def test_synthetic_host_batch_is_consistent(oracle):
    """The generator's packed batch decodes (the oracle checks every mismatch pair against the
    read's bases), is sorted by start, and its records index disjoint, in-range slices."""
    from avocado_b200 import batch as B
    p = B.synth_params(seed=11, n_reads_total=5000, contig_len=40000)
    b = B.synth_host(p)
    m = b.meta
    assert np.all(np.diff(m["start"]) >= 0)
    assert np.all(m["end"] > m["start"])
    assert np.array_equal(m["op_off"][1:], (m["op_off"] + m["n_ops"])[:-1])
    assert int(m["op_off"][-1]) + int(m["n_ops"][-1]) == b.n_ops
    assert np.all(np.diff(m["refb_off"].astype(np.int64)) >= 0) and int(m["refb_off"][-1]) <= b.n_refb
    rs = oracle_reads_from_batch(oracle, b)
    # the reference span of the operators equals end - start for every read
    for i in range(0, b.n_reads, 97):
        d = oracle.read_as_dict(rs, i)
        span = sum(int(n) for n, c in re.findall(r"(\d+)([=XD])", d["cigar"]))
        assert span == d["end"] - d["start"]
    s = b.slice(100, 300)
    rs2 = oracle_reads_from_batch(oracle, s)
    for j in (0, 57, 199):
        assert oracle.read_as_dict(rs2, j) == oracle.read_as_dict(rs, 100 + j)


def test_wire_columns_hold_the_same_information():
    """avo_pack_wire: the compact records drop only the offsets (prefix sums) and the operators
    keep their bits; batches that do not fit are refused, not truncated."""
    from avocado_b200 import _lib as L, batch as B
    p = B.synth_params(seed=4, n_reads_total=3000, contig_len=20000)
    b = B.synth_host(p)
    assert b.with_wire()
    m, w = b.meta, b.wire_meta
    for f in ("start", "l_seq", "n_ops", "mapq", "flags", "qhead"):
        assert np.array_equal(m[f].astype(np.int64), w[f].astype(np.int64)), f
    assert np.array_equal(m["end"].astype(np.int64), w["start"].astype(np.int64) + w["span"])
    blocks = (w["l_seq"].astype(np.int64) + 9) // 10
    assert np.array_equal(np.concatenate([[0], np.cumsum(blocks)[:-1]]), m["seq_off"])
    assert np.array_equal(np.concatenate([[0], np.cumsum(w["n_ops"].astype(np.int64))[:-1]]), m["op_off"])
    assert np.array_equal(np.concatenate([[0], np.cumsum(w["n_refb"].astype(np.int64))[:-1]]), m["refb_off"])
    assert np.array_equal(b.wire_ops[:b.n_ops].astype(np.uint32), b.ops[:b.n_ops])
    long_op = B.synth_host(p)
    long_op.ops[0] = (5000 << 4) | 0
    assert not long_op.with_wire()
    gap = B.synth_host(p)
    gap.meta["op_off"][5] += 1
    assert not gap.with_wire()


def _same_batch(a, b):
    assert (a.n_reads, a.n_blocks, a.n_ops, a.n_refb) == (b.n_reads, b.n_blocks, b.n_ops, b.n_refb)
    for f in a.meta.dtype.names:
        assert np.array_equal(a.meta[f], b.meta[f]), f
    assert np.array_equal(a.payload[:a.n_blocks * 16], b.payload[:a.n_blocks * 16])
    assert np.array_equal(a.ops[:a.n_ops], b.ops[:a.n_ops])
    assert np.array_equal(a.refb[:a.n_refb], b.refb[:a.n_refb])


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_text_round_trip_reproduces_the_packed_batch(threads):
    """pack(text(batch)) == batch, byte for byte: CIGAR M runs + MD carry everything the operators do
    (ObservationOperator.scala:42-171), and the threaded packer writes what the serial one does."""
    from avocado_b200 import batch as B
    n = 60000
    p = B.synth_params(contig_len=n * 150 // 40, n_reads_total=n, softclip_frac=0.1, indel_rate=5e-4)
    b = B.synth_host(p)
    cols = B.text_columns_of(b)
    assert cols["cigar_off"][-1] > 4 * n and cols["md_off"][-1] > 3 * n
    _same_batch(b, B.pack_text_columns(cols, threads=threads))
    # the same text with 32-bit positions / offsets and 8-bit mapq (avo_text_reads.narrow)
    slim = B.narrow_text_columns(cols)
    assert slim["start"].dtype == np.int32 and slim["seq_off"].dtype == np.uint32 and slim["mapq"].dtype == np.uint8
    _same_batch(b, B.pack_text_columns(slim, threads=threads))


def test_narrow_text_columns_keep_an_absent_mapq_absent():
    """mapq = None travels as a cleared flag bit in the narrow form; both forms pack to the same records
    (PrefilterReads keeps reads without a mapq, PrefilterReads.scala:201-209)."""
    from avocado_b200 import batch as B
    from avocado_b200.prefilter import PrefilterReadsArgs, prefilter_struct
    from avocado_b200.records import AlignmentRecord
    recs = [AlignmentRecord(readMapped=True, contigName="1", start=10 * i, end=10 * i + 20, sequence="ACGT" * 5, qual="I" * 20,
                            cigar="20M", mismatchingPositions="20", mapq=(None if i % 3 == 0 else 5 + 10 * (i % 5)),
                            readName="r%d" % i) for i in range(40)]
    cols = B.text_columns(recs)
    pf = prefilter_struct(PrefilterReadsArgs(minMappingQuality=10), "1")
    wide = B.pack_text_columns(cols, prefilter=pf)
    slim = B.pack_text_columns(B.narrow_text_columns(cols), prefilter=pf)
    assert 0 < wide.n_reads < 40
    _same_batch(wide, slim)
    assert np.array_equal(wide.source_index, slim.source_index)


def _walk_events(b):
    """DiscoverVariants.scala:112-252 per read, in Python, over a packed host batch -> (positions, pq words, indel events, first)."""
    m = b.meta
    pos_l, pq_l, ind, first = [], [], [], []
    for r in range(b.n_reads):
        if r % 32 == 0:
            first.append(len(pos_l))
        pos, rb, ff, qh = int(m["start"][r]), int(m["refb_off"][r]), True, int(m["qhead"][r])
        for k in range(int(m["n_ops"][r])):
            op = int(b.ops[int(m["op_off"][r]) + k])
            ln, code = op >> 4, op & 15
            if code == 1:
                for i in range(ln):
                    q = (qh >> (8 * i)) & 255 if i < 4 else int(b.payload[(int(m["seq_off"][r]) + i // 10) * 16 + 5 + i % 10])
                    pos_l.append(pos + i)
                    pq_l.append((int(b.refb[rb + i]) << 8) | q)
            if code in (2, 3) and not ff:
                ind.append((r, k))
            if code in (0, 1, 3):
                pos += ln
            rb += ln if code == 1 else (ln + 1) // 2 if code == 3 else 0
            if code <= 1:
                ff = False
    first.append(len(pos_l))
    return pos_l, pq_l, ind, first


def test_host_events_of_messy_reads_and_of_an_empty_batch():
    """Fuzzed operator mixes (mismatch runs longer than the four head qualities, insertions before the first aligned base,
    reads without operators) and the empty batch."""
    from avocado_b200 import batch as B
    from avocado_b200.genotyping import derive_events
    from helpers import fuzz_records
    recs = [r for r in fuzz_records(77, 3000) if r.readMapped]
    b = derive_events(B.pack_reads(recs, sort=True))
    pos_l, pq_l, ind, first = _walk_events(b)
    ev = b.events
    assert (ev["n_snv"], ev["n_indel"]) == (len(pos_l), len(ind)) and len(pos_l) > 500 and len(ind) > 50
    assert np.array_equal(ev["snv_pos"][:len(pos_l)], np.array(pos_l, np.int32))
    assert np.array_equal(ev["snv_pq"][:len(pq_l)], np.array(pq_l, np.uint16))
    assert np.array_equal(ev["snv_first"], np.array(first, np.uint32))
    assert [(int(x), int(y)) for x, y in zip(ev["indel_read"][:len(ind)], ev["indel_op"][:len(ind)])] == ind
    ops = [int(b.ops[int(b.meta["op_off"][r]) + k]) for r in range(b.n_reads) for k in range(int(b.meta["n_ops"][r]))]
    assert max(o >> 4 for o in ops if o & 15 == 1) > 4 and (b.meta["n_ops"] == 0).any()
    empty = derive_events(B.pack_reads([]))
    assert (empty.events["n_snv"], empty.events["n_indel"]) == (0, 0) and int(empty.events["snv_first"][0]) == 0


def test_host_events_are_the_walk_of_the_operators():
    """avo_derive_events on a host batch (no device needed) against the per-read walk of DiscoverVariants.scala:112-252
    written out in Python: every mismatching base with the quality qual(i) the reference tests, every I / D operator
    after the first aligned one, in read order; snv_first indexes every 32nd read; exact bounds."""
    from avocado_b200 import batch as B
    from avocado_b200.genotyping import derive_events
    p = B.synth_params(seed=5, n_reads_total=6000, contig_len=30000, first_pos=100, indel_rate=2e-3, softclip_frac=0.1)
    b = derive_events(B.synth_host(p))
    ev, m = b.events, b.meta
    pos_l, pq_l, ind, first = [], [], [], []
    for r in range(b.n_reads):
        if r % 32 == 0:
            first.append(len(pos_l))
        pos, rb, ff, qh = int(m["start"][r]), int(m["refb_off"][r]), True, int(m["qhead"][r])
        for k in range(int(m["n_ops"][r])):
            op = int(b.ops[int(m["op_off"][r]) + k])
            ln, code = op >> 4, op & 15
            if code == 1:
                for i in range(ln):
                    q = (qh >> (8 * i)) & 255 if i < 4 else int(b.payload[(int(m["seq_off"][r]) + i // 10) * 16 + 5 + i % 10])
                    pos_l.append(pos + i)
                    pq_l.append((int(b.refb[rb + i]) << 8) | q)
            if code in (2, 3) and not ff:
                ind.append((r, k))
            if code in (0, 1, 3):
                pos += ln
            rb += ln if code == 1 else (ln + 1) // 2 if code == 3 else 0
            if code <= 1:
                ff = False
    first.append(len(pos_l))
    assert (ev["n_snv"], ev["n_indel"]) == (len(pos_l), len(ind)) and len(ind) > 100 and len(pos_l) > 1000
    assert np.array_equal(ev["snv_pos"][:len(pos_l)], np.array(pos_l, np.int32))
    assert np.array_equal(ev["snv_pq"][:len(pq_l)], np.array(pq_l, np.uint16))
    assert np.array_equal(ev["snv_first"], np.array(first, np.uint32))
    assert [(int(x), int(y)) for x, y in zip(ev["indel_read"][:len(ind)], ev["indel_op"][:len(ind)])] == ind
    st, en = m["start"].astype(np.int64), m["end"].astype(np.int64)
    assert tuple(ev["stats"]) == (int(st.min()), int(en.max()), int((en - st).max()))
    b.meta["start"][100] -= 1000
    with pytest.raises(Exception):
        derive_events(b)


def test_threaded_packer_matches_serial_on_the_reference_fixtures():
    from avocado_b200 import batch as B
    recs = [r for name in FIXTURES[:4] for r in load_fixture(name)[0] if r.readMapped]
    recs = recs * 12                       # enough reads for several chunks
    recs.sort(key=lambda r: r.start)
    cols = B.text_columns(recs)
    keep = np.array([(i % 7) != 3 for i in range(len(recs))], np.uint8)
    for k in (None, keep):
        _same_batch(B.pack_text_columns(cols, threads=1, keep=k), B.pack_text_columns(cols, threads=4, keep=k))


def test_packer_agrees_with_the_oracle_on_fuzzed_cigar_and_md(oracle):
    """extractAlignmentOperators (ObservationOperator.scala:42-171) on consistent, slightly broken
    and random CIGAR / MD strings: the packed operators are the oracle's, and reads the oracle
    throws on are packed with zero operators."""
    from avocado_b200 import batch as B
    from helpers import fuzz_records, ops_string
    recs = [r for r in fuzz_records(11, 6000) if r.readMapped]
    b = B.pack_reads(recs, sort=False)
    assert b.n_reads == len(recs)
    thrown = kept = 0
    for i, r in enumerate(recs):
        try:
            want = oracle.extract_alignment_operators(r.as_dict())
        except RuntimeError:
            want = ""
            thrown += 1
        need = sum(int(x[:-1].split(":")[0][:-1] if ":" in x else x[:-1]) for x in want.split(",") if x and
                   (x.split(":")[0][-1] in "=XIS")) if want else 0
        if want and (need > len(r.sequence) or need > len(r.qual)):
            want = ""            # the packer drops reads whose strings are shorter than the CIGAR
        if want and re.search(r"(?<![0-9])0+[MIDX=]", r.cigar):
            want = ""            # ... and reads with a zero-length non-clip element (htsjdk calls them invalid)
        # the packed form holds BAM's 4-bit codes: MD letters outside them (U, X) are stored as N
        want = ",".join(x.split(":")[0] + ":" + re.sub("[UX]", "N", x.split(":")[1]) if ":" in x else x
                        for x in want.split(","))
        want = ",".join(x for x in want.split(",") if x not in ("0S", "0H"))   # empty clips are not stored
        got = ops_string(b, i)
        assert got == want, (i, r.cigar, r.mismatchingPositions, got, want)
        kept += bool(want)
    assert thrown > 300 and kept > 3000


def _expand_wire6(b):
    """numpy restatement of the device-side expansion of the 6-byte form -> META_DTYPE records"""
    from avocado_b200 import batch as B
    w = b.wire6_meta
    n = len(w)
    k = np.arange(b.n_ops)
    ops = (b.wire6_oplen[:b.n_ops].astype(np.int64) << 4) | ((b.wire6_opcode[k >> 1] >> ((k & 1) * 4)) & 15)
    assert np.array_equal(ops, b.ops[:b.n_ops].astype(np.int64))
    m = np.zeros(n, B.META_DTYPE)
    d = (w["dstart_flags"] >> 2).astype(np.int64)
    m["start"] = b.wire6_base + np.cumsum(d)
    m["flags"] = w["dstart_flags"] & 3
    m["n_ops"], m["mapq"] = w["n_ops"], w["mapq"]
    m["qhead"] = w["q0"].astype(np.uint32) | (w["q1"].astype(np.uint32) << 8)
    op_off = np.concatenate([[0], np.cumsum(w["n_ops"].astype(np.int64))])
    m["op_off"] = op_off[:-1]
    ln, code = ops >> 4, ops & 15
    per = lambda mask, val: np.add.reduceat(np.where(mask, val, 0), np.minimum(op_off[:-1], max(len(ops) - 1, 0))) * (w["n_ops"] > 0)
    span = per((code <= 1) | (code == 3), ln)
    lseq = per((code <= 1) | (code == 2) | (code == 4), ln)
    nrefb = per(code == 1, ln) + per(code == 3, (ln + 1) >> 1)
    end = m["start"].astype(np.int64) + span
    for e in b.wire6_exc:
        lseq[e["index"]] = e["l_seq"]
        end[e["index"]] = e["end"]
    m["end"], m["l_seq"] = end, lseq
    m["seq_off"] = np.concatenate([[0], np.cumsum((lseq + 9) // 10)[:-1]])
    m["refb_off"] = np.concatenate([[0], np.cumsum(nrefb)[:-1]])
    return m


def test_wire6_columns_hold_the_same_information():
    """avo_pack_wire6: 6 bytes per read; starts as differences, end / l_seq / refb bytes implied by the
    operators, the reads for which that fails listed as exceptions."""
    from avocado_b200 import batch as B
    from helpers import fuzz_records
    p = B.synth_params(seed=4, n_reads_total=3000, contig_len=20000, softclip_frac=0.2)
    b = B.synth_host(p)
    assert b.with_wire6() and len(b.wire6_exc) == 0
    got = _expand_wire6(b)
    for f in b.meta.dtype.names:
        want = b.meta[f] & 0xFFFF if f == "qhead" else b.meta[f]      # two of the four head qualities travel
        assert np.array_equal(got[f], want), f
    # fuzzed reads: failed extractions (no operators) and over-long strings become exceptions
    from helpers import short_operators
    recs = short_operators([r for r in fuzz_records(5, 3000) if r.readMapped])
    fz = B.pack_reads(recs, sort=False)
    assert fz.with_wire6(max_exceptions=fz.n_reads)
    assert 0 < len(fz.wire6_exc) < fz.n_reads
    got = _expand_wire6(fz)
    for f in fz.meta.dtype.names:
        want = fz.meta[f] & 0xFFFF if f == "qhead" else fz.meta[f]
        assert np.array_equal(got[f], want), f
    assert not fz.with_wire6(max_exceptions=3)
    far = B.synth_host(p)
    far.meta["start"][100:] += 20000                  # a start difference that does not fit 14 bits
    far.meta["end"][100:] += 20000
    assert not far.with_wire6()
    unsorted = B.synth_host(p)
    unsorted.meta["start"][7] -= 500
    assert not unsorted.with_wire6()
    long_op = B.synth_host(p)
    long_op.ops[0] = (300 << 4) | 0                   # an operator that needs more than a length byte
    assert not long_op.with_wire6() and long_op.with_wire() is not None


def test_headers_are_plain_c(tmp_path):
    """The boundary is a C ABI: both headers and the C example compile as C11 (no C++, no torch types)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    obj = str(tmp_path / "example.o")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-c",
                           os.path.join(root, "examples", "discover_and_call.c"), "-o", obj])
    assert os.path.getsize(obj) > 0


def test_ctypes_mirrors_have_the_sizes_of_the_c_structs(tmp_path):
    """Every struct that crosses the C ABI: sizeof in C (gcc on include/*.h) == ctypes.sizeof of the mirror."""
    import subprocess
    from avocado_b200 import _lib as L
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pairs = {
        "avo_read_batch": L.ReadBatchStruct, "avo_sites": L.SitesStruct, "avo_cnv": L.CnvStruct,
        "avo_params": L.ParamsStruct, "avo_sites_out": L.SitesOutStruct, "avo_site_results": L.SiteResultsStruct,
        "avo_ref_model_out": L.RefModelOutStruct, "avo_timing": L.TimingStruct, "avo_text_reads": L.TextReadsStruct,
        "avo_text_out": L.TextOutStruct, "avo_packed_out": L.PackedOutStruct, "avo_filter_params": L.FilterParamsStruct,
        "avo_filter_out": L.FilterOutStruct, "avo_sample_rows": L.SampleRowsStruct, "avo_square_out": L.SquareOutStruct,
        "avo_trio": L.TrioStruct, "avo_joint_in": L.JointInStruct, "avo_joint_out": L.JointOutStruct,
        "avo_synth_params": L.SynthParamsStruct, "avo_depths_in": L.DepthsInStruct, "avo_site_depths": L.SiteDepthsStruct,
        "avo_shard_plan": L.ShardPlanStruct, "avo_read_events": L.ReadEventsStruct,
    }
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "avocado_b200.h"\n#include "avocado_b200_synth.h"\nint main(void) {\n' +
                   "".join('  printf("%s %%zu\\n", sizeof(%s));\n' % (n, n) for n in pairs) +
                   '  printf("avo_read_meta %zu\\navo_read_wire %zu\\navo_read_wire6 %zu\\n", sizeof(avo_read_meta), '
                   'sizeof(avo_read_wire), sizeof(avo_read_wire6));\n  return 0;\n}\n')
    exe = str(tmp_path / "sizes")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", exe])
    got = dict(line.split() for line in subprocess.check_output([exe], text=True).splitlines())
    for name, mirror in pairs.items():
        assert int(got[name]) == C.sizeof(mirror), name
    assert (int(got["avo_read_meta"]), int(got["avo_read_wire"]), int(got["avo_read_wire6"])) == (32, 16, 6)


def test_jni_glue_compiles_against_the_headers():
    """examples/jni/avocado_b200_jni.c (the binding INTEGRATION.md describes) against include/*.h and a
    stand-in for the JDK's jni.h with the JDK's names and call shapes: the sketch cannot rot."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(root, "tests", "mock_jni"),
                           "-I", os.path.join(root, "include"), os.path.join(root, "examples", "jni", "avocado_b200_jni.c")])
