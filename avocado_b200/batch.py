"""Columnar read batches and site tables (the data that crosses the C ABI), the host packer
wrapper, and the synthetic generator wrappers.

A :class:`ReadBatch` / :class:`SiteTable` holds either numpy arrays (host) or torch CUDA tensors
(device); ``as_struct`` produces the ctypes struct the library consumes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L
from .records import AlignmentRecord, Variant

NIB = "=ACMGRSVTWYHKDBN"
_NIB_OF = {c: i for i, c in enumerate(NIB)}

# avo_read_meta (include/avocado_b200.h): one 32-byte record per read
META_DTYPE = np.dtype([("start", "<i4"), ("end", "<i4"), ("seq_off", "<u4"), ("op_off", "<u4"),
                       ("refb_off", "<u4"), ("n_ops", "<u2"), ("mapq", "u1"), ("flags", "u1"),
                       ("qhead", "<u4"), ("l_seq", "<u4")])
assert META_DTYPE.itemsize == 32

# avo_read_wire: the compact 16-byte wire form of the record (offsets are rebuilt on the device)
WIRE_DTYPE = np.dtype([("start", "<i4"), ("span", "<u2"), ("l_seq", "<u2"), ("n_ops", "u1"), ("n_refb", "u1"),
                       ("mapq", "u1"), ("flags", "u1"), ("qhead", "<u4")])
assert WIRE_DTYPE.itemsize == 16

# avo_read_wire6 / avo_wire_exception: the 6-byte form (start differences; sizes implied by the operators)
WIRE6_DTYPE = np.dtype([("dstart_flags", "<u2"), ("n_ops", "u1"), ("mapq", "u1"), ("q0", "u1"), ("q1", "u1")])
WIRE_EXC_DTYPE = np.dtype([("index", "<u4"), ("l_seq", "<u4"), ("end", "<i4"), ("reserved", "<u4")])
assert WIRE6_DTYPE.itemsize == 6 and WIRE_EXC_DTYPE.itemsize == 16

READ_COLUMNS = [("meta", META_DTYPE), ("payload", np.uint8), ("ops", np.uint32), ("refb", np.uint8)]
BLOCK_BASES, BLOCK_BYTES = 10, 16     # payload: 10 bases (4 bit) + their 10 qualities per 16-byte block


def blocks_of(l_seq):
    return (l_seq + BLOCK_BASES - 1) // BLOCK_BASES


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    return a.data_ptr()  # torch tensor


def _is_device(a):
    return not isinstance(a, np.ndarray)


@dataclass
class ReadBatch:
    """Reads of one contig, sorted by start (include/avocado_b200.h: avo_read_batch)."""
    n_reads: int
    n_blocks: int
    n_ops: int
    n_refb: int
    contig_id: int = 0
    meta: object = None      # host: structured array of META_DTYPE; device: uint8 tensor [n * 32]
    payload: object = None   # [n_blocks * 32] bytes
    ops: object = None
    refb: object = None
    source_index: Optional[np.ndarray] = None
    stats: Optional[tuple] = None   # (min_start, max_end, max_span): avo_batch_stats, optional
    wire_meta: Optional[np.ndarray] = None   # host batches: compact wire columns (avo_pack_wire)
    wire_ops: Optional[np.ndarray] = None
    wire6_meta: Optional[np.ndarray] = None  # host batches: the 6-byte form (avo_pack_wire6) ...
    wire6_oplen: Optional[np.ndarray] = None   # ... with its operators: a length byte
    wire6_opcode: Optional[np.ndarray] = None  # ... and a code nibble each
    wire6_exc: Optional[np.ndarray] = None
    wire6_base: int = 0
    # discovery's candidate events (Context.derive_events): dict(snv_pos, snv_pq, snv_first, indel_read, indel_op,
    # n_snv, n_indel, stats) in the batch's memory; a device batch that has them is discovered by counting them
    events: Optional[dict] = None

    def with_wire6(self, alloc=None, max_exceptions=None):
        """Adds the 6-byte wire columns (host batches): start differences, sizes implied by the
        operators, exceptions listed.  Returns False when the batch does not fit the form."""
        assert not self.on_device
        alloc = alloc or (lambda n: np.zeros(n, np.uint8))
        cap = max(16, self.n_reads // 64) if max_exceptions is None else max_exceptions
        wm = alloc(max(self.n_reads, 1) * 6).view(WIRE6_DTYPE)
        wl = alloc(max(self.n_ops, 1))
        wc = alloc(max((self.n_ops + 1) // 2, 1))
        ex = alloc(max(cap, 1) * 16).view(WIRE_EXC_DTYPE)
        ne, base = C.c_int64(), C.c_int32()
        s = self.as_struct()
        rc = L.load().avo_pack_wire6(C.byref(s), wm.ctypes.data, wl.ctypes.data, wc.ctypes.data, ex.ctypes.data, cap,
                                     C.byref(ne), C.byref(base))
        if rc == L.AVO_E_UNSUPPORTED:
            return False
        if rc != 0:
            raise L.AvocadoError(rc, "avo_pack_wire6 failed")
        self.wire6_meta, self.wire6_oplen, self.wire6_opcode = wm[:self.n_reads], wl, wc
        self.wire6_exc, self.wire6_base = ex[:ne.value], base.value
        return True

    def with_wire(self, alloc=None):
        """Adds the compact wire columns (host batches): the library then copies 16-byte records
        and 16-bit operators instead of `meta` / `ops`.  `alloc(n_bytes) -> uint8 ndarray` lets the
        caller supply pinned memory.  Returns False when the batch does not fit the compact form."""
        assert not self.on_device
        alloc = alloc or (lambda n: np.zeros(n, np.uint8))
        wm = alloc(max(self.n_reads, 1) * 16).view(WIRE_DTYPE)
        wo = alloc(max(self.n_ops, 1) * 2).view(np.uint16)
        s = self.as_struct()
        rc = L.load().avo_pack_wire(C.byref(s), wm.ctypes.data, wo.ctypes.data)
        if rc == L.AVO_E_UNSUPPORTED:
            return False
        if rc != 0:
            raise L.AvocadoError(rc, "avo_pack_wire failed")
        self.wire_meta, self.wire_ops = wm[:self.n_reads], wo[:max(self.n_ops, 1)]
        return True

    def with_stats(self):
        """Fills `stats` (what a packer knows for free) so that the library can skip its own
        statistics pass; the discovery kernel still verifies the bounds and the sort order."""
        if self.n_reads == 0:
            self.stats = None
        elif self.on_device:
            m = self.meta[:self.n_reads * 32].view(__import__("torch").int32).view(-1, 8)
            st, en = m[:, 0], m[:, 1]
            self.stats = (int(st.min()), int(en.max()), int((en - st).max()))
        else:
            st, en = self.meta["start"].astype(np.int64), self.meta["end"].astype(np.int64)
            self.stats = (int(st.min()), int(en.max()), int((en - st).max()))
        return self

    @property
    def on_device(self):
        return _is_device(self.meta)

    def __getattr__(self, name):
        # per-read scalar columns (host batches): batch.start, batch.end, batch.mapq, ...
        if name in META_DTYPE.names:
            meta = self.__dict__.get("meta")
            if isinstance(meta, np.ndarray):
                return meta[name]
        raise AttributeError(name)

    def as_struct(self):
        s = L.ReadBatchStruct()
        s.n_reads, s.n_blocks, s.n_ops, s.n_refb = self.n_reads, self.n_blocks, self.n_ops, self.n_refb
        s.contig_id = self.contig_id
        s.mem = L.MEM_DEVICE if self.on_device else L.MEM_HOST
        if self.stats is not None:
            s.stats_min_start, s.stats_max_end, s.stats_max_span = self.stats
            s.stats_valid = 1
        for name, _ in READ_COLUMNS:
            setattr(s, name, _ptr(getattr(self, name)))
        if self.wire_meta is not None and self.wire_ops is not None and not self.on_device:
            s.wire_meta, s.wire_ops = _ptr(self.wire_meta), _ptr(self.wire_ops)
        if self.wire6_meta is not None and self.wire6_oplen is not None and not self.on_device:
            s.wire6_meta = _ptr(self.wire6_meta)
            s.wire6_oplen, s.wire6_opcode = _ptr(self.wire6_oplen), _ptr(self.wire6_opcode)
            s.wire6_exc = _ptr(self.wire6_exc) if len(self.wire6_exc) else None
            s.n_wire6_exc, s.wire6_base = len(self.wire6_exc), self.wire6_base
        if self.events is not None:
            # one struct per events dict, kept alive with the batch: a caller may cache the batch struct
            # (ShardedPipeline does) while other calls build theirs
            held = self.__dict__.get("_events_struct")
            if held is None or held[0] is not self.events:
                held = self._events_struct = (self.events, self.events_struct())
            s.events = C.cast(C.pointer(held[1]), C.c_void_p)
        return s

    def events_struct(self):
        ev = self.events
        e = L.ReadEventsStruct()
        e.n_snv, e.n_indel = ev["n_snv"], ev["n_indel"]
        for k in ("snv_pos", "snv_pq", "snv_first", "indel_read", "indel_op"):
            setattr(e, k, _ptr(ev[k]))
        e.stats_min_start, e.stats_max_end, e.stats_max_span = ev["stats"]
        e.stats_valid = 1
        return e

    def nbytes(self):
        """Bytes of the packed batch = the algorithmic bytes of one full scan."""
        tot = 0
        for name, _ in READ_COLUMNS:
            a = getattr(self, name)
            tot += a.nbytes if isinstance(a, np.ndarray) else a.numel() * a.element_size()
        return tot

    def events_nbytes(self):
        """Bytes of the candidate events a device batch carries (0 without)."""
        if self.events is None:
            return 0
        return sum((a.nbytes if isinstance(a, np.ndarray) else a.numel() * a.element_size())
                   for a in self.events.values() if hasattr(a, "shape"))

    def to_device(self, device="cuda:0", pinned=False):
        import torch
        kw = {}
        out = ReadBatch(self.n_reads, self.n_blocks, self.n_ops, self.n_refb, self.contig_id)
        for name, _ in READ_COLUMNS:
            a = getattr(self, name)
            t = torch.from_numpy(_torch_view(a))
            if pinned:
                t = t.pin_memory()
            setattr(out, name, t.to(device, non_blocking=pinned, **kw))
        out.stats = self.stats
        if self.events is not None:
            out.events = {k: (torch.from_numpy(_torch_view(v)).to(device) if isinstance(v, np.ndarray) else v)
                          for k, v in self.events.items()}
        return out

    def to_host(self):
        if not self.on_device:
            return self
        out = ReadBatch(self.n_reads, self.n_blocks, self.n_ops, self.n_refb, self.contig_id)
        for name, dt in READ_COLUMNS:
            setattr(out, name, _from_torch(getattr(self, name), dt))
        out.stats = self.stats
        return out

    def slice(self, lo, hi):
        """Host-side sub-batch of reads [lo, hi) (offsets rebased)."""
        assert not self.on_device
        m = self.meta[lo:hi].copy()
        if hi > lo:
            last = self.meta[hi - 1]
            b0 = int(m["seq_off"][0]); b1 = int(last["seq_off"]) + blocks_of(int(last["l_seq"]))
            o0 = int(m["op_off"][0]); o1 = int(last["op_off"]) + int(last["n_ops"])
            f0 = int(m["refb_off"][0])
            f1 = int(self.meta[hi]["refb_off"]) if hi < self.n_reads else self.n_refb
        else:
            b0 = b1 = o0 = o1 = f0 = f1 = 0
        m["seq_off"] -= np.uint32(b0); m["op_off"] -= np.uint32(o0); m["refb_off"] -= np.uint32(f0)
        out = ReadBatch(hi - lo, b1 - b0, o1 - o0, f1 - f0, self.contig_id)
        out.meta = m
        out.payload = np.ascontiguousarray(self.payload[b0 * BLOCK_BYTES:b1 * BLOCK_BYTES])
        out.ops = np.ascontiguousarray(self.ops[o0:o1])
        out.refb = np.ascontiguousarray(self.refb[f0:f1])
        return out.with_stats() if self.stats is not None else out


_MERGE_CTX = {}


def _merge_ctx(device):
    from .genotyping import Context
    if device not in _MERGE_CTX:
        _MERGE_CTX[device] = Context(device)
    return _MERGE_CTX[device]


def merge_batches(batches: Sequence[ReadBatch], ctx=None, stats=None) -> ReadBatch:
    """The union of several batches of ONE contig as a batch of its own, stably sorted by start
    (`reads.union(...)` of CLI/TrioGenotyper.scala:216: discovery over a trio counts the candidates
    of all three samples together).  Payload, operator and mismatch columns are concatenated, the
    records are merged and their three offsets rebased.  Host batches: numpy.  Device batches: one kernel
    (avo_merge_meta, on `ctx`'s stream) merges the records; `stats` = (min_start, max_end, max_span) of the
    union if the caller knows them (saves a reduction and its synchronisation)."""
    assert len(batches) > 0 and len({b.on_device for b in batches}) == 1
    assert len({b.contig_id for b in batches}) == 1
    dev = batches[0].on_device
    out = ReadBatch(sum(b.n_reads for b in batches), sum(b.n_blocks for b in batches),
                    sum(b.n_ops for b in batches), sum(b.n_refb for b in batches), batches[0].contig_id)
    blk = np.cumsum([0] + [b.n_blocks for b in batches])
    ops = np.cumsum([0] + [b.n_ops for b in batches])
    rfb = np.cumsum([0] + [b.n_refb for b in batches])
    if dev:
        import torch
        # columns that already lie back to back in one allocation (a shim that uploads the samples into one
        # buffer; bench.py's trio mode) are used in place, otherwise concatenated
        def joined(name, unit, pad):
            parts = [getattr(b, name) for b in batches]
            sizes = [b.n_blocks * BLOCK_BYTES if name == "payload" else (b.n_ops if name == "ops" else b.n_refb) for b in batches]
            adj = all(parts[k].data_ptr() + sizes[k] * parts[k].element_size() == parts[k + 1].data_ptr()
                      for k in range(len(parts) - 1))
            if adj and getattr(batches[0], "_arena", None) is not None:
                return batches[0]._arena[name]
            return torch.cat([p[:n] for p, n in zip(parts, sizes)] + [parts[0][:0].new_zeros(pad)])
        out.payload, out.ops, out.refb = joined("payload", 1, 16), joined("ops", 4, 4), joined("refb", 1, 16)
        own_ctx = ctx is None
        ctx = ctx or _merge_ctx(batches[0].meta.device.index or 0)
        merged = torch.empty(max(out.n_reads, 1) * 32, dtype=torch.uint8, device=batches[0].meta.device)
        K = len(batches)
        ptrs = (C.c_void_p * K)(*[b.meta.data_ptr() for b in batches])
        cnt = (C.c_int64 * K)(*[b.n_reads for b in batches])
        u32 = lambda a: (C.c_uint32 * K)(*[int(x) & 0xFFFFFFFF for x in a[:K]])
        ctx._check(ctx.lib.avo_merge_meta(ctx.h, K, ptrs, cnt, u32(blk), u32(ops), u32(rfb), merged.data_ptr()))
        out.meta = merged
        if own_ctx:                       # the helper context has a stream of its own: wait for the merge here
            torch.cuda.synchronize(batches[0].meta.device)
        if stats is None and all(b.stats is not None for b in batches):
            stats = (min(b.stats[0] for b in batches), max(b.stats[1] for b in batches), max(b.stats[2] for b in batches))
        if stats is not None:
            out.stats = stats
            return out
    else:
        recs = []
        for k, b in enumerate(batches):
            m = b.meta[:b.n_reads].copy()
            m["seq_off"] += np.uint32(blk[k]); m["op_off"] += np.uint32(ops[k]); m["refb_off"] += np.uint32(rfb[k])
            recs.append(m)
        m = np.concatenate(recs)
        out.meta = np.ascontiguousarray(m[np.argsort(m["start"], kind="stable")])
        out.payload = np.concatenate([b.payload[:b.n_blocks * BLOCK_BYTES] for b in batches] + [np.zeros(16, np.uint8)])
        out.ops = np.concatenate([b.ops[:b.n_ops] for b in batches] + [np.zeros(4, np.uint32)])
        out.refb = np.concatenate([b.refb[:b.n_refb] for b in batches] + [np.zeros(16, np.uint8)])
    return out.with_stats()


_TORCH_VIEW = {np.dtype(np.uint32): np.int32, np.dtype(np.uint16): np.int16}


def _torch_view(a: np.ndarray) -> np.ndarray:
    """torch has no uint32: reinterpret as int32 (bit pattern preserved); records as bytes."""
    if a.dtype == META_DTYPE:
        return a.view(np.uint8)
    v = _TORCH_VIEW.get(a.dtype)
    return a.view(v) if v is not None else a


def _from_torch(t, dtype) -> np.ndarray:
    a = t.detach().cpu().numpy()
    if np.dtype(dtype) == META_DTYPE:
        return a.view(np.uint8)[:(a.nbytes // 32) * 32].view(META_DTYPE)
    return a.view(dtype) if a.dtype != np.dtype(dtype) else a


def _device_empty(n, dtype, device):
    import torch
    tdt = {np.int32: torch.int32, np.uint32: torch.int32, np.uint8: torch.uint8, np.uint16: torch.int16,
           np.float64: torch.float64, np.float32: torch.float32, np.int64: torch.int64}[dtype]
    return torch.empty(max(int(n), 1), dtype=tdt, device=device)


# ---------------------------------------------------------------------------------------------
# Site tables
# ---------------------------------------------------------------------------------------------
@dataclass
class SiteTable:
    """Variant sites sorted by (contig, start, start+ref_len) (avo_sites)."""
    n: int
    n_allele_nibbles: int
    start: object
    ref_len: object
    alt_len: object
    allele_off: object
    alleles4: object
    contig: object = None
    count: object = None

    @property
    def on_device(self):
        return _is_device(self.start)

    def as_struct(self):
        s = L.SitesStruct()
        s.n, s.n_allele_nibbles = self.n, self.n_allele_nibbles
        s.mem = L.MEM_DEVICE if self.on_device else L.MEM_HOST
        for name in ("contig", "start", "ref_len", "alt_len", "allele_off", "alleles4"):
            setattr(s, name, _ptr(getattr(self, name)))
        return s

    def to_host(self):
        if not self.on_device:
            return self
        f = _from_torch
        return SiteTable(self.n, self.n_allele_nibbles, f(self.start, np.int32), f(self.ref_len, np.int32),
                         f(self.alt_len, np.int32), f(self.allele_off, np.uint32), f(self.alleles4, np.uint8),
                         None if self.contig is None else f(self.contig, np.int32),
                         None if self.count is None else f(self.count, np.int32))

    def alleles(self, i):
        """(ref, alt) strings of site i (host tables only)."""
        off = int(self.allele_off[i])
        rl, al = int(self.ref_len[i]), int(self.alt_len[i])
        s = [NIB[(self.alleles4[(off + k) >> 1] >> (0 if (off + k) & 1 else 4)) & 15] for k in range(rl + al)]
        return "".join(s[:rl]), "".join(s[rl:])

    def to_variants(self, contig_names: Sequence[str] = ("0",), contig_id=0) -> List[Variant]:
        t = self.to_host()
        out = []
        for i in range(t.n):
            ref, alt = t.alleles(i)
            c = contig_names[int(t.contig[i])] if t.contig is not None else contig_names[contig_id]
            out.append(Variant(c, int(t.start[i]), int(t.start[i]) + len(ref), ref, alt))
        return out


def sites_from_variants(variants: Sequence[Variant], contig_rank=None) -> SiteTable:
    """Builds a host :class:`SiteTable` sorted the way ``Forest.apply`` sorts its keys
    (TreeRegionJoin.scala:43-50: by contig, start, end).  ``contig_rank`` maps contig name -> rank
    (default: ranks by name order, the ReferenceRegion ordering)."""
    if contig_rank is None:
        names = sorted({v.contigName for v in variants})
        contig_rank = {c: i for i, c in enumerate(names)}
    for v in variants:
        if v.end != v.start + len(v.referenceAllele):
            raise ValueError("variant end must equal start + len(referenceAllele): %r" % (v,))
        if v.alternateAllele is None:
            raise ValueError("variants to call need an alternate allele")
    order = sorted(range(len(variants)), key=lambda i: (contig_rank[variants[i].contigName],
                                                         variants[i].start, variants[i].end))
    n = len(order)
    start = np.zeros(n, np.int32)
    ref_len = np.zeros(n, np.int32)
    alt_len = np.zeros(n, np.int32)
    contig = np.zeros(n, np.int32)
    allele_off = np.zeros(n + 1, np.uint32)
    nibs: List[int] = []
    for k, i in enumerate(order):
        v = variants[i]
        start[k], ref_len[k], alt_len[k] = v.start, len(v.referenceAllele), len(v.alternateAllele)
        contig[k] = contig_rank[v.contigName]
        allele_off[k] = len(nibs)
        nibs.extend(_NIB_OF.get(c, 15) for c in (v.referenceAllele + v.alternateAllele).upper())
        if len(nibs) & 1:
            nibs.append(0)
    allele_off[n] = len(nibs)
    packed = np.zeros(max(len(nibs) // 2, 1), np.uint8)
    if nibs:
        a = np.array(nibs, np.uint8)
        packed[:len(nibs) // 2] = (a[0::2] << 4) | a[1::2]
    t = SiteTable(n, len(nibs), start, ref_len, alt_len, allele_off, packed, contig)
    t.order = order  # input index of each sorted row
    return t


# ---------------------------------------------------------------------------------------------
# Packer (host): AlignmentRecords -> ReadBatch through avo_pack_reads
# ---------------------------------------------------------------------------------------------
def _blob(strings):
    offs = np.zeros(len(strings) + 1, np.int64)
    enc = [s.encode() if s is not None else b"" for s in strings]
    np.cumsum([len(e) for e in enc], out=offs[1:])
    return b"".join(enc), offs


def text_columns(records: Sequence[AlignmentRecord]):
    """The text columns avo_pack_reads consumes (also what the oracle's make_reads takes)."""
    n = len(records)
    start = np.array([-1 if r.start is None else r.start for r in records], np.int64).reshape(n)
    end = np.array([-1 if r.end is None else r.end for r in records], np.int64).reshape(n)
    mapq = np.array([-1 if r.mapq is None else r.mapq for r in records], np.int32).reshape(n)
    flags = np.array([(1 if r.readMapped else 0) | (2 if r.readNegativeStrand else 0) |
                      (4 if r.cigar is not None else 0) | (8 if r.mismatchingPositions is not None else 0) |
                      (16 if r.mapq is not None else 0) | (32 if getattr(r, "primaryAlignment", True) else 0) |
                      (64 if getattr(r, "duplicateRead", False) else 0) for r in records], np.uint8).reshape(n)
    cigar, cigar_off = _blob([r.cigar for r in records])
    md, md_off = _blob([r.mismatchingPositions for r in records])
    seq, seq_off = _blob([r.sequence for r in records])
    qual, qual_off = _blob([r.qual for r in records])
    return dict(start=start, end=end, mapq=mapq, flags=flags, cigar=cigar, cigar_off=cigar_off, md=md,
                md_off=md_off, seq=seq, seq_off=seq_off, qual=qual, qual_off=qual_off)


def narrow_text_columns(cols):
    """The same text columns with 32-bit numeric columns (avo_text_reads.narrow): start / end int32, the
    offsets uint32, mapq uint8 (an absent mapq loses its flag bit instead of being -1).  23 bytes per read
    less to move when the text is packed on the device."""
    flags = np.array(cols["flags"], np.uint8, copy=True)
    mapq = np.asarray(cols["mapq"])
    flags[mapq < 0] &= np.uint8(0xFF ^ 16)
    out = dict(cols)
    out.update(narrow=True, flags=flags, start=np.asarray(cols["start"]).astype(np.int32),
               end=np.asarray(cols["end"]).astype(np.int32), mapq=np.clip(mapq, 0, 255).astype(np.uint8))
    shared = cols["qual_off"] is cols["seq_off"]
    for k in ("cigar_off", "md_off", "seq_off", "qual_off"):
        a = np.asarray(cols[k])
        if a.size and int(a[-1]) >= 1 << 32:
            raise ValueError("text blobs of 4 GB or more need the 64-bit columns")
        out[k] = a.astype(np.uint32)
    if shared:
        out["qual_off"] = out["seq_off"]
    return out


def pack_reads(records: Sequence[AlignmentRecord], contig_id=0, keep=None, sort=True) -> ReadBatch:
    """Packs mapped reads of ONE contig, (stably) sorted by start, into a host ReadBatch.
    ``keep`` optionally selects reads (``prefilter.prefilter_mask`` gives the PrefilterReads predicate)."""
    lib = L.load()
    idx = [i for i, r in enumerate(records) if r.readMapped and r.start is not None and
           (keep is None or keep[i])]
    if sort:
        idx.sort(key=lambda i: records[i].start)   # Python's sort is stable
    recs = [records[i] for i in idx]
    cols = text_columns(recs)
    b = pack_text_columns(cols, contig_id=contig_id)
    b.source_index = np.array([idx[j] for j in b.source_index], np.int64)
    return b


def pack_text_columns(cols, contig_id=0, threads=1, keep=None, out=None, prefilter=None) -> ReadBatch:
    """avo_pack_reads_mt over text columns (the dict ``text_columns`` / ``text_columns_of`` return):
    PrefilterReads' ``keep`` mask + extractAlignmentOperators + payload packing, on ``threads`` host
    threads.  The reads must already be those of one contig, sorted by start.  ``out`` optionally
    holds reusable output arrays {meta, payload, ops, refb, source_index} (e.g. page-locked)."""
    lib = L.load()
    t = L.TextReadsStruct()
    t.n = len(cols["start"])
    t.narrow = 1 if cols.get("narrow") else 0
    keepalive = []
    for name in ("start", "end", "mapq"):
        setattr(t, name, cols[name].ctypes.data)
    t.rec_flags = cols["flags"].ctypes.data
    for name in ("cigar", "md", "seq", "qual"):
        blob = cols[name]
        if isinstance(blob, (bytes, bytearray)):
            blob = np.frombuffer(bytes(blob) + b"\0", np.uint8)
        keepalive.append(blob)
        setattr(t, name, blob.ctypes.data)
        setattr(t, name + "_off", cols[name + "_off"].ctypes.data)
    if keep is not None:
        keep = np.ascontiguousarray(keep, np.uint8)
        t.keep = keep.ctypes.data
    else:
        t.keep = None
    if prefilter is not None:       # an L.PrefilterStruct: PrefilterReads' read predicates, evaluated by the packer
        t.prefilter = C.cast(C.pointer(prefilter), C.c_void_p)
    nr, nb, no, nf = (C.c_int64() for _ in range(4))
    rc = lib.avo_pack_bounds(C.byref(t), C.byref(nr), C.byref(nb), C.byref(no), C.byref(nf))
    if rc != 0:
        raise L.AvocadoError(rc, "avo_pack_bounds failed")
    arrs = out if out is not None else {}
    want = dict(meta=(max(nr.value, 1), META_DTYPE), payload=(max(nb.value, 1) * BLOCK_BYTES, np.uint8),
                ops=(max(no.value, 1), np.uint32), refb=(max(nf.value, 1), np.uint8),
                source_index=(max(nr.value, 1), np.int64))
    for k, (cnt, dt) in want.items():
        if k not in arrs or arrs[k].size < cnt:
            arrs[k] = np.empty(cnt, dt)
    out = L.PackedOutStruct()
    out.n_reads, out.n_blocks, out.n_ops, out.n_refb = (arrs["meta"].size, arrs["payload"].size // BLOCK_BYTES,
                                                        arrs["ops"].size, arrs["refb"].size)
    for k, a in arrs.items():
        setattr(out, k, a.ctypes.data)
    rc = lib.avo_pack_reads_mt(C.byref(t), C.byref(out), int(threads))
    if rc != 0:
        raise L.AvocadoError(rc, "avo_pack_reads failed")
    b = ReadBatch(out.n_reads, out.n_blocks, out.n_ops, out.n_refb, contig_id)
    b.meta = arrs["meta"][:out.n_reads]
    b.payload = arrs["payload"][:max(out.n_blocks, 1) * BLOCK_BYTES]
    b.ops = arrs["ops"][:max(out.n_ops, 1)]
    b.refb = arrs["refb"][:max(out.n_refb, 1)]
    b.source_index = arrs["source_index"][:out.n_reads]
    return b.with_stats()


def text_columns_of(batch: ReadBatch, first=0, n=None):
    """Text columns (CIGAR with M runs, MD, sequence, phred+33 qualities) of reads [first, first+n)
    of a packed host batch (avo_synth_text): bench / test infrastructure."""
    lib = L.load()
    n = batch.n_reads - first if n is None else n
    rb = batch.as_struct()
    out = L.TextOutStruct()
    rc = lib.avo_synth_text(C.byref(rb), first, n, C.byref(out))
    if rc != 0:
        raise L.AvocadoError(rc, "avo_synth_text(count) failed")
    cols = dict(start=np.empty(n, np.int64), end=np.empty(n, np.int64), mapq=np.empty(n, np.int32),
                flags=np.empty(n, np.uint8), cigar=np.empty(out.cigar_bytes + 1, np.uint8),
                cigar_off=np.empty(n + 1, np.int64), md=np.empty(out.md_bytes + 1, np.uint8),
                md_off=np.empty(n + 1, np.int64), seq=np.empty(out.seq_bytes + 1, np.uint8),
                qual=np.empty(out.seq_bytes + 1, np.uint8), seq_off=np.empty(n + 1, np.int64))
    for k in ("start", "end", "mapq", "cigar", "cigar_off", "md", "md_off", "seq", "qual", "seq_off"):
        setattr(out, k, cols[k].ctypes.data)
    out.rec_flags = cols["flags"].ctypes.data
    rc = lib.avo_synth_text(C.byref(rb), first, n, C.byref(out))
    if rc != 0:
        raise L.AvocadoError(rc, "avo_synth_text failed")
    cols["qual_off"] = cols["seq_off"]
    return cols


# ---------------------------------------------------------------------------------------------
# Synthetic generator wrappers (bench / test infrastructure)
# ---------------------------------------------------------------------------------------------
def synth_params(**kw) -> "L.SynthParamsStruct":
    lib = L.load()
    p = L.SynthParamsStruct()
    lib.avo_synth_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError("unknown synth parameter %r" % k)
        setattr(p, k, v)
    return p


def _packed_out(arrs, n, nb, no, nf):
    out = L.PackedOutStruct()
    out.n_reads, out.n_blocks, out.n_ops, out.n_refb = n, nb, no, nf
    for k, a in arrs.items():
        setattr(out, k, _ptr(a))
    out.source_index = None
    return out


def synth_host(p, first=0, n=None, contig_id=0) -> ReadBatch:
    lib = L.load()
    n = p.n_reads_total - first if n is None else n
    no, nf = C.c_int64(), C.c_int64()
    rc = lib.avo_synth_host(C.byref(p), first, n, C.byref(no), C.byref(nf), None)
    if rc != 0:
        raise L.AvocadoError(rc, "avo_synth_host(count) failed")
    nb = n * blocks_of(p.read_len)
    arrs = dict(meta=np.zeros(n, META_DTYPE), payload=np.zeros(max(nb, 1) * BLOCK_BYTES, np.uint8),
                ops=np.zeros(max(no.value, 1), np.uint32),
                refb=np.zeros(max(nf.value, 1), np.uint8))
    out = _packed_out(arrs, n, nb, no.value, nf.value)
    rc = lib.avo_synth_host(C.byref(p), first, n, None, None, C.byref(out))
    if rc != 0:
        raise L.AvocadoError(rc, "avo_synth_host failed")
    b = ReadBatch(n, nb, no.value, nf.value, contig_id)
    for k, a in arrs.items():
        setattr(b, k, a)
    return b.with_stats()


def synth_device(p, first=0, n=None, contig_id=0, device="cuda:0") -> ReadBatch:
    import torch
    lib = L.load()
    n = p.n_reads_total - first if n is None else n
    with torch.cuda.device(device):
        op_off = _device_empty(n + 1, np.uint32, device)
        refb_off = _device_empty(n + 1, np.uint32, device)
        no, nf = C.c_int64(), C.c_int64()
        stream = torch.cuda.current_stream().cuda_stream
        rc = lib.avo_synth_device(C.byref(p), first, n, op_off.data_ptr(), refb_off.data_ptr(),
                                  C.byref(no), C.byref(nf), None, stream)
        if rc != 0:
            raise L.AvocadoError(rc, "avo_synth_device(count) failed")
        nb = n * blocks_of(p.read_len)
        arrs = dict(meta=_device_empty(n * 32, np.uint8, device),
                    payload=_device_empty(nb * BLOCK_BYTES, np.uint8, device),
                    ops=_device_empty(no.value, np.uint32, device),
                    refb=_device_empty(nf.value, np.uint8, device))
        out = _packed_out(arrs, n, nb, no.value, nf.value)
        rc = lib.avo_synth_device(C.byref(p), first, n, op_off.data_ptr(), refb_off.data_ptr(), None,
                                  None, C.byref(out), stream)
        if rc != 0:
            raise L.AvocadoError(rc, "avo_synth_device failed")
    b = ReadBatch(n, nb, no.value, nf.value, contig_id)
    for k, a in arrs.items():
        setattr(b, k, a)
    return b.with_stats()
