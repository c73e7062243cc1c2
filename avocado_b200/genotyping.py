"""Host-side mirror of the reference's entry points for the hot path, over the C ABI.

    DiscoverVariants(reads, optPhredThreshold, optMinObservations)
        -> CORE/genotyping/DiscoverVariants.scala:51-61
    BiallelicGenotyper.call(reads, variants, copyNumber, scoreAllSites, maxQuality, maxMapQ)
        -> CORE/genotyping/BiallelicGenotyper.scala:88-135
    BiallelicGenotyper.discoverAndCall(...)            -> BG:156-184

(The reference's host language is Scala; no JVM exists in this image, so this mirror is Python
and the JNI shim a maintainer would add is in INTEGRATION.md.)  Names, argument meaning and error
behaviour follow the reference: exactly one sample is required, an empty variant set raises,
malformed reads are skipped.  All compute happens in libavocado_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from .batch import ReadBatch, SiteTable, pack_reads, sites_from_variants, _ptr, _device_empty
from .records import ALT, OTHER_ALT, REF, AlignmentRecord, Genotype, Variant


@dataclass
class CopyNumberMap:
    """CORE/models/CopyNumberMap.scala:75-111: base ploidy + CNVs as (contig, start, end, copyNumber)."""
    basePloidy: int = 2
    cnvs: List[Tuple[str, int, int, int]] = field(default_factory=list)

    @staticmethod
    def empty(basePloidy: int) -> "CopyNumberMap":
        return CopyNumberMap(basePloidy, [])

    @staticmethod
    def from_features(basePloidy: int, features: Sequence[Tuple[str, int, int, str]]) -> "CopyNumberMap":
        """features: (contig, start, end, featureType); DUP -> +1, DEL -> -1 (CopyNumberMap.scala:45-64)"""
        out = []
        for c, s, e, t in features:
            if t == "DUP":
                out.append((c, s, e, basePloidy + 1))
            elif t == "DEL":
                out.append((c, s, e, basePloidy - 1))
        return CopyNumberMap(basePloidy, out)

    @property
    def minPloidy(self):
        return min([self.basePloidy] + [c[3] for c in self.cnvs])

    @property
    def maxPloidy(self):
        return max([self.basePloidy] + [c[3] for c in self.cnvs])

    @property
    def variantsByContig(self) -> Dict[str, List[Tuple[int, int, int]]]:
        """contig -> [(start, end, copyNumber)] sorted by region (CopyNumberMap.scala:56-61)."""
        out: Dict[str, List[Tuple[int, int, int]]] = {}
        for c, s, e, n in self.cnvs:
            out.setdefault(c, []).append((s, e, n))
        return {c: sorted(v, key=lambda t: (t[0], t[1])) for c, v in out.items()}

    def overlappingVariants(self, contig: str, start: int, end: int) -> List[Tuple[int, int, int]]:
        """CopyNumberMap.scala:103-111: dropWhile(!overlaps).takeWhile(overlaps) over the contig's sorted list."""
        out, seen = [], False
        for s, e, n in self.variantsByContig.get(contig, []):
            hit = e > start and s < end
            if hit:
                out.append((s, e, n))
                seen = True
            elif seen:
                break
        return out


class Context:
    """One avo_ctx: bound to one GPU, not thread-safe."""

    def __init__(self, device: int = 0):
        self.lib = L.load()
        h = C.c_void_p()
        rc = self.lib.avo_ctx_create(device, C.byref(h))
        if rc != 0:
            raise L.AvocadoError(rc, "avo_ctx_create(device=%d) failed: a CUDA device is required "
                                     "(there is no CPU fallback)" % device)
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.avo_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise L.AvocadoError(rc, (self.lib.avo_last_error(self.h) or b"").decode())

    def set_stream(self, cuda_stream: Optional[int]):
        self._check(self.lib.avo_ctx_set_stream(self.h, cuda_stream))

    def timing(self) -> dict:
        t = L.TimingStruct()
        self.lib.avo_get_timing(self.h, C.byref(t))
        return {k: getattr(t, k) for k, _ in L.TimingStruct._fields_}

    # ---- discovery's candidate events ----------------------------------------------------------------
    def derive_events(self, batch: ReadBatch) -> ReadBatch:
        """avo_derive_events on this context's GPU (device batches) or on the host (host batches)."""
        return derive_events(batch, self)

    # ---- the packer on the device ---------------------------------------------------------------
    def pack_text_device(self, cols, contig_id=0, keep=None, out=None, source_index=False, prefilter=None) -> ReadBatch:
        """avo_pack_reads_device: text columns (``batch.text_columns`` layout; numpy arrays, ideally
        page-locked, or torch device tensors) -> a packed ReadBatch in device memory.  ``out``
        optionally supplies reusable device buffers {meta, payload, ops, refb} of sufficient size.
        ``prefilter``: an ``L.PrefilterStruct`` (``prefilter.prefilter_struct``): PrefilterReads' read
        predicates, evaluated by the packer kernels."""
        import torch
        on_dev = not isinstance(cols["start"], np.ndarray)
        n = int(cols["start"].shape[0])
        t = L.TextReadsStruct()
        t.n = n
        hold = []

        def col(name, dtype):
            a = cols[name]
            if isinstance(a, (bytes, bytearray)):
                a = np.frombuffer(bytes(a) + b"\0", np.uint8)
            hold.append(a)
            return _ptr(a)

        def last(name):
            a = cols[name]
            return int(a[n]) if n or len(a) else 0

        t.narrow = 1 if cols.get("narrow") else 0       # batch.narrow_text_columns: 32-bit numeric columns
        t.start, t.end, t.mapq = col("start", np.int64), col("end", np.int64), col("mapq", np.int32)
        t.rec_flags = col("flags", np.uint8)
        for name in ("cigar", "md", "seq", "qual"):
            setattr(t, name, col(name, np.uint8))
            setattr(t, name + "_off", col(name + "_off", np.int64))
        if keep is not None:
            keep = keep if not isinstance(keep, np.ndarray) else np.ascontiguousarray(keep, np.uint8)
            hold.append(keep)
            t.keep = _ptr(keep)
        else:
            t.keep = None
        if prefilter is not None:
            hold.append(prefilter)
            t.prefilter = C.cast(C.pointer(prefilter), C.c_void_p)
        # upper bounds from the blob sizes (what avo_pack_bounds sums read by read)
        cb, mb, sb = last("cigar_off"), last("md_off"), last("seq_off")
        cap = dict(meta=n, payload=sb // 10 + n, ops=cb // 2 + n + 2 * mb, refb=mb + n)
        dev = "cuda:%d" % self.device
        if out is None:
            out = {}
        with torch.cuda.device(dev):
            for k, unit, dt in (("meta", 32, np.uint8), ("payload", 16, np.uint8), ("ops", 1, np.uint32), ("refb", 1, np.uint8)):
                if k not in out or out[k].numel() < max(cap[k] * unit, 1):
                    out[k] = _device_empty(cap[k] * unit, dt, dev)
            if source_index and ("source_index" not in out or out["source_index"].numel() < max(n, 1)):
                out["source_index"] = _device_empty(n, np.int64, dev)
        po = L.PackedOutStruct()
        po.n_reads, po.n_blocks, po.n_ops, po.n_refb = (out["meta"].numel() // 32, out["payload"].numel() // 16,
                                                        out["ops"].numel(), out["refb"].numel())
        po.meta, po.payload, po.ops, po.refb = (out[k].data_ptr() for k in ("meta", "payload", "ops", "refb"))
        po.source_index = out["source_index"].data_ptr() if source_index else None
        self._check(self.lib.avo_pack_reads_device(self.h, C.byref(t), L.MEM_DEVICE if on_dev else L.MEM_HOST,
                                                   C.byref(po)))
        b = ReadBatch(po.n_reads, po.n_blocks, po.n_ops, po.n_refb, contig_id)
        b.meta = out["meta"][:max(po.n_reads, 1) * 32]
        b.payload = out["payload"][:max(po.n_blocks, 1) * 16]
        b.ops = out["ops"][:max(po.n_ops, 1)]
        b.refb = out["refb"][:max(po.n_refb, 1)]
        if source_index:
            b.source_index = out["source_index"][:po.n_reads]
        return b

    # ---- raw entry points over packed data ------------------------------------------------------
    host_payload = 0    # avo_params.host_payload: 0 = pinned host payload read in place, 1 = copy,
    #                     2 = read in place by discovery, gathered by host threads for the call stage
    gather_threads = 0  # avo_params.gather_threads (0 = the library's default)

    def _params(self, ploidy=2, max_quality=93, max_mapq=93, phred=0, min_obs=None, score_all=False):
        p = L.ParamsStruct()
        p.host_payload = self.host_payload
        p.gather_threads = self.gather_threads
        p.ploidy, p.score_all_sites = ploidy, int(bool(score_all))
        p.max_quality, p.max_mapq = max_quality, max_mapq
        p.min_phred_discover = phred
        p.min_obs_discover = -1 if min_obs is None else min_obs
        return p

    @staticmethod
    def _cnv(cnvs, contig_rank):
        if not cnvs:
            return None, None
        # contigs the caller did not rank can never match a batch; they still count towards
        # min/max ploidy (CopyNumberMap.scala:82-95)
        rows = sorted((contig_rank.get(c, 1 << 30), s, e, cn) for c, s, e, cn in cnvs)
        a = np.array(rows, np.int32).reshape(-1, 4)
        cols = [np.ascontiguousarray(a[:, k]) for k in range(4)]
        s = L.CnvStruct()
        s.n = len(rows)
        s.contig, s.start, s.end, s.copy_number = (c.ctypes.data for c in cols)
        return s, cols

    def _alloc_results(self, n, ns, device=None):
        res = L.SiteResultsStruct()
        res.n_sites, res.n_states = n, ns
        res.mem = L.MEM_DEVICE if device is not None else L.MEM_HOST
        arrs = {}
        for name, dt, w in L.RESULT_COLUMNS:
            width = ns if w == "ns" else w
            if device is None:
                a = np.zeros(max(n * width, 1), np.dtype(dt))
            else:
                a = _device_empty(max(n * width, 1), np.dtype(dt).type, device)
            arrs[name] = a
            setattr(res, name, _ptr(a))
        return res, arrs

    @staticmethod
    def _shape_results(arrs, n, ns):
        out = {}
        for name, dt, w in L.RESULT_COLUMNS:
            width = ns if w == "ns" else w
            a = arrs[name][:n * width]
            out[name] = a.reshape(n, width) if width != 1 else a
        return out

    def discover_packed(self, batch: ReadBatch, phred=0, min_obs=None, capacity=None,
                        device_out=False) -> SiteTable:
        """avo_discover on a packed batch -> sorted SiteTable (host arrays unless device_out)."""
        p = self._params(phred=phred, min_obs=min_obs)
        bs = batch.as_struct()
        cap = capacity or 1 << 16
        acap = 8 * cap
        for _ in range(3):
            out = L.SitesOutStruct()
            out.capacity, out.allele_capacity = cap, acap
            dev = "cuda:%d" % self.device if device_out else None
            mk = (lambda n, dt: _device_empty(n, dt, dev)) if device_out else (lambda n, dt: np.zeros(n, dt))
            arrs = dict(start=mk(cap, np.int32), ref_len=mk(cap, np.int32), alt_len=mk(cap, np.int32),
                        allele_off=mk(cap + 1, np.uint32), alleles4=mk(acap // 2 + 1, np.uint8),
                        count=mk(cap, np.int32))
            out.mem = L.MEM_DEVICE if device_out else L.MEM_HOST
            for k, a in arrs.items():
                setattr(out, k, _ptr(a))
            rc = self.lib.avo_discover(self.h, C.byref(bs), C.byref(p), C.byref(out))
            if rc == L.AVO_E_CAPACITY:
                cap, acap = out.n + 16, out.n_allele_nibbles + 64
                continue
            self._check(rc)
            n, nn = out.n, out.n_allele_nibbles
            return SiteTable(n, nn, arrs["start"][:max(n, 1)][:n], arrs["ref_len"][:n], arrs["alt_len"][:n],
                             arrs["allele_off"][:n + 1], arrs["alleles4"][:(nn + 1) // 2 or 1], None,
                             arrs["count"][:n])
        raise L.AvocadoError(L.AVO_E_CAPACITY, "discover output kept growing")

    def call_packed(self, batch: ReadBatch, sites: SiteTable, ploidy=2, cnvs=None, contig_rank=None,
                    max_quality=93, max_mapq=93, device_out=False) -> Dict[str, np.ndarray]:
        """avo_call -> dict of per-site columns (RESULT_COLUMNS), rows aligned with `sites`."""
        p = self._params(ploidy=ploidy, max_quality=max_quality, max_mapq=max_mapq)
        cn = CopyNumberMap(ploidy, list(cnvs or []))
        ns = cn.maxPloidy + 1
        cnv_s, keep = self._cnv(cn.cnvs, contig_rank or {})
        res, arrs = self._alloc_results(sites.n, ns, "cuda:%d" % self.device if device_out else None)
        bs, ss = batch.as_struct(), sites.as_struct()
        rc = self.lib.avo_call(self.h, C.byref(bs), C.byref(ss), C.byref(cnv_s) if cnv_s else None,
                               C.byref(p), C.byref(res))
        self._check(rc)
        return self._shape_results(arrs, sites.n, ns) if not device_out else arrs

    def call_all_sites_packed(self, batch: ReadBatch, sites: SiteTable, ploidy=2, cnvs=None,
                              contig_rank=None, max_quality=93, max_mapq=93, capacity=None,
                              device_out=False):
        """avo_call_all_sites (scoreAllSites / gVCF mode) -> (variant columns, ref-model start
        positions, ref-model columns)."""
        p = self._params(ploidy=ploidy, max_quality=max_quality, max_mapq=max_mapq, score_all=True)
        cn = CopyNumberMap(ploidy, list(cnvs or []))
        ns = cn.maxPloidy + 1
        cnv_s, keep = self._cnv(cn.cnvs, contig_rank or {})
        dev = "cuda:%d" % self.device if device_out else None
        res, arrs = self._alloc_results(sites.n, ns, dev)
        bs, ss = batch.as_struct(), sites.as_struct()
        cap = capacity or 1 << 16      # two-phase: a too small guess costs one planning pass
        for _ in range(3):
            rres, rarrs = self._alloc_results(cap, ns, dev)
            start = _device_empty(cap, np.int32, dev) if device_out else np.zeros(max(cap, 1), np.int32)
            ro = L.RefModelOutStruct()
            ro.capacity, ro.start, ro.rows = cap, _ptr(start), rres
            rc = self.lib.avo_call_all_sites(self.h, C.byref(bs), C.byref(ss), C.byref(cnv_s) if cnv_s else None,
                                             C.byref(p), C.byref(res), C.byref(ro))
            if rc == L.AVO_E_CAPACITY:
                cap = ro.n + 16
                continue
            self._check(rc)
            n = ro.n
            cols = self._shape_results(arrs, sites.n, ns) if not device_out else arrs
            rcols = self._shape_results(rarrs, n, ns) if not device_out else rarrs
            return cols, start[:n], rcols
        raise L.AvocadoError(L.AVO_E_CAPACITY, "ref-model output kept growing")

    def discover_and_call_packed(self, batch: ReadBatch, ploidy=2, phred=0, min_obs=None, cnvs=None,
                                 contig_rank=None, max_quality=93, max_mapq=93, capacity=None,
                                 device_out=False, copy=True):
        """avo_discover_and_call -> (SiteTable, result columns).  Output buffers are cached in the
        context and reused by the next call; with copy=False the returned arrays are views of them."""
        p = self._params(ploidy=ploidy, max_quality=max_quality, max_mapq=max_mapq, phred=phred,
                         min_obs=min_obs)
        cn = CopyNumberMap(ploidy, list(cnvs or []))
        ns = cn.maxPloidy + 1
        cnv_s, keep = self._cnv(cn.cnvs, contig_rank or {})
        bs = batch.as_struct()
        cap = capacity or 1 << 16
        acap = 8 * cap
        dev = "cuda:%d" % self.device if device_out else None
        mk = (lambda n, dt: _device_empty(n, dt, dev)) if device_out else (lambda n, dt: np.zeros(n, dt))
        for _ in range(3):
            # output buffers are cached across calls (a shim would reuse its direct buffers too)
            key = (cap, acap, ns, bool(device_out))
            caches = self.__dict__.setdefault("_dc_cache", {})
            if key not in caches:
                sarr = dict(start=mk(cap, np.int32), ref_len=mk(cap, np.int32), alt_len=mk(cap, np.int32),
                            allele_off=mk(cap + 1, np.uint32), alleles4=mk(acap // 2 + 1, np.uint8),
                            count=mk(cap, np.int32))
                res, arrs = self._alloc_results(cap, ns, dev)
                for k in [k for k in caches if k[3] == key[3]]:
                    del caches[k]          # keep one host and one device set
                caches[key] = (sarr, res, arrs)
            sarr, res, arrs = caches[key]
            out = L.SitesOutStruct()
            out.capacity, out.allele_capacity = cap, acap
            out.mem = L.MEM_DEVICE if device_out else L.MEM_HOST
            for k, a in sarr.items():
                setattr(out, k, _ptr(a))
            res.n_sites = cap
            rc = self.lib.avo_discover_and_call(self.h, C.byref(bs), C.byref(cnv_s) if cnv_s else None,
                                                C.byref(p), C.byref(out), C.byref(res))
            if rc == L.AVO_E_CAPACITY:
                cap, acap = max(out.n, res.n_sites) + 16, out.n_allele_nibbles + 64
                continue
            self._check(rc)
            n, nn = out.n, out.n_allele_nibbles
            # the views of the cached buffers only depend on (n, nn): reuse them when a call returns
            # as many rows as the one before (slicing ~25 tensors costs more than the call's launches)
            views = self.__dict__.setdefault("_dc_views", {})
            if not copy and views.get("key") == (key, n, nn):
                return views["sites"], views["cols"]
            sites = SiteTable(n, nn, sarr["start"][:n], sarr["ref_len"][:n], sarr["alt_len"][:n],
                              sarr["allele_off"][:n + 1], sarr["alleles4"][:(nn + 1) // 2 or 1], None,
                              sarr["count"][:n])
            cols = {}
            for name, dt, w in L.RESULT_COLUMNS:
                width = ns if w == "ns" else w
                a = arrs[name][:n * width]
                cols[name] = a.reshape(n, width) if width != 1 else a
            if not copy:
                views.update(key=(key, n, nn), sites=sites, cols=cols)
            if copy and not device_out:
                cols = {k: v.copy() for k, v in cols.items()}
                sites = SiteTable(n, nn, sites.start.copy(), sites.ref_len.copy(), sites.alt_len.copy(),
                                  sites.allele_off.copy(), sites.alleles4.copy(), None, sites.count.copy())
            return sites, cols
        raise L.AvocadoError(L.AVO_E_CAPACITY, "discover_and_call output kept growing")


def derive_events(batch: ReadBatch, ctx: "Optional[Context]" = None) -> ReadBatch:
    """avo_derive_events: attaches the candidate events (every mismatching base, every I / D operator after
    the first aligned one, in read order) to ``batch`` -- in its own memory -- and returns it.  A DEVICE
    batch that carries them is discovered by counting them (k_discover_events) instead of walking its reads;
    packers call this once per batch.  Host batches need no context.  Raises AVO_E_UNSORTED for an unsorted
    batch."""
    lib = L.load()
    if batch.on_device and ctx is None:
        raise ValueError("a device batch needs the Context of its GPU")
    n_w = (batch.n_reads + 31) // 32
    dev = ("cuda:%d" % ctx.device) if batch.on_device else None
    mk = (lambda n, dt: _device_empty(max(n, 1), dt, dev)) if dev else (lambda n, dt: np.zeros(max(n, 1), dt))
    caps = (batch.n_refb, batch.n_ops)
    for _ in range(2):
        arrs = dict(snv_pos=mk(caps[0], np.int32), snv_pq=mk(caps[0], np.uint16), snv_first=mk(n_w + 1, np.uint32),
                    indel_read=mk(caps[1], np.int32), indel_op=mk(caps[1], np.uint32))
        e = L.ReadEventsStruct()
        e.n_snv, e.n_indel = caps
        for k, a in arrs.items():
            setattr(e, k, _ptr(a))
        saved, batch.events = batch.events, None
        try:
            bs = batch.as_struct()
        finally:
            batch.events = saved
        rc = lib.avo_derive_events(ctx.h if (ctx is not None and batch.on_device) else None, C.byref(bs), C.byref(e))
        if rc == L.AVO_E_CAPACITY:
            caps = (int(e.n_snv), int(e.n_indel))
            continue
        if rc != 0:
            msg = (lib.avo_last_error(ctx.h) or b"").decode() if (ctx is not None and batch.on_device) else "avo_derive_events failed"
            raise L.AvocadoError(rc, msg)
        ns, ni = int(e.n_snv), int(e.n_indel)
        batch.events = dict(snv_pos=arrs["snv_pos"][:max(ns, 1)], snv_pq=arrs["snv_pq"][:max(ns, 1)],
                            snv_first=arrs["snv_first"], indel_read=arrs["indel_read"][:max(ni, 1)],
                            indel_op=arrs["indel_op"][:max(ni, 1)], n_snv=ns, n_indel=ni,
                            stats=(e.stats_min_start, e.stats_max_end, e.stats_max_span))
        return batch
    raise L.AvocadoError(L.AVO_E_CAPACITY, "avo_derive_events kept asking for more room")


_default_ctx: Dict[int, Context] = {}


def default_context(device=0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


# ---------------------------------------------------------------------------------------------
# Record-level API (the reference's signatures)
# ---------------------------------------------------------------------------------------------
def _by_contig(reads: Sequence[AlignmentRecord]):
    groups: Dict[str, List[AlignmentRecord]] = {}
    for r in reads:
        if r.readMapped and r.contigName is not None and r.start is not None:
            groups.setdefault(r.contigName, []).append(r)
    return groups


def _contig_ranks(names) -> Dict[str, int]:
    # ReferenceRegion ordering compares contig names as strings (ADAM RegionOrdering)
    return {c: i for i, c in enumerate(sorted(set(names)))}


def DiscoverVariants(reads: Sequence[AlignmentRecord], optPhredThreshold: Optional[int] = None,
                     optMinObservations: Optional[int] = None, ctx: Optional[Context] = None
                     ) -> List[Variant]:
    """DiscoverVariants.apply (DiscoverVariants.scala:51-61, 71-102)."""
    ctx = ctx or default_context()
    out: List[Variant] = []
    groups = _by_contig(reads)
    for contig in sorted(groups):
        batch = pack_reads(groups[contig])
        sites = ctx.discover_packed(batch, phred=optPhredThreshold or 0, min_obs=optMinObservations)
        out.extend(sites.to_variants([contig]))
    return out


def _genotypes_from_columns(sites: SiteTable, cols, variants: Sequence[Variant], sample: Optional[str]
                            ) -> List[Genotype]:
    out = []
    for i in range(sites.n):
        if not cols["emitted"][i]:
            continue      # the reference only emits sites that received an observation (BG:475-500)
        v = variants[i]
        cn = int(cols["copy_number"][i])
        alleles = [ALT] * int(cols["n_alt"][i]) + [REF] * int(cols["n_ref"][i]) + \
                  [OTHER_ALT] * int(cols["n_other"][i])                                 # BG:691-697
        obs = {k: (cols[k][i].tolist() if hasattr(cols[k][i], "tolist") else cols[k][i])
               for k in ("allele_fwd", "other_fwd", "square_mapq", "ref_ll", "allele_ll", "other_ll",
                         "nonref_ll", "allele_cov", "other_cov", "total_cov", "first_is_ref",
                         "copy_number", "qual")}
        out.append(Genotype(
            variant=v, contigName=v.contigName, start=v.start, end=v.end, sampleId=sample,
            alleles=alleles, genotypeQuality=int(cols["gq"][i]),
            genotypeLikelihoods=cols["gl"][i][:cn + 1].tolist(),
            nonReferenceLikelihoods=cols["nonref_gl"][i][:cn + 1].tolist(),
            strandBiasComponents=cols["sb"][i].tolist(), readDepth=int(cols["total_cov"][i]),
            referenceReadDepth=int(cols["other_cov"][i]), alternateReadDepth=int(cols["allele_cov"][i]),
            rmsMapQ=float(cols["rms_mapq"][i]), fisherStrandBiasPValue=float(cols["fisher"][i]),
            observation=obs))
    return out


def _one_sample(reads, samples):
    if samples is None:
        samples = sorted({r.recordGroupSample for r in reads if r.recordGroupSample is not None}) or [None]
    if len(set(samples)) != 1:                                                          # BG:102-105
        raise ValueError("Currently, we only support a single sample. Saw: %s." %
                         ", ".join(map(str, samples)))
    return list(samples)[0]


class BiallelicGenotyper:
    """CORE/genotyping/BiallelicGenotyper.scala"""

    @staticmethod
    def call(reads: Sequence[AlignmentRecord], variants: Sequence[Variant],
             copyNumber: Optional[CopyNumberMap] = None, scoreAllSites: bool = False,
             maxQuality: int = 93, maxMapQ: int = 93, samples=None, ctx: Optional[Context] = None
             ) -> List[Genotype]:
        ctx = ctx or default_context()
        copyNumber = copyNumber or CopyNumberMap.empty(2)
        sample = _one_sample(reads, samples)
        if len(variants) == 0:
            raise L.AvocadoError(L.AVO_E_EMPTY_SITES, "empty variant set: the reference throws "
                                 "ArrayIndexOutOfBounds at TreeRegionJoin.scala:77")
        groups = _by_contig(reads)
        ranks = _contig_ranks([v.contigName for v in variants] + list(groups))
        table = sites_from_variants(list(variants), ranks)
        sorted_variants = [variants[i] for i in table.order]
        out: List[Genotype] = []
        merged = None
        ref_model: List[Genotype] = []
        for contig in sorted(groups):
            batch = pack_reads(groups[contig], contig_id=ranks[contig])
            if scoreAllSites:                                                       # BG:223-275
                cols, rstart, rcols = ctx.call_all_sites_packed(
                    batch, table, ploidy=copyNumber.basePloidy, cnvs=copyNumber.cnvs, contig_rank=ranks,
                    max_quality=maxQuality, max_mapq=maxMapQ)
                rv = [Variant(contig, int(s), int(s) + 1, "N", None) for s in rstart]
                rt = SiteTable(len(rv), 0, rstart, None, None, None, None)
                ref_model.extend(_genotypes_from_columns(rt, rcols, rv, sample))
            else:
                cols = ctx.call_packed(batch, table, ploidy=copyNumber.basePloidy, cnvs=copyNumber.cnvs,
                                       contig_rank=ranks, max_quality=maxQuality, max_mapq=maxMapQ)
            # a site belongs to one contig, so per-contig results never overlap
            if merged is None:
                merged = {k: v.copy() for k, v in cols.items()}
            else:
                m = cols["emitted"].astype(bool)
                for k in merged:
                    merged[k][m] = cols[k][m]
        if merged is None:
            return []
        return ref_model + _genotypes_from_columns(table, merged, sorted_variants, sample)

    @staticmethod
    def discoverAndCall(reads: Sequence[AlignmentRecord], copyNumber: Optional[CopyNumberMap] = None,
                        scoreAllSites: bool = False, optPhredThreshold: Optional[int] = None,
                        optMinObservations: Optional[int] = None, maxQuality: int = 93,
                        maxMapQ: int = 93, samples=None, ctx: Optional[Context] = None) -> List[Genotype]:
        """BG:156-184.  With one contig the fused avo_discover_and_call keeps the discovered
        sites on the device; with several contigs discovery runs per contig and call() then sees
        the global table, as the reference's Forest does."""
        ctx = ctx or default_context()
        copyNumber = copyNumber or CopyNumberMap.empty(2)
        sample = _one_sample(reads, samples)
        groups = _by_contig(reads)
        if len(groups) == 1 and not scoreAllSites:
            (contig, recs), = groups.items()
            ranks = {contig: 0}
            batch = pack_reads(recs, contig_id=0)
            sites, cols = ctx.discover_and_call_packed(
                batch, ploidy=copyNumber.basePloidy, phred=optPhredThreshold or 0,
                min_obs=optMinObservations, cnvs=copyNumber.cnvs, contig_rank=ranks,
                max_quality=maxQuality, max_mapq=maxMapQ)
            return _genotypes_from_columns(sites, cols, sites.to_variants([contig]), sample)
        variants = DiscoverVariants(reads, optPhredThreshold, optMinObservations, ctx=ctx)
        return BiallelicGenotyper.call(reads, variants, copyNumber, scoreAllSites, maxQuality, maxMapQ,
                                       samples=[sample], ctx=ctx)
