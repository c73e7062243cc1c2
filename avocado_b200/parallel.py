"""Multi-GPU form of the hot path: one process per GPU, work sharded by genomic region
(SURVEY.md 8e).  Nothing on the data path is exchanged between ranks except

  * the discovered site tables (all-gather, so that every rank joins its reads against the
    GLOBAL sorted table, which the Forest look-up of TreeRegionJoin.scala:74-119 depends on), and
  * the per-site result rows of the sites each rank owns (all-gather to every rank).

Rank g owns the sites that START in [lo_g, hi_g) and is handed every read that can overlap one of
them or contribute a discovery candidate to one of them (read-overlap halo), so its candidate
counts and per-site sums are complete and no cross-rank reduction exists.  The reads of a rank
are a contiguous slice of the sorted batch, so per-site sums keep the batch's read order.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .batch import ReadBatch, SiteTable


@dataclass
class RegionShard:
    rank: int
    lo: int          # owned sites start in [lo, hi)
    hi: int
    read_lo: int     # reads [read_lo, read_hi) of the sorted batch
    read_hi: int
    max_site_len: int = 1 << 12


def plan_region_shards(start: np.ndarray, end: np.ndarray, world: int, max_site_len: int = 1 << 12
                       ) -> List[RegionShard]:
    """Cuts a sorted batch into `world` shards of (nearly) equal read count; boundaries are read
    starts.  Rank g owns the sites starting in [lo_g, hi_g) and gets the reads with
    lo_g - max_span <= start < hi_g + max_site_len: everything that can overlap an owned site or
    put a discovery candidate on an owned position.  `max_site_len` bounds the reference length of
    a site (checked against the discovered table by discover_and_call_sharded)."""
    n = len(start)
    start = np.asarray(start, np.int64)
    span = int((np.asarray(end, np.int64) - start).max()) if n else 0
    NEG, POS = -(1 << 40), 1 << 40
    cuts = [NEG] + [int(start[(g * n) // world]) if n else 0 for g in range(1, world)] + [POS]
    for g in range(2, world):
        cuts[g] = max(cuts[g], cuts[g - 1])
    shards = []
    for g in range(world):
        lo, hi = cuts[g], cuts[g + 1]
        r_lo = 0 if g == 0 else int(np.searchsorted(start, lo - span, side="left"))
        r_hi = n if g + 1 == world else int(np.searchsorted(start, hi + max_site_len, side="left"))
        shards.append(RegionShard(g, lo, hi, r_lo, max(r_hi, r_lo), max_site_len))
    return shards


def _cat_sites(tables: Sequence[SiteTable]) -> Tuple[SiteTable, List[int]]:
    """Concatenates per-rank (already position-ordered, disjoint) site tables; returns the merged
    table and the row offset of every part."""
    offs, rows, nibs = [], 0, 0
    st, rl, al, ao, a4, ct = [], [], [], [], [], []
    for t in tables:
        offs.append(rows)
        st.append(np.asarray(t.start[:t.n], np.int32)); rl.append(np.asarray(t.ref_len[:t.n], np.int32))
        al.append(np.asarray(t.alt_len[:t.n], np.int32))
        ao.append(np.asarray(t.allele_off[:t.n], np.uint32) + np.uint32(nibs))
        assert t.n_allele_nibbles % 2 == 0
        a4.append(np.asarray(t.alleles4[:t.n_allele_nibbles // 2], np.uint8))
        ct.append(np.asarray(t.count[:t.n], np.int32) if t.count is not None else np.zeros(t.n, np.int32))
        rows += t.n
        nibs += t.n_allele_nibbles
    ao.append(np.array([nibs], np.uint32))
    cat = lambda xs, dt: np.ascontiguousarray(np.concatenate(xs)) if xs else np.zeros(0, dt)
    alle = cat(a4, np.uint8)
    if alle.size == 0:
        alle = np.zeros(1, np.uint8)
    merged = SiteTable(rows, nibs, cat(st, np.int32), cat(rl, np.int32), cat(al, np.int32),
                       cat(ao, np.uint32), alle, None, cat(ct, np.int32))
    return merged, offs


def _take_sites(t: SiteTable, keep: np.ndarray) -> SiteTable:
    """Rows `keep` (a boolean mask) of a host site table, alleles repacked."""
    idx = np.nonzero(keep)[0]
    nibs: List[int] = []
    ao = np.zeros(len(idx) + 1, np.uint32)
    for k, i in enumerate(idx):
        off, ln = int(t.allele_off[i]), int(t.ref_len[i]) + int(t.alt_len[i])
        ao[k] = len(nibs)
        for j in range(ln):
            b = int(t.alleles4[(off + j) >> 1])
            nibs.append((b & 15) if (off + j) & 1 else (b >> 4))
        if len(nibs) & 1:
            nibs.append(0)
    ao[len(idx)] = len(nibs)
    a = np.array(nibs, np.uint8) if nibs else np.zeros(0, np.uint8)
    packed = ((a[0::2] << 4) | a[1::2]).astype(np.uint8) if len(a) else np.zeros(1, np.uint8)
    return SiteTable(len(idx), len(nibs), np.ascontiguousarray(t.start[idx]), np.ascontiguousarray(t.ref_len[idx]),
                     np.ascontiguousarray(t.alt_len[idx]), ao, packed, None,
                     None if t.count is None else np.ascontiguousarray(t.count[idx]))


def _all_gather_objects(obj, group=None):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [obj]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, obj, group=group)
    return out


def _all_gather_rows(x, device=None, group=None):
    """All-gather of a per-rank array with a different number of rows on every rank (padded to the
    maximum, one collective per column; NCCL when `device` is a CUDA device, else the group's CPU
    backend).  Returns the list of per-rank numpy arrays."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [np.asarray(x)]
    world = dist.get_world_size(group)
    x = np.ascontiguousarray(x)
    view = x.view(np.int32) if x.dtype == np.uint32 else x
    t = torch.from_numpy(view)
    if device is not None:
        t = t.to(device)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    parts = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return [p[:s].cpu().numpy().view(x.dtype) for p, s in zip(parts, sizes)]


class ShardedPipeline:
    """Region-sharded discoverAndCall with everything on the device (the C ABI's avo_shard_discover /
    avo_shard_call[_direct] / avo_shard_finish): an NCCL all-gather of fixed-size blocks holding the owned
    slice of every rank's site table, then the result rows of the owned sites -- as a second all-gather
    (rows_to="all"), a gather to rank 0 ("root"), or stored by every rank's last kernel straight into rank
    0's columns over NVLink peer memory ("direct").  The host waits once per step, in avo_shard_finish.  `gather(out, inp)` is the collective: torch.distributed's all_gather_into_tensor over
    NCCL by default; tests emulate several ranks on one GPU by passing their own."""

    def __init__(self, ctx, rank: int, world: int, own_lo: int, own_hi: int, ploidy: int = 2, site_cap: int = 1 << 15,
                 allele_cap: int = 1 << 17, global_cap: Optional[int] = None, gather=None, group=None, stream=None,
                 rows_to: str = "all", direct_root=None):
        import torch
        from . import _lib as L
        self.ctx, self.L, self.torch = ctx, L, torch
        self.dev = "cuda:%d" % ctx.device
        self.plan = L.ShardPlanStruct()
        self.plan.rank, self.plan.world, self.plan.own_lo, self.plan.own_hi = rank, world, own_lo, own_hi
        self.plan.n_states = ploidy + 1
        self.ploidy = ploidy
        self.global_cap = global_cap
        self.group = group
        self.gather = gather or self._nccl_gather
        # "all": every rank ends with every row (all-gather of the rows blocks).  "root": the rows are
        # gathered to rank 0 only (the driver's `collect`); the other ranks all-gather just the 32-byte
        # headers and end with the global table and their OWN rows.
        # "direct": nothing is gathered but 32-byte headers -- every rank's last kernel stores its owned rows
        # straight into rank 0's result columns over NVLink (avo_shard_call_direct; the columns live in a
        # CUDA-IPC block the other ranks map).  Rank 0 ends with every row, the others with the table.
        # `direct_root` (tests: several ranks emulated in ONE process): rank 0's pipeline, whose block the
        # other ranks then address directly (rank 0 itself passes True).
        assert rows_to in ("all", "root", "direct")
        self.rows_to = rows_to if (gather is None or rows_to == "direct") else "all"
        self._emul_root = direct_root if direct_root is not True else None
        self._emulated = direct_root is not None
        self._block = None            # rank 0: the peer-visible block; others: its mapping
        self._block_for = None
        self.nccl_bytes = 0
        # the phases and the collectives are ordered on ONE stream of the pipeline's own (the context's
        # default is a stream torch does not know; torch's default stream has no handle to pass)
        self.stream = stream or torch.cuda.Stream(self.dev)
        ctx.set_stream(self.stream.cuda_stream)
        self._size(site_cap, allele_cap)

    def _nccl_gather(self, out, inp):
        import torch.distributed as dist
        dist.all_gather_into_tensor(out, inp, group=self.group)

    def _gather_rows(self):
        """The second collective: the rows blocks to every rank, or to rank 0 (+ the headers to everyone), or
        the headers alone (direct rows)."""
        if self.rows_to == "all":
            self.gather(self.r_all, self.r_send)
            return
        import torch.distributed as dist
        torch, p, W = self.torch, self.plan, self.plan.world
        dist.all_gather_into_tensor(self.rh_all, self.r_send[:32], group=self.group)
        if self.rows_to == "direct":
            self._place_headers()
            return
        blocks = self.r_all.view(W, self.rb)
        if p.rank == 0:
            dist.gather(self.r_send, gather_list=list(blocks.unbind(0)), dst=0, group=self.group)
        else:
            dist.gather(self.r_send, dst=0, group=self.group)
            self._place_own_rows()

    def _place_headers(self):
        """Direct rows: the gathered headers (self.rh_all) stand in for the rows blocks."""
        W = self.plan.world
        self.r_all.view(W, self.rb)[:, :32].copy_(self.rh_all.view(W, 32))

    # ---- direct rows: rank 0's result columns in a block the other ranks can store into -------------------
    def _results_layout(self):
        offs, o = {}, 0
        for name, dt, w in self.L.RESULT_COLUMNS:
            width = self.plan.n_states if w == "ns" else w
            offs[name] = o
            o += (max(self.gcap * width, 1) * np.dtype(dt).itemsize + 255) // 256 * 256
        return offs, o

    def _results_at(self, base):
        res = self.L.SiteResultsStruct()
        res.n_sites, res.n_states, res.mem = self.gcap, self.plan.n_states, self.L.MEM_DEVICE
        offs, _ = self._results_layout()
        for name, off in offs.items():
            setattr(res, name, base + off)
        return res

    def _direct_release(self):
        import ctypes as C
        if self._block is None:
            return
        ctx, lib = self.ctx, self.ctx.lib
        if not self._emulated and self.plan.world > 1:
            import torch.distributed as dist
            self.stream.synchronize()
            if self.plan.rank != 0:
                ctx._check(lib.avo_peer_close(ctx.h, C.c_void_p(self._block)))
            dist.barrier(group=self.group)           # every mapping is gone before the owner frees
        if self.plan.rank == 0:
            self.torch.cuda.synchronize()
            ctx._check(lib.avo_peer_free(ctx.h, C.c_void_p(self._block)))
        self._block = None

    def _direct_setup(self):
        import ctypes as C
        ctx, lib, p, torch = self.ctx, self.ctx.lib, self.plan, self.torch
        self._direct_release()
        offs, total = self._results_layout()
        handle = C.create_string_buffer(64)
        ptr = C.c_void_p()
        if p.rank == 0:
            ctx._check(lib.avo_peer_alloc(ctx.h, total, C.byref(ptr), handle))
            self._block = ptr.value
        if not self._emulated and p.world > 1:
            import torch.distributed as dist
            obj = [handle.raw if p.rank == 0 else None]
            dist.broadcast_object_list(obj, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                                       group=self.group)
            if p.rank != 0:
                ctx._check(lib.avo_peer_open(ctx.h, obj[0], C.byref(ptr)))
                self._block = ptr.value
        self.cols = {}
        if p.rank == 0:
            class _Raw:                                  # the block as a torch tensor (no copy)
                pass
            raw = _Raw()
            raw.__cuda_array_interface__ = {"shape": (total,), "typestr": "|u1", "data": (self._block, False), "version": 2}
            whole = torch.as_tensor(raw, device=self.dev)
            for name, dt, w in self.L.RESULT_COLUMNS:
                width = p.n_states if w == "ns" else w
                nbytes = max(self.gcap * width, 1) * np.dtype(dt).itemsize
                self.cols[name] = whole[offs[name]:offs[name] + nbytes].view(getattr(torch, dt))
            self.res = self._results_at(self._block)
            self._block_for = self._block
        elif not self._emulated:
            self.res = self._results_at(self._block)
            self._block_for = self._block
        else:
            self.res, self._block_for = None, None       # bound to the root's block at call time

    def close(self):
        """Releases the peer-visible block of rows_to="direct" (collective: every rank calls it)."""
        self._direct_release()

    def _place_own_rows(self):
        """A rank that is not the root of the rows gather: its own block in place; of the other ranks only
        the header (self.rh_all), with a row count of 0 -- their flags still decide the retry."""
        torch, p, W = self.torch, self.plan, self.plan.world
        blocks = self.r_all.view(W, self.rb)
        h = self.rh_all.view(W, 32).clone()
        others = torch.arange(W, device=self.dev) != p.rank
        h[:, :4] = torch.where(others[:, None], torch.zeros_like(h[:, :4]), h[:, :4])
        blocks[:, :32].copy_(h)
        blocks[p.rank].copy_(self.r_send)

    def _size(self, site_cap, allele_cap):
        import ctypes as C
        torch, L, p = self.torch, self.L, self.plan
        p.site_cap, p.allele_cap, p.row_cap = int(site_cap), (int(allele_cap) + 15) // 16 * 16, int(site_cap)
        lib = self.ctx.lib
        self.tb = lib.avo_shard_table_block_bytes(C.byref(p))
        self.rb = lib.avo_shard_rows_block_bytes(C.byref(p))
        W = p.world
        self.t_send = torch.empty(self.tb, dtype=torch.uint8, device=self.dev)
        self.t_all = torch.empty(self.tb * W, dtype=torch.uint8, device=self.dev)
        self.r_send = torch.empty(self.rb, dtype=torch.uint8, device=self.dev)
        self.r_all = torch.empty(self.rb * W, dtype=torch.uint8, device=self.dev)
        self.rh_all = torch.empty(32 * W, dtype=torch.uint8, device=self.dev)
        gcap = self.global_cap or p.site_cap * W
        self.gcap, self.gacap = int(gcap), int(p.allele_cap * W * 2)
        from .batch import _device_empty
        mk = lambda n, dt: _device_empty(n, dt, self.dev)
        self.sites = dict(start=mk(self.gcap, np.int32), ref_len=mk(self.gcap, np.int32), alt_len=mk(self.gcap, np.int32),
                          allele_off=mk(self.gcap + 1, np.uint32), alleles4=mk(self.gacap // 2 + 16, np.uint8),
                          count=mk(self.gcap, np.int32))
        if self.rows_to == "direct":
            self._direct_setup()
        else:
            self.res, self.cols = self.ctx._alloc_results(self.gcap, p.n_states, self.dev)
        self._views = None

    # ---- the three phases (step() runs them around the two collectives; tests drive them by hand) ----
    def phase_discover(self, batch: ReadBatch, phred: int = 0, min_obs: Optional[int] = None):
        import ctypes as C
        ctx, p = self.ctx, self.plan
        key = (phred, min_obs, batch.n_reads)
        if getattr(self, "_bs_for", None) is not batch or self._bs_key != key:    # the same batch again: same structs
            self._params = ctx._params(ploidy=self.ploidy, phred=phred, min_obs=min_obs)
            self._bs = batch.as_struct()
            self._bs_for, self._bs_key = batch, key
        ctx._check(ctx.lib.avo_shard_discover(ctx.h, C.byref(self._bs), C.byref(self._params), C.byref(p),
                                              self.t_send.data_ptr()))

    def phase_call(self):
        import ctypes as C
        from ._lib import MEM_DEVICE, SitesOutStruct
        from .batch import _ptr
        ctx, p = self.ctx, self.plan
        out = SitesOutStruct()
        out.capacity, out.allele_capacity, out.mem = self.gcap, self.gacap, MEM_DEVICE
        for k, a in self.sites.items():
            setattr(out, k, _ptr(a))
        self._out = out
        if self.rows_to == "direct":
            if self._emul_root is not None and self._block_for != self._emul_root._block:
                self._block_for = self._emul_root._block           # (tests) the root pipeline's block, same process
                self.res = self._results_at(self._block_for)
            self.res.n_sites = self.gcap
            ctx._check(ctx.lib.avo_shard_call_direct(ctx.h, None, C.byref(self._params), C.byref(p), self.t_all.data_ptr(),
                                                     C.byref(out), self.r_send.data_ptr(), C.byref(self.res)))
            return
        self.res.n_sites = self.gcap
        ctx._check(ctx.lib.avo_shard_call(ctx.h, None, C.byref(self._params), C.byref(p), self.t_all.data_ptr(),
                                          C.byref(out), self.r_send.data_ptr()))

    def phase_finish(self):
        """-> (table, cols) or None when every rank has to run the step again (capacities grown)."""
        import ctypes as C
        from ._lib import AVO_E_CAPACITY, AVO_E_RETRY
        ctx, p, out = self.ctx, self.plan, self._out
        need = self._need = self.L.ShardPlanStruct()
        rc = ctx.lib.avo_shard_finish(ctx.h, C.byref(p), self.t_all.data_ptr(), self.r_all.data_ptr(), C.byref(out),
                                      C.byref(self.res), C.byref(need))
        # bytes this rank receives per step
        rows_in = self.rows_to == "all" or (self.rows_to == "root" and p.rank == 0)
        self.nccl_bytes = self.tb * p.world + (self.rb * p.world if rows_in else 32 * p.world)
        if rc == AVO_E_RETRY:
            return None
        if rc == AVO_E_CAPACITY:           # the same answer on every rank: all of them grow and run again
            self.global_cap = max(self.gcap, int(out.n) + int(out.n) // 8 + 256)
            self._size(max(p.site_cap, need.site_cap), max(p.allele_cap, need.allele_cap))
            return None
        ctx._check(rc)
        n, nn = int(out.n), int(out.n_allele_nibbles)
        if self._views is not None and self._views[0] == (n, nn):     # same shapes as the previous step: same views
            return self._views[1]
        s = self.sites
        table = SiteTable(n, nn, s["start"][:n], s["ref_len"][:n], s["alt_len"][:n], s["allele_off"][:n + 1],
                          s["alleles4"][:(nn + 1) // 2 or 1], None, s["count"][:n])
        cols = {}
        for name, dt, w in self.L.RESULT_COLUMNS:
            if name not in self.cols:              # direct rows: only rank 0 holds them
                continue
            width = p.n_states if w == "ns" else w
            a = self.cols[name][:n * width]
            cols[name] = a.reshape(n, width) if width != 1 else a
        self._views = ((n, nn), (table, cols))
        return table, cols

    def step(self, batch: ReadBatch, phred: int = 0, min_obs: Optional[int] = None, max_tries: int = 6):
        """One sharded discoverAndCall over this rank's reads (own interval + halo).  Returns, on every
        rank, (global SiteTable, dict of result columns for all its rows) as device tensors -- views of the
        pipeline's buffers, valid until the next step.  With rows_to="root" only rank 0's columns hold
        every row; on the other ranks the rows outside [own_lo, own_hi) are not filled."""
        from ._lib import AVO_E_CAPACITY, AvocadoError
        if getattr(self, "_shrink_to", None):        # (the previous step's result views end here)
            self.global_cap = None
            self._size(*self._shrink_to)
            self._shrink_to = None
        for _ in range(max_tries):
            with self.torch.cuda.stream(self.stream):
                self.phase_discover(batch, phred, min_obs)
                self.gather(self.t_all, self.t_send)
                self.phase_call()
                self._gather_rows()
                got = self.phase_finish()
            if got is not None:
                # blocks much larger than what the ranks own waste NVLink bytes: every rank sees the same
                # `needed` (it follows from the gathered headers) and shrinks the same way before its next step
                need, p = self._need, self.plan
                if p.site_cap > need.site_cap + need.site_cap // 4 or p.allele_cap > 2 * need.allele_cap:
                    self._shrink_to = (int(need.site_cap), int(need.allele_cap))
                return got
        raise AvocadoError(AVO_E_CAPACITY, "sharded step kept asking for larger buffers")

    def step_phases_ms(self, batch: ReadBatch, phred: int = 0, min_obs: Optional[int] = None):
        """One steady-state step with CUDA events between its phases (a measurement aid; a step that has
        to retry raises).  Returns the milliseconds of discover / table gather / call / rows gather / finish."""
        ev = [self.torch.cuda.Event(enable_timing=True) for _ in range(6)]
        with self.torch.cuda.stream(self.stream):
            ev[0].record()
            self.phase_discover(batch, phred, min_obs); ev[1].record()
            self.gather(self.t_all, self.t_send); ev[2].record()
            self.phase_call(); ev[3].record()
            self._gather_rows(); ev[4].record()
            got = self.phase_finish(); ev[5].record()
        self.stream.synchronize()
        if got is None:
            raise RuntimeError("step_phases_ms needs a pipeline in steady state")
        names = ("discover", "gather_tables", "call", "gather_rows", "finish")
        return {n: ev[i].elapsed_time(ev[i + 1]) for i, n in enumerate(names)}


class GpuEngine:
    """The product engine: avo_discover / avo_call on this rank's GPU."""

    def __init__(self, ctx):
        self.ctx = ctx

    def discover(self, batch: ReadBatch, phred: int, min_obs: Optional[int]) -> SiteTable:
        return self.ctx.discover_packed(batch, phred=phred, min_obs=min_obs)

    def call(self, batch: ReadBatch, sites: SiteTable, ploidy: int) -> Dict[str, np.ndarray]:
        return self.ctx.call_packed(batch, sites, ploidy=ploidy)


def discover_and_call_sharded(engine, batch: ReadBatch, shard: RegionShard, phred: int = 0,
                              min_obs: Optional[int] = None, ploidy: int = 2, group=None,
                              collective_device=None):
    """BiallelicGenotyper.discoverAndCall (BG:156-184) over region shards.  `batch` holds THIS
    rank's reads (shard.read_lo .. shard.read_hi of the sorted whole).  Returns, on every rank,
    (global SiteTable, dict of result columns for all its rows)."""
    local = engine.discover(batch, phred, min_obs).to_host()
    st = np.asarray(local.start[:local.n], np.int64)
    owned = _take_sites(local, (st >= shard.lo) & (st < shard.hi))
    parts = _all_gather_objects(owned, group)                    # replicate the global sorted table
    table, offs = _cat_sites(parts)
    if table.n == 0:
        return table, {}
    if int(np.max(table.ref_len[:table.n])) > shard.max_site_len:
        raise ValueError("a discovered site is longer than the shard plan's max_site_len=%d: "
                         "re-plan with a larger halo" % shard.max_site_len)
    cols = engine.call(batch, table, ploidy)
    my = slice(offs[shard.rank], offs[shard.rank] + parts[shard.rank].n)
    merged = {}
    for k, v in cols.items():                                    # gather the rows each rank owns
        rows = _all_gather_rows(np.asarray(v)[my], collective_device, group)
        merged[k] = np.concatenate(rows) if len(rows) > 1 else rows[0]
    return table, merged
