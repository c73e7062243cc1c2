"""Host-side logic of the multi-GPU path on CPU: the region-shard plan, the site-table merge and
the collectives (world_size 2, gloo).  The per-rank compute is supplied by the oracle here (test
infrastructure), so these tests check the sharding itself: halos, ownership, ordering, gathers."""
import os
import socket
import sys

import numpy as np
import pytest

from helpers import ORACLE_OF, oracle_reads_from_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleEngine:
    """discover / call of one rank computed by the CPU oracle (checker standing in for the GPU)."""

    def __init__(self, oracle):
        self.oracle = oracle

    def discover(self, batch, phred, min_obs):
        from avocado_b200.batch import sites_from_variants
        from avocado_b200.records import Variant
        rs = oracle_reads_from_batch(self.oracle, batch)
        v = sorted(self.oracle.discover(rs, phred, min_obs), key=lambda t: (t[1], len(t[3]), len(t[4]), t[3] + t[4]))
        t = sites_from_variants([Variant(c, s, e, r, a) for c, s, e, r, a, _ in v], {"ctg": 0})
        t.count = np.array([n for *_, n in v], np.int32)
        t.contig = None
        return t

    def call(self, batch, sites, ploidy):
        from avocado_b200 import _lib as L
        rs = oracle_reads_from_batch(self.oracle, batch)
        variants = [("ctg", v.start, v.end, v.referenceAllele, v.alternateAllele) for v in sites.to_variants(["ctg"])]
        res = self.oracle.call(rs, variants, ploidy, [], False, 93, 93, 1)
        row = {(int(s), r, a): j for j, (s, r, a) in enumerate(zip(res["start"], res["referenceAllele"],
                                                                    res["alternateAllele"]))}
        ns = ploidy + 1
        cols = {}
        for name, dt, w in L.RESULT_COLUMNS:
            width = ns if w == "ns" else w
            cols[name] = np.zeros((sites.n, width) if width != 1 else sites.n, dt)
        for i, v in enumerate(variants):
            j = row.get((v[1], v[3], v[4]))
            if j is None:
                continue
            cols["emitted"][i] = 1
            for name, oname in ORACLE_OF.items():
                val = np.asarray(res[oname][j])
                cols[name][i] = val[:cols[name][i].size] if val.ndim else val
        return cols


def test_shard_plan_covers_every_site_and_read():
    from avocado_b200.parallel import plan_region_shards
    rng = np.random.default_rng(3)
    start = np.sort(rng.integers(1000, 200000, 5000)).astype(np.int32)
    end = start + rng.integers(100, 180, 5000).astype(np.int32)
    for world in (1, 2, 3, 8):
        sh = plan_region_shards(start, end, world, max_site_len=64)
        assert sh[0].read_lo == 0 and sh[-1].read_hi == len(start)
        assert all(a.hi == b.lo for a, b in zip(sh, sh[1:]))                # owned intervals tile the line
        sizes = [s.read_hi - s.read_lo for s in sh]
        assert max(sizes) <= len(start) / world * 1.3 + 200                 # balanced up to the halo
        span = int((end - start).max())
        for s in sh:
            mine = np.zeros(len(start), bool)
            mine[s.read_lo:s.read_hi] = True
            # every read that can overlap an owned site [p, p + 64) with lo <= p < hi is present
            need = (end.astype(np.int64) > s.lo) & (start.astype(np.int64) < s.hi + 64)
            assert not np.any(need & ~mine)


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import avocado_oracle as oracle
    from avocado_b200 import batch as B
    from avocado_b200.parallel import discover_and_call_sharded, plan_region_shards
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        p = B.synth_params(seed=21, n_reads_total=6000, contig_len=30000, first_pos=500)
        whole = B.synth_host(p)
        shards = plan_region_shards(whole.meta["start"], whole.meta["end"], world, max_site_len=64)
        sh = shards[rank]
        mine = whole.slice(sh.read_lo, sh.read_hi)
        table, cols = discover_and_call_sharded(OracleEngine(oracle), mine, sh, phred=25, min_obs=5)
        if rank == 0:
            np.savez(out_path, start=table.start, ref_len=table.ref_len, alt_len=table.alt_len,
                     count=table.count, **{"c_" + k: v for k, v in cols.items()})
    finally:
        dist.destroy_process_group()


def test_region_sharded_discover_and_call_equals_one_rank(oracle, tmp_path):
    import torch.multiprocessing as mp
    from avocado_b200 import batch as B
    from avocado_b200.parallel import RegionShard, discover_and_call_sharded
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    p = B.synth_params(seed=21, n_reads_total=6000, contig_len=30000, first_pos=500)
    whole = B.synth_host(p)
    one = RegionShard(0, -(1 << 40), 1 << 40, 0, whole.n_reads, 64)
    table, cols = discover_and_call_sharded(OracleEngine(oracle), whole, one, phred=25, min_obs=5)
    assert table.n > 20
    for k in ("start", "ref_len", "alt_len", "count"):
        assert np.array_equal(got[k], getattr(table, k)), k
    for k, v in cols.items():
        assert np.array_equal(got["c_" + k], v, equal_nan=True), k      # bit-identical, incl. fp sums


def test_joint_counts_all_reduce_over_gloo(tmp_path):
    """The cross-sample exchange of joint calling: per-rank allele counts summed by all_reduce give
    the cohort counts (avocado_b200.joint.joint_call_sharded does exactly this around avo_joint)."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "counts.npy")
    mp.spawn(_count_worker, args=(2, port, out), nprocs=2, join=True)
    rng = np.random.default_rng(5)
    n_alt = rng.integers(0, 3, (6, 300)).astype(np.int32)
    assert np.array_equal(np.load(out), n_alt.sum(0))


def _count_worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        n_alt = rng.integers(0, 3, (6, 300)).astype(np.int32)
        mine = n_alt[rank * 3:(rank + 1) * 3].sum(0).astype(np.int32)      # what avo_joint_counts returns
        t = torch.from_numpy(mine)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if rank == 0:
            np.save(out_path, t.numpy())
    finally:
        dist.destroy_process_group()


def test_depth_roll_up_reduced_over_gloo(tmp_path):
    """The exchange of the depth roll-up for a cohort sharded by sample (joint.reduce_site_depths): per-rank
    summaries (the oracle's VariantSummary stands in for avo_joint_depths, as in the other tests of this file)
    reduced over two ranks equal the summary over all samples, unset fields and single-genotype sites included."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "depths.npz")
    mp.spawn(_depth_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    want, count = _depth_columns(*_depth_inputs(), 0, 6)
    for k, v in want.items():
        assert np.array_equal(got[k], v), k
    assert np.array_equal(got["annotated"], (count > 1).astype(np.uint8))
    assert (got["read_depth"] == -1).any() and (got["annotated"] == 0).any() and (got["fwd_read_depth"] >= 0).any()


def _depth_inputs():
    rng = np.random.default_rng(21)
    S, N = 6, 200
    present = rng.random((S, N)) > 0.5
    present[:, :10] = False
    present[4, 5:10] = True
    dp = np.where(rng.random((S, N)) > 0.5, rng.integers(0, 90, (S, N)), -1)
    refdp = np.where(rng.random((S, N)) > 0.5, rng.integers(0, 90, (S, N)), -1)
    sb = rng.integers(0, 30, (S, N, 4))
    sb[rng.random((S, N)) > 0.4] = -1
    return present, dp, refdp, sb


def _depth_columns(present, dp, refdp, sb, lo, hi):
    """What avo_joint_depths returns for samples [lo, hi): the oracle's merged VariantSummary per site."""
    import avocado_oracle as oracle
    from avocado_b200 import _lib as L
    names = dict(zip(L.SITE_DEPTH_COLUMNS, ("readDepth", "referenceReadDepth", "forwardReadDepth", "reverseReadDepth",
                                            "forwardReferenceReadDepth", "reverseReferenceReadDepth")))
    N = present.shape[1]
    cols = {c: -np.ones(N, np.int32) for c in L.SITE_DEPTH_COLUMNS}
    opt = lambda v: None if v < 0 else int(v)
    for i in range(N):
        idx = [s for s in range(lo, hi) if present[s, i]]
        if not idx:
            continue
        o = oracle.variant_summary([(opt(dp[s, i]), opt(refdp[s, i]), [] if sb[s, i, 0] < 0 else sb[s, i].tolist())
                                    for s in idx])
        for c, n in names.items():
            cols[c][i] = -1 if o[n] is None else o[n]
    return cols, present[lo:hi].sum(0).astype(np.int32)


def _depth_worker(rank, world, port, out_path):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
    from avocado_b200.joint import reduce_site_depths
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        local, count = _depth_columns(*_depth_inputs(), rank * 3, rank * 3 + 3)
        out = reduce_site_depths(local, count)
        if rank == 0:
            np.savez(out_path, **out)
    finally:
        dist.destroy_process_group()
