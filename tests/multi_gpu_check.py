"""Run under torchrun with >= 2 ranks (one per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Checks, on real GPUs, that (1) region-sharded discoverAndCall (avocado_b200.parallel.ShardedPipeline:
avo_shard_discover / _call / _finish around two NCCL all-gathers, everything device-resident; then with the
rows gathered to rank 0 only, then stored into rank 0's columns over CUDA-IPC peer memory) returns on
every rank exactly the rows a single GPU computes for the whole batch, and (2) sample-sharded joint
calling (avocado_b200.joint.joint_call_sharded: all-reduce of allele counts and of posterior sums
over NCCL) reproduces the single-GPU cohort result, and (3) so does the sample-sharded depth roll-up
(joint_depths_sharded: one all-reduce of 13 int32 vectors).  Exits non-zero on any mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from avocado_b200 import batch as B
    from avocado_b200.genotyping import Context
    from avocado_b200.joint import joint_call, joint_call_sharded
    from avocado_b200.parallel import ShardedPipeline, plan_region_shards
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    dev = torch.device("cuda", local)

    # ---- (1) region-sharded discoverAndCall: the device-resident pipeline over NCCL ---------------------
    p = B.synth_params(seed=33, n_reads_total=400000, contig_len=2000000, first_pos=10000)
    whole = B.synth_host(p)
    shards = plan_region_shards(whole.meta["start"], whole.meta["end"], world, max_site_len=64)
    sh = shards[rank]
    mine = whole.slice(sh.read_lo, sh.read_hi).to_device(dev)
    pipe = ShardedPipeline(ctx, rank, world, max(sh.lo, -(1 << 31) + 1), min(sh.hi, (1 << 31) - 1),
                           site_cap=256, allele_cap=256)        # too small on purpose: every rank grows together
    for _ in range(3):                                          # grow, shrink to fit, steady state
        table, cols = pipe.step(mine, phred=25, min_obs=5)
    table = table.to_host()
    cols = {k: v.cpu().numpy() for k, v in cols.items()}
    # the single-GPU answer, computed locally without collectives
    c1x = Context(local)
    t1, c1 = c1x.discover_and_call_packed(whole, phred=25, min_obs=5)
    c1x.close()
    assert table.n == t1.n and table.n > 1000, (table.n, t1.n)
    for k in ("start", "ref_len", "alt_len", "count", "allele_off"):
        assert np.array_equal(np.asarray(getattr(table, k)), np.asarray(getattr(t1, k))), k
    for k, v in c1.items():
        assert np.array_equal(cols[k], np.asarray(v), equal_nan=True), "column %s differs on rank %d" % (k, rank)
    nccl_bytes = pipe.nccl_bytes
    # the same with the rows gathered to rank 0 only: the others keep the table and their own rows
    pipe_r = ShardedPipeline(ctx, rank, world, max(sh.lo, -(1 << 31) + 1), min(sh.hi, (1 << 31) - 1),
                             site_cap=256, allele_cap=256, rows_to="root")
    for _ in range(3):
        table_r, cols_r = pipe_r.step(mine, phred=25, min_obs=5)
    table_r = table_r.to_host()
    st = np.asarray(t1.start)
    assert table_r.n == t1.n and np.array_equal(np.asarray(table_r.start), st)
    own = np.ones(t1.n, bool) if rank == 0 else (st >= sh.lo) & (st < sh.hi)
    for k, v in c1.items():
        v = np.asarray(v)
        per = v.size // t1.n
        m = np.repeat(own, per) if per > 1 else own
        assert np.array_equal(cols_r[k].cpu().numpy().ravel()[m], v.ravel()[m], equal_nan=True), "root-gather column %s differs on rank %d" % (k, rank)
    assert rank == 0 or pipe_r.nccl_bytes < nccl_bytes
    # and with no rows gathered at all: every rank stores its owned rows into rank 0's columns over NVLink
    pipe_d = ShardedPipeline(ctx, rank, world, max(sh.lo, -(1 << 31) + 1), min(sh.hi, (1 << 31) - 1),
                             site_cap=256, allele_cap=256, rows_to="direct")
    for _ in range(4):
        table_d, cols_d = pipe_d.step(mine, phred=25, min_obs=5)
    assert table_d.n == t1.n and np.array_equal(table_d.to_host().start, st)
    if rank == 0:
        for k, v in c1.items():
            assert np.array_equal(cols_d[k].cpu().numpy(), np.asarray(v), equal_nan=True), "direct rows: column %s differs" % k
    else:
        assert cols_d == {}
    assert pipe_d.nccl_bytes <= pipe_r.nccl_bytes and (rank != 0 or pipe_d.nccl_bytes < pipe_r.nccl_bytes or world == 1)
    pipe_d.close()

    # ---- (2) sample-sharded joint calling ------------------------------------------------------
    rng = np.random.default_rng(11)
    S, N, ns = 4 * world, 20000, 3
    n_alt = rng.integers(0, 3, (S, N)).astype(np.int32)
    n_alt[:, rng.random(N) < 0.3] = 0
    n_ref = (2 - n_alt).astype(np.int32)
    n_other = np.zeros((S, N), np.int32)
    gl = -np.abs(rng.normal(0, 6, (S, N, ns)))
    lo, hi = rank * 4, (rank + 1) * 4
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    local_cols = dict(n_sites=N, n_samples=hi - lo, n_states=ns, n_alt=cu(n_alt[lo:hi].ravel()),
                      n_ref=cu(n_ref[lo:hi].ravel()), n_other=cu(n_other[lo:hi].ravel()), gl=cu(gl[lo:hi].ravel()))
    out = joint_call_sharded(ctx, local_cols)
    ref = joint_call(ctx, N, S, ns, n_alt.ravel(), n_ref.ravel(), n_other.ravel(), gl.ravel())
    for k in ("flags", "posteriors", "priors", "n_alt", "n_ref", "gq"):
        assert np.array_equal(out[k].cpu().numpy(), ref[k][lo:hi], equal_nan=True), k
    q, q1 = out["quality"].cpu().numpy(), ref["quality"]
    em = ref["site_emitted"].astype(bool)
    assert np.allclose(q[em], q1[em], rtol=1e-9, atol=1e-9)
    assert np.array_equal(out["alt_alleles"].cpu().numpy(), n_alt.sum(0))
    # ---- (3) the depth roll-up of the same cohort, sharded by sample -----------------------------
    from avocado_b200.joint import joint_depths, joint_depths_sharded
    present = (rng.random((S, N)) > 0.4).astype(np.uint8)
    present[:, :50] = 0
    present[1, 25:50] = 1
    dp = np.where(rng.random((S, N)) > 0.3, rng.integers(0, 90, (S, N)), -1).astype(np.int32)
    refdp = np.where(rng.random((S, N)) > 0.3, rng.integers(0, 90, (S, N)), -1).astype(np.int32)
    sb = rng.integers(0, 30, (S, N, 4)).astype(np.int32)
    sb[rng.random((S, N)) > 0.5] = -1
    got = joint_depths_sharded(ctx, N, hi - lo, cu(dp[lo:hi].ravel()), cu(refdp[lo:hi].ravel()), cu(sb[lo:hi].ravel()),
                               cu(present[lo:hi].ravel()))
    want = joint_depths(ctx, N, S, dp.ravel(), refdp.ravel(), sb.ravel(), present.ravel())
    for k, v in want.items():
        assert np.array_equal(got[k].cpu().numpy(), np.asarray(v)), "depth column %s differs on rank %d" % (k, rank)
    dist.barrier()
    if rank == 0:
        print("multi_gpu_check ok: world=%d, %d sites region-sharded (%d NCCL bytes received per rank per step), %d x %d joint"
              % (world, table.n, nccl_bytes, S, N))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
