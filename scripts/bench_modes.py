"""GPU box: throughput of the other modes of the path at BASELINE.json shapes (not the headline
bench line): scoreAllSites at configs[1], the filter pass over its rows, and joint calling at the
configs[3] shape (100 samples x 1.5e5 sites), single GPU or sample-sharded under torchrun.
usage: python scripts/bench_modes.py [--reads R] > gpurun_out/modes.json
       python -m torch.distributed.run --nproc-per-node N ... scripts/bench_modes.py --joint-only"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from avocado_b200 import batch as B
from avocado_b200.filters import FilterArgs, filter_columns
from avocado_b200.genotyping import Context
from avocado_b200.joint import joint_call, joint_call_sharded

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=25_600_000)
ap.add_argument("--joint-only", action="store_true")
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = Context(local)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
out = {"n_gpus": world}


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), r


if not args.joint_only:
    n = args.reads
    p = B.synth_params(seed=0xA70CAD0 + 1, contig_len=int(64_000_000 * n / 25_600_000), n_reads_total=n)
    dev = B.synth_device(p)
    sites = ctx.discover_packed(dev, phred=25, min_obs=5, device_out=True)
    cols, rstart, rcols = ctx.call_all_sites_packed(dev, sites, device_out=True)      # sizes the output
    cap = len(rstart) + 16
    del cols, rcols
    ms, (cols, rstart, rcols) = timed(lambda: ctx.call_all_sites_packed(dev, sites, capacity=cap, device_out=True), reps=3, warm=1)
    n_rows = len(rstart)
    ms_dev = ctx.timing()["ms_total"]
    out["score_all_sites"] = {"reads": n, "variant_rows": sites.n, "reference_model_rows": n_rows,
                              "ms_library": ms_dev, "reads_per_s": n / (ms_dev / 1e3), "ms_with_python_wrapper": ms,
                              "note": "avo_call_all_sites, device-resident input and output, configs[1] shape; "
                                      "ms_library = the library's own CUDA events around the call, the wrapper "
                                      "figure adds allocation of the 3.7 GB of output tensors"}
    ns = 3
    keep = {k: rcols[k][:n_rows] for k in ("emitted", "n_alt", "n_ref", "n_other", "gq", "total_cov", "allele_cov", "fisher", "rms_mapq")}
    ref_len = torch.ones(n_rows, dtype=torch.int32, device="cuda")
    alt_len = torch.full((n_rows,), -1, dtype=torch.int32, device="cuda")
    ms_f, _ = timed(lambda: filter_columns(ctx, keep, ref_len, alt_len, FilterArgs(), False, False))
    out["filter_pass"] = {"rows": n_rows, "ms": ms_f, "rows_per_s": n_rows / (ms_f / 1e3),
                          "note": "avo_filter_genotypes over the reference-model rows, device-resident"}
    del dev, cols, rcols, keep

# joint calling at the configs[3] shape: 100 samples x 1.5e5 union sites, sharded by sample over the ranks
S, N, ns = 100, 150_000, 3
rng = np.random.default_rng(3)
lo, hi = rank * S // world, (rank + 1) * S // world
maf = rng.beta(0.5, 5.0, N)
n_alt = (rng.random((S, N, 2)) < maf[None, :, None]).sum(2).astype(np.int32)[lo:hi]
gl = -np.abs(rng.normal(0, 6, (hi - lo, N, ns)))
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
cols = dict(n_sites=N, n_samples=hi - lo, n_states=ns, n_alt=cu(n_alt.ravel()), n_ref=cu((2 - n_alt).ravel()),
            n_other=cu(np.zeros_like(n_alt).ravel()), gl=cu(gl.ravel()))
def per_call(fn, reps=30, warm_s=1.0):
    """Median wall time per call (host clock around a synchronised call) and median of the library's own
    CUDA-event total, after at least warm_s seconds of calls (an idle GPU needs that to clock up)."""
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < warm_s:
        fn()
        torch.cuda.synchronize()
    wall, lib = [], []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        wall.append((time.perf_counter() - t) * 1e3)
        lib.append(ctx.timing()["ms_total"])
    return float(np.median(wall)), float(np.median(lib))


if world > 1:
    ms_j, _ = timed(lambda: joint_call_sharded(ctx, cols))
    ms_lib = None
else:
    ms_j, ms_lib = per_call(lambda: joint_call(ctx, **cols))
out["joint"] = {"samples": S, "sites": N, "ms": ms_j, "ms_library_events": ms_lib, "genotypes_per_s": S * N / (ms_j / 1e3),
                "note": "avo_joint (+ avo_joint_counts and two NCCL all-reduces when sharded), device-resident, "
                        "incl. output allocation by the Python wrapper"}
# the depth roll-up of the same matrix (avo_joint_depths): 28 B in per genotype, 25 B out per site
from avocado_b200.joint import joint_depths  # noqa: E402
n_loc = hi - lo
dp = cu(rng.integers(0, 60, (n_loc, N)).astype(np.int32).ravel())
refdp = cu(rng.integers(0, 40, (n_loc, N)).astype(np.int32).ravel())
sbc = cu(rng.integers(0, 20, (n_loc, N, 4)).astype(np.int32).ravel())
ms_d, ms_d_lib = per_call(lambda: joint_depths(ctx, N, n_loc, dp, refdp, sbc))
bytes_d = n_loc * N * 24 + N * 25
out["joint_depths"] = {"samples": n_loc, "sites": N, "ms": ms_d, "ms_library_events": ms_d_lib,
                       "gbps": bytes_d / (ms_d_lib / 1e3) / 1e9,
                       "note": "avo_joint_depths on this rank's samples, device-resident, incl. output allocation "
                               "by the Python wrapper; gbps = (24 B per genotype in + 25 B per site out) / library-event time"}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
