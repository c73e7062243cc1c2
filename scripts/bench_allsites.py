"""GPU box: scoreAllSites (gVCF mode) at the configs[1] size, device-resident input and output.
usage: python scripts/bench_allsites.py [reads]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avocado_b200 import batch as B
from avocado_b200.genotyping import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25_600_000
p = B.synth_params(seed=0xA70CAD0 + 1, contig_len=int(64_000_000 * n / 25_600_000), n_reads_total=n)
dev = B.synth_device(p)
ctx = Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
sites = ctx.discover_packed(dev, phred=25, min_obs=5, device_out=True)
cap = None
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cols, rstart, rcols = ctx.call_all_sites_packed(dev, sites, capacity=cap, device_out=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    cap = len(rstart) + 16
    t = ctx.timing()
    print("iter %d: %d variant sites, %d reference-model rows, wall %.2f ms (incl. output allocation), device total %.2f ms, "
          "call tiles %.2f ms -> %.3e reads/s" % (it, sites.n, len(rstart), dt * 1e3, t["ms_total"], t["ms_call_tiles"], n / (t["ms_total"] / 1e3)))
    del cols, rcols
