import sys, os
sys.path.insert(0, os.getcwd())
import torch
from avocado_b200 import batch as B
from avocado_b200.genotyping import Context
import bench
p = B.synth_params(**bench.synth_dict(25_600_000, 64_000_000, 1))
dev = B.synth_device(p, first=0, n=25_600_000, device="cuda:0")
ctx = Context(0)
ctx.derive_events(dev)
for _ in range(4):
    ctx.discover_and_call_packed(dev, ploidy=2, phred=25, min_obs=5, capacity=1 << 18, device_out=True, copy=False)
torch.cuda.synchronize()
