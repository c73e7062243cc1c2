import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from avocado_b200 import batch as B
from avocado_b200.genotyping import Context
from avocado_b200.parallel import ShardedPipeline, plan_region_shards
p = B.synth_params(seed=33, n_reads_total=80000, contig_len=400000, first_pos=10000)
whole = B.synth_host(p)
world=2
shards = plan_region_shards(whole.meta["start"], whole.meta["end"], world, max_site_len=64)
ctxs=[Context(0) for _ in range(world)]
pipes=[];batches=[]
for sh in shards:
    lo=max(sh.lo,-(1<<31)+1); hi=min(sh.hi,(1<<31)-1)
    print("shard",sh)
    pipes.append(ShardedPipeline(ctxs[sh.rank], sh.rank, world, lo, hi, site_cap=4096, allele_cap=16384, gather=lambda o,i:None))
    batches.append(whole.slice(sh.read_lo, sh.read_hi).to_device("cuda:0"))
for pp,b in zip(pipes,batches): pp.phase_discover(b,25,5)
allt=torch.cat([pp.t_send for pp in pipes])
for r,pp in enumerate(pipes):
    print("table hdr", r, pp.t_send[:32].view(torch.int32).tolist())
for pp in pipes:
    pp.t_all.copy_(allt); pp.phase_call()
    torch.cuda.synchronize()
    print("rows hdr", pp.plan.rank, pp.r_send[:32].view(torch.int32).tolist(), "timing", {k:v for k,v in ctxs[pp.plan.rank].timing().items() if 'ms' in k or 'launch' in k})
    # emitted column of packed block
    em = pp.r_send[32:32+4096]
    print(" packed emitted sum", int(em.sum()))
allr=torch.cat([pp.r_send for pp in pipes])
for pp in pipes:
    pp.r_all.copy_(allr); g=pp.phase_finish()
    t,c=g
    e=c["emitted"].cpu().numpy()
    print("rank",pp.plan.rank,"n",t.n,"emitted",e.sum(), "first zero", int(np.argmin(e)) if (e==0).any() else -1)
