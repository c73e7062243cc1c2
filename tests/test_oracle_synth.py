"""The oracle carries its own copy of the workload generator (same source: avocado_b200/csrc/avo_synth_core.h),
so that bench.py's CPU arms build their input without loading the product library.  Both must produce
identical bytes -- otherwise the GPU and CPU arms would not be looking at the same reads."""
import numpy as np


def test_oracle_generator_reproduces_the_product_generator(oracle):
    from avocado_b200 import batch as B
    kw = dict(seed=0xA70CAD0 + 1, n_reads_total=60000, contig_len=150000, first_pos=77, sample=2)
    b = B.synth_host(B.synth_params(**kw), first=1000, n=30000)
    d = oracle.synth_packed(kw, 1000, 30000, 4)
    m = b.meta
    for k in ("start", "end", "mapq", "flags", "seq_off", "l_seq", "op_off", "n_ops", "refb_off", "qhead"):
        assert np.array_equal(np.asarray(m[k]).astype(np.int64), d[k].astype(np.int64)), k
    assert np.array_equal(b.payload[:b.n_blocks * 16], d["payload"])
    assert np.array_equal(b.ops[:b.n_ops], d["ops"]) and np.array_equal(b.refb[:b.n_refb], d["refb"])
    rs = oracle.synth_reads(kw, 1000, 30000, "ctg", 3)
    rs2 = oracle.reads_from_packed("ctg", m["start"], m["end"], m["mapq"], m["flags"], m["seq_off"], m["l_seq"], b.payload,
                                   m["op_off"], m["n_ops"], b.ops, m["refb_off"], b.refb)
    assert len(rs) == len(rs2) == 30000
    for i in range(0, 30000, 97):
        assert oracle.read_as_dict(rs, i) == oracle.read_as_dict(rs2, i)
    # the text reads discover the same sites with any thread count
    assert oracle.discover(rs, 25, 5, 1) == oracle.discover(rs2, 25, 5, 3)
