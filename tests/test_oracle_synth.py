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


def test_reference_arm_of_the_bench_runs_without_the_product_library():
    """`bench.py --impl reference` (the CPU arm the driver times beside the GPU arm): one JSON line with the
    contract's keys, produced by a process that never maps libavocado_b200.so or imports torch."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', "
            "'--cpu-sample', '20000']; runpy.run_path(%r, run_name='__main__'); "
            "maps = open('/proc/self/maps').read(); "
            "sys.stderr.write('LOADED_PRODUCT=%%d TORCH=%%d\\n' %% ('libavocado_b200' in maps, 'torch' in sys.modules))"
            % os.path.join(root, "bench.py"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "reads/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]
    assert "LOADED_PRODUCT=0 TORCH=0" in r.stderr, r.stderr[-500:]
