#!/usr/bin/env python
"""bench.py -- reads genotyped per second for avocado's genotyping hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--reads R]

One "step" = BiallelicGenotyper.discoverAndCall (DiscoverVariants, then call against the discovered
sites; the CLI default, CLI/BiallelicGenotyper.scala:255-263) over one synthetic batch of the
BASELINE.json configs[1] shape: single-sample 60x chr20 (64 Mb), 150 bp reads = 25.6 M reads.

  value     reads/s with the packed batch already resident in HBM (device pointers through the C ABI)
  e2e       the same C-ABI call with HOST (pinned) buffers, the batch cut into region bins that a few
            host threads (one context + stream each) pipeline: H2D of the records / operators /
            mismatch bytes (compact wire form), in-place payload reads over PCIe, D2H of the site
            table + all per-site results are inside the timed region; `single_call` = one bin;
            `from_text` = a bounded sample of the bins starting from AlignmentRecord TEXT columns
            (CIGAR, MD, sequence, qualities): avo_pack_reads_device + avo_discover_and_call, with the
            threaded host packer (avo_pack_reads_mt) timed beside it
  roofline  the longest kernel (k_discover_tiles, or the call stage's kernels together): measured
            DRAM bytes per launch (ncu, profiles/traffic.json) / its CUDA-event duration, because the
            kernels are sparse (SURVEY.md 8d anti-inflation rule); `full_scan_step` = every packed
            input byte once over the whole step; peak = MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle (a C++ restatement of the reference; the Scala/Spark reference
            cannot run here: no JVM) on a bounded sample of the same workload, 1 thread

With torchrun (N ranks) every rank processes its own region shard of a proportionally larger
genome (weak scaling, no data-path collective); times are max over ranks.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READS_C2 = 25_600_000
CONTIG_C2 = 64_000_000
SITE_ROW_BYTES = 208   # SURVEY.md 8(d) B_site at P = 2


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=READS_C2, help="reads per GPU (default: configs[1])")
    ap.add_argument("--cpu-sample", type=int, default=8_000_000, help="reads in the CPU baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-bins", type=int, default=8, help="region bins the e2e leg is cut into")
    ap.add_argument("--e2e-threads", type=int, default=0,
                    help="host threads (one context + stream each) that work through the e2e bins "
                         "(0 = min(8, host threads / (2 x ranks)))")
    ap.add_argument("--e2e-payload", type=int, default=-1, help="avo_params.host_payload of the e2e leg: 0 in place, "
                    "1 copy, 2 in place for discovery + host-gathered blocks for the call stage "
                    "(-1 = 2 when the rank has 16 or more host threads to itself, else 0)")
    ap.add_argument("--gather-threads", type=int, default=2,
                    help="avo_params.gather_threads: host threads gathering payload blocks per call")
    ap.add_argument("--no-text", action="store_true", help="skip the from-text leg of e2e")
    ap.add_argument("--text-reads", type=int, default=6_400_000, help="reads of the from-text sample")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def workload(args, world):
    """Synthetic parameters of this job: each rank owns `reads` reads over a 64 Mb-per-rank genome."""
    per = args.reads
    contig = int(CONTIG_C2 * (per / READS_C2))
    return per, contig


def synth_for_rank(B, per, contig, rank, world):
    # one stateless genome of world*contig bases; rank r owns reads [r*per, (r+1)*per), which start
    # in [r*contig, (r+1)*contig): region sharding with no shared sites
    p = B.synth_params(seed=0xA70CAD0 + 1, contig_len=contig * world, n_reads_total=per * world)
    return p


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = "/tmp/avo_clocks_%d_%d.csv" % (os.getpid(), gpu_index)

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def oracle_sample(B, p, first, n, contig_id=0):
    """Builds the oracle's text reads for reads [first, first+n) of the workload (device generation,
    D2H, unpack).  Returns (ReadSet, n)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import avocado_oracle as oracle
    dev = B.synth_device(p, first=first, n=n)
    h = dev.to_host()
    del dev
    m = h.meta
    rs = oracle.reads_from_packed("chr20", m["start"], m["end"], m["mapq"], m["flags"], m["seq_off"],
                                  m["l_seq"], h.payload, m["op_off"], m["n_ops"], h.ops,
                                  m["refb_off"], h.refb)
    return oracle, rs


def oracle_step(oracle, rs, threads):
    v = oracle.discover(rs, 25, 5, threads)
    res = oracle.call(rs, [(c, s, e, r, a) for c, s, e, r, a, _ in v], 2, [], False, 93, 93, threads)
    return len(v), len(res["start"])


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The Scala/Spark reference
    cannot be built or run in this image (no JVM), so this arm times the oracle port
    with all host threads on a bounded sample per step.  Rank 0 only."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    import torch
    from avocado_b200 import batch as B
    per, contig = workload(args, world)
    p = synth_for_rank(B, per, contig, 0, world)
    n = min(args.cpu_sample, per)
    threads = os.cpu_count() or 1
    oracle, rs = oracle_sample(B, p, 0, n)
    for _ in range(args.warmup):
        oracle_step(oracle, rs, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        nv, ns = oracle_step(oracle, rs, threads)
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt
    sample = "first %d reads of the workload (%.1f Mb of the synthetic chr20), discoverAndCall, oracle port, %d threads" % (
        n, contig * n / per / 1e6, threads)
    line = {
        "impl": "reference", "metric": "reads genotyped/sec", "value": val, "unit": "reads/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, per, contig, world),
        "cpu_baseline": {"value": val, "unit": "reads/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the Scala/Spark reference cannot run here (no JVM); this is the C++ oracle port of its semantics",
    }
    emit(line)


def config_dict(args, per, contig, world):
    return {"workload": "single-sample germline discoverAndCall, synthetic 60x chr20-shaped (%.0f Mb, %d x 150 bp reads) per GPU"
                        % (contig / 1e6, per),
            "reads_per_gpu": per, "contig_mb_per_gpu": contig / 1e6, "read_len": 150, "ploidy": 2,
            "min_phred_discover": 25, "min_obs_discover": 5,
            "parallelism": "region-sharded x%d, no data-path collective" % world,
            "l2": "inputs (%.1f GB per GPU) are far larger than the 126 MB L2; no flush needed" % (per * 270 / 1e9)}


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly one JSON line: everything libraries print there (NCCL's version banner,
    torchrun notices) is sent to stderr instead."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    claim_stdout()
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    rank, world, local = dist_env()
    cpus_per_rank = max(1, (os.cpu_count() or 8) // world)
    if args.e2e_payload < 0:        # gathering on the host needs cores; reading in place needs none
        args.e2e_payload = 2 if cpus_per_rank >= 16 else 0
    import numpy as np
    import torch
    from avocado_b200 import batch as B
    from avocado_b200.genotyping import Context
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    per, contig = workload(args, world)
    site_cap = max(1 << 18, per // 128)      # rows for discovered sites (about 1 per 360 reads)
    p = synth_for_rank(B, per, contig, rank, world)
    dev = B.synth_device(p, first=rank * per, n=per, device="cuda:%d" % local)
    torch.cuda.synchronize()
    ctx = Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    def step_device():
        return ctx.discover_and_call_packed(dev, ploidy=2, phred=25, min_obs=5, capacity=site_cap,
                                            device_out=True, copy=False)

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    # the pinned host copy of the batch (for the e2e leg) is set-up, done before any timed region so
    # that the two timed regions run back to back under the clock sampler
    host = None if args.no_e2e else ReadBatchPinned(dev, B)

    # host batches only; the single call has the host to itself and gathers with all its threads
    ctx.host_payload = args.e2e_payload
    ctx.gather_threads = max(args.gather_threads, (os.cpu_count() or 8) // max(1, int(os.environ.get("WORLD_SIZE", "1"))))

    def step_host():
        return ctx.discover_and_call_packed(host.batch, ploidy=2, phred=25, min_obs=5,
                                            capacity=site_cap, copy=False)

    # ---- device-resident throughput (value) -----------------------------------------------------
    for _ in range(args.warmup):
        sites, cols = step_device()
    if host is not None:
        for _ in range(max(1, min(args.warmup, 2))):
            step_host()
        for _ in range(2):
            sites, cols = step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tim = []
    e0.record(stream)
    for _ in range(args.steps):
        sites, cols = step_device()
        tim.append(ctx.timing())
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    n_sites_dev = sites.n
    emitted_dev = int(cols["emitted"].sum().item())
    pairs_dev = int(cols["total_cov"].sum().item())     # (read, site) observations that were scored
    t_dev = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if use_dist:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_max = float(t_dev.item())
    total_reads = per * world
    value = total_reads * args.steps / (ms_max / 1e3)

    # ---- roofline of the dominant kernel (this rank) ---------------------------------------------
    # Both tile kernels are sparse: they stream the 32-byte read records (+ operators, mismatch
    # bytes) and gather single sectors of bases / qualities only at variant sites and indels.
    # SURVEY.md 8(d) anti-inflation rule: the full-scan byte count may be claimed only when measured
    # DRAM traffic is >= 0.9 of it; it is not, so `achieved` is measured DRAM bytes (ncu, profiles/)
    # per launch over the kernel's CUDA-event time, and the full-scan figure is given for the step.
    ms_call = statistics.mean(t["ms_call_tiles"] for t in tim)
    ms_disc = statistics.mean(t["ms_discover_tiles"] for t in tim)
    batch_bytes = dev.nbytes()
    n_sites = n_sites_dev
    peak, peak_src = measured_peak()
    # ms_call_tiles spans the call stage's kernels (k_site_buckets, k_call_join, k_call_observe,
    # k_call_reduce); k_discover_tiles is one kernel
    top, top_ms = ("call_kernels", ms_call) if ms_call >= ms_disc else ("k_discover_tiles", ms_disc)
    small_cols = sum(getattr(dev, c).numel() * getattr(dev, c).element_size() for c in ("meta", "ops", "refb"))
    # bytes the algorithm has to touch: discovery = records + operators + mismatch bytes once;
    # call = the record of every read within reach of a site is not known a priori, so the same
    # streaming term is used as an upper bound, plus 64 B per (read, site) pair and the result rows
    alg = small_cols + (n_sites * 16 if top == "k_discover_tiles" else pairs_dev * 64 + n_sites * SITE_ROW_BYTES)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)[top]
        traffic = int(tj["dram_bytes"] * (per / tj["reads"]))
    except Exception:
        pass
    basis = traffic if traffic is not None else alg
    achieved = basis / (top_ms / 1e3) / 1e9
    step_ms = ms_max / args.steps
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "ms_per_launch": top_ms,
                "basis": "measured DRAM bytes per launch (profiles/traffic.json, ncu) / CUDA-event time"
                         if traffic is not None else "algorithmic bytes / CUDA-event time",
                "full_scan_step": {"bytes": batch_bytes, "gbps": batch_bytes / (step_ms / 1e3) / 1e9,
                                   "frac": batch_bytes / (step_ms / 1e3) / 1e9 / peak,
                                   "note": "every packed input byte once (SURVEY.md 8d B_alg) over the WHOLE step "
                                           "(discover + call + set-up): the reads/s figure expressed as the "
                                           "bandwidth a dense scan would need"},
                "note": "sparse kernels: latency / issue bound, not bandwidth bound (profiles/README.md)",
                "other_kernel": {"call_kernels_ms": ms_call, "k_discover_tiles_ms": ms_disc}}

    # ---- end to end through the C ABI with host buffers -------------------------------------------
    # The pinned host batch is cut into region bins (owned interval + read halo, avocado_b200.parallel)
    # and every bin goes through avo_discover_and_call on its own context / stream / host thread, the
    # way executor threads of the JVM shim would: the H2D copies of one bin (PCIe bandwidth bound)
    # overlap the in-place payload gathers of another (PCIe request-latency bound).  The single-call
    # figure (one bin) is reported beside it.
    e2e = None
    if host is not None:
        from concurrent.futures import ThreadPoolExecutor
        from avocado_b200.parallel import plan_region_shards
        K = max(1, args.e2e_bins)
        hb = host.batch
        host_wire = host.wire
        shards = plan_region_shards(hb.meta["start"], hb.meta["end"], K, max_site_len=64)
        bins = [host.pinned_slice(sh.read_lo, sh.read_hi) for sh in shards]
        T = args.e2e_threads or (max(2, min(8, cpus_per_rank // 2)) if args.e2e_payload == 2 else 4)
        T = max(1, min(T, K))
        ctxs = [Context(local) for _ in range(T)]
        for c in ctxs:                 # the call stage's payload blocks are gathered by host threads
            c.host_payload, c.gather_threads = args.e2e_payload, args.gather_threads
        pool = ThreadPoolExecutor(T)

        def run_thread(t, steps=1):   # thread t takes bins t, t + T, ... on its own context / stream
            out = []
            for _step in range(steps):
                out.append([])
                for i in range(t, K, T):
                    w0 = time.perf_counter()
                    s_i, c_i = ctxs[t].discover_and_call_packed(bins[i], ploidy=2, phred=25, min_obs=5,
                                                                capacity=site_cap, copy=False)
                    w1 = time.perf_counter()
                    st = np.asarray(s_i.start[:s_i.n], np.int64)
                    tm = ctxs[t].timing()
                    tm["wall"] = (i, t, w0, w1)
                    out[-1].append((int(((st >= shards[i].lo) & (st < shards[i].hi)).sum()), tm))
            return out

        def step_bins(steps=1):
            """`steps` passes over all bins.  Every host thread runs its bins of successive passes back
            to back, like executor threads working through a stream of partitions: there is no barrier
            between passes, so the H2D phase of one bin overlaps the call stage of another."""
            parts = list(pool.map(lambda t: run_thread(t, steps), range(T)))
            return [[x for part in parts for x in part[k]] for k in range(steps)]

        step_bins(2)
        barrier()
        e0.record(stream)
        t_b = step_bins(args.steps)
        barrier()
        e1.record(stream)
        torch.cuda.synchronize()
        ms_b = e0.elapsed_time(e1)
        assert all(sum(n for n, _ in tb) == n_sites_dev for tb in t_b), "region bins must reproduce the site count"
        # one bin = the plain single call
        barrier()
        e0.record(stream)
        t_h = []
        for _ in range(args.steps):
            s2, c2 = step_host()
            t_h.append(ctx.timing())
        e1.record(stream)
        barrier()
        ms_h = e0.elapsed_time(e1)
        assert s2.n == n_sites_dev
        t_dev = torch.tensor([ms_b, ms_h], dtype=torch.float64, device="cuda")
        if use_dist:
            dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        ms_b, ms_h = float(t_dev[0].item()), float(t_dev[1].item())
        last = [t for _, t in t_b[-1]]
        if os.environ.get("AVO_BENCH_DEBUG"):
            w00 = min(t["wall"][2] for t in last)
            for t in last:
                print("bin %d thread %d: wall %.2f..%.2f  total %.2f h2d %.2f discover %.2f call %.2f d2h %.2f blocks %d" % (
                    t["wall"][0], t["wall"][1], (t["wall"][2] - w00) * 1e3, (t["wall"][3] - w00) * 1e3,
                    t["ms_total"], t["ms_h2d"], t["ms_discover"], t["ms_observe"], t["ms_d2h"], t["gathered_blocks"]),
                    file=sys.stderr)
        e2e = {"value": total_reads * args.steps / (ms_b / 1e3), "unit": "reads/s",
               "h2d_bytes_per_step": int(sum(t["h2d_bytes"] for t in last)) * world,
               "d2h_bytes_per_step": int(sum(t["d2h_bytes"] for t in last)) * world,
               "ms_per_step": ms_b / args.steps,
               "region_bins": K, "host_threads": T,
               "payload_in_place": bool(last[-1]["payload_in_place"]),
               "payload_mode": {0: "copied", 1: "read in place over PCIe",
                                2: "discovery reads in place, call-stage blocks gathered by host threads"}[
                                    int(last[-1]["payload_in_place"])],
               "gathered_blocks_per_step": int(sum(t["gathered_blocks"] for t in last)) * world,
               "gather_threads_per_call": args.gather_threads, "wire_form": ("wire6" if hb.wire6_meta is not None else "wire16") if host_wire else None,
               "single_call": {"value": total_reads * args.steps / (ms_h / 1e3), "ms_per_step": ms_h / args.steps,
                               "ms_h2d": statistics.mean(t["ms_h2d"] for t in t_h),
                               "ms_discover": statistics.mean(t["ms_discover"] for t in t_h),
                               "ms_call": statistics.mean(t["ms_observe"] for t in t_h),
                               "ms_d2h": statistics.mean(t["ms_d2h"] for t in t_h)},
               "note": "pinned host batch through the C ABI (avo_discover_and_call), %d region bins worked through by %d "
                       "host threads (one context + stream each; the K steps are K passes over the bins with no barrier "
                       "between passes, like executor threads on a stream of partitions): records, operators and mismatch "
                       "bytes are copied in the compact wire form, bases and qualities stay in host memory: discovery reads "
                       "the few blocks it needs in place over PCIe, the call stage's blocks are listed by the device, "
                       "gathered by host threads and sent as one dense copy (all inside h2d / d2h_bytes_per_step); D2H of "
                       "the site table and all result columns included" % (K, T)}
        # ---- the same from TEXT records (N1: the packer inside the timed region) -------------------
        # CIGAR / MD / sequence / quality strings of a bounded sample of the bins, in pinned memory.
        # Device route: the text crosses PCIe, avo_pack_reads_device packs it in HBM and
        # avo_discover_and_call runs on the device-resident batch.  Host route beside it:
        # avo_pack_reads_mt on all host threads (pack only; the packed batch would then take the e2e path).
        if not args.no_text:
            n_text_bins, acc = 0, 0
            while n_text_bins < K and acc < args.text_reads:
                acc += bins[n_text_bins].n_reads
                n_text_bins += 1
            keep_alive = []

            def pin(a):
                t = torch.from_numpy(a.view(np.uint8) if a.dtype.names else a).pin_memory()
                keep_alive.append(t)
                return t.numpy().view(a.dtype)
            tcols = []
            for i in range(n_text_bins):
                c = B.text_columns_of(bins[i])
                qo = c.pop("qual_off")
                c = {k: pin(v) for k, v in c.items()}
                c["qual_off"] = c["seq_off"] if qo is not None else None
                tcols.append(c)
            text_bytes = sum(v.nbytes for c in tcols for k, v in c.items() if k != "qual_off")
            reuse = [dict() for _ in range(T)]

            def run_text_thread(t):
                out = []
                for i in range(t, n_text_bins, T):
                    pb = ctxs[t].pack_text_device(tcols[i], out=reuse[t])
                    tp = ctxs[t].timing()
                    s_i, c_i = ctxs[t].discover_and_call_packed(pb, ploidy=2, phred=25, min_obs=5,
                                                                capacity=site_cap, copy=False)
                    out.append((pb.n_reads, s_i.n, tp, ctxs[t].timing()))
                return out

            def step_text():
                return [x for part in pool.map(run_text_thread, range(T)) for x in part]

            for _ in range(2):
                step_text()
            barrier()
            e0.record(stream)
            for _ in range(args.steps):
                t_t = step_text()
            barrier()
            e1.record(stream)
            torch.cuda.synchronize()
            t_txt = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
            if use_dist:
                dist.all_reduce(t_txt, op=dist.ReduceOp.MAX)
            ms_t = float(t_txt.item())
            n_text = sum(x[0] for x in t_t)
            assert n_text == acc
            # host packer, all threads, into reused page-locked output arrays (rank 0 only)
            host_packer = None
            if rank == 0:
                hthreads = os.cpu_count() or 1
                hout = {}
                B.pack_text_columns(tcols[0], threads=hthreads, out=hout)
                hout = {k: pin(v) for k, v in hout.items()}
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    B.pack_text_columns(tcols[0], threads=hthreads, out=hout)
                dt_h = (time.perf_counter() - t0) / args.steps
                t0 = time.perf_counter()
                B.pack_text_columns(tcols[0], threads=1, out=hout)
                dt_1 = time.perf_counter() - t0
                host_packer = {"threads": hthreads, "reads_per_s": tcols[0]["start"].shape[0] / dt_h,
                               "reads_per_s_1_thread": tcols[0]["start"].shape[0] / dt_1,
                               "note": "avo_pack_reads_mt on one bin, pack only"}
            e2e["from_text"] = {
                "value": n_text * world / (ms_t / 1e3), "unit": "reads/s", "ms_per_step": ms_t,
                "sample": "%d of the %d region bins (%d reads), %d host threads" % (n_text_bins, K, n_text, T),
                "h2d_bytes_per_step": int(sum(x[2]["h2d_bytes"] + x[3]["h2d_bytes"] for x in t_t)),
                "d2h_bytes_per_step": int(sum(x[3]["d2h_bytes"] for x in t_t)),
                "text_bytes_per_read": text_bytes / n_text,
                "pack_ms_per_bin": statistics.mean(x[2]["ms_total"] for x in t_t),
                "pack_h2d_ms_per_bin": statistics.mean(x[2]["ms_h2d"] for x in t_t),
                "host_packer": host_packer,
                "note": "AlignmentRecord text columns (CIGAR, MD, sequence, qualities) in pinned host memory -> "
                        "avo_pack_reads_device (H2D + pack in HBM) -> avo_discover_and_call on the device batch -> "
                        "D2H of sites and result columns"}
            del tcols, keep_alive
        pool.shutdown()
        for c in ctxs:
            c.close()
        del host, bins
    clocks = sampler.stop() if rank == 0 else None

    # ---- CPU baseline on a bounded sample (rank 0, N = 1) ------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n = min(args.cpu_sample, per)
        oracle, rs = oracle_sample(B, p, 0, n)
        t0 = time.perf_counter()
        nv, ns = oracle_step(oracle, rs, 1)
        dt = time.perf_counter() - t0
        cpu = {"value": n / dt, "unit": "reads/s", "cores": 1, "kind": "port",
               "sample": "first %d reads of the workload (%.1f Mb), discoverAndCall, %.1f s of CPU work; "
                         "C++ oracle port - the Scala/Spark reference cannot run here (no JVM)" %
                         (n, contig * n / per / 1e6, dt),
               "host_cores": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": "reads genotyped/sec", "value": value, "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(args, per, contig, world),
            "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(sum(t["n_launches"] for t in tim)),
            "roofline": roofline, "cpu_baseline": cpu,
            "stages_ms": {"discover": statistics.mean(t["ms_discover"] for t in tim),
                          "call": statistics.mean(t["ms_observe"] for t in tim),
                          "d2h": statistics.mean(t["ms_d2h"] for t in tim),
                          "call_total": statistics.mean(t["ms_total"] for t in tim)},
            "sites_discovered": int(n_sites), "sites_emitted": emitted_dev,
        }
        emit(line)
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()


class ReadBatchPinned:
    """The rank's batch in pinned host memory (what a JVM shim's direct ByteBuffers look like)."""

    def __init__(self, dev, B):
        import torch
        self.tensors = {}
        b = B.ReadBatch(dev.n_reads, dev.n_blocks, dev.n_ops, dev.n_refb, dev.contig_id)
        for name, dt in B.READ_COLUMNS:
            t = getattr(dev, name)
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t)
            self.tensors[name] = h
            a = h.numpy()
            setattr(b, name, a.view(dt) if a.dtype != dt else a)
        torch.cuda.synchronize()
        b.stats = dev.stats
        self.batch = b
        self.B = B
        self.wire = self.add_wire(b, "all")

    def pinned_slice(self, lo, hi):
        """Reads [lo, hi) as a batch of their own: views of the pinned columns, offsets rebased in a
        pinned copy of the records."""
        import torch
        s = self.batch.slice(lo, hi)
        m = torch.from_numpy(s.meta.view("uint8")).pin_memory()
        self.tensors[("meta", lo, hi)] = m
        s.meta = m.numpy().view(self.B.META_DTYPE)
        self.add_wire(s, (lo, hi))
        return s

    def add_wire(self, batch, key):
        """The packer's compact wire columns (avo_pack_wire), in pinned memory."""
        import torch

        def alloc(n):
            t = torch.empty(n, dtype=torch.uint8, pin_memory=True)
            self.tensors[("wire", key, len(self.tensors))] = t
            return t.numpy()
        return batch.with_wire6(alloc) or batch.with_wire(alloc)


if __name__ == "__main__":
    main()
