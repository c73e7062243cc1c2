#!/usr/bin/env python
"""bench.py -- reads genotyped per second for avocado's genotyping hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode MODE] [--reads R]

Default mode: one "step" = BiallelicGenotyper.discoverAndCall (DiscoverVariants, then call against the
discovered sites; the CLI default, CLI/BiallelicGenotyper.scala:255-263) over one synthetic batch of the
BASELINE.json configs[1] shape per GPU: single-sample 60x chr20 (64 Mb), 150 bp reads = 25.6 M reads.

  value     reads/s with the packed batch already resident in HBM (device pointers through the C ABI).
            N = 1: avo_discover_and_call, one launch chain, no host round trip.  N > 1 (torchrun): the
            genome is N times larger, cut into N regions WITH read-overlap halos; every step runs the
            region-sharded path (avo_shard_discover / _call_direct / _finish): an NCCL all-gather of the
            owned slice of every rank's site table; then every rank's last kernel stores the rows of its
            owned sites straight into rank 0's result columns over NVLink peer memory, and only 32-byte
            headers are all-gathered.  Both exchanges are inside the timed region; every rank
            ends with the global table, rank 0 with all rows.  (AVO_ROWS_TO=root|all: NCCL gathers.)
  e2e       the same metric from HOST buffers holding what the reference's job holds: AlignmentRecord
            TEXT columns (CIGAR, MD, sequence, qualities) in pinned memory -> H2D -> device packer
            (avo_pack_reads_device) -> avo_discover_and_call -> D2H of the site table and every result
            column, the batch cut into region bins that a few host threads (one context + stream each)
            pipeline.  `e2e.from_packed` = the same starting from an already packed and wire-encoded
            pinned batch (what a shim that packs on the JVM side would hand over); the cost of that
            packing is reported beside it, not hidden.
  roofline  the longest kernel (k_discover_tiles): measured DRAM bytes per launch (ncu,
            profiles/traffic.json) / its CUDA-event duration, because the kernels are sparse (SURVEY.md
            8d anti-inflation rule); peak = MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle (a C++ restatement of the reference; the Scala/Spark reference cannot
            run here: no JVM) on a bounded sample of the same workload, 1 thread -- and `parity`: its
            rows compared with the GPU's rows of the same reads

Other modes (one JSON line each, `mode` says which; BASELINE.json configs[2]..[4] and the gVCF route):
  --mode trio      3 samples x 30x of one region: union batch -> discovery -> 3 calls -> TrioCaller
  --mode joint     100 samples x 1.5e5 sites, samples sharded over the ranks, two fused all-reduces
  --mode wgs       6 batches of the configs[1] shape per GPU (3.1 Gb / 8 GPUs), streamed through one context
  --mode allsites  scoreAllSites (gVCF rows) at the configs[1] shape
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
SEED = 0xA70CAD0 + 1
# AVO_EVENTS=0 skips the extra `with_events` measurement (the same K steps over the batch after avo_derive_events)
EVENTS = os.environ.get("AVO_EVENTS", "1") != "0"
MIN_MAPQ = 10          # PrefilterReadsArgs.minMappingQuality, the CLI default (applied by the e2e legs that start from text)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="single", choices=["single", "trio", "joint", "wgs", "allsites"])
    ap.add_argument("--reads", type=int, default=READS_C2, help="reads per GPU (default: configs[1])")
    ap.add_argument("--cpu-sample", type=int, default=8_000_000, help="reads in the CPU baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-bins", type=int, default=8, help="region bins the e2e legs are cut into")
    ap.add_argument("--e2e-threads", type=int, default=0,
                    help="host threads (one context + stream each) that work through the e2e bins "
                         "(0 = min(8, host threads / (2 x ranks)))")
    ap.add_argument("--e2e-payload", type=int, default=-1, help="avo_params.host_payload of the from-packed leg: 0 in "
                    "place, 1 copy, 2 in place for discovery + host-gathered blocks for the call stage "
                    "(-1 = 2 when the rank has 16 or more host threads to itself, else 0)")
    ap.add_argument("--gather-threads", type=int, default=2,
                    help="avo_params.gather_threads: host threads gathering payload blocks per call")
    ap.add_argument("--text-bins", type=int, default=0, help="bins of the from-text leg (0 = all that fit in host memory)")
    ap.add_argument("--no-packed", action="store_true", help="skip the from-packed leg of e2e")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def workload(args):
    """Every rank owns `reads` reads over a 64 Mb-per-rank stretch of one genome."""
    per = args.reads
    return per, int(CONTIG_C2 * (per / READS_C2))


def synth_dict(per, contig, world, sample=0):
    """The workload as the generator's parameters: one stateless genome of world*contig bases; rank r owns
    reads [r*per, (r+1)*per), which start in [r*contig, (r+1)*contig).  (The same dict feeds the product's
    generator and the oracle's.)"""
    return dict(seed=SEED, contig_len=contig * world, n_reads_total=per * world, sample=sample)


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


# ---- the CPU arms: the oracle on its own input (nothing of the product is loaded for them) ----------
def load_oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import avocado_oracle as oracle
    return oracle


def oracle_step(oracle, rs, threads):
    v = oracle.discover(rs, 25, 5, threads)
    res = oracle.call(rs, [(c, s, e, r, a) for c, s, e, r, a, _ in v], 2, [], False, 93, 93, threads)
    return v, res


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The Scala/Spark reference
    cannot be built or run in this image (no JVM), so this arm times the oracle port
    with all host threads on a bounded sample per step, from TEXT records (CIGAR / MD parsed every step,
    as the reference does).  Rank 0 only; neither torch nor the product library is loaded."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    per, contig = workload(args)
    n = min(args.cpu_sample, per)
    threads = os.cpu_count() or 1
    oracle = load_oracle()
    # PrefilterReads with the CLI's defaults (mapq > 10), as in the GPU arm's e2e leg
    rs = oracle.synth_reads(synth_dict(per, contig, world), 0, n, "chr20", threads, MIN_MAPQ)
    for _ in range(args.warmup):
        oracle_step(oracle, rs, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(oracle, rs, threads)
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt
    sample = ("first %d reads of the workload (%.1f Mb of the synthetic chr20), PrefilterReads (mapq > %d) + discoverAndCall from "
              "text records, oracle port, %d threads" % (n, contig * n / per / 1e6, MIN_MAPQ, threads))
    emit({
        "impl": "reference", "metric": "reads genotyped/sec", "value": val, "unit": "reads/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, per, contig, world),
        "cpu_baseline": {"value": val, "unit": "reads/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the Scala/Spark reference cannot run here (no JVM); this is the C++ oracle port of its semantics on its own "
                "copy of the workload generator (libavocado_b200.so is not loaded by this arm)",
    })


def config_dict(args, per, contig, world):
    par = ("one GPU" if world == 1 else
           "region-sharded x%d with read-overlap halos: NCCL all-gather of the owned site-table slices; the result rows of the "
           "owned sites are stored by each rank's last kernel straight into rank 0's columns over NVLink peer memory "
           "(CUDA IPC), and only their 32-byte headers are all-gathered; all inside the timed step; every rank ends with the "
           "global table, rank 0 with all rows" % world)
    return {"workload": "single-sample germline discoverAndCall, synthetic 60x chr20-shaped (%.0f Mb, %d x 150 bp reads) per GPU"
                        % (contig / 1e6, per),
            "reads_per_gpu": per, "contig_mb_per_gpu": contig / 1e6, "read_len": 150, "ploidy": 2,
            "min_phred_discover": 25, "min_obs_discover": 5, "parallelism": par,
            "batch": "packed batch (records, operators, mismatch bytes, payload blocks): discovery walks the reads (k_discover_tiles); "
                     "`with_events` repeats the measurement on the batch after avo_derive_events",
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


# ---- parity of the GPU rows with the oracle's, on the reads of the CPU sample ----------------------------
def parity_check(np, sites, cols, want_v, want, limit):
    """GPU rows (whole batch) against oracle rows (first n reads) for the sites the sample decides alone:
    those that end at or before `limit`, the start of the first read outside the sample (every read that
    can touch them is in the sample).  Integers / alleles exact; floating point bit-identical (the product
    sums in the oracle's read order) with the max relative error reported."""
    host = sites.to_host()
    start = np.asarray(host.start[:host.n], np.int64)
    end = start + np.asarray(host.ref_len[:host.n], np.int64)
    rows = np.nonzero(end <= limit)[0]
    keys = {}
    for i in rows:
        ref, alt = host.alleles(int(i))
        keys[(int(start[i]), ref, alt)] = int(i)
    okeys = {(int(s), r, a): j for j, (c, s, e, r, a, n) in enumerate(want_v) if e <= limit}
    out = {"sample_limit": int(limit), "sites": len(okeys), "discovery_identical": set(keys) == set(okeys)}
    counts_ok = all(int(host.count[keys[k]]) == int(want_v[j][5]) for k, j in okeys.items() if k in keys)
    out["counts_identical"] = bool(counts_ok)
    wkey = {(int(s), r, a): j for j, (s, r, a) in enumerate(zip(want["start"], want["referenceAllele"], want["alternateAllele"]))}
    int_cols = {"allele_fwd": "alleleForwardStrand", "other_fwd": "otherForwardStrand", "allele_cov": "alleleCoverage",
                "other_cov": "otherCoverage", "total_cov": "totalCoverage", "first_is_ref": "isRef", "copy_number": "copyNumber",
                "n_alt": "nAlt", "n_ref": "nRef", "n_other": "nOther", "gq": "genotypeQuality", "sb": "strandBiasComponents"}
    fp_cols = {"square_mapq": "squareMapQ", "ref_ll": "referenceLogLikelihoods", "allele_ll": "alleleLogLikelihoods",
               "other_ll": "otherLogLikelihoods", "nonref_ll": "nonRefLogLikelihoods", "qual": "qual",
               "fisher": "fisherStrandBiasPValue", "rms_mapq": "rmsMapQ", "gl": "genotypeLikelihoods",
               "nonref_gl": "nonReferenceLikelihoods"}
    both = [k for k in okeys if k in keys and k in wkey and cols["emitted"][keys[k]]]
    gi = np.array([keys[k] for k in both], np.int64)
    oi = np.array([wkey[k] for k in both], np.int64)
    out["rows_compared"] = len(both)
    out["emitted_identical"] = {k for k in okeys if k in keys and cols["emitted"][keys[k]]} == {k for k in okeys if k in wkey}
    ints_ok, bits_ok, max_rel = True, True, 0.0
    if len(both):
        for c, o in int_cols.items():
            ints_ok &= bool(np.array_equal(np.asarray(cols[c])[gi].astype(np.int64), np.asarray(want[o])[oi].astype(np.int64)))
        for c, o in fp_cols.items():
            g = np.asarray(cols[c])[gi].astype(np.float64)
            w = np.asarray(want[o])[oi].astype(np.float64)
            if g.ndim == 2 and c in ("gl", "nonref_gl"):
                cn = np.asarray(cols["copy_number"])[gi]
                m = np.arange(g.shape[1])[None, :] <= cn[:, None]
                g, w = g[m], w[m]
            fin = np.isfinite(g) & np.isfinite(w)
            bits_ok &= bool(np.array_equal(g, w, equal_nan=True))
            if fin.any():
                rel = np.abs(g[fin] - w[fin]) / np.maximum(np.abs(w[fin]), 1e-300)
                rel[g[fin] == w[fin]] = 0.0
                max_rel = max(max_rel, float(rel.max()))
    out.update(integer_columns_identical=bool(ints_ok), columns_bit_identical=bool(ints_ok and bits_ok), max_rel_err=max_rel,
               tolerance="1e-6 relative for floating point (north_star); integers, alleles and genotypes exact")
    return out


# ---- pinned host copies of a device batch -------------------------------------------------------------
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

    def pinned_slice(self, lo, hi, wire=True):
        """Reads [lo, hi) as a batch of their own: views of the pinned columns, offsets rebased in a
        pinned copy of the records."""
        import torch
        s = self.batch.slice(lo, hi)
        m = torch.from_numpy(s.meta.view("uint8")).pin_memory()
        self.tensors[("meta", lo, hi)] = m
        s.meta = m.numpy().view(self.B.META_DTYPE)
        if wire:
            self.add_wire(s, (lo, hi))
        return s

    def add_wire(self, batch, key):
        """The packer's compact wire columns (avo_pack_wire6 / avo_pack_wire), in pinned memory."""
        import torch

        def alloc(n):
            t = torch.empty(n, dtype=torch.uint8, pin_memory=True)
            self.tensors[("wire", key, len(self.tensors))] = t
            return t.numpy()
        return batch.with_wire6(alloc) or batch.with_wire(alloc)


NEAR_CPUS = None


def bind_near_gpu(torch, local):
    """One process per GPU: run on the cores next to that GPU (NVML's CPU affinity), so that the page-locked buffers
    this rank allocates -- first touch -- live in the memory its GPU reads over the shortest path.  What
    `numactl --cpunodebind` / an executor's core binding does in a deployment; best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        bus = "%08x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


def main():
    claim_stdout()
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.mode != "single":
        return {"trio": run_trio, "joint": run_joint, "wgs": run_wgs, "allsites": run_allsites}[args.mode](args)
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
    global NEAR_CPUS
    NEAR_CPUS = bind_near_gpu(torch, local) if (use_dist and os.environ.get("AVO_BIND", "1") != "0") else None
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    per, contig = workload(args)
    total = per * world
    site_cap = max(1 << 18, per // 128)      # rows for discovered sites (about 1 per 360 reads)
    p = B.synth_params(**synth_dict(per, contig, world))
    # this rank's reads: its own [rank*per, (rank+1)*per) plus, when sharded, the halo -- reads of the
    # neighbours that can reach into its interval (before) or start inside its last sites (after)
    max_span_guess, max_site = 150 + 2 * 20 + 2, 64
    dens = per / contig                      # reads per base
    halo_lo = min(rank * per, int(dens * (max_span_guess + 64)) + 64) if use_dist else 0
    halo_hi = min(total - (rank + 1) * per, int(dens * (max_site + 64)) + 64) if use_dist else 0
    first = rank * per - halo_lo
    dev = B.synth_device(p, first=first, n=per + halo_lo + halo_hi, device="cuda:%d" % local)
    torch.cuda.synchronize()
    ctx = Context(local)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)

    pipe = None
    if use_dist:
        from avocado_b200.parallel import ShardedPipeline
        starts = dev.meta[:dev.n_reads * 32].view(torch.int32).view(-1, 8)[:, 0]
        own_lo = int(starts[halo_lo].item()) if rank > 0 else -(1 << 31) + 1
        own_hi = int(starts[halo_lo + per].item()) if rank + 1 < world else (1 << 31) - 1
        span = int((dev.meta[:dev.n_reads * 32].view(torch.int32).view(-1, 8)[:, 1] - starts).max().item())
        assert rank == 0 or int(starts[0].item()) <= own_lo - span, "halo too short for the batch's longest read span"
        pipe = ShardedPipeline(ctx, rank, world, own_lo, own_hi, site_cap=max(1 << 12, per // 256), allele_cap=max(1 << 14, per // 64),
                               stream=stream, rows_to=os.environ.get("AVO_ROWS_TO", "direct"))

    def step_device():
        if pipe is not None:
            return pipe.step(dev, phred=25, min_obs=5)
        return ctx.discover_and_call_packed(dev, ploidy=2, phred=25, min_obs=5, capacity=site_cap,
                                            device_out=True, copy=False)

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (value) -----------------------------------------------------
    for _ in range(args.warmup):
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
    # (region-sharded steps with direct rows: only rank 0 holds the result columns)
    emitted_dev = int(cols["emitted"].sum().item()) if "emitted" in cols else -1
    pairs_dev = int(cols["total_cov"].sum().item()) if "total_cov" in cols else -1     # (read, site) observations that were scored
    t_dev = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if use_dist:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_max = float(t_dev.item())
    value = total * args.steps / (ms_max / 1e3)
    step_ms = ms_max / args.steps

    # ---- the same steps over the batch WITH its candidate events (reported beside `value`, not as it) -------
    # avo_derive_events walks the reads once and writes every mismatching base and I / D operator as an event;
    # discovery then counts events instead of walking reads.  Deriving costs about as much as one walking
    # discovery, so it pays for a batch that is genotyped more than once (or whose producer emits the events
    # while it packs); its one-off cost is reported here and is NOT part of these steps -- which is why this
    # is not `value`.
    with_events = None
    if EVENTS:
        ctx.derive_events(dev)            # (first call: allocates the event arrays)
        barrier()
        t0 = time.perf_counter()
        ctx.derive_events(dev)
        torch.cuda.synchronize()
        derive_ms = (time.perf_counter() - t0) * 1e3
        if pipe is not None:
            pipe._bs_for = None           # (the pipeline caches the batch's struct: build it again, with the events)
        for _ in range(args.warmup):
            step_device()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tim_e = []
        f0.record(stream)
        for _ in range(args.steps):
            sites_e, cols_e = step_device()
            tim_e.append(ctx.timing())
        f1.record(stream)
        barrier()
        t_e = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device="cuda")
        if use_dist:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        ms_e = float(t_e.item())
        same = bool(sites_e.n == n_sites_dev and ("emitted" not in cols_e or int(cols_e["emitted"].sum().item()) == emitted_dev))
        with_events = {"value": total * args.steps / (ms_e / 1e3), "unit": "reads/s", "ms_per_step": ms_e / args.steps,
                       "derive_ms_once_per_batch": derive_ms, "events_bytes_per_gpu": int(dev.events_nbytes()),
                       "snv_events": int(dev.events["n_snv"]), "indel_events": int(dev.events["n_indel"]),
                       "k_discover_events_ms": statistics.mean(t["ms_discover_tiles"] for t in tim_e),
                       "k_call_pairs_ms": statistics.mean(t["ms_call_pairs"] for t in tim_e),
                       "same_sites_and_emitted_rows": same,
                       "note": "the batch after avo_derive_events (6 B per mismatching base + 8 B per I / D operator, indexed per 32 reads): "
                               "k_discover_events counts them instead of k_discover_tiles walking the reads; the derivation (a walk "
                               "of its own, outside these steps) is why this is reported beside `value` and not as it"}
        dev.events = None
        if pipe is not None:
            pipe._bs_for = None           # (the pipeline cached the batch's struct with the events attached)

    # ---- roofline of the dominant kernel (this rank) ---------------------------------------------
    # Both tile kernels are sparse: they stream the 32-byte read records (+ operators, mismatch
    # bytes) and gather single sectors of bases / qualities only at variant sites and indels.
    # SURVEY.md 8(d) anti-inflation rule: the full-scan byte count may be claimed only when measured
    # DRAM traffic is >= 0.9 of it; it is not, so `achieved` is measured DRAM bytes (ncu, profiles/)
    # per launch over the kernel's CUDA-event time.
    ms_call = statistics.mean(t["ms_call_tiles"] for t in tim)
    ms_disc = statistics.mean(t["ms_discover_tiles"] for t in tim)
    batch_bytes = dev.nbytes()
    peak, peak_src = measured_peak()
    ms_pairs = statistics.mean(t["ms_call_pairs"] for t in tim)
    disc_name = "k_discover_events" if dev.events is not None else "k_discover_tiles"
    # the dominant single kernel: the discovery kernel or k_call_pairs (join + observation), each timed by its own events
    top, top_ms = ("k_call_pairs", ms_pairs) if ms_pairs >= ms_disc else (disc_name, ms_disc)
    small_cols = sum(getattr(dev, c).numel() * getattr(dev, c).element_size() for c in ("meta", "ops", "refb"))
    n_local_sites = n_sites_dev // world
    if top == "k_discover_tiles":
        alg, alg_basis = small_cols + n_local_sites * 16, ("records + operators + mismatch bytes streamed once + 16 B per site "
                                                           "(about 44 B/read; the payload is only touched at indels): DESIGN.md section 4")
    elif top == "k_discover_events":
        alg, alg_basis = dev.events_nbytes() + n_local_sites * 16, "the candidate events streamed once (6 B per mismatching base) + 16 B per site"
    else:
        alg, alg_basis = max(pairs_dev, 0) * 80, ("per scored (read, site) pair: the 32-byte record, ~12 B of operators, one 32-byte "
                                                  "sector holding the payload block, the 4-byte slot (rank 0's pairs)")
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)[top]
        traffic = int(tj["dram_bytes"] * (dev.n_reads / tj["reads"]))
    except Exception:
        pass
    basis = traffic if traffic is not None else alg
    achieved = basis / (top_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "ms_per_launch": top_ms,
                "basis": "measured DRAM bytes per launch (profiles/traffic.json, ncu) / CUDA-event time"
                         if traffic is not None else "algorithmic bytes / CUDA-event time",
                "algorithmic_basis": alg_basis,
                "whole_step": {"ms": step_ms, "note": "all kernels of the step; its DRAM bytes are in profiles/ (ncu launch list)"},
                "note": "sparse kernels: latency / issue bound, not bandwidth bound (profiles/README.md)",
                "other_kernel": {"call_kernels_ms": ms_call, "k_call_pairs_ms": ms_pairs, disc_name + "_ms": ms_disc}}
    try:        # the other large kernel, for the reader who wants both fractions on one line
        other = "k_call_pairs" if top != "k_call_pairs" else disc_name
        other_ms = ms_pairs if other == "k_call_pairs" else ms_disc
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj2 = json.load(fh)[other]
        tr2 = tj2["dram_bytes"] * (dev.n_reads / tj2["reads"])
        roofline["second_kernel"] = {"kernel": other, "ms_per_launch": other_ms, "traffic": int(tr2),
                                     "achieved": tr2 / (other_ms / 1e3) / 1e9, "frac": tr2 / (other_ms / 1e3) / 1e9 / peak}
    except Exception:
        pass
    sharding = None
    if pipe is not None:
        for _ in range(2):                      # outside the timed region: where a step's time goes, per phase
            phases = pipe.step_phases_ms(dev, phred=25, min_obs=5)
        sharding = {"halo_reads_before": halo_lo, "halo_reads_after": halo_hi, "nccl_bytes_received_per_rank_per_step": int(pipe.nccl_bytes),
                    "table_block_bytes": int(pipe.tb), "rows_block_bytes": int(pipe.rb), "collectives_per_step": 2,
                    "rows_to": pipe.rows_to, "phases_ms_rank0": {k: round(v, 4) for k, v in phases.items()},
                    "note": "all_gather_into_tensor of the owned table slices and of 32-byte headers; rows_to=direct: the result rows reach "
                            "rank 0 as peer-memory stores of k_shard_pack_rows (not counted in the NCCL bytes), root / all: NCCL gather / "
                            "all-gather of the rows blocks; one host synchronisation per step"}

    # ---- end to end through the C ABI with host buffers -------------------------------------------
    e2e = None
    if not args.no_e2e:
        own = dev if not use_dist else B.synth_device(p, first=rank * per, n=per, device="cuda:%d" % local)
        e2e = run_e2e(args, np, torch, B, Context, own, per, local, world, rank, cpus_per_rank, site_cap, barrier,
                      n_sites_dev if not use_dist else None)
        del own
    clocks = sampler.stop() if rank == 0 else None

    # ---- CPU baseline on a bounded sample (rank 0, N = 1) + parity of the GPU rows with it -----------
    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n = min(args.cpu_sample, per)
        oracle = load_oracle()
        rs = oracle.synth_reads(synth_dict(per, contig, world), 0, n, "chr20", min(8, os.cpu_count() or 1))
        t0 = time.perf_counter()
        want_v, want = oracle_step(oracle, rs, 1)
        dt = time.perf_counter() - t0
        cpu = {"value": n / dt, "unit": "reads/s", "cores": 1, "kind": "port",
               "sample": "first %d reads of the workload (%.1f Mb), discoverAndCall from text records, %.1f s of CPU work; "
                         "C++ oracle port - the Scala/Spark reference cannot run here (no JVM)" % (n, contig * n / per / 1e6, dt),
               "host_cores": os.cpu_count()}
        # the rows of the same reads: sites that end at or before the start of the first read outside the sample
        limit = int(dev.meta[:dev.n_reads * 32].view(torch.int32).view(-1, 8)[n, 0].item()) if n < per else (1 << 40)
        hcols = {k: v.cpu().numpy() for k, v in cols.items()}
        parity = parity_check(np, sites, hcols, want_v, want, limit)

    if rank == 0:
        line = {
            "metric": "reads genotyped/sec", "value": value, "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(args, per, contig, world),
            "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(sum(t["n_launches"] for t in tim)),
            "roofline": roofline, "with_events": with_events, "cpu_baseline": cpu, "parity": parity, "sharding": sharding,
            "stages_ms": {"discover": statistics.mean(t["ms_discover"] for t in tim),
                          "call": statistics.mean(t["ms_observe"] for t in tim),
                          "d2h": statistics.mean(t["ms_d2h"] for t in tim),
                          "call_total": statistics.mean(t["ms_total"] for t in tim)},
            "sites_discovered": int(n_sites_dev), "sites_emitted": emitted_dev,
        }
        emit(line)
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, np, torch, B, Context, own, per, local, world, rank, cpus_per_rank, site_cap, barrier, want_sites):
    """The end-to-end legs.  The rank's OWN reads (no halo: every rank calls its region on its own, results
    end on its host) in pinned host memory, cut into region bins (owned interval + read halo) that T host
    threads work through on their own context / stream, the way executor threads of the JVM shim would: the
    H2D of one bin overlaps the kernels of another.  K steps = K passes over the bins, no barrier between."""
    from concurrent.futures import ThreadPoolExecutor
    from avocado_b200.parallel import plan_region_shards
    import torch.distributed as dist
    use_dist = world > 1
    host = ReadBatchPinned(own, B)
    K = max(1, args.e2e_bins)
    hb = host.batch
    shards = plan_region_shards(hb.meta["start"], hb.meta["end"], K, max_site_len=64)
    T = args.e2e_threads or max(2, min(8, cpus_per_rank // 2))
    T = max(1, min(T, K))
    ctxs = [Context(local) for _ in range(T)]
    pool = ThreadPoolExecutor(T)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.current_stream()

    def timed_passes(run_thread, steps):
        """`steps` passes over the bins by T threads; returns (ms per step as max over ranks, per-step outputs)."""
        list(pool.map(lambda t: run_thread(t, 2), range(T)))
        barrier()
        e0.record(stream)
        parts = list(pool.map(lambda t: run_thread(t, steps), range(T)))
        barrier()
        e1.record(stream)
        torch.cuda.synchronize()
        t_ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
        if use_dist:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        return float(t_ms.item()), [[x for part in parts for x in part[k]] for k in range(steps)]

    # ---- from TEXT records: the like-for-like leg (what the reference's job holds in memory) -------------
    import psutil
    text_bytes_est = 370 * per
    n_text_bins = args.text_bins or K
    while n_text_bins > 1 and text_bytes_est * n_text_bins / K > 0.35 * psutil.virtual_memory().available / max(1, min(world, 8)):
        n_text_bins -= 1
    keep_alive = []

    def pin(a):
        t = torch.from_numpy(a.view(np.uint8) if a.dtype.names else a).pin_memory()
        keep_alive.append(t)
        return t.numpy().view(a.dtype)

    def text_of(i):
        sl = hb.slice(shards[i].read_lo, shards[i].read_hi)
        c = B.narrow_text_columns(B.text_columns_of(sl))     # 32-bit positions and offsets, 8-bit mapq (avo_text_reads.narrow)
        qo = c.pop("qual_off")
        c = {k: (pin(v) if isinstance(v, np.ndarray) else v) for k, v in c.items()}
        c["qual_off"] = c["seq_off"] if qo is not None else None
        return c
    tcols = list(pool.map(text_of, range(n_text_bins)))
    text_bytes = sum(v.nbytes for c in tcols for k, v in c.items() if k != "qual_off" and isinstance(v, np.ndarray))
    n_text = sum(len(c["start"]) for c in tcols)
    reuse = [dict() for _ in range(T)]
    for c in ctxs:
        c.host_payload, c.gather_threads = 0, args.gather_threads
    from avocado_b200.prefilter import PrefilterReadsArgs, prefilter_struct
    pf = prefilter_struct(PrefilterReadsArgs(minMappingQuality=MIN_MAPQ), "chr20")     # the CLI's defaults

    def run_text_thread(t, steps=1):
        out = []
        for _step in range(steps):
            out.append([])
            for i in range(t, n_text_bins, T):
                pb = ctxs[t].pack_text_device(tcols[i], out=reuse[t], prefilter=pf)
                tp = ctxs[t].timing()

                s_i, c_i = ctxs[t].discover_and_call_packed(pb, ploidy=2, phred=25, min_obs=5, capacity=site_cap, copy=False)
                st = np.asarray(s_i.start[:s_i.n], np.int64)
                out[-1].append((pb.n_reads, int(((st >= shards[i].lo) & (st < shards[i].hi)).sum()), tp, ctxs[t].timing()))
        return out

    ms_t, t_t = timed_passes(run_text_thread, args.steps)
    last = t_t[-1]
    owned_sites = sum(x[1] for x in last)
    kept_reads = sum(x[0] for x in last)
    # the threaded host packer beside it (rank 0; pack only)
    host_packer = None
    if rank == 0:
        hthreads = os.cpu_count() or 1
        hout = {}
        B.pack_text_columns(tcols[0], threads=hthreads, out=hout)
        hout = {k: pin(v) for k, v in hout.items()}
        t0 = time.perf_counter()
        for _ in range(3):
            B.pack_text_columns(tcols[0], threads=hthreads, out=hout)
        dt_h = (time.perf_counter() - t0) / 3
        host_packer = {"threads": hthreads, "reads_per_s": len(tcols[0]["start"]) / dt_h,
                       "note": "avo_pack_reads_mt on one bin, pack only (the alternative to packing on the device)"}
    e2e = {"value": n_text * world / (ms_t / 1e3), "unit": "reads/s", "ms_per_step": ms_t,
           "h2d_bytes_per_step": int(sum(x[2]["h2d_bytes"] + x[3]["h2d_bytes"] for x in last)) * world,
           "d2h_bytes_per_step": int(sum(x[3]["d2h_bytes"] for x in last)) * world,
           "cpus_bound_to": ("%d cores next to the GPU (NVML affinity)" % len(NEAR_CPUS)) if NEAR_CPUS else "all",
           "input": "AlignmentRecord text columns (CIGAR, MD, sequence, phred+33 qualities) in pinned host memory",
           "reads_per_step_per_gpu": n_text, "reads_kept_by_prefilter": kept_reads, "sites_called_per_gpu": owned_sites,
           "region_bins": n_text_bins, "of_bins": K, "host_threads": T,
           "text_bytes_per_read": text_bytes / n_text,
           "pack_ms_per_bin": statistics.mean(x[2]["ms_total"] for x in last),
           "pack_h2d_ms_per_bin": statistics.mean(x[2]["ms_h2d"] for x in last),
           "call_ms_per_bin": statistics.mean(x[3]["ms_total"] for x in last),
           "host_packer": host_packer,
           "note": "pinned text columns -> avo_pack_reads_device (H2D + PrefilterReads' read predicates with the CLI defaults, "
                   "mapq > 10, in the packer kernels + CIGAR/MD merge + "
                   "payload packing in HBM) -> avo_discover_and_call on the device batch -> D2H of sites and every result "
                   "column; %d region bins worked through by %d host threads (one context + stream each; the K steps are K "
                   "passes over the bins with no barrier between passes, like executor threads on a stream of partitions). "
                   "PCIe-bound: %.0f B/read cross the bus." % (n_text_bins, T, text_bytes / n_text)}
    del tcols, keep_alive, reuse

    # ---- from an already packed, wire-encoded pinned batch ---------------------------------------------
    if not args.no_packed:
        t0 = time.perf_counter()
        bins = [host.pinned_slice(sh.read_lo, sh.read_hi) for sh in shards]
        dt_wire = time.perf_counter() - t0
        for c in ctxs:
            c.host_payload, c.gather_threads = args.e2e_payload, args.gather_threads

        def run_thread(t, steps=1):   # thread t takes bins t, t + T, ... on its own context / stream
            out = []
            for _step in range(steps):
                out.append([])
                for i in range(t, K, T):
                    s_i, c_i = ctxs[t].discover_and_call_packed(bins[i], ploidy=2, phred=25, min_obs=5,
                                                                capacity=site_cap, copy=False)
                    st = np.asarray(s_i.start[:s_i.n], np.int64)
                    out[-1].append((int(((st >= shards[i].lo) & (st < shards[i].hi)).sum()), ctxs[t].timing()))
            return out
        ms_b, t_b = timed_passes(run_thread, args.steps)
        lastb = [t for _, t in t_b[-1]]
        if want_sites is not None:
            assert all(sum(n for n, _ in tb) == want_sites for tb in t_b), "region bins must reproduce the site count"
        e2e["from_packed"] = {
            "value": per * world / (ms_b / 1e3), "unit": "reads/s", "ms_per_step": ms_b,
            "h2d_bytes_per_step": int(sum(t["h2d_bytes"] for t in lastb)) * world,
            "d2h_bytes_per_step": int(sum(t["d2h_bytes"] for t in lastb)) * world,
            "payload_mode": {0: "copied", 1: "read in place over PCIe",
                             2: "discovery reads in place, call-stage blocks gathered by host threads"}[int(lastb[-1]["payload_in_place"])],
            "gathered_blocks_per_step": int(sum(t["gathered_blocks"] for t in lastb)) * world,
            "wire_form": "wire6" if bins[0].wire6_meta is not None else ("wire16" if bins[0].wire_meta is not None else None),
            "setup_not_timed": {"slice_and_wire_encode_s": dt_wire,
                                "note": "cutting the packed batch into bins and avo_pack_wire6 on one host thread; packing the "
                                        "records themselves from text is the host_packer figure above"},
            "note": "NOT the like-for-like figure: the batch is already packed (CIGAR/MD merged, bases 4-bit) and wire-encoded in "
                    "pinned memory before the timed region, as a shim that packs on the JVM side would hand it over; records, "
                    "operators and mismatch bytes cross PCIe in the compact wire form, bases and qualities stay in host memory"}
        del bins
    pool.shutdown()
    for c in ctxs:
        c.close()
    return e2e


# ---- other modes ------------------------------------------------------------------------------------------
def _mode_setup(args):
    import torch
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from avocado_b200.genotyping import Context
    ctx = Context(local)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    return torch, rank, world, local, dist, ctx, stream


def _timed(torch, dist, stream, fn, steps, warmup):
    for _ in range(warmup):
        r = fn()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        r = fn()
    e1.record(stream)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), r


def _base_line(args, world, mode, metric, value, unit, ms, workload_text, **extra):
    line = {"mode": mode, "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": workload_text}}
    line.update(extra)
    return line


def run_wgs(args):
    """BASELINE.json configs[4]: 60x whole-genome-shaped (3.1 Gb = 1.24e9 reads) region-partitioned over 8 GPUs =
    six batches of the configs[1] shape per GPU, resident in HBM, streamed through one context."""
    torch, rank, world, local, dist, ctx, stream = _mode_setup(args)
    from avocado_b200 import batch as B
    per, contig = workload(args)
    n_batches = 6
    batches = []
    for b in range(n_batches):
        p = B.synth_params(seed=SEED + 17 * b, contig_len=contig, n_reads_total=per, sample=0)
        batches.append(B.synth_device(p, device="cuda:%d" % local))
    torch.cuda.synchronize()
    site_cap = max(1 << 18, per // 128)

    def step():
        n = 0
        for b in batches:
            s, c = ctx.discover_and_call_packed(b, ploidy=2, phred=25, min_obs=5, capacity=site_cap, device_out=True, copy=False)
            n += s.n
        return n
    ms, n_sites = _timed(torch, dist, stream, step, args.steps, args.warmup)
    if rank == 0:
        reads = per * n_batches * world
        emit(_base_line(args, world, "wgs", "reads genotyped/sec", reads / (ms / 1e3), "reads/s", ms,
                        "60x whole-genome-shaped single sample: %d batches of %.0f Mb / %d reads per GPU (3.1 Gb over 8 GPUs), "
                        "contig-partitioned: no exchange between ranks, %.0f GB of packed reads resident per GPU"
                        % (n_batches, contig / 1e6, per, sum(b.nbytes() for b in batches) / 1e9),
                        scaling="weak", reads_per_step=reads, sites_per_gpu=int(n_sites), batches_per_gpu=n_batches))
    if dist:
        dist.destroy_process_group()


def run_allsites(args):
    """scoreAllSites (BG:223-275) at the configs[1] shape: reference-model rows for every covered position."""
    torch, rank, world, local, dist, ctx, stream = _mode_setup(args)
    from avocado_b200 import batch as B
    per, contig = workload(args)
    p = B.synth_params(**synth_dict(per, contig, world))
    dev = B.synth_device(p, first=rank * per, n=per, device="cuda:%d" % local)
    sites = ctx.discover_packed(dev, phred=25, min_obs=5, device_out=True)
    cols, rstart, rcols = ctx.call_all_sites_packed(dev, sites, device_out=True)      # sizes the output
    cap = len(rstart) + 16
    del cols, rcols
    lib_ms = []

    def step():
        r = ctx.call_all_sites_packed(dev, sites, capacity=cap, device_out=True)
        lib_ms.append(ctx.timing()["ms_total"])
        return len(r[1])
    ms, n_rows = _timed(torch, dist, stream, step, max(1, args.steps // 2), 1)
    ms_lib = statistics.median(lib_ms[1:]) if len(lib_ms) > 1 else lib_ms[0]
    if rank == 0:
        out_bytes = n_rows * SITE_ROW_BYTES
        emit(_base_line(args, world, "allsites", "reads genotyped/sec (scoreAllSites)", per * world / (ms_lib / 1e3), "reads/s", ms_lib,
                        "scoreAllSites at the configs[1] shape: %d reads, %d variant rows, %d reference-model rows per GPU" % (per, sites.n, n_rows),
                        scaling="weak", reference_model_rows=int(n_rows), ms_with_python_wrapper=ms,
                        roofline={"bound": "hbm", "kernel": "k_as_tiles (+ plan kernels)", "unit": "GB/s",
                                  "achieved": (dev.nbytes() + out_bytes) / (ms_lib / 1e3) / 1e9, "peak": measured_peak()[0],
                                  "frac": (dev.nbytes() + out_bytes) / (ms_lib / 1e3) / 1e9 / measured_peak()[0],
                                  "algorithmic_bytes_per_launch": dev.nbytes() + out_bytes, "traffic": None,
                                  "note": "packed batch read once + 208 B per emitted row over the library's event time of the whole call"}))
    if dist:
        dist.destroy_process_group()


def run_joint(args):
    """BASELINE.json configs[3]: 100 samples x 8x low coverage, ~1.5e5 union sites: JointAnnotatorCaller over the
    (samples x sites) matrix, samples sharded over the ranks, the cross-sample per-site statistics all-reduced over
    NVLink: both count vectors as ONE message, then the posterior(0) sums."""
    torch, rank, world, local, dist, ctx, stream = _mode_setup(args)
    import numpy as np
    from avocado_b200.joint import joint_call, joint_call_sharded
    S, N, ns = 100, 150_000, 3
    rng = np.random.default_rng(3)
    lo, hi = rank * S // world, (rank + 1) * S // world
    maf = rng.beta(0.5, 5.0, N)
    n_alt = (rng.random((S, N, 2)) < maf[None, :, None]).sum(2).astype(np.int32)[lo:hi]
    gl = -np.abs(rng.normal(0, 6, (S, N, ns)))[lo:hi]
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    cols = dict(n_sites=N, n_samples=hi - lo, n_states=ns, n_alt=cu(n_alt.ravel()), n_ref=cu((2 - n_alt).ravel()),
                n_other=cu(np.zeros_like(n_alt).ravel()), gl=cu(gl.ravel()))
    with torch.cuda.stream(stream):
        fn = (lambda: joint_call_sharded(ctx, cols)) if world > 1 else (lambda: joint_call(ctx, **cols))
        ms, _ = _timed(torch, dist, stream, fn, max(args.steps, 20), max(args.warmup, 20))
    if rank == 0:
        emit(_base_line(args, world, "joint", "genotypes re-called/sec", S * N / (ms / 1e3), "genotypes/s", ms,
                        "joint calling, %d samples x %d sites (BASELINE configs[3] shape), device-resident genotype likelihoods, "
                        "samples sharded x%d" % (S, N, world), scaling="strong",
                        nccl={"all_reduces_per_step": 2 if world > 1 else 0, "bytes": [2 * N * 4, N * 8] if world > 1 else []},
                        note="avo_joint_counts -> all-reduce(both int32 count vectors as one %d-byte message) -> avo_joint -> "
                             "all-reduce(fp64 posterior sums, %d bytes); times include the Python wrapper's output allocation"
                             % (2 * N * 4, N * 8)))
    if dist:
        dist.destroy_process_group()


def run_trio(args):
    """BASELINE.json configs[2]: trio calling, 3 samples x 30x over one region per GPU (chr1 region-sharded: 248 Mb / N
    per rank; --reads sets the reads per sample per GPU): CLI/TrioGenotyper.scala:216-273 -- the three samples' reads are
    merged, DiscoverVariants runs once over the union, every sample is called against that table, TrioCaller
    post-processes.  Regions are whole intervals of the contig, so no exchange is needed beyond the result gather."""
    torch, rank, world, local, dist, ctx, stream = _mode_setup(args)
    from avocado_b200 import batch as B
    from avocado_b200.trio import trio_call_packed
    per = args.reads if args.reads != READS_C2 else 6_200_000          # 31 Mb x 30x = chr1 / 8 per sample
    contig = int(per * 150 / 30)
    parts = []
    for s in range(3):
        p = B.synth_params(seed=SEED + 5, contig_len=contig * world, n_reads_total=per * world, sample=s + 1)
        parts.append(B.synth_device(p, first=rank * per, n=per, device="cuda:%d" % local))
    # the three samples' columns back to back in one allocation (as a shim uploading into one buffer would
    # leave them): the union batch then points at the same memory and only the records are merged
    arena = {name: torch.cat([getattr(b, name)[:n] for b, n in zip(parts, sizes)] + [getattr(parts[0], name)[:0].new_zeros(16)])
             for name, sizes in (("payload", [b.n_blocks * 16 for b in parts]), ("ops", [b.n_ops for b in parts]),
                                 ("refb", [b.n_refb for b in parts]))}
    samples, o = [], dict(payload=0, ops=0, refb=0)
    for b in parts:
        v = B.ReadBatch(b.n_reads, b.n_blocks, b.n_ops, b.n_refb, b.contig_id, meta=b.meta, stats=b.stats)
        for name, n in (("payload", b.n_blocks * 16), ("ops", b.n_ops), ("refb", b.n_refb)):
            setattr(v, name, arena[name][o[name]:o[name] + n])
            o[name] += n
        samples.append(v)
    samples[0]._arena = arena
    del parts
    torch.cuda.synchronize()
    cap = max(1 << 17, per // 64)
    stage = {}

    def step():
        with torch.cuda.stream(stream):
            t0 = time.perf_counter()
            union = B.merge_batches(samples, ctx=ctx)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            table = ctx.discover_packed(union, phred=25, min_obs=5, capacity=cap, device_out=True)
            t2 = time.perf_counter()
            cnt = {}
            for s, b in enumerate(samples):
                cols = ctx.call_packed(b, table, device_out=True)
                for k in ("n_alt", "n_ref", "n_other", "emitted"):
                    cnt.setdefault(k, []).append(cols[k][:table.n])
            t3 = time.perf_counter()
            cat = {k: torch.cat(v) for k, v in cnt.items()}
            out = trio_call_packed(ctx, table.n, cat["n_alt"], cat["n_ref"], cat["n_other"], cat["emitted"])
            torch.cuda.synchronize(); t4 = time.perf_counter()
            stage.update(merge=(t1 - t0) * 1e3, discover=(t2 - t1) * 1e3, call_x3=(t3 - t2) * 1e3, trio=(t4 - t3) * 1e3)
            return table.n, int(out["kept"].sum().item())
    ms, (n_sites, kept) = _timed(torch, dist, stream, step, max(2, args.steps // 2), 2)
    if rank == 0:
        reads = 3 * per * world
        emit(_base_line(args, world, "trio", "reads genotyped/sec (trio)", reads / (ms / 1e3), "reads/s", ms,
                        "trio calling: 3 samples x 30x, %.0f Mb region and %d reads per sample per GPU (chr1-shaped when sharded x8), "
                        "union discovery -> 3 calls -> TrioCaller" % (contig / 1e6, per),
                        scaling="weak", reads_per_step=reads, sites_per_gpu=int(n_sites), trio_sites_kept_per_gpu=int(kept),
                        stages_ms=stage,
                        note="merge = sorted union of the three samples' RECORDS (avo_merge_meta, one kernel; bases, qualities and "
                             "operators stay where they are: the samples lie back to back in one allocation); the reference's RDD union "
                             "is free, a sorted batch is this design's requirement; stage times are host clocks around synchronised calls"))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
