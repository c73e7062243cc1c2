// avo_call.cu -- BiallelicGenotyper.call on the GPU: read<->site join, per-read observation and
// site classification, score lookup, per-site reduction and genotype emission.
//
// Reference semantics (file:line under avocado-core/src/main/scala/org/bdgenomics/avocado/):
//   join            util/TreeRegionJoin.scala:74-119, 175-203        -> forest_get (avo_device.cuh)
//   observations    genotyping/Observer.scala:48-140                 -> classify() (avo_observe.cuh)
//   classification  genotyping/BiallelicGenotyper.scala:195-393      -> classify(), observe_read()
//   scoring         genotyping/ScoredObservation.scala:45-90, BG:444-472  -> LUT built by the host
//   reduction       BG:475-500                                        -> k_call_reduce
//   genotype        BG:579-797, util/LogPhred.scala:38-40             -> finalize_site()
//
// Two kernels (DESIGN.md "call"):
//  k_call_join     one thread per read: index test + Forest.get; joined reads go to a list.
//  k_call_observe  one thread per joined read.  Reads are sorted by start, so the reads that can overlap site j
//                  are one contiguous index range [rbase_j, rend_j) (k_site_buckets); site j owns a
//                  bucket of that many 32-bit slots and read r writes its record for site j into
//                  slot r - rbase_j.  Every record therefore has a fixed place that does not depend
//                  on scheduling: no atomics, no ordering constraint, every read is classified
//                  exactly once against all of its sites.
//  k_call_reduce   one THREAD per site scans its bucket in slot order = batch read order, adds the
//                  table rows in fp64 in that order (the oracle's order), and emits the row.
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "avo_device.cuh"
#include "avo_observe.cuh"

namespace avo {

#ifndef AVO_OBS_THREADS
#define AVO_OBS_THREADS 256
#endif
constexpr int OBS_THREADS = AVO_OBS_THREADS;
constexpr int MAX_LOCAL_SITES = 32;  // per-read category cache
#ifndef AVO_CALL_MINB
#define AVO_CALL_MINB 6
#endif

// record word: [0] valid, [2:1] category, [3] forward strand, [10:4] q', [17:11] mq',
//              [21:18] copy number - min ploidy.  bits [21:4] are the row index of the score table.
__device__ __forceinline__ uint32_t make_record(uint32_t cat, bool fwd, int q, int mq, int cni) {
  return 1u | (cat << 1) | ((uint32_t)fwd << 3) | ((uint32_t)q << 4) | ((uint32_t)mq << 11) |
         ((uint32_t)cni << 18);
}

// ---- buckets ------------------------------------------------------------------------------------
// rbase[j] = first read that can overlap site j (start > site.start - max_span),
// len[j]   = reads from there up to the first read starting at or after the site's end.
// first read with start >= key, searched inside the key's bin of the coarse index
__device__ __forceinline__ int32_t first_read_at(const DReads& R, const DIndex& X, int64_t key) {
  if (key <= X.win0) return 0;
  const int32_t bin = idx_bin_floor(X, key);
  int32_t lo = __ldg(X.ridx + bin), hi = __ldg(X.ridx + min(bin + 1, X.nb));
  const int32_t k32 = (int32_t)min(key, (int64_t)INT32_MAX);
  while (lo < hi) {
    const int32_t mid = lo + ((hi - lo) >> 1);
    if (meta_start(R.meta, mid) < k32) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void k_site_buckets(DReads R, DSites S_in, DIndex X, int32_t max_span, int32_t n_grid,
                               int32_t* __restrict__ rbase, uint32_t* __restrict__ len) {
  const DSites S = resolve_sites(S_in);
  const int32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_grid) return;
  const int32_t j = S_in.c_lo + idx;           // the launch covers the host's range (a header table: its capacity)
  if (j < S.c_lo || j >= S.c_hi) { len[j] = 0; return; }     // rows no read of the batch can touch; the scan's sentinel
  const int32_t a = first_read_at(R, X, (int64_t)__ldg(S.start + j) - max_span + 1);
  const int32_t b = max(a, first_read_at(R, X, (int64_t)__ldg(S.end + j)));
  rbase[j] = a;
  len[j] = (uint32_t)(b - a);
}

// Everything one joined read contributes: one record per site of its Forest range [a, b) that it
// observes, written to that site's bucket (BG:195-393 incl. the other-alt pass and copy number).
// pair_blk0 (host-gathered payload only): for site a + k of the read, the slot of the blocks gathered
// for that (read, site) pair in the compact array minus the first block gathered -- what takes the
// place of the read's seq_off while that site is classified.  PAIRS is a template parameter so that
// the device-resident route compiles to a kernel that never looks at pair_blk0.

// read_records calls emit(j, record) for EVERY site j of [a, b), record 0 where the read contributes none.
// The read's record travels in: `ma` (its first 16 bytes), `nmf` (the word holding n_ops | mapq << 16 | flags << 24)
// and w0, its first operator (0 when it has none) -- loaded by the caller as early as it knows the read.
template <bool PAIRS, typename EmitF>
__device__ __forceinline__ void read_records(const DReads& R, const DSites& S, const DCnv& C, const DCallParams& P,
                                             const MetaA& ma, uint32_t nmf, uint32_t w0, int32_t a, int32_t b,
                                             const uint32_t* __restrict__ pair_blk0, EmitF&& emit) {
  const uint32_t n_ops = nmf & 0xFFFFu, flags = nmf >> 24;
  auto none = [&]() { for (int32_t j = a; j < b; ++j) emit(j, 0u); };
  // observeRead throws for such reads (BG:385-391: skipped); no operators -> no observations
  if ((flags & AVO_READ_SKIP_CALL) || n_ops == 0) { none(); return; }
  ReadCtx rc;
  rc.start = ma.start;
  rc.end = ma.end;
  rc.ob = ma.op_off;
  rc.no = n_ops;
  rc.w0 = w0;
  rc.sb = ma.seq_off;
  rc.mapq = (nmf >> 16) & 0xFFu;
  rc.fwd = !(flags & AVO_READ_NEGATIVE_STRAND);
  auto classify_at = [&](int32_t j) {
    if (PAIRS) rc.sb = pair_blk0[j - a];
    return classify_site(R, rc, S, j);
  };

  const int ns = b - a;
  uint16_t cq[MAX_LOCAL_SITES];                           // category | (quality + 1) << 3 (quality -1 = none .. 255)
  const bool cached = ns <= MAX_LOCAL_SITES;
  if (cached) {
    for (int j = 0; j < ns; ++j) {
      const Cls c = classify_at(a + j);
      if (c.cat == CAT_EXC) { none(); return; }           // BG:385-391: the whole read is skipped
      cq[j] = (uint16_t)(c.cat | ((uint32_t)(c.q + 1) << 3));
    }
  } else {
    for (int j = 0; j < ns; ++j)
      if (classify_at(a + j).cat == CAT_EXC) { none(); return; }
  }

  const int mq = min(P.max_mapq, (int)rc.mapq);           // BG:452
  for (int32_t j = a; j < b; ++j) {
    Cls c;
    if (cached) { c.cat = cq[j - a] & 7u; c.q = (int32_t)(cq[j - a] >> 3) - 1; } else c = classify_at(j);
    if (c.cat == CAT_NONE) { emit(j, 0u); continue; }
    const int32_t js = __ldg(S.start + j), je = __ldg(S.end + j);
    if (c.cat == CAT_NONREF) {                            // BG:359-373 other-alt pass
      for (int32_t k = a; k < b; ++k) {
        if (!(__ldg(S.start + k) < je && __ldg(S.end + k) > js)) continue;
        const uint32_t kc = cached ? (cq[k - a] & 7u) : classify_at(k).cat;
        if (kc == CAT_ALT) { c.cat = CAT_OTHER; break; }
      }
    }
    const int q = (c.q < 0) ? P.max_quality : min(P.max_quality, c.q);  // BG:451 least() skips null
    if (q < 1 || mq < 1) { emit(j, 0u); continue; }       // no such row in the score table: dropped
    const int cn = copy_number_for(C, P.base_ploidy, rc.start, rc.end, js, js + __ldg(S.ref_len + j));
    emit(j, make_record(c.cat, rc.fwd, q, mq, cn - P.min_ploidy));
  }
}

template <bool PAIRS>
__device__ void observe_read(const DReads& R, const DSites& S, const DCnv& C, const DCallParams& P,
                             int32_t r, int32_t a, int32_t b, const int32_t* __restrict__ rbase,
                             const uint64_t* __restrict__ boff, uint32_t* __restrict__ buckets,
                             uint64_t n_slots, const uint32_t* __restrict__ pair_blk0) {
  const MetaA ma = load_meta_a(R.meta, r);
  const uint32_t nmf = meta_nmf(R.meta, r);
  const uint32_t w0 = (nmf & 0xFFFFu) ? __ldg(R.ops + ma.op_off) : 0u;
  read_records<PAIRS>(R, S, C, P, ma, nmf, w0, a, b, pair_blk0, [&](int32_t j, uint32_t rec) {
    if (!rec) return;                                     // (the buckets were cleared)
    // n_slots may be a guess carried over from the previous call (the host checks it afterwards)
    const uint64_t at = __ldg(boff + j) + (uint32_t)(r - __ldg(rbase + j));
    if (at < n_slots) buckets[at] = rec;
  });
}

struct JoinedRead { int32_t r, a, b; };

// Join (TreeRegionJoin.scala:175-203): persistent warps stride over blocks of 32 reads.  A cheap
// index test first (most reads have no site in reach), then Forest.get; the joined reads (~15 %) are
// staged in a per-warp shared-memory buffer and appended to a global list JOIN_FLUSH at a time
// (one global atomic per flush; the list order is irrelevant: every record has its own slot).
// No block barrier: a CTA-wide append would put two barriers and a global atomic on the critical
// path of every 512 reads.
// sites that start in the index bins [lo, hi] touches, or an earlier site reaching past lo: [hLo, hHi)
// holds every site with start in [lo, hi]; false when no site can overlap [lo, hi)
__device__ __forceinline__ bool site_hint(const DSites& S, const DIndex& X, int32_t lo, int64_t hi, int32_t& hLo,
                                          int32_t& hHi) {
  hLo = (lo <= X.win0) ? S.c_lo : __ldg(X.sidx + idx_bin_floor(X, lo));
  const int64_t d = hi - X.win0;
  const int64_t k = d < 0 ? 0 : d / IDX_G + 1;
  hHi = (k >= X.nb) ? S.c_hi : __ldg(X.sidx + (int32_t)k);
  return hLo < hHi || (hLo > S.c_lo && __ldg(S.pme + hLo - 1) > lo);
}

constexpr int JOIN_THREADS = 256;
constexpr int JOIN_WARPS = JOIN_THREADS / 32;
constexpr int JOIN_FLUSH = 96;
constexpr int JOIN_CAP = JOIN_FLUSH + 32;

__global__ void __launch_bounds__(JOIN_THREADS)
k_call_join(DReads R, DSites S_in, DIndex X, int32_t max_span, JoinedRead* __restrict__ list,
            int32_t* __restrict__ n_list) {
  __shared__ JoinedRead stage[JOIN_WARPS][JOIN_CAP];
  const DSites S = resolve_sites(S_in);
  // nothing to join; or the table did not fit its arrays (its index points past them: the host reruns)
  if (S.c_hi <= S.c_lo || (S.hdr && S.hdr->n > S.cap)) return;
  const int lane = threadIdx.x & 31;
  JoinedRead* st = stage[threadIdx.x >> 5];
  const unsigned lt = (1u << lane) - 1u;
  int cnt = 0;
  auto flush = [&]() {
    int32_t base = 0;
    if (lane == 0) base = atomicAdd(n_list, cnt);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    for (int i = lane; i < cnt; i += 32) list[base + i] = st[i];
    cnt = 0;
    __syncwarp();
  };
  auto may_join = [&](int32_t lo, int64_t hi, int32_t& hLo, int32_t& hHi) -> bool { return site_hint(S, X, lo, hi, hLo, hHi); };
  const int64_t n_blocks = (R.n + 31) / 32;               // blocks of 32 consecutive reads
  const int64_t n_super = (n_blocks + 31) / 32;           // 32 blocks per warp iteration
  const int64_t stride = (int64_t)gridDim.x * JOIN_WARPS;
  for (int64_t sb = (int64_t)blockIdx.x * JOIN_WARPS + (threadIdx.x >> 5); sb < n_super; sb += stride) {
    // ---- one LANE per block: the reads of a block are neighbours (sorted by start), so the block
    // is tested like one long read [first start, last start + max_span): 4 of 5 blocks have no site
    // in reach at 1.1 sites / kb and are done after two loads
    bool may = false;
    {
      const int64_t blk = sb * 32 + lane;
      if (blk < n_blocks) {
        const int64_t r0 = blk * 32, r1 = min(r0 + 31, R.n - 1);
        int32_t hLo, hHi;
        const int32_t bmin = __ldg(R.coords + r0).x;
        const int64_t bmax = (int64_t)__ldg(R.coords + r1).x + max_span;     // >= every end of the block
        may = may_join(bmin, bmax, hLo, hHi);
        if (may && hHi - hLo <= 8 && !(hLo > S.c_lo && __ldg(S.pme + hLo - 1) > bmin)) {
          // the bins are 256 bp wide: look at the few sites they hold
          may = false;
          for (int32_t k = hLo; k < hHi; ++k)
            may = may || ((int64_t)__ldg(S.start + k) < bmax && __ldg(S.end + k) > bmin);
        }
      }
    }
    unsigned todo = __ballot_sync(0xFFFFFFFFu, may);
    // ---- the other blocks, one lane per read
    while (todo) {
      const int bi = __ffs(todo) - 1;
      todo &= todo - 1;
      const int64_t r = (sb * 32 + bi) * 32 + lane;
      int32_t a = 0, b = 0;
      bool joined = false;
      if (r < R.n) {
        const int2 se = __ldg(R.coords + r);                                // start, end
        int32_t hLo, hHi;
        if (may_join(se.x, (int64_t)se.y, hLo, hHi)) {
          forest_get(S, R.contig, se.x, se.y, hLo, hHi, a, b);
          joined = a < b;                                   // no overlapping variant: read dropped
        }
      }
      const unsigned m = __ballot_sync(0xFFFFFFFFu, joined);
      if (joined) {
        JoinedRead j; j.r = (int32_t)r; j.a = a; j.b = b;
        st[cnt + __popc(m & lt)] = j;
      }
      cnt += __popc(m);
      __syncwarp();
      if (cnt >= JOIN_FLUSH) flush();
    }
  }
  if (cnt > 0) flush();
}

// Observation + classification: one thread per joined read.  GATHERED: the payload is the dense array
// the host gathered (compact), addressed per (read, site) pair through pbase / blk0.
template <bool GATHERED>
__global__ void __launch_bounds__(OBS_THREADS, AVO_CALL_MINB)
k_call_observe(DReads R, DSites S_in, DCnv C, DCallParams P, const JoinedRead* __restrict__ list,
               const int32_t* __restrict__ n_list, const int32_t* __restrict__ rbase,
               const uint64_t* __restrict__ boff, uint32_t* __restrict__ buckets, uint64_t n_slots,
               const uint8_t* __restrict__ compact, const uint32_t* __restrict__ pbase,
               const uint32_t* __restrict__ blk0) {
  const DSites S = resolve_sites(S_in);
  const int32_t n = *n_list;
  if (GATHERED) R.payload = compact;
  for (int32_t i = blockIdx.x * OBS_THREADS + threadIdx.x; i < n; i += gridDim.x * OBS_THREADS) {
    const JoinedRead j = list[i];
    observe_read<GATHERED>(R, S, C, P, j.r, j.a, j.b, rbase, boff, buckets, n_slots,
                           GATHERED ? blk0 + pbase[i] : nullptr);
  }
}

// Join + observation in one pass over the bucket SLOTS (device-resident batches).  Slot s of site j's bucket
// belongs to read r = rbase[j] + (s - boff[j]): the thread of that slot tests the pair, runs Forest.get for
// the read and writes the slot -- the record, or 0 -- so the buckets need no clearing, no list of joined
// reads is built, the 25.6 M reads that touch no site are never looked at, and the lanes of a warp hold
// consecutive reads of ONE site (coalesced records, the same control flow).  A read whose Forest range
// [a, b) has several sites is classified against all of them by each of its (at most PAIR_SITES) slot
// threads (the other-alt pass and the skipped-read rule need the read's other sites); past PAIR_SITES
// the thread of the read's first site writes all of the read's slots and the others stand back.
constexpr int PAIR_SITES = 4;

// Forest.get for a read KNOWN to overlap site j: the fast path of forest_get (avo_device.cuh) found by
// walking from j -- the first site with start >= rs lies at or just before j, the last with start < re
// just after it -- instead of two binary searches; the literal replay where the fast path does not apply.
__device__ __forceinline__ void forest_get_near(const DSites& S, const DIndex& X, int32_t c, int32_t rs, int32_t re,
                                                int32_t j, int32_t& a, int32_t& b) {
  if (__ldg(S.start + j) >= rs) {
    int32_t i0 = j;
    while (i0 > S.c_lo && __ldg(S.start + i0 - 1) >= rs) --i0;
    if (!(i0 > S.c_lo && __ldg(S.pme + i0 - 1) > rs)) {
      a = i0;
      b = j + 1;
      while (b < S.c_hi && __ldg(S.start + b) < re) ++b;
      return;
    }
  }
  int32_t hLo, hHi;
  a = b = 0;
  if (site_hint(S, X, rs, (int64_t)re, hLo, hHi)) forest_get(S, c, rs, re, hLo, hHi, a, b);
}

__global__ void __launch_bounds__(OBS_THREADS, AVO_CALL_MINB)
k_call_pairs(DReads R, DSites S_in, DIndex X, DCnv C, DCallParams P, const int32_t* __restrict__ rbase,
             const uint64_t* __restrict__ boff, uint32_t* __restrict__ buckets, uint64_t n_slots) {
  const DSites S = resolve_sites(S_in);
  // nothing to join; or the table did not fit its arrays (its index points past them: the host reruns)
  if (S.c_hi <= S.c_lo || (S.hdr && S.hdr->n > S.cap)) return;
  const uint64_t total = min((uint64_t)__ldg(boff + S.c_hi), n_slots);
  const uint64_t stride = (uint64_t)gridDim.x * OBS_THREADS;
  const int lane = threadIdx.x & 31;
  // a warp takes 32 consecutive slots per iteration (the loop bound is the same for all of its lanes)
  for (uint64_t s0 = (uint64_t)blockIdx.x * OBS_THREADS + (threadIdx.x & ~31); s0 < total; s0 += stride) {
    // the last site whose bucket starts at or before s0: a 32-way search by the whole warp (three rounds
    // of one load each for 32 k sites, instead of fifteen dependent loads per lane)
    int32_t lo = S.c_lo, hi = S.c_hi;
    while (hi - lo > 1) {
      const int32_t step = (hi - lo + 31) >> 5;
      const int32_t p = lo + lane * step;
      const bool le = p < hi && __ldg(boff + p) <= s0;
      const unsigned m = __ballot_sync(0xFFFFFFFFu, le);          // lanes 0 .. t (boff[lo] <= s0 always)
      const int t = 31 - __clz((int)m);
      lo += t * step;
      hi = min(hi, lo + step);
    }
    const uint64_t s = s0 + lane;
    if (s >= total) continue;
    // the lane's own site: a few steps further (32 slots rarely span more than two or three buckets)
    int32_t j = lo;
    for (int k = 0; k < 4 && j + 1 < S.c_hi && __ldg(boff + j + 1) <= s; ++k) ++j;
    if (j + 1 < S.c_hi && __ldg(boff + j + 1) <= s) {
      int32_t l2 = j + 1, h2 = S.c_hi;                    // many short or empty buckets: finish by bisection
      while (h2 - l2 > 1) {
        const int32_t mid = l2 + ((h2 - l2) >> 1);
        if (__ldg(boff + mid) <= s) l2 = mid; else h2 = mid;
      }
      j = l2;
    }
    const int32_t r = __ldg(rbase + j) + (int32_t)(s - __ldg(boff + j));
    uint32_t rec = 0;
    bool mine = true;                                     // this thread writes slot s
    if (r >= R.n) { buckets[s] = 0; continue; }
    // the read's record and first operator are requested now: nearly every slot's read overlaps its site, and
    // the Forest range below is found while they travel
    const MetaA ma = load_meta_a(R.meta, r);
    const uint32_t nmf = meta_nmf(R.meta, r);
    const int32_t j_end = __ldg(S.end + j), j_start = __ldg(S.start + j);
    const uint32_t w0 = (nmf & 0xFFFFu) ? __ldg(R.ops + ma.op_off) : 0u;
    if (ma.start < j_end && ma.end > j_start) {
      int32_t a, b;
      forest_get_near(S, X, R.contig, ma.start, ma.end, j, a, b);
      if (a <= j && j < b) {
        if (b - a <= PAIR_SITES) {
          read_records<false>(R, S, C, P, ma, nmf, w0, a, b, nullptr, [&](int32_t k, uint32_t v) { if (k == j) rec = v; });
        } else {
          mine = false;
          if (a == j)
            read_records<false>(R, S, C, P, ma, nmf, w0, a, b, nullptr, [&](int32_t k, uint32_t v) {
              const uint64_t at = __ldg(boff + k) + (uint32_t)(r - __ldg(rbase + k));
              if (at < n_slots) buckets[at] = v;
            });
        }
      }
    }
    if (mine) buckets[s] = rec;
  }
}

// ---- host-gathered payload ------------------------------------------------------------------------
// A page-locked host payload can be read in place, one PCIe request per 16-byte block; past a few
// million blocks the link's request rate is the whole cost of a call.  Instead the device lists the
// blocks the joined reads need (k_call_ranges + scan + k_call_blocklist), the host copies them into
// a dense array with its own cores, and one DMA brings them over.
//
// Read indices classify() touches for site [vs, ve): the aligned bases at reference positions in
// [vs, ve) and the bases of insertions at positions in (vs, ve] (avo_observe.cuh): one range of
// blocks per (read, site) pair.  pbase[i] = pairs of the joined reads before i (pair = site of the
// read's Forest range, observed or not).
__device__ bool read_index_range(const DReads& R, uint32_t ob, uint32_t no, int32_t start, int32_t vs, int32_t ve,
                                 uint32_t* lo_out, uint32_t* hi_out) {
  uint32_t lo = 0xFFFFFFFFu, hi = 0;
  bool any = false;
  int32_t pos = start;
  uint32_t ridx = 0;
  for (uint32_t k = 0; k < no; ++k) {
    const uint32_t op = __ldg(R.ops + ob + k);
    const int len = (int)(op >> 4);
    const uint32_t code = op & 15u;
    if (code == AVO_OP_SOFTCLIP) { ridx += len; continue; }
    if (code == AVO_OP_HARDCLIP) continue;
    if (code == AVO_OP_MATCH || code == AVO_OP_MISMATCH) {
      const int32_t l = max(pos, vs), h = min(pos + len, ve);
      if (l < h) {
        lo = min(lo, ridx + (uint32_t)(l - pos));
        hi = max(hi, ridx + (uint32_t)(h - 1 - pos));
        any = true;
      }
      pos += len; ridx += len;
    } else if (code == AVO_OP_INS) {
      if (pos - 1 < ve && pos > vs && len > 0) {
        lo = min(lo, ridx);
        hi = max(hi, ridx + (uint32_t)len - 1);
        any = true;
      }
      ridx += len;
    } else if (code == AVO_OP_DEL) {
      pos += len;
    }
    if (pos > ve) break;
  }
  *lo_out = lo; *hi_out = hi;
  return any;
}

struct PairsOf {
  __host__ __device__ uint32_t operator()(const JoinedRead& j) const { return (uint32_t)(j.b - j.a); }
};

__global__ void __launch_bounds__(256)
k_call_ranges(DReads R, DSites S, const JoinedRead* __restrict__ list, int32_t n, const uint32_t* __restrict__ pbase,
              uint32_t* __restrict__ nblk, uint32_t* __restrict__ clo) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) { nblk[pbase[n]] = 0; return; }     // the scan's last element is the total
  const JoinedRead j = list[i];
  const MetaA ma = load_meta_a(R.meta, j.r);
  const MetaB mb = load_meta_b(R.meta, j.r);
  const bool skip = (mb.flags & AVO_READ_SKIP_CALL) || mb.n_ops == 0;
  // never past the read's own blocks (an operator list longer than the sequence is caught by the
  // packer; this keeps a malformed batch from indexing another read's payload)
  const uint32_t last = mb.l_seq ? (mb.l_seq - 1) / AVO_BLOCK_BASES : 0u;
  for (int32_t s = j.a; s < j.b; ++s) {
    uint32_t cnt = 0, c0 = 0, lo, hi;
    const int32_t vs = __ldg(S.start + s);
    if (!skip && read_index_range(R, ma.op_off, mb.n_ops, ma.start, vs, vs + __ldg(S.ref_len + s), &lo, &hi)) {
      c0 = min(lo / AVO_BLOCK_BASES, last);
      cnt = min(hi / AVO_BLOCK_BASES, last) - c0 + 1;
    }
    const uint32_t p = pbase[i] + (uint32_t)(s - j.a);
    nblk[p] = cnt;
    clo[p] = c0;
  }
}

__global__ void __launch_bounds__(256)
k_call_blocklist(DReads R, const JoinedRead* __restrict__ list, int32_t n, const uint32_t* __restrict__ pbase,
                 const uint32_t* __restrict__ nblk, const uint32_t* __restrict__ slot, const uint32_t* __restrict__ clo,
                 uint32_t* __restrict__ idx, uint32_t* __restrict__ blk0) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const JoinedRead j = list[i];
  const uint32_t seq_off = load_meta_a(R.meta, j.r).seq_off;
  for (uint32_t p = pbase[i]; p < pbase[i] + (uint32_t)(j.b - j.a); ++p) {
    const uint32_t s = slot[p], c0 = clo[p], cnt = nblk[p];
    for (uint32_t k = 0; k < cnt; ++k) idx[s + k] = seq_off + c0 + k;
    blk0[p] = s - c0;  // block c of the read lives at compact[(s - c0) + c]: 32-bit wrap-around is fine
  }
}

// ---- per-site ordered reduction + genotype ---------------------------------------------------------
// One thread per site.  Slots are scanned in order (= batch read order); Spark sums every column
// for every row, the arrays a record does not select receive + 0.0, which leaves them bit-identical,
// so only the selected category's array is touched.
template <int NS>
__global__ void __launch_bounds__(128)
k_call_reduce(DSites S_in, DCallParams P, DResults O, const uint64_t* __restrict__ boff,
              const uint32_t* __restrict__ blen, const uint32_t* __restrict__ buckets, uint64_t n_slots) {
  const DSites S = resolve_sites(S_in);
  const int32_t j = S_in.c_lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (j < S.c_lo || j >= S.c_hi) return;
  const uint32_t* __restrict__ bk = buckets + (size_t)__ldg(boff + j);
  const uint32_t n = __ldg(blen + j);
  if (__ldg(boff + j) + n > n_slots) return;   // guessed capacity too small: the host reruns the stage
  const int ns = P.ns;
  SiteAcc A;
  for (int c = 0; c < 4; ++c)
    for (int g = 0; g < NSMAX; ++g) A.ll[c][g] = 0.0;
  double acc[4][NS];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int g = 0; g < NS; ++g) acc[c][g] = 0.0;
  int32_t c_tot = 0, c_acov = 0, c_ocov = 0, c_afw = 0, c_ofw = 0;
  unsigned long long sq = 0;
  int seen = 0, first_is_ref = 0, first_cn = 0;
  // Four slots per round: their words, then their table rows, are requested together (a thread's chain of
  // slot -> row loads is the kernel's whole latency: 71 k threads are a third of a wave); the sums are still
  // added one record at a time, in slot order.
  constexpr int RU = 4;
  for (uint32_t i0 = 0; i0 < n; i0 += RU) {
    uint32_t ws[RU];
#pragma unroll
    for (int u = 0; u < RU; ++u) ws[u] = (i0 + u < n) ? __ldg(bk + i0 + u) : 0u;
    double vs[RU][NS];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      // row = (copy-number index, mq', q') packed exactly as bits [21:4] of the record word
      const double* __restrict__ row = P.lut + (size_t)((ws[u] >> 4) & 0x3FFFFu) * (uint32_t)ns;
#pragma unroll
      for (int g = 0; g < NS; ++g) vs[u][g] = ((ws[u] & 1u) && g < ns) ? __ldg(row + g) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const uint32_t w = ws[u];
      if (!(w & 1u)) continue;
      const uint32_t cat = (w >> 1) & 3u;
      const uint32_t fwd = (w >> 3) & 1u;
      const uint32_t mq = (w >> 11) & 127u;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (cat == (uint32_t)c) {
#pragma unroll
          for (int g = 0; g < NS; ++g) acc[c][g] += vs[u][g];
        }
      ++c_tot;                                                   // ScoredObservation.scala:88
      const uint32_t is_ref = cat == CAT_REF, is_al = (cat == CAT_ALT) | (cat == CAT_NONREF);
      c_ocov += is_ref; c_ofw += is_ref & fwd;                   // :87, :80
      c_acov += is_al; c_afw += is_al & fwd;                     // :86, :79
      sq += (unsigned long long)(mq * mq);                       // :81
      if (!seen) {                                               // first("isRef"), first("copyNumber")
        seen = 1;
        first_is_ref = (int)is_ref;
        first_cn = (int)((w >> 18) & 15u) + P.min_ploidy;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int g = 0; g < NS; ++g) A.ll[c][g] = acc[c][g];
  A.sq = sq;
  A.allele_fwd = c_afw; A.other_fwd = c_ofw; A.allele_cov = c_acov; A.other_cov = c_ocov; A.total_cov = c_tot;
  A.first_is_ref = first_is_ref; A.copy_number = first_cn; A.seen = seen;
  finalize_site<true>(P, O, j, A);        // a site nobody observed gets a row of zeros: no clearing pass
}

// ---- launchers --------------------------------------------------------------------------------------
// bucket lengths are 32-bit, their prefix sums 64-bit (a batch with one very long read span can need
// more than 2^32 slots; the allocation then fails cleanly instead of the offsets wrapping)
struct LenTo64 {
  __host__ __device__ unsigned long long operator()(uint32_t v) const { return (unsigned long long)v; }
};
size_t call_scan_temp_bytes(int32_t n) {
  size_t a = 0;
  cub::TransformInputIterator<unsigned long long, LenTo64, const uint32_t*> it(nullptr, LenTo64());
  cub::DeviceScan::ExclusiveSum((void*)nullptr, a, it, (unsigned long long*)nullptr, n + 1);
  return a + 256;
}

// rbase / blen / boff: [S.n + 1] scratch.  boff[c_hi] (device) = total number of slots.
cudaError_t launch_call_buckets(const DReads& R, const DSites& S, const DIndex& X, const BatchStats& hstats, int32_t* rbase,
                                uint32_t* blen, uint64_t* boff, void* d_temp, size_t temp_bytes,
                                cudaStream_t s, int* n_launches) {
  const int32_t n = S.c_hi - S.c_lo;      // rows to cover (a header table: its capacity)
  if (n <= 0) return cudaSuccess;
  k_site_buckets<<<(n + 1 + 127) / 128, 128, 0, s>>>(R, S, X, hstats.max_span, n + 1, rbase, blen);
  cub::TransformInputIterator<unsigned long long, LenTo64, const uint32_t*> it(blen + S.c_lo, LenTo64());
  cudaError_t e = cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, it, (unsigned long long*)(boff + S.c_lo), n + 1, s);
  if (e != cudaSuccess) return e;
  if (n_launches) *n_launches += 2;
  return cudaGetLastError();
}

cudaError_t launch_call_join(const DReads& R, const DSites& S, const DIndex& X, int32_t max_span, uint32_t* buckets,
                             size_t n_slots, void* joined_list, int32_t* d_counter, int num_sms, cudaStream_t s,
                             int* n_launches) {
  cudaError_t e = cudaMemsetAsync(d_counter, 0, sizeof(int32_t), s);
  if (e != cudaSuccess) return e;
  if (R.n == 0 || S.c_hi <= S.c_lo) return cudaSuccess;
  if (n_slots) {
    e = cudaMemsetAsync(buckets, 0, n_slots * sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
  }
  int bps = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_call_join, JOIN_THREADS, 0);
  if (e != cudaSuccess) return e;
  // a warp takes 1024 reads per iteration; at least four iterations per warp keep the tail short
  const int64_t n_super = ((R.n + 31) / 32 + 31) / 32;
  const int64_t want = (n_super + 4 * JOIN_WARPS - 1) / (4 * JOIN_WARPS);
  const int grid = (int)max((int64_t)1, min(want, (int64_t)num_sms * max(1, bps)));
  k_call_join<<<grid, JOIN_THREADS, 0, s>>>(R, S, X, max_span, (JoinedRead*)joined_list, d_counter);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

// device-resident batches: join + observation over the bucket slots, then the per-site reduction
cudaError_t launch_call_pairs_reduce(const DReads& R, const DSites& S, const DIndex& X, const DCnv& C, const DCallParams& P,
                                     const DResults& out, const int32_t* rbase, const uint32_t* blen, const uint64_t* boff,
                                     uint32_t* buckets, size_t n_slots, int num_sms, cudaStream_t s, int* n_launches,
                                     cudaEvent_t before_pairs, cudaEvent_t after_pairs) {
  if (R.n == 0 || S.c_hi <= S.c_lo) return cudaSuccess;
  int blocks_per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_call_pairs, OBS_THREADS, 0);
  if (e != cudaSuccess) return e;
  if (before_pairs && (e = cudaEventRecord(before_pairs, s)) != cudaSuccess) return e;
  k_call_pairs<<<num_sms * max(1, blocks_per_sm), OBS_THREADS, 0, s>>>(R, S, X, C, P, rbase, boff, buckets, (uint64_t)n_slots);
  if (after_pairs && (e = cudaEventRecord(after_pairs, s)) != cudaSuccess) return e;
  const int32_t n = S.c_hi - S.c_lo;
  if (P.ns <= 3)
    k_call_reduce<3><<<(n + 127) / 128, 128, 0, s>>>(S, P, out, boff, blen, buckets, (uint64_t)n_slots);
  else
    k_call_reduce<NSMAX><<<(n + 127) / 128, 128, 0, s>>>(S, P, out, boff, blen, buckets, (uint64_t)n_slots);
  if (n_launches) *n_launches += 2;
  return cudaGetLastError();
}

size_t call_gather_temp_bytes(int64_t n) {
  size_t a = 0;
  cub::DeviceScan::ExclusiveSum((void*)nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)(n + 1));
  return a + 256;
}

// pbase: [n_joined + 1] pairs before every joined read; pbase[n_joined] (device) = number of pairs
cudaError_t launch_call_pairs(const void* joined_list, int32_t n_joined, uint32_t* pbase, void* d_temp,
                              size_t temp_bytes, cudaStream_t s, int* n_launches) {
  cub::TransformInputIterator<uint32_t, PairsOf, const JoinedRead*> it((const JoinedRead*)joined_list, PairsOf());
  // the element past the list is never read by the exclusive sum's last output
  cudaError_t e = cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, it, pbase, n_joined + 1, s);
  if (n_launches) *n_launches += 1;
  return e != cudaSuccess ? e : cudaGetLastError();
}

// nblk / slot: [n_pairs + 1], clo: [n_pairs]; slot[n_pairs] (device) = number of blocks to gather
cudaError_t launch_call_gather_plan(const DReads& R, const DSites& S, const void* joined_list, int32_t n_joined,
                                    const uint32_t* pbase, int64_t n_pairs, uint32_t* nblk, uint32_t* slot,
                                    uint32_t* clo, void* d_temp, size_t temp_bytes, cudaStream_t s, int* n_launches) {
  k_call_ranges<<<(n_joined + 1 + 255) / 256, 256, 0, s>>>(R, S, (const JoinedRead*)joined_list, n_joined, pbase, nblk,
                                                          clo);
  cudaError_t e = cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, (const uint32_t*)nblk, slot, (int)(n_pairs + 1), s);
  if (e != cudaSuccess) return e;
  if (n_launches) *n_launches += 2;
  return cudaGetLastError();
}

cudaError_t launch_call_blocklist(const DReads& R, const void* joined_list, int32_t n_joined, const uint32_t* pbase,
                                  const uint32_t* nblk, const uint32_t* slot, const uint32_t* clo, uint32_t* idx,
                                  uint32_t* blk0, cudaStream_t s, int* n_launches) {
  if (n_joined <= 0) return cudaSuccess;
  k_call_blocklist<<<(n_joined + 255) / 256, 256, 0, s>>>(R, (const JoinedRead*)joined_list, n_joined, pbase, nblk, slot,
                                                          clo, idx, blk0);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

// compact / pbase / blk0: the host-gathered payload blocks, the first pair of every joined read and
// every pair's block offset, or NULL
cudaError_t launch_call_observe(const DReads& R, const DSites& S, const DCnv& C, const DCallParams& P,
                                const DResults& out, const int32_t* rbase, const uint32_t* blen,
                                const uint64_t* boff, uint32_t* buckets, size_t n_slots, const void* joined_list,
                                const int32_t* d_counter, const uint8_t* compact, const uint32_t* pbase,
                                const uint32_t* blk0, int num_sms, cudaStream_t s, int* n_launches) {
  if (R.n == 0 || S.c_hi <= S.c_lo) return cudaSuccess;
  int blocks_per_sm = 0;
  auto kernel = compact ? k_call_observe<true> : k_call_observe<false>;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, OBS_THREADS, 0);
  if (e != cudaSuccess) return e;
  kernel<<<num_sms * max(1, blocks_per_sm), OBS_THREADS, 0, s>>>(
      R, S, C, P, (const JoinedRead*)joined_list, d_counter, rbase, boff, buckets, (uint64_t)n_slots, compact, pbase,
      blk0);
  const int32_t n = S.c_hi - S.c_lo;
  if (P.ns <= 3)
    k_call_reduce<3><<<(n + 127) / 128, 128, 0, s>>>(S, P, out, boff, blen, buckets, (uint64_t)n_slots);
  else
    k_call_reduce<NSMAX><<<(n + 127) / 128, 128, 0, s>>>(S, P, out, boff, blen, buckets, (uint64_t)n_slots);
  if (n_launches) *n_launches += 2;
  return cudaGetLastError();
}

}  // namespace avo
