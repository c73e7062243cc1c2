// avo_call.cu -- BiallelicGenotyper.call on the GPU: read<->site join, per-read observation and
// site classification, score lookup, per-site reduction and genotype emission in ONE kernel.
//
// Reference semantics (file:line under avocado-core/src/main/scala/org/bdgenomics/avocado/):
//   join            util/TreeRegionJoin.scala:74-119, 175-203        -> forest_get (avo_device.cuh)
//   observations    genotyping/Observer.scala:48-140                 -> classify()
//   classification  genotyping/BiallelicGenotyper.scala:195-393      -> classify(), process_read()
//   scoring         genotyping/ScoredObservation.scala:45-90, BG:444-472  -> LUT built by the host
//   reduction       BG:475-500                                        -> phase B below
//   genotype        BG:579-797, util/LogPhred.scala:38-40             -> finalize_site()
//
// Layout of the work (DESIGN.md "call kernel"): the genome is cut into windows of CALL_W bases; a
// CTA that takes window w OWNS the sites whose start lies in it, and walks every read that can
// overlap one of them (reads are sorted by start, so that is one contiguous index range, found by
// binary search, including a look-back of max read span).  Because each site is reduced by exactly
// one CTA, in read order, the fp64 likelihood sums are produced in the batch's read order --
// the same order the oracle uses -- and no atomics or cross-CTA combine are needed.
#include "avo_device.cuh"
#include "avo_observe.cuh"

namespace avo {

constexpr int CALL_THREADS = 256;
constexpr int CALL_WARPS = CALL_THREADS / 32;
#ifndef AVO_CALL_W
#define AVO_CALL_W 1024
#endif
#ifndef AVO_CALL_MINB
#define AVO_CALL_MINB 4
#endif
constexpr int CALL_W = AVO_CALL_W;  // window width in bases (one warp per window)
constexpr int SCHUNK = 8;           // owned sites reduced at a time
constexpr int MAX_LOCAL_SITES = 32; // per-read category cache

// record word: [3:0] site slot in chunk, [5:4] category, [6] forward strand, [13:7] q', [20:14] mq',
//              [24:21] copy number - min ploidy
__device__ __forceinline__ uint32_t make_record(int slot, uint32_t cat, bool fwd, int q, int mq, int cni) {
  return (uint32_t)slot | (cat << 4) | ((uint32_t)fwd << 6) | ((uint32_t)q << 7) |
         ((uint32_t)mq << 14) | ((uint32_t)cni << 21);
}

// Join part (TreeRegionJoin): the read's Forest.get range [a, b); false when the read yields
// nothing for the owned site chunk [cLo, cHi).
__device__ __forceinline__ bool join_read(const DReads& R, const DSites& S, int32_t r, int32_t cLo,
                                          int32_t cHi, int32_t hLo, int32_t hHi, int32_t& a, int32_t& b) {
  const int2 se = __ldg(reinterpret_cast<const int2*>(R.meta + r));   // start, end
  forest_get(S, R.contig, se.x, se.y, hLo, hHi, a, b);
  if (a >= b) return false;                               // no overlapping variant: read dropped
  return !(b <= cLo || a >= cHi);                         // nothing of this chunk
}

// Everything one joined read contributes to the owned site chunk [cLo, cHi): returns the number of
// records written to recs[] (at most SCHUNK).
__device__ int classify_read(const DReads& R, const DSites& S, const DCnv& C, const DCallParams& P,
                             int32_t r, int32_t a, int32_t b, int32_t cLo, int32_t cHi,
                             uint32_t* recs) {
  const MetaA ma = load_meta_a(R.meta, r);
  const MetaB mb = load_meta_b(R.meta, r);
  // observeRead throws for such reads (BG:385-391: skipped); no operators -> no observations
  if ((mb.flags & AVO_READ_SKIP_CALL) || mb.n_ops == 0) return 0;
  ReadCtx rc;
  rc.start = ma.start;
  rc.end = ma.end;
  rc.ob = ma.op_off;
  rc.no = mb.n_ops;
  rc.sb = ma.seq_off;
  rc.mapq = mb.mapq;
  rc.fwd = !(mb.flags & AVO_READ_NEGATIVE_STRAND);

  const int ns = b - a;
  uint32_t cats[MAX_LOCAL_SITES];
  int32_t qs[MAX_LOCAL_SITES];
  const bool cached = ns <= MAX_LOCAL_SITES;
  if (cached) {
    for (int j = 0; j < ns; ++j) {
      Cls c = classify_site(R, rc, S, a + j);
      if (c.cat == CAT_EXC) return 0;                     // BG:385-391: the whole read is skipped
      cats[j] = c.cat;
      qs[j] = c.q;
    }
  } else {
    for (int j = 0; j < ns; ++j)
      if (classify_site(R, rc, S, a + j).cat == CAT_EXC) return 0;
  }

  const int mq = min(P.max_mapq, (int)rc.mapq);           // BG:452
  int nrec = 0;
  const int jLo = max(a, cLo), jHi = min(b, cHi);
  for (int32_t j = jLo; j < jHi; ++j) {
    Cls c;
    if (cached) { c.cat = cats[j - a]; c.q = qs[j - a]; } else c = classify_site(R, rc, S, j);
    if (c.cat == CAT_NONE) continue;
    const int32_t js = __ldg(S.start + j), je = __ldg(S.end + j);
    if (c.cat == CAT_NONREF) {                            // BG:359-373 other-alt pass
      for (int32_t k = a; k < b; ++k) {
        if (!(__ldg(S.start + k) < je && __ldg(S.end + k) > js)) continue;
        const uint32_t kc = cached ? cats[k - a] : classify_site(R, rc, S, k).cat;
        if (kc == CAT_ALT) { c.cat = CAT_OTHER; break; }
      }
    }
    const int q = (c.q < 0) ? P.max_quality : min(P.max_quality, c.q);  // BG:451 least() skips null
    if (q < 1 || mq < 1) continue;                        // no such row in the score table: dropped
    const int cn = copy_number_for(C, P.base_ploidy, rc.start, rc.end, js, js + __ldg(S.ref_len + j));
    recs[nrec++] = make_record(j - cLo, c.cat, rc.fwd, q, mq, cn - P.min_ploidy);
  }
  return nrec;
}

constexpr int WREC_CAP = 512;   // records per warp region; >= 32 * SCHUNK guarantees progress
constexpr int QCAP = 64;

// Everything a warp needs in shared memory: the kernel has NO block-level barrier; every warp is
// an independent worker that pulls window tiles from a global counter.
struct __align__(16) WarpSmem {
  SiteAcc acc[SCHUNK];
  uint32_t rec[WREC_CAP];       // records of the current read segment, in read order
  int32_t q_read[QCAP];         // joined reads awaiting classification
  int32_t q_a[QCAP];
  int32_t q_b[QCAP];
};

// Phase A: reads [bLo, bHi) in order.  Lanes first run the join for 32 reads at a time and compact
// the joined ones into the warp's queue; whenever 32 are pending (or the range is exhausted) each
// lane classifies one of them, so that the expensive part runs at full SIMD width.  Records land
// in ws.rec in read order.  Returns the record count, or -1 if the region overflowed.
__device__ __forceinline__ int phase_a_warp(const DReads& R, const DSites& S, const DCnv& C,
                                            const DCallParams& P, WarpSmem& ws, int lane, int32_t bLo,
                                            int32_t bHi, int32_t cLo, int32_t cHi, int32_t hLo,
                                            int32_t hHi) {
  int wc = 0, qn = 0;
  bool ovf = false;
  int32_t base = bLo;
  const unsigned lt = (1u << lane) - 1u;
  while (base < bHi || qn > 0) {
    if (base < bHi) {
      const int32_t r = base + lane;
      int32_t a = 0, b = 0;
      const bool joined = (r < bHi) && join_read(R, S, r, cLo, cHi, hLo, hHi, a, b);
      const unsigned m = __ballot_sync(0xFFFFFFFFu, joined);
      if (joined) {
        const int slot = qn + __popc(m & lt);
        ws.q_read[slot] = r; ws.q_a[slot] = a; ws.q_b[slot] = b;
      }
      qn += __popc(m);
      base += 32;
      __syncwarp();
    }
    if (qn >= 32 || (base >= bHi && qn > 0)) {
      const int take = min(qn, 32);
      uint32_t recs[SCHUNK];
      int nrec = 0;
      if (lane < take)
        nrec = classify_read(R, S, C, P, ws.q_read[lane], ws.q_a[lane], ws.q_b[lane], cLo, cHi, recs);
      int incl = nrec;
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
      }
      const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
      const int off = wc + incl - nrec;
      if (wc + total <= WREC_CAP) {
        for (int i = 0; i < nrec; ++i) ws.rec[off + i] = recs[i];
      } else {
        ovf = true;
      }
      wc += total;
      // pop the classified entries
      const int rest = qn - take;
      int32_t tr = 0, ta = 0, tb = 0;
      if (lane < rest) { tr = ws.q_read[take + lane]; ta = ws.q_a[take + lane]; tb = ws.q_b[take + lane]; }
      __syncwarp();
      if (lane < rest) { ws.q_read[lane] = tr; ws.q_a[lane] = ta; ws.q_b[lane] = tb; }
      qn = rest;
      __syncwarp();
    }
  }
  return ovf ? -1 : wc;
}

// Phase B for one site slot: scan the records in order, accumulate in read order.  Lane
// (cat * ns + g) owns state g of category cat's likelihood array.
__device__ __forceinline__ void phase_b_slot(const DCallParams& P, WarpSmem& ws, int cnt, int slot,
                                             int lane) {
  SiteAcc& A = ws.acc[slot];
  int c_afw = 0, c_ofw = 0, c_acov = 0, c_ocov = 0, c_tot = 0;
  unsigned long long sq = 0;
  const int ns = P.ns;
  const bool ll_lane = lane < 4 * ns;
  const uint32_t mycat = (uint32_t)(lane / ns);
  const int myg = lane - (int)mycat * ns;
  double acc = ll_lane ? A.ll[mycat][myg] : 0.0;
  int seen = A.seen, first_is_ref = A.first_is_ref, first_cn = A.copy_number;
  for (int i0 = 0; i0 < cnt; i0 += 32) {
    const int i = i0 + lane;
    const uint32_t w = (i < cnt) ? ws.rec[i] : 0xFFFFFFFFu;
    const bool mine = (i < cnt) && ((w & 15u) == (uint32_t)slot);
    const unsigned m = __ballot_sync(0xFFFFFFFFu, mine);
    if (m == 0) continue;
    const uint32_t cat = (w >> 4) & 3u;
    const bool fwd = (w >> 6) & 1u;
    const unsigned isRef = __ballot_sync(0xFFFFFFFFu, mine && cat == CAT_REF);
    const unsigned isAl = __ballot_sync(0xFFFFFFFFu, mine && (cat == CAT_ALT || cat == CAT_NONREF));
    const unsigned fw = __ballot_sync(0xFFFFFFFFu, mine && fwd);
    c_tot += __popc(m);                                        // ScoredObservation.scala:88
    c_ocov += __popc(isRef);                                   // :87
    c_acov += __popc(isAl);                                    // :86
    c_ofw += __popc(isRef & fw);                               // :80
    c_afw += __popc(isAl & fw);                                // :79
    if (mine) { const unsigned long long q = (w >> 14) & 127u; sq += q * q; }  // :81
    if (!seen) {                                               // first("isRef"), first("copyNumber")
      const int f = __ffs(m) - 1;
      const uint32_t wf = __shfl_sync(0xFFFFFFFFu, w, f);
      first_is_ref = (((wf >> 4) & 3u) == CAT_REF);
      first_cn = (int)((wf >> 21) & 15u) + P.min_ploidy;
      seen = 1;
    }
    // ordered fp64 accumulation.  Spark sums every column for every row; the arrays a record
    // does not select receive + 0.0, which leaves them bit-identical.
    unsigned mm = m;
    while (mm) {
      const int src = __ffs(mm) - 1;
      mm &= mm - 1;
      const uint32_t x = __shfl_sync(0xFFFFFFFFu, w, src);
      const bool sel = ll_lane && ((x >> 4) & 3u) == mycat;
      // row = (copy-number index, mq', q') packed exactly as bits [24:7] of the record word
      const double v = sel ? __ldg(P.lut + ((x >> 7) & 0x3FFFFu) * (uint32_t)ns + (uint32_t)myg) : 0.0;
      acc += v;
    }
  }
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xFFFFFFFFu, sq, o);
  if (ll_lane) A.ll[mycat][myg] = acc;
  if (lane == 0) {
    A.sq += sq;
    A.allele_fwd += c_afw; A.other_fwd += c_ofw; A.allele_cov += c_acov;
    A.other_cov += c_ocov; A.total_cov += c_tot;
    A.seen = seen; A.first_is_ref = first_is_ref; A.copy_number = first_cn;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(CALL_THREADS, AVO_CALL_MINB)
k_call_tiles(DReads R, DSites S, DCnv C, DCallParams P, DResults O, DIndex X, int32_t n_tiles,
             int32_t max_span, int32_t* __restrict__ tile_counter) {
  __shared__ WarpSmem smem[CALL_WARPS];
  const int lane = threadIdx.x & 31;
  WarpSmem& ws = smem[threadIdx.x >> 5];

  for (;;) {
    int32_t tile = 0;
    if (lane == 0) tile = atomicAdd(tile_counter, 1);
    tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
    if (tile >= n_tiles) break;

    // owned sites: start in [win0 + tile*CALL_W, win0 + (tile+1)*CALL_W).  The first tile also owns
    // the sites that start before the first read but reach into the batch (s_begin), the last tile
    // everything up to the end of the last read (s_end).  Sites outside [s_begin, s_end) cannot
    // overlap any read; their (pre-zeroed) rows stay "not emitted".
    constexpr int BPT = CALL_W / IDX_G;
    const int32_t sLo = (tile == 0) ? __ldg(X.s_range) : __ldg(X.sidx + min(tile * BPT, X.nb));
    const int32_t sHi = (tile == n_tiles - 1) ? __ldg(X.s_range + 1)
                                              : __ldg(X.sidx + min((tile + 1) * BPT, X.nb));

    for (int32_t cLo = sLo; cLo < sHi; cLo += SCHUNK) {
      const int32_t cHi = min(cLo + SCHUNK, sHi);
      for (int i = lane; i < (int)(sizeof(SiteAcc) * SCHUNK / 4); i += 32)
        reinterpret_cast<uint32_t*>(ws.acc)[i] = 0u;
      // reads that can overlap a site of the chunk: start < max site end, start >= first site start - max_span
      int32_t maxEnd = INT32_MIN;
      for (int32_t j = cLo; j < cHi; ++j) maxEnd = max(maxEnd, __ldg(S.end + j));
      const int64_t loKey = (int64_t)__ldg(S.start + cLo) - max_span;
      // over-approximate (by < one index bin each side) read range and site search range
      const int32_t rLo = __ldg(X.ridx + idx_bin_floor(X, loKey));
      const int32_t rHi = __ldg(X.ridx + idx_bin_ceil(X, maxEnd));
      const int32_t hLo = (loKey <= X.win0) ? S.c_lo : __ldg(X.sidx + idx_bin_floor(X, loKey));
      // reads of the range start before maxEnd + IDX_G and end within max_span of their start
      const int32_t hb = idx_bin_ceil(X, (int64_t)maxEnd + max_span) + 1;
      const int32_t hHi = (hb >= X.nb) ? S.c_hi : __ldg(X.sidx + hb);
      __syncwarp();

      // segments of reads: phase A then phase B.  If the record region overflows the segment is
      // halved and redone; a 32-read segment cannot overflow (32 * SCHUNK <= WREC_CAP).
      int32_t seg = max(rHi - rLo, 32);
      int32_t pos = rLo;
      while (pos < rHi) {
        const int32_t n = min(seg, rHi - pos);
        const int cnt = phase_a_warp(R, S, C, P, ws, lane, pos, pos + n, cLo, cHi, hLo, hHi);
        if (cnt < 0) { seg = max(32, seg / 2); continue; }
        __syncwarp();
        if (cnt > 0)
          for (int slot = 0; slot < cHi - cLo; ++slot) phase_b_slot(P, ws, cnt, slot, lane);
        pos += n;
      }
      // emit the chunk's rows
      __syncwarp();
      if (lane < cHi - cLo) finalize_site(P, O, cLo + lane, ws.acc[lane]);
      __syncwarp();
    }
  }
}

cudaError_t launch_call_tiles(const DReads& R, const DSites& S, const DCnv& C, const DCallParams& P,
                              const DResults& out, const DIndex& X, const BatchStats& hstats,
                              int32_t* d_tile_counter, int num_sms, cudaStream_t s, int* n_launches) {
  cudaError_t e = cudaMemsetAsync(d_tile_counter, 0, sizeof(int32_t), s);
  if (e != cudaSuccess) return e;
  if (R.n == 0 || S.c_hi <= S.c_lo) return cudaSuccess;
  const int64_t span = (int64_t)hstats.max_end - (int64_t)hstats.min_start;
  const int32_t n_tiles = (int32_t)max((int64_t)1, (span + CALL_W - 1) / CALL_W);
  int blocks_per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_call_tiles, CALL_THREADS, 0);
  if (e != cudaSuccess) return e;
  // persistent grid; every warp pulls tiles, so a grid needs at most n_tiles / CALL_WARPS blocks
  const int grid = (int)min((int64_t)(n_tiles + CALL_WARPS - 1) / CALL_WARPS,
                            (int64_t)num_sms * max(1, blocks_per_sm));
  k_call_tiles<<<grid, CALL_THREADS, 0, s>>>(R, S, C, P, out, X, n_tiles, hstats.max_span,
                                             d_tile_counter);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

}  // namespace avo
