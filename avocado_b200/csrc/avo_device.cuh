// avo_device.cuh -- small device helpers shared by the kernels.
#pragma once

#include "avo_internal.h"

namespace avo {

__device__ __forceinline__ uint32_t nib_at(const uint8_t* __restrict__ p, uint32_t i) {
  uint32_t b = __ldg(p + (i >> 1));
  return (i & 1u) ? (b & 15u) : (b >> 4);
}

// Payload accessors: base / quality i of the read whose first block is blk0.
__device__ __forceinline__ const uint8_t* pl_block(const uint8_t* __restrict__ pl, uint32_t blk0, uint32_t i,
                                                    uint32_t& w) {
  const uint32_t c = i / (uint32_t)AVO_BLOCK_BASES;
  w = i - c * (uint32_t)AVO_BLOCK_BASES;
  return pl + ((size_t)(blk0 + c) << 4);
}
__device__ __forceinline__ uint32_t pl_base(const uint8_t* __restrict__ pl, uint32_t blk0, uint32_t i) {
  uint32_t w;
  const uint8_t* b = pl_block(pl, blk0, i, w);
  const uint32_t v = __ldg(b + (w >> 1));
  return (w & 1u) ? (v & 15u) : (v >> 4);
}
__device__ __forceinline__ uint32_t pl_qual(const uint8_t* __restrict__ pl, uint32_t blk0, uint32_t i) {
  uint32_t w;
  const uint8_t* b = pl_block(pl, blk0, i, w);
  return __ldg(b + 5 + w);
}
// both with one 16-byte load (one memory / PCIe request per observed position)
__device__ __forceinline__ void pl_base_qual(const uint8_t* __restrict__ pl, uint32_t blk0, uint32_t i,
                                             uint32_t& base, uint32_t& qual) {
  uint32_t w;
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(pl_block(pl, blk0, i, w)));
  auto byte_at = [&](uint32_t k) -> uint32_t {
    const uint32_t word = (k < 4u) ? v.x : (k < 8u) ? v.y : (k < 12u) ? v.z : v.w;
    return (word >> (8u * (k & 3u))) & 255u;
  };
  const uint32_t bb = byte_at(w >> 1);
  base = (w & 1u) ? (bb & 15u) : (bb >> 4);
  qual = byte_at(5u + w);
}

// The 32-byte read record is fetched as two 16-byte halves.
struct MetaA { int32_t start, end; uint32_t seq_off, op_off; };
struct MetaB { uint32_t refb_off, n_ops, mapq, flags, qhead, l_seq; };

__device__ __forceinline__ MetaA load_meta_a(const avo_read_meta* __restrict__ m, int64_t r) {
  const int4 v = __ldg(reinterpret_cast<const int4*>(m + r));
  MetaA a; a.start = v.x; a.end = v.y; a.seq_off = (uint32_t)v.z; a.op_off = (uint32_t)v.w;
  return a;
}
__device__ __forceinline__ MetaB load_meta_b(const avo_read_meta* __restrict__ m, int64_t r) {
  const int4 v = __ldg(reinterpret_cast<const int4*>(m + r) + 1);
  MetaB b; b.refb_off = (uint32_t)v.x;
  b.n_ops = (uint32_t)v.y & 0xFFFFu; b.mapq = ((uint32_t)v.y >> 16) & 0xFFu; b.flags = (uint32_t)v.y >> 24;
  b.qhead = (uint32_t)v.z; b.l_seq = (uint32_t)v.w;
  return b;
}
// n_ops | mapq << 16 | flags << 24: the one word of the record's second half the call stage needs
__device__ __forceinline__ uint32_t meta_nmf(const avo_read_meta* __restrict__ m, int64_t r) {
  return (uint32_t)__ldg(reinterpret_cast<const int32_t*>(m + r) + 5);
}
__device__ __forceinline__ int32_t meta_start(const avo_read_meta* __restrict__ m, int64_t r) {
  return __ldg(&m[r].start);
}

// A table assembled earlier in the same launch chain carries its sizes in device memory (DSiteHdr);
// kernels call this first.  Rows beyond the arrays' capacity do not exist for the kernels (the host
// sees the true count in the header and reruns with larger arrays).
__device__ __forceinline__ DSites resolve_sites(DSites S) {
  if (S.hdr) {
    const int32_t n = min(S.hdr->n, S.cap);
    S.n = n; S.midpoint = S.hdr->midpoint;
    // the rows a read of THIS batch can touch: [s_begin, s_end) (a table discovered from the batch itself:
    // all of it; the global table of a sharded run: the neighbourhood of the rank's interval)
    S.c_lo = max(0, min(S.hdr->s_begin, n));
    S.c_hi = max(S.c_lo, min(S.hdr->s_end, n));
  }
  return S;
}

__device__ __forceinline__ int32_t site_contig(const DSites& S, int32_t k) {
  return S.contig ? __ldg(S.contig + k) : 0;
}

// first index in [lo, hi) with a[i] >= v   (a non-decreasing)
__device__ __forceinline__ int32_t lower_bound_i32(const int32_t* __restrict__ a, int32_t lo,
                                                   int32_t hi, int32_t v) {
  while (lo < hi) {
    int32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// bin of position x in the coarse index, clamped to [0, nb]
__device__ __forceinline__ int32_t idx_bin_floor(const DIndex& X, int64_t x) {
  const int64_t d = x - (int64_t)X.win0;
  if (d <= 0) return 0;
  const int64_t k = d / IDX_G;
  return (int32_t)(k > X.nb ? X.nb : k);
}
__device__ __forceinline__ int32_t idx_bin_ceil(const DIndex& X, int64_t x) {
  const int64_t d = x - (int64_t)X.win0;
  if (d <= 0) return 0;
  const int64_t k = (d + IDX_G - 1) / IDX_G;
  return (int32_t)(k > X.nb ? X.nb : k);
}

// first read in [lo, hi) with start >= v   (reads sorted by start)
__device__ __forceinline__ int64_t lower_bound_reads(const avo_read_meta* __restrict__ m, int64_t lo,
                                                     int64_t hi, int32_t v) {
  while (lo < hi) {
    int64_t mid = lo + ((hi - lo) >> 1);
    if (meta_start(m, mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------------
// Forest.get (CORE/util/TreeRegionJoin.scala:74-119) for the read region [rs, re) on contig c.
// Returns the contiguous index range [a, b) of the sorted site table the reference would return.
//
// Fast path: when no site that starts before rs reaches past rs (prefix-max-end test), the
// overlapping set is exactly the sites with start in [rs, re), which is contiguous, and the
// reference's probe sequence is guaranteed to hit it ("join").  Otherwise the
// reference's result depends on which element its power-of-two descent probes first, so the
// descent is replayed literally.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool site_overlaps(const DSites& S, int32_t k, int32_t c, int32_t rs,
                                              int32_t re) {
  return site_contig(S, k) == c && re > __ldg(S.start + k) && rs < __ldg(S.end + k);
}

// [hLo, hHi) is a caller-supplied index range known to contain every site with start in [rs, re]
// (from the coarse index); it only narrows the two lower-bound searches.
__device__ inline void forest_get(const DSites& S, int32_t c, int32_t rs, int32_t re, int32_t hLo,
                                  int32_t hHi, int32_t& a, int32_t& b) {
  int32_t i0 = lower_bound_i32(S.start, hLo, hHi, rs);
  if (!(i0 > S.c_lo && __ldg(S.pme + i0 - 1) > rs)) {
    a = i0;
    b = lower_bound_i32(S.start, i0, hHi, re);
    return;
  }
  // literal replay of binarySearch (:74-92)
  int32_t idx = 0, step = S.midpoint;
  int32_t hit = -1;
  for (;;) {
    if (site_overlaps(S, idx, c, rs, re)) { hit = idx; break; }
    if (step == 0) break;
    int32_t s = idx + step;
    bool stay;
    if (s >= S.n) {
      stay = true;
    } else if (site_overlaps(S, s, c, rs, re)) {
      stay = false;
    } else {
      // rr.compareTo(array(s)) < 0 on (contig, start, end)
      int32_t kc = site_contig(S, s), ks = __ldg(S.start + s), ke = __ldg(S.end + s);
      bool lt = (c != kc) ? (c < kc) : (rs != ks) ? (rs < ks) : (re < ke);
      stay = lt;
    }
    if (!stay) idx = s;
    step >>= 1;
  }
  if (hit < 0) { a = 0; b = 0; return; }
  // expand (:94-105): left from hit, right from hit+1, while overlapping
  int32_t l = hit;
  while (l - 1 >= 0 && site_overlaps(S, l - 1, c, rs, re)) --l;
  int32_t r = hit + 1;
  while (r < S.n && site_overlaps(S, r, c, rs, re)) ++r;
  a = l;
  b = r;
}

}  // namespace avo
