// avo_allsites.cu -- the reference-model rows of scoreAllSites (gVCF mode).
//
// Reference semantics (avocado-core/.../genotyping/BiallelicGenotyper.scala:223-275, 299-306): in
// scoreAllSites mode every observation of a joined read (Observer.scala:72-138: one per aligned
// base, one per insertion anchored on the base to its left, one per deletion) that overlaps none
// of the read's variants becomes a site of its own, DiscoveredVariant(rr) = (contig, rr.start,
// "N", no alternate), end = start + 1 (DiscoveredVariant.scala:59-61); the observation is scored
// as REF when it is a sequence match and as a plain (ALT-flagged) observation otherwise, and
// the rows of one position are summed like any other site (BG:475-500).  Observations that do
// overlap a variant go to that variant exactly as in variant-only mode (avo_call.cu).
//
// Work layout: one thread per reference position (a warp owns 32 consecutive positions); the
// warp walks the joined reads that overlap its positions IN READ ORDER, every lane keeping the
// sums of its own position in registers, so the fp64 sums come out in the batch's read order
// with no atomics.  The operator walk is warp-uniform (all lanes walk the same read); base and
// quality loads are coalesced (adjacent lanes = adjacent bases).
#include <cub/device/device_scan.cuh>

#include "avo_observe.cuh"

namespace avo {

constexpr int AS_THREADS = 256;
constexpr int AS_WIN = 256;          // positions per row-offset window (one popcount + scan entry)

struct JoinedRead { int32_t r, a, b; };

// index range of the site table that contains every site overlapping [rs, re)'s start range
__device__ __forceinline__ void site_hint_range(const DIndex& X, const DSites& S, int32_t rs, int32_t re,
                                                int32_t& hLo, int32_t& hHi) {
  hLo = (rs <= X.win0) ? S.c_lo : __ldg(X.sidx + idx_bin_floor(X, rs));
  const int64_t d = (int64_t)re - X.win0;
  const int64_t k = d < 0 ? 0 : d / IDX_G + 1;
  hHi = (k >= X.nb) ? S.c_hi : __ldg(X.sidx + (int32_t)k);
}

// TreeRegionJoin + the whole-read drop of BG:385-391: is read r joined, and to which table range?
__device__ bool as_join(const DReads& R, const DSites& S, const DIndex& X, int32_t r, int32_t& a, int32_t& b) {
  const MetaA ma = load_meta_a(R.meta, r);
  const MetaB mb = load_meta_b(R.meta, r);
  if ((mb.flags & AVO_READ_SKIP_CALL) || mb.n_ops == 0) return false;
  int32_t hLo, hHi;
  site_hint_range(X, S, ma.start, ma.end, hLo, hHi);
  forest_get(S, R.contig, ma.start, ma.end, hLo, hHi, a, b);
  if (a >= b) return false;
  // leadBaseObserved(0) on an empty allele throws for an insertion variant next to a deletion:
  // the reference then skips the whole read
  ReadCtx rc;
  rc.start = ma.start; rc.end = ma.end; rc.ob = ma.op_off; rc.no = mb.n_ops; rc.sb = ma.seq_off;
  rc.mapq = mb.mapq; rc.fwd = true;
  rc.w0 = __ldg(R.ops + ma.op_off);
  for (int32_t j = a; j < b; ++j) {
    if (__ldg(S.ref_len + j) == 1 && __ldg(S.alt_len + j) > 1 && classify_site(R, rc, S, j).cat == CAT_EXC)
      return false;
  }
  return true;
}

__global__ void k_as_flags(DReads R, DSites S, DIndex X, int32_t* __restrict__ flag) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= R.n) return;
  int32_t a, b;
  flag[r] = as_join(R, S, X, (int32_t)r, a, b) ? 1 : 0;
}

__device__ __forceinline__ void set_bits(uint32_t* __restrict__ bm, int64_t lo, int64_t hi) {   // [lo, hi)
  for (int64_t w = lo >> 5; w <= (hi - 1) >> 5; ++w) {
    const int64_t b0 = max(lo, w << 5) - (w << 5), b1 = min(hi, (w + 1) << 5) - (w << 5);
    const uint32_t m = (b1 - b0 == 32) ? 0xFFFFFFFFu : (((1u << (b1 - b0)) - 1u) << b0);
    atomicOr(bm + w, m);
  }
}

// joined list (in read order) + coverage bitmap of the positions that may receive a row
__global__ void k_as_scatter(DReads R, DSites S, DIndex X, const int32_t* __restrict__ flag,
                             const int32_t* __restrict__ jpos, JoinedRead* __restrict__ jr,
                             int2* __restrict__ jcoord, uint32_t* __restrict__ cover, int64_t p0, int64_t nbits) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= R.n || !flag[r]) return;
  int32_t a, b;
  as_join(R, S, X, (int32_t)r, a, b);
  JoinedRead j; j.r = (int32_t)r; j.a = a; j.b = b;
  jr[jpos[r]] = j;
  const MetaA ma = load_meta_a(R.meta, r);
  jcoord[jpos[r]] = make_int2(ma.start, ma.end);     // dense (start, end) of the joined reads
  // observations are keyed at [start - 1, end): an insertion before the first aligned base is
  // anchored one position to the left of the read
  const int64_t lo = max((int64_t)ma.start - 1 - p0, (int64_t)0), hi = min((int64_t)ma.end - p0, nbits);
  if (lo < hi) set_bits(cover, lo, hi);
}

// bitmap of the positions covered by some site of the batch's contig (fast path of the
// "overlaps no variant" test)
__global__ void k_as_site_cover(DSites S, uint32_t* __restrict__ bm, int64_t p0, int64_t nbits) {
  const int32_t k = S.c_lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= S.c_hi) return;
  const int64_t lo = max((int64_t)__ldg(S.start + k) - p0, (int64_t)0), hi = min((int64_t)__ldg(S.end + k) - p0, nbits);
  if (lo < hi) set_bits(bm, lo, hi);
}

__global__ void k_as_window_counts(const uint32_t* __restrict__ cover, int32_t n_win, int32_t* __restrict__ counts) {
  const int32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_win) return;
  int c = 0;
  for (int i = 0; i < AS_WIN / 32; ++i) c += __popc(cover[(int64_t)w * (AS_WIN / 32) + i]);
  counts[w] = c;
}

// jfine[k] = first joined read whose start is >= p0 + 32 k   (k = 0..n_fine; jfine[n_fine] = n_joined).
// The joined list is in read order = sorted by start: read j fills the entries between its
// predecessor's start and its own.
__global__ void k_as_jfine(const int2* __restrict__ jcoord, int32_t n_joined, int64_t p0, int64_t n_fine,
                           int32_t* __restrict__ jfine) {
  const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n_joined) return;
  const int64_t b_prev = (j == 0) ? -1 : min(((int64_t)jcoord[j - 1].x - p0) >> 5, n_fine);
  const int64_t b_cur = (j == n_joined) ? n_fine : min(((int64_t)jcoord[j].x - p0) >> 5, n_fine);
  for (int64_t b = b_prev + 1; b <= b_cur; ++b) jfine[b] = j;
}

template <int NS>
__global__ void __launch_bounds__(AS_THREADS)
k_as_tiles(DReads R, DSites S, DCnv C, DCallParams P, DResults O, int32_t* __restrict__ o_start, DIndex X,
           const JoinedRead* __restrict__ jr, const int2* __restrict__ jcoord, const int32_t* __restrict__ jfine,
           const uint32_t* __restrict__ cover, const uint32_t* __restrict__ site_cover,
           const int32_t* __restrict__ win_base, int64_t p0, int64_t nbits, int32_t max_span) {
  const int lane = threadIdx.x & 31;
  const int64_t wbit = ((int64_t)blockIdx.x * AS_THREADS + (threadIdx.x & ~31));   // first bit of the warp
  if (wbit >= nbits) return;
  const uint32_t cov = cover[wbit >> 5];
  if (cov == 0u) return;                                   // no joined read reaches these positions
  const uint32_t scov = site_cover[wbit >> 5];
  const int64_t wlo = p0 + wbit;                           // reference position of lane 0
  const int32_t p = (int32_t)(wlo + lane);
  const int ns = P.ns;

  double acc_ref[NS], acc_alt[NS];
#pragma unroll
  for (int g = 0; g < NS; ++g) { acc_ref[g] = 0.0; acc_alt[g] = 0.0; }
  int32_t c_tot = 0, c_acov = 0, c_ocov = 0, c_afw = 0, c_ofw = 0;
  unsigned long long sq = 0;
  int seen = 0, first_is_ref = 0, first_cn = 0;

  // one observation of read (rs, re, [a, b)) at region [p, p + w): BG:254-262 + 299-306 + 444-500
  auto observe = [&](int32_t rs, int32_t re, int32_t a, int32_t b, int32_t w, bool maybe_site, bool is_ref,
                     int32_t q, uint32_t mapq, bool fwd) {
    if (maybe_site) {                                      // overlaps one of the read's variants?
      for (int32_t k = a; k < b; ++k)
        if (p < __ldg(S.end + k) && p + w > __ldg(S.start + k)) return;   // it belongs to that variant
    }
    const int qq = (q < 0) ? P.max_quality : min(P.max_quality, q);       // BG:451 least() skips null
    const int mq = min(P.max_mapq, (int)mapq);                             // BG:452
    if (qq < 1 || mq < 1) return;                          // no such row in the score table: dropped
    const int cn = copy_number_for(C, P.base_ploidy, rs, re, p, p + 1);
    const double* row = P.lut + ((((size_t)(cn - P.min_ploidy) * 128 + mq) * 128) + qq) * ns;
    if (is_ref) {
#pragma unroll
      for (int g = 0; g < NS; ++g) if (g < ns) acc_ref[g] += __ldg(row + g);
      ++c_ocov; c_ofw += fwd;
    } else {
#pragma unroll
      for (int g = 0; g < NS; ++g) if (g < ns) acc_alt[g] += __ldg(row + g);
      ++c_acov; c_afw += fwd;
    }
    ++c_tot;
    sq += (unsigned long long)(mq * mq);
    if (!seen) { seen = 1; first_is_ref = is_ref; first_cn = cn; }
  };

  // joined reads that can reach the warp's positions: start in (wlo - max_span, wlo + 32]
  const int64_t fb = (wbit - (int64_t)max_span - 1) >> 5;
  const int32_t jLo = __ldg(jfine + (fb < 0 ? 0 : fb));
  const int32_t jHi = __ldg(jfine + min((wbit >> 5) + 2, nbits >> 5));
  for (int32_t j = jLo; j < jHi; ++j) {
    const int2 se = __ldg(jcoord + j);
    if ((int64_t)se.y <= wlo || (int64_t)se.x - 1 >= wlo + 32) continue;          // warp-uniform
    const JoinedRead J = jr[j];
    const MetaA ma = load_meta_a(R.meta, J.r);
    const MetaB mb = load_meta_b(R.meta, J.r);
    const bool fwd = !(mb.flags & AVO_READ_NEGATIVE_STRAND);
    int32_t pos = ma.start;
    uint32_t ridx = 0;
    for (uint32_t k = 0; k < mb.n_ops; ++k) {              // warp-uniform walk (Observer.scala:62-139)
      const uint32_t op = __ldg(R.ops + ma.op_off + k);
      const int32_t len = (int32_t)(op >> 4);
      const uint32_t code = op & 15u;
      if (code == AVO_OP_SOFTCLIP) { ridx += len; continue; }
      if (code == AVO_OP_HARDCLIP) continue;
      if (code == AVO_OP_MATCH || code == AVO_OP_MISMATCH) {
        if (p >= pos && p < pos + len) {
          observe(ma.start, ma.end, J.a, J.b, 1, (scov >> lane) & 1u, code == AVO_OP_MATCH,
                  (int32_t)pl_qual(R.payload, ma.seq_off, ridx + (uint32_t)(p - pos)), mb.mapq, fwd);
        }
        pos += len; ridx += len;
      } else if (code == AVO_OP_INS) {
        if (p == pos - 1)
          observe(ma.start, ma.end, J.a, J.b, 1, (scov >> lane) & 1u, false,
                  ins_quality(R.payload, ma.seq_off, ridx, len), mb.mapq, fwd);
        ridx += len;
      } else if (code == AVO_OP_DEL) {
        if (p == pos) observe(ma.start, ma.end, J.a, J.b, len, true, false, -1, mb.mapq, fwd);
        pos += len;
      }
      if ((int64_t)pos - 1 >= wlo + 32) break;             // nothing later reaches the warp
    }
  }

  if (!((cov >> lane) & 1u)) return;
  const int64_t win = wbit / AS_WIN;
  int32_t row = win_base[win];
  for (int64_t w = win * (AS_WIN / 32); w < (wbit >> 5); ++w) row += __popc(cover[w]);
  row += __popc(cov & ((1u << lane) - 1u));
  o_start[row] = p;
  SiteAcc A;
  for (int g = 0; g < NSMAX; ++g) { A.ll[CAT_REF][g] = 0.0; A.ll[CAT_ALT][g] = 0.0; A.ll[CAT_OTHER][g] = 0.0; A.ll[CAT_NONREF][g] = 0.0; }
#pragma unroll
  for (int g = 0; g < NS; ++g) { A.ll[CAT_REF][g] = acc_ref[g]; A.ll[CAT_ALT][g] = acc_alt[g]; }
  A.sq = sq;
  A.allele_fwd = c_afw; A.other_fwd = c_ofw; A.allele_cov = c_acov; A.other_cov = c_ocov; A.total_cov = c_tot;
  A.first_is_ref = first_is_ref; A.copy_number = first_cn; A.seen = seen;
  finalize_site<true>(P, O, row, A);     // rows are written completely: the output needs no clearing
}

// ---- host side ------------------------------------------------------------------------------
size_t allsites_scan_temp_bytes(int64_t n) {
  size_t a = 0;
  cub::DeviceScan::ExclusiveSum((void*)nullptr, a, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  return a + 256;
}

cudaError_t launch_allsites_plan(const DReads& R, const DSites& S, const DIndex& X, const AllSitesPlan& A,
                                 void* d_temp, size_t temp_bytes, cudaStream_t s, int* n_launches) {
  cudaError_t e;
  const int64_t words = (A.nbits + 31) / 32;
  if ((e = cudaMemsetAsync(A.cover, 0, (size_t)words * 4, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(A.site_cover, 0, (size_t)words * 4, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(A.flag + R.n, 0, 4, s)) != cudaSuccess) return e;
  const unsigned gr = (unsigned)((R.n + 255) / 256);
  if (R.n > 0) k_as_flags<<<gr, 256, 0, s>>>(R, S, X, A.flag);
  // jpos[r] = joined reads before r; jpos[n] = number of joined reads
  e = cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, A.flag, A.jpos, (int)(R.n + 1), s);
  if (e != cudaSuccess) return e;
  if (R.n > 0)
    k_as_scatter<<<gr, 256, 0, s>>>(R, S, X, A.flag, A.jpos, (JoinedRead*)A.jr, A.jcoord, A.cover, A.p0, A.nbits);
  if (S.c_hi > S.c_lo)
    k_as_site_cover<<<(S.c_hi - S.c_lo + 255) / 256, 256, 0, s>>>(S, A.site_cover, A.p0, A.nbits);
  if ((e = cudaMemsetAsync(A.win_count + A.n_win, 0, 4, s)) != cudaSuccess) return e;
  k_as_window_counts<<<(A.n_win + 255) / 256, 256, 0, s>>>(A.cover, A.n_win, A.win_count);
  e = cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, A.win_count, A.win_base, A.n_win + 1, s);
  if (e != cudaSuccess) return e;
  if (n_launches) *n_launches += 7;
  return cudaGetLastError();
}

cudaError_t launch_allsites_tiles(const DReads& R, const DSites& S, const DCnv& C, const DCallParams& P,
                                  const DResults& O, int32_t* o_start, const DIndex& X,
                                  const AllSitesPlan& A, int32_t n_joined, int32_t max_span,
                                  cudaStream_t s, int* n_launches) {
  k_as_jfine<<<(n_joined + 1 + 255) / 256, 256, 0, s>>>(A.jcoord, n_joined, A.p0, A.nbits >> 5, A.jfine);
  const unsigned grid = (unsigned)((A.nbits + AS_THREADS - 1) / AS_THREADS);
  if (P.ns <= 3)
    k_as_tiles<3><<<grid, AS_THREADS, 0, s>>>(R, S, C, P, O, o_start, X, (const JoinedRead*)A.jr, A.jcoord, A.jfine,
                                              A.cover, A.site_cover, A.win_base, A.p0, A.nbits, max_span);
  else
    k_as_tiles<NSMAX><<<grid, AS_THREADS, 0, s>>>(R, S, C, P, O, o_start, X, (const JoinedRead*)A.jr, A.jcoord, A.jfine,
                                                  A.cover, A.site_cover, A.win_base, A.p0, A.nbits, max_span);
  if (n_launches) *n_launches += 2;
  return cudaGetLastError();
}

}  // namespace avo
