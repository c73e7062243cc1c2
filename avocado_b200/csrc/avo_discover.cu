// avo_discover.cu -- DiscoverVariants on the GPU.
//
// Reference semantics (avocado-core/src/main/scala/org/bdgenomics/avocado/genotyping/
// DiscoverVariants.scala): per-read candidate emission :112-252 (walk_discover below), then
// `distinct` or groupBy(contig,start,ref,alt).count().where(count > minObservations) :71-102.
//
// SNV candidates (the bulk: every sequencing error produces one) never leave the SM: a CTA owns a
// window of DISC_W reference positions, counts candidates per position in shared memory (pass 1),
// and only for the few positions whose total exceeds the threshold builds exact (ref,alt)
// histograms (pass 2).  Indel candidates are rare; they go to a global list and are grouped
// exactly by a verified hash table (k_indel_count).  Survivors are sorted and materialised by
// assemble_sites().
#include <cub/block/block_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "avo_device.cuh"

namespace avo {

constexpr int DISC_THREADS = 256;
constexpr int DISC_W = 2048;      // reference positions owned by one tile
constexpr int DISC_HB = 8;        // hot positions histogrammed per pass-2 batch
constexpr int DISC_LIST = 3072;   // SNV candidates of one tile kept in shared memory
constexpr uint32_t NOT_HOT = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t mix32(uint32_t h, uint32_t v) {
  h ^= v;
  h *= 0x9E3779B1u;
  h ^= h >> 15;
  return h;
}

// One I / D operator of one read, queued by the thread that walked the read and handled after the
// walk by whichever thread picks it up (indels are rare; handling them inline would stall the
// other 31 reads of the warp on their scattered quality / base loads).
struct IndelEvent {
  int32_t pos;       // pos - 1 (the anchor position)
  uint32_t bidx;     // I: global index of the first inserted base; D: of the anchor base
  uint32_t rb;       // D: refb nibble index of the deleted bases
  uint32_t lenkind;  // (len << 3) | (lastRef lives in refb ? 4 : 0) | kind   kind: 1 I, 2 D
  uint32_t last;     // location of lastRef (bases4 index or refb nibble index)
};

// DiscoverVariants.scala:112-252 for one read.  snp(pos, (ref << 4) | alt) is called for every
// mismatching base that passes the phred threshold, indel(IndelEvent) for every I / D operator
// after the first aligned base (its quality test runs in disc_handle_indel).
template <typename SnpF, typename IndelF>
__device__ __forceinline__ void walk_discover(const DReads& R, const MetaA& a, const MetaB& m, int32_t thr,
                                              SnpF&& snp, IndelF&& indel) {
  const uint32_t no = m.n_ops;
  const uint32_t sb = a.seq_off;
  uint32_t rb = m.refb_off;          // byte index into refb
  int32_t pos = a.start;
  uint32_t idx = 0, last = 0;
  bool ff = true, last_refb = false;
  for (uint32_t k = 0; k < no; ++k) {
    const uint32_t op = __ldg(R.ops + a.op_off + k);
    const uint32_t len = op >> 4, code = op & 15u;
    if (ff && (code == AVO_OP_MATCH || code == AVO_OP_MISMATCH)) ff = false;
    if (ff) {                                                             // fastForward (:139-172)
      if (code == AVO_OP_SOFTCLIP || code == AVO_OP_INS) idx += len;
      else if (code == AVO_OP_DEL) { pos += (int32_t)len; rb += (len + 1) >> 1; }
    } else if (code == AVO_OP_MATCH) {                                    // :194-197
      last = sb + idx + len - 1; last_refb = false;
      pos += (int32_t)len; idx += len;
    } else if (code == AVO_OP_MISMATCH) {                                 // :198-210
      for (uint32_t i = 0; i < len; ++i) {
        // qual(i), not qual(idx + i): the first four qualities of the read travel in the record
        const int32_t q = (i < 4) ? (int32_t)((m.qhead >> (8 * i)) & 255u) : (int32_t)__ldg(R.quals + sb + i);
        if (q >= thr) snp(pos + (int32_t)i, (uint32_t)__ldg(R.refb + rb + i));
      }
      last = 2 * (rb + len - 1); last_refb = true;
      rb += len; pos += (int32_t)len; idx += len;
    } else if (code == AVO_OP_INS) {                                      // :215-228
      IndelEvent e;
      e.pos = pos - 1; e.bidx = sb + idx; e.rb = 0;
      e.lenkind = (len << 3) | (last_refb ? 4u : 0u) | 1u; e.last = last;
      indel(e);
      idx += len;
    } else if (code == AVO_OP_DEL) {                                      // :229-242
      IndelEvent e;
      e.pos = pos - 1; e.bidx = sb + idx - 1; e.rb = 2 * rb;
      e.lenkind = (len << 3) | (last_refb ? 4u : 0u) | 2u; e.last = last;
      indel(e);
      pos += (int32_t)len;
      last = 2 * rb + len - 1; last_refb = true;
      rb += (len + 1) >> 1;
    }                                                                     // Clipped: no-op (:191-193)
  }
}

constexpr int DISC_WARPS = DISC_THREADS / 32;
constexpr int DISC_Q = 256;             // indel events of one tile kept in shared memory

struct __align__(16) DiscSmem {
  uint32_t cnt[DISC_W];                 // pass 1: candidates per position; then hot index or NOT_HOT
  uint32_t hist[DISC_HB][256];          // pass 2: exact (ref<<4|alt) counts of the batch's hot positions
  uint32_t list[DISC_LIST];             // SNV candidates of the tile: (pos_off << 8) | (ref << 4) | alt
  IndelEvent q[DISC_Q];
  typename cub::BlockScan<int, DISC_THREADS>::TempStorage scan;
  int32_t tile;
  int32_t n_list;
  int32_t n_q;
  int32_t hotpos[DISC_HB];              // window offsets of the current batch's hot positions
};

__device__ __forceinline__ void disc_indel_append(const DReads& R, const DIndelList& L, int32_t pos,
                                                  int32_t ref_len, int32_t alt_len, uint32_t lastRef,
                                                  uint32_t anchor, uint32_t src, bool is_del) {
  const int32_t slot = atomicAdd(L.n, 1);
  if (slot >= L.capacity) return;   // counted; the host grows the list and reruns
  uint32_t h = mix32(0x811C9DC5u, (uint32_t)pos);
  h = mix32(h, ((uint32_t)ref_len << 16) | (uint32_t)alt_len);
  h = mix32(h, (lastRef << 4) | anchor);
  const int n = is_del ? ref_len - 1 : alt_len - 1;
  const uint8_t* buf = is_del ? R.refb : R.bases4;
  uint32_t acc = 0;
  for (int i = 0; i < n; ++i) {
    acc = (acc << 4) | nib_at(buf, src + i);
    if ((i & 7) == 7) { h = mix32(h, acc); acc = 0; }
  }
  h = mix32(h, acc);
  L.pos[slot] = pos;
  L.lens[slot] = ((uint32_t)ref_len << 16) | (uint32_t)alt_len;
  L.nibs[slot] = (lastRef << 4) | anchor;
  L.src[slot] = src;
  L.hash[slot] = h;
}

// candidate emission for one I / D operator (DiscoverVariants.scala:215-242)
__device__ __forceinline__ void disc_handle_indel(const DReads& R, const DIndelList& L, const IndelEvent& e,
                                                  int32_t thr) {
  const uint32_t kind = e.lenkind & 3u, len = e.lenkind >> 3;
  const bool last_refb = e.lenkind & 4u;
  if (kind == 1) {                                                          // :215-228
    int sum = 0;
    for (uint32_t j = 0; j <= len; ++j) sum += __ldg(R.quals + e.bidx - 1 + j);
    if (sum / (int)len < thr) return;
    const uint32_t lastRef = last_refb ? nib_at(R.refb, e.last) : nib_at(R.bases4, e.last);
    disc_indel_append(R, L, e.pos, 1, (int32_t)len + 1, lastRef, nib_at(R.bases4, e.bidx - 1), e.bidx, false);
  } else {                                                                  // :229-242
    if ((int32_t)__ldg(R.quals + e.bidx) < thr) return;
    const uint32_t lastRef = last_refb ? nib_at(R.refb, e.last) : nib_at(R.bases4, e.last);
    disc_indel_append(R, L, e.pos, (int32_t)len + 1, 1, lastRef, nib_at(R.bases4, e.bidx), e.rb, true);
  }
}

__global__ void __launch_bounds__(DISC_THREADS, 4)
k_discover_tiles(DReads R, DIndex X, int32_t thr, int32_t min_count /* keep count > min_count */,
                 int32_t n_tiles, int32_t max_span, DDiscoverOut out, DIndelList L,
                 int32_t* __restrict__ tile_counter) {
  __shared__ DiscSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  constexpr int BPT = DISC_W / IDX_G;
  constexpr int PER_T = DISC_W / DISC_THREADS;

  for (;;) {
    __syncthreads();
    if (tid == 0) { sm.tile = atomicAdd(tile_counter, 1); sm.n_list = 0; sm.n_q = 0; }
    __syncthreads();
    const int32_t tile = sm.tile;
    if (tile >= n_tiles) break;
    const int64_t wlo = (int64_t)X.win0 + (int64_t)tile * DISC_W;
    // candidates of a read lie in [start, end): reads with start < whi and start > wlo - max_span
    const int32_t rLo = __ldg(X.ridx + idx_bin_floor(X, wlo - max_span));
    const int32_t rHi = __ldg(X.ridx + min((tile + 1) * BPT, X.nb));
    if (rLo >= rHi) continue;

    for (int i = tid; i < DISC_W; i += DISC_THREADS) sm.cnt[i] = 0;
    __syncthreads();

    // ---- pass 1 (the only walk over the reads): one read per thread.  Per-position SNV candidate
    // counts and the tile's SNV candidate list in shared memory; I / D operators to the event queue
    for (int32_t r = rLo + tid; r < rHi; r += DISC_THREADS) {
      const MetaA a = load_meta_a(R.meta, r);
      if ((int64_t)a.end <= wlo) continue;             // ends before the window
      const MetaB m = load_meta_b(R.meta, r);
      walk_discover(
          R, a, m, thr,
          [&](int32_t pos, uint32_t pair) {
            const int64_t o = (int64_t)pos - wlo;
            if (o < 0 || o >= DISC_W) return;
            atomicAdd(&sm.cnt[o], 1u);
            const unsigned am = __activemask();
            const int leader = __ffs(am) - 1;
            int slot = 0;
            if (lane == leader) slot = atomicAdd(&sm.n_list, __popc(am));
            slot = __shfl_sync(am, slot, leader) + __popc(am & lt);
            if (slot < DISC_LIST) sm.list[slot] = ((uint32_t)o << 8) | (pair & 255u);
          },
          [&](const IndelEvent& e) {
            const int64_t o = (int64_t)e.pos - wlo;
            if (o < 0 || o >= DISC_W) return;
            const int slot = atomicAdd(&sm.n_q, 1);
            if (slot < DISC_Q) sm.q[slot] = e; else disc_handle_indel(R, L, e, thr);
          });
    }
    __syncthreads();
    {
      const int nq = min(sm.n_q, DISC_Q);
      for (int i = tid; i < nq; i += DISC_THREADS) disc_handle_indel(R, L, sm.q[i], thr);
    }
    const int n_list = sm.n_list;
    const bool list_ok = n_list <= DISC_LIST;   // else: fall back to re-walking the reads in pass 2

    // ---- hot positions: ordered compaction; cnt[] becomes the hot index ------------------------
    int nh = 0;
    uint32_t hot_mask = 0;
    for (int i = 0; i < PER_T; ++i) {
      const bool hot = (int32_t)sm.cnt[tid * PER_T + i] > min_count;
      hot_mask |= (uint32_t)hot << i;
      nh += hot;
    }
    int hoff, htotal;
    cub::BlockScan<int, DISC_THREADS>(sm.scan).ExclusiveSum(nh, hoff, htotal);
    for (int i = 0; i < PER_T; ++i) {
      const bool hot = (hot_mask >> i) & 1u;
      sm.cnt[tid * PER_T + i] = hot ? (uint32_t)hoff : NOT_HOT;
      hoff += hot;
    }
    __syncthreads();

    // ---- pass 2: exact (ref, alt) counts at the hot positions, DISC_HB positions at a time ------
    for (int hb = 0; hb < htotal; hb += DISC_HB) {
      for (int i = tid; i < DISC_HB * 256; i += DISC_THREADS) (&sm.hist[0][0])[i] = 0;
      __syncthreads();
      if (list_ok) {
        for (int i = tid; i < n_list; i += DISC_THREADS) {
          const uint32_t e = sm.list[i];
          const uint32_t h = sm.cnt[e >> 8];
          if (h == NOT_HOT || (int)h < hb || (int)h >= hb + DISC_HB) continue;
          atomicAdd(&sm.hist[h - hb][e & 255u], 1u);
        }
      } else {
        for (int32_t r = rLo + tid; r < rHi; r += DISC_THREADS) {
          const MetaA a = load_meta_a(R.meta, r);
          const MetaB m = load_meta_b(R.meta, r);
          walk_discover(
              R, a, m, thr,
              [&](int32_t pos, uint32_t pair) {
                const int64_t o = (int64_t)pos - wlo;
                if (o < 0 || o >= DISC_W) return;
                const uint32_t h = sm.cnt[o];
                if (h == NOT_HOT || (int)h < hb || (int)h >= hb + DISC_HB) return;
                atomicAdd(&sm.hist[h - hb][pair & 255u], 1u);
              },
              [&](const IndelEvent&) {});
        }
      }
      __syncthreads();
      // emit: a warp scans the 256 (ref, alt) bins of one hot position of this batch
      for (int i = tid; i < DISC_W; i += DISC_THREADS) {
        const uint32_t h = sm.cnt[i];
        if (h != NOT_HOT && (int)h >= hb && (int)h < hb + DISC_HB) sm.hotpos[h - hb] = i;
      }
      __syncthreads();
      for (int s = warp; s < min(DISC_HB, htotal - hb); s += DISC_WARPS) {
        const int i = sm.hotpos[s];
        const uint32_t h = (uint32_t)(hb + s);
        for (int k = lane; k < 256; k += 32) {
          const int32_t c = (int32_t)sm.hist[h - hb][k];
          if (c > min_count) {
            const int32_t slot = atomicAdd(out.n, 1);
            if (slot < out.capacity) {
              out.start[slot] = (int32_t)(wlo + i);
              out.ref_len[slot] = 1;
              out.alt_len[slot] = 1;
              out.count[slot] = c;
              out.src[slot] = (1ull << 63) | (uint64_t)k;
            }
          }
        }
      }
      __syncthreads();
    }
  }
}

cudaError_t launch_discover_tiles(const DReads& R, int32_t thr, int32_t min_obs, const DIndex& X,
                                  const BatchStats& hstats, const DDiscoverOut& out,
                                  const DIndelList& indels, int32_t* d_tile_counter, int num_sms,
                                  cudaStream_t s, int* n_launches) {
  cudaError_t e = cudaMemsetAsync(d_tile_counter, 0, sizeof(int32_t), s);
  if (e != cudaSuccess) return e;
  if (R.n == 0) return cudaSuccess;
  const int64_t span = (int64_t)hstats.max_end - (int64_t)hstats.min_start;
  const int32_t n_tiles = (int32_t)max((int64_t)1, (span + DISC_W - 1) / DISC_W);
  int blocks_per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_discover_tiles, DISC_THREADS, 0);
  if (e != cudaSuccess) return e;
  const int grid = (int)min((int64_t)n_tiles, (int64_t)num_sms * max(1, blocks_per_sm));
  const int32_t min_count = min_obs < 0 ? 0 : min_obs;  // distinct == count > 0
  k_discover_tiles<<<grid, DISC_THREADS, 0, s>>>(R, X, thr, min_count, n_tiles, hstats.max_span, out,
                                                 indels, d_tile_counter);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

// ---- indel grouping: verified hash table --------------------------------------------------------
// table word = (hash32 << 32) | (representative record index + 1); 0 = empty
__device__ __forceinline__ bool indel_equal(const DReads& R, const DIndelList& L, int32_t a, int32_t b) {
  if (L.pos[a] != L.pos[b] || L.lens[a] != L.lens[b] || L.nibs[a] != L.nibs[b]) return false;
  const uint32_t lens = L.lens[a];
  const int ref_len = (int)(lens >> 16), alt_len = (int)(lens & 0xFFFFu);
  const bool is_del = ref_len > 1;
  const int n = is_del ? ref_len - 1 : alt_len - 1;
  const uint8_t* buf = is_del ? R.refb : R.bases4;
  const uint32_t sa = L.src[a], sb = L.src[b];
  for (int i = 0; i < n; ++i)
    if (nib_at(buf, sa + i) != nib_at(buf, sb + i)) return false;
  return true;
}

__global__ void k_indel_count(DReads R, DIndelList L, int32_t n, unsigned long long* __restrict__ table,
                              int32_t* __restrict__ tcount, uint32_t mask) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t h = L.hash[i];
  uint32_t slot = (h * 0x85EBCA6Bu) & mask;
  const unsigned long long mine = ((unsigned long long)h << 32) | (unsigned long long)(i + 1);
  for (;;) {
    unsigned long long w = table[slot];
    if (w == 0ull) {
      w = atomicCAS(&table[slot], 0ull, mine);
      if (w == 0ull) { atomicAdd(&tcount[slot], 1); return; }
    }
    if ((uint32_t)(w >> 32) == h && indel_equal(R, L, i, (int32_t)(w & 0xFFFFFFFFull) - 1)) {
      atomicAdd(&tcount[slot], 1);
      return;
    }
    slot = (slot + 1) & mask;
  }
}

__global__ void k_indel_emit(DIndelList L, const unsigned long long* __restrict__ table,
                             const int32_t* __restrict__ tcount, uint32_t tsize, int32_t min_count,
                             DDiscoverOut out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tsize) return;
  const unsigned long long w = table[i];
  if (w == 0ull) return;
  const int32_t c = tcount[i];
  if (c <= min_count) return;
  const int32_t rec = (int32_t)(w & 0xFFFFFFFFull) - 1;
  const int32_t slot = atomicAdd(out.n, 1);
  if (slot >= out.capacity) return;
  const uint32_t lens = L.lens[rec];
  out.start[slot] = L.pos[rec];
  out.ref_len[slot] = (int32_t)(lens >> 16);
  out.alt_len[slot] = (int32_t)(lens & 0xFFFFu);
  out.count[slot] = c;
  out.src[slot] = (uint64_t)rec;
}

cudaError_t launch_indel_group(const DReads& R, const DIndelList& L, int32_t n_indel,
                               unsigned long long* table, int32_t* tcount, uint32_t tsize,
                               int32_t min_obs, const DDiscoverOut& out, cudaStream_t s,
                               int* n_launches) {
  if (n_indel == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(table, 0, sizeof(unsigned long long) * tsize, s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(tcount, 0, sizeof(int32_t) * tsize, s);
  if (e != cudaSuccess) return e;
  const int32_t min_count = min_obs < 0 ? 0 : min_obs;
  k_indel_count<<<(n_indel + 255) / 256, 256, 0, s>>>(R, L, n_indel, table, tcount, tsize - 1);
  k_indel_emit<<<(tsize + 255) / 256, 256, 0, s>>>(L, table, tcount, tsize, min_count, out);
  if (n_launches) *n_launches += 2;
  return cudaGetLastError();
}

// ---- assembling the sorted site table -----------------------------------------------------------
// nibble t of survivor e's allele string (ref nibbles then alt nibbles)
__device__ __forceinline__ uint32_t survivor_nib(const DReads& R, const DIndelList& L,
                                                 const DDiscoverOut& U, int32_t e, int t) {
  const uint64_t src = U.src[e];
  if (src >> 63) {                     // SNV: (ref << 4) | alt
    const uint32_t k = (uint32_t)src;
    return t == 0 ? ((k >> 4) & 15u) : (k & 15u);
  }
  const int32_t rec = (int32_t)src;
  const int ref_len = U.ref_len[e];
  const uint32_t nibs = L.nibs[rec];
  if (ref_len > 1) {                   // deletion: lastRef + deleted ref bases | anchor
    if (t == 0) return nibs >> 4;
    if (t < ref_len) return nib_at(R.refb, L.src[rec] + (uint32_t)(t - 1));
    return nibs & 15u;
  }
  if (t == 0) return nibs >> 4;        // insertion: lastRef | anchor + inserted bases
  if (t == 1) return nibs & 15u;
  return nib_at(R.bases4, L.src[rec] + (uint32_t)(t - 2));
}

__global__ void k_survivor_keys(DDiscoverOut U, int32_t n, unsigned long long* __restrict__ keys,
                                int32_t* __restrict__ vals) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // (start, ref_len, alt_len); start is biased so that negative positions sort first
  keys[i] = ((unsigned long long)((uint32_t)U.start[i] ^ 0x80000000u) << 32) |
            ((unsigned long long)(uint32_t)min(U.ref_len[i], 0xFFFF) << 16) |
            (unsigned long long)(uint32_t)min(U.alt_len[i], 0xFFFF);
  vals[i] = i;
}

// order ties of (start, ref_len, alt_len) by allele content so that the output is deterministic;
// also emits the nibble footprint (even-aligned) of every sorted site
__global__ void k_survivor_ties(DReads R, DIndelList L, DDiscoverOut U, int32_t n,
                                const unsigned long long* __restrict__ keys,
                                int32_t* __restrict__ vals, uint32_t* __restrict__ foot) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  {
    const int32_t e = vals[i];
    foot[i] = (uint32_t)((U.ref_len[e] + U.alt_len[e] + 1) & ~1);
  }
  if (i > 0 && keys[i - 1] == keys[i]) return;   // not a run head
  int32_t j = i + 1;
  while (j < n && keys[j] == keys[i]) ++j;
  if (j - i < 2) return;
  // insertion sort of vals[i..j) by content (runs are tiny: alleles sharing position and lengths)
  for (int32_t a = i + 1; a < j; ++a) {
    const int32_t va = vals[a];
    int32_t b = a - 1;
    while (b >= i) {
      const int32_t vb = vals[b];
      const int len = U.ref_len[va] + U.alt_len[va];
      int cmp = 0;
      for (int t = 0; t < len && cmp == 0; ++t) {
        const uint32_t x = survivor_nib(R, L, U, vb, t), y = survivor_nib(R, L, U, va, t);
        cmp = (x > y) - (x < y);
      }
      if (cmp <= 0) break;
      vals[b + 1] = vb;
      --b;
    }
    vals[b + 1] = va;
  }
}

__global__ void k_survivor_fill(DReads R, DIndelList L, DDiscoverOut U, int32_t n,
                                const int32_t* __restrict__ vals, const uint32_t* __restrict__ aoff,
                                int32_t* __restrict__ o_start, int32_t* __restrict__ o_ref_len,
                                int32_t* __restrict__ o_alt_len, uint32_t* __restrict__ o_aoff,
                                uint8_t* __restrict__ o_alleles, int32_t* __restrict__ o_count,
                                uint32_t total_nibbles) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) o_aoff[n] = total_nibbles;
  if (i >= n) return;
  const int32_t e = vals[i];
  const int ref_len = U.ref_len[e], alt_len = U.alt_len[e];
  o_start[i] = U.start[e];
  o_ref_len[i] = ref_len;
  o_alt_len[i] = alt_len;
  o_count[i] = U.count[e];
  const uint32_t off = aoff[i];
  o_aoff[i] = off;
  const int len = ref_len + alt_len;
  for (int t = 0; t < len; t += 2) {
    const uint32_t hi = survivor_nib(R, L, U, e, t);
    const uint32_t lo = (t + 1 < len) ? survivor_nib(R, L, U, e, t + 1) : 0u;
    o_alleles[(off + (uint32_t)t) >> 1] = (uint8_t)((hi << 4) | lo);
  }
}

size_t assemble_temp_bytes(int32_t n) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs((void*)nullptr, a, (const unsigned long long*)nullptr,
                                  (unsigned long long*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, n);
  cub::DeviceScan::ExclusiveSum((void*)nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr, n);
  return (a > b ? a : b) + 256;
}

// Sorts the n survivors and writes the site table.  keys_a/keys_b, vals_a/vals_b, foot, aoff are
// scratch arrays of n (aoff: n) entries.  *d_total receives the nibble total (device scalar).
cudaError_t launch_assemble_sites(const DReads& R, const DIndelList& L, const DDiscoverOut& U,
                                  int32_t n, unsigned long long* keys_a, unsigned long long* keys_b,
                                  int32_t* vals_a, int32_t* vals_b, uint32_t* foot, uint32_t* aoff,
                                  void* d_temp, size_t temp_bytes, cudaStream_t s, int* n_launches) {
  if (n == 0) return cudaSuccess;
  const int g = (n + 255) / 256;
  k_survivor_keys<<<g, 256, 0, s>>>(U, n, keys_a, vals_a);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, keys_a, keys_b, vals_a, vals_b,
                                                  n, 0, 64, s);
  if (e != cudaSuccess) return e;
  k_survivor_ties<<<g, 256, 0, s>>>(R, L, U, n, keys_b, vals_b, foot);
  e = cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, foot, aoff, n, s);
  if (e != cudaSuccess) return e;
  if (n_launches) *n_launches += 4;
  return cudaGetLastError();
}

cudaError_t launch_fill_sites(const DReads& R, const DIndelList& L, const DDiscoverOut& U, int32_t n,
                              const int32_t* vals, const uint32_t* aoff, int32_t* o_start,
                              int32_t* o_ref_len, int32_t* o_alt_len, uint32_t* o_aoff,
                              uint8_t* o_alleles, int32_t* o_count, uint32_t total_nibbles,
                              cudaStream_t s, int* n_launches) {
  k_survivor_fill<<<(max(n, 1) + 255) / 256, 256, 0, s>>>(R, L, U, n, vals, aoff, o_start, o_ref_len,
                                                          o_alt_len, o_aoff, o_alleles, o_count,
                                                          total_nibbles);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

}  // namespace avo
