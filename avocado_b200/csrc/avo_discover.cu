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

// Tile shape (measured, profiles/README.md r2): 1536 resident threads per SM in any split run at
// the same occupancy; the smallest CTAs lose least to the barriers between a tile's phases.
#ifndef AVO_DISC_THREADS
#define AVO_DISC_THREADS 256
#endif
constexpr int DISC_THREADS = AVO_DISC_THREADS;
#ifndef AVO_DISC_W
#define AVO_DISC_W 2048
#endif
#ifndef AVO_DISC_MINB
#define AVO_DISC_MINB 6
#endif
constexpr int DISC_W = AVO_DISC_W;  // reference positions owned by one tile
#ifndef AVO_DISC_HB
#define AVO_DISC_HB 2
#endif
constexpr int DISC_HB = AVO_DISC_HB;        // hot positions histogrammed per pass-2 batch
constexpr uint32_t NOT_HOT = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t mix32(uint32_t h, uint32_t v) {
  h ^= v;
  h *= 0x9E3779B1u;
  h ^= h >> 15;
  return h;
}

// DiscoverVariants.scala:112-252, one operator at a time.  The hot loop carries only what SNV
// candidates need (reference position, refb cursor, the fastForward flag) and is written without
// branches on the operator kind (lanes of a warp hold different kinds); only candidate emission
// diverges: snp(pos, (ref << 4) | alt) for every mismatching base that passes the phred threshold,
// and indel(k, anchor) for every I / D operator after the first aligned base.  Indel candidates are
// rare; their read index, lastRef and quality window are recomputed by disc_handle_indel.
struct WalkState {
  int32_t pos;       // reference position of the next aligned base
  uint32_t rb;       // byte index into refb
  bool ff;           // still in fastForward (:139-172): no aligned operator seen yet
};

__device__ __forceinline__ WalkState walk_begin(const MetaA& a, const MetaB& m) {
  WalkState st;
  st.pos = a.start; st.rb = m.refb_off; st.ff = true;
  return st;
}

// SNV candidates of one mismatch run (:198-210)
template <typename SnpF>
__device__ __forceinline__ void walk_mismatch(const DReads& R, uint32_t len, int32_t pos, uint32_t rb,
                                              uint32_t sb, uint32_t qhead, int32_t thr, SnpF&& snp) {
  // qual(i), not qual(idx + i): the first four qualities of the read travel in the record
  if ((int32_t)(qhead & 255u) >= thr) snp(pos, (uint32_t)__ldg(R.refb + rb));
  const uint32_t in_head = R.qhead_short ? 2u : 4u;      // the 6-byte wire form carries two of them
  for (uint32_t i = 1; i < len; ++i) {
    const int32_t q = (i < in_head) ? (int32_t)((qhead >> (8 * i)) & 255u) : (int32_t)pl_qual(R.payload, sb, i);
    if (q >= thr) snp(pos + (int32_t)i, (uint32_t)__ldg(R.refb + rb + i));
  }
}

template <typename SnpF, typename IndelF>
__device__ __forceinline__ void walk_step(const DReads& R, uint32_t op, uint32_t k, uint32_t sb,
                                          uint32_t qhead, int32_t thr, WalkState& st, SnpF&& snp,
                                          IndelF&& indel) {
  const uint32_t len = op >> 4, code = op & 15u;
  const bool isX = code == AVO_OP_MISMATCH, isD = code == AVO_OP_DEL;
  const bool aligned = code <= AVO_OP_MISMATCH;                           // MATCH or MISMATCH
  if (isX) walk_mismatch(R, len, st.pos, st.rb, sb, qhead, thr, snp);
  if ((code == AVO_OP_INS || isD) && !st.ff) indel(k, st.pos - 1);        // :215-242
  st.pos += (aligned || isD) ? (int32_t)len : 0;
  st.rb += isX ? len : isD ? ((len + 1) >> 1) : 0u;
  st.ff = st.ff && !aligned;
}

// the whole read (long reads and the overflow paths)
template <typename SnpF, typename IndelF>
__device__ __forceinline__ void walk_discover(const DReads& R, const MetaA& a, const MetaB& m, int32_t thr,
                                              SnpF&& snp, IndelF&& indel) {
  WalkState st = walk_begin(a, m);
  for (uint32_t k = 0; k < m.n_ops; ++k)
    walk_step(R, __ldg(R.ops + a.op_off + k), k, a.seq_off, m.qhead, thr, st, snp, indel);
}

// An I / D operator awaiting its quality test: (read, operator index)
struct IndelEvent { int32_t r; uint32_t k; };

constexpr int DISC_WARPS = DISC_THREADS / 32;
constexpr int DISC_Q = 256;             // indel events of one tile kept in shared memory
constexpr int DISC_SHORT = 3;           // reads with at most this many operators are walked at once
constexpr int DISC_LONGQ = 1024;        // reads with more operators: walked afterwards, lanes full
#ifdef AVO_DISC_CQ
// Experiment (off by default, NOT yet run on a GPU; profiles/r6_discover_regions.txt): SNV candidates are
// queued per warp while the reads are walked and counted afterwards by all 32 lanes, instead of being
// counted on the spot by the ~10 lanes that hold a mismatch at that step of their walk.
constexpr int DISC_CQ = 96;             // queued candidates per warp
constexpr int DISC_CQ_DRAIN = 48;       // counted once this many are waiting
#endif

struct __align__(16) DiscSmem {
  // fast path: SNV candidate counts per position and alternate base, two 16-bit fields per word
  // (cnt[0]: A | C << 16, cnt[1]: G | T << 16), and the OR of the candidates' reference nibbles,
  // 4 bits per position.  slow path: cnt[0] = candidates per position, then hot index or NOT_HOT.
  uint32_t cnt[2][DISC_W];
  uint32_t refm[DISC_W / 8];
  uint32_t hist[DISC_HB][256];          // slow path: exact (ref<<4|alt) counts of the hot positions
  IndelEvent q[DISC_Q];
  int32_t longq[DISC_LONGQ];
  typename cub::BlockScan<int, DISC_THREADS>::TempStorage scan;
  int32_t tile;
  int32_t n_q;
  int32_t n_long;
  int32_t exotic;                       // a candidate with a non-ACGT base was seen: use the slow path
  int32_t hotpos[DISC_HB];              // window offsets of the current batch's hot positions
#ifdef AVO_DISC_CQ
  uint32_t cq[DISC_WARPS][DISC_CQ];     // (window offset << 8) | (ref << 4 | alt)
  int32_t cq_n[DISC_WARPS];             // candidates appended since the warp's last drain (may exceed DISC_CQ)
#endif
};

__device__ __forceinline__ void disc_indel_append(const DReads& R, const DIndelList& L, int32_t pos,
                                                  int32_t ref_len, int32_t alt_len, uint32_t lastRef,
                                                  uint32_t anchor, uint32_t src, uint32_t blk, bool is_del) {
  const int32_t slot = atomicAdd(L.n, 1);
  if (slot >= L.capacity) return;   // counted; the host grows the list and reruns
  uint32_t h = mix32(0x811C9DC5u, (uint32_t)pos);
  h = mix32(h, ((uint32_t)ref_len << 16) | (uint32_t)alt_len);
  h = mix32(h, (lastRef << 4) | anchor);
  const int n = is_del ? ref_len - 1 : alt_len - 1;
  uint32_t acc = 0;
  for (int i = 0; i < n; ++i) {
    acc = (acc << 4) | (is_del ? nib_at(R.refb, src + i) : pl_base(R.payload, blk, src + i));
    if ((i & 7) == 7) { h = mix32(h, acc); acc = 0; }
  }
  h = mix32(h, acc);
  L.pos[slot] = pos;
  L.lens[slot] = ((uint32_t)ref_len << 16) | (uint32_t)alt_len;
  L.nibs[slot] = (lastRef << 4) | anchor;
  L.src[slot] = src;
  L.blk[slot] = blk;
  L.hash[slot] = h;
}

// candidate emission for operator k (an I or D) of read r: DiscoverVariants.scala:177-242 replayed
// up to that operator to recover the read index, the refb cursor and lastRef (:186)
__device__ void disc_handle_indel(const DReads& R, const DIndelList& L, int32_t r, uint32_t k, int32_t thr) {
  const MetaA a = load_meta_a(R.meta, r);
  const MetaB m = load_meta_b(R.meta, r);
  const uint32_t sb = a.seq_off;
  int32_t pos = a.start;
  uint32_t idx = 0, rb = m.refb_off, last = 0;
  bool ff = true, last_refb = false;
  for (uint32_t j = 0; j < k; ++j) {
    const uint32_t op = __ldg(R.ops + a.op_off + j);
    const uint32_t len = op >> 4, code = op & 15u;
    if (ff && code <= AVO_OP_MISMATCH) ff = false;
    if (ff) {                                                             // fastForward (:139-172)
      if (code == AVO_OP_SOFTCLIP || code == AVO_OP_INS) idx += len;
      else if (code == AVO_OP_DEL) { pos += (int32_t)len; rb += (len + 1) >> 1; }
    } else if (code == AVO_OP_MATCH) {                                    // :194-197
      last = idx + len - 1; last_refb = false;     // index within the read
      pos += (int32_t)len; idx += len;
    } else if (code == AVO_OP_MISMATCH) {                                 // :198-210
      last = 2 * (rb + len - 1); last_refb = true;
      rb += len; pos += (int32_t)len; idx += len;
    } else if (code == AVO_OP_INS) {
      idx += len;
    } else if (code == AVO_OP_DEL) {
      pos += (int32_t)len;
      last = 2 * rb + len - 1; last_refb = true;
      rb += (len + 1) >> 1;
    }                                                                     // Clipped: no-op (:191-193)
  }
  const uint32_t op = __ldg(R.ops + a.op_off + k);
  const uint32_t len = op >> 4;
  const uint32_t lastRef = last_refb ? nib_at(R.refb, last) : pl_base(R.payload, sb, last);
  const uint32_t anchor = pl_base(R.payload, sb, idx - 1);
  if ((op & 15u) == AVO_OP_INS) {                                         // :215-228
    int sum = 0;
    for (uint32_t j = 0; j <= len; ++j) sum += (int)pl_qual(R.payload, sb, idx - 1 + j);
    if (sum / (int)len < thr) return;
    disc_indel_append(R, L, pos - 1, 1, (int32_t)len + 1, lastRef, anchor, idx, sb, false);
  } else {                                                                // :229-242
    if ((int32_t)pl_qual(R.payload, sb, idx - 1) < thr) return;
    disc_indel_append(R, L, pos - 1, (int32_t)len + 1, 1, lastRef, anchor, 2 * rb, 0u, true);
  }
}

__device__ __forceinline__ void disc_emit_snv(const DDiscoverOut& out, int32_t pos, uint32_t pair, int32_t c) {
  const int32_t slot = atomicAdd(out.n, 1);
  if (slot < out.capacity) {
    out.start[slot] = pos;
    out.ref_len[slot] = 1;
    out.alt_len[slot] = 1;
    out.count[slot] = c;
    out.src[slot] = (1ull << 63) | (uint64_t)pair;
  }
}

__global__ void __launch_bounds__(DISC_THREADS, AVO_DISC_MINB)
k_discover_tiles(DReads R, DIndex X, int32_t thr, int32_t min_count /* keep count > min_count */,
                 int32_t n_tiles, int32_t max_span, int32_t max_end, int32_t verify, DDiscoverOut out,
                 DIndelList L, int32_t* __restrict__ tile_counter, int32_t* __restrict__ flags) {
  extern __shared__ __align__(16) unsigned char disc_smem_raw[];     // sizeof(DiscSmem), may exceed 48 KB
  DiscSmem& sm = *reinterpret_cast<DiscSmem*>(disc_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int BPT = DISC_W / IDX_G;
  constexpr int PER_T = DISC_W / DISC_THREADS;

  for (int i = tid; i < 2 * DISC_W; i += DISC_THREADS) (&sm.cnt[0][0])[i] = 0;
  for (int i = tid; i < DISC_W / 8; i += DISC_THREADS) sm.refm[i] = 0;
#ifdef AVO_DISC_CQ
  if (lane == 0) sm.cq_n[warp] = 0;
#endif
  for (;;) {
    __syncthreads();
    if (tid == 0) { sm.tile = atomicAdd(tile_counter, 1); sm.n_q = 0; sm.n_long = 0; sm.exotic = 0; }
    __syncthreads();
    const int32_t tile = sm.tile;
    if (tile >= n_tiles) break;
    const int64_t wlo = (int64_t)X.win0 + (int64_t)tile * DISC_W;
    // candidates of a read lie in [start, end): reads with start < whi and start > wlo - max_span
    const int32_t rLo = __ldg(X.ridx + idx_bin_floor(X, wlo - max_span));
    const int32_t rHi = __ldg(X.ridx + min((tile + 1) * BPT, X.nb));
    // the records whose start lies in this window: every record is owned by exactly one tile as
    // long as the index is monotone (checked here), so the checks below see every record once
    const int32_t rOwn = __ldg(X.ridx + min(tile * BPT, X.nb));
    if (verify && tid == 0) {
      if (rOwn > rHi || (tile == 0 && rOwn != 0) || (tile == 0 && R.n > 0 && meta_start(R.meta, 0) < X.win0))
        atomicOr(flags, DISC_FLAG_UNSORTED);
    }
    if (rLo >= rHi) continue;

    // ---- the walk over the reads: one read per thread.  SNV candidates are counted per position
    // and alternate base in shared memory; I / D operators go to the event queue.  Most reads are
    // one match, or one mismatch run after the leading match, and are handled on the spot; the
    // others are queued and walked afterwards so that no lane idles while a neighbour works
    // through a long operator list.
#ifdef AVO_DISC_CQ
    auto count_candidate = [&](uint32_t o, uint32_t pair) {
      const uint32_t refn = (pair >> 4) & 15u, altn = pair & 15u;
      if (!(((0x116u >> refn) & (0x116u >> altn)) & 1u)) { sm.exotic = 1; return; }
      const int ai = (int)((altn >> 1) - (altn >> 3));
      atomicAdd(&sm.cnt[ai >> 1][o], 1u << ((ai & 1) * 16));
      atomicOr(&sm.refm[o >> 3], refn << ((o & 7) * 4));
    };
    // wlo fits in 32 bits (win0 <= wlo < max_end); a candidate of a read in [rLo, rHi) lies within
    // max_span + DISC_W of it, so the unsigned difference cannot wrap into the window
    const uint32_t wlo32 = (uint32_t)(int32_t)wlo;
    auto snp = [&](int32_t pos, uint32_t pair) {
      const uint32_t o = (uint32_t)pos - wlo32;
      if (o >= (uint32_t)DISC_W) return;
      // the lanes that arrive here together take consecutive slots: one shared atomic per group
      const unsigned grp = __activemask();
      const int leader = __ffs(grp) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(&sm.cq_n[warp], __popc(grp));
      base = __shfl_sync(grp, base, leader);
      const int slot = base + __popc(grp & ((1u << lane) - 1u));
      if (slot < DISC_CQ) sm.cq[warp][slot] = (o << 8) | (pair & 255u);
      else count_candidate(o, pair);                       // queue full: counted on the spot
    };
    auto drain = [&]() {                                   // every lane of the warp must call it
      __syncwarp();
      const int n = min(sm.cq_n[warp], DISC_CQ);
      for (int i = lane; i < n; i += 32) {
        const uint32_t e = sm.cq[warp][i];
        count_candidate(e >> 8, e & 255u);
      }
      __syncwarp();
      if (lane == 0) sm.cq_n[warp] = 0;
      __syncwarp();
    };
    for (int32_t r0 = rLo; r0 < rHi; r0 += DISC_THREADS) {   // the same trip count for every lane of a warp
      const int32_t r = r0 + tid;
      if (r < rHi) do {
      const MetaA a = load_meta_a(R.meta, r);
#else
    auto snp = [&](int32_t pos, uint32_t pair) {
      const int64_t o = (int64_t)pos - wlo;
      if (o < 0 || o >= DISC_W) return;
      const uint32_t refn = (pair >> 4) & 15u, altn = pair & 15u;
      // A C G T are the nibbles 1 2 4 8 (bits of 0x116); anything else takes the exact path
      if (!(((0x116u >> refn) & (0x116u >> altn)) & 1u)) { sm.exotic = 1; return; }
      const int ai = (int)((altn >> 1) - (altn >> 3));                // 1 2 4 8 -> 0 1 2 3
      atomicAdd(&sm.cnt[ai >> 1][o], 1u << ((ai & 1) * 16));
      atomicOr(&sm.refm[o >> 3], refn << ((o & 7) * 4));
    };
    for (int32_t r = rLo + tid; r < rHi; r += DISC_THREADS) {
      const MetaA a = load_meta_a(R.meta, r);
#endif
      const MetaB m = load_meta_b(R.meta, r);
      if (r >= rOwn) {
        if (R.coords) R.coords[r] = make_int2(a.start, a.end);   // dense (start, end) copy for the join
        if (verify) {     // sort order and the caller's bounds (avo_batch_stats)
          if (r > 0 && meta_start(R.meta, r - 1) > a.start) atomicOr(flags, DISC_FLAG_UNSORTED);
          if (a.end > max_end || (int64_t)a.end - a.start > max_span) atomicOr(flags, DISC_FLAG_BAD_STATS);
        }
      }
      if ((int64_t)a.end <= wlo || m.n_ops == 0) continue;   // ends before the window / no operators
      const uint32_t w0 = __ldg(R.ops + a.op_off);
      const uint32_t w1 = (m.n_ops > 1) ? __ldg(R.ops + a.op_off + 1) : 0u;
      // the common shapes "=", "= X", "= X =" and "= X S"
      const bool simple = m.n_ops <= DISC_SHORT && (w0 & 15u) == AVO_OP_MATCH &&
                          (m.n_ops == 1 || (w1 & 15u) == AVO_OP_MISMATCH);
      if (simple) {
        if (m.n_ops > 1 && (m.n_ops == 2 || ((__ldg(R.ops + a.op_off + 2) & 15u) | 4u) == 4u))
          walk_mismatch(R, w1 >> 4, a.start + (int32_t)(w0 >> 4), m.refb_off, a.seq_off, m.qhead, thr, snp);
        else if (m.n_ops > 1) goto general;
        continue;
      }
    general:
      {
#ifdef AVO_DISC_AGG
        // experiment (off by default, not yet run on a GPU): one shared atomic per group of lanes, not per lane
        const unsigned grp = __activemask();
        const int leader = __ffs(grp) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&sm.n_long, __popc(grp));
        base = __shfl_sync(grp, base, leader);
        const int slot = base + __popc(grp & ((1u << lane) - 1u));
#else
        const int slot = atomicAdd(&sm.n_long, 1);
#endif
        if (slot < DISC_LONGQ) { sm.longq[slot] = r; continue; }
      }
      walk_discover(R, a, m, thr, snp, [&](uint32_t k, int32_t anchor) {
        if ((int64_t)anchor >= wlo && (int64_t)anchor < wlo + DISC_W) disc_handle_indel(R, L, r, k, thr);
      });
#ifdef AVO_DISC_CQ
      } while (0);                       // `continue` above leaves the read, not the warp's iteration
      __syncwarp();
      if (sm.cq_n[warp] >= DISC_CQ_DRAIN) drain();
#endif
    }
#ifdef AVO_DISC_CQ
    drain();
#endif
    __syncthreads();
    {
      const int nl = min(sm.n_long, DISC_LONGQ);
      for (int i = tid; i < nl; i += DISC_THREADS) {
        const int32_t r = sm.longq[i];
        walk_discover(R, load_meta_a(R.meta, r), load_meta_b(R.meta, r), thr, snp,
                      [&](uint32_t k, int32_t anchor) {
                        if ((int64_t)anchor < wlo || (int64_t)anchor >= wlo + DISC_W) return;
                        const int slot = atomicAdd(&sm.n_q, 1);
                        if (slot < DISC_Q) { sm.q[slot].r = r; sm.q[slot].k = k; }
                        else disc_handle_indel(R, L, r, k, thr);
                      });
      }
#ifdef AVO_DISC_CQ
      drain();                           // before the barrier: the emission below reads the counters
#endif
    }
    __syncthreads();
    {
      const int nq = min(sm.n_q, DISC_Q);
      for (int i = tid; i < nq; i += DISC_THREADS) disc_handle_indel(R, L, sm.q[i].r, sm.q[i].k, thr);
    }

    // ---- emission.  Fast path: every candidate had ACGT bases and each position saw one reference
    // base, so the per-alt counters ARE the (ref, alt) counts.  The counters are cleared as they
    // are read.  (16-bit fields cannot overflow: fewer than 65536 reads reach the tile.)
    // A position is hot when one of its 16-bit counters exceeds min_count (tested four positions
    // at a time; fields stay below 32768: fewer than 32768 reads reach the tile).
    bool slow = sm.exotic != 0 || (rHi - rLo) >= 32768 || min_count >= 32767;
    const uint32_t bias = (uint32_t)(0x7FFF - min_count) * 0x00010001u;
    uint32_t hot = 0;                      // bit (4 j + e): position 4 (tid + j * DISC_THREADS) + e is hot
    if (!slow) {
      bool bad = false;
#pragma unroll
      for (int j = 0; j < PER_T / 4; ++j) {
        const int o4 = tid + j * DISC_THREADS;        // 16 contiguous bytes per lane: no bank conflicts
        const uint4 vAC = reinterpret_cast<const uint4*>(sm.cnt[0])[o4];
        const uint4 vGT = reinterpret_cast<const uint4*>(sm.cnt[1])[o4];
        const uint32_t wa[4] = {vAC.x, vAC.y, vAC.z, vAC.w};
        const uint32_t wg[4] = {vGT.x, vGT.y, vGT.z, vGT.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (((wa[e] + bias) | (wg[e] + bias)) & 0x80008000u) hot |= 1u << (4 * j + e);
      }
      for (uint32_t h = hot; h; h &= h - 1) {
        const int b = __ffs(h) - 1;
        const int o = 4 * (tid + (b >> 2) * DISC_THREADS) + (b & 3);
        bad = bad || __popc((sm.refm[o >> 3] >> ((o & 7) * 4)) & 15u) != 1;
      }
      slow = __syncthreads_or(bad);
    }
    if (!slow) {
      for (uint32_t h = hot; h; h &= h - 1) {
        const int b = __ffs(h) - 1;
        const int o = 4 * (tid + (b >> 2) * DISC_THREADS) + (b & 3);
        const uint32_t wAC = sm.cnt[0][o], wGT = sm.cnt[1][o];
        const uint32_t refn = (sm.refm[o >> 3] >> ((o & 7) * 4)) & 15u;
        const int32_t c[4] = {(int32_t)(wAC & 0xFFFFu), (int32_t)(wAC >> 16), (int32_t)(wGT & 0xFFFFu),
                              (int32_t)(wGT >> 16)};
        for (int ai = 0; ai < 4; ++ai)
          if (c[ai] > min_count) disc_emit_snv(out, (int32_t)(wlo + o), (refn << 4) | (1u << ai), c[ai]);
      }
      __syncthreads();
      const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int j = 0; j < PER_T / 4; ++j) {
        reinterpret_cast<uint4*>(sm.cnt[0])[tid + j * DISC_THREADS] = z;
        reinterpret_cast<uint4*>(sm.cnt[1])[tid + j * DISC_THREADS] = z;
      }
      for (int i = tid; i < DISC_W / 8; i += DISC_THREADS) sm.refm[i] = 0;
      continue;
    }

    // ---- slow path (non-ACGT candidate bases, inconsistent reference bases, or extreme depth):
    // exact.  Re-walk for per-position candidate totals, then (ref, alt) histograms of the positions
    // whose total exceeds the threshold, DISC_HB positions at a time.
    __syncthreads();
    for (int i = tid; i < 2 * DISC_W; i += DISC_THREADS) (&sm.cnt[0][0])[i] = 0;
    for (int i = tid; i < DISC_W / 8; i += DISC_THREADS) sm.refm[i] = 0;
    __syncthreads();
    for (int32_t r = rLo + tid; r < rHi; r += DISC_THREADS) {
      walk_discover(
          R, load_meta_a(R.meta, r), load_meta_b(R.meta, r), thr,
          [&](int32_t pos, uint32_t) {
            const int64_t o = (int64_t)pos - wlo;
            if (o >= 0 && o < DISC_W) atomicAdd(&sm.cnt[0][o], 1u);
          },
          [&](uint32_t, int32_t) {});
    }
    __syncthreads();
    int nh = 0;
    uint32_t hot_mask = 0;
    for (int i = 0; i < PER_T; ++i) {
      const bool hot = (int32_t)sm.cnt[0][tid * PER_T + i] > min_count;
      hot_mask |= (uint32_t)hot << i;
      nh += hot;
    }
    int hoff, htotal;
    cub::BlockScan<int, DISC_THREADS>(sm.scan).ExclusiveSum(nh, hoff, htotal);
    for (int i = 0; i < PER_T; ++i) {
      const bool hot = (hot_mask >> i) & 1u;
      sm.cnt[0][tid * PER_T + i] = hot ? (uint32_t)hoff : NOT_HOT;
      hoff += hot;
    }
    __syncthreads();
    for (int hb = 0; hb < htotal; hb += DISC_HB) {
      for (int i = tid; i < DISC_HB * 256; i += DISC_THREADS) (&sm.hist[0][0])[i] = 0;
      __syncthreads();
      for (int32_t r = rLo + tid; r < rHi; r += DISC_THREADS) {
        walk_discover(
            R, load_meta_a(R.meta, r), load_meta_b(R.meta, r), thr,
            [&](int32_t pos, uint32_t pair) {
              const int64_t o = (int64_t)pos - wlo;
              if (o < 0 || o >= DISC_W) return;
              const uint32_t h = sm.cnt[0][o];
              if (h == NOT_HOT || (int)h < hb || (int)h >= hb + DISC_HB) return;
              atomicAdd(&sm.hist[h - hb][pair & 255u], 1u);
            },
            [&](uint32_t, int32_t) {});
      }
      __syncthreads();
      for (int i = tid; i < DISC_W; i += DISC_THREADS) {
        const uint32_t h = sm.cnt[0][i];
        if (h != NOT_HOT && (int)h >= hb && (int)h < hb + DISC_HB) sm.hotpos[h - hb] = i;
      }
      __syncthreads();
      // a warp scans the 256 (ref, alt) bins of one hot position of this batch
      for (int s = warp; s < min(DISC_HB, htotal - hb); s += DISC_WARPS) {
        const int i = sm.hotpos[s];
        for (int k = lane; k < 256; k += 32) {
          const int32_t c = (int32_t)sm.hist[s][k];
          if (c > min_count) disc_emit_snv(out, (int32_t)(wlo + i), (uint32_t)k, c);
        }
      }
      __syncthreads();
    }
    for (int i = tid; i < DISC_W; i += DISC_THREADS) sm.cnt[0][i] = 0;
  }
}

cudaError_t launch_discover_tiles(const DReads& R, int32_t thr, int32_t min_obs, const DIndex& X,
                                  const BatchStats& hstats, int verify, const DDiscoverOut& out,
                                  const DIndelList& indels, int32_t* d_tile_counter, int32_t* d_flags,
                                  int num_sms, cudaStream_t s, int* n_launches) {
  cudaError_t e = cudaMemsetAsync(d_tile_counter, 0, sizeof(int32_t), s);
  if (e != cudaSuccess) return e;
  if (R.n == 0) return cudaSuccess;
  const int64_t span = (int64_t)hstats.max_end - (int64_t)hstats.min_start;
  const int32_t n_tiles = (int32_t)max((int64_t)1, (span + DISC_W - 1) / DISC_W);
  e = cudaFuncSetAttribute(k_discover_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DiscSmem));
  if (e != cudaSuccess) return e;
  int blocks_per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_discover_tiles, DISC_THREADS, sizeof(DiscSmem));
  if (e != cudaSuccess) return e;
  const int grid = (int)min((int64_t)n_tiles, (int64_t)num_sms * max(1, blocks_per_sm));
  const int32_t min_count = min_obs < 0 ? 0 : min_obs;  // distinct == count > 0
  k_discover_tiles<<<grid, DISC_THREADS, sizeof(DiscSmem), s>>>(R, X, thr, min_count, n_tiles, hstats.max_span,
                                                 hstats.max_end, verify, out, indels, d_tile_counter,
                                                 d_flags);
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
  const uint32_t sa = L.src[a], sb = L.src[b];
  if (is_del) {
    for (int i = 0; i < n; ++i)
      if (nib_at(R.refb, sa + i) != nib_at(R.refb, sb + i)) return false;
  } else {
    const uint32_t ba = L.blk[a], bb = L.blk[b];
    for (int i = 0; i < n; ++i)
      if (pl_base(R.payload, ba, sa + i) != pl_base(R.payload, bb, sb + i)) return false;
  }
  return true;
}

__global__ void k_indel_count(DReads R, DIndelList L, unsigned long long* __restrict__ table,
                              int32_t* __restrict__ tcount, uint32_t mask, int32_t* __restrict__ flags) {
  const int32_t n = min(*L.n, L.capacity);
  if ((uint32_t)n > (mask + 1) / 2) {          // the probe loops need free slots
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(flags, DISC_FLAG_TABLE_FULL);
    return;
  }
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t h = L.hash[i];
    uint32_t slot = (h * 0x85EBCA6Bu) & mask;
    const unsigned long long mine = ((unsigned long long)h << 32) | (unsigned long long)(i + 1);
    for (;;) {
      unsigned long long w = table[slot];
      if (w == 0ull) {
        w = atomicCAS(&table[slot], 0ull, mine);
        if (w == 0ull) { atomicAdd(&tcount[slot], 1); break; }
      }
      if ((uint32_t)(w >> 32) == h && indel_equal(R, L, i, (int32_t)(w & 0xFFFFFFFFull) - 1)) {
        atomicAdd(&tcount[slot], 1);
        break;
      }
      slot = (slot + 1) & mask;
    }
  }
}

__global__ void k_indel_emit(DIndelList L, const unsigned long long* __restrict__ table,
                             const int32_t* __restrict__ tcount, uint32_t tsize, int32_t min_count,
                             DDiscoverOut out, int32_t* __restrict__ nib, int32_t* __restrict__ nsurv) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tsize) return;
  const unsigned long long w = table[i];
  if (w == 0ull) return;
  const int32_t c = tcount[i];
  if (c <= min_count) return;
  const int32_t rec = (int32_t)(w & 0xFFFFFFFFull) - 1;
  const uint32_t lens = L.lens[rec];
  atomicAdd(nib, (int32_t)(((lens >> 16) + (lens & 0xFFFFu) + 1u) & ~1u));
  atomicAdd(nsurv, 1);
  const int32_t slot = atomicAdd(out.n, 1);
  if (slot >= out.capacity) return;
  out.start[slot] = L.pos[rec];
  out.ref_len[slot] = (int32_t)(lens >> 16);
  out.alt_len[slot] = (int32_t)(lens & 0xFFFFu);
  out.count[slot] = c;
  out.src[slot] = (uint64_t)rec;
}

cudaError_t launch_indel_group(const DReads& R, const DIndelList& L, unsigned long long* table,
                               int32_t* tcount, uint32_t tsize, int32_t min_obs,
                               const DDiscoverOut& out, int32_t* d_nib, int32_t* d_nsurv,
                               int32_t* d_flags, int num_sms, cudaStream_t s, int* n_launches) {
  cudaError_t e = cudaMemsetAsync(table, 0, sizeof(unsigned long long) * tsize, s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(tcount, 0, sizeof(int32_t) * tsize, s);
  if (e != cudaSuccess) return e;
  const int32_t min_count = min_obs < 0 ? 0 : min_obs;
  k_indel_count<<<num_sms * 4, 256, 0, s>>>(R, L, table, tcount, tsize - 1, d_flags);
  k_indel_emit<<<(tsize + 255) / 256, 256, 0, s>>>(L, table, tcount, tsize, min_count, out, d_nib, d_nsurv);
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
  return pl_base(R.payload, L.blk[rec], L.src[rec] + (uint32_t)(t - 2));
}

__global__ void k_survivor_keys(DDiscoverOut U, int32_t n, int32_t min_start,
                                unsigned long long* __restrict__ keys, int32_t* __restrict__ vals) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = (unsigned long long)(uint32_t)(U.start[i] - min_start);   // candidates lie at or after min_start
  vals[i] = i;
}

// (ref_len, alt_len, allele content) order of two survivors that share a start
__device__ int survivor_cmp(const DReads& R, const DIndelList& L, const DDiscoverOut& U, int32_t x, int32_t y) {
  const int rx = U.ref_len[x], ry = U.ref_len[y];
  if (rx != ry) return rx < ry ? -1 : 1;
  const int ax = U.alt_len[x], ay = U.alt_len[y];
  if (ax != ay) return ax < ay ? -1 : 1;
  for (int t = 0; t < rx + ax; ++t) {
    const uint32_t p = survivor_nib(R, L, U, x, t), q = survivor_nib(R, L, U, y, t);
    if (p != q) return p < q ? -1 : 1;
  }
  return 0;
}

// The radix sort orders by start only; the (tiny) runs of survivors that share a start are put in
// (ref_len, alt_len, allele content) order here, so that the table is sorted the way Forest sorts
// its keys and is deterministic.  Also emits the nibble footprint (even-aligned) of every sorted site.
__global__ void k_survivor_ties(DReads R, DIndelList L, DDiscoverOut U, int32_t n,
                                const unsigned long long* __restrict__ keys,
                                int32_t* __restrict__ vals, uint32_t* __restrict__ foot) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i > 0 && keys[i - 1] == keys[i]) return;   // not a run head
  int32_t j = i + 1;
  while (j < n && keys[j] == keys[i]) ++j;
  for (int32_t a = i + 1; a < j; ++a) {          // insertion sort of vals[i..j)
    const int32_t va = vals[a];
    int32_t b = a - 1;
    while (b >= i) {
      const int32_t vb = vals[b];
      if (survivor_cmp(R, L, U, vb, va) <= 0) break;
      vals[b + 1] = vb;
      --b;
    }
    vals[b + 1] = va;
  }
  for (int32_t a = i; a < j; ++a) {
    const int32_t e = vals[a];
    foot[a] = (uint32_t)((U.ref_len[e] + U.alt_len[e] + 1) & ~1);
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
                                  int32_t n, int32_t min_start, int32_t max_end,
                                  unsigned long long* keys_a, unsigned long long* keys_b,
                                  int32_t* vals_a, int32_t* vals_b, uint32_t* foot, uint32_t* aoff,
                                  void* d_temp, size_t temp_bytes, cudaStream_t s, int* n_launches) {
  if (n == 0) return cudaSuccess;
  const int g = (n + 255) / 256;
  k_survivor_keys<<<g, 256, 0, s>>>(U, n, min_start, keys_a, vals_a);
  int bits = 1;
  while (bits < 32 && (((int64_t)max_end - (int64_t)min_start) >> bits) != 0) ++bits;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, keys_a, keys_b, vals_a, vals_b,
                                                  n, 0, bits, s);
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
