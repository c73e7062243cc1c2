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
};

__device__ __forceinline__ void disc_indel_append(const DReads& R, const DIndelList& L, int32_t pos,
                                                  int32_t ref_len, int32_t alt_len, uint32_t lastRef,
                                                  uint32_t anchor, uint32_t src, uint32_t blk, bool is_del) {
  const int32_t slot = atomicAdd(L.n, 1);
  if (slot >= L.capacity) return;   // counted; the host grows the list and reruns
  uint32_t h = mix32(0x811C9DC5u, (uint32_t)pos);
  h = mix32(h, (uint32_t)ref_len);
  h = mix32(h, (uint32_t)alt_len);
  h = mix32(h, (lastRef << 4) | anchor);
  const int n = is_del ? ref_len - 1 : alt_len - 1;
  uint32_t acc = 0;
  for (int i = 0; i < n; ++i) {
    acc = (acc << 4) | (is_del ? nib_at(R.refb, src + i) : pl_base(R.payload, blk, src + i));
    if ((i & 7) == 7) { h = mix32(h, acc); acc = 0; }
  }
  h = mix32(h, acc);
  L.pos[slot] = pos;
  L.ref_len[slot] = ref_len;
  L.alt_len[slot] = alt_len;
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

// A survivor (count > minObservations): appended to the list and linked into the tile it starts in;
// k_assemble_sites turns the per-tile lists into the sorted table without a sort or a host round trip.
__device__ __forceinline__ void disc_emit(const DDiscoverOut& out, int32_t pos, int32_t ref_len, int32_t alt_len,
                                          int32_t c, uint64_t src) {
  const int32_t slot = atomicAdd(out.n, 1);
  if (slot >= out.capacity) return;      // counted: the host grows the list and reruns
  out.start[slot] = pos;
  out.ref_len[slot] = ref_len;
  out.alt_len[slot] = alt_len;
  out.count[slot] = c;
  out.src[slot] = src;
  TileBucket* tb = out.tiles + (pos - out.win0) / DISC_W;
  out.next[slot] = atomicExch(&tb->head, slot + 1);
  atomicAdd(&tb->cnt, 1);
  atomicAdd(&tb->nib, (uint32_t)(ref_len + alt_len + 1) & ~1u);
  atomicMax(&tb->max_rel_end, pos + ref_len - out.win0);
}
__device__ __forceinline__ void disc_emit_snv(const DDiscoverOut& out, int32_t pos, uint32_t pair, int32_t c) {
  disc_emit(out, pos, 1, 1, c, (1ull << 63) | (uint64_t)pair);
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
    // wlo fits 32 bits (win0 <= wlo < max_end) and a candidate lies within max_span + DISC_W of it, so
    // one unsigned compare is the window test (a position below wlo wraps far past DISC_W)
    const int32_t wlo32 = (int32_t)wlo;
    auto snp = [&](int32_t pos, uint32_t pair) {
      const uint32_t o = (uint32_t)pos - (uint32_t)wlo32;
      if (o >= (uint32_t)DISC_W) return;
      const uint32_t refn = (pair >> 4) & 15u, altn = pair & 15u;
      // A C G T are the nibbles 1 2 4 8 (bits of 0x116); anything else takes the exact path
      if (!(((0x116u >> refn) & (0x116u >> altn)) & 1u)) { sm.exotic = 1; return; }
      const int ai = (int)((altn >> 1) - (altn >> 3));                // 1 2 4 8 -> 0 1 2 3
      atomicAdd(&sm.cnt[ai >> 1][o], 1u << ((ai & 1) * 16));
      atomicOr(&sm.refm[o >> 3], refn << ((o & 7) * 4));
    };
    for (int32_t r = rLo + tid; r < rHi; r += DISC_THREADS) {
      const MetaA a = load_meta_a(R.meta, r);
      const MetaB m = load_meta_b(R.meta, r);
      // the previous read's start: the neighbouring lane has it (the lanes still in the loop are 0..k)
      int32_t prev_start = __shfl_up_sync(__activemask(), a.start, 1);
      if (lane == 0) prev_start = r > 0 ? meta_start(R.meta, r - 1) : a.start;
      if (r >= rOwn) {
        if (R.coords) R.coords[r] = make_int2(a.start, a.end);   // dense (start, end) copy for the join
        if (verify) {     // sort order and the caller's bounds (avo_batch_stats)
          if (prev_start > a.start) atomicOr(flags, DISC_FLAG_UNSORTED);
          if (a.end > max_end || a.end - a.start > max_span || a.end < a.start) atomicOr(flags, DISC_FLAG_BAD_STATS);
        }
      }
      if (a.end <= wlo32 || m.n_ops == 0) continue;   // ends before the window / no operators
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
        const int slot = atomicAdd(&sm.n_long, 1);
        if (slot < DISC_LONGQ) { sm.longq[slot] = r; continue; }
      }
      walk_discover(R, a, m, thr, snp, [&](uint32_t k, int32_t anchor) {
        if ((int64_t)anchor >= wlo && (int64_t)anchor < wlo + DISC_W) disc_handle_indel(R, L, r, k, thr);
      });
    }
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

int32_t discover_tiles(const BatchStats& hstats);

cudaError_t launch_discover_tiles(const DReads& R, int32_t thr, int32_t min_obs, const DIndex& X,
                                  const BatchStats& hstats, int verify, const DDiscoverOut& out,
                                  const DIndelList& indels, int32_t* d_tile_counter, int32_t* d_flags,
                                  int num_sms, int* blocks_per_sm_cache /* per context; 0 = not yet known */,
                                  cudaStream_t s, int* n_launches) {
  cudaError_t e = cudaMemsetAsync(d_tile_counter, 0, sizeof(int32_t), s);
  if (e != cudaSuccess) return e;
  if (R.n == 0) return cudaSuccess;
  const int32_t n_tiles = discover_tiles(hstats);
  if (*blocks_per_sm_cache == 0) {       // once per context (the attribute belongs to the device's context)
    e = cudaFuncSetAttribute(k_discover_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DiscSmem));
    if (e != cudaSuccess) return e;
    int b = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_discover_tiles, DISC_THREADS, sizeof(DiscSmem));
    if (e != cudaSuccess) return e;
    *blocks_per_sm_cache = max(1, b);
  }
  const int blocks_per_sm = *blocks_per_sm_cache;
  const int grid = (int)min((int64_t)n_tiles, (int64_t)num_sms * max(1, blocks_per_sm));
  const int32_t min_count = min_obs < 0 ? 0 : min_obs;  // distinct == count > 0
  k_discover_tiles<<<grid, DISC_THREADS, sizeof(DiscSmem), s>>>(R, X, thr, min_count, n_tiles, hstats.max_span,
                                                 hstats.max_end, verify, out, indels, d_tile_counter,
                                                 d_flags);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

// ---- candidate events: derived once per batch (avo_derive_events), counted per call ---------------------
// The per-read part of DiscoverVariants.variantsInRead with nothing filtered: every mismatching base with the
// quality the reference tests (qual(i), i counted inside the run: walk_mismatch), every I / D operator after
// the first aligned one.
template <typename SnvF, typename IndelF>
__device__ __forceinline__ void walk_events(const DReads& R, const MetaA& a, const MetaB& m, SnvF&& snv, IndelF&& indel) {
  int32_t pos = a.start;
  uint32_t rb = m.refb_off;
  bool ff = true;
  for (uint32_t k = 0; k < m.n_ops; ++k) {
    const uint32_t op = __ldg(R.ops + a.op_off + k);
    const uint32_t len = op >> 4, code = op & 15u;
    const bool isX = code == AVO_OP_MISMATCH, isD = code == AVO_OP_DEL, aligned = code <= AVO_OP_MISMATCH;
    if (isX)
      for (uint32_t i = 0; i < len; ++i) {
        const uint32_t q = (i < 4u) ? ((m.qhead >> (8 * i)) & 255u) : pl_qual(R.payload, a.seq_off, i);
        snv(pos + (int32_t)i, (uint32_t)__ldg(R.refb + rb + i), q);
      }
    if ((code == AVO_OP_INS || isD) && !ff) indel(k);                     // :215-242
    pos += (aligned || isD) ? (int32_t)len : 0;
    rb += isX ? len : isD ? ((len + 1) >> 1) : 0u;
    ff = ff && !aligned;
  }
}

// one thread per read, a warp = 32 consecutive reads = one entry of the per-32-reads counts
__global__ void __launch_bounds__(256)
k_events_count(DReads R, uint32_t* __restrict__ c_snv, uint32_t* __restrict__ c_ind, BatchStats* __restrict__ st) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t ns = 0, ni = 0;
  int32_t mn = INT32_MAX, mxe = INT32_MIN, mxs = 0, uns = 0;
  if (r < R.n) {
    const MetaA a = load_meta_a(R.meta, r);
    const MetaB m = load_meta_b(R.meta, r);
    mn = a.start; mxe = a.end; mxs = a.end - a.start;
    if (r > 0 && meta_start(R.meta, r - 1) > a.start) uns = 1;
    walk_events(R, a, m, [&](int32_t, uint32_t, uint32_t) { ++ns; }, [&](uint32_t) { ++ni; });
  }
  for (int o = 16; o > 0; o >>= 1) {
    ns += __shfl_xor_sync(0xFFFFFFFFu, ns, o);
    ni += __shfl_xor_sync(0xFFFFFFFFu, ni, o);
    mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o));
    mxe = max(mxe, __shfl_xor_sync(0xFFFFFFFFu, mxe, o));
    mxs = max(mxs, __shfl_xor_sync(0xFFFFFFFFu, mxs, o));
    uns |= __shfl_xor_sync(0xFFFFFFFFu, uns, o);
  }
  if ((threadIdx.x & 31) == 0 && (r & ~(int64_t)31) < R.n) {
    c_snv[r >> 5] = ns;
    c_ind[r >> 5] = ni;
    atomicMin(&st->min_start, mn);
    atomicMax(&st->max_end, mxe);
    atomicMax(&st->max_span, mxs);
    if (uns) atomicOr(&st->unsorted, 1);
  }
}

__global__ void __launch_bounds__(256)
k_events_write(DReads R, const uint32_t* __restrict__ snv_first, const uint32_t* __restrict__ ind_first, int64_t snv_cap,
               int64_t ind_cap, int32_t* __restrict__ snv_pos, uint16_t* __restrict__ snv_pq,
               int32_t* __restrict__ indel_read, uint32_t* __restrict__ indel_op) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  if ((r & ~(int64_t)31) >= R.n) return;                       // (whole warps only)
  MetaA a{};
  MetaB m{};
  uint32_t ns = 0, ni = 0;
  if (r < R.n) {
    a = load_meta_a(R.meta, r);
    m = load_meta_b(R.meta, r);
    walk_events(R, a, m, [&](int32_t, uint32_t, uint32_t) { ++ns; }, [&](uint32_t) { ++ni; });
  }
  uint32_t ps = ns, pi = ni;                                   // inclusive scans over the warp
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t ys = __shfl_up_sync(0xFFFFFFFFu, ps, d), yi = __shfl_up_sync(0xFFFFFFFFu, pi, d);
    if (lane >= d) { ps += ys; pi += yi; }
  }
  if (r >= R.n) return;
  int64_t es = (int64_t)__ldg(snv_first + (r >> 5)) + (ps - ns), ei = (int64_t)__ldg(ind_first + (r >> 5)) + (pi - ni);
  walk_events(R, a, m,
              [&](int32_t pos, uint32_t pair, uint32_t q) {
                if (es < snv_cap) { snv_pos[es] = pos; snv_pq[es] = (uint16_t)((pair << 8) | q); }
                ++es;
              },
              [&](uint32_t k) {
                if (ei < ind_cap) { indel_read[ei] = (int32_t)r; indel_op[ei] = k; }
                ++ei;
              });
}

cudaError_t launch_events_count(const DReads& R, uint32_t* c_snv, uint32_t* c_ind, BatchStats* d_stats, cudaStream_t s) {
  if (R.n == 0) return cudaSuccess;
  k_events_count<<<(unsigned)((R.n + 255) / 256), 256, 0, s>>>(R, c_snv, c_ind, d_stats);
  return cudaGetLastError();
}

size_t events_scan_temp_bytes(int64_t n) {
  size_t a = 0;
  cub::DeviceScan::ExclusiveSum((void*)nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
  return a + 256;
}
cudaError_t launch_events_scan(const uint32_t* counts, uint32_t* first, int64_t n, void* d_temp, size_t temp_bytes, cudaStream_t s) {
  return cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, counts, first, (int)n, s);
}

cudaError_t launch_events_write(const DReads& R, const uint32_t* snv_first, const uint32_t* ind_first, int64_t snv_cap,
                                int64_t ind_cap, int32_t* snv_pos, uint16_t* snv_pq, int32_t* indel_read, uint32_t* indel_op,
                                cudaStream_t s) {
  if (R.n == 0) return cudaSuccess;
  k_events_write<<<(unsigned)((R.n + 255) / 256), 256, 0, s>>>(R, snv_first, ind_first, snv_cap, ind_cap, snv_pos, snv_pq,
                                                             indel_read, indel_op);
  return cudaGetLastError();
}

// Discovery over a batch that carries its events.  Tiles of EV_W positions (larger than k_discover_tiles': a
// tile's work is only its events, a few hundred per 2048 positions, so the per-tile costs -- barriers, the
// bounds look-ups, clearing the counters -- are what has to be amortised); the candidates of a tile are the
// events of the reads that can reach it -- [snv_first[rLo / 32], snv_first[ceil(rHi / 32)]): a superset whose
// surplus (events of up to 31 reads on either side) lies outside the window and is dropped by the window test
// -- streamed with coalesced 4 + 2 byte loads.  The same 16-bit counters per (position, alternate base) as
// k_discover_tiles; the event that lifts a counter over minObservations notes (position, base) in a list, so
// emission looks at the survivors instead of scanning the window (the scan remains for lists that overflow:
// minObservations unset makes every candidate a survivor).
#ifndef AVO_EV_W
#define AVO_EV_W 2048
#endif
#ifndef AVO_EV_THREADS
#define AVO_EV_THREADS 128
#endif
constexpr int EV_W = AVO_EV_W;
constexpr int EV_THREADS = AVO_EV_THREADS;
constexpr int EV_WARPS = EV_THREADS / 32;
constexpr int EV_HOT = 1024;
constexpr int EV_UNROLL = 4;          // events requested per thread before the first is counted
static_assert(EV_W % DISC_W == 0 && EV_W % (4 * EV_THREADS) == 0, "event tiles are whole discovery tiles");
struct __align__(16) EvSmem {
  uint32_t cnt[2][EV_W];
  uint32_t refm[EV_W / 8];
  uint32_t hist[DISC_HB][256];
  typename cub::BlockScan<int, EV_THREADS>::TempStorage scan;
  int64_t bounds[4];                    // (first, last) event of the current / next tile
  int32_t hot[EV_HOT];                  // (window offset << 2) | alternate base index
  int32_t n_hot;
  int32_t exotic;
  int32_t hotpos[DISC_HB];
};

__global__ void __launch_bounds__(EV_THREADS)
k_discover_events(DReads R, DIndex X, int32_t thr, int32_t min_count, int32_t n_tiles, int32_t max_span, DDiscoverOut out) {
  extern __shared__ __align__(16) unsigned char disc_smem_raw[];
  EvSmem& sm = *reinterpret_cast<EvSmem*>(disc_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int BPT = EV_W / IDX_G;
  constexpr int PER_T = EV_W / EV_THREADS;
  const int64_t n_w = (R.n + 31) >> 5;

  auto clear_counters = [&]() {
    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < PER_T / 4; ++j) {
      reinterpret_cast<uint4*>(sm.cnt[0])[tid + j * EV_THREADS] = z;
      reinterpret_cast<uint4*>(sm.cnt[1])[tid + j * EV_THREADS] = z;
    }
    for (int i = tid; i < EV_W / 8; i += EV_THREADS) sm.refm[i] = 0;
  };
  // lanes 0 and 1 fetch the NEXT tile's bounds (read index -> event index: two dependent look-ups) while the
  // current tile is counted; tiles are taken in strides of the grid (their work is uniform)
  auto fetch_bounds = [&](int32_t t, int slot) {
    if (tid >= 2) return;
    int64_t v = 0;
    if (t < n_tiles) {
      const int64_t w = (int64_t)X.win0 + (int64_t)t * EV_W;
      if (tid == 0) {
        const int32_t rLo = __ldg(X.ridx + idx_bin_floor(X, w - max_span));
        v = __ldg(R.snv_first + (rLo >> 5));
      } else {
        const int32_t rHi = __ldg(X.ridx + (int32_t)min((int64_t)(t + 1) * BPT, (int64_t)X.nb));
        v = min((int64_t)__ldg(R.snv_first + min(((int64_t)rHi + 31) >> 5, n_w)), R.n_snv);
      }
    }
    sm.bounds[slot * 2 + tid] = v;
  };
  clear_counters();
  fetch_bounds(blockIdx.x, 0);
  if (tid == 0) { sm.exotic = 0; sm.n_hot = 0; }
  int it = 0;
  for (int32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    __syncthreads();                      // counters cleared, flags reset, this tile's bounds in place
    const int64_t eLo = sm.bounds[(it & 1) * 2], eHi = sm.bounds[(it & 1) * 2 + 1];
    fetch_bounds(tile + (int32_t)gridDim.x, (it + 1) & 1);
    if (eLo >= eHi) continue;
    const int64_t wlo = (int64_t)X.win0 + (int64_t)tile * EV_W;
    const int32_t wlo32 = (int32_t)wlo;
    const bool counted = (eHi - eLo) < 32768 && min_count < 32767;      // 16-bit fields cannot overflow
    if (counted) {
      for (int64_t e0 = eLo + tid; e0 < eHi; e0 += EV_UNROLL * EV_THREADS) {
       uint32_t pqs[EV_UNROLL], pos_s[EV_UNROLL];
#pragma unroll
       for (int u = 0; u < EV_UNROLL; ++u) {
         const int64_t e = e0 + (int64_t)u * EV_THREADS;
         const bool in = e < eHi;
         pqs[u] = in ? (uint32_t)__ldg(R.snv_pq + e) : 0u;
         pos_s[u] = in ? (uint32_t)__ldg(R.snv_pos + e) : (uint32_t)wlo32 - 1u;      // (outside every window)
       }
#pragma unroll
       for (int u = 0; u < EV_UNROLL; ++u) {
        const uint32_t pq = pqs[u];
        const uint32_t o = pos_s[u] - (uint32_t)wlo32;
        if (o >= (uint32_t)EV_W) continue;
        if ((int32_t)(pq & 255u) < thr) continue;
        const uint32_t refn = (pq >> 12) & 15u, altn = (pq >> 8) & 15u;
        // A C G T are the nibbles 1 2 4 8 (bits of 0x116); anything else takes the exact path
        if (!(((0x116u >> refn) & (0x116u >> altn)) & 1u)) { sm.exotic = 1; continue; }
        const int ai = (int)((altn >> 1) - (altn >> 3));                // 1 2 4 8 -> 0 1 2 3
        const int sh = (ai & 1) * 16;
        const uint32_t old = atomicAdd(&sm.cnt[ai >> 1][o], 1u << sh);
        atomicOr(&sm.refm[o >> 3], refn << ((o & 7) * 4));
        if ((int32_t)((old >> sh) & 0xFFFFu) == min_count) {             // this event makes it a survivor
          const int k = atomicAdd(&sm.n_hot, 1);
          if (k < EV_HOT) sm.hot[k] = (int32_t)((o << 2) | (uint32_t)ai);
        }
       }
      }
    }
    __syncthreads();
    bool slow = !counted || sm.exotic != 0;
    const int n_hot = sm.n_hot;
    if (!slow && n_hot <= EV_HOT) {
      // ---- emission from the survivor list (a position whose reads disagree about its reference base: exact path)
      bool bad = false;
      for (int k = tid; k < n_hot; k += EV_THREADS) {
        const uint32_t o = (uint32_t)sm.hot[k] >> 2;
        bad = bad || __popc((sm.refm[o >> 3] >> ((o & 7) * 4)) & 15u) != 1;
      }
      slow = __syncthreads_or(bad);
      if (!slow) {
        for (int k = tid; k < n_hot; k += EV_THREADS) {
          const uint32_t o = (uint32_t)sm.hot[k] >> 2;
          const int ai = sm.hot[k] & 3;
          const uint32_t refn = (sm.refm[o >> 3] >> ((o & 7) * 4)) & 15u;
          const int32_t c = (int32_t)((sm.cnt[ai >> 1][o] >> ((ai & 1) * 16)) & 0xFFFFu);
          disc_emit_snv(out, (int32_t)(wlo + o), (refn << 4) | (1u << ai), c);
        }
        __syncthreads();
        clear_counters();
        if (tid == 0) { sm.exotic = 0; sm.n_hot = 0; }
        continue;
      }
    } else if (!slow) {
      // ---- more survivors than the list holds: scan the window (four positions per word operation)
      const uint32_t bias = (uint32_t)(0x7FFF - min_count) * 0x00010001u;
      bool bad = false;
      for (int j = 0; j < PER_T / 4; ++j) {
        const int o4 = tid + j * EV_THREADS;
        const uint4 vAC = reinterpret_cast<const uint4*>(sm.cnt[0])[o4];
        const uint4 vGT = reinterpret_cast<const uint4*>(sm.cnt[1])[o4];
        const uint32_t wa[4] = {vAC.x, vAC.y, vAC.z, vAC.w};
        const uint32_t wg[4] = {vGT.x, vGT.y, vGT.z, vGT.w};
        for (int e = 0; e < 4; ++e)
          if (((wa[e] + bias) | (wg[e] + bias)) & 0x80008000u) {
            const int o = 4 * o4 + e;
            bad = bad || __popc((sm.refm[o >> 3] >> ((o & 7) * 4)) & 15u) != 1;
          }
      }
      slow = __syncthreads_or(bad);
      if (!slow) {
        for (int j = 0; j < PER_T / 4; ++j) {
          const int o4 = tid + j * EV_THREADS;
          for (int e = 0; e < 4; ++e) {
            const int o = 4 * o4 + e;
            const uint32_t wAC = sm.cnt[0][o], wGT = sm.cnt[1][o];
            if (!(((wAC + bias) | (wGT + bias)) & 0x80008000u)) continue;
            const uint32_t refn = (sm.refm[o >> 3] >> ((o & 7) * 4)) & 15u;
            const int32_t c[4] = {(int32_t)(wAC & 0xFFFFu), (int32_t)(wAC >> 16), (int32_t)(wGT & 0xFFFFu),
                                  (int32_t)(wGT >> 16)};
            for (int ai = 0; ai < 4; ++ai)
              if (c[ai] > min_count) disc_emit_snv(out, (int32_t)(wlo + o), (refn << 4) | (1u << ai), c[ai]);
          }
        }
        __syncthreads();
        clear_counters();
        if (tid == 0) { sm.exotic = 0; sm.n_hot = 0; }
        continue;
      }
    }

    // ---- slow path (non-ACGT candidate bases, inconsistent reference bases, or extreme depth): exact.
    // Per-position candidate totals, then (ref, alt) histograms of the positions whose total exceeds the
    // threshold, DISC_HB positions at a time -- over the same event range.
    auto each = [&](auto&& f) {
      for (int64_t e = eLo + tid; e < eHi; e += EV_THREADS) {
        const uint32_t pq = __ldg(R.snv_pq + e);
        if ((int32_t)(pq & 255u) < thr) continue;
        const uint32_t o = (uint32_t)__ldg(R.snv_pos + e) - (uint32_t)wlo32;
        if (o < (uint32_t)EV_W) f(o, pq >> 8);
      }
    };
    __syncthreads();
    clear_counters();
    __syncthreads();
    each([&](uint32_t o, uint32_t) { atomicAdd(&sm.cnt[0][o], 1u); });
    __syncthreads();
    int nh = 0;
    uint32_t hot_mask = 0;
    static_assert(PER_T <= 32, "one bit per owned position");
    for (int i = 0; i < PER_T; ++i) {
      const bool is_hot = (int32_t)sm.cnt[0][tid * PER_T + i] > min_count;
      hot_mask |= (uint32_t)is_hot << i;
      nh += is_hot;
    }
    int hoff, htotal;
    cub::BlockScan<int, EV_THREADS>(sm.scan).ExclusiveSum(nh, hoff, htotal);
    for (int i = 0; i < PER_T; ++i) {
      const bool is_hot = (hot_mask >> i) & 1u;
      sm.cnt[0][tid * PER_T + i] = is_hot ? (uint32_t)hoff : NOT_HOT;
      hoff += is_hot;
    }
    __syncthreads();
    for (int hb = 0; hb < htotal; hb += DISC_HB) {
      for (int i = tid; i < DISC_HB * 256; i += EV_THREADS) (&sm.hist[0][0])[i] = 0;
      __syncthreads();
      each([&](uint32_t o, uint32_t pair) {
        const uint32_t h = sm.cnt[0][o];
        if (h == NOT_HOT || (int)h < hb || (int)h >= hb + DISC_HB) return;
        atomicAdd(&sm.hist[h - hb][pair & 255u], 1u);
      });
      __syncthreads();
      for (int i = tid; i < EV_W; i += EV_THREADS) {
        const uint32_t h = sm.cnt[0][i];
        if (h != NOT_HOT && (int)h >= hb && (int)h < hb + DISC_HB) sm.hotpos[h - hb] = i;
      }
      __syncthreads();
      for (int s = warp; s < min(DISC_HB, htotal - hb); s += EV_WARPS) {
        const int i = sm.hotpos[s];
        for (int k = lane; k < 256; k += 32) {
          const int32_t c = (int32_t)sm.hist[s][k];
          if (c > min_count) disc_emit_snv(out, (int32_t)(wlo + i), (uint32_t)k, c);
        }
      }
      __syncthreads();
    }
    __syncthreads();
    clear_counters();
    if (tid == 0) { sm.exotic = 0; sm.n_hot = 0; }
  }
}

// every I / D event once: quality test and candidate (disc_handle_indel replays the read up to the operator)
__global__ void __launch_bounds__(256)
k_indel_events(DReads R, DIndelList L, int32_t thr) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= R.n_indel) return;
  const int32_t r = __ldg(R.indel_read + e);
  if (r < 0 || r >= R.n) return;
  disc_handle_indel(R, L, r, __ldg(R.indel_op + e), thr);
}

cudaError_t launch_discover_events(const DReads& R, int32_t thr, int32_t min_obs, const DIndex& X, const BatchStats& hstats,
                                   const DDiscoverOut& out, const DIndelList& indels, int32_t* d_tile_counter, int32_t* d_flags,
                                   int num_sms, int* blocks_per_sm_cache, cudaStream_t s, int* n_launches) {
  (void)d_flags;
  cudaError_t e = cudaMemsetAsync(d_tile_counter, 0, sizeof(int32_t), s);
  if (e != cudaSuccess) return e;
  if (R.n == 0) return cudaSuccess;
  const int32_t n_tiles = (discover_tiles(hstats) + EV_W / DISC_W - 1) / (EV_W / DISC_W);
  if (*blocks_per_sm_cache == 0) {
    e = cudaFuncSetAttribute(k_discover_events, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EvSmem));
    if (e != cudaSuccess) return e;
    int b = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_discover_events, EV_THREADS, sizeof(EvSmem));
    if (e != cudaSuccess) return e;
    *blocks_per_sm_cache = max(1, b);
  }
  const int grid = (int)min((int64_t)n_tiles, (int64_t)num_sms * max(1, *blocks_per_sm_cache));
  const int32_t min_count = min_obs < 0 ? 0 : min_obs;
  k_discover_events<<<grid, EV_THREADS, sizeof(EvSmem), s>>>(R, X, thr, min_count, n_tiles, hstats.max_span, out);
  if (R.n_indel > 0) k_indel_events<<<(unsigned)((R.n_indel + 255) / 256), 256, 0, s>>>(R, indels, thr);
  if (n_launches) *n_launches += R.n_indel > 0 ? 2 : 1;
  return cudaGetLastError();
}

// ---- indel grouping: verified hash table --------------------------------------------------------
// table word = (hash32 << 32) | (representative record index + 1); 0 = empty
__device__ __forceinline__ bool indel_equal(const DReads& R, const DIndelList& L, int32_t a, int32_t b) {
  if (L.pos[a] != L.pos[b] || L.ref_len[a] != L.ref_len[b] || L.alt_len[a] != L.alt_len[b] || L.nibs[a] != L.nibs[b])
    return false;
  const int ref_len = L.ref_len[a], alt_len = L.alt_len[a];
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
                             DDiscoverOut out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tsize) return;
  const unsigned long long w = table[i];
  if (w == 0ull) return;
  const int32_t c = tcount[i];
  if (c <= min_count) return;
  const int32_t rec = (int32_t)(w & 0xFFFFFFFFull) - 1;
  disc_emit(out, L.pos[rec], L.ref_len[rec], L.alt_len[rec], c, (uint64_t)rec);
}

cudaError_t launch_indel_group(const DReads& R, const DIndelList& L, unsigned long long* table,
                               int32_t* tcount, uint32_t tsize, int32_t min_obs,
                               const DDiscoverOut& out, int32_t* d_flags, int num_sms, cudaStream_t s,
                               int* n_launches) {
  // (table and tcount are zeroed by the caller together with the other per-call scratch)
  const int32_t min_count = min_obs < 0 ? 0 : min_obs;
  k_indel_count<<<num_sms * 4, 256, 0, s>>>(R, L, table, tcount, tsize - 1, d_flags);
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
  return pl_base(R.payload, L.blk[rec], L.src[rec] + (uint32_t)(t - 2));
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

// full order of the table: (start, ref_len, alt_len, allele content) -- the way Forest sorts its keys
// (contig, start, end), made deterministic for sites that share a region
__device__ __forceinline__ int survivor_order(const DReads& R, const DIndelList& L, const DDiscoverOut& U, int32_t x,
                                              int32_t y) {
  const int32_t sx = U.start[x], sy = U.start[y];
  if (sx != sy) return sx < sy ? -1 : 1;
  return survivor_cmp(R, L, U, x, y);
}

int32_t discover_tiles(const BatchStats& hstats) {
  const int64_t span = (int64_t)hstats.max_end - (int64_t)hstats.min_start;
  return (int32_t)max((int64_t)1, (span + DISC_W - 1) / DISC_W);
}

// One thread per tile.  A CTA first adds up the buckets of all tiles before its own (the buckets are
// a few hundred kilobytes in L2: cheaper than a scan kernel and its launches), then scans its own 256,
// then every thread puts the survivors of its tile (two or three; an insertion sort) into the table.
constexpr int ASM_THREADS = 256;
constexpr int ASM_K = 32;
__global__ void __launch_bounds__(ASM_THREADS)
k_assemble_sites(DReads R, DIndelList L, DDiscoverOut U, int32_t n_tiles, int32_t nb, DSiteTableOut T) {
  __shared__ unsigned long long s_before[ASM_THREADS / 32], s_sum[ASM_THREADS / 32];
  __shared__ int32_t s_bmax[ASM_THREADS / 32], s_max[ASM_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int32_t t0 = blockIdx.x * ASM_THREADS;
  const uint4* __restrict__ B = reinterpret_cast<const uint4*>(U.tiles);
  static_assert(sizeof(TileBucket) == 16, "one 16-byte load per bucket");
  // ---- rows (high word), nibbles (low word) and the furthest site end of the tiles before this CTA's
  unsigned long long acc = 0;
  int32_t mx = 0;
  for (int32_t i = tid; i < t0; i += ASM_THREADS) {
    const uint4 b = __ldg(B + i);
    acc += ((unsigned long long)b.y << 32) | (unsigned long long)b.z;
    mx = max(mx, (int32_t)b.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
  }
  // ---- this thread's tile, and the scan over the CTA's tiles
  const int32_t t = t0 + tid;
  uint4 mine = make_uint4(0, 0, 0, 0);
  if (t < n_tiles) mine = __ldg(B + t);
  const unsigned long long v = ((unsigned long long)mine.y << 32) | (unsigned long long)mine.z;
  unsigned long long x = v;              // inclusive over the warp
  int32_t pm = (int32_t)mine.w;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, d);
    const int32_t ym = __shfl_up_sync(0xFFFFFFFFu, pm, d);
    if (lane >= d) { x += y; pm = max(pm, ym); }
  }
  const int32_t pm_prev = __shfl_up_sync(0xFFFFFFFFu, pm, 1);
  if (lane == 31) { s_sum[warp] = x; s_max[warp] = pm; }
  if (lane == 0) { s_before[warp] = acc; s_bmax[warp] = mx; }
  __syncthreads();
  unsigned long long base = 0;
  int32_t bmax = 0;
  for (int w = 0; w < ASM_THREADS / 32; ++w) {
    base += s_before[w];
    bmax = max(bmax, s_bmax[w]);
    if (w < warp) { base += s_sum[w]; bmax = max(bmax, s_max[w]); }
  }
  const unsigned long long off = base + (x - v);                 // exclusive: rows << 32 | nibbles before this tile
  const int32_t pme_in = max(bmax, lane == 0 ? 0 : pm_prev);     // furthest end (relative to win0) before this tile
  if (t >= n_tiles) return;
  const int32_t row0 = (int32_t)(off >> 32);
  const uint32_t nib0 = (uint32_t)off;
  const int32_t cnt = (int32_t)mine.y;
  // ---- the tile's survivors in table order
  int32_t ids[ASM_K];
  const bool small = cnt <= ASM_K;
  if (small) {
    int n = 0;
    for (int32_t e = (int32_t)mine.x; e != 0 && n < cnt; e = U.next[e - 1]) {
      const int32_t id = e - 1;
      int k = n++;
      while (k > 0 && survivor_order(R, L, U, ids[k - 1], id) > 0) { ids[k] = ids[k - 1]; --k; }
      ids[k] = id;
    }
  }
  int32_t pme = pme_in;
  uint32_t nib = nib0;
  int32_t last = -1;                     // more than ASM_K survivors: selection by repeated minimum
  int bin = 0;                           // next bin of this tile whose first site is not yet known
  constexpr int BPT = DISC_W / IDX_G;
  const int64_t tile_lo = (int64_t)U.win0 + (int64_t)t * DISC_W;
  for (int k = 0; k < cnt; ++k) {
    int32_t id;
    if (small) {
      id = ids[k];
    } else {
      id = -1;
      for (int32_t e = (int32_t)mine.x; e != 0; e = U.next[e - 1]) {
        const int32_t c = e - 1;
        if (last >= 0 && survivor_order(R, L, U, c, last) <= 0) continue;
        if (id < 0 || survivor_order(R, L, U, c, id) < 0) id = c;
      }
      last = id;
    }
    const int32_t row = row0 + k;
    const int32_t st = U.start[id], rl = U.ref_len[id], al = U.alt_len[id];
    while (bin < BPT && tile_lo + (int64_t)bin * IDX_G <= (int64_t)st) {     // bins that begin at or before this site
      if (t * BPT + bin <= nb) T.sidx[t * BPT + bin] = row;
      ++bin;
    }
    pme = max(pme, st + rl - U.win0);
    const uint32_t foot = (uint32_t)(rl + al + 1) & ~1u;
    if (row < T.cap) {
      T.start[row] = st; T.ref_len[row] = rl; T.alt_len[row] = al; T.count[row] = U.count[id];
      T.end[row] = st + rl; T.pme[row] = pme + U.win0;
      T.allele_off[row] = nib;
      if (nib + foot <= T.nib_cap) {
        const int len = rl + al;
        for (int q = 0; q < len; q += 2) {
          const uint32_t hi = survivor_nib(R, L, U, id, q);
          const uint32_t lo = (q + 1 < len) ? survivor_nib(R, L, U, id, q + 1) : 0u;
          T.alleles4[(nib + (uint32_t)q) >> 1] = (uint8_t)((hi << 4) | lo);
        }
      }
    }
    nib += foot;
  }
  for (; bin < BPT; ++bin)
    if (t * BPT + bin <= nb) T.sidx[t * BPT + bin] = row0 + cnt;
  if (t == n_tiles - 1) {                // the table's end: sizes, the sentinel offsets
    const int32_t n = row0 + cnt;
    for (int k = n_tiles * BPT; k <= nb; ++k) T.sidx[k] = n;
    if (n <= T.cap) T.allele_off[n] = nib;
    DSiteHdr h;
    h.n = n; h.n_nibbles = nib; h.s_begin = 0; h.s_end = n;
    int32_t mp = 1;                                              // TreeRegionJoin.scala:66-72
    while (!(2 * (int64_t)mp >= (int64_t)n)) mp *= 2;
    h.midpoint = mp;
    h.reserved[0] = h.reserved[1] = h.reserved[2] = 0;
    *T.hdr = h;
  }
}

cudaError_t launch_assemble_sites(const DReads& R, const DIndelList& L, const DDiscoverOut& U,
                                  const BatchStats& hstats, int32_t nb, const DSiteTableOut& T,
                                  cudaStream_t s, int* n_launches) {
  const int32_t n_tiles = discover_tiles(hstats);
  k_assemble_sites<<<(n_tiles + ASM_THREADS - 1) / ASM_THREADS, ASM_THREADS, 0, s>>>(R, L, U, n_tiles, nb, T);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

}  // namespace avo
