// avo_discover.cu -- DiscoverVariants on the GPU.
//
// Reference semantics (avocado-core/src/main/scala/org/bdgenomics/avocado/genotyping/
// DiscoverVariants.scala): per-read candidate emission :112-252 (the operator walk below), then
// `distinct` or groupBy(contig,start,ref,alt).count().where(count > minObservations) :71-102.
//
// k_discover_tiles is a persistent, warp-specialised kernel.  A CTA owns windows ("tiles") of DISC_W
// reference positions, pulled from a global counter.  The reads that can have candidates in a tile are
// one contiguous index range of the (sorted) batch, and so are their operators and mismatch bytes: one
// PRODUCER warp streams them into shared memory in chunks of DISC_CHUNK reads with bulk async copies
// (cp.async.bulk, completion on an mbarrier), a ring of DISC_STAGES stages deep and running ahead of
// the tile boundaries, while the CONSUMER warps walk one read per thread out of shared memory.  SNV
// candidates (the bulk: every sequencing error produces one) never leave the SM: they are queued in
// shared memory during the walk and counted per position and alternate base by all lanes when the
// tile's last chunk is done; a position whose count exceeds the threshold is emitted.  Tiles with
// non-ACGT candidate bases, reads that disagree about a reference base, or more candidates than the
// queue holds take an exact (ref, alt) histogram path.  Indel candidates are rare; they go to a
// global list and are grouped exactly by a verified hash table (k_indel_count).  Survivors are sorted
// and materialised by assemble_sites().
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "avo_device.cuh"

namespace avo {

// Tile / pipeline shape (sweeps: profiles/README.md).
#ifndef AVO_DISC_CONSUMERS
#define AVO_DISC_CONSUMERS 256
#endif
#ifndef AVO_DISC_W
#define AVO_DISC_W 2048
#endif
#ifndef AVO_DISC_MINB
#define AVO_DISC_MINB 3
#endif
#ifndef AVO_DISC_STAGES
#define AVO_DISC_STAGES 3
#endif
#ifndef AVO_DISC_OPS_CAP
#define AVO_DISC_OPS_CAP 1280           // operators staged per chunk (about 3 per read); the rest is read in place
#endif
#ifndef AVO_DISC_RB_CAP
#define AVO_DISC_RB_CAP 768             // mismatch / deletion bytes staged per chunk (about 1 per read)
#endif
#ifndef AVO_DISC_CQ_ROWS
#define AVO_DISC_CQ_ROWS 8              // SNV candidates one consumer thread queues per tile (then: shared overflow)
#endif
constexpr int DISC_CONSUMERS = AVO_DISC_CONSUMERS;
constexpr int DISC_THREADS = DISC_CONSUMERS + 32;      // + the producer warp
constexpr int DISC_CWARPS = DISC_CONSUMERS / 32;
constexpr int DISC_CHUNK = DISC_CONSUMERS;             // reads per chunk: one per consumer thread
constexpr int DISC_W = AVO_DISC_W;                     // reference positions owned by one tile
constexpr int DISC_STAGES = AVO_DISC_STAGES;
constexpr int DISC_OPS_CAP = AVO_DISC_OPS_CAP;
constexpr int DISC_RB_CAP = AVO_DISC_RB_CAP;
constexpr int DISC_LONGQ = 1024;                       // reads of one tile awaiting the second pass
constexpr int DISC_IQ = 128;                           // indel events of one tile kept in shared memory
constexpr int DISC_HB = 2;                             // hot positions histogrammed per exact-path batch
constexpr uint32_t NOT_HOT = 0xFFFFFFFFu;
static_assert(DISC_W % (4 * DISC_CONSUMERS) == 0, "the emission scan reads 4 positions per 16-byte load");
static_assert(DISC_W % IDX_G == 0, "tiles are whole index bins");
static_assert(DISC_OPS_CAP % 4 == 0 && DISC_RB_CAP % 16 == 0, "bulk copies move 16-byte units");

__device__ __forceinline__ uint32_t mix32(uint32_t h, uint32_t v) {
  h ^= v;
  h *= 0x9E3779B1u;
  h ^= h >> 15;
  return h;
}

// ---- mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP) ---------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_addr(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}
// the producer's form: a failed probe backs off instead of taking issue slots from the consumers
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_addr(bar);
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) return;
    __nanosleep(100);
  }
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
// barrier among the consumer threads only (the producer warp never joins it)
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(DISC_CONSUMERS) : "memory"); }

// DiscoverVariants.scala:112-252, one operator at a time, reading the batch in place (global memory):
// the exact path, the indel handler and reads the staging does not cover use these.  The hot loop
// carries only what SNV candidates need (reference position, refb cursor, the fastForward flag);
// only candidate emission diverges: snp(pos, (ref << 4) | alt) for every mismatching base that
// passes the phred threshold, and indel(k, anchor) for every I / D operator after the first aligned
// base.  Indel candidates are rare; their read index, lastRef and quality window are recomputed by
// disc_handle_indel.
struct WalkState {
  int32_t pos;       // reference position of the next aligned base
  uint32_t rb;       // byte index into refb
  bool ff;           // still in fastForward (:139-172): no aligned operator seen yet
};

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
__device__ __forceinline__ void walk_discover(const DReads& R, const MetaA& a, const MetaB& m, int32_t thr,
                                              SnpF&& snp, IndelF&& indel) {
  WalkState st;
  st.pos = a.start; st.rb = m.refb_off; st.ff = true;
  for (uint32_t k = 0; k < m.n_ops; ++k) {
    const uint32_t op = __ldg(R.ops + a.op_off + k);
    const uint32_t len = op >> 4, code = op & 15u;
    const bool isX = code == AVO_OP_MISMATCH, isD = code == AVO_OP_DEL;
    const bool aligned = code <= AVO_OP_MISMATCH;                           // MATCH or MISMATCH
    if (isX) walk_mismatch(R, len, st.pos, st.rb, a.seq_off, m.qhead, thr, snp);
    if ((code == AVO_OP_INS || isD) && !st.ff) indel(k, st.pos - 1);        // :215-242
    st.pos += (aligned || isD) ? (int32_t)len : 0;
    st.rb += isX ? len : isD ? ((len + 1) >> 1) : 0u;
    st.ff = st.ff && !aligned;
  }
}

// An I / D operator awaiting its quality test: (read, operator index)
struct IndelEvent { int32_t r; uint32_t k; };

// What the producer tells the consumers about a staged chunk
struct ChunkHeader {
  int32_t tile;          // < 0: no more work
  int32_t r0, n;         // reads [r0, r0 + n)
  int32_t pre;           // 1: the record of read r0 - 1 was copied in front (sort check)
  int32_t r_own;         // reads from here on start inside the tile's window (each record has one owner tile)
  int32_t r_lo, r_hi;    // the tile's whole read range (exact path)
  int32_t wlo;           // first reference position of the window
  int32_t last;          // 1: the tile's last chunk
  uint32_t op_lo, op_n;  // operators [op_lo, op_lo + op_n) are staged
  uint32_t rb_lo, rb_n;  // mismatch bytes [rb_lo, rb_lo + rb_n) are staged
};

struct __align__(128) DiscStage {
  uint4 meta[2 * (DISC_CHUNK + 2)];     // 32-byte records: [pre] + n + [one past the chunk]
  uint32_t ops[DISC_OPS_CAP];
  uint8_t refb[DISC_RB_CAP];
  ChunkHeader hdr;
};

constexpr int DISC_CQ_ROWS = AVO_DISC_CQ_ROWS;         // private candidate slots per consumer thread and tile
constexpr int DISC_CQ_OVER = 256;                      // shared overflow of the private slots

struct __align__(128) DiscSmem {
  DiscStage stage[DISC_STAGES];
  // SNV candidate counts per position and alternate base, two 16-bit fields per word
  // (cnt[0]: A | C << 16, cnt[1]: G | T << 16).  exact path: cnt[0] = candidates per position, then
  // hot index or NOT_HOT; cnt[1] holds its (ref << 4 | alt) histograms.
  uint32_t cnt[2][DISC_W];
  uint8_t refs[DISC_W];                 // reference nibble of the position's candidates (0: none yet)
  // queued candidates, (window offset << 8) | (ref << 4 | alt): slot k of consumer thread t is
  // cq[k * DISC_CONSUMERS + t] (no atomics, no bank conflicts); what does not fit goes to cq_over
  uint32_t cq[DISC_CQ_ROWS * DISC_CONSUMERS];
  uint32_t cq_over[DISC_CQ_OVER];
  int32_t longq[DISC_LONGQ];            // reads of the tile that are walked operator by operator (second pass)
  IndelEvent iq[2][DISC_IQ];            // by tile parity, like the counters below
  uint64_t meta_full[DISC_STAGES];      // the chunk's records have landed
  uint64_t data_full[DISC_STAGES];      // ... and its operators / mismatch bytes (and the header is final)
  uint64_t empty[DISC_STAGES];          // every consumer warp is done with the stage
  int32_t scan[DISC_CWARPS];
  int32_t n_over[2], n_iq[2], n_long[2];
  int32_t exotic[2];                    // a candidate with a non-ACGT base was seen: exact path
  int32_t bad[2];                       // reads disagree about a reference base: exact path
  int32_t hotpos[DISC_HB];
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
__device__ __noinline__ void disc_handle_indel(const DReads& R, const DIndelList& L, int32_t r, uint32_t k, int32_t thr) {
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

// A C G T are the nibbles 1 2 4 8 (bits of 0x116)
__device__ __forceinline__ bool acgt_pair(uint32_t pair) {
  return (((0x116u >> ((pair >> 4) & 15u)) & (0x116u >> (pair & 15u))) & 1u) != 0;
}

// exclusive prefix sum over the consumer threads (ctid = consumer thread index)
__device__ __forceinline__ int consumer_scan(DiscSmem& sm, int ctid, int v, int& total) {
  const int lane = ctid & 31, w = ctid >> 5;
  int x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int y = __shfl_up_sync(0xFFFFFFFFu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) sm.scan[w] = x;
  consumer_sync();
  int base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < DISC_CWARPS; ++i) {
    const int t = sm.scan[i];
    if (i < w) base += t;
    tot += t;
  }
  consumer_sync();
  total = tot;
  return base + x - v;
}

// ---- the exact path of one tile (non-ACGT candidate bases, inconsistent reference bases, a full
// candidate queue, or extreme depth).  Re-walks the tile's reads in place for per-position candidate
// totals, then builds (ref, alt) histograms of the positions whose total exceeds the threshold,
// DISC_HB positions at a time.  Entered and left by all consumer threads; leaves the counters clear.
__device__ __noinline__ void disc_exact_tile(const DReads& R, DiscSmem& sm, int ctid, int32_t wlo, int32_t r_lo,
                                             int32_t r_hi, int32_t thr, int32_t min_count, const DDiscoverOut& out) {
  constexpr int PER_T = DISC_W / DISC_CONSUMERS;
  uint32_t (*hist)[256] = reinterpret_cast<uint32_t (*)[256]>(&sm.cnt[1][0]);
  static_assert(DISC_HB * 256 <= DISC_W, "the histograms live in cnt[1]");
  const int lane = ctid & 31, warp = ctid >> 5;
  consumer_sync();
  for (int i = ctid; i < 2 * DISC_W; i += DISC_CONSUMERS) (&sm.cnt[0][0])[i] = 0;
  consumer_sync();
  for (int32_t r = r_lo + ctid; r < r_hi; r += DISC_CONSUMERS) {
    walk_discover(
        R, load_meta_a(R.meta, r), load_meta_b(R.meta, r), thr,
        [&](int32_t pos, uint32_t) {
          const uint32_t o = (uint32_t)pos - (uint32_t)wlo;
          if (pos >= wlo && o < (uint32_t)DISC_W) atomicAdd(&sm.cnt[0][o], 1u);
        },
        [&](uint32_t, int32_t) {});
  }
  consumer_sync();
  int nh = 0;
  uint32_t hot_mask = 0;
  for (int i = 0; i < PER_T; ++i) {
    const bool hot = (int32_t)sm.cnt[0][ctid * PER_T + i] > min_count;
    hot_mask |= (uint32_t)hot << i;
    nh += hot;
  }
  int htotal;
  int hoff = consumer_scan(sm, ctid, nh, htotal);
  for (int i = 0; i < PER_T; ++i) {
    const bool hot = (hot_mask >> i) & 1u;
    sm.cnt[0][ctid * PER_T + i] = hot ? (uint32_t)hoff : NOT_HOT;
    hoff += hot;
  }
  consumer_sync();
  for (int hb = 0; hb < htotal; hb += DISC_HB) {
    for (int i = ctid; i < DISC_HB * 256; i += DISC_CONSUMERS) (&hist[0][0])[i] = 0;
    consumer_sync();
    for (int32_t r = r_lo + ctid; r < r_hi; r += DISC_CONSUMERS) {
      walk_discover(
          R, load_meta_a(R.meta, r), load_meta_b(R.meta, r), thr,
          [&](int32_t pos, uint32_t pair) {
            const uint32_t o = (uint32_t)pos - (uint32_t)wlo;
            if (pos < wlo || o >= (uint32_t)DISC_W) return;
            const uint32_t h = sm.cnt[0][o];
            if (h == NOT_HOT || (int)h < hb || (int)h >= hb + DISC_HB) return;
            atomicAdd(&hist[h - hb][pair & 255u], 1u);
          },
          [&](uint32_t, int32_t) {});
    }
    consumer_sync();
    for (int i = ctid; i < DISC_W; i += DISC_CONSUMERS) {
      const uint32_t h = sm.cnt[0][i];
      if (h != NOT_HOT && (int)h >= hb && (int)h < hb + DISC_HB) sm.hotpos[h - hb] = i;
    }
    consumer_sync();
    // a warp scans the 256 (ref, alt) bins of one hot position of this batch
    for (int s = warp; s < min(DISC_HB, htotal - hb); s += DISC_CWARPS) {
      const int i = sm.hotpos[s];
      for (int k = lane; k < 256; k += 32) {
        const int32_t c = (int32_t)hist[s][k];
        if (c > min_count) disc_emit_snv(out, wlo + i, (uint32_t)k, c);
      }
    }
    consumer_sync();
  }
  for (int i = ctid; i < 2 * DISC_W; i += DISC_CONSUMERS) (&sm.cnt[0][0])[i] = 0;
  for (int i = ctid; i < DISC_W / 4; i += DISC_CONSUMERS) reinterpret_cast<uint32_t*>(sm.refs)[i] = 0;
}

// ---- the producer warp (one lane works): tiles from the global counter, their reads in chunks -----
__device__ void disc_producer(const DReads& R, const DIndex& X, DiscSmem& sm, int32_t n_tiles, int32_t max_span,
                              int32_t verify, int32_t* __restrict__ tile_counter, int32_t* __restrict__ flags) {
  constexpr int BPT = DISC_W / IDX_G;
  // 16-byte aligned columns can be bulk-copied; anything else is read in place by the consumers
  const bool ops_ok = ((uintptr_t)R.ops & 15u) == 0, rb_ok = ((uintptr_t)R.refb & 15u) == 0;
  uint32_t j = 0;            // chunks issued so far
  bool pending = false;      // chunk j - 1 still waits for its operators / mismatch bytes
  auto issue_data = [&](uint32_t jj) {
    DiscStage& st = sm.stage[jj % DISC_STAGES];
    mbar_wait_backoff(&sm.meta_full[jj % DISC_STAGES], (jj / DISC_STAGES) & 1u);
    ChunkHeader& h = st.hdr;
    const uint4 first = st.meta[2 * h.pre], first_b = st.meta[2 * h.pre + 1];
    const bool post = (int64_t)h.r0 + h.n < R.n;
    const uint32_t op_lo = first.w, rb_lo = first_b.x;
    uint32_t op_hi = post ? st.meta[2 * (h.pre + h.n)].w : (uint32_t)R.n_ops;
    uint32_t rb_hi = post ? st.meta[2 * (h.pre + h.n) + 1].x : (uint32_t)R.n_refb;
    // spans of 16-byte units inside the arrays (a batch whose offsets do not ascend - a merged union -
    // stages a clamped window; what falls outside is read in place)
    uint32_t a_lo = op_lo & ~3u, a_hi = (op_hi + 3u) & ~3u;
    const uint32_t a_end = (uint32_t)R.n_ops & ~3u;
    if (op_hi < op_lo || a_hi - a_lo > (uint32_t)DISC_OPS_CAP) a_hi = a_lo + (uint32_t)DISC_OPS_CAP;
    if (a_hi > a_end) a_hi = a_end;
    if (!ops_ok || a_hi < a_lo || a_lo > a_end) a_hi = a_lo;
    uint32_t b_lo = rb_lo & ~15u, b_hi = (rb_hi + 15u) & ~15u;
    const uint32_t b_end = (uint32_t)R.n_refb & ~15u;
    if (rb_hi < rb_lo || b_hi - b_lo > (uint32_t)DISC_RB_CAP) b_hi = b_lo + (uint32_t)DISC_RB_CAP;
    if (b_hi > b_end) b_hi = b_end;
    if (!rb_ok || b_hi < b_lo || b_lo > b_end) b_hi = b_lo;
    h.op_lo = a_lo; h.op_n = a_hi - a_lo;
    h.rb_lo = b_lo; h.rb_n = b_hi - b_lo;
    uint64_t* bar = &sm.data_full[jj % DISC_STAGES];
    mbar_arrive_expect_tx(bar, (a_hi - a_lo) * 4u + (b_hi - b_lo));
    if (a_hi > a_lo) bulk_g2s(st.ops, R.ops + a_lo, (a_hi - a_lo) * 4u, bar);
    if (b_hi > b_lo) bulk_g2s(st.refb, R.refb + b_lo, b_hi - b_lo, bar);
  };
  // Tile ids are claimed two ahead and the index entries of a tile are loaded one tile ahead, so that
  // neither the atomic's nor the loads' latency sits between two chunks.
  struct TileInfo { int32_t tile, r_lo, r_hi, r_own; };
  auto tile_info = [&](int32_t tile) {
    TileInfo t;
    t.tile = tile; t.r_lo = t.r_hi = t.r_own = 0;
    if (tile < n_tiles) {
      const int64_t wlo = (int64_t)X.win0 + (int64_t)tile * DISC_W;
      // candidates of a read lie in [start, end): reads with start < whi and start > wlo - max_span
      t.r_lo = __ldg(X.ridx + idx_bin_floor(X, wlo - max_span));
      t.r_hi = __ldg(X.ridx + min((tile + 1) * BPT, X.nb));
      // the records whose start lies in this window: every record is owned by exactly one tile as
      // long as the index is monotone (checked below), so the consumers' checks see every record once
      t.r_own = __ldg(X.ridx + min(tile * BPT, X.nb));
    }
    return t;
  };
  TileInfo cur = tile_info(atomicAdd(tile_counter, 1));
  int32_t id_next = atomicAdd(tile_counter, 1);
  while (cur.tile < n_tiles) {
    const TileInfo nxt = tile_info(id_next);          // loads in flight while this tile's chunks are issued
    id_next = atomicAdd(tile_counter, 1);
    const int32_t tile = cur.tile, r_lo = cur.r_lo, r_hi = cur.r_hi, r_own = cur.r_own;
    const int64_t wlo = (int64_t)X.win0 + (int64_t)tile * DISC_W;
    if (verify) {
      if (r_own > r_hi || (tile == 0 && r_own != 0) || (tile == 0 && R.n > 0 && meta_start(R.meta, 0) < X.win0))
        atomicOr(flags, DISC_FLAG_UNSORTED);
    }
    for (int32_t r0 = r_lo; r0 < r_hi; r0 += DISC_CHUNK) {
      const uint32_t s = j % DISC_STAGES;
      mbar_wait_backoff(&sm.empty[s], ((j / DISC_STAGES) & 1u) ^ 1u);
      DiscStage& st = sm.stage[s];
      ChunkHeader& h = st.hdr;
      h.tile = tile; h.r0 = r0; h.n = min(DISC_CHUNK, r_hi - r0);
      h.pre = r0 > 0; h.r_own = r_own; h.r_lo = r_lo; h.r_hi = r_hi; h.wlo = (int32_t)wlo;
      h.last = r0 + DISC_CHUNK >= r_hi;
      const int32_t n_rec = h.n + h.pre + (((int64_t)r0 + h.n < R.n) ? 1 : 0);
      mbar_arrive_expect_tx(&sm.meta_full[s], (uint32_t)n_rec * 32u);
      bulk_g2s(st.meta, R.meta + (r0 - h.pre), (uint32_t)n_rec * 32u, &sm.meta_full[s]);
      if (pending) issue_data(j - 1);
      pending = true;
      ++j;
    }
    cur = nxt;
  }
  if (pending) issue_data(j - 1);
  // terminator
  const uint32_t s = j % DISC_STAGES;
  mbar_wait_backoff(&sm.empty[s], ((j / DISC_STAGES) & 1u) ^ 1u);
  sm.stage[s].hdr.tile = -1;
  mbar_arrive(&sm.meta_full[s]);
  mbar_arrive(&sm.data_full[s]);
}

__global__ void __launch_bounds__(DISC_THREADS, AVO_DISC_MINB)
k_discover_tiles(DReads R, DIndex X, int32_t thr, int32_t min_count /* keep count > min_count */,
                 int32_t n_tiles, int32_t max_span, int32_t max_end, int32_t verify, DDiscoverOut out,
                 DIndelList L, int32_t* __restrict__ tile_counter, int32_t* __restrict__ flags) {
  extern __shared__ __align__(128) unsigned char disc_smem_raw[];     // sizeof(DiscSmem), more than 48 KB
  DiscSmem& sm = *reinterpret_cast<DiscSmem*>(disc_smem_raw);
  const int tid = threadIdx.x;
  constexpr int PER_T = DISC_W / DISC_CONSUMERS;

  for (int i = tid; i < 2 * DISC_W; i += DISC_THREADS) (&sm.cnt[0][0])[i] = 0;
  for (int i = tid; i < DISC_W / 4; i += DISC_THREADS) reinterpret_cast<uint32_t*>(sm.refs)[i] = 0;
  if (tid == 0) {
    for (int s = 0; s < DISC_STAGES; ++s) {
      mbar_init(&sm.meta_full[s], 1);
      mbar_init(&sm.data_full[s], 1);
      mbar_init(&sm.empty[s], DISC_CWARPS);
    }
    for (int q = 0; q < 2; ++q) { sm.n_over[q] = 0; sm.n_iq[q] = 0; sm.n_long[q] = 0; sm.exotic[q] = 0; sm.bad[q] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (tid >= DISC_CONSUMERS) {                  // ---- producer warp
    if (tid == DISC_CONSUMERS) disc_producer(R, X, sm, n_tiles, max_span, verify, tile_counter, flags);
    return;
  }

  // ---- consumer warps: one read per thread out of the staged chunk
  const int ctid = tid, lane = tid & 31;
  const uint32_t in_head = R.qhead_short ? 2u : 4u;
  uint32_t my_n = 0;                      // candidates in this thread's private slots (current tile)
  uint32_t par_t = 0;                     // parity of the current tile (indel queue / flags)
  // one SNV candidate: into the thread's own slots, else the shared overflow, else the exact path
  auto push = [&](uint32_t e) {
    if (my_n < (uint32_t)DISC_CQ_ROWS) {
      sm.cq[my_n * DISC_CONSUMERS + ctid] = e;
      ++my_n;
    } else {
      const int slot = atomicAdd(&sm.n_over[par_t], 1);
      if (slot < DISC_CQ_OVER) sm.cq_over[slot] = e;
    }
  };
  auto snp = [&](int32_t wlo, int32_t pos, uint32_t pair) {
    const uint32_t o = (uint32_t)pos - (uint32_t)wlo;
    if (pos >= wlo && o < (uint32_t)DISC_W) push((o << 8) | (pair & 255u));
  };
  auto indel_event = [&](int32_t wlo, int32_t r, uint32_t k, int32_t anchor) {
    if (anchor < wlo || (uint32_t)(anchor - wlo) >= (uint32_t)DISC_W) return;
    const int slot = atomicAdd(&sm.n_iq[par_t], 1);
    if (slot < DISC_IQ) { sm.iq[par_t][slot].r = r; sm.iq[par_t][slot].k = k; }
    else disc_handle_indel(R, L, r, k, thr);
  };
  // count one queued candidate (all lanes; the tile's walk is over for this thread)
  auto count_candidate = [&](uint32_t e) {
    const uint32_t o = e >> 8, refn = (e >> 4) & 15u, altn = e & 15u;
    if (!acgt_pair(e)) { sm.exotic[par_t] = 1; return; }
    const int ai = (int)((altn >> 1) - (altn >> 3));                // 1 2 4 8 -> 0 1 2 3
    atomicAdd(&sm.cnt[ai >> 1][o], 1u << ((ai & 1) * 16));
    const uint32_t v = sm.refs[o];                                  // racy on purpose: settled after the barrier
    if (v == 0) sm.refs[o] = (uint8_t)refn;
    else if (v != refn) sm.bad[par_t] = 1;
  };
  auto check_candidate = [&](uint32_t e) {
    if (acgt_pair(e) && sm.refs[e >> 8] != ((e >> 4) & 15u)) sm.bad[par_t] = 1;
  };

  for (uint32_t j = 0;; ++j) {
    const uint32_t s = j % DISC_STAGES, par = (j / DISC_STAGES) & 1u;
    mbar_wait(&sm.meta_full[s], par);
    mbar_wait(&sm.data_full[s], par);
    const DiscStage& st = sm.stage[s];
    const ChunkHeader& h = st.hdr;
    if (h.tile < 0) break;
    const int32_t wlo = h.wlo;
    const int32_t r_lo = h.r_lo, r_hi = h.r_hi;
    const bool last = h.last != 0;
    // ---- pass 1: every thread looks at its read.  Most reads are one match, or one mismatch run after
    // the leading match ("=", "= X", "= X =", "= X S": about three quarters), and are finished on the
    // spot in straight-line code; the others are queued and walked operator by operator in pass 2 by
    // as many threads as there are such reads, so that no lane idles while a neighbour loops.
    bool is_long = false;
    if (ctid < h.n) {
      const int32_t r = h.r0 + ctid;
      const int mi = 2 * (ctid + h.pre);
      const uint4 va = st.meta[mi], vb = st.meta[mi + 1];
      const int32_t start = (int32_t)va.x, end = (int32_t)va.y;
      const uint32_t n_ops = vb.y & 0xFFFFu;
      if (r >= h.r_own) {
        if (R.coords) R.coords[r] = make_int2(start, end);   // dense (start, end) copy for the join
        if (verify) {     // sort order and the caller's bounds (avo_batch_stats)
          if (r > 0 && (int32_t)st.meta[mi - 2].x > start) atomicOr(flags, DISC_FLAG_UNSORTED);
          if (end > max_end || (int64_t)end - start > max_span) atomicOr(flags, DISC_FLAG_BAD_STATS);
        }
      }
      if (end > wlo && n_ops != 0) {                        // else: ends before the window / no operators
        const uint32_t o0 = va.w - h.op_lo;
        if (n_ops <= 3u && o0 <= h.op_n && n_ops <= h.op_n - o0) {
          const uint32_t w0 = st.ops[o0];
          const uint32_t w1 = n_ops > 1u ? st.ops[o0 + 1] : (uint32_t)(0x10u | AVO_OP_MISMATCH);
          const uint32_t w2 = n_ops > 2u ? st.ops[o0 + 2] : 0u;
          const uint32_t f0 = vb.x - h.rb_lo;
          // "=", "= X", "= X =", "= X S" with a one-base mismatch whose byte is staged
          if ((w0 & 15u) == AVO_OP_MATCH && w1 == (0x10u | AVO_OP_MISMATCH) && ((w2 & 15u) | 4u) == 4u &&
              (n_ops == 1u || f0 < h.rb_n)) {
            if (n_ops > 1u && (int32_t)(vb.z & 255u) >= thr)   // qual(0): the first quality of the read (:199)
              snp(wlo, start + (int32_t)(w0 >> 4), st.refb[f0]);
          } else {
            is_long = true;
          }
        } else {
          is_long = true;
        }
      }
    }
    {
      const unsigned m = __ballot_sync(0xFFFFFFFFu, is_long);
      if (m) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&sm.n_long[par_t], __popc(m));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        const int slot = base + __popc(m & ((1u << lane) - 1u));
        if (is_long) {
          if (slot < DISC_LONGQ) {
            sm.longq[slot] = h.r0 + ctid;
          } else {                                          // queue full: walked here, in place
            const int32_t r = h.r0 + ctid;
            walk_discover(R, load_meta_a(R.meta, r), load_meta_b(R.meta, r), thr,
                          [&](int32_t pos, uint32_t pair) { snp(wlo, pos, pair); },
                          [&](uint32_t k, int32_t anchor) { indel_event(wlo, r, k, anchor); });
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[s]);
    if (!last) continue;

    // ---- the tile's last chunk is walked.  Barrier 1: every thread has left the previous tile (its
    // counters are clear, its flags read) and queued this tile's candidates; the other parity's flags
    // (last used by the previous tile, next by the next one) are reset here.
    consumer_sync();
    if (ctid == 0) {
      sm.n_over[par_t ^ 1u] = 0; sm.n_iq[par_t ^ 1u] = 0; sm.n_long[par_t ^ 1u] = 0;
      sm.exotic[par_t ^ 1u] = 0; sm.bad[par_t ^ 1u] = 0;
    }
    // pass 2: the queued reads, one per thread, read in place (their lines were just streamed through L2)
    {
      const int nl = min(sm.n_long[par_t], DISC_LONGQ);
      for (int i = ctid; i < nl; i += DISC_CONSUMERS) {
        const int32_t r = sm.longq[i];
        walk_discover(R, load_meta_a(R.meta, r), load_meta_b(R.meta, r), thr,
                      [&](int32_t pos, uint32_t pair) { snp(wlo, pos, pair); },
                      [&](uint32_t k, int32_t anchor) { indel_event(wlo, r, k, anchor); });
      }
    }
    for (uint32_t k = 0; k < my_n; ++k) count_candidate(sm.cq[k * DISC_CONSUMERS + ctid]);
    consumer_sync();                                   // (indel events and overflow of pass 2 are queued)
    const int n_over_all = sm.n_over[par_t];
    const int n_over = min(n_over_all, DISC_CQ_OVER);
    for (int i = ctid; i < n_over; i += DISC_CONSUMERS) count_candidate(sm.cq_over[i]);
    {
      const int nq = min(sm.n_iq[par_t], DISC_IQ);
      for (int i = ctid; i < nq; i += DISC_CONSUMERS) disc_handle_indel(R, L, sm.iq[par_t][i].r, sm.iq[par_t][i].k, thr);
    }
    consumer_sync();                                   // barrier 2: every candidate is counted
    // every candidate agrees with the reference nibble its position ended up with, or the tile is exact-path
    for (uint32_t k = 0; k < my_n; ++k) check_candidate(sm.cq[k * DISC_CONSUMERS + ctid]);
    for (int i = ctid; i < n_over; i += DISC_CONSUMERS) check_candidate(sm.cq_over[i]);
    my_n = 0;
    // A position is hot when one of its 16-bit counters exceeds min_count (tested four positions at a
    // time; fields stay below 32768: fewer than 32768 reads reach a fast-path tile).
    const uint32_t bias = (uint32_t)(0x7FFF - min(min_count, 0x7FFF)) * 0x00010001u;
    uint32_t hot = 0;                      // bit (4 j + e): position 4 (ctid + j * DISC_CONSUMERS) + e is hot
#pragma unroll
    for (int jj = 0; jj < PER_T / 4; ++jj) {
      const int o4 = ctid + jj * DISC_CONSUMERS;      // 16 contiguous bytes per lane: no bank conflicts
      const uint4 vAC = reinterpret_cast<const uint4*>(sm.cnt[0])[o4];
      const uint4 vGT = reinterpret_cast<const uint4*>(sm.cnt[1])[o4];
      const uint32_t wa[4] = {vAC.x, vAC.y, vAC.z, vAC.w};
      const uint32_t wg[4] = {vGT.x, vGT.y, vGT.z, vGT.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (((wa[e] + bias) | (wg[e] + bias)) & 0x80008000u) hot |= 1u << (4 * jj + e);
    }
    consumer_sync();                                   // barrier 3: the tile's flags are final
    const bool slow = sm.exotic[par_t] != 0 || sm.bad[par_t] != 0 || n_over_all > DISC_CQ_OVER ||
                      (r_hi - r_lo) >= 32768 || min_count >= 32767;
    if (!slow) {
      for (uint32_t hh = hot; hh; hh &= hh - 1) {
        const int b = __ffs(hh) - 1;
        const int o = 4 * (ctid + (b >> 2) * DISC_CONSUMERS) + (b & 3);
        const uint32_t wAC = sm.cnt[0][o], wGT = sm.cnt[1][o];
        const uint32_t refn = sm.refs[o];
        const int32_t c[4] = {(int32_t)(wAC & 0xFFFFu), (int32_t)(wAC >> 16), (int32_t)(wGT & 0xFFFFu),
                              (int32_t)(wGT >> 16)};
        for (int ai = 0; ai < 4; ++ai)
          if (c[ai] > min_count) disc_emit_snv(out, wlo + o, (refn << 4) | (1u << ai), c[ai]);
      }
      // every thread clears the positions it scanned (the next tile's counting starts after its own
      // first barrier, which this thread reaches only after this)
      const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int jj = 0; jj < PER_T / 4; ++jj) {
        reinterpret_cast<uint4*>(sm.cnt[0])[ctid + jj * DISC_CONSUMERS] = z;
        reinterpret_cast<uint4*>(sm.cnt[1])[ctid + jj * DISC_CONSUMERS] = z;
        reinterpret_cast<uint32_t*>(sm.refs)[ctid + jj * DISC_CONSUMERS] = 0;
      }
    } else {
      disc_exact_tile(R, sm, ctid, wlo, r_lo, r_hi, thr, min_count, out);
    }
    par_t ^= 1u;
  }
}

cudaError_t launch_discover_tiles(const DReads& R, int32_t thr, int32_t min_obs, const DIndex& X,
                                  const BatchStats& hstats, int verify, const DDiscoverOut& out,
                                  const DIndelList& indels, int32_t* d_tile_counter, int32_t* d_flags,
                                  int num_sms, int* blocks_per_sm /* per context; 0 = not yet known */,
                                  cudaStream_t s, int* n_launches) {
  cudaError_t e = cudaMemsetAsync(d_tile_counter, 0, sizeof(int32_t), s);
  if (e != cudaSuccess) return e;
  if (R.n == 0) return cudaSuccess;
  const int64_t span = (int64_t)hstats.max_end - (int64_t)hstats.min_start;
  const int32_t n_tiles = (int32_t)max((int64_t)1, (span + DISC_W - 1) / DISC_W);
  if (*blocks_per_sm == 0) {             // once per context (the attribute belongs to the device's context)
    e = cudaFuncSetAttribute(k_discover_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DiscSmem));
    if (e != cudaSuccess) return e;
    int b = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_discover_tiles, DISC_THREADS, sizeof(DiscSmem));
    if (e != cudaSuccess) return e;
    *blocks_per_sm = max(1, b);
  }
  const int grid = (int)min((int64_t)n_tiles, (int64_t)num_sms * *blocks_per_sm);
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
                             DDiscoverOut out, int32_t* __restrict__ nib, int32_t* __restrict__ nsurv) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tsize) return;
  const unsigned long long w = table[i];
  if (w == 0ull) return;
  const int32_t c = tcount[i];
  if (c <= min_count) return;
  const int32_t rec = (int32_t)(w & 0xFFFFFFFFull) - 1;
  const int32_t ref_len = L.ref_len[rec], alt_len = L.alt_len[rec];
  atomicAdd(nib, (int32_t)(((uint32_t)ref_len + (uint32_t)alt_len + 1u) & ~1u));
  atomicAdd(nsurv, 1);
  const int32_t slot = atomicAdd(out.n, 1);
  if (slot >= out.capacity) return;
  out.start[slot] = L.pos[rec];
  out.ref_len[slot] = ref_len;
  out.alt_len[slot] = alt_len;
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
