// avo_packdev.cu -- the packer on the device: AlignmentRecord text columns (CIGAR, MD tag,
// sequence, phred+33 qualities) -> the packed batch of include/avocado_b200.h, written in HBM.
//
// Same contract as the host packer (avo_pack.cpp), byte for byte:
// ObservationOperator.extractAlignmentOperators
// (avocado-core/src/main/scala/org/bdgenomics/avocado/models/ObservationOperator.scala:42-171) merges
// the CIGAR with the MD tag into =/X/I/D/S/H runs; a read whose operators cannot be extracted is
// packed with zero operators (DiscoverVariants.scala:119-127, BiallelicGenotyper.scala:385-391); a
// read observeRead would throw on keeps its operators for discovery and gets AVO_READ_SKIP_CALL.
//
// Text arrives at PCIe rate (about 400 bytes per 150-bp read); three kernels turn it into the
// packed columns:
//   k_pack_count    one thread per read: validate + walk CIGAR and MD, count operators / bytes
//   (one cub scan over the four counts gives every read its record, block, operator and byte offset)
//   k_pack_records  one thread per read: walk again, write operators, mismatch bytes and the record
//   k_pack_payload  16 lanes per read: one lane per 16-byte payload block (10 bases + 10 qualities)
#include <cub/cub.cuh>

#include "avo_internal.h"

namespace avo {

namespace {

// BAM 4-bit base codes "=ACMGRSVTWYHKDBN", either case; everything else is N (15)
__device__ __forceinline__ uint32_t nib_of(unsigned char c) {
  switch (c | 0x20) {
    case 'a': return 1; case 'c': return 2; case 'm': return 3; case 'g': return 4;
    case 'r': return 5; case 's': return 6; case 'v': return 7; case 't': return 8;
    case 'w': return 9; case 'y': return 10; case 'h': return 11; case 'k': return 12;
    case 'd': return 13; case 'b': return 14; default: return c == '=' ? 0u : 15u;
  }
}

__device__ __forceinline__ bool is_digit(char c) { return (unsigned)(c - '0') <= 9u; }

// the letters ADAM's MdTag accepts, either case
__device__ __forceinline__ bool is_md_base(char c) {
  const unsigned u = (unsigned)((c | 0x20) - 'a');
  //                         a      b      c      d      g      h      k       m       n       r       s       t       u       v       w       x
  constexpr unsigned MASK = 1u | 1u << 1 | 1u << 2 | 1u << 3 | 1u << 6 | 1u << 7 | 1u << 10 | 1u << 12 | 1u << 13 | 1u << 17 |
                            1u << 18 | 1u << 19 | 1u << 20 | 1u << 21 | 1u << 22 | 1u << 23;
  return u < 26u && ((MASK >> u) & 1u) && ((c | 0x20) >= 'a');
}

__device__ __forceinline__ char upper(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }

// MD grammar: number (('^')? bases+ number)*; every number at most 2^31 - 1
__device__ bool md_valid(const char* s, int n) {
  if (n == 0) return true;
  int off = 0;
  auto number = [&]() -> bool {
    int j = off;
    uint64_t v = 0;
    while (j < n && is_digit(s[j])) { v = v * 10 + (uint64_t)(s[j] - '0'); if (v > 0x7FFFFFFFu) return false; ++j; }
    if (j == off) return false;
    off = j;
    return true;
  };
  if (!number()) return false;
  while (off < n) {
    if (s[off] == '^') ++off;
    int j = off;
    while (j < n && is_md_base(s[j])) ++j;
    if (j == off) return false;
    off = j;
    if (!number()) return false;
  }
  return true;
}

// Cursor over the events of a VALID MD tag: (alignment offset, reference base, deleted?) in
// offset order -- the entries of MdTag's mismatch and deletion maps.
struct MdCursor {
  const char* s;
  int n, off;
  uint64_t ref;     // alignment offset of the base at s[off]
  bool del;
  __device__ void number() {
    uint64_t v = 0;
    while (off < n && is_digit(s[off])) { v = v * 10 + (uint64_t)(s[off] - '0'); ++off; }
    ref += v;
    del = false;
    if (off < n && s[off] == '^') { del = true; ++off; }
  }
  __device__ void begin(const char* s_, int n_) { s = s_; n = n_; off = 0; ref = 0; del = false; if (n) number(); }
  __device__ bool valid() const { return off < n; }
  __device__ char base() const { return upper(s[off]); }
  __device__ void next() {
    ++off; ++ref;
    if (off < n && is_digit(s[off])) number();
  }
};

__device__ bool cigar_valid(const char* s, int n) {
  int i = 0;
  while (i < n) {
    if (!is_digit(s[i])) return false;
    uint64_t v = 0;
    while (i < n && is_digit(s[i])) { v = v * 10 + (uint64_t)(s[i] - '0'); if (v > 0x0FFFFFFFu) return false; ++i; }
    if (i >= n) return false;
    const char op = s[i++];
    if (!(op == 'M' || op == 'I' || op == 'D' || op == 'N' || op == 'S' || op == 'H' || op == 'P' || op == '=' ||
          op == 'X'))
      return false;
  }
  return true;
}

struct Extracted { uint32_t n_ops, n_refb, need; bool ok; };

// extractAlignmentOperators for one read.  WRITE = false counts; WRITE = true also stores the
// operators and the reference bytes (mismatch: (ref << 4) | read base; deletion: nibbles, high first).
template <bool WRITE>
__device__ Extracted extract_ops(const char* cig, int ncig, bool has_cigar, const char* md, int nmd, bool has_md,
                                 const char* seq, int64_t lseq, uint32_t* __restrict__ ops,
                                 uint8_t* __restrict__ refb) {
  Extracted x{0, 0, 0, false};
  if (!has_cigar || (ncig == 1 && cig[0] == '*') || !has_md) return x;
  if (!cigar_valid(cig, ncig) || !md_valid(md, nmd)) return x;
  MdCursor ev;
  ev.begin(md, nmd);
  uint64_t idx = 0;
  uint32_t need = 0, no = 0, nf = 0;
  // discovery's own read index (DiscoverVariants.scala:139-172, 189-193): soft clips count before
  // the first aligned operator and are ignored after it; it picks the read base stored with a mismatch
  uint64_t didx = 0;
  bool aligned_seen = false;
  auto push = [&](uint32_t len, uint32_t code) { if (WRITE) ops[no] = (len << 4) | code; ++no; };
  auto read_nib = [&](uint64_t ri) -> uint32_t { return (int64_t)ri < lseq ? nib_of((unsigned char)seq[ri]) : 15u; };
  int i = 0;
  while (i < ncig) {
    uint32_t len = 0;
    while (is_digit(cig[i])) { len = len * 10 + (uint32_t)(cig[i] - '0'); ++i; }
    const char op = cig[i++];
    if (len == 0 && op != 'S' && op != 'H') return x;    // 0-length M asserts; 0I divides by zero
    switch (op) {
      case 'M': {
        const uint64_t hi = idx + len;
        uint64_t pos = idx;
        while (ev.valid() && ev.ref < idx) ev.next();
        while (ev.valid() && ev.ref < hi) {
          if (ev.del) { ev.next(); continue; }
          const uint64_t k = ev.ref;
          if (k > pos) push((uint32_t)(k - pos), AVO_OP_MATCH);
          uint32_t run = 0;
          while (ev.valid() && !ev.del && ev.ref == k + run && k + run < hi) {
            if (WRITE) refb[nf] = (uint8_t)((nib_of((unsigned char)ev.base()) << 4) | read_nib(didx + (k - idx) + run));
            ++nf; ++run;
            ev.next();
          }
          push(run, AVO_OP_MISMATCH);
          pos = k + run;
        }
        if (pos < hi) push((uint32_t)(hi - pos), AVO_OP_MATCH);
        idx += len; need += len; didx += len; aligned_seen = true;
        break;
      }
      case '=':
        push(len, AVO_OP_MATCH);
        idx += len; need += len; didx += len; aligned_seen = true;
        break;
      case 'X':
        for (uint32_t k = 0; k < len; ++k) {
          while (ev.valid() && ev.ref < idx + k) ev.next();
          if (!ev.valid() || ev.ref != idx + k || ev.del) return x;
          if (WRITE) refb[nf] = (uint8_t)((nib_of((unsigned char)ev.base()) << 4) | read_nib(didx + k));
          ++nf;
        }
        push(len, AVO_OP_MISMATCH);
        idx += len; need += len; didx += len; aligned_seen = true;
        break;
      case 'I':
        push(len, AVO_OP_INS);
        need += len; didx += len;
        break;
      case 'D':
        for (uint32_t k = 0; k < len; ++k) {
          while (ev.valid() && ev.ref < idx + k) ev.next();
          if (!ev.valid() || ev.ref != idx + k || !ev.del) return x;
          const uint32_t nb = nib_of((unsigned char)ev.base());
          if (WRITE) {
            if (k & 1) refb[nf - 1] |= (uint8_t)nb; else refb[nf] = (uint8_t)(nb << 4);
          }
          if (!(k & 1)) ++nf;
        }
        push(len, AVO_OP_DEL);
        idx += len;
        break;
      case 'S':
        if (len) push(len, AVO_OP_SOFTCLIP);
        need += len;
        if (!aligned_seen) didx += len;
        break;
      case 'H':
        if (len) push(len, AVO_OP_HARDCLIP);
        break;
      default:
        return x;   // N, P: "Unsupported CIGAR element"
    }
  }
  x.n_ops = no; x.n_refb = nf; x.need = need; x.ok = true;
  return x;
}

__device__ __forceinline__ bool kept(const DText& T, int64_t i) {
  const uint8_t rf = T.rec_flags[i];
  if (!(rf & 1)) return false;
  if (AVO_T_I64(T, start, i) < 0) return false;
  if (T.prefilter) {                                     // PrefilterReads.scala:142-209, as in avo_pack.cpp
    if (!T.pf.contig_kept) return false;
    if (!(rf & 32) && !T.pf.keep_non_primary) return false;
    if ((rf & 16) && AVO_T_MAPQ(T, i) >= 0 && !(AVO_T_MAPQ(T, i) > T.pf.min_mapq)) return false;
    if ((rf & 64) && !T.pf.keep_duplicates) return false;
  }
  return T.keep ? T.keep[i] != 0 : true;
}

// what both passes need to know about read i
struct ReadText {
  const char *cig, *md, *seq, *qual;
  int ncig, nmd;
  int64_t lseq, lqual;
  uint8_t rf;
};

__device__ __forceinline__ ReadText read_text(const DText& T, int64_t i) {
  ReadText t;
  const int64_t c0 = AVO_T_OFF(T, cigar_off, i), m0 = AVO_T_OFF(T, md_off, i), s0 = AVO_T_OFF(T, seq_off, i),
                q0 = AVO_T_OFF(T, qual_off, i);
  t.cig = T.cigar + c0; t.ncig = (int)(AVO_T_OFF(T, cigar_off, i + 1) - c0);
  t.md = T.md + m0; t.nmd = (int)(AVO_T_OFF(T, md_off, i + 1) - m0);
  t.seq = T.seq + s0; t.lseq = AVO_T_OFF(T, seq_off, i + 1) - s0;
  t.qual = T.qual + q0; t.lqual = AVO_T_OFF(T, qual_off, i + 1) - q0;
  t.rf = T.rec_flags[i];
  return t;
}

template <bool WRITE>
__device__ __forceinline__ Extracted extract_read(const ReadText& t, uint32_t* ops, uint8_t* refb, uint8_t* flags_out) {
  Extracted x = extract_ops<WRITE>(t.cig, t.ncig, t.rf & 4, t.md, t.nmd, t.rf & 8, t.seq, t.lseq, ops, refb);
  // a read whose strings are shorter than the CIGAR's read length is dropped from discovery;
  // more than 65535 operators do not fit the record
  if (x.ok && ((int64_t)x.need > t.lseq || (int64_t)x.need > t.lqual)) x.ok = false;
  if (flags_out)
    *flags_out = (((int64_t)x.need > t.lseq || (int64_t)x.need > t.lqual) ? AVO_READ_SKIP_CALL : 0) |
                 ((x.ok && x.n_ops > 0xFFFFu) ? AVO_READ_OPS_DROPPED : 0);      // flagged, not silent
  if (!x.ok || x.n_ops > 0xFFFFu) { x.n_ops = 0; x.n_refb = 0; }
  return x;
}

__global__ void __launch_bounds__(128) k_pack_count(DText T, uint4* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > T.n) return;
  uint4 c = make_uint4(0, 0, 0, 0);
  if (i < T.n && kept(T, i)) {
    const ReadText t = read_text(T, i);
    const Extracted x = extract_read<false>(t, nullptr, nullptr, nullptr);
    c = make_uint4(1u, (uint32_t)((t.lseq + AVO_BLOCK_BASES - 1) / AVO_BLOCK_BASES), x.n_ops, x.n_refb);
  }
  counts[i] = c;     // counts[n] = 0: the scan's last element is the total
}

__global__ void __launch_bounds__(128) k_pack_records(DText T, const PackOffsets* __restrict__ off, DPacked out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T.n || !kept(T, i)) return;
  const PackOffsets o = off[i];
  const ReadText t = read_text(T, i);
  // the counting pass decided whether the read keeps its operators
  const bool keep_ops = off[i + 1].o > o.o || off[i + 1].f > o.f;
  uint8_t skip = 0;
  Extracted x{0, 0, 0, false};
  if (keep_ops) x = extract_read<true>(t, out.ops + o.o, out.refb + o.f, &skip);
  else x = extract_read<false>(t, nullptr, nullptr, &skip);
  const bool has_mapq = (t.rf & 16) && AVO_T_MAPQ(T, i) >= 0;
  const int mq = has_mapq ? AVO_T_MAPQ(T, i) : 0;
  uint32_t qhead = 0;
  for (int k = 0; k < 4 && k < t.lseq; ++k) {
    int q = (k < t.lqual) ? (int)(unsigned char)t.qual[k] - 33 : 0;
    q = q < 0 ? 0 : q;
    qhead |= (uint32_t)q << (8 * k);
  }
  const uint8_t flags = (uint8_t)(((t.rf & 2) ? AVO_READ_NEGATIVE_STRAND : 0) | skip | (has_mapq ? 0 : AVO_READ_SKIP_CALL));
  // the record as two 16-byte stores: (start, end, seq_off, op_off | refb_off, n_ops/mapq/flags, qhead, l_seq)
  uint4* rec = reinterpret_cast<uint4*>(out.meta + o.k);
  rec[0] = make_uint4((uint32_t)(int32_t)AVO_T_I64(T, start, i), (uint32_t)(int32_t)AVO_T_I64(T, end, i), (uint32_t)o.b, (uint32_t)o.o);
  rec[1] = make_uint4((uint32_t)o.f, (x.n_ops & 0xFFFFu) | ((uint32_t)(mq > 255 ? 255 : mq) << 16) | ((uint32_t)flags << 24),
                      qhead, (uint32_t)t.lseq);
  if (out.source_index) out.source_index[o.k] = i;
}

// 16 lanes per read, one lane per payload block: bytes 0..4 hold 10 bases (high nibble first),
// bytes 5..14 their qualities, byte 15 is zero.
__global__ void __launch_bounds__(256) k_pack_payload(DText T, const PackOffsets* __restrict__ off, DPacked out) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int sub = threadIdx.x & 15;
  if (i >= T.n || !kept(T, i)) return;
  const int64_t s0 = AVO_T_OFF(T, seq_off, i), q0 = AVO_T_OFF(T, qual_off, i);
  const int64_t lseq = AVO_T_OFF(T, seq_off, i + 1) - s0, lqual = AVO_T_OFF(T, qual_off, i + 1) - q0;
  const unsigned char* seq = reinterpret_cast<const unsigned char*>(T.seq) + s0;
  const unsigned char* qual = reinterpret_cast<const unsigned char*>(T.qual) + q0;
  const int64_t lpad = (lseq + AVO_BLOCK_BASES - 1) / AVO_BLOCK_BASES;
  uint4* dst = reinterpret_cast<uint4*>(out.payload) + off[i].b;
  for (int64_t c = sub; c < lpad; c += 16) {
    const int64_t k0 = c * AVO_BLOCK_BASES;
    uint8_t b[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) b[j] = 0;
#pragma unroll
    for (int w = 0; w < AVO_BLOCK_BASES; ++w) {
      const int64_t k = k0 + w;
      if (k < lseq) {
        const uint32_t nb = nib_of(__ldg(seq + k));
        b[w >> 1] |= (uint8_t)((w & 1) ? nb : nb << 4);
        int q = (k < lqual) ? (int)__ldg(qual + k) - 33 : 0;
        b[5 + w] = (uint8_t)(q < 0 ? 0 : q);
      }
    }
    uint4 v;
    v.x = b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24);
    v.y = b[4] | (b[5] << 8) | (b[6] << 16) | ((uint32_t)b[7] << 24);
    v.z = b[8] | (b[9] << 8) | (b[10] << 16) | ((uint32_t)b[11] << 24);
    v.w = b[12] | (b[13] << 8) | (b[14] << 16) | ((uint32_t)b[15] << 24);
    dst[c] = v;
  }
}

struct CountToOffsets {
  __host__ __device__ PackOffsets operator()(const uint4& c) const { return PackOffsets{c.x, c.y, c.z, c.w}; }
};
struct AddOffsets {
  __host__ __device__ PackOffsets operator()(const PackOffsets& a, const PackOffsets& b) const {
    return PackOffsets{a.k + b.k, a.b + b.b, a.o + b.o, a.f + b.f};
  }
};

}  // namespace

size_t pack_scan_temp_bytes(int64_t n) {
  size_t tb = 0;
  cub::TransformInputIterator<PackOffsets, CountToOffsets, const uint4*> it(nullptr, CountToOffsets());
  cub::DeviceScan::ExclusiveScan(nullptr, tb, it, (PackOffsets*)nullptr, AddOffsets(), PackOffsets{0, 0, 0, 0},
                                 (int)(n + 1));
  return tb;
}

cudaError_t launch_pack_count(const DText& T, uint4* counts, PackOffsets* offsets, void* temp, size_t temp_bytes,
                              cudaStream_t s, int* n_launches) {
  const int grid = (int)((T.n + 1 + 127) / 128);
  k_pack_count<<<grid, 128, 0, s>>>(T, counts);
  cub::TransformInputIterator<PackOffsets, CountToOffsets, const uint4*> it(counts, CountToOffsets());
  cudaError_t e = cub::DeviceScan::ExclusiveScan(temp, temp_bytes, it, offsets, AddOffsets(), PackOffsets{0, 0, 0, 0},
                                                 (int)(T.n + 1), s);
  if (n_launches) *n_launches += 2;
  return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_pack_write(const DText& T, const PackOffsets* offsets, const DPacked& out, cudaStream_t s,
                              int* n_launches) {
  if (T.n == 0) return cudaSuccess;
  k_pack_records<<<(int)((T.n + 127) / 128), 128, 0, s>>>(T, offsets, out);
  k_pack_payload<<<(int)((T.n * 16 + 255) / 256), 256, 0, s>>>(T, offsets, out);
  if (n_launches) *n_launches += 2;
  return cudaGetLastError();
}

}  // namespace avo
