// avo_pack.cpp -- host-side packer: AlignmentRecord text columns -> the columnar batch of
// include/avocado_b200.h.  This is where ObservationOperator.extractAlignmentOperators
// (avocado-core/src/main/scala/org/bdgenomics/avocado/models/ObservationOperator.scala:42-171)
// happens in this build: the CIGAR is merged with the MD tag once, on the host, into =/X/I/D/S/H
// runs, and the kernels never see text.
//
// Failure behaviour mirrors the reference: a read whose operators cannot be extracted (CIGAR
// null or "*", MD null, unsupported CIGAR op, X/D without an MD base, malformed MD) contributes
// nothing to discovery (DiscoverVariants.scala:119-127) and nothing to calling
// (BiallelicGenotyper.scala:385-391): it is packed with zero operators.  A read that the
// reference's observeRead would throw on later (null mapq, sequence or qualities shorter than the
// CIGAR) keeps its operators for discovery and gets AVO_READ_SKIP_CALL.
#include <cctype>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/avocado_b200.h"

// numeric text columns in either width (avo_text_reads.narrow)
static inline int64_t t_i64(const avo_text_reads* in, const int64_t* col, int64_t i) {
  return in->narrow ? (int64_t)reinterpret_cast<const int32_t*>(col)[i] : col[i];
}
static inline int64_t t_off(const avo_text_reads* in, const int64_t* col, int64_t i) {
  return in->narrow ? (int64_t)reinterpret_cast<const uint32_t*>(col)[i] : col[i];
}
static inline int32_t t_mapq(const avo_text_reads* in, int64_t i) {
  return in->narrow ? (int32_t)reinterpret_cast<const uint8_t*>(in->mapq)[i] : in->mapq[i];
}

namespace {

// BAM 4-bit base codes "=ACMGRSVTWYHKDBN"
struct NibTable {
  uint8_t t[256];
  NibTable() {
    memset(t, 15, sizeof t);
    const char* codes = "=ACMGRSVTWYHKDBN";
    for (int i = 0; i < 16; ++i) {
      t[(unsigned char)codes[i]] = (uint8_t)i;
      t[(unsigned char)tolower(codes[i])] = (uint8_t)i;
    }
  }
};
const NibTable NIB;

struct Elem { uint32_t len; char op; };

bool parse_cigar(const char* s, size_t n, std::vector<Elem>& out) {
  out.clear();
  size_t i = 0;
  while (i < n) {
    if (!isdigit((unsigned char)s[i])) return false;
    uint64_t v = 0;
    while (i < n && isdigit((unsigned char)s[i])) { v = v * 10 + (uint64_t)(s[i] - '0'); if (v > 0x0FFFFFFFu) return false; ++i; }
    if (i >= n) return false;
    const char op = s[i++];
    if (!strchr("MIDNSHP=X", op)) return false;
    out.push_back({(uint32_t)v, op});
  }
  return true;
}

struct MdBase {
  bool ok[256];
  MdBase() {
    memset(ok, 0, sizeof ok);
    for (const char* c = "ACGTNUKMRSWBVHDX"; *c; ++c) { ok[(unsigned char)*c] = true; ok[(unsigned char)tolower(*c)] = true; }
  }
};
const MdBase MD_BASE;

// two sequence characters -> one payload byte (high nibble first)
struct PairTable {
  uint8_t t[65536];
  PairTable() {
    for (int a = 0; a < 256; ++a)
      for (int b = 0; b < 256; ++b) t[a | (b << 8)] = (uint8_t)((NIB.t[a] << 4) | NIB.t[b]);
  }
};
const PairTable PAIR;

// One MD event: the reference base recorded for alignment offset `off` (ADAM MdTag semantics: see
// oracle/avocado_oracle.cpp): a mismatch, or a deleted base.
struct MdEvent { uint64_t off; char base; bool del; };

// MD tag -> events in offset order.  false: malformed (the reference's MdTag would throw)
bool parse_md(const char* s, size_t n, std::vector<MdEvent>& ev) {
  ev.clear();
  if (n == 0) return true;
  size_t off = 0;
  uint64_t ref = 0;
  auto read_matches = [&]() -> bool {
    size_t j = off;
    uint64_t v = 0;
    while (j < n && (unsigned)(s[j] - '0') <= 9u) { v = v * 10 + (uint64_t)(s[j] - '0'); if (v > 0x7FFFFFFFu) return false; ++j; }
    if (j == off) return false;
    off = j;
    ref += v;
    return true;
  };
  if (!read_matches()) return false;
  while (off < n) {
    bool is_del = false;
    if (s[off] == '^') { is_del = true; ++off; }
    size_t j = off;
    while (j < n && MD_BASE.ok[(unsigned char)s[j]]) ++j;
    if (j == off) return false;
    for (size_t k = off; k < j; ++k) ev.push_back({ref++, (char)toupper((unsigned char)s[k]), is_del});
    off = j;
    if (!read_matches()) return false;
  }
  return true;
}

struct Scratch {
  std::vector<Elem> cigar;
  std::vector<MdEvent> md;
  std::vector<uint32_t> ops;
  std::vector<uint8_t> refb;
};

// returns false when the reference's extractAlignmentOperators would throw or return empty
// refb bytes: MISMATCH -> one (reference base << 4) | read base byte per position, DEL -> the deleted
// reference bases as nibbles, high first, padded to whole bytes.
// `seq`/`lseq` give the read bases for the pairs (0 = N when the sequence is too short; such reads
// are dropped by the caller anyway).
// The CIGAR is walked with a cursor into the MD events: an M element takes the mismatches recorded
// inside its span (deleted bases recorded there are ignored, as a lookup in MdTag's mismatch map
// would), X needs a mismatch and D a deleted base at every one of its offsets.
bool extract_ops(const char* cig, size_t ncig, bool has_cigar, const char* md, size_t nmd, bool has_md,
                 const char* seq, int64_t lseq, Scratch& sc, uint32_t* read_len_needed) {
  sc.ops.clear(); sc.refb.clear();
  *read_len_needed = 0;
  if (!has_cigar || (ncig == 1 && cig[0] == '*') || !has_md) return false;
  if (!parse_cigar(cig, ncig, sc.cigar)) return false;
  if (!parse_md(md, nmd, sc.md)) return false;
  const MdEvent* ev = sc.md.data();
  const size_t nev = sc.md.size();
  size_t cur = 0;       // first event at or after the current alignment offset
  uint64_t idx = 0;
  uint32_t need = 0;    // read bases consumed so far == index of the next read base
  // The read base stored next to a mismatch's reference base is discovery's alternate allele, and
  // discovery indexes the read its own way (DiscoverVariants.scala:139-172, 189-193): soft clips
  // count before the first aligned operator and are ignored after it.
  uint64_t didx = 0;
  bool aligned_seen = false;
  auto push = [&](uint32_t len, uint32_t code) { sc.ops.push_back((len << 4) | code); };
  auto read_nib = [&](uint64_t ri) -> uint8_t {
    return (int64_t)ri < lseq ? NIB.t[(unsigned char)seq[ri]] : (uint8_t)15;
  };
  // the event recorded for offset `o`, or NULL
  auto at = [&](uint64_t o) -> const MdEvent* {
    while (cur < nev && ev[cur].off < o) ++cur;
    return (cur < nev && ev[cur].off == o) ? &ev[cur] : nullptr;
  };
  for (const Elem& e : sc.cigar) {
    if (e.len == 0 && e.op != 'S' && e.op != 'H') return false;  // 0-length M asserts; 0I divides by zero
    switch (e.op) {
      case 'M': {
        const uint64_t hi = idx + e.len;
        uint64_t pos = idx;                                  // next offset not yet covered by an operator
        while (cur < nev && ev[cur].off < idx) ++cur;
        while (cur < nev && ev[cur].off < hi) {
          if (ev[cur].del) { ++cur; continue; }
          const uint64_t k = ev[cur].off;
          if (k > pos) push((uint32_t)(k - pos), AVO_OP_MATCH);
          uint32_t run = 0;
          while (cur < nev && !ev[cur].del && ev[cur].off == k + run && k + run < hi) {
            sc.refb.push_back((uint8_t)((NIB.t[(unsigned char)ev[cur].base] << 4) | read_nib(didx + (k - idx) + run)));
            ++run; ++cur;
          }
          push(run, AVO_OP_MISMATCH);
          pos = k + run;
        }
        if (pos < hi) push((uint32_t)(hi - pos), AVO_OP_MATCH);
        idx += e.len; need += e.len; didx += e.len; aligned_seen = true;
        break;
      }
      case '=':
        push(e.len, AVO_OP_MATCH);
        idx += e.len; need += e.len; didx += e.len; aligned_seen = true;
        break;
      case 'X':
        for (uint32_t k = 0; k < e.len; ++k) {
          const MdEvent* m = at(idx + k);
          if (!m || m->del) return false;
          sc.refb.push_back((uint8_t)((NIB.t[(unsigned char)m->base] << 4) | read_nib(didx + k)));
        }
        push(e.len, AVO_OP_MISMATCH);
        idx += e.len; need += e.len; didx += e.len; aligned_seen = true;
        break;
      case 'I':
        push(e.len, AVO_OP_INS);
        need += e.len; didx += e.len;
        break;
      case 'D':
        for (uint32_t k = 0; k < e.len; ++k) {
          const MdEvent* m = at(idx + k);
          if (!m || !m->del) return false;
          const uint8_t nb = NIB.t[(unsigned char)m->base];
          if (k & 1) sc.refb.back() |= nb; else sc.refb.push_back((uint8_t)(nb << 4));
        }
        push(e.len, AVO_OP_DEL);
        idx += e.len;
        break;
      case 'S':
        if (e.len) push(e.len, AVO_OP_SOFTCLIP);
        need += e.len;
        if (!aligned_seen) didx += e.len;
        break;
      case 'H':
        if (e.len) push(e.len, AVO_OP_HARDCLIP);
        break;
      default:
        return false;  // N, P: "Unsupported CIGAR element"
    }
  }
  *read_len_needed = need;
  return true;
}

bool kept(const avo_text_reads* in, int64_t i) {
  const uint8_t rf = in->rec_flags[i];
  if (!(rf & 1)) return false;                           // unmapped reads never reach the path
  if (t_i64(in, in->start, i) < 0) return false;
  if (const avo_prefilter* f = in->prefilter) {          // PrefilterReads.scala:142-209
    if (!f->contig_kept) return false;
    if (!(rf & 32) && !f->keep_non_primary) return false;                 // filterMapped
    if ((rf & 16) && t_mapq(in, i) >= 0 && !(t_mapq(in, i) > f->min_mapq)) return false;   // filterMappingQuality
    if ((rf & 64) && !f->keep_duplicates) return false;                   // filterUnique
  }
  return in->keep ? in->keep[i] != 0 : true;
}

// Packing is split over contiguous chunks of the input, one per thread: payload blocks and records
// go straight to their final place (their offsets only need the sequence lengths), operators and
// mismatch bytes are collected per chunk and copied once the chunk totals are known.
struct Chunk {
  int64_t i0 = 0, i1 = 0;        // input reads
  int64_t r0 = 0, n_kept = 0;    // packed reads
  uint64_t blk0 = 0, n_blk = 0;  // payload blocks
  uint64_t op0 = 0, fb0 = 0;     // final offsets of the chunk's operators / mismatch bytes
  std::vector<uint32_t> ops;
  std::vector<uint8_t> refb;
  int rc = AVO_OK;
};

void pack_chunk(const avo_text_reads* in, avo_packed_out* out, Chunk& ch) {
  Scratch sc;
  int64_t r = ch.r0;
  uint64_t nb = ch.blk0;
  ch.ops.reserve((size_t)ch.n_kept * 4);
  ch.refb.reserve((size_t)ch.n_kept * 2);
  for (int64_t i = ch.i0; i < ch.i1; ++i) {
    if (!kept(in, i)) continue;
    const uint8_t rf = in->rec_flags[i];
    const char* seq = in->seq + t_off(in, in->seq_off, i);
    const int64_t lseq = t_off(in, in->seq_off, i + 1) - t_off(in, in->seq_off, i);
    const char* qual = in->qual + t_off(in, in->qual_off, i);
    const int64_t lqual = t_off(in, in->qual_off, i + 1) - t_off(in, in->qual_off, i);
    uint32_t need = 0;
    bool ok = extract_ops(in->cigar + t_off(in, in->cigar_off, i), (size_t)(t_off(in, in->cigar_off, i + 1) - t_off(in, in->cigar_off, i)),
                          rf & 4, in->md + t_off(in, in->md_off, i), (size_t)(t_off(in, in->md_off, i + 1) - t_off(in, in->md_off, i)),
                          rf & 8, seq, lseq, sc, &need);
    uint8_t flags = (rf & 2) ? AVO_READ_NEGATIVE_STRAND : 0;
    const bool has_mapq = (rf & 16) && t_mapq(in, i) >= 0;
    // observeRead indexes sequence and qualities up to the CIGAR's read length
    if (!has_mapq || (int64_t)need > lseq || (int64_t)need > lqual) flags |= AVO_READ_SKIP_CALL;
    // discovery additionally indexes qual(i) from 0 and sequence up to the same length: a read
    // whose strings are too short would crash the reference job; it is dropped here instead
    if (ok && ((int64_t)need > lseq || (int64_t)need > lqual)) ok = false;
    if (ok && sc.ops.size() > 0xFFFFu) flags |= AVO_READ_OPS_DROPPED;   // does not fit the record: flagged, not silent
    if (!ok || sc.ops.size() > 0xFFFFu) { sc.ops.clear(); sc.refb.clear(); }
    const uint64_t lpad = ((uint64_t)lseq + AVO_BLOCK_BASES - 1) / AVO_BLOCK_BASES;   // payload blocks of the read
    if (ch.ops.size() + sc.ops.size() > 0xFFFFFFFFull || ch.refb.size() + sc.refb.size() > 0xFFFFFFFFull) {
      ch.rc = AVO_E_INVALID;   // split the batch
      return;
    }
    avo_read_meta& m = out->meta[r];
    m.start = (int32_t)t_i64(in, in->start, i);
    m.end = (int32_t)t_i64(in, in->end, i);
    const int mq = has_mapq ? t_mapq(in, i) : 0;
    m.mapq = (uint8_t)(mq > 255 ? 255 : mq);
    m.flags = flags;
    m.seq_off = (uint32_t)nb;
    m.op_off = (uint32_t)ch.ops.size();        // chunk-relative until the totals are known
    m.refb_off = (uint32_t)ch.refb.size();
    m.n_ops = (uint16_t)sc.ops.size();
    m.l_seq = (uint32_t)lseq;
    if (out->source_index) out->source_index[r] = i;
    // payload: 10 bases in bytes 0..4 (high nibble first), their qualities in bytes 5..14, byte 15 = 0
    uint8_t* blk = out->payload + (nb << 4);
    const int64_t full = (lseq < lqual ? lseq : lqual) / AVO_BLOCK_BASES;   // blocks with 10 bases and 10 qualities
    for (int64_t c = 0; c < full; ++c, blk += AVO_BLOCK_BYTES) {
      const unsigned char* sp = (const unsigned char*)seq + c * AVO_BLOCK_BASES;
      const unsigned char* qp = (const unsigned char*)qual + c * AVO_BLOCK_BASES;
      for (int j = 0; j < 5; ++j) blk[j] = PAIR.t[sp[2 * j] | (sp[2 * j + 1] << 8)];
      for (int w = 0; w < AVO_BLOCK_BASES; ++w) blk[5 + w] = (uint8_t)(qp[w] < 33 ? 0 : qp[w] - 33);
      blk[15] = 0;
    }
    if ((uint64_t)full < lpad) {
      memset(blk, 0, (size_t)(lpad - (uint64_t)full) << 4);
      for (int64_t k = full * AVO_BLOCK_BASES; k < lseq; ++k) {
        const uint64_t c = (uint64_t)k / AVO_BLOCK_BASES, w = (uint64_t)k - c * AVO_BLOCK_BASES;
        uint8_t* b = out->payload + ((nb + c) << 4);
        const uint8_t nibv = NIB.t[(unsigned char)seq[k]];
        b[w >> 1] |= (w & 1) ? nibv : (uint8_t)(nibv << 4);
        int q = (k < lqual) ? (int)(unsigned char)qual[k] - 33 : 0;
        b[5 + w] = (uint8_t)(q < 0 ? 0 : q);
      }
    }
    m.qhead = 0;
    for (int64_t k = 0; k < 4 && k < lseq; ++k)
      m.qhead |= (uint32_t)out->payload[(nb << 4) + 5 + (uint64_t)k] << (8 * k);
    ch.ops.insert(ch.ops.end(), sc.ops.begin(), sc.ops.end());
    ch.refb.insert(ch.refb.end(), sc.refb.begin(), sc.refb.end());
    nb += lpad;
    ++r;
  }
}

template <typename F>
void run_chunks(std::vector<Chunk>& chunks, F&& f) {
  if (chunks.size() == 1) { f(chunks[0]); return; }
  std::vector<std::thread> pool;
  pool.reserve(chunks.size());
  for (Chunk& ch : chunks) pool.emplace_back([&f, &ch]() { f(ch); });
  for (std::thread& t : pool) t.join();
}

}  // namespace

extern "C" {

int avo_pack_wire(const avo_read_batch* full, avo_read_wire* wire_meta, uint16_t* wire_ops) {
  if (!full || full->mem != AVO_MEM_HOST || (full->n_reads > 0 && (!full->meta || !wire_meta)) ||
      (full->n_ops > 0 && (!full->ops || !wire_ops)))
    return AVO_E_INVALID;
  // the compact form drops the offsets: the columns must be laid out read after read without gaps
  uint64_t blk = 0, op = 0, fb = 0;
  for (int64_t r = 0; r < full->n_reads; ++r) {
    const avo_read_meta& m = full->meta[r];
    const uint64_t nfb = (r + 1 < full->n_reads ? (uint64_t)full->meta[r + 1].refb_off : (uint64_t)full->n_refb) - m.refb_off;
    const int64_t span = (int64_t)m.end - (int64_t)m.start;
    if (m.seq_off != blk || m.op_off != op || m.refb_off != fb) return AVO_E_UNSUPPORTED;
    if (span < 0 || span > 0xFFFF || m.l_seq > 0xFFFFu || m.n_ops > 0xFFu || nfb > 0xFFu) return AVO_E_UNSUPPORTED;
    avo_read_wire& w = wire_meta[r];
    w.start = m.start; w.span = (uint16_t)span; w.l_seq = (uint16_t)m.l_seq; w.n_ops = (uint8_t)m.n_ops;
    w.n_refb = (uint8_t)nfb; w.mapq = m.mapq; w.flags = m.flags; w.qhead = m.qhead;
    blk += (m.l_seq + AVO_BLOCK_BASES - 1) / AVO_BLOCK_BASES; op += m.n_ops; fb += nfb;
  }
  if ((int64_t)blk != full->n_blocks || (int64_t)op != full->n_ops || (int64_t)fb != full->n_refb)
    return AVO_E_UNSUPPORTED;
  for (int64_t i = 0; i < full->n_ops; ++i) {
    if (full->ops[i] > 0xFFFFu) return AVO_E_UNSUPPORTED;     // operator of 4096 bases or more
    wire_ops[i] = (uint16_t)full->ops[i];
  }
  return AVO_OK;
}

int avo_pack_wire6(const avo_read_batch* full, avo_read_wire6* wire6_meta, uint8_t* wire6_oplen,
                   uint8_t* wire6_opcode, avo_wire_exception* exc, int64_t exc_capacity, int64_t* n_exc,
                   int32_t* base) {
  if (!full || full->mem != AVO_MEM_HOST || !n_exc || !base || (full->n_reads > 0 && (!full->meta || !wire6_meta)) ||
      (full->n_ops > 0 && (!full->ops || !wire6_oplen || !wire6_opcode)) || (exc_capacity > 0 && !exc) ||
      exc_capacity < 0)
    return AVO_E_INVALID;
  if (full->n_ops > 0) memset(wire6_opcode, 0, (size_t)(full->n_ops + 1) / 2);
  uint64_t blk = 0, op = 0, fb = 0;
  int64_t ne = 0;
  int32_t prev = full->n_reads ? full->meta[0].start : 0;
  *base = prev;
  for (int64_t r = 0; r < full->n_reads; ++r) {
    const avo_read_meta& m = full->meta[r];
    const uint64_t nfb = (r + 1 < full->n_reads ? (uint64_t)full->meta[r + 1].refb_off : (uint64_t)full->n_refb) - m.refb_off;
    if (m.seq_off != blk || m.op_off != op || m.refb_off != fb) return AVO_E_UNSUPPORTED;
    const int64_t d = (int64_t)m.start - (int64_t)prev;
    if (d < 0 || d > 0x3FFF || m.n_ops > 0xFFu || m.flags > 3u) return AVO_E_UNSUPPORTED;
    // what the operators imply
    uint64_t span = 0, lseq = 0, nrefb = 0;
    for (uint32_t k = 0; k < m.n_ops; ++k) {
      const uint32_t o = full->ops[m.op_off + k], len = o >> 4, code = o & 15u;
      if (len > 0xFFu) return AVO_E_UNSUPPORTED;      // operator of 256 bases or more
      if (code == AVO_OP_MATCH || code == AVO_OP_MISMATCH || code == AVO_OP_DEL) span += len;
      if (code == AVO_OP_MATCH || code == AVO_OP_MISMATCH || code == AVO_OP_INS || code == AVO_OP_SOFTCLIP) lseq += len;
      if (code == AVO_OP_MISMATCH) nrefb += len;
      if (code == AVO_OP_DEL) nrefb += (len + 1) >> 1;
    }
    if (nrefb != nfb) return AVO_E_UNSUPPORTED;
    if ((int64_t)m.start + (int64_t)span != (int64_t)m.end || lseq != m.l_seq) {
      if (ne >= exc_capacity) return AVO_E_UNSUPPORTED;
      exc[ne].index = (uint32_t)r; exc[ne].l_seq = m.l_seq; exc[ne].end = m.end; exc[ne].reserved = 0;
      ++ne;
    }
    avo_read_wire6& w = wire6_meta[r];
    w.dstart_flags = (uint16_t)(((uint32_t)d << 2) | m.flags);
    w.n_ops = (uint8_t)m.n_ops; w.mapq = m.mapq; w.q0 = (uint8_t)(m.qhead & 255u); w.q1 = (uint8_t)((m.qhead >> 8) & 255u);
    for (uint32_t k = 0; k < m.n_ops; ++k) {
      const uint64_t i = (uint64_t)m.op_off + k;
      const uint32_t o = full->ops[i];
      wire6_oplen[i] = (uint8_t)(o >> 4);
      wire6_opcode[i >> 1] |= (uint8_t)((o & 15u) << ((i & 1) * 4));
    }
    prev = m.start;
    blk += (m.l_seq + AVO_BLOCK_BASES - 1) / AVO_BLOCK_BASES; op += m.n_ops; fb += nfb;
  }
  if ((int64_t)blk != full->n_blocks || (int64_t)op != full->n_ops || (int64_t)fb != full->n_refb)
    return AVO_E_UNSUPPORTED;
  *n_exc = ne;
  return AVO_OK;
}

int avo_pack_bounds(const avo_text_reads* in, int64_t* n_reads, int64_t* n_blocks, int64_t* n_ops,
                    int64_t* n_refb) {
  if (!in || !n_reads || !n_blocks || !n_ops || !n_refb) return AVO_E_INVALID;
  int64_t r = 0, b = 0, o = 0, f = 0;
  for (int64_t i = 0; i < in->n; ++i) {
    if (!kept(in, i)) continue;
    ++r;
    b += ((t_off(in, in->seq_off, i + 1) - t_off(in, in->seq_off, i)) + AVO_BLOCK_BASES - 1) / AVO_BLOCK_BASES;   // payload blocks
    // every op needs at least one CIGAR character pair or one MD character
    const int64_t c = t_off(in, in->cigar_off, i + 1) - t_off(in, in->cigar_off, i);
    const int64_t m = t_off(in, in->md_off, i + 1) - t_off(in, in->md_off, i);
    o += c / 2 + 1 + 2 * m;
    f += m + 1;
  }
  *n_reads = r; *n_blocks = b; *n_ops = o; *n_refb = f;
  return AVO_OK;
}

int avo_pack_reads_mt(const avo_text_reads* in, avo_packed_out* out, int n_threads) {
  if (!in || !out || in->n < 0) return AVO_E_INVALID;
  int64_t T = n_threads < 1 ? 1 : (n_threads > 256 ? 256 : n_threads);
  if (T > (in->n + 4095) / 4096) T = (in->n + 4095) / 4096;     // not worth a thread below 4096 reads
  if (T < 1) T = 1;
  std::vector<Chunk> chunks((size_t)T);
  for (int64_t t = 0; t < T; ++t) { chunks[t].i0 = in->n * t / T; chunks[t].i1 = in->n * (t + 1) / T; }
  // records and payload blocks per chunk: sequence lengths only
  run_chunks(chunks, [in](Chunk& ch) {
    for (int64_t i = ch.i0; i < ch.i1; ++i) {
      if (!kept(in, i)) continue;
      ++ch.n_kept;
      ch.n_blk += (uint64_t)((t_off(in, in->seq_off, i + 1) - t_off(in, in->seq_off, i)) + AVO_BLOCK_BASES - 1) / AVO_BLOCK_BASES;
    }
  });
  int64_t r = 0;
  uint64_t nb = 0;
  for (Chunk& ch : chunks) { ch.r0 = r; ch.blk0 = nb; r += ch.n_kept; nb += ch.n_blk; }
  if (r > out->n_reads || (int64_t)nb > out->n_blocks) return AVO_E_CAPACITY;
  if (nb > 0xFFFFFFFFull) return AVO_E_INVALID;   // split the batch
  run_chunks(chunks, [in, out](Chunk& ch) { pack_chunk(in, out, ch); });
  uint64_t no = 0, nf = 0;
  for (Chunk& ch : chunks) {
    if (ch.rc != AVO_OK) return ch.rc;
    ch.op0 = no; ch.fb0 = nf; no += ch.ops.size(); nf += ch.refb.size();
  }
  if ((int64_t)no > out->n_ops || (int64_t)nf > out->n_refb) return AVO_E_CAPACITY;
  if (no > 0xFFFFFFFFull || nf > 0xFFFFFFFFull) return AVO_E_INVALID;
  run_chunks(chunks, [out](Chunk& ch) {
    if (!ch.ops.empty()) memcpy(out->ops + ch.op0, ch.ops.data(), ch.ops.size() * sizeof(uint32_t));
    if (!ch.refb.empty()) memcpy(out->refb + ch.fb0, ch.refb.data(), ch.refb.size());
    if (ch.op0 || ch.fb0)
      for (int64_t k = ch.r0; k < ch.r0 + ch.n_kept; ++k) {
        out->meta[k].op_off += (uint32_t)ch.op0;
        out->meta[k].refb_off += (uint32_t)ch.fb0;
      }
  });
  out->n_reads = r; out->n_blocks = (int64_t)nb; out->n_ops = (int64_t)no; out->n_refb = (int64_t)nf;
  return AVO_OK;
}

int avo_pack_reads(const avo_text_reads* in, avo_packed_out* out) { return avo_pack_reads_mt(in, out, 1); }

}  // extern "C"
