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
#include <vector>

#include "../../include/avocado_b200.h"

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

bool is_md_base(char c) { return strchr("ACGTNUKMRSWBVHDX", c) != nullptr; }

// MD tag -> per-reference-offset maps (ADAM MdTag semantics: see oracle/avocado_oracle.cpp)
bool parse_md(const char* s, size_t n, std::vector<char>& mism, std::vector<char>& del) {
  mism.clear(); del.clear();
  if (n == 0) return true;
  size_t off = 0;
  size_t ref = 0;
  auto grow = [&](size_t upto) { if (mism.size() < upto) { mism.resize(upto, 0); del.resize(upto, 0); } };
  auto read_matches = [&]() -> bool {
    size_t j = off;
    uint64_t v = 0;
    while (j < n && isdigit((unsigned char)s[j])) { v = v * 10 + (uint64_t)(s[j] - '0'); if (v > 0x7FFFFFFFu) return false; ++j; }
    if (j == off) return false;
    off = j;
    ref += (size_t)v;
    return true;
  };
  if (!read_matches()) return false;
  while (off < n) {
    bool is_del = false;
    if (s[off] == '^') { is_del = true; ++off; }
    size_t j = off;
    while (j < n && is_md_base((char)toupper((unsigned char)s[j]))) ++j;
    if (j == off) return false;
    grow(ref + (j - off));
    for (size_t k = off; k < j; ++k) {
      const char c = (char)toupper((unsigned char)s[k]);
      if (is_del) del[ref] = c; else mism[ref] = c;
      ++ref;
    }
    off = j;
    if (!read_matches()) return false;
  }
  return true;
}

struct Scratch {
  std::vector<Elem> cigar;
  std::vector<char> mism, del;
  std::vector<uint32_t> ops;
  std::vector<uint8_t> refb;
};

inline char lookup(const std::vector<char>& m, size_t i) { return i < m.size() ? m[i] : 0; }

// returns false when the reference's extractAlignmentOperators would throw or return empty
// refb bytes: MISMATCH -> one (reference base << 4) | read base byte per position, DEL -> the deleted
// reference bases as nibbles, high first, padded to whole bytes.
// `seq`/`lseq` give the read bases for the pairs (0 = N when the sequence is too short; such reads
// are dropped by the caller anyway).
bool extract_ops(const char* cig, size_t ncig, bool has_cigar, const char* md, size_t nmd, bool has_md,
                 const char* seq, int64_t lseq, Scratch& sc, uint32_t* read_len_needed) {
  sc.ops.clear(); sc.refb.clear();
  *read_len_needed = 0;
  if (!has_cigar || (ncig == 1 && cig[0] == '*') || !has_md) return false;
  if (!parse_cigar(cig, ncig, sc.cigar)) return false;
  if (!parse_md(md, nmd, sc.mism, sc.del)) return false;
  size_t idx = 0;
  uint32_t need = 0;   // read bases consumed so far == index of the next read base
  auto push = [&](uint32_t len, uint32_t code) { sc.ops.push_back((len << 4) | code); };
  auto read_nib = [&](uint32_t ri) -> uint8_t {
    return (int64_t)ri < lseq ? NIB.t[(unsigned char)seq[ri]] : (uint8_t)15;
  };
  for (const Elem& e : sc.cigar) {
    if (e.len == 0 && e.op != 'S' && e.op != 'H') return false;  // 0-length M asserts; 0I divides by zero
    switch (e.op) {
      case 'M': {
        uint32_t run = 0;
        bool in_mis = false;
        for (uint32_t k = 0; k < e.len; ++k) {
          const char b = lookup(sc.mism, idx + k);
          const bool mis = b != 0;
          if (k > 0 && mis != in_mis) { push(run, in_mis ? AVO_OP_MISMATCH : AVO_OP_MATCH); run = 0; }
          in_mis = mis;
          ++run;
          if (mis) sc.refb.push_back((uint8_t)((NIB.t[(unsigned char)b] << 4) | read_nib(need + k)));
        }
        push(run, in_mis ? AVO_OP_MISMATCH : AVO_OP_MATCH);
        idx += e.len; need += e.len;
        break;
      }
      case '=':
        push(e.len, AVO_OP_MATCH);
        idx += e.len; need += e.len;
        break;
      case 'X':
        for (uint32_t k = 0; k < e.len; ++k) {
          const char b = lookup(sc.mism, idx + k);
          if (!b) return false;
          sc.refb.push_back((uint8_t)((NIB.t[(unsigned char)b] << 4) | read_nib(need + k)));
        }
        push(e.len, AVO_OP_MISMATCH);
        idx += e.len; need += e.len;
        break;
      case 'I':
        push(e.len, AVO_OP_INS);
        need += e.len;
        break;
      case 'D':
        for (uint32_t k = 0; k < e.len; ++k) {
          const char b = lookup(sc.del, idx + k);
          if (!b) return false;
          const uint8_t nb = NIB.t[(unsigned char)b];
          if (k & 1) sc.refb.back() |= nb; else sc.refb.push_back((uint8_t)(nb << 4));
        }
        push(e.len, AVO_OP_DEL);
        idx += e.len;
        break;
      case 'S':
        if (e.len) push(e.len, AVO_OP_SOFTCLIP);
        need += e.len;
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
  if (!(in->rec_flags[i] & 1)) return false;             // unmapped reads never reach the path
  if (in->start[i] < 0) return false;
  return in->keep ? in->keep[i] != 0 : true;
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

int avo_pack_bounds(const avo_text_reads* in, int64_t* n_reads, int64_t* n_blocks, int64_t* n_ops,
                    int64_t* n_refb) {
  if (!in || !n_reads || !n_blocks || !n_ops || !n_refb) return AVO_E_INVALID;
  int64_t r = 0, b = 0, o = 0, f = 0;
  for (int64_t i = 0; i < in->n; ++i) {
    if (!kept(in, i)) continue;
    ++r;
    b += ((in->seq_off[i + 1] - in->seq_off[i]) + AVO_BLOCK_BASES - 1) / AVO_BLOCK_BASES;   // payload blocks
    // every op needs at least one CIGAR character pair or one MD character
    const int64_t c = in->cigar_off[i + 1] - in->cigar_off[i];
    const int64_t m = in->md_off[i + 1] - in->md_off[i];
    o += c / 2 + 1 + 2 * m;
    f += m + 1;
  }
  *n_reads = r; *n_blocks = b; *n_ops = o; *n_refb = f;
  return AVO_OK;
}

int avo_pack_reads(const avo_text_reads* in, avo_packed_out* out) {
  if (!in || !out) return AVO_E_INVALID;
  Scratch sc;
  int64_t r = 0;
  uint64_t nb = 0, no = 0, nf = 0;
  for (int64_t i = 0; i < in->n; ++i) {
    if (!kept(in, i)) continue;
    if (r >= out->n_reads) return AVO_E_CAPACITY;
    const uint8_t rf = in->rec_flags[i];
    const char* seq = in->seq + in->seq_off[i];
    const int64_t lseq = in->seq_off[i + 1] - in->seq_off[i];
    const char* qual = in->qual + in->qual_off[i];
    const int64_t lqual = in->qual_off[i + 1] - in->qual_off[i];
    uint32_t need = 0;
    bool ok = extract_ops(in->cigar + in->cigar_off[i], (size_t)(in->cigar_off[i + 1] - in->cigar_off[i]),
                          rf & 4, in->md + in->md_off[i], (size_t)(in->md_off[i + 1] - in->md_off[i]),
                          rf & 8, seq, lseq, sc, &need);
    uint8_t flags = (rf & 2) ? AVO_READ_NEGATIVE_STRAND : 0;
    const bool has_mapq = (rf & 16) && in->mapq[i] >= 0;
    // observeRead indexes sequence and qualities up to the CIGAR's read length
    if (!has_mapq || (int64_t)need > lseq || (int64_t)need > lqual) flags |= AVO_READ_SKIP_CALL;
    // discovery additionally indexes qual(i) from 0 and sequence up to the same length: a read
    // whose strings are too short would crash the reference job; it is dropped here instead
    if (ok && ((int64_t)need > lseq || (int64_t)need > lqual)) ok = false;
    if (!ok || sc.ops.size() > 0xFFFFu) { sc.ops.clear(); sc.refb.clear(); }
    const uint64_t lpad = ((uint64_t)lseq + AVO_BLOCK_BASES - 1) / AVO_BLOCK_BASES;   // payload blocks of the read
    if ((int64_t)(nb + lpad) > out->n_blocks || (int64_t)(no + sc.ops.size()) > out->n_ops ||
        (int64_t)(nf + sc.refb.size()) > out->n_refb)
      return AVO_E_CAPACITY;
    if (nb + lpad > 0xFFFFFFFFull) return AVO_E_INVALID;   // split the batch
    avo_read_meta& m = out->meta[r];
    m.start = (int32_t)in->start[i];
    m.end = (int32_t)in->end[i];
    const int mq = has_mapq ? in->mapq[i] : 0;
    m.mapq = (uint8_t)(mq > 255 ? 255 : mq);
    m.flags = flags;
    m.seq_off = (uint32_t)nb;
    m.op_off = (uint32_t)no;
    m.refb_off = (uint32_t)nf;
    m.n_ops = (uint16_t)sc.ops.size();
    m.l_seq = (uint32_t)lseq;
    m.qhead = 0;
    if (out->source_index) out->source_index[r] = i;
    memset(out->payload + (nb << 4), 0, (size_t)lpad << 4);
    for (int64_t k = 0; k < lseq; ++k) {
      const uint64_t c = (uint64_t)k / AVO_BLOCK_BASES, w = (uint64_t)k - c * AVO_BLOCK_BASES;
      uint8_t* blk = out->payload + ((nb + c) << 4);
      const uint8_t nibv = NIB.t[(unsigned char)seq[k]];
      blk[w >> 1] |= (w & 1) ? nibv : (uint8_t)(nibv << 4);
      int q = (k < lqual) ? (int)(unsigned char)qual[k] - 33 : 0;
      q = q < 0 ? 0 : (q > 255 ? 255 : q);
      blk[5 + w] = (uint8_t)q;
      if (k < 4) m.qhead |= (uint32_t)q << (8 * k);
    }
    for (size_t k = 0; k < sc.ops.size(); ++k) out->ops[no + k] = sc.ops[k];
    for (size_t k = 0; k < sc.refb.size(); ++k) out->refb[nf + k] = sc.refb[k];
    nb += lpad; no += sc.ops.size(); nf += sc.refb.size();
    ++r;
  }
  out->n_reads = r; out->n_blocks = (int64_t)nb; out->n_ops = (int64_t)no; out->n_refb = (int64_t)nf;
  return AVO_OK;
}

}  // extern "C"
