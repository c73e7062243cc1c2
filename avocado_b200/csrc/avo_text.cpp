// avo_text.cpp -- packed host batch -> AlignmentRecord text columns (CIGAR, MD, sequence, qualities).
// Bench / test infrastructure (include/avocado_b200_synth.h): it gives the packer benchmark and the
// pack(text(batch)) == batch round-trip test text records of the synthetic workload.  =/X runs are
// merged back into CIGAR M and the reference bases of X / D operators go into the MD tag.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/avocado_b200.h"
#include "../../include/avocado_b200_synth.h"

namespace {

const char* const NIB_CHARS = "=ACMGRSVTWYHKDBN";

inline void put_num(std::string& s, uint64_t v) {
  char buf[24];
  const int n = snprintf(buf, sizeof buf, "%llu", (unsigned long long)v);
  s.append(buf, (size_t)n);
}

// one read; returns the read bases the operators consume
void read_text(const avo_read_batch* b, int64_t r, std::string& cigar, std::string& md) {
  const avo_read_meta& m = b->meta[r];
  cigar.clear(); md.clear();
  uint32_t m_run = 0;       // pending CIGAR M length
  uint64_t md_run = 0;      // pending MD match count
  uint32_t rb = m.refb_off;
  bool md_any = false;
  auto flush_m = [&]() { if (m_run) { put_num(cigar, m_run); cigar.push_back('M'); m_run = 0; } };
  for (uint32_t k = 0; k < m.n_ops; ++k) {
    const uint32_t op = b->ops[m.op_off + k], len = op >> 4, code = op & 15u;
    switch (code) {
      case AVO_OP_MATCH: m_run += len; md_run += len; md_any = true; break;
      case AVO_OP_MISMATCH:
        m_run += len;
        for (uint32_t i = 0; i < len; ++i) {
          put_num(md, md_run); md_run = 0;
          md.push_back(NIB_CHARS[b->refb[rb + i] >> 4]);
        }
        rb += len; md_any = true;
        break;
      case AVO_OP_INS: flush_m(); put_num(cigar, len); cigar.push_back('I'); break;
      case AVO_OP_DEL:
        flush_m(); put_num(cigar, len); cigar.push_back('D');
        put_num(md, md_run); md_run = 0;
        md.push_back('^');
        for (uint32_t i = 0; i < len; ++i) {
          const uint8_t byte = b->refb[rb + (i >> 1)];
          md.push_back(NIB_CHARS[(i & 1) ? (byte & 15u) : (byte >> 4)]);
        }
        rb += (len + 1) >> 1; md_any = true;
        break;
      case AVO_OP_SOFTCLIP: flush_m(); put_num(cigar, len); cigar.push_back('S'); break;
      case AVO_OP_HARDCLIP: flush_m(); put_num(cigar, len); cigar.push_back('H'); break;
      default: break;
    }
  }
  flush_m();
  if (md_any || m.n_ops) put_num(md, md_run);
}

}  // namespace

extern "C" int avo_synth_text(const avo_read_batch* b, int64_t first, int64_t n, avo_text_out* out) {
  if (!b || !out || b->mem != AVO_MEM_HOST || first < 0 || n < 0 || first + n > b->n_reads) return AVO_E_INVALID;
  std::string cigar, md;
  int64_t nc = 0, nm = 0, ns = 0;
  const bool write = out->cigar != nullptr;
  for (int64_t j = 0; j < n; ++j) {
    const avo_read_meta& m = b->meta[first + j];
    read_text(b, first + j, cigar, md);
    if (write) {
      if (nc + (int64_t)cigar.size() > out->cigar_bytes || nm + (int64_t)md.size() > out->md_bytes ||
          ns + (int64_t)m.l_seq > out->seq_bytes)
        return AVO_E_CAPACITY;
      out->start[j] = m.start; out->end[j] = m.end; out->mapq[j] = m.mapq;
      out->rec_flags[j] = (uint8_t)(1 | ((m.flags & AVO_READ_NEGATIVE_STRAND) ? 2 : 0) | (m.n_ops ? 4 | 8 : 0) | 16 | 32 /* primary */);
      out->cigar_off[j] = nc; out->md_off[j] = nm; out->seq_off[j] = ns;
      memcpy(out->cigar + nc, cigar.data(), cigar.size());
      memcpy(out->md + nm, md.data(), md.size());
      for (uint32_t i = 0; i < m.l_seq; ++i) {
        const uint8_t* blk = b->payload + (((uint64_t)m.seq_off + i / AVO_BLOCK_BASES) << 4);
        const uint32_t w = i % AVO_BLOCK_BASES;
        const uint8_t byte = blk[w >> 1];
        out->seq[ns + i] = NIB_CHARS[(w & 1) ? (byte & 15u) : (byte >> 4)];
        out->qual[ns + i] = (char)(blk[5 + w] + 33);
      }
    }
    nc += (int64_t)cigar.size(); nm += (int64_t)md.size(); ns += m.l_seq;
  }
  if (write) { out->cigar_off[n] = nc; out->md_off[n] = nm; out->seq_off[n] = ns; }
  out->cigar_bytes = nc; out->md_bytes = nm; out->seq_bytes = ns;
  return AVO_OK;
}
