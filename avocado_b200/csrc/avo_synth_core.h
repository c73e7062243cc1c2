// avo_synth_core.h -- the seeded, stateless synthetic read generator (bench / test infrastructure; not
// part of the reference): pure functions of (seed, read index) or (seed, reference position) through a
// counter-based hash.  One source for three builds -- the device kernels and the host entry point of
// libavocado_b200.so (avo_synth.cu), and the CPU oracle's own generator (oracle/avocado_oracle.cpp:
// bench.py's reference arm makes its input without loading the product library) -- which therefore
// produce identical bytes.  Shapes follow SURVEY.md 8(d).
#pragma once

#include <stdint.h>

#include "../../include/avocado_b200.h"
#include "../../include/avocado_b200_synth.h"

namespace avo {

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

HD uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
HD uint64_t hash3(uint64_t seed, uint64_t stream, uint64_t a, uint64_t b) {
  return mix64(mix64(mix64(seed ^ (stream * 0xD6E8FEB86659FD93ull)) ^ a) ^ (b * 0xA0761D6478BD642Full));
}
HD double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

HD uint32_t code_of(uint32_t two_bit) { return 1u << two_bit; }   // A=1 C=2 G=4 T=8

// seed of the per-sample streams (genotypes, reads); the default sample keeps the batch seed
HD uint64_t sample_seed(const avo_synth_params& P) {
  return P.sample ? mix64(P.seed ^ (P.sample * 0x9E3779B97F4A7C15ull)) : P.seed;
}

HD uint32_t ref_base(const avo_synth_params& P, int64_t p) {
  return (uint32_t)(hash3(P.seed, 1, (uint64_t)p, 0) & 3u);
}

struct TruthVar {
  int kind;      // 0 none, 1 SNV, 2 insertion after p, 3 deletion of p+1..p+len
  int len;
  uint32_t alt[2];  // SNV: 2-bit alt base per haplotype (equal unless tri-allelic)
  bool on[2];       // present on haplotype 0 / 1
};

HD TruthVar truth_at(const avo_synth_params& P, int64_t p) {
  TruthVar v;
  v.kind = 0; v.len = 0; v.alt[0] = v.alt[1] = 0; v.on[0] = v.on[1] = false;
  const uint64_t h = hash3(P.seed, 2, (uint64_t)p, 0);
  const double u = u01(h);
  if (u >= P.snp_rate + P.indel_rate) return v;
  const uint64_t g = hash3(P.seed, 3, (uint64_t)p, 0);
  const uint64_t gs = P.sample ? hash3(sample_seed(P), 3, (uint64_t)p, 0) : g;   // this sample's genotype
  const bool hom = (gs % 3u) == 0;                // het : hom-alt = 2 : 1
  const int hap = (int)((gs >> 8) & 1u);
  const bool carrier = !P.sample || ((gs >> 40) & 3u) != 0;
  v.on[0] = carrier && (hom || hap == 0);
  v.on[1] = carrier && (hom || hap == 1);
  if (u < P.snp_rate) {
    v.kind = 1;
    const uint32_t r = ref_base(P, p);
    const uint32_t a0 = (r + 1u + (uint32_t)((g >> 16) % 3u)) & 3u;
    v.alt[0] = v.alt[1] = a0;
    if (u01(hash3(P.seed, 4, (uint64_t)p, 0)) < P.triallelic_frac) {
      // two different alts, one per haplotype, both non-reference
      uint32_t a1 = (a0 + 1u) & 3u;
      if (a1 == r) a1 = (a1 + 1u) & 3u;
      v.alt[1] = a1;
      v.on[0] = v.on[1] = carrier;
    }
  } else {
    v.kind = ((g >> 20) & 1u) ? 2 : 3;
    int len = 1;
    uint64_t bits = g >> 24;
    while (len < P.max_indel && (bits & 1u)) { ++len; bits >>= 1; }
    v.len = len;
  }
  return v;
}

HD int draw_qual(uint64_t h) {
  const double u = u01(h);
  if (u < 0.0001) return 0;
  if (u < 0.0051) return 2;
  if (u < 0.0351) return 12;
  if (u < 0.1001) return 23;
  if (u < 0.3501) return 32;
  if (u < 0.8001) return 37;
  return 41;
}

HD double err_prob(int q) {   // 10^(-q/10) for the 7 levels above
  switch (q) {
    case 0: return 1.0;
    case 2: return 0.630957344480193;
    case 12: return 0.0630957344480193;
    case 23: return 0.00501187233627272;
    case 32: return 0.000630957344480193;
    case 37: return 0.000199526231496888;
    default: return 0.0000794328234724281;
  }
}

HD int draw_mapq(uint64_t h) {
  const double u = u01(h);
  if (u < 0.90) return 60;
  if (u < 0.94) return 40;
  if (u < 0.97) return 20;
  if (u < 0.99) return 5;
  return 0;
}

HD int64_t read_start(const avo_synth_params& P, int64_t i) {
  const uint64_t SS = sample_seed(P);
  const int64_t usable = (int64_t)P.contig_len - P.read_len - 2 * P.max_indel - 2;
  const double f = ((double)i + u01(hash3(SS, 5, (uint64_t)i, 0))) / (double)P.n_reads_total;
  int64_t s = (int64_t)(f * (double)usable);
  if (s < 0) s = 0;
  if (s > usable) s = usable;
  return s + P.first_pos;
}

// Sink concept: base(nibble, qual), op(len, code), pair(ref, read) for a mismatching base,
// delbase(nibble)... delend() for a deletion (include/avocado_b200.h: avo_read_batch.refb).
// One read, fully determined by i.
template <typename Sink>
HD void synth_read(const avo_synth_params& P, int64_t i, Sink& out, int32_t* o_start, int32_t* o_end,
                   uint8_t* o_mapq, uint8_t* o_flags) {
  const uint64_t SS = sample_seed(P);
  const uint64_t hr = hash3(SS, 6, (uint64_t)i, 0);
  const int hap = (int)(hr & 1u);
  const bool neg = (hr >> 1) & 1u;
  const int L = P.read_len;
  int clip_lead = 0, clip_trail = 0;
  if (u01(hash3(SS, 7, (uint64_t)i, 0)) < P.softclip_frac) {
    const int c = 1 + (int)((hr >> 8) % 20u);
    if ((hr >> 2) & 1u) clip_lead = c; else clip_trail = c;
  }
  const int64_t start = read_start(P, i);
  int64_t p = start;
  int emitted = 0;           // read bases emitted so far
  // pending run of '=' / 'X'
  int run_len = 0;
  uint32_t run_code = AVO_OP_MATCH;
  auto flush = [&]() { if (run_len) { out.op((uint32_t)run_len, run_code); run_len = 0; } };
  auto seq_base = [&](uint32_t truth2, int k) -> uint32_t {   // apply sequencing error; emit base
    const int q = draw_qual(hash3(SS, 8, (uint64_t)i, (uint64_t)k));
    uint32_t b = truth2;
    const uint64_t he = hash3(SS, 9, (uint64_t)i, (uint64_t)k);
    if (u01(he) < err_prob(q)) b = (truth2 + 1u + (uint32_t)((he >> 56) % 3u)) & 3u;
    out.base(code_of(b), (uint32_t)q);
    return b;
  };
  if (clip_lead) {
    for (int k = 0; k < clip_lead; ++k)
      seq_base((uint32_t)(hash3(SS, 10, (uint64_t)i, (uint64_t)k) & 3u), emitted++);
    out.op((uint32_t)clip_lead, AVO_OP_SOFTCLIP);
  }
  const int body = L - clip_trail;   // read bases that take part in the alignment (incl. lead clip)
  while (emitted < body) {
    const TruthVar v = truth_at(P, p);
    const uint32_t r2 = ref_base(P, p);
    uint32_t t2 = r2;
    if (v.kind == 1 && v.on[hap]) t2 = v.alt[hap];
    const uint32_t b2 = seq_base(t2, emitted++);
    const uint32_t code = (b2 == r2) ? AVO_OP_MATCH : AVO_OP_MISMATCH;
    if (run_len && code != run_code) flush();
    run_code = code;
    ++run_len;
    if (code == AVO_OP_MISMATCH) out.pair(code_of(r2), code_of(b2));
    ++p;
    if (emitted >= body) break;
    if (v.kind == 2 && v.on[hap]) {                 // insertion after the anchor base
      flush();
      const int n = v.len < body - emitted ? v.len : body - emitted;
      for (int k = 0; k < n; ++k)
        seq_base((uint32_t)(hash3(P.seed, 11, (uint64_t)(p - 1), (uint64_t)k) & 3u), emitted++);
      // an insertion that runs into the end of the read is reported as a soft clip, as aligners do
      out.op((uint32_t)n, (emitted >= body) ? AVO_OP_SOFTCLIP : AVO_OP_INS);
    } else if (v.kind == 3 && v.on[hap]) {          // deletion of p .. p+len-1 (anchor was p-1)
      flush();
      for (int k = 0; k < v.len; ++k) out.delbase(code_of(ref_base(P, p + k)));
      out.delend();
      out.op((uint32_t)v.len, AVO_OP_DEL);
      p += v.len;
    }
  }
  flush();
  if (clip_trail) {
    for (int k = 0; k < clip_trail; ++k)
      seq_base((uint32_t)(hash3(SS, 10, (uint64_t)i, (uint64_t)(100 + k)) & 3u), emitted++);
    out.op((uint32_t)clip_trail, AVO_OP_SOFTCLIP);
  }
  *o_start = (int32_t)start;
  *o_end = (int32_t)p;
  *o_mapq = (uint8_t)draw_mapq(hash3(SS, 12, (uint64_t)i, 0));
  *o_flags = neg ? AVO_READ_NEGATIVE_STRAND : 0;
}

struct CountSink {
  uint32_t n_ops = 0, n_refb = 0, pend = 0;   // n_refb in bytes
  HD void base(uint32_t, uint32_t) {}
  HD void op(uint32_t, uint32_t) { ++n_ops; }
  HD void pair(uint32_t, uint32_t) { ++n_refb; }
  HD void delbase(uint32_t) { ++pend; }
  HD void delend() { n_refb += (pend + 1) >> 1; pend = 0; }
};

struct WriteSink {
  uint8_t* payload; uint32_t* ops; uint8_t* refb;
  uint64_t blk0; // the read's first payload block
  uint64_t oi;   // global op index
  uint64_t fi;   // global refb byte index
  uint32_t pend_b = 0, pend_f = 0, n_pend = 0;
  uint32_t nb = 0, qhead = 0;   // bases written so far; phred of the first four
  HD void base(uint32_t nib, uint32_t q) {
    if (nb < 4) qhead |= q << (8 * nb);
    const uint32_t c = nb / AVO_BLOCK_BASES, w = nb - c * AVO_BLOCK_BASES;
    uint8_t* b = payload + ((blk0 + c) << 4);
    b[5 + w] = (uint8_t)q;
    if (w & 1) b[w >> 1] = (uint8_t)((pend_b << 4) | nib); else { pend_b = nib; b[w >> 1] = (uint8_t)(nib << 4); }
    ++nb;
  }
  HD void op(uint32_t len, uint32_t code) { ops[oi++] = (len << 4) | code; }
  HD void pair(uint32_t refn, uint32_t readn) { refb[fi++] = (uint8_t)((refn << 4) | readn); }
  HD void delbase(uint32_t nib) {
    if (n_pend & 1) refb[fi++] = (uint8_t)((pend_f << 4) | nib); else pend_f = nib;
    ++n_pend;
  }
  HD void delend() {
    if (n_pend & 1) refb[fi++] = (uint8_t)(pend_f << 4);
    n_pend = 0;
  }
  HD void begin(uint32_t n_blocks) {     // unused slots and the two pad bytes of every block are zero
    uint64_t* z = reinterpret_cast<uint64_t*>(payload + (blk0 << 4));
    for (uint32_t k = 0; k < 2 * n_blocks; ++k) z[k] = 0;
  }
  HD void finish() {}
};

HD uint32_t blocks_of(uint32_t l_seq) { return (l_seq + AVO_BLOCK_BASES - 1) / AVO_BLOCK_BASES; }

HD void fill_meta(avo_read_meta& m, int32_t start, int32_t end, uint8_t mapq, uint8_t flags,
                  uint32_t seq_off, uint32_t op_off, uint32_t refb_off, uint32_t n_ops, uint32_t qhead,
                  uint32_t l_seq) {
  m.start = start; m.end = end; m.seq_off = seq_off; m.op_off = op_off; m.refb_off = refb_off;
  m.n_ops = (uint16_t)n_ops; m.mapq = mapq; m.flags = flags; m.qhead = qhead; m.l_seq = l_seq;
}


}  // namespace avo
