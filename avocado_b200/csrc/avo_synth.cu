// avo_synth.cu -- seeded, stateless synthetic read generator (bench / test infrastructure; not part
// of the reference).  Shapes follow SURVEY.md 8(d): i.i.d. reference, truth SNVs 1e-3/bp and indels
// 1e-4/bp (length geometric p = 0.5, capped at 20), het:hom 2:1, 1 % of SNV sites with two alts,
// fixed-length reads sorted by start, per-base qualities from a 6-level distribution plus a rare
// Q0, sequencing errors at 10^(-q/10), mapQ from a 5-level distribution, 2 % soft-clipped reads.
//
// Everything is a pure function of (seed, read index) or (seed, reference position) through a
// counter-based hash, so the same core runs on the device (full-size benchmarks never cross PCIe)
// and on the host (small cases that are handed to the CPU oracle), and produces identical bytes.
#include <cub/device/device_scan.cuh>

#include <vector>

#include "avo_internal.h"
#include "avo_synth.h"

namespace avo {

#define HD __host__ __device__ __forceinline__

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
      const int n = min(v.len, body - emitted);
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
    uint4* z = reinterpret_cast<uint4*>(payload + (blk0 << 4));
    for (uint32_t k = 0; k < n_blocks; ++k) z[k] = make_uint4(0, 0, 0, 0);
  }
  HD void finish() {}
};

HD uint32_t blocks_of(uint32_t l_seq) { return (l_seq + AVO_BLOCK_BASES - 1) / AVO_BLOCK_BASES; }

__global__ void k_synth_count(avo_synth_params P, int64_t first, int64_t n, uint32_t* __restrict__ n_ops,
                              uint32_t* __restrict__ n_refb) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= n) return;
  CountSink c;
  int32_t s, e; uint8_t mq, fl;
  synth_read(P, first + j, c, &s, &e, &mq, &fl);
  n_ops[j] = c.n_ops;
  n_refb[j] = c.n_refb;
}

HD void fill_meta(avo_read_meta& m, int32_t start, int32_t end, uint8_t mapq, uint8_t flags,
                  uint32_t seq_off, uint32_t op_off, uint32_t refb_off, uint32_t n_ops, uint32_t qhead,
                  uint32_t l_seq) {
  m.start = start; m.end = end; m.seq_off = seq_off; m.op_off = op_off; m.refb_off = refb_off;
  m.n_ops = (uint16_t)n_ops; m.mapq = mapq; m.flags = flags; m.qhead = qhead; m.l_seq = l_seq;
}

__global__ void k_synth_write(avo_synth_params P, int64_t first, int64_t n, avo_read_meta* __restrict__ meta,
                              uint8_t* __restrict__ payload,
                              const uint32_t* __restrict__ op_off, uint32_t* __restrict__ ops,
                              const uint32_t* __restrict__ refb_off, uint8_t* __restrict__ refb) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t lpad = blocks_of((uint32_t)P.read_len);
  WriteSink w;
  w.payload = payload; w.ops = ops; w.refb = refb;
  w.blk0 = (uint64_t)j * lpad; w.oi = op_off[j]; w.fi = refb_off[j];
  w.begin(lpad);
  int32_t s, e; uint8_t mq, fl;
  synth_read(P, first + j, w, &s, &e, &mq, &fl);
  w.finish();
  fill_meta(meta[j], s, e, mq, fl, (uint32_t)(j * lpad), op_off[j], refb_off[j], op_off[j + 1] - op_off[j],
            w.qhead, (uint32_t)P.read_len);
}

}  // namespace avo

using namespace avo;

extern "C" {

int avo_synth_default_params(avo_synth_params* p) {
  if (!p) return AVO_E_INVALID;
  p->seed = 0xA70CAD0ull;
  p->contig_len = 1000000;
  p->first_pos = 0;
  p->n_reads_total = 200000;
  p->read_len = 150;
  p->max_indel = 20;
  p->snp_rate = 1.0e-3;
  p->indel_rate = 1.0e-4;
  p->triallelic_frac = 0.01;
  p->softclip_frac = 0.02;
  p->sample = 0;
  return AVO_OK;
}

static int check_synth(const avo_synth_params* p, int64_t first, int64_t n) {
  if (!p || n < 0 || first < 0 || first + n > p->n_reads_total) return AVO_E_INVALID;
  if (p->read_len < 30 || p->read_len > 1000 || p->max_indel < 1 || p->max_indel > 20) return AVO_E_INVALID;
  if ((int64_t)p->contig_len < (int64_t)p->read_len + 2 * p->max_indel + 4) return AVO_E_INVALID;
  const uint64_t lpad = (uint64_t)blocks_of((uint32_t)p->read_len);
  if ((uint64_t)n * lpad > 0xFFFFFFFFull) return AVO_E_INVALID;
  return AVO_OK;
}

// Host generation of reads [first, first + n): fills n_ops / n_refb and, when `out` arrays are
// non-NULL, the batch.  Call once with out == NULL to size, then with buffers.
int avo_synth_host(const avo_synth_params* p, int64_t first, int64_t n, int64_t* n_ops, int64_t* n_refb,
                   avo_packed_out* out) {
  int rc = check_synth(p, first, n);
  if (rc != AVO_OK) return rc;
  const uint32_t lpad = blocks_of((uint32_t)p->read_len);
  uint64_t to = 0, tf = 0;
  if (!out) {
    for (int64_t j = 0; j < n; ++j) {
      CountSink c;
      int32_t s, e; uint8_t mq, fl;
      synth_read(*p, first + j, c, &s, &e, &mq, &fl);
      to += c.n_ops; tf += c.n_refb;
    }
    if (n_ops) *n_ops = (int64_t)to;
    if (n_refb) *n_refb = (int64_t)tf;
    return AVO_OK;
  }
  for (int64_t j = 0; j < n; ++j) {
    CountSink c;
    int32_t s, e; uint8_t mq, fl;
    synth_read(*p, first + j, c, &s, &e, &mq, &fl);
    if ((int64_t)(to + c.n_ops) > out->n_ops || (int64_t)(tf + c.n_refb) > out->n_refb ||
        j >= out->n_reads)
      return AVO_E_CAPACITY;
    WriteSink w;
    w.payload = out->payload; w.ops = out->ops; w.refb = out->refb;
    w.blk0 = (uint64_t)j * lpad; w.oi = to; w.fi = tf;
    w.begin(lpad);
    synth_read(*p, first + j, w, &s, &e, &mq, &fl);
    w.finish();
    fill_meta(out->meta[j], s, e, mq, fl, (uint32_t)(j * lpad), (uint32_t)to, (uint32_t)tf, c.n_ops,
              w.qhead, (uint32_t)p->read_len);
    to += c.n_ops; tf += c.n_refb;
  }
  out->n_reads = n; out->n_blocks = (int64_t)n * lpad; out->n_ops = (int64_t)to; out->n_refb = (int64_t)tf;
  if (n_ops) *n_ops = (int64_t)to;
  if (n_refb) *n_refb = (int64_t)tf;
  return AVO_OK;
}

// Device generation.  Phase 1 (out == NULL): op_off_dev / refb_off_dev ([n+1] device uint32
// arrays) receive the offsets and *n_ops / *n_refb the totals.  Phase 2: the caller passes device
// buffers sized from phase 1 in `out` (all pointers are device pointers) together with the same
// offset arrays.
int avo_synth_device(const avo_synth_params* p, int64_t first, int64_t n, uint32_t* op_off_dev,
                     uint32_t* refb_off_dev, int64_t* n_ops, int64_t* n_refb, avo_packed_out* out,
                     void* cuda_stream) {
  int rc = check_synth(p, first, n);
  if (rc != AVO_OK) return rc;
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const int grid = (int)((n + 1 + 127) / 128);
  if (!out) {
    if (!op_off_dev || !refb_off_dev || !n_ops || !n_refb) return AVO_E_INVALID;
    uint32_t *cnt_o = nullptr, *cnt_f = nullptr;
    void* temp = nullptr;
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)(n + 1));
    if (cudaMalloc(&cnt_o, (size_t)(n + 1) * 4) != cudaSuccess) return AVO_E_NOMEM;
    if (cudaMalloc(&cnt_f, (size_t)(n + 1) * 4) != cudaSuccess) { cudaFree(cnt_o); return AVO_E_NOMEM; }
    if (cudaMalloc(&temp, tb + 16) != cudaSuccess) { cudaFree(cnt_o); cudaFree(cnt_f); return AVO_E_NOMEM; }
    cudaMemsetAsync(cnt_o, 0, (size_t)(n + 1) * 4, s);
    cudaMemsetAsync(cnt_f, 0, (size_t)(n + 1) * 4, s);
    if (n > 0) k_synth_count<<<grid, 128, 0, s>>>(*p, first, n, cnt_o, cnt_f);
    cub::DeviceScan::ExclusiveSum(temp, tb, cnt_o, op_off_dev, (int)(n + 1), s);
    cub::DeviceScan::ExclusiveSum(temp, tb, cnt_f, refb_off_dev, (int)(n + 1), s);
    uint32_t tot[2] = {0, 0};
    cudaMemcpyAsync(&tot[0], op_off_dev + n, 4, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(&tot[1], refb_off_dev + n, 4, cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(cnt_o); cudaFree(cnt_f); cudaFree(temp);
    if (e != cudaSuccess) { cudaGetLastError(); return AVO_E_CUDA; }
    *n_ops = tot[0]; *n_refb = tot[1];
    return AVO_OK;
  }
  if (n > 0)
    k_synth_write<<<grid, 128, 0, s>>>(*p, first, n, out->meta, out->payload, op_off_dev, out->ops,
                                       refb_off_dev, out->refb);
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { cudaGetLastError(); return AVO_E_CUDA; }
  return AVO_OK;
}

}  // extern "C"
