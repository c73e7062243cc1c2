// avo_api.cu -- the C ABI of include/avocado_b200.h: context, device memory, host<->device
// staging, the score table, and the orchestration of the kernels in avo_setup.cu,
// avo_discover.cu and avo_call.cu.  No CPU fallback exists anywhere in this file: every entry
// point needs a CUDA device and reports AVO_E_CUDA when the runtime fails.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <thread>

#include <nvtx3/nvToolsExt.h>     // header-only: ranges cost nothing unless a profiler is attached

#include "avo_internal.h"

using namespace avo;

namespace {

// NVTX ranges named after the reference's own timers (avocado-core/.../avocado/Timers.scala:25-69), so that
// an nsys trace of a JVM shim around this library reads like the reference's instrumentation report.
struct Range {
  explicit Range(const char* name) { nvtxRangePushA(name); }
  ~Range() { nvtxRangePop(); }
};
#define T_DISCOVER_AND_CALL "Discover variants and call genotypes"          /* Timers.DiscoverAndCall */
#define T_CALL_GENOTYPES "Call genotypes"                                    /* Timers.CallGenotypes */
#define T_DISCOVERING "Discovering variants in reads"                        /* Timers.DiscoveringVariants */
#define T_BUILDING_TREES "Building interval tree"                            /* Timers.BuildingTrees: the sorted site table + its index */
#define T_TREE_JOIN "Running broadcast join with interval tree"              /* Timers.TreeJoin */
#define T_OBSERVE_READS "Observing variants in reads"                        /* Timers.ObserveReads */
#define T_EMIT_GENOTYPES "Emit observations as genotype calls"               /* Timers.EmitGenotypes */
#define T_EXTRACT_ALIGNMENT "Parsing alignment (CIGAR/MD tag)"               /* Timers.ExtractingAlignment */

constexpr int32_t LNK_N = 1 << 16;

enum Slot {
  SL_META, SL_COORDS, SL_PAYLOAD, SL_OPS, SL_REFB, SL_WIRE_META, SL_WIRE_OPS, SL_WIRE_OPC, SL_WIRE_OFFS, SL_WIRE_SIZES, SL_WIRE_EXC,
  SL_S_CONTIG, SL_S_START, SL_S_REFLEN, SL_S_ALTLEN, SL_S_AOFF, SL_S_ALLELES, SL_S_END, SL_S_PME,
  SL_S_COUNT,
  SL_CNV, SL_STATS, SL_SMALL, SL_RIDX, SL_SIDX, SL_TEMP, SL_LUT, SL_ARENA, SL_U_NEXT, SL_PACK, SL_G_END, SL_G_PME, SL_G_SIDX, SL_G_COUNT,
  SL_RES,                      // one block holding every result column
  SL_U_START, SL_U_REFLEN, SL_U_ALTLEN, SL_U_COUNT, SL_U_SRC,
  SL_L_POS, SL_L_REFLEN, SL_L_ALTLEN, SL_L_NIBS, SL_L_SRC, SL_L_BLK, SL_L_HASH,
  SL_C_RBASE, SL_C_BLEN, SL_C_BOFF, SL_C_BUCKETS,
  SL_G_PBASE, SL_G_NBLK, SL_G_SLOT, SL_G_CLO, SL_G_IDX, SL_G_BLK0, SL_G_COMPACT,
  SL_AS_FLAG, SL_AS_JPOS, SL_AS_JR, SL_AS_JIDX, SL_AS_JCOORD, SL_AS_COVER, SL_AS_SCOVER, SL_AS_WCOUNT, SL_AS_WBASE,
  SL_AS_RES, SL_AS_START,
  SL_F_IN, SL_F_OUT,
  SL_T_START, SL_T_END, SL_T_MAPQ, SL_T_FLAGS, SL_T_CIGAR, SL_T_CIGAR_OFF, SL_T_MD, SL_T_MD_OFF, SL_T_SEQ,
  SL_T_SEQ_OFF, SL_T_QUAL, SL_T_QUAL_OFF, SL_T_KEEP, SL_T_COUNTS, SL_T_OFFSETS, SL_EV_CSNV, SL_EV_CIND, SL_EV_IFIRST,
  SL_J_PRESENT, SL_J_NALT, SL_J_NREF, SL_J_NOTHER, SL_J_NOCALL, SL_J_GL, SL_J_ALT, SL_J_CALLED, SL_J_OUT,
  SL_JD_DP, SL_JD_REFDP, SL_JD_SB, SL_JD_OUT,
  SL_TABLE, SL_TCOUNT, SL_KEYS_A, SL_KEYS_B, SL_VALS_A, SL_VALS_B, SL_FOOT, SL_AOFF,
  SL_COUNT_
};

struct Buf { void* p = nullptr; size_t cap = 0; };

}  // namespace

struct ShardState;
void free_shard_state(struct avo_ctx* ctx);

struct avo_ctx {
  int device = 0;
  int num_sms = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;
  Buf bufs[SL_COUNT_];
  double* d_lnk = nullptr;
  double* d_lnfact = nullptr;
  double* d_lgam = nullptr;        // lgamma(k), k = 0..AVO_MAX_PLOIDY + 2 (host libm)
  // score table cache key
  int lut_minp = -1, lut_maxp = -1, lut_maxq = -1, lut_maxmq = -1;
  avo_timing timing{};
  cudaEvent_t ev[10] = {};         // [8], [9]: around k_call_pairs
  bool pairs_timed = false;
  int64_t indel_guess = 0;         // indel candidates seen by the previous discovery (sizes the table)
  int64_t site_guess = 0;          // survivors / allele nibbles of the previous discovery (+ margin): the
  int64_t nib_guess = 0;           // capacities the next one runs with (checked when its counts arrive)
  int disc_blocks_per_sm = 0;      // resident CTAs of k_discover_tiles per SM (occupancy query, once)
  int ev_blocks_per_sm = 0;        // ... of k_discover_events
  ShardState* shard = nullptr;     // between avo_shard_discover / _call / _finish
  BatchStats* h_stats = nullptr;   // pinned
  int32_t* h_small = nullptr;      // pinned scratch (16 ints)
  // device -> host results travel through one page-locked staging buffer (a few large DMA copies
  // instead of one pageable copy per column) and are scattered to the caller's arrays after the sync
  struct Pending { void* dst; size_t off; size_t bytes; };
  uint8_t* h_stage = nullptr;
  size_t h_stage_cap = 0, h_stage_used = 0, h_stage_want = 0;
  std::vector<Pending> pending;
  // host-gathered payload (avo_params.host_payload == 2): set per call by upload_reads
  const uint8_t* gather_src = nullptr;   // the caller's host payload, read by the library's threads
  int gather_threads = 0;
  size_t slots_guess = 0;          // bucket slots of the previous call stage (+ 1/8): capacity of the next
  uint32_t* h_gidx = nullptr;      // pinned: block indices from the device
  uint8_t* h_gblk = nullptr;       // pinned: the gathered blocks
  size_t h_gcap = 0;               // capacity of both, in blocks
};

namespace {

int fail(avo_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return fail(ctx, AVO_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
  } while (0)

#define CKS(call)                                   \
  do {                                              \
    int s__ = (call);                               \
    if (s__ != AVO_OK) return s__;                  \
  } while (0)

int get_buf(avo_ctx* ctx, int slot, size_t bytes, void** out) {
  Buf& b = ctx->bufs[slot];
  if (bytes == 0) bytes = 16;
  if (b.cap < bytes) {
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(ctx, AVO_E_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
  }
  *out = b.p;
  return AVO_OK;
}

template <typename T>
int stage_in(avo_ctx* ctx, int slot, const T* src, size_t count, int mem, const T** out) {
  if (mem == AVO_MEM_DEVICE) { *out = src; return AVO_OK; }
  void* d;
  CKS(get_buf(ctx, slot, count * sizeof(T), &d));
  if (count) {
    CK(cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    ctx->timing.h2d_bytes += (int64_t)(count * sizeof(T));
  }
  *out = static_cast<const T*>(d);
  return AVO_OK;
}

// Device -> host copy of `total` bytes whose pieces go to nsc host arrays: staged when the staging
// buffer has room (it grows to the previous call's demand at the next begin_call), direct otherwise.
struct Scatter { void* dst; size_t off; size_t bytes; };
int d2h_block(avo_ctx* ctx, const void* d_src, size_t total, const Scatter* sc, int nsc) {
  if (total == 0) return AVO_OK;
  constexpr size_t STAGE_MAX = (size_t)128 << 20;     // gVCF-sized outputs go straight to the caller
  const size_t at = (ctx->h_stage_used + 255) & ~(size_t)255;
  if (at + total <= STAGE_MAX) ctx->h_stage_want = std::max(ctx->h_stage_want, at + total);
  ctx->timing.d2h_bytes += (int64_t)total;
  if (at + total <= ctx->h_stage_cap) {
    CK(cudaMemcpyAsync(ctx->h_stage + at, d_src, total, cudaMemcpyDeviceToHost, ctx->stream));
    for (int k = 0; k < nsc; ++k)
      if (sc[k].bytes) ctx->pending.push_back({sc[k].dst, at + sc[k].off, sc[k].bytes});
    ctx->h_stage_used = at + total;
    return AVO_OK;
  }
  for (int k = 0; k < nsc; ++k)
    if (sc[k].bytes)
      CK(cudaMemcpyAsync(sc[k].dst, (const char*)d_src + sc[k].off, sc[k].bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return AVO_OK;
}

// after the stream is idle
void scatter_pending(avo_ctx* ctx) {
  for (const auto& p : ctx->pending) memcpy(p.dst, ctx->h_stage + p.off, p.bytes);
  ctx->pending.clear();
}

// Page-locked host memory is mapped into the device address space (UVA): returns the pointer the
// kernels can dereference, or null when `p` is pageable / unknown to the runtime.
template <typename T>
const T* mapped_alias(const T* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
  return static_cast<const T*>(a.devicePointer);
}

int validate_reads(avo_ctx* ctx, const avo_read_batch* r) {
  if (!r) return fail(ctx, AVO_E_INVALID, "reads is NULL");
  if (r->n_reads < 0 || r->n_reads >= (int64_t)INT32_MAX)
    return fail(ctx, AVO_E_INVALID, "n_reads out of range (must be < 2^31 per batch)");
  if (r->n_blocks < 0 || r->n_blocks > (int64_t)UINT32_MAX || r->n_ops < 0 ||
      r->n_ops > (int64_t)UINT32_MAX || r->n_refb < 0 || r->n_refb > (int64_t)INT32_MAX)
    return fail(ctx, AVO_E_INVALID, "batch totals must fit 32-bit offsets; split the batch");
  const bool wire6 = r->wire6_meta && (r->n_ops == 0 || (r->wire6_oplen && r->wire6_opcode));
  const bool wire = wire6 || (r->wire_meta && r->wire_ops);
  if (r->wire6_meta && !wire6) return fail(ctx, AVO_E_INVALID, "wire6_meta needs wire6_oplen and wire6_opcode");
  if (!r->wire6_meta && (r->wire_meta != nullptr) != (r->wire_ops != nullptr) && r->n_ops > 0)
    return fail(ctx, AVO_E_INVALID, "wire_meta and wire_ops go together");
  if (r->wire6_meta && (r->n_wire6_exc < 0 || r->n_wire6_exc > r->n_reads || (r->n_wire6_exc > 0 && !r->wire6_exc)))
    return fail(ctx, AVO_E_INVALID, "bad wire6 exception list");
  if (wire && r->mem != AVO_MEM_HOST) return fail(ctx, AVO_E_INVALID, "wire columns are for host batches");
  if (r->n_reads > 0 && !r->meta && !wire) return fail(ctx, AVO_E_INVALID, "read batch has no meta column");
  if ((r->n_blocks > 0 && !r->payload) || (r->n_ops > 0 && !r->ops && !wire) ||
      (r->n_refb > 0 && !r->refb))
    return fail(ctx, AVO_E_INVALID, "read batch has NULL columns");
  if (r->mem != AVO_MEM_HOST && r->mem != AVO_MEM_DEVICE) return fail(ctx, AVO_E_INVALID, "bad reads->mem");
  if (((uintptr_t)r->payload & 15u) || ((uintptr_t)r->meta & 15u) || ((uintptr_t)r->wire_meta & 15u) ||
      ((uintptr_t)r->wire6_meta & 1u) || ((uintptr_t)r->wire6_exc & 15u))
    return fail(ctx, AVO_E_INVALID, "reads->payload and reads->meta must be 16-byte aligned");
  return AVO_OK;
}

// sparse: the mode reads the payload only at variant sites and indels, so a page-locked host
// batch keeps it where it is and the kernels fetch single sectors over PCIe
int upload_reads(avo_ctx* ctx, const avo_read_batch* r, bool sparse, DReads* R, const avo_params* params = nullptr) {
  const size_t n = (size_t)r->n_reads;
  ctx->gather_src = nullptr;
  R->n = r->n_reads;
  R->contig = r->contig_id;
  R->qhead_short = 0;
  R->n_ops = r->n_ops;
  R->n_refb = r->n_refb;
  R->coords = nullptr;             // the dense (start, end) copy: only the host-gathered route's join reads it
  R->n_snv = R->n_indel = 0;
  R->snv_pos = nullptr; R->snv_pq = nullptr; R->snv_first = nullptr; R->indel_read = nullptr; R->indel_op = nullptr;
  if (r->events && r->mem == AVO_MEM_DEVICE && r->n_reads > 0) {     // candidate events: discovery counts them
    const avo_read_events* e = r->events;
    if (e->n_snv < 0 || e->n_indel < 0 || !e->snv_first || (e->n_snv > 0 && (!e->snv_pos || !e->snv_pq)) ||
        (e->n_indel > 0 && (!e->indel_read || !e->indel_op)) || !e->stats.valid)
      return fail(ctx, AVO_E_INVALID, "reads->events is incomplete (fill it with avo_derive_events)");
    R->n_snv = e->n_snv; R->n_indel = e->n_indel;
    R->snv_pos = e->snv_pos; R->snv_pq = e->snv_pq; R->snv_first = e->snv_first;
    R->indel_read = e->indel_read; R->indel_op = e->indel_op;
  }
  if (r->wire6_meta && r->mem == AVO_MEM_HOST) {
    // 6-byte records: starts, ends, lengths and offsets are rebuilt on the device; two head qualities
    const avo_read_wire6* dw;
    const uint8_t *dol = nullptr, *doc = nullptr;
    const avo_wire_exception* dexc = nullptr;
    CKS(stage_in(ctx, SL_WIRE_META, r->wire6_meta, n, r->mem, &dw));
    if (r->n_ops > 0) {
      CKS(stage_in(ctx, SL_WIRE_OPS, r->wire6_oplen, (size_t)r->n_ops, r->mem, &dol));
      CKS(stage_in(ctx, SL_WIRE_OPC, r->wire6_opcode, (size_t)(r->n_ops + 1) / 2, r->mem, &doc));
    }
    if (r->n_wire6_exc > 0) CKS(stage_in(ctx, SL_WIRE_EXC, r->wire6_exc, (size_t)r->n_wire6_exc, r->mem, &dexc));
    void *dm, *dops, *doffs, *dsizes, *temp;
    CKS(get_buf(ctx, SL_META, n * sizeof(avo_read_meta), &dm));
    CKS(get_buf(ctx, SL_OPS, (size_t)r->n_ops * 4, &dops));
    CKS(get_buf(ctx, SL_WIRE_OFFS, n * 12, &doffs));
    CKS(get_buf(ctx, SL_WIRE_SIZES, n * 8, &dsizes));
    const size_t tb = wire6_temp_bytes(r->n_reads);
    CKS(get_buf(ctx, SL_TEMP, tb, &temp));
    CK(launch_wire6_expand(dw, dol, doc, dexc, r->n_wire6_exc, r->wire6_base, r->n_reads, r->n_ops, doffs, dsizes,
                           (avo_read_meta*)dm, (uint32_t*)dops, temp, tb, ctx->stream, &ctx->timing.n_launches));
    R->meta = (const avo_read_meta*)dm;
    R->ops = (const uint32_t*)dops;
    R->qhead_short = 1;
  } else if (r->wire_meta && r->wire_ops && r->mem == AVO_MEM_HOST) {
    // compact wire columns: half the bytes over PCIe; offsets rebuilt by a scan on the device
    const avo_read_wire* dw;
    const uint16_t* dwo;
    CKS(stage_in(ctx, SL_WIRE_META, r->wire_meta, n, r->mem, &dw));
    CKS(stage_in(ctx, SL_WIRE_OPS, r->wire_ops, (size_t)r->n_ops, r->mem, &dwo));
    void *dm, *dops, *doffs, *temp;
    CKS(get_buf(ctx, SL_META, n * sizeof(avo_read_meta), &dm));
    CKS(get_buf(ctx, SL_OPS, (size_t)r->n_ops * 4, &dops));
    CKS(get_buf(ctx, SL_WIRE_OFFS, n * 12, &doffs));
    const size_t tb = wire_temp_bytes(r->n_reads);
    CKS(get_buf(ctx, SL_TEMP, tb, &temp));
    CK(launch_wire_expand(dw, dwo, r->n_reads, r->n_ops, doffs, (avo_read_meta*)dm, (uint32_t*)dops, temp, tb,
                          ctx->stream));
    ctx->timing.n_launches += 3;
    R->meta = (const avo_read_meta*)dm;
    R->ops = (const uint32_t*)dops;
  } else {
    CKS(stage_in(ctx, SL_META, r->meta, n, r->mem, &R->meta));
    CKS(stage_in(ctx, SL_OPS, r->ops, (size_t)r->n_ops, r->mem, &R->ops));
  }
  CKS(stage_in(ctx, SL_REFB, r->refb, (size_t)r->n_refb, r->mem, &R->refb));
  if (sparse && r->mem == AVO_MEM_HOST && r->n_blocks > 0) {
    const uint8_t* b = mapped_alias(r->payload);
    if (b) {
      R->payload = b;
      ctx->timing.payload_in_place = 1;
      if (params && params->host_payload == 2) {     // the call stage gathers its blocks on the host
        ctx->gather_src = r->payload;
        void* c;
        CKS(get_buf(ctx, SL_COORDS, n * sizeof(int2), &c));
        R->coords = (int2*)c;
        ctx->gather_threads = params->gather_threads > 0 ? std::min(params->gather_threads, 64) : 4;
        ctx->timing.payload_in_place = 2;
      }
      return AVO_OK;
    }
  }
  CKS(stage_in(ctx, SL_PAYLOAD, r->payload, (size_t)r->n_blocks * AVO_BLOCK_BYTES, r->mem, &R->payload));
  return AVO_OK;
}

int begin_call(avo_ctx* ctx) {
  if (!ctx) return AVO_E_INVALID;
  ctx->err.clear();
  CK(cudaSetDevice(ctx->device));
  memset(&ctx->timing, 0, sizeof ctx->timing);
  ctx->pending.clear();
  ctx->h_stage_used = 0;
  if (ctx->h_stage_want > ctx->h_stage_cap) {
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    ctx->h_stage = nullptr; ctx->h_stage_cap = 0;
    const size_t want = ctx->h_stage_want + ctx->h_stage_want / 4 + 4096;
    if (cudaMallocHost(&ctx->h_stage, want) == cudaSuccess) ctx->h_stage_cap = want;
    else cudaGetLastError();              // no staging: copies go straight to the caller's arrays
  }
  ctx->h_stage_want = 0;
  CK(cudaEventRecord(ctx->ev[0], ctx->stream));
  return AVO_OK;
}

float ev_ms(avo_ctx* ctx, int a, int b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]) != cudaSuccess) { cudaGetLastError(); return 0.f; }
  return ms;
}

// The caller's bounds (avo_batch_stats), to be verified on the device by the discovery kernel.
bool stats_from_hints(const avo_read_batch* r, BatchStats* h) {
  if (r->events && r->mem == AVO_MEM_DEVICE && r->events->stats.valid && r->n_reads > 0) {
    // measured by avo_derive_events together with the sort order: exact
    h->min_start = r->events->stats.min_start; h->max_end = r->events->stats.max_end; h->max_span = r->events->stats.max_span;
    h->unsorted = 0;
    return true;
  }
  if (!r->stats.valid || r->n_reads == 0) return false;
  if (r->stats.max_end < r->stats.min_start || r->stats.max_span < 0) return false;
  h->min_start = r->stats.min_start; h->max_end = r->stats.max_end; h->max_span = r->stats.max_span;
  h->unsorted = 0;
  return true;
}

// batch statistics + sortedness (one small D2H + sync)
int batch_stats(avo_ctx* ctx, const DReads& R, BatchStats* h) {
  void* d;
  CKS(get_buf(ctx, SL_STATS, sizeof(BatchStats), &d));
  CK(launch_batch_stats(R, (BatchStats*)d, ctx->stream));
  ctx->timing.n_launches += (R.n > 0) ? 2 : 1;
  CK(cudaMemcpyAsync(ctx->h_stats, d, sizeof(BatchStats), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *h = *ctx->h_stats;
  if (R.n == 0) { h->min_start = 0; h->max_end = 0; h->max_span = 0; }
  if (h->unsorted) return fail(ctx, AVO_E_UNSORTED, "reads are not sorted by start");
  return AVO_OK;
}

// The score table of ScoredObservation.createScores (ScoredObservation.scala:101-165) reduced to
// what the join can reach: L[g] = ln((m-g) eps + g (1-eps)), eps = 1 - s_map s_base
// (Observer.scala:151-185), for m in minPloidy..maxPloidy, q in 0..maxQ, mq in 0..maxMapQ.  Built
// with the host libm so that table values do not depend on the device's log/pow.
int ensure_lut(avo_ctx* ctx, int minp, int maxp, int maxq, int maxmq, const double** out) {
  const int ns = maxp + 1;
  // layout [copy number - minp][mq' : 128][q' : 128][ns]: the kernel indexes it with the bit
  // field (cn index << 14 | mq' << 7 | q') taken straight from its record word
  const size_t count = (size_t)(maxp - minp + 1) * 128 * 128 * ns;
  void* d;
  if (ctx->lut_minp == minp && ctx->lut_maxp == maxp && ctx->lut_maxq == maxq &&
      ctx->lut_maxmq == maxmq && ctx->bufs[SL_LUT].p) {
    *out = (const double*)ctx->bufs[SL_LUT].p;
    return AVO_OK;
  }
  CKS(get_buf(ctx, SL_LUT, count * sizeof(double), &d));
  std::vector<double> lut(count);
  for (int m = minp; m <= maxp; ++m)
    for (int q = 0; q <= maxq; ++q)
      for (int mq = 0; mq <= maxmq; ++mq) {
        const double s_map = 1.0 - std::pow(10.0, -(double)mq / 10.0);   // PhredUtils (ADAM)
        const double s_base = 1.0 - std::pow(10.0, -(double)q / 10.0);
        const double ome = s_map * s_base;
        const double eps = 1.0 - ome;
        double* row = &lut[((((size_t)(m - minp) * 128 + mq) * 128) + q) * ns];
        for (int g = 0; g < ns; ++g)
          row[g] = (g <= m && m > 0) ? std::log((m - g) * eps + g * ome) : 0.0;
      }
  CK(cudaMemcpyAsync(d, lut.data(), count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));   // `lut` goes out of scope
  ctx->lut_minp = minp; ctx->lut_maxp = maxp; ctx->lut_maxq = maxq; ctx->lut_maxmq = maxmq;
  *out = (const double*)d;
  return AVO_OK;
}

// ---- device-resident site table ------------------------------------------------------------------
struct DevSites {
  int32_t n = 0;                // rows (tables assembled on the device: known after the chain's sync)
  uint32_t n_nibbles = 0;
  const int32_t* contig = nullptr;
  const int32_t* start = nullptr;
  const int32_t* ref_len = nullptr;
  const int32_t* alt_len = nullptr;
  const uint32_t* allele_off = nullptr;
  const uint8_t* alleles4 = nullptr;
  const int32_t* count = nullptr;
  // tables assembled on the device in this launch chain: capacities, the header, derived columns
  const DSiteHdr* hdr = nullptr;
  int32_t cap = 0;
  uint32_t nib_cap = 0;
  const int32_t* end = nullptr;
  const int32_t* pme = nullptr;
  const int32_t* sidx = nullptr;
  bool caller_arrays = false;   // the columns ARE the caller's device arrays (nothing to export)
};

// words of the SL_SMALL block (int32): what the kernels count for the host
enum SmallWord {
  SW_UNSORTED = 0, SW_CRANGE = 1 /* 2 */, SW_SRANGE = 4 /* 2 */, SW_TILE = 8, SW_FLAGS = 9, SW_NOUT = 10, SW_NINDEL = 11,
  SW_JOIN = 14, SW_HDR = 16 /* DSiteHdr: 8 words */, SW_HDR2 = 24 /* the global table's, sharded runs */, SW_COUNT_ = 32
};

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct ResLayout {
  size_t off[24];
  size_t total;
};

// one allocation holding all result columns (order = avo_site_results fields)
ResLayout res_layout(int64_t n, int ns) {
  const size_t sz[24] = {
      (size_t)n,                                   // emitted u8
      (size_t)n * 4, (size_t)n * 4, (size_t)n * 4, (size_t)n * 4, (size_t)n * 4,  // 5 x i32
      (size_t)n * 8,                               // square_mapq
      (size_t)n * ns * 8, (size_t)n * ns * 8, (size_t)n * ns * 8, (size_t)n * ns * 8,  // 4 x ll
      (size_t)n,                                   // first_is_ref
      (size_t)n * 4,                               // copy_number
      (size_t)n * 4, (size_t)n * 4, (size_t)n * 4, // n_alt n_ref n_other
      (size_t)n * 8,                               // qual
      (size_t)n * 4,                               // gq
      (size_t)n * 16,                              // sb
      (size_t)n * 4, (size_t)n * 4,                // fisher rms
      (size_t)n * ns * 8, (size_t)n * ns * 8,      // gl nonref_gl
      0};
  ResLayout L;
  size_t o = 0;
  for (int i = 0; i < 23; ++i) { L.off[i] = o; o += align_up(sz[i], 256); }
  L.off[23] = o;
  L.total = o;
  return L;
}

void res_pointers(void* base, const ResLayout& L, int ns, DResults* D) {
  char* b = (char*)base;
  D->ns = ns;
  D->emitted = (uint8_t*)(b + L.off[0]);
  D->allele_fwd = (int32_t*)(b + L.off[1]); D->other_fwd = (int32_t*)(b + L.off[2]);
  D->allele_cov = (int32_t*)(b + L.off[3]); D->other_cov = (int32_t*)(b + L.off[4]);
  D->total_cov = (int32_t*)(b + L.off[5]);
  D->square_mapq = (double*)(b + L.off[6]);
  D->ref_ll = (double*)(b + L.off[7]); D->allele_ll = (double*)(b + L.off[8]);
  D->other_ll = (double*)(b + L.off[9]); D->nonref_ll = (double*)(b + L.off[10]);
  D->first_is_ref = (uint8_t*)(b + L.off[11]);
  D->copy_number = (int32_t*)(b + L.off[12]);
  D->n_alt = (int32_t*)(b + L.off[13]); D->n_ref = (int32_t*)(b + L.off[14]);
  D->n_other = (int32_t*)(b + L.off[15]);
  D->qual = (double*)(b + L.off[16]);
  D->gq = (int32_t*)(b + L.off[17]);
  D->sb = (int32_t*)(b + L.off[18]);
  D->fisher = (float*)(b + L.off[19]); D->rms_mapq = (float*)(b + L.off[20]);
  D->gl = (double*)(b + L.off[21]); D->nonref_gl = (double*)(b + L.off[22]);
}

int check_results(avo_ctx* ctx, const avo_site_results* o) {
  if (!o) return fail(ctx, AVO_E_INVALID, "results is NULL");
  if (!o->emitted || !o->allele_fwd || !o->other_fwd || !o->allele_cov || !o->other_cov ||
      !o->total_cov || !o->square_mapq || !o->ref_ll || !o->allele_ll || !o->other_ll ||
      !o->nonref_ll || !o->first_is_ref || !o->copy_number || !o->n_alt || !o->n_ref ||
      !o->n_other || !o->qual || !o->gq || !o->sb || !o->fisher || !o->rms_mapq || !o->gl ||
      !o->nonref_gl)
    return fail(ctx, AVO_E_INVALID, "results has NULL columns");
  return AVO_OK;
}

// copy the device result block into the caller's arrays (host or device)
int export_results(avo_ctx* ctx, const DResults& D, int64_t n, int ns, avo_site_results* o) {
  const cudaMemcpyKind kind = (o->mem == AVO_MEM_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  struct { void* dst; const void* src; size_t bytes; } c[] = {
      {o->emitted, D.emitted, (size_t)n}, {o->allele_fwd, D.allele_fwd, (size_t)n * 4},
      {o->other_fwd, D.other_fwd, (size_t)n * 4}, {o->allele_cov, D.allele_cov, (size_t)n * 4},
      {o->other_cov, D.other_cov, (size_t)n * 4}, {o->total_cov, D.total_cov, (size_t)n * 4},
      {o->square_mapq, D.square_mapq, (size_t)n * 8}, {o->ref_ll, D.ref_ll, (size_t)n * ns * 8},
      {o->allele_ll, D.allele_ll, (size_t)n * ns * 8}, {o->other_ll, D.other_ll, (size_t)n * ns * 8},
      {o->nonref_ll, D.nonref_ll, (size_t)n * ns * 8}, {o->first_is_ref, D.first_is_ref, (size_t)n},
      {o->copy_number, D.copy_number, (size_t)n * 4}, {o->n_alt, D.n_alt, (size_t)n * 4},
      {o->n_ref, D.n_ref, (size_t)n * 4}, {o->n_other, D.n_other, (size_t)n * 4},
      {o->qual, D.qual, (size_t)n * 8}, {o->gq, D.gq, (size_t)n * 4}, {o->sb, D.sb, (size_t)n * 16},
      {o->fisher, D.fisher, (size_t)n * 4}, {o->rms_mapq, D.rms_mapq, (size_t)n * 4},
      {o->gl, D.gl, (size_t)n * ns * 8}, {o->nonref_gl, D.nonref_gl, (size_t)n * ns * 8}};
  if (kind == cudaMemcpyDeviceToDevice) {
    ColumnCopies cc{};
    for (auto& x : c) {
      if (x.bytes == 0) continue;
      cc.dst[cc.n] = x.dst; cc.src[cc.n] = x.src; cc.bytes[cc.n] = x.bytes; ++cc.n;
    }
    CK(launch_copy_columns(cc, ctx->stream));
    ctx->timing.n_launches += 1;
    return AVO_OK;
  }
  // the columns live in one device block (res_layout): one DMA, scattered after the sync
  Scatter sc[23];
  int nsc = 0;
  size_t end = 0;
  const char* base = (const char*)D.emitted;
  for (auto& x : c) {
    if (x.bytes == 0) continue;
    const size_t off = (size_t)((const char*)x.src - base);
    sc[nsc++] = Scatter{x.dst, off, x.bytes};
    end = std::max(end, off + x.bytes);
  }
  return d2h_block(ctx, base, end, sc, nsc);
}

int check_params(avo_ctx* ctx, const avo_params* p) {
  if (!p) return fail(ctx, AVO_E_INVALID, "params is NULL");
  if (p->ploidy < 1 || p->ploidy > AVO_MAX_PLOIDY)
    return fail(ctx, AVO_E_UNSUPPORTED, "ploidy %d outside 1..%d", p->ploidy, AVO_MAX_PLOIDY);
  if (p->max_quality < 1 || p->max_quality > 127 || p->max_mapq < 1 || p->max_mapq > 127)
    return fail(ctx, AVO_E_UNSUPPORTED, "max_quality / max_mapq must be in 1..127");
  return AVO_OK;
}

// BiallelicGenotyper.call against a device-resident site table.  `fresh` = the table was just
// produced by discover_internal on the same batch: it is sorted and single-contig by construction
// and the read index in SL_RIDX is current.
int allsites_internal(avo_ctx* ctx, const DReads& R, const BatchStats& hs, const DSites& S, const DCnv& C,
                      const DCallParams& P, const DIndex& X, avo_ref_model_out* ro);

// host_payload == 2: the device lists the payload blocks the joined reads need, host threads copy
// them out of the caller's payload into a dense page-locked array, one DMA brings them back.
int gather_payload(avo_ctx* ctx, const DReads& R, const DSites& S, const void* d_jlist, const int32_t* d_n_joined,
                   const uint8_t** compact, const uint32_t** pbase, const uint32_t** blk0, int* nl) {
  CK(cudaMemcpyAsync(ctx->h_small, d_n_joined, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const int32_t nj = ctx->h_small[0];
  if (nj <= 0) return AVO_OK;
  void *d_pbase, *d_nblk, *d_slot, *d_clo, *d_temp, *d_idx, *d_b0, *d_cmp;
  CKS(get_buf(ctx, SL_G_PBASE, (size_t)(nj + 1) * 4, &d_pbase));
  size_t tb = call_gather_temp_bytes(nj);
  CKS(get_buf(ctx, SL_TEMP, tb, &d_temp));
  CK(launch_call_pairs(d_jlist, nj, (uint32_t*)d_pbase, d_temp, tb, ctx->stream, nl));
  CK(cudaMemcpyAsync(ctx->h_small, (uint32_t*)d_pbase + nj, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const int64_t np = (int64_t)(uint32_t)ctx->h_small[0];      // (read, site) pairs
  CKS(get_buf(ctx, SL_G_NBLK, (size_t)(np + 1) * 4, &d_nblk));
  CKS(get_buf(ctx, SL_G_SLOT, (size_t)(np + 1) * 4, &d_slot));
  CKS(get_buf(ctx, SL_G_CLO, (size_t)(np + 1) * 4, &d_clo));
  tb = call_gather_temp_bytes(np);
  CKS(get_buf(ctx, SL_TEMP, tb, &d_temp));
  CK(launch_call_gather_plan(R, S, d_jlist, nj, (const uint32_t*)d_pbase, np, (uint32_t*)d_nblk, (uint32_t*)d_slot,
                             (uint32_t*)d_clo, d_temp, tb, ctx->stream, nl));
  CK(cudaMemcpyAsync(ctx->h_small, (uint32_t*)d_slot + np, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const size_t total = (size_t)(uint32_t)ctx->h_small[0];
  CKS(get_buf(ctx, SL_G_BLK0, (size_t)(np + 1) * 4, &d_b0));
  CKS(get_buf(ctx, SL_G_IDX, total * 4, &d_idx));
  CKS(get_buf(ctx, SL_G_COMPACT, total * AVO_BLOCK_BYTES, &d_cmp));
  CK(launch_call_blocklist(R, d_jlist, nj, (const uint32_t*)d_pbase, (const uint32_t*)d_nblk, (const uint32_t*)d_slot,
                           (const uint32_t*)d_clo, (uint32_t*)d_idx, (uint32_t*)d_b0, ctx->stream, nl));
  *pbase = (const uint32_t*)d_pbase;
  *blk0 = (const uint32_t*)d_b0;
  *compact = (const uint8_t*)d_cmp;
  if (total == 0) return AVO_OK;
  if (ctx->h_gcap < total) {
    if (ctx->h_gidx) cudaFreeHost(ctx->h_gidx);
    if (ctx->h_gblk) cudaFreeHost(ctx->h_gblk);
    ctx->h_gidx = nullptr; ctx->h_gblk = nullptr; ctx->h_gcap = 0;
    const size_t want = total + total / 4 + 1024;
    if (cudaMallocHost(&ctx->h_gidx, want * 4) != cudaSuccess || cudaMallocHost(&ctx->h_gblk, want * AVO_BLOCK_BYTES) != cudaSuccess) {
      cudaGetLastError();
      return fail(ctx, AVO_E_NOMEM, "page-locked staging for %zu payload blocks", want);
    }
    ctx->h_gcap = want;
  }
  CK(cudaMemcpyAsync(ctx->h_gidx, d_idx, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.d2h_bytes += (int64_t)(total * 4);
  {
    struct Blk { uint64_t a, b; };
    const Blk* src = reinterpret_cast<const Blk*>(ctx->gather_src);
    Blk* dst = reinterpret_cast<Blk*>(ctx->h_gblk);
    const uint32_t* idx = ctx->h_gidx;
    const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)ctx->gather_threads, total / 4096 + 1));
    auto work = [=](int t) {
      const size_t lo = total * (size_t)t / (size_t)T, hi = total * (size_t)(t + 1) / (size_t)T;
      for (size_t k = lo; k < hi; ++k) {
        if (k + 24 < hi) __builtin_prefetch(src + idx[k + 24], 0, 0);
        dst[k] = src[idx[k]];
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
    work(0);
    for (std::thread& th : pool) th.join();
  }
  CK(cudaMemcpyAsync(d_cmp, ctx->h_gblk, total * AVO_BLOCK_BYTES, cudaMemcpyHostToDevice, ctx->stream));
  ctx->timing.h2d_bytes += (int64_t)(total * AVO_BLOCK_BYTES);
  ctx->timing.gathered_blocks = (int64_t)total;
  return AVO_OK;
}

// `fresh`: the table was assembled by discover_enqueue on the same batch earlier in this launch chain:
// sorted and single-contig by construction, sizes in its device header, derived columns and site index
// already there -- nothing here waits for the device; *slots_used (pinned) receives the bucket slots the
// stage needed, to be compared with *slots_cap after the chain's sync.
int call_internal(avo_ctx* ctx, const DReads& R, const BatchStats& hs, const DevSites& T,
                  const avo_cnv* cnv, const avo_params* p, bool fresh, avo_site_results* out,
                  avo_ref_model_out* ref_out = nullptr, DResults* res_out = nullptr, size_t* slots_cap = nullptr) {
  Range r_call(T_CALL_GENOTYPES);
  if (!fresh && T.n == 0)
    return fail(ctx, AVO_E_EMPTY_SITES,
                "empty variant set: the reference throws here (TreeRegionJoin.scala:77)");
  // copy-number map: ploidy range over ALL CNVs (CopyNumberMap.min/maxPloidy), list of this contig
  int minp = p->ploidy, maxp = p->ploidy;
  std::vector<int32_t> cs, ce, cc;
  if (cnv && cnv->n > 0) {
    if (!cnv->contig || !cnv->start || !cnv->end || !cnv->copy_number)
      return fail(ctx, AVO_E_INVALID, "cnv has NULL columns");
    for (int64_t i = 0; i < cnv->n; ++i) {
      minp = std::min(minp, cnv->copy_number[i]);
      maxp = std::max(maxp, cnv->copy_number[i]);
      if (cnv->contig[i] == R.contig) {
        cs.push_back(cnv->start[i]); ce.push_back(cnv->end[i]); cc.push_back(cnv->copy_number[i]);
      }
    }
  }
  if (minp < 1 || maxp > AVO_MAX_PLOIDY)
    return fail(ctx, AVO_E_UNSUPPORTED, "copy numbers must stay within 1..%d", AVO_MAX_PLOIDY);
  const int ns = maxp + 1;
  if (out->n_states != ns)
    return fail(ctx, AVO_E_INVALID, "results->n_states must be max ploidy + 1 = %d", ns);
  if (!fresh && out->n_sites != T.n) return fail(ctx, AVO_E_INVALID, "results->n_sites != number of sites");
  // rows the stage may touch: the table's (a fresh table: its capacity, bounded by the caller's rows)
  const int32_t rows = fresh ? (int32_t)std::min<int64_t>(T.cap, out->mem == AVO_MEM_DEVICE ? out->n_sites : T.cap) : T.n;

  DCnv C{};
  C.n = (int32_t)cs.size();
  if (C.n) {
    void* d;
    CKS(get_buf(ctx, SL_CNV, (size_t)C.n * 12, &d));
    int32_t* b = (int32_t*)d;
    CK(cudaMemcpyAsync(b, cs.data(), (size_t)C.n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(b + C.n, ce.data(), (size_t)C.n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(b + 2 * C.n, cc.data(), (size_t)C.n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    C.start = b; C.end = b + C.n; C.cn = b + 2 * C.n;
  }

  DCallParams P{};
  P.base_ploidy = p->ploidy; P.min_ploidy = minp; P.max_ploidy = maxp; P.ns = ns;
  P.max_quality = p->max_quality; P.max_mapq = p->max_mapq;
  CKS(ensure_lut(ctx, minp, maxp, p->max_quality, p->max_mapq, &P.lut));
  P.lnk = ctx->d_lnk; P.lnfact = ctx->d_lnfact; P.lnk_n = LNK_N;
  P.ten_div_log10 = 10.0 / std::log(10.0);
  P.log10 = std::log(10.0);

  void *d_small, *d_temp;
  CKS(get_buf(ctx, SL_SMALL, SW_COUNT_ * 4, &d_small));
  int32_t* small = (int32_t*)d_small;
  DSites S{};
  S.n = rows;
  S.contig = T.contig; S.start = T.start; S.ref_len = T.ref_len; S.alt_len = T.alt_len;
  S.allele_off = T.allele_off; S.alleles4 = T.alleles4;
  const int32_t nb = index_bins(hs);
  void *d_ridx, *d_sidx;
  CKS(get_buf(ctx, SL_RIDX, (size_t)(nb + 1) * 4, &d_ridx));
  DIndex X{};
  if (fresh) {
    S.hdr = T.hdr; S.cap = rows; S.c_lo = 0; S.c_hi = rows; S.midpoint = 1;
    S.end = T.end; S.pme = T.pme;
    X = DIndex{hs.min_start, nb, (const int32_t*)d_ridx, T.sidx, &T.hdr->s_begin};
  } else {
    // derived site columns + sortedness + contig range
    void *d_end, *d_pme;
    CKS(get_buf(ctx, SL_S_END, (size_t)T.n * 4, &d_end));
    CKS(get_buf(ctx, SL_S_PME, (size_t)T.n * 4, &d_pme));
    const size_t tb = site_setup_temp_bytes(T.n);
    CKS(get_buf(ctx, SL_TEMP, tb, &d_temp));
    CK(launch_site_setup(T.n, T.contig, T.start, T.ref_len, (int32_t*)d_end, (int32_t*)d_pme, small + SW_UNSORTED,
                         d_temp, tb, ctx->stream));
    ctx->timing.n_launches += 2;
    CK(launch_contig_range(T.n, T.contig, R.contig, small + SW_CRANGE, ctx->stream));
    ctx->timing.n_launches += 1;
    CK(cudaMemcpyAsync(ctx->h_small, small, 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_small[0]) return fail(ctx, AVO_E_UNSORTED, "sites are not sorted by (contig, start, end)");
    S.c_lo = ctx->h_small[1]; S.c_hi = ctx->h_small[2];
    S.end = (const int32_t*)d_end; S.pme = (const int32_t*)d_pme;
    int mp = 1;                                                  // TreeRegionJoin.scala:66-72
    while (!(2 * (int64_t)mp >= (int64_t)S.n)) mp *= 2;
    S.midpoint = mp;
    CKS(get_buf(ctx, SL_SIDX, (size_t)(nb + 1) * 4, &d_sidx));
    CK(launch_build_index(R, S, hs, (int32_t*)d_ridx, (int32_t*)d_sidx, small + SW_SRANGE, ctx->stream));
    ctx->timing.n_launches += 1;
    X = DIndex{hs.min_start, nb, (const int32_t*)d_ridx, (const int32_t*)d_sidx, small + SW_SRANGE};
  }

  // results: straight into the caller's device arrays (fresh tables), else one block holding every column
  DResults D;
  void* d_res = nullptr;
  const ResLayout L = res_layout(rows, ns);
  if (fresh && out->mem == AVO_MEM_DEVICE) {
    D.ns = ns;
    D.emitted = out->emitted; D.allele_fwd = out->allele_fwd; D.other_fwd = out->other_fwd;
    D.allele_cov = out->allele_cov; D.other_cov = out->other_cov; D.total_cov = out->total_cov;
    D.square_mapq = out->square_mapq; D.ref_ll = out->ref_ll; D.allele_ll = out->allele_ll;
    D.other_ll = out->other_ll; D.nonref_ll = out->nonref_ll; D.first_is_ref = out->first_is_ref;
    D.copy_number = out->copy_number; D.n_alt = out->n_alt; D.n_ref = out->n_ref; D.n_other = out->n_other;
    D.qual = out->qual; D.gq = out->gq; D.sb = out->sb; D.fisher = out->fisher; D.rms_mapq = out->rms_mapq;
    D.gl = out->gl; D.nonref_gl = out->nonref_gl;
  } else {
    CKS(get_buf(ctx, SL_RES, L.total, &d_res));
    // a multi-contig table has rows this batch never touches; a fresh table's rows are all written
    if (!fresh) CK(cudaMemsetAsync(d_res, 0, L.total, ctx->stream));
    res_pointers(d_res, L, ns, &D);
  }
  if (res_out) *res_out = D;

  // per-site buckets: one 32-bit slot per read that can overlap the site
  void *d_rbase, *d_blen, *d_boff, *d_buckets;
  CKS(get_buf(ctx, SL_C_RBASE, (size_t)(rows + 1) * 4, &d_rbase));
  CKS(get_buf(ctx, SL_C_BLEN, (size_t)(rows + 1) * 4, &d_blen));
  CKS(get_buf(ctx, SL_C_BOFF, (size_t)(rows + 1) * 8, &d_boff));
  const size_t tb2 = call_scan_temp_bytes(rows);
  CKS(get_buf(ctx, SL_TEMP, tb2, &d_temp));
  CK(cudaEventRecord(ctx->ev[4], ctx->stream));
  int nl = 0;
  const bool work = R.n > 0 && S.c_hi > S.c_lo;
  if (work)
    CK(launch_call_buckets(R, S, X, hs, (int32_t*)d_rbase, (uint32_t*)d_blen, (uint64_t*)d_boff, d_temp, tb2,
                           ctx->stream, &nl));
  void* d_jlist = nullptr;                 // the list of joined reads: only the host-gathered route builds one
  if (ctx->gather_src)
    CKS(get_buf(ctx, SL_AS_JR, (size_t)(R.n + 1) * 12 /* + 1: the pair scan loads one element past the list */, &d_jlist));
  // The number of bucket slots is known on the device.  Instead of waiting for it, the stage runs
  // with the previous call's count (+ 1/8) as capacity, the kernels never write past it, and the
  // count is checked with the results; a stage that ran out of slots is rerun with the exact count.
  uint64_t* h_slots = reinterpret_cast<uint64_t*>(ctx->h_small + 48);
  for (int attempt = 0; attempt < 2; ++attempt) {
    const bool speculative = work && !ctx->gather_src && (fresh || (attempt == 0 && ctx->slots_guess > 0));
    size_t n_slots = 0;
    if (work) {
      CK(cudaMemcpyAsync(h_slots, (uint64_t*)d_boff + S.c_hi, 8, cudaMemcpyDeviceToHost, ctx->stream));
      if (speculative) {
        if (ctx->slots_guess == 0) ctx->slots_guess = (size_t)R.n / 2 + (1u << 20);   // first call of the context
        n_slots = ctx->slots_guess;
      } else {
        CK(cudaStreamSynchronize(ctx->stream));
        n_slots = (size_t)*h_slots;
        ctx->slots_guess = n_slots + n_slots / 8 + 1024;
      }
    }
    CKS(get_buf(ctx, SL_C_BUCKETS, n_slots * 4, &d_buckets));
    ctx->pairs_timed = false;
    if (ctx->gather_src) {
      // host-gathered payload: the joined reads are listed first (the host copies the blocks they need)
      nvtxRangePushA(T_TREE_JOIN);
      CK(launch_call_join(R, S, X, hs.max_span, (uint32_t*)d_buckets, n_slots, d_jlist, small + SW_JOIN, ctx->num_sms, ctx->stream, &nl));
      const uint8_t* d_compact = nullptr;
      const uint32_t *d_pbase = nullptr, *d_blk0 = nullptr;
      nvtxRangePop();
      if (work) CKS(gather_payload(ctx, R, S, d_jlist, small + SW_JOIN, &d_compact, &d_pbase, &d_blk0, &nl));
      Range r_obs(T_OBSERVE_READS " / " T_EMIT_GENOTYPES);
      CK(launch_call_observe(R, S, C, P, D, (const int32_t*)d_rbase, (const uint32_t*)d_blen, (const uint64_t*)d_boff,
                             (uint32_t*)d_buckets, n_slots, d_jlist, small + SW_JOIN, d_compact, d_pbase, d_blk0, ctx->num_sms,
                             ctx->stream, &nl));
    } else {
      // device-resident payload: join and observation are one pass over the bucket slots
      Range r_obs(T_TREE_JOIN " / " T_OBSERVE_READS " / " T_EMIT_GENOTYPES);
      CK(launch_call_pairs_reduce(R, S, X, C, P, D, (const int32_t*)d_rbase, (const uint32_t*)d_blen, (const uint64_t*)d_boff,
                                  (uint32_t*)d_buckets, n_slots, ctx->num_sms, ctx->stream, &nl, ctx->ev[8], ctx->ev[9]));
      ctx->pairs_timed = work;
    }
    CK(cudaEventRecord(ctx->ev[5], ctx->stream));
    if (fresh) {                            // the caller checks *h_slots against *slots_cap after its sync
      if (slots_cap) *slots_cap = work ? n_slots : SIZE_MAX;
      break;
    }
    const size_t mark = ctx->pending.size(), used = ctx->h_stage_used;
    CKS(export_results(ctx, D, T.n, ns, out));
    if (!speculative) break;
    CK(cudaStreamSynchronize(ctx->stream));
    const size_t actual = (size_t)*h_slots;
    ctx->slots_guess = actual + actual / 8 + 1024;
    if (actual <= n_slots) break;
    ctx->pending.resize(mark);              // too few slots: forget the exported rows and run again
    ctx->h_stage_used = used;
    CK(cudaMemsetAsync(d_res, 0, L.total, ctx->stream));
  }
  ctx->timing.n_launches += nl;
  ctx->timing.n_tile_launches += nl;
  if (ref_out) CKS(allsites_internal(ctx, R, hs, S, C, P, X, ref_out));
  return AVO_OK;
}

// The reference-model rows of scoreAllSites (avo_allsites.cu) for the batch and site table of a
// call_internal in progress.
int allsites_internal(avo_ctx* ctx, const DReads& R, const BatchStats& hs, const DSites& S, const DCnv& C,
                      const DCallParams& P, const DIndex& X, avo_ref_model_out* ro) {
  avo_site_results* rows = &ro->rows;
  if (rows->mem != AVO_MEM_HOST && rows->mem != AVO_MEM_DEVICE) return fail(ctx, AVO_E_INVALID, "bad ref_model->rows.mem");
  if (rows->n_states != P.ns) return fail(ctx, AVO_E_INVALID, "ref_model->rows.n_states must be %d", P.ns);
  ro->n = 0;
  if (R.n == 0) return AVO_OK;
  AllSitesPlan A{};
  A.p0 = (int64_t)hs.min_start - 1;
  A.nbits = ((int64_t)hs.max_end - A.p0 + 255) / 256 * 256;
  if (A.nbits <= 0) return AVO_OK;
  A.n_win = (int32_t)(A.nbits / 256);
  void* q;
  CKS(get_buf(ctx, SL_AS_FLAG, (size_t)(R.n + 1) * 4, &q)); A.flag = (int32_t*)q;
  CKS(get_buf(ctx, SL_AS_JPOS, (size_t)(R.n + 1) * 4, &q)); A.jpos = (int32_t*)q;
  CKS(get_buf(ctx, SL_AS_JR, (size_t)R.n * 12, &q)); A.jr = q;
  CKS(get_buf(ctx, SL_AS_JIDX, (size_t)(A.nbits / 32 + 1) * 4, &q)); A.jfine = (int32_t*)q;
  CKS(get_buf(ctx, SL_AS_JCOORD, (size_t)R.n * 8, &q)); A.jcoord = (int2*)q;
  CKS(get_buf(ctx, SL_AS_COVER, (size_t)A.nbits / 8, &q)); A.cover = (uint32_t*)q;
  CKS(get_buf(ctx, SL_AS_SCOVER, (size_t)A.nbits / 8, &q)); A.site_cover = (uint32_t*)q;
  CKS(get_buf(ctx, SL_AS_WCOUNT, (size_t)(A.n_win + 1) * 4, &q)); A.win_count = (int32_t*)q;
  CKS(get_buf(ctx, SL_AS_WBASE, (size_t)(A.n_win + 1) * 4, &q)); A.win_base = (int32_t*)q;
  const size_t tb = allsites_scan_temp_bytes(std::max<int64_t>(R.n + 1, A.n_win + 1));
  void* temp;
  CKS(get_buf(ctx, SL_TEMP, tb, &temp));
  int nl = 0;
  CK(launch_allsites_plan(R, S, X, A, temp, tb, ctx->stream, &nl));
  CK(cudaMemcpyAsync(ctx->h_small, A.jpos + R.n, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ctx->h_small + 1, A.win_base + A.n_win, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const int32_t n_joined = ctx->h_small[0];
  const int64_t total = ctx->h_small[1];
  ro->n = total;
  ctx->timing.n_launches += nl;
  if (total > ro->capacity)
    return fail(ctx, AVO_E_CAPACITY, "ref_model too small: need %lld rows", (long long)total);
  if (total == 0) return AVO_OK;
  if (!ro->start) return fail(ctx, AVO_E_INVALID, "ref_model->start is NULL");
  rows->n_sites = total;
  CKS(check_results(ctx, rows));
  nl = 0;
  if (rows->mem == AVO_MEM_DEVICE) {
    // device output: the kernel writes every column of every row straight into the caller's arrays
    DResults D;
    D.ns = P.ns;
    D.emitted = rows->emitted; D.allele_fwd = rows->allele_fwd; D.other_fwd = rows->other_fwd;
    D.allele_cov = rows->allele_cov; D.other_cov = rows->other_cov; D.total_cov = rows->total_cov;
    D.square_mapq = rows->square_mapq; D.ref_ll = rows->ref_ll; D.allele_ll = rows->allele_ll;
    D.other_ll = rows->other_ll; D.nonref_ll = rows->nonref_ll; D.first_is_ref = rows->first_is_ref;
    D.copy_number = rows->copy_number; D.n_alt = rows->n_alt; D.n_ref = rows->n_ref; D.n_other = rows->n_other;
    D.qual = rows->qual; D.gq = rows->gq; D.sb = rows->sb; D.fisher = rows->fisher; D.rms_mapq = rows->rms_mapq;
    D.gl = rows->gl; D.nonref_gl = rows->nonref_gl;
    CK(launch_allsites_tiles(R, S, C, P, D, ro->start, X, A, n_joined, hs.max_span, ctx->stream, &nl));
    ctx->timing.n_launches += nl;
    return AVO_OK;
  }
  const ResLayout L = res_layout(total, P.ns);
  void *d_res, *d_start;
  CKS(get_buf(ctx, SL_AS_RES, L.total, &d_res));
  CKS(get_buf(ctx, SL_AS_START, (size_t)total * 4, &d_start));
  DResults D;
  res_pointers(d_res, L, P.ns, &D);
  CK(launch_allsites_tiles(R, S, C, P, D, (int32_t*)d_start, X, A, n_joined, hs.max_span, ctx->stream, &nl));
  ctx->timing.n_launches += nl;
  CKS(export_results(ctx, D, total, P.ns, rows));
  CK(cudaMemcpyAsync(ro->start, d_start, (size_t)total * 4, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->timing.d2h_bytes += total * 4;
  return AVO_OK;
}

// ---- DiscoverVariants on a device batch, as one launch chain with no host round trip ---------------
// index -> tile kernel -> indel grouping -> k_assemble_sites.  Every size the kernels need lives in
// device memory; the host supplies CAPACITIES (the previous call's counts plus a margin, or the
// caller's arrays) and reads the counts back when the whole chain -- for avo_discover_and_call that
// includes the call stage -- has been queued.  discover_check then says whether a capacity was too small.
struct DiscCaps { int64_t ucap = 0, lcap = 0, tguess = 0, site_cap = 0, nib_cap = 0; };

void initial_caps(avo_ctx* ctx, const DReads& R, DiscCaps* c) {
  c->ucap = ctx->site_guess > 0 ? ctx->site_guess : std::max<int64_t>(1 << 16, R.n / 8);
  c->lcap = std::max<int64_t>(1 << 16, std::max<int64_t>(R.n / 8, ctx->indel_guess + ctx->indel_guess / 8 + 1024));
  c->tguess = std::max<int64_t>(1 << 15, ctx->indel_guess + ctx->indel_guess / 4);
  c->site_cap = c->ucap;
  c->nib_cap = ctx->nib_guess > 0 ? ctx->nib_guess : 4 * c->ucap;
}

// dst: where the table goes.  sites_out in device memory with all columns -> the caller's arrays, else
// context buffers of the guessed capacities.
int discover_enqueue(avo_ctx* ctx, const DReads& R, const BatchStats& hs, bool verify, const avo_params* p,
                     const DiscCaps& c, const avo_sites_out* o, DevSites* T) {
  void *d_small, *d_ridx, *d_sidx, *q;
  CKS(get_buf(ctx, SL_SMALL, SW_COUNT_ * 4, &d_small));
  int32_t* small = (int32_t*)d_small;
  const int32_t nb = index_bins(hs);
  CKS(get_buf(ctx, SL_RIDX, (size_t)(nb + 1) * 4, &d_ridx));
  CKS(get_buf(ctx, SL_SIDX, (size_t)(nb + 1) * 4, &d_sidx));
  DSites none{};
  DIndex X{hs.min_start, nb, (const int32_t*)d_ridx, nullptr, nullptr};

  DDiscoverOut U{};
  DIndelList Lst{};
  const int64_t ucap = c.ucap, lcap = c.lcap;
  U.capacity = (int32_t)std::min<int64_t>(ucap, INT32_MAX - 1);
  CKS(get_buf(ctx, SL_U_START, (size_t)ucap * 4, &q)); U.start = (int32_t*)q;
  CKS(get_buf(ctx, SL_U_REFLEN, (size_t)ucap * 4, &q)); U.ref_len = (int32_t*)q;
  CKS(get_buf(ctx, SL_U_ALTLEN, (size_t)ucap * 4, &q)); U.alt_len = (int32_t*)q;
  CKS(get_buf(ctx, SL_U_COUNT, (size_t)ucap * 4, &q)); U.count = (int32_t*)q;
  CKS(get_buf(ctx, SL_U_SRC, (size_t)ucap * 8, &q)); U.src = (uint64_t*)q;
  CKS(get_buf(ctx, SL_U_NEXT, (size_t)ucap * 4, &q)); U.next = (int32_t*)q;
  U.n = small + SW_NOUT;
  U.win0 = hs.min_start;
  Lst.capacity = (int32_t)std::min<int64_t>(lcap, INT32_MAX - 1);
  CKS(get_buf(ctx, SL_L_POS, (size_t)lcap * 4, &q)); Lst.pos = (int32_t*)q;
  CKS(get_buf(ctx, SL_L_REFLEN, (size_t)lcap * 4, &q)); Lst.ref_len = (int32_t*)q;
  CKS(get_buf(ctx, SL_L_ALTLEN, (size_t)lcap * 4, &q)); Lst.alt_len = (int32_t*)q;
  CKS(get_buf(ctx, SL_L_NIBS, (size_t)lcap * 4, &q)); Lst.nibs = (uint32_t*)q;
  CKS(get_buf(ctx, SL_L_SRC, (size_t)lcap * 4, &q)); Lst.src = (uint32_t*)q;
  CKS(get_buf(ctx, SL_L_BLK, (size_t)lcap * 4, &q)); Lst.blk = (uint32_t*)q;
  CKS(get_buf(ctx, SL_L_HASH, (size_t)lcap * 4, &q)); Lst.hash = (uint32_t*)q;
  Lst.n = small + SW_NINDEL;
  uint32_t tsize = 1024;
  while ((int64_t)tsize < 2 * c.tguess) tsize <<= 1;
  // one arena, one memset: the tile buckets, the indel hash table and its counts
  const int32_t n_tiles = discover_tiles(hs);
  const size_t a_tiles = align_up((size_t)n_tiles * sizeof(TileBucket), 256), a_table = (size_t)tsize * 8, a_tcount = (size_t)tsize * 4;
  CKS(get_buf(ctx, SL_ARENA, a_tiles + a_table + a_tcount, &q));
  char* arena = (char*)q;
  U.tiles = (TileBucket*)arena;
  unsigned long long* d_table = (unsigned long long*)(arena + a_tiles);
  int32_t* d_tcount = (int32_t*)(arena + a_tiles + a_table);
  CK(cudaMemsetAsync(arena, 0, a_tiles + a_table + a_tcount, ctx->stream));
  CK(cudaMemsetAsync(small + SW_TILE, 0, (SW_COUNT_ - SW_TILE) * sizeof(int32_t), ctx->stream));

  // where the table goes
  DSiteTableOut D{};
  const bool to_caller = o && o->mem == AVO_MEM_DEVICE && o->start && o->ref_len && o->alt_len && o->allele_off && o->alleles4 &&
                         o->capacity > 0;
  if (to_caller) {
    D.cap = (int32_t)std::min<int64_t>(o->capacity, INT32_MAX - 1);
    D.nib_cap = (uint32_t)std::min<int64_t>(std::max<int64_t>(o->allele_capacity, 0), UINT32_MAX);
    D.start = o->start; D.ref_len = o->ref_len; D.alt_len = o->alt_len; D.allele_off = o->allele_off; D.alleles4 = o->alleles4;
    D.count = o->count;
  } else {
    D.cap = (int32_t)std::min<int64_t>(c.site_cap, INT32_MAX - 1);
    D.nib_cap = (uint32_t)std::min<int64_t>(c.nib_cap, UINT32_MAX);
    CKS(get_buf(ctx, SL_S_START, (size_t)D.cap * 4, &q)); D.start = (int32_t*)q;
    CKS(get_buf(ctx, SL_S_REFLEN, (size_t)D.cap * 4, &q)); D.ref_len = (int32_t*)q;
    CKS(get_buf(ctx, SL_S_ALTLEN, (size_t)D.cap * 4, &q)); D.alt_len = (int32_t*)q;
    CKS(get_buf(ctx, SL_S_AOFF, (size_t)(D.cap + 1) * 4, &q)); D.allele_off = (uint32_t*)q;
    CKS(get_buf(ctx, SL_S_ALLELES, (size_t)D.nib_cap / 2 + 16, &q)); D.alleles4 = (uint8_t*)q;
  }
  if (!D.count) { CKS(get_buf(ctx, SL_S_COUNT, (size_t)D.cap * 4, &q)); D.count = (int32_t*)q; }
  CKS(get_buf(ctx, SL_S_END, (size_t)D.cap * 4, &q)); D.end = (int32_t*)q;
  CKS(get_buf(ctx, SL_S_PME, (size_t)D.cap * 4, &q)); D.pme = (int32_t*)q;
  D.sidx = (int32_t*)d_sidx;
  D.hdr = (DSiteHdr*)(small + SW_HDR);

  Range r_disc(T_DISCOVERING);
  CK(launch_build_index(R, none, hs, (int32_t*)d_ridx, nullptr, nullptr, ctx->stream));
  ctx->timing.n_launches += 1;
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  int nl = 0;
  if (R.snv_first)
    CK(launch_discover_events(R, p->min_phred_discover, p->min_obs_discover, X, hs, U, Lst, small + SW_TILE, small + SW_FLAGS,
                              ctx->num_sms, &ctx->ev_blocks_per_sm, ctx->stream, &nl));
  else
    CK(launch_discover_tiles(R, p->min_phred_discover, p->min_obs_discover, X, hs, verify ? 1 : 0, U, Lst,
                             small + SW_TILE, small + SW_FLAGS, ctx->num_sms, &ctx->disc_blocks_per_sm, ctx->stream, &nl));
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->timing.n_tile_launches += nl;
  CK(launch_indel_group(R, Lst, d_table, d_tcount, tsize, p->min_obs_discover, U, small + SW_FLAGS, ctx->num_sms,
                        ctx->stream, &nl));
  {
    Range r_tree(T_BUILDING_TREES);
    CK(launch_assemble_sites(R, Lst, U, hs, nb, D, ctx->stream, &nl));
  }
  ctx->timing.n_launches += nl;

  T->n = 0; T->n_nibbles = 0; T->contig = nullptr;
  T->start = D.start; T->ref_len = D.ref_len; T->alt_len = D.alt_len; T->allele_off = D.allele_off;
  T->alleles4 = D.alleles4; T->count = D.count;
  T->hdr = D.hdr; T->cap = D.cap; T->nib_cap = D.nib_cap; T->end = D.end; T->pme = D.pme; T->sidx = D.sidx;
  T->caller_arrays = to_caller;
  return AVO_OK;
}

// queues the copy of the counters (to be read after the next synchronisation)
int discover_fetch_counts(avo_ctx* ctx) {
  void* d_small;
  CKS(get_buf(ctx, SL_SMALL, SW_COUNT_ * 4, &d_small));
  CK(cudaMemcpyAsync(ctx->h_small, (int32_t*)d_small + SW_TILE, (SW_COUNT_ - SW_TILE) * sizeof(int32_t),
                     cudaMemcpyDeviceToHost, ctx->stream));
  return AVO_OK;
}

// After the sync.  AVO_OK: *again says whether a capacity of `c` was too small (c is then grown).
int discover_check(avo_ctx* ctx, DiscCaps* c, DevSites* T, bool* again) {
  const int32_t* w = ctx->h_small - SW_TILE;          // w[SW_x] = word SW_x of the device block
  const int32_t flags = w[SW_FLAGS];
  const int64_t n_out = w[SW_NOUT], n_indel = w[SW_NINDEL];
  DSiteHdr h;
  memcpy(&h, w + SW_HDR, sizeof h);
  *again = false;
  if (flags & DISC_FLAG_UNSORTED) return fail(ctx, AVO_E_UNSORTED, "reads are not sorted by start");
  if (flags & DISC_FLAG_BAD_STATS)
    return fail(ctx, AVO_E_INVALID, "reads->stats does not bound the batch (max_end / max_span too small)");
  ctx->indel_guess = n_indel;
  if (n_indel > c->lcap) { c->lcap = n_indel + n_indel / 8 + 1024; *again = true; }
  if (flags & DISC_FLAG_TABLE_FULL) { c->tguess = std::max<int64_t>(2 * c->tguess, n_indel); *again = true; }
  // without the table the survivors of the indel pass are unknown: at most n_indel
  const int64_t need_u = n_out + ((flags & DISC_FLAG_TABLE_FULL) ? n_indel : 0);
  if (need_u > c->ucap) { c->ucap = need_u + need_u / 8 + 1024; *again = true; }
  if (!*again) {
    ctx->site_guess = (int64_t)h.n + h.n / 8 + 1024;
    ctx->nib_guess = (int64_t)h.n_nibbles + h.n_nibbles / 8 + 4096;
    if (!T->caller_arrays) {           // the context's own arrays: grow and rerun; the caller's: reported
      if (h.n > c->site_cap) { c->site_cap = ctx->site_guess; *again = true; }
      if ((int64_t)h.n_nibbles > c->nib_cap) { c->nib_cap = ctx->nib_guess; *again = true; }
    }
    T->n = h.n;
    T->n_nibbles = h.n_nibbles;
  }
  return AVO_OK;
}

int export_sites(avo_ctx* ctx, const DevSites& T, avo_sites_out* o) {
  if (!o) return fail(ctx, AVO_E_INVALID, "sites_out is NULL");
  const bool fits = (int64_t)T.n <= o->capacity && (int64_t)T.n_nibbles <= o->allele_capacity;
  o->n = T.n;
  o->n_allele_nibbles = T.n_nibbles;
  if (!fits)
    return fail(ctx, AVO_E_CAPACITY, "sites_out too small: need %d entries, %u allele nibbles", T.n,
                T.n_nibbles);
  if (T.n == 0 || T.caller_arrays) return AVO_OK;
  if (!o->start || !o->ref_len || !o->alt_len || !o->allele_off || !o->alleles4)
    return fail(ctx, AVO_E_INVALID, "sites_out has NULL columns");
  const cudaMemcpyKind kind = (o->mem == AVO_MEM_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  const size_t n = (size_t)T.n;
  if (kind == cudaMemcpyDeviceToDevice) {
    ColumnCopies cc{};
    auto add = [&](void* d, const void* s_, size_t b) { cc.dst[cc.n] = d; cc.src[cc.n] = s_; cc.bytes[cc.n] = b; ++cc.n; };
    add(o->start, T.start, n * 4); add(o->ref_len, T.ref_len, n * 4); add(o->alt_len, T.alt_len, n * 4);
    add(o->allele_off, T.allele_off, (n + 1) * 4); add(o->alleles4, T.alleles4, ((size_t)T.n_nibbles + 1) / 2);
    if (o->count) add(o->count, T.count, n * 4);
    CK(launch_copy_columns(cc, ctx->stream));
    ctx->timing.n_launches += 1;
    return AVO_OK;
  }
  auto one = [&](void* dst, const void* src, size_t bytes) {
    const Scatter sc{dst, 0, bytes};
    return d2h_block(ctx, src, bytes, &sc, 1);
  };
  CKS(one(o->start, T.start, n * 4));
  CKS(one(o->ref_len, T.ref_len, n * 4));
  CKS(one(o->alt_len, T.alt_len, n * 4));
  CKS(one(o->allele_off, T.allele_off, (n + 1) * 4));
  CKS(one(o->alleles4, T.alleles4, ((size_t)T.n_nibbles + 1) / 2));
  if (o->count) CKS(one(o->count, T.count, n * 4));
  return AVO_OK;
}

int finish_call(avo_ctx* ctx, bool did_discover, bool did_call) {
  CK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  scatter_pending(ctx);
  avo_timing& t = ctx->timing;
  t.ms_total = ev_ms(ctx, 0, 7);
  t.ms_h2d = ev_ms(ctx, 0, 1);
  if (did_discover) t.ms_discover = ev_ms(ctx, 1, 6);
  if (did_call) t.ms_observe = did_discover ? ev_ms(ctx, 6, 5) : ev_ms(ctx, 1, 5);
  if (did_discover) t.ms_discover_tiles = ev_ms(ctx, 2, 3);
  if (did_call) t.ms_call_tiles = ev_ms(ctx, 4, 5);
  if (did_call && ctx->pairs_timed) t.ms_call_pairs = ev_ms(ctx, 8, 9);
  t.ms_d2h = did_call ? ev_ms(ctx, 5, 7) : ev_ms(ctx, 6, 7);
  return AVO_OK;
}

}  // namespace

// what avo_shard_discover leaves for avo_shard_call / avo_shard_finish
struct ShardState {
  DReads R{};
  BatchStats hs{};
  DevSites T_loc, T_glob;
  DResults D{};
  DiscCaps caps;
  size_t slots_cap = SIZE_MAX;
  int ns = 0;
  int phase = 0;
  ShardHdr* d_hdrs = nullptr;      // the 2 x world block headers of a step, collected by the unpack kernel
  ShardHdr* h_hdrs = nullptr;      // (pinned) so that they reach the host in ONE copy
};
void free_shard_state(avo_ctx* ctx) {
  if (ctx->shard) {
    if (ctx->shard->d_hdrs) cudaFree(ctx->shard->d_hdrs);
    if (ctx->shard->h_hdrs) cudaFreeHost(ctx->shard->h_hdrs);
  }
  delete ctx->shard;
  ctx->shard = nullptr;
}

// ---- public entry points -------------------------------------------------------------------------
extern "C" {

const char* avo_version(void) { return "avocado_b200 0.1 (sm_100a)"; }

int avo_ctx_create(int device, avo_ctx** out) {
  if (!out) return AVO_E_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || device < 0 || device >= count) { cudaGetLastError(); return AVO_E_CUDA; }
  avo_ctx* ctx = new (std::nothrow) avo_ctx();
  if (!ctx) return AVO_E_NOMEM;
  ctx->device = device;
  auto bail = [&](int code) { avo_ctx_destroy(ctx); return code; };
  if (cudaSetDevice(device) != cudaSuccess) return bail(AVO_E_CUDA);
  // How host threads wait for the device (cudaStreamSynchronize): the CUDA default spins.  With many
  // contexts per process (one per executor thread) AVO_SYNC=blocking|yield frees the cores for the
  // threads that gather payload blocks.  Process-wide (it is a device flag); unset = leave as is.
  if (const char* mode = std::getenv("AVO_SYNC")) {
    const std::string m(mode);
    const unsigned f = m == "blocking" ? cudaDeviceScheduleBlockingSync : m == "yield" ? cudaDeviceScheduleYield
                     : m == "spin" ? cudaDeviceScheduleSpin : cudaDeviceScheduleAuto;
    cudaSetDeviceFlags(f);
    cudaGetLastError();
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(AVO_E_CUDA);
  ctx->num_sms = prop.multiProcessorCount;
  // the kernels gather single 32-byte sectors (one quality / base sector per mismatch or site):
  // ask L2 to fetch 32 B on a miss instead of the default 64 B (a hint; ignored if unsupported)
#ifndef AVO_L2_FETCH
#define AVO_L2_FETCH 32
#endif
  cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, AVO_L2_FETCH);
  cudaGetLastError();
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(AVO_E_CUDA);
  ctx->stream = ctx->own_stream;
  for (auto& ev : ctx->ev)
    if (cudaEventCreate(&ev) != cudaSuccess) return bail(AVO_E_CUDA);
  if (cudaMallocHost(&ctx->h_stats, sizeof(BatchStats)) != cudaSuccess) return bail(AVO_E_CUDA);
  if (cudaMallocHost(&ctx->h_small, 64 * sizeof(int32_t)) != cudaSuccess) return bail(AVO_E_CUDA);
  // ln table (host libm) and logFactorial table (device, reference summation order)
  std::vector<double> lnk(LNK_N);
  lnk[0] = 0.0;
  for (int k = 1; k < LNK_N; ++k) lnk[k] = std::log((double)k);
  if (cudaMalloc(&ctx->d_lnk, LNK_N * sizeof(double)) != cudaSuccess) return bail(AVO_E_NOMEM);
  if (cudaMalloc(&ctx->d_lnfact, LNK_N * sizeof(double)) != cudaSuccess) return bail(AVO_E_NOMEM);
  if (cudaMemcpy(ctx->d_lnk, lnk.data(), LNK_N * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
    return bail(AVO_E_CUDA);
  if (launch_lnfact_table(ctx->d_lnk, ctx->d_lnfact, LNK_N, ctx->stream) != cudaSuccess) return bail(AVO_E_CUDA);
  {
    double lg[AVO_MAX_PLOIDY + 3];
    for (int k = 0; k < AVO_MAX_PLOIDY + 3; ++k) lg[k] = k == 0 ? 0.0 : std::lgamma((double)k);
    if (cudaMalloc(&ctx->d_lgam, sizeof lg) != cudaSuccess) return bail(AVO_E_NOMEM);
    if (cudaMemcpy(ctx->d_lgam, lg, sizeof lg, cudaMemcpyHostToDevice) != cudaSuccess) return bail(AVO_E_CUDA);
  }
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return bail(AVO_E_CUDA);
  *out = ctx;
  return AVO_OK;
}

void avo_ctx_destroy(avo_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (auto& b : ctx->bufs) if (b.p) cudaFree(b.p);
  if (ctx->d_lnk) cudaFree(ctx->d_lnk);
  if (ctx->d_lnfact) cudaFree(ctx->d_lnfact);
  if (ctx->d_lgam) cudaFree(ctx->d_lgam);
  if (ctx->h_stats) cudaFreeHost(ctx->h_stats);
  if (ctx->h_small) cudaFreeHost(ctx->h_small);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->h_gidx) cudaFreeHost(ctx->h_gidx);
  if (ctx->h_gblk) cudaFreeHost(ctx->h_gblk);
  free_shard_state(ctx);
  for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  cudaGetLastError();
  delete ctx;
}

const char* avo_last_error(const avo_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }

int avo_ctx_set_stream(avo_ctx* ctx, void* cuda_stream) {
  if (!ctx) return AVO_E_INVALID;
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return AVO_OK;
}

int avo_get_timing(const avo_ctx* ctx, avo_timing* out) {
  if (!ctx || !out) return AVO_E_INVALID;
  *out = ctx->timing;
  return AVO_OK;
}

int avo_discover(avo_ctx* ctx, const avo_read_batch* reads, const avo_params* params,
                 avo_sites_out* out) {
  CKS(begin_call(ctx));
  CKS(validate_reads(ctx, reads));
  if (!params) return fail(ctx, AVO_E_INVALID, "params is NULL");
  if (!out) return fail(ctx, AVO_E_INVALID, "sites_out is NULL");
  DReads R;
  CKS(upload_reads(ctx, reads, params->host_payload != 1, &R));
  CK(cudaEventRecord(ctx->ev[1], ctx->stream));
  BatchStats hs;
  const bool hinted = stats_from_hints(reads, &hs);
  if (!hinted) CKS(batch_stats(ctx, R, &hs));
  DevSites T;
  DiscCaps caps;
  initial_caps(ctx, R, &caps);
  for (int attempt = 0;; ++attempt) {
    if (attempt > 5) return fail(ctx, AVO_E_NOMEM, "discovery buffers kept overflowing");
    CKS(discover_enqueue(ctx, R, hs, hinted, params, caps, out, &T));
    CKS(discover_fetch_counts(ctx));
    CK(cudaStreamSynchronize(ctx->stream));
    bool again = false;
    CKS(discover_check(ctx, &caps, &T, &again));
    if (!again) break;
  }
  CK(cudaEventRecord(ctx->ev[6], ctx->stream));
  const int rc = export_sites(ctx, T, out);
  CKS(finish_call(ctx, true, false));
  return rc;
}

static int call_entry(avo_ctx* ctx, const avo_read_batch* reads, const avo_sites* sites, const avo_cnv* cnv,
                      const avo_params* params, avo_site_results* out, avo_ref_model_out* ref_model) {
  CKS(begin_call(ctx));
  CKS(validate_reads(ctx, reads));
  CKS(check_params(ctx, params));
  CKS(check_results(ctx, out));
  if (!sites) return fail(ctx, AVO_E_INVALID, "sites is NULL");
  if (sites->n < 0 || sites->n >= INT32_MAX) return fail(ctx, AVO_E_INVALID, "sites->n out of range");
  if (sites->n == 0)
    return fail(ctx, AVO_E_EMPTY_SITES,
                "empty variant set: the reference throws here (TreeRegionJoin.scala:77)");
  if (!sites->start || !sites->ref_len || !sites->alt_len || !sites->allele_off || !sites->alleles4)
    return fail(ctx, AVO_E_INVALID, "sites has NULL columns");
  DReads R;
  CKS(upload_reads(ctx, reads, params->host_payload != 1 && !ref_model, &R, params));
  DevSites T;
  T.n = (int32_t)sites->n;
  T.n_nibbles = (uint32_t)sites->n_allele_nibbles;
  const size_t n = (size_t)sites->n;
  if (sites->contig) CKS(stage_in(ctx, SL_S_CONTIG, sites->contig, n, sites->mem, &T.contig));
  CKS(stage_in(ctx, SL_S_START, sites->start, n, sites->mem, &T.start));
  CKS(stage_in(ctx, SL_S_REFLEN, sites->ref_len, n, sites->mem, &T.ref_len));
  CKS(stage_in(ctx, SL_S_ALTLEN, sites->alt_len, n, sites->mem, &T.alt_len));
  CKS(stage_in(ctx, SL_S_AOFF, sites->allele_off, n + 1, sites->mem, &T.allele_off));
  CKS(stage_in(ctx, SL_S_ALLELES, sites->alleles4, (size_t)(sites->n_allele_nibbles + 1) / 2,
               sites->mem, &T.alleles4));
  CK(cudaEventRecord(ctx->ev[1], ctx->stream));
  BatchStats hs;
  CKS(batch_stats(ctx, R, &hs));
  CKS(call_internal(ctx, R, hs, T, cnv, params, false, out, ref_model));
  CKS(finish_call(ctx, false, true));
  return AVO_OK;
}

int avo_call(avo_ctx* ctx, const avo_read_batch* reads, const avo_sites* sites, const avo_cnv* cnv,
             const avo_params* params, avo_site_results* out) {
  if (params && params->score_all_sites)
    return fail(ctx, AVO_E_INVALID, "score_all_sites needs the reference-model output: call avo_call_all_sites");
  return call_entry(ctx, reads, sites, cnv, params, out, nullptr);
}

int avo_call_all_sites(avo_ctx* ctx, const avo_read_batch* reads, const avo_sites* sites,
                       const avo_cnv* cnv, const avo_params* params, avo_site_results* out,
                       avo_ref_model_out* ref_model) {
  if (!ref_model) return fail(ctx, AVO_E_INVALID, "ref_model is NULL");
  return call_entry(ctx, reads, sites, cnv, params, out, ref_model);
}

int avo_discover_and_call(avo_ctx* ctx, const avo_read_batch* reads, const avo_cnv* cnv,
                          const avo_params* params, avo_sites_out* sites_out,
                          avo_site_results* results) {
  Range r_all(T_DISCOVER_AND_CALL);
  CKS(begin_call(ctx));
  CKS(validate_reads(ctx, reads));
  CKS(check_params(ctx, params));
  if (params->score_all_sites)
    return fail(ctx, AVO_E_INVALID, "score_all_sites: discover with avo_discover, then avo_call_all_sites");
  CKS(check_results(ctx, results));
  if (!sites_out) return fail(ctx, AVO_E_INVALID, "sites_out is NULL");
  if (results->mem != AVO_MEM_HOST && results->mem != AVO_MEM_DEVICE) return fail(ctx, AVO_E_INVALID, "bad results->mem");
  DReads R;
  CKS(upload_reads(ctx, reads, params->host_payload != 1 && !params->score_all_sites, &R, params));
  CK(cudaEventRecord(ctx->ev[1], ctx->stream));
  BatchStats hs;
  const bool hinted = stats_from_hints(reads, &hs);
  if (!hinted) CKS(batch_stats(ctx, R, &hs));
  // One launch chain: discovery, the table, the call stage.  Capacities are the previous call's counts
  // plus a margin; all counts come back with one copy when everything is queued.
  const int64_t res_cap = results->n_sites;
  const bool gathered = ctx->gather_src != nullptr;    // host-gathered payload: the call stage talks to the host anyway
  DevSites T;
  DResults D;
  DiscCaps caps;
  initial_caps(ctx, R, &caps);
  uint64_t* h_slots = reinterpret_cast<uint64_t*>(ctx->h_small + 48);
  for (int attempt = 0;; ++attempt) {
    if (attempt > 6) return fail(ctx, AVO_E_NOMEM, "discovery / call buffers kept overflowing");
    CKS(discover_enqueue(ctx, R, hs, hinted, params, caps, sites_out, &T));
    CK(cudaEventRecord(ctx->ev[6], ctx->stream));
    size_t slots_cap = SIZE_MAX;
    if (gathered) {      // the table's size first (this route synchronises inside the call stage as well)
      CKS(discover_fetch_counts(ctx));
      CK(cudaStreamSynchronize(ctx->stream));
      bool again = false;
      CKS(discover_check(ctx, &caps, &T, &again));
      if (again) continue;
      if (T.n == 0) break;
    }
    *h_slots = 0;
    CKS(call_internal(ctx, R, hs, T, cnv, params, true, results, nullptr, &D, &slots_cap));
    CKS(discover_fetch_counts(ctx));
    CK(cudaStreamSynchronize(ctx->stream));
    bool again = false;
    CKS(discover_check(ctx, &caps, &T, &again));
    if (slots_cap != SIZE_MAX) {
      const size_t actual = (size_t)*h_slots;
      ctx->slots_guess = actual + actual / 8 + 1024;
      if (actual > slots_cap) again = true;
    }
    if (!again) break;
  }
  results->n_sites = T.n;
  int rc = export_sites(ctx, T, sites_out);
  if (rc == AVO_OK && res_cap < T.n)
    rc = fail(ctx, AVO_E_CAPACITY, "results too small: need %d rows", T.n);
  if (rc != AVO_OK) { finish_call(ctx, true, false); return rc; }
  if (T.n == 0) {
    finish_call(ctx, true, false);
    return fail(ctx, AVO_E_EMPTY_SITES,
                "no variants discovered: the reference throws here (TreeRegionJoin.scala:77)");
  }
  if (results->mem == AVO_MEM_HOST) {
    // the rows of every column, packed by one launch, cross PCIe as one copy
    const int ns = D.ns;
    const size_t n = (size_t)T.n;
    struct { void* dst; const void* src; size_t bytes; } c[] = {
        {results->emitted, D.emitted, n}, {results->allele_fwd, D.allele_fwd, n * 4},
        {results->other_fwd, D.other_fwd, n * 4}, {results->allele_cov, D.allele_cov, n * 4},
        {results->other_cov, D.other_cov, n * 4}, {results->total_cov, D.total_cov, n * 4},
        {results->square_mapq, D.square_mapq, n * 8}, {results->ref_ll, D.ref_ll, n * ns * 8},
        {results->allele_ll, D.allele_ll, n * ns * 8}, {results->other_ll, D.other_ll, n * ns * 8},
        {results->nonref_ll, D.nonref_ll, n * ns * 8}, {results->first_is_ref, D.first_is_ref, n},
        {results->copy_number, D.copy_number, n * 4}, {results->n_alt, D.n_alt, n * 4},
        {results->n_ref, D.n_ref, n * 4}, {results->n_other, D.n_other, n * 4},
        {results->qual, D.qual, n * 8}, {results->gq, D.gq, n * 4}, {results->sb, D.sb, n * 16},
        {results->fisher, D.fisher, n * 4}, {results->rms_mapq, D.rms_mapq, n * 4},
        {results->gl, D.gl, n * ns * 8}, {results->nonref_gl, D.nonref_gl, n * ns * 8}};
    size_t total = 0;
    for (auto& x : c) total += align_up(x.bytes, 16);
    void* d_pack;
    CKS(get_buf(ctx, SL_PACK, total, &d_pack));
    ColumnCopies cc{};
    Scatter sc[23];
    size_t off = 0;
    for (auto& x : c) {
      cc.dst[cc.n] = (char*)d_pack + off; cc.src[cc.n] = x.src; cc.bytes[cc.n] = x.bytes;
      sc[cc.n] = Scatter{x.dst, off, x.bytes};
      ++cc.n;
      off += align_up(x.bytes, 16);
    }
    CK(launch_copy_columns(cc, ctx->stream));
    ctx->timing.n_launches += 1;
    CKS(d2h_block(ctx, d_pack, total, sc, cc.n));
  }
  CKS(finish_call(ctx, true, true));
  return AVO_OK;
}

int avo_merge_meta(avo_ctx* ctx, int32_t n_parts, const avo_read_meta* const* part_meta, const int64_t* part_reads,
                   const uint32_t* block_base, const uint32_t* op_base, const uint32_t* refb_base, avo_read_meta* out_meta) {
  if (!ctx) return AVO_E_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (n_parts < 1 || n_parts > AVO_MAX_MERGE || !part_meta || !part_reads || !block_base || !op_base || !refb_base || !out_meta)
    return fail(ctx, AVO_E_INVALID, "avo_merge_meta: 1..%d parts, no NULL arguments", AVO_MAX_MERGE);
  MergeParts M{};
  M.k = n_parts;
  int64_t total = 0;
  for (int k = 0; k < n_parts; ++k) {
    if (part_reads[k] < 0 || (part_reads[k] > 0 && !part_meta[k])) return fail(ctx, AVO_E_INVALID, "avo_merge_meta: bad part %d", k);
    M.meta[k] = part_meta[k]; M.n[k] = part_reads[k];
    M.block_base[k] = block_base[k]; M.op_base[k] = op_base[k]; M.refb_base[k] = refb_base[k];
    total += part_reads[k];
  }
  if (total >= (int64_t)INT32_MAX) return fail(ctx, AVO_E_INVALID, "the union must stay below 2^31 reads");
  CK(launch_merge_meta(M, out_meta, ctx->stream));
  ctx->timing.n_launches += 1;
  return AVO_OK;
}

// ---- region-sharded discoverAndCall: three phases around the framework's two all-gathers ------------
static int check_plan(avo_ctx* ctx, const avo_shard_plan* pl) {
  if (!pl) return fail(ctx, AVO_E_INVALID, "shard plan is NULL");
  if (pl->world < 1 || pl->world > AVO_MAX_SHARDS || pl->rank < 0 || pl->rank >= pl->world)
    return fail(ctx, AVO_E_INVALID, "shard plan: bad rank / world (at most %d ranks)", AVO_MAX_SHARDS);
  if (pl->site_cap < 1 || pl->site_cap >= INT32_MAX / 8 || pl->row_cap < 1 || pl->row_cap >= INT32_MAX / 8 ||
      pl->allele_cap < 16 || (pl->allele_cap & 15) || pl->allele_cap >= INT32_MAX)
    return fail(ctx, AVO_E_INVALID, "shard plan: bad capacities (allele_cap must be a multiple of 16)");
  if (pl->own_lo > pl->own_hi) return fail(ctx, AVO_E_INVALID, "shard plan: own_lo > own_hi");
  return AVO_OK;
}

int64_t avo_shard_table_block_bytes(const avo_shard_plan* pl) {
  return pl ? (int64_t)shard_table_block_bytes(pl->site_cap, pl->allele_cap) : 0;
}
int64_t avo_shard_rows_block_bytes(const avo_shard_plan* pl) {
  return pl ? (int64_t)shard_rows_block_bytes(pl->row_cap, pl->n_states) : 0;
}

int avo_shard_discover(avo_ctx* ctx, const avo_read_batch* reads, const avo_params* params,
                       const avo_shard_plan* pl, void* table_block) {
  CKS(begin_call(ctx));
  CKS(validate_reads(ctx, reads));
  CKS(check_params(ctx, params));
  CKS(check_plan(ctx, pl));
  if (!table_block || ((uintptr_t)table_block & 15u)) return fail(ctx, AVO_E_INVALID, "table_block must be a 16-byte aligned device buffer");
  if (!ctx->shard) ctx->shard = new (std::nothrow) ShardState();
  if (!ctx->shard) return fail(ctx, AVO_E_NOMEM, "out of host memory");
  ShardState& st = *ctx->shard;
  st.phase = 0;
  if (!st.d_hdrs) CK(cudaMalloc(&st.d_hdrs, 2 * AVO_MAX_SHARDS * sizeof(ShardHdr)));
  if (!st.h_hdrs) CK(cudaHostAlloc(&st.h_hdrs, 2 * AVO_MAX_SHARDS * sizeof(ShardHdr), cudaHostAllocDefault));
  CKS(upload_reads(ctx, reads, params->host_payload != 1, &st.R, params));
  CK(cudaEventRecord(ctx->ev[1], ctx->stream));
  const bool hinted = stats_from_hints(reads, &st.hs);
  if (!hinted) CKS(batch_stats(ctx, st.R, &st.hs));
  initial_caps(ctx, st.R, &st.caps);
  CKS(discover_enqueue(ctx, st.R, st.hs, hinted, params, st.caps, nullptr, &st.T_loc));
  void* d_small;
  CKS(get_buf(ctx, SL_SMALL, SW_COUNT_ * 4, &d_small));
  ShardTableIn T{st.T_loc.hdr, st.T_loc.cap, st.T_loc.nib_cap, st.T_loc.start, st.T_loc.ref_len, st.T_loc.alt_len,
                 st.T_loc.count, st.T_loc.pme, st.T_loc.allele_off, st.T_loc.alleles4};
  ShardLocalCaps C{SW_FLAGS, SW_NOUT, SW_NINDEL, (int32_t)std::min<int64_t>(st.caps.ucap, INT32_MAX - 1),
                   (int32_t)std::min<int64_t>(st.caps.lcap, INT32_MAX - 1)};
  CK(launch_shard_owned(T, pl->own_lo, pl->own_hi, (const int32_t*)d_small, C, table_block, (int32_t)pl->site_cap,
                        (int32_t)pl->allele_cap, ctx->num_sms, ctx->stream));
  ctx->timing.n_launches += 1;
  CK(cudaEventRecord(ctx->ev[6], ctx->stream));
  st.phase = 1;
  return AVO_OK;
}

int avo_shard_call(avo_ctx* ctx, const avo_cnv* cnv, const avo_params* params, const avo_shard_plan* pl,
                   const void* gathered_tables, avo_sites_out* gs, void* rows_block) {
  return avo_shard_call_direct(ctx, cnv, params, pl, gathered_tables, gs, rows_block, nullptr);
}

int avo_shard_call_direct(avo_ctx* ctx, const avo_cnv* cnv, const avo_params* params, const avo_shard_plan* pl,
                          const void* gathered_tables, avo_sites_out* gs, void* rows_block,
                          const avo_site_results* root_rows) {
  if (!ctx) return AVO_E_INVALID;
  CK(cudaSetDevice(ctx->device));
  CKS(check_params(ctx, params));
  CKS(check_plan(ctx, pl));
  if (!ctx->shard || ctx->shard->phase != 1) return fail(ctx, AVO_E_INVALID, "avo_shard_call follows avo_shard_discover");
  if (!gathered_tables || !rows_block || ((uintptr_t)rows_block & 15u)) return fail(ctx, AVO_E_INVALID, "NULL / unaligned block");
  if (!gs || gs->mem != AVO_MEM_DEVICE || !gs->start || !gs->ref_len || !gs->alt_len || !gs->allele_off || !gs->alleles4 ||
      gs->capacity < 1)
    return fail(ctx, AVO_E_INVALID, "global_sites must be device arrays with a capacity");
  ShardState& st = *ctx->shard;
  void *d_small, *q;
  CKS(get_buf(ctx, SL_SMALL, SW_COUNT_ * 4, &d_small));
  int32_t* small = (int32_t*)d_small;
  const int32_t nb = index_bins(st.hs);
  DSiteTableOut G{};
  G.cap = (int32_t)std::min<int64_t>(gs->capacity, INT32_MAX - 1);
  G.nib_cap = (uint32_t)std::min<int64_t>(std::max<int64_t>(gs->allele_capacity, 0), UINT32_MAX);
  G.start = gs->start; G.ref_len = gs->ref_len; G.alt_len = gs->alt_len; G.allele_off = gs->allele_off; G.alleles4 = gs->alleles4;
  G.count = gs->count;
  CKS(get_buf(ctx, SL_G_END, (size_t)G.cap * 4, &q)); G.end = (int32_t*)q;
  CKS(get_buf(ctx, SL_G_PME, (size_t)G.cap * 4, &q)); G.pme = (int32_t*)q;
  CKS(get_buf(ctx, SL_G_SIDX, (size_t)(nb + 1) * 4, &q)); G.sidx = (int32_t*)q;
  G.hdr = (DSiteHdr*)(small + SW_HDR2);
  if (root_rows) {
    CKS(check_results(ctx, root_rows));
    if (root_rows->mem != AVO_MEM_DEVICE) return fail(ctx, AVO_E_INVALID, "root_rows must be device arrays (local or peer-mapped)");
  }
  CK(launch_shard_concat(gathered_tables, pl->world, (int32_t)pl->site_cap, (int32_t)pl->allele_cap, G, ctx->num_sms, ctx->stream));
  DevSites& T = st.T_glob;
  T = DevSites();
  T.start = G.start; T.ref_len = G.ref_len; T.alt_len = G.alt_len; T.allele_off = G.allele_off; T.alleles4 = G.alleles4;
  T.count = G.count; T.hdr = G.hdr; T.cap = G.cap; T.nib_cap = G.nib_cap; T.end = G.end; T.pme = G.pme; T.sidx = G.sidx;
  T.caller_arrays = true;
  {
    DSites S{};
    S.hdr = T.hdr; S.cap = T.cap; S.n = T.cap; S.c_hi = T.cap; S.start = T.start; S.pme = T.pme;
    CK(launch_shard_site_index(S, st.hs, nb, G.sidx, G.hdr, ctx->stream));
  }
  ctx->timing.n_launches += 2;
  // the call stage against the global table; rows into the context's block (host-style destination)
  avo_site_results fake{};
  fake.mem = AVO_MEM_HOST; fake.n_states = pl->n_states; fake.n_sites = T.cap;
  uint64_t* h_slots = reinterpret_cast<uint64_t*>(ctx->h_small + 48);
  *h_slots = 0;
  CKS(call_internal(ctx, st.R, st.hs, T, cnv, params, true, &fake, nullptr, &st.D, &st.slots_cap));
  st.ns = st.D.ns;
  DResults to{};                 // direct rows: the owned rows go from the context's block into root_rows, local or peer memory
  if (root_rows) {
    if (root_rows->n_states != st.ns) return fail(ctx, AVO_E_INVALID, "root_rows->n_states must be %d", st.ns);
    const avo_site_results* g = root_rows;
    to.ns = st.ns;
    to.emitted = g->emitted; to.allele_fwd = g->allele_fwd; to.other_fwd = g->other_fwd; to.allele_cov = g->allele_cov;
    to.other_cov = g->other_cov; to.total_cov = g->total_cov; to.square_mapq = g->square_mapq; to.ref_ll = g->ref_ll;
    to.allele_ll = g->allele_ll; to.other_ll = g->other_ll; to.nonref_ll = g->nonref_ll; to.first_is_ref = g->first_is_ref;
    to.copy_number = g->copy_number; to.n_alt = g->n_alt; to.n_ref = g->n_ref; to.n_other = g->n_other; to.qual = g->qual;
    to.gq = g->gq; to.sb = g->sb; to.fisher = g->fisher; to.rms_mapq = g->rms_mapq; to.gl = g->gl; to.nonref_gl = g->nonref_gl;
  }
  void* d_boff = ctx->bufs[SL_C_BOFF].p;
  const uint64_t* d_slots = st.slots_cap != SIZE_MAX ? (const uint64_t*)d_boff + std::min<int64_t>(T.cap, fake.n_sites) : nullptr;
  CK(launch_shard_pack_rows(st.D, st.ns, gathered_tables, pl->world, pl->rank, (int32_t)pl->site_cap, (int32_t)pl->allele_cap,
                            d_slots, st.slots_cap == SIZE_MAX ? 0 : (uint64_t)st.slots_cap, rows_block, (int32_t)pl->row_cap,
                            root_rows ? &to : nullptr, root_rows ? root_rows->n_sites : 0, ctx->num_sms, ctx->stream));
  ctx->timing.n_launches += 1;
  st.phase = 2;
  return AVO_OK;
}

// host form of walk_events (avo_discover.cu): the same events in the same order
static int derive_events_host(avo_ctx* ctx, const avo_read_batch* r, avo_read_events* ev) {
  if (r->wire_meta || r->wire6_meta) return fail(ctx, AVO_E_INVALID, "avo_derive_events needs the full records of a host batch");
  const int64_t n = r->n_reads, scap = ev->n_snv, icap = ev->n_indel;
  int64_t ns = 0, ni = 0;
  int32_t mn = INT32_MAX, mxe = INT32_MIN, mxs = 0;
  for (int64_t i = 0; i < n; ++i) {
    const avo_read_meta& m = r->meta[i];
    if ((i & 31) == 0) ev->snv_first[i >> 5] = (uint32_t)ns;
    if (i > 0 && r->meta[i - 1].start > m.start) return fail(ctx, AVO_E_UNSORTED, "reads are not sorted by start");
    mn = std::min(mn, m.start); mxe = std::max(mxe, m.end); mxs = std::max(mxs, m.end - m.start);
    int32_t pos = m.start;
    uint32_t rb = m.refb_off;
    bool ff = true;
    for (uint32_t k = 0; k < m.n_ops; ++k) {
      const uint32_t op = r->ops[m.op_off + k], len = op >> 4, code = op & 15u;
      const bool isX = code == AVO_OP_MISMATCH, isD = code == AVO_OP_DEL, aligned = code <= AVO_OP_MISMATCH;
      if (isX)
        for (uint32_t j = 0; j < len; ++j) {
          const uint32_t q = (j < 4u) ? ((m.qhead >> (8 * j)) & 255u)
                                      : r->payload[((size_t)(m.seq_off + j / AVO_BLOCK_BASES) << 4) + 5 + j % AVO_BLOCK_BASES];
          if (ns < scap) { ev->snv_pos[ns] = pos + (int32_t)j; ev->snv_pq[ns] = (uint16_t)(((uint32_t)r->refb[rb + j] << 8) | q); }
          ++ns;
        }
      if ((code == AVO_OP_INS || isD) && !ff) {
        if (ni < icap) { ev->indel_read[ni] = (int32_t)i; ev->indel_op[ni] = k; }
        ++ni;
      }
      pos += (aligned || isD) ? (int32_t)len : 0;
      rb += isX ? len : isD ? ((len + 1) >> 1) : 0u;
      ff = ff && !aligned;
    }
  }
  ev->snv_first[(n + 31) >> 5] = (uint32_t)ns;
  ev->n_snv = ns; ev->n_indel = ni;
  if (ns > scap || ni > icap) return fail(ctx, AVO_E_CAPACITY, "events need %lld + %lld entries", (long long)ns, (long long)ni);
  ev->stats.min_start = n ? mn : 0; ev->stats.max_end = n ? mxe : 0; ev->stats.max_span = mxs; ev->stats.valid = 1;
  return AVO_OK;
}

int avo_derive_events(avo_ctx* ctx, const avo_read_batch* reads, avo_read_events* ev) {
  // host batches need no device: ctx may be NULL for them (a host-side packer's last step)
  if (!ctx && !(reads && reads->mem == AVO_MEM_HOST)) return AVO_E_INVALID;
  if (ctx) CKS(begin_call(ctx));
  CKS(validate_reads(ctx, reads));
  if (!ev || !ev->snv_first || ev->n_snv < 0 || ev->n_indel < 0 || (ev->n_snv > 0 && (!ev->snv_pos || !ev->snv_pq)) ||
      (ev->n_indel > 0 && (!ev->indel_read || !ev->indel_op)))
    return fail(ctx, AVO_E_INVALID, "events needs its arrays and their capacities");
  ev->stats.valid = 0;
  const int64_t n = reads->n_reads;
  if (reads->mem == AVO_MEM_HOST && n > 0) return derive_events_host(ctx, reads, ev);
  if (n == 0) {
    if (reads->mem == AVO_MEM_DEVICE) CK(cudaMemsetAsync(ev->snv_first, 0, sizeof(uint32_t), ctx->stream));
    else ev->snv_first[0] = 0;
    ev->n_snv = ev->n_indel = 0;
    ev->stats.min_start = ev->stats.max_end = ev->stats.max_span = 0; ev->stats.valid = 1;
    if (reads->mem == AVO_MEM_DEVICE) CK(cudaStreamSynchronize(ctx->stream));
    return AVO_OK;
  }
  CK(cudaSetDevice(ctx->device));
  DReads R{};
  R.n = n; R.contig = reads->contig_id; R.meta = reads->meta; R.payload = reads->payload; R.ops = reads->ops; R.refb = reads->refb;
  R.n_ops = reads->n_ops; R.n_refb = reads->n_refb;
  const int64_t nw = (n + 31) >> 5;
  void *c_snv, *c_ind, *i_first, *d_stats, *d_temp;
  CKS(get_buf(ctx, SL_EV_CSNV, (size_t)(nw + 1) * 4, &c_snv));
  CKS(get_buf(ctx, SL_EV_CIND, (size_t)(nw + 1) * 4, &c_ind));
  CKS(get_buf(ctx, SL_EV_IFIRST, (size_t)(nw + 1) * 4, &i_first));
  CKS(get_buf(ctx, SL_STATS, sizeof(BatchStats), &d_stats));
  const size_t tb = events_scan_temp_bytes(nw + 1);
  CKS(get_buf(ctx, SL_TEMP, tb, &d_temp));
  const BatchStats init{INT32_MAX, INT32_MIN, 0, 0};
  CK(cudaMemcpyAsync(d_stats, &init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync((uint32_t*)c_snv + nw, 0, 4, ctx->stream));
  CK(cudaMemsetAsync((uint32_t*)c_ind + nw, 0, 4, ctx->stream));
  CK(launch_events_count(R, (uint32_t*)c_snv, (uint32_t*)c_ind, (BatchStats*)d_stats, ctx->stream));
  CK(launch_events_scan((const uint32_t*)c_snv, ev->snv_first, nw + 1, d_temp, tb, ctx->stream));
  CK(launch_events_scan((const uint32_t*)c_ind, (uint32_t*)i_first, nw + 1, d_temp, tb, ctx->stream));
  uint32_t tot[2];
  CK(cudaMemcpyAsync(&tot[0], ev->snv_first + nw, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(&tot[1], (uint32_t*)i_first + nw, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ctx->h_stats, d_stats, sizeof(BatchStats), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.n_launches += 3;
  if (ctx->h_stats->unsorted) return fail(ctx, AVO_E_UNSORTED, "reads are not sorted by start");
  const int64_t scap = ev->n_snv, icap = ev->n_indel;
  ev->n_snv = tot[0]; ev->n_indel = tot[1];
  if ((int64_t)tot[0] > scap || (int64_t)tot[1] > icap)
    return fail(ctx, AVO_E_CAPACITY, "events need %u + %u entries", tot[0], tot[1]);
  CK(launch_events_write(R, ev->snv_first, (const uint32_t*)i_first, scap, icap, ev->snv_pos, ev->snv_pq, ev->indel_read,
                         ev->indel_op, ctx->stream));
  ctx->timing.n_launches += 1;
  CK(cudaStreamSynchronize(ctx->stream));
  ev->stats.min_start = ctx->h_stats->min_start; ev->stats.max_end = ctx->h_stats->max_end;
  ev->stats.max_span = ctx->h_stats->max_span; ev->stats.valid = 1;
  return AVO_OK;
}

int avo_peer_alloc(avo_ctx* ctx, int64_t bytes, void** ptr, unsigned char* handle) {
  if (!ctx || !ptr || !handle || bytes <= 0) return ctx ? fail(ctx, AVO_E_INVALID, "bad argument") : AVO_E_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == AVO_PEER_HANDLE_BYTES, "CUDA IPC handle size");
  CK(cudaSetDevice(ctx->device));
  *ptr = nullptr;
  CK(cudaMalloc(ptr, (size_t)bytes));
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, *ptr);
  if (e != cudaSuccess) { cudaFree(*ptr); *ptr = nullptr; CK(e); }
  memcpy(handle, &h, sizeof h);
  return AVO_OK;
}

int avo_peer_open(avo_ctx* ctx, const unsigned char* handle, void** ptr) {
  if (!ctx || !ptr || !handle) return ctx ? fail(ctx, AVO_E_INVALID, "bad argument") : AVO_E_INVALID;
  CK(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  *ptr = nullptr;
  CK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return AVO_OK;
}

int avo_peer_close(avo_ctx* ctx, void* ptr) {
  if (!ctx) return AVO_E_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (ptr) CK(cudaIpcCloseMemHandle(ptr));
  return AVO_OK;
}

int avo_peer_free(avo_ctx* ctx, void* ptr) {
  if (!ctx) return AVO_E_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (ptr) CK(cudaFree(ptr));
  return AVO_OK;
}

int avo_shard_finish(avo_ctx* ctx, const avo_shard_plan* pl, const void* gathered_tables, const void* gathered_rows,
                     avo_sites_out* gs, avo_site_results* gr, avo_shard_plan* needed) {
  if (!ctx) return AVO_E_INVALID;
  CK(cudaSetDevice(ctx->device));
  CKS(check_plan(ctx, pl));
  if (!ctx->shard || ctx->shard->phase != 2) return fail(ctx, AVO_E_INVALID, "avo_shard_finish follows avo_shard_call");
  if (!gathered_tables || !gathered_rows || !gs) return fail(ctx, AVO_E_INVALID, "NULL argument");
  CKS(check_results(ctx, gr));
  if (gr->mem != AVO_MEM_DEVICE) return fail(ctx, AVO_E_INVALID, "global_results must be device arrays");
  ShardState& st = *ctx->shard;
  st.phase = 0;
  if (gr->n_states != st.ns) return fail(ctx, AVO_E_INVALID, "global_results->n_states must be %d", st.ns);
  DResults O;
  O.ns = st.ns;
  O.emitted = gr->emitted; O.allele_fwd = gr->allele_fwd; O.other_fwd = gr->other_fwd; O.allele_cov = gr->allele_cov;
  O.other_cov = gr->other_cov; O.total_cov = gr->total_cov; O.square_mapq = gr->square_mapq; O.ref_ll = gr->ref_ll;
  O.allele_ll = gr->allele_ll; O.other_ll = gr->other_ll; O.nonref_ll = gr->nonref_ll; O.first_is_ref = gr->first_is_ref;
  O.copy_number = gr->copy_number; O.n_alt = gr->n_alt; O.n_ref = gr->n_ref; O.n_other = gr->n_other; O.qual = gr->qual;
  O.gq = gr->gq; O.sb = gr->sb; O.fisher = gr->fisher; O.rms_mapq = gr->rms_mapq; O.gl = gr->gl; O.nonref_gl = gr->nonref_gl;
  const int64_t out_rows = gr->n_sites;
  CK(launch_shard_unpack_rows(gathered_rows, gathered_tables, (int32_t)pl->site_cap, (int32_t)pl->allele_cap, pl->world,
                              (int32_t)pl->row_cap, st.ns, O, out_rows, st.d_hdrs, ctx->num_sms, ctx->stream));
  ctx->timing.n_launches += 1;
  // headers of both gathers (collected by the kernel: one copy) + this rank's own counters, one synchronisation
  CK(cudaMemcpyAsync(st.h_hdrs, st.d_hdrs, 2 * (size_t)pl->world * sizeof(ShardHdr), cudaMemcpyDeviceToHost, ctx->stream));
  const ShardHdr *ht = st.h_hdrs, *hr = st.h_hdrs + pl->world;
  CKS(discover_fetch_counts(ctx));
  CK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  // this rank's own capacities (grown for the next run if short; the flags in its header told everyone)
  bool again_local = false;
  {
    DevSites t = st.T_loc;
    const int rc = discover_check(ctx, &st.caps, &t, &again_local);
    if (rc != AVO_OK && rc != AVO_E_UNSORTED && rc != AVO_E_INVALID) return rc;
    if (again_local) {            // remember the larger capacities for the next avo_shard_discover
      ctx->site_guess = std::max(ctx->site_guess, st.caps.ucap);
      ctx->indel_guess = std::max<int64_t>(ctx->indel_guess, st.caps.lcap);
      ctx->nib_guess = std::max(ctx->nib_guess, st.caps.nib_cap);
    }
    if (st.slots_cap != SIZE_MAX) {
      const size_t actual = (size_t)*reinterpret_cast<uint64_t*>(ctx->h_small + 48);
      ctx->slots_guess = actual + actual / 8 + 1024;
    }
  }
  int32_t flags = 0;
  int64_t rows = 0, bytes = 0, max_rows = 0, max_bytes = 0;
  for (int r = 0; r < pl->world; ++r) {
    flags |= ht[r].flags | hr[r].flags;
    rows += ht[r].n_rows; bytes += ht[r].n_bytes;
    max_rows = std::max<int64_t>(max_rows, ht[r].n_rows); max_bytes = std::max<int64_t>(max_bytes, ht[r].n_bytes);
  }
  ctx->timing.ms_total = ev_ms(ctx, 0, 7);
  ctx->timing.ms_discover = ev_ms(ctx, 1, 6);
  ctx->timing.ms_discover_tiles = ev_ms(ctx, 2, 3);
  ctx->timing.ms_call_tiles = ev_ms(ctx, 4, 5);
  if (ctx->pairs_timed) ctx->timing.ms_call_pairs = ev_ms(ctx, 8, 9);
  if (flags & SHARD_UNSORTED) return fail(ctx, AVO_E_UNSORTED, "reads are not sorted by start (on some rank)");
  if (flags & SHARD_BAD_STATS) return fail(ctx, AVO_E_INVALID, "reads->stats does not bound the batch (on some rank)");
  if (needed) {
    *needed = *pl;
    needed->site_cap = max_rows + max_rows / 8 + 256;
    needed->allele_cap = (max_bytes + max_bytes / 8 + 4096 + 15) & ~(int64_t)15;
    needed->row_cap = needed->site_cap;
  }
  gs->n = rows;
  gs->n_allele_nibbles = 2 * bytes;
  gr->n_sites = rows;
  if (flags & SHARD_BLOCK_FULL)
    return fail(ctx, AVO_E_CAPACITY, "shard blocks too small: a rank owns %lld rows / %lld allele bytes", (long long)max_rows,
                (long long)max_bytes);
  if (flags & SHARD_RETRY) return fail(ctx, AVO_E_RETRY, "a rank ran out of an internal capacity; run the step again");
  if (rows > gs->capacity || 2 * bytes > gs->allele_capacity || rows > out_rows)
    return fail(ctx, AVO_E_CAPACITY, "global tables too small: need %lld rows, %lld allele nibbles", (long long)rows,
                (long long)(2 * bytes));
  return AVO_OK;
}

// ---- PrefilterReads' mask + extractAlignmentOperators + payload packing, on the device -----------------
int avo_pack_reads_device(avo_ctx* ctx, const avo_text_reads* in, int32_t in_mem, avo_packed_out* out) {
  Range r_pack(T_EXTRACT_ALIGNMENT);
  CKS(begin_call(ctx));
  if (!in || !out || in->n < 0 || in->n >= (int64_t)INT32_MAX) return fail(ctx, AVO_E_INVALID, "bad text batch");
  if (in_mem != AVO_MEM_HOST && in_mem != AVO_MEM_DEVICE) return fail(ctx, AVO_E_INVALID, "bad in_mem");
  const size_t n = (size_t)in->n;
  if (n && (!in->start || !in->end || !in->mapq || !in->rec_flags || !in->cigar_off || !in->md_off ||
            !in->seq_off || !in->qual_off))
    return fail(ctx, AVO_E_INVALID, "text batch has NULL columns");
  if (!out->meta || !out->payload || !out->ops || !out->refb)
    return fail(ctx, AVO_E_INVALID, "output columns must be device buffers sized from avo_pack_bounds");
  if ((uintptr_t)out->meta & 15u || (uintptr_t)out->payload & 15u)
    return fail(ctx, AVO_E_INVALID, "out->meta and out->payload must be 16-byte aligned");
  if (n == 0) { out->n_reads = out->n_blocks = out->n_ops = out->n_refb = 0; return AVO_OK; }
  // blob sizes = the last offsets (64-bit columns, or 32-bit ones: avo_text_reads.narrow)
  const bool narrow = in->narrow != 0;
  int64_t tot[4];
  uint32_t tot32[4] = {0, 0, 0, 0};
  const int64_t* offs[4] = {in->cigar_off, in->md_off, in->seq_off, in->qual_off};
  for (int k = 0; k < 4; ++k) {
    if (in_mem == AVO_MEM_HOST) tot[k] = narrow ? (int64_t)reinterpret_cast<const uint32_t*>(offs[k])[n] : offs[k][n];
    else if (narrow) CK(cudaMemcpyAsync(&tot32[k], reinterpret_cast<const uint32_t*>(offs[k]) + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    else CK(cudaMemcpyAsync(&tot[k], offs[k] + n, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (in_mem == AVO_MEM_DEVICE) {
    CK(cudaStreamSynchronize(ctx->stream));
    if (narrow) for (int k = 0; k < 4; ++k) tot[k] = tot32[k];
  }
  for (int k = 0; k < 4; ++k)
    if (tot[k] < 0) return fail(ctx, AVO_E_INVALID, "negative blob size");
  if ((tot[0] && !in->cigar) || (tot[1] && !in->md) || (tot[2] && !in->seq) || (tot[3] && !in->qual))
    return fail(ctx, AVO_E_INVALID, "text batch has NULL blobs");
  DText T;
  T.n = in->n;
  T.narrow = narrow ? 1 : 0;
  CK(cudaEventRecord(ctx->ev[0], ctx->stream));
  // numeric columns travel as bytes of their own width
  auto stage_num = [&](int slot, const void* src, size_t count, size_t wide, size_t slim, const void** dst) -> int {
    const uint8_t* d = nullptr;
    const int rc = stage_in(ctx, slot, reinterpret_cast<const uint8_t*>(src), count * (narrow ? slim : wide), in_mem, &d);
    *dst = d;
    return rc;
  };
  const void* q = nullptr;
  CKS(stage_num(SL_T_START, in->start, n, 8, 4, &q)); T.start = (const int64_t*)q;
  CKS(stage_num(SL_T_END, in->end, n, 8, 4, &q)); T.end = (const int64_t*)q;
  CKS(stage_num(SL_T_MAPQ, in->mapq, n, 4, 1, &q)); T.mapq = (const int32_t*)q;
  CKS(stage_in(ctx, SL_T_FLAGS, in->rec_flags, n, in_mem, &T.rec_flags));
  CKS(stage_in(ctx, SL_T_CIGAR, in->cigar, (size_t)tot[0], in_mem, &T.cigar));
  CKS(stage_num(SL_T_CIGAR_OFF, in->cigar_off, n + 1, 8, 4, &q)); T.cigar_off = (const int64_t*)q;
  CKS(stage_in(ctx, SL_T_MD, in->md, (size_t)tot[1], in_mem, &T.md));
  CKS(stage_num(SL_T_MD_OFF, in->md_off, n + 1, 8, 4, &q)); T.md_off = (const int64_t*)q;
  CKS(stage_in(ctx, SL_T_SEQ, in->seq, (size_t)tot[2], in_mem, &T.seq));
  CKS(stage_num(SL_T_SEQ_OFF, in->seq_off, n + 1, 8, 4, &q)); T.seq_off = (const int64_t*)q;
  if (in->qual == in->seq) T.qual = T.seq; else CKS(stage_in(ctx, SL_T_QUAL, in->qual, (size_t)tot[3], in_mem, &T.qual));
  if (in->qual_off == in->seq_off) T.qual_off = T.seq_off;
  else { CKS(stage_num(SL_T_QUAL_OFF, in->qual_off, n + 1, 8, 4, &q)); T.qual_off = (const int64_t*)q; }
  T.keep = nullptr;
  T.prefilter = in->prefilter ? 1 : 0;
  if (in->prefilter) T.pf = *in->prefilter; else memset(&T.pf, 0, sizeof T.pf);
  if (in->keep) CKS(stage_in(ctx, SL_T_KEEP, in->keep, n, in_mem, &T.keep));
  CK(cudaEventRecord(ctx->ev[1], ctx->stream));
  void *d_counts, *d_offsets, *d_temp;
  const size_t tb = pack_scan_temp_bytes(in->n);
  CKS(get_buf(ctx, SL_T_COUNTS, (n + 1) * sizeof(uint4), &d_counts));
  CKS(get_buf(ctx, SL_T_OFFSETS, (n + 1) * sizeof(PackOffsets), &d_offsets));
  CKS(get_buf(ctx, SL_TEMP, tb, &d_temp));
  CK(launch_pack_count(T, (uint4*)d_counts, (PackOffsets*)d_offsets, d_temp, tb, ctx->stream, &ctx->timing.n_launches));
  PackOffsets total;
  CK(cudaMemcpyAsync(&total, (PackOffsets*)d_offsets + n, sizeof total, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (total.b > 0xFFFFFFFFull || total.o > 0xFFFFFFFFull || total.f > 0x7FFFFFFFull)
    return fail(ctx, AVO_E_INVALID, "batch totals must fit 32-bit offsets; split the batch");
  if ((int64_t)total.k > out->n_reads || (int64_t)total.b > out->n_blocks || (int64_t)total.o > out->n_ops ||
      (int64_t)total.f > out->n_refb)
    return fail(ctx, AVO_E_CAPACITY, "packed output too small: need %llu reads, %llu blocks, %llu operators, %llu bytes",
                (unsigned long long)total.k, (unsigned long long)total.b, (unsigned long long)total.o,
                (unsigned long long)total.f);
  DPacked P;
  P.meta = out->meta; P.payload = out->payload; P.ops = out->ops; P.refb = out->refb; P.source_index = out->source_index;
  CK(launch_pack_write(T, (const PackOffsets*)d_offsets, P, ctx->stream, &ctx->timing.n_launches));
  CK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  out->n_reads = (int64_t)total.k; out->n_blocks = (int64_t)total.b; out->n_ops = (int64_t)total.o;
  out->n_refb = (int64_t)total.f;
  ctx->timing.ms_total = ev_ms(ctx, 0, 7);
  ctx->timing.ms_h2d = ev_ms(ctx, 0, 1);
  return AVO_OK;
}

// ---- SquareOffReferenceModel.squareOffSite ----------------------------------------------------------
int avo_square_off(avo_ctx* ctx, const avo_sites* cohort, const avo_sample_rows* sample, avo_square_out* out) {
  CKS(begin_call(ctx));
  if (!cohort || !sample || !out || !sample->variants || !sample->variant_rows || !sample->ref_model_rows)
    return fail(ctx, AVO_E_INVALID, "NULL argument");
  const avo_sites* V = sample->variants;
  const avo_site_results* VR = sample->variant_rows;
  const avo_site_results* RR = sample->ref_model_rows;
  const int mem = out->mem;
  if ((mem != AVO_MEM_HOST && mem != AVO_MEM_DEVICE) || cohort->mem != mem || V->mem != mem || VR->mem != mem ||
      RR->mem != mem)
    return fail(ctx, AVO_E_INVALID, "all tables must live in the memory out->mem names");
  if (out->n_sites != cohort->n || VR->n_sites != V->n || cohort->n >= INT32_MAX || V->n >= INT32_MAX ||
      sample->n_ref_model >= INT32_MAX || sample->n_ref_model < 0)
    return fail(ctx, AVO_E_INVALID, "inconsistent table sizes");
  const int ns = out->n_states;
  if (ns < 2 || ns > AVO_MAX_PLOIDY + 1 || VR->n_states != ns || (sample->n_ref_model > 0 && RR->n_states != ns))
    return fail(ctx, AVO_E_INVALID, "n_states must agree");
  const size_t N = (size_t)cohort->n, nv = (size_t)V->n, nr = (size_t)sample->n_ref_model;
  if (N == 0) return AVO_OK;
  // host tables are staged through stream-ordered allocations (this entry point is not a hot loop)
  std::vector<void*> tmp;
  auto release = [&]() { for (void* q : tmp) cudaFreeAsync(q, ctx->stream); };
  cudaError_t cerr = cudaSuccess;
  auto in = [&](const void* src, size_t bytes) -> const void* {
    if (mem == AVO_MEM_DEVICE || bytes == 0 || !src) return src;
    void* d = nullptr;
    if (cerr == cudaSuccess) cerr = cudaMallocAsync(&d, bytes, ctx->stream);
    if (cerr == cudaSuccess) { tmp.push_back(d); cerr = cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, ctx->stream); }
    ctx->timing.h2d_bytes += (int64_t)bytes;
    return d;
  };
  // the look-back bounds: the longest variant and the longest reference-model block.  Host tables are
  // scanned here; device tables by one small reduction each (discovery no longer caps allele lengths)
  int32_t max_ref_len = 1, max_block_len = 1;
  if (mem == AVO_MEM_HOST) {
    for (size_t k = 0; k < nv; ++k) max_ref_len = std::max(max_ref_len, V->ref_len[k]);
    if (sample->ref_model_len) {
      max_block_len = sample->ref_model_max_len;
      for (size_t k = 0; k < nr; ++k) {
        if (sample->ref_model_len[k] < 1) return fail(ctx, AVO_E_INVALID, "ref_model_len[%zu] < 1", k);
        max_block_len = std::max(max_block_len, sample->ref_model_len[k]);
      }
    }
  } else {
    void* d_small;
    CKS(get_buf(ctx, SL_SMALL, SW_COUNT_ * 4, &d_small));
    int32_t* mx = (int32_t*)d_small + SW_TILE;
    const int32_t one[2] = {1, 1};
    CK(cudaMemcpyAsync(mx, one, sizeof one, cudaMemcpyHostToDevice, ctx->stream));
    CK(launch_max_i32(V->ref_len, (int64_t)nv, mx, ctx->stream));
    const bool scan_blocks = sample->ref_model_len && sample->ref_model_max_len <= 0;
    if (scan_blocks) CK(launch_max_i32(sample->ref_model_len, (int64_t)nr, mx + 1, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_small, mx, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    max_ref_len = std::max(1, ctx->h_small[0]);
    if (sample->ref_model_len) max_block_len = scan_blocks ? std::max(1, ctx->h_small[1]) : sample->ref_model_max_len;
  }
  DSquareIn I{};
  I.n_sites = (int32_t)N; I.n_v = (int32_t)nv; I.n_r = (int32_t)nr; I.ns = ns; I.v_max_ref_len = max_ref_len;
  I.r_max_len = max_block_len;
  I.r_len = (const int32_t*)in(sample->ref_model_len, nr * 4);
  I.c_start = (const int32_t*)in(cohort->start, N * 4); I.c_ref_len = (const int32_t*)in(cohort->ref_len, N * 4);
  I.c_alt_len = (const int32_t*)in(cohort->alt_len, N * 4);
  I.c_allele_off = (const uint32_t*)in(cohort->allele_off, (N + 1) * 4);
  I.c_alleles4 = (const uint8_t*)in(cohort->alleles4, (size_t)(cohort->n_allele_nibbles + 1) / 2);
  I.v_start = (const int32_t*)in(V->start, nv * 4); I.v_ref_len = (const int32_t*)in(V->ref_len, nv * 4);
  I.v_alt_len = (const int32_t*)in(V->alt_len, nv * 4); I.v_allele_off = (const uint32_t*)in(V->allele_off, (nv + 1) * 4);
  I.v_alleles4 = (const uint8_t*)in(V->alleles4, (size_t)(V->n_allele_nibbles + 1) / 2);
  I.v_emitted = (const uint8_t*)in(VR->emitted, nv); I.v_n_alt = (const int32_t*)in(VR->n_alt, nv * 4);
  I.v_n_ref = (const int32_t*)in(VR->n_ref, nv * 4); I.v_n_other = (const int32_t*)in(VR->n_other, nv * 4);
  I.v_gl = (const double*)in(VR->gl, nv * ns * 8); I.v_nonref_gl = (const double*)in(VR->nonref_gl, nv * ns * 8);
  I.r_start = (const int32_t*)in(sample->ref_model_start, nr * 4); I.r_emitted = (const uint8_t*)in(RR->emitted, nr);
  I.r_n_alt = (const int32_t*)in(RR->n_alt, nr * 4); I.r_n_ref = (const int32_t*)in(RR->n_ref, nr * 4);
  I.r_n_other = (const int32_t*)in(RR->n_other, nr * 4); I.r_nonref_gl = (const double*)in(RR->nonref_gl, nr * ns * 8);
  DSquareOut O{};
  const size_t osz[6] = {N, N * 4, N * 4, N * 4, N * 4, N * ns * 8};
  void* odst[6] = {out->kind, out->src, out->n_alt, out->n_ref, out->n_other, out->gl};
  void* odev[6];
  for (int k = 0; k < 6; ++k) {
    if (!odst[k]) { release(); return fail(ctx, AVO_E_INVALID, "square-off output has NULL columns"); }
    odev[k] = odst[k];
    if (mem == AVO_MEM_HOST && cerr == cudaSuccess) {
      cerr = cudaMallocAsync(&odev[k], osz[k], ctx->stream);
      if (cerr == cudaSuccess) tmp.push_back(odev[k]);
    }
  }
  if (cerr != cudaSuccess) { release(); return fail(ctx, AVO_E_CUDA, "staging failed: %s", cudaGetErrorString(cerr)); }
  O.kind = (uint8_t*)odev[0]; O.src = (int32_t*)odev[1]; O.n_alt = (int32_t*)odev[2]; O.n_ref = (int32_t*)odev[3];
  O.n_other = (int32_t*)odev[4]; O.gl = (double*)odev[5];
  CK(cudaEventRecord(ctx->ev[4], ctx->stream));
  CK(launch_square_off(I, O, ctx->stream));
  CK(cudaEventRecord(ctx->ev[5], ctx->stream));
  ctx->timing.n_launches += 1;
  if (mem == AVO_MEM_HOST)
    for (int k = 0; k < 6; ++k) {
      CK(cudaMemcpyAsync(odst[k], odev[k], osz[k], cudaMemcpyDeviceToHost, ctx->stream));
      ctx->timing.d2h_bytes += (int64_t)osz[k];
    }
  release();
  CK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.ms_total = ev_ms(ctx, 0, 7);
  ctx->timing.ms_observe = ev_ms(ctx, 4, 5);
  return AVO_OK;
}

// ---- TrioCaller ------------------------------------------------------------------------------------
int avo_trio_call(avo_ctx* ctx, avo_trio* t) {
  CKS(begin_call(ctx));
  if (!t) return fail(ctx, AVO_E_INVALID, "trio is NULL");
  if (t->n_sites < 0 || (t->mem != AVO_MEM_HOST && t->mem != AVO_MEM_DEVICE)) return fail(ctx, AVO_E_INVALID, "bad trio");
  const size_t n = (size_t)t->n_sites;
  if (n == 0) return AVO_OK;
  if (!t->n_alt || !t->n_ref || !t->n_other || !t->kept || !t->child_action || !t->filled)
    return fail(ctx, AVO_E_INVALID, "trio has NULL columns");
  std::vector<void*> tmp;
  cudaError_t cerr = cudaSuccess;
  auto in = [&](const void* src, size_t bytes) -> const void* {
    if (t->mem == AVO_MEM_DEVICE || !src) return src;
    void* d = nullptr;
    if (cerr == cudaSuccess) cerr = cudaMallocAsync(&d, bytes, ctx->stream);
    if (cerr == cudaSuccess) { tmp.push_back(d); cerr = cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, ctx->stream); }
    return d;
  };
  DTrioIn I{};
  I.n_sites = t->n_sites;
  I.present = (const uint8_t*)in(t->present, 3 * n);
  I.n_alt = (const int32_t*)in(t->n_alt, 12 * n); I.n_ref = (const int32_t*)in(t->n_ref, 12 * n);
  I.n_other = (const int32_t*)in(t->n_other, 12 * n); I.n_nocall = (const int32_t*)in(t->n_nocall, 12 * n);
  void* odst[3] = {t->kept, t->child_action, t->filled};
  void* odev[3] = {t->kept, t->child_action, t->filled};
  if (t->mem == AVO_MEM_HOST)
    for (int k = 0; k < 3 && cerr == cudaSuccess; ++k) {
      cerr = cudaMallocAsync(&odev[k], n, ctx->stream);
      if (cerr == cudaSuccess) tmp.push_back(odev[k]);
    }
  if (cerr != cudaSuccess) {
    for (void* q : tmp) cudaFreeAsync(q, ctx->stream);
    return fail(ctx, AVO_E_CUDA, "staging failed: %s", cudaGetErrorString(cerr));
  }
  DTrioOut O{(uint8_t*)odev[0], (uint8_t*)odev[1], (uint8_t*)odev[2]};
  cerr = launch_trio(I, O, ctx->stream);
  ctx->timing.n_launches += 1;
  if (cerr == cudaSuccess && t->mem == AVO_MEM_HOST)
    for (int k = 0; k < 3 && cerr == cudaSuccess; ++k)
      cerr = cudaMemcpyAsync(odst[k], odev[k], n, cudaMemcpyDeviceToHost, ctx->stream);
  for (void* q : tmp) cudaFreeAsync(q, ctx->stream);
  if (cerr != cudaSuccess) return fail(ctx, AVO_E_CUDA, "trio: %s", cudaGetErrorString(cerr));
  CK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.ms_total = ev_ms(ctx, 0, 7);
  return AVO_OK;
}

// ---- RewriteHets + HardFilterGenotypes ---------------------------------------------------------------
int avo_filter_genotypes(avo_ctx* ctx, const avo_site_results* rows, const int32_t* ref_len,
                         const int32_t* alt_len, const avo_filter_params* params, avo_filter_out* out) {
  CKS(begin_call(ctx));
  if (!rows || !params || !out) return fail(ctx, AVO_E_INVALID, "NULL argument");
  if (rows->n_sites < 0 || out->n != rows->n_sites) return fail(ctx, AVO_E_INVALID, "out->n != rows->n_sites");
  if (out->mem != rows->mem || (rows->mem != AVO_MEM_HOST && rows->mem != AVO_MEM_DEVICE))
    return fail(ctx, AVO_E_INVALID, "rows and out must live in the same memory");
  const size_t n = (size_t)rows->n_sites;
  if (n == 0) return AVO_OK;
  if (!ref_len || !rows->emitted || !rows->n_alt || !rows->n_ref || !rows->n_other || !rows->gq ||
      !rows->total_cov || !rows->allele_cov || !rows->fisher || !rows->rms_mapq || !out->emit ||
      !out->rewritten || !out->filters_applied || !out->filters_failed || !out->n_alt || !out->n_ref ||
      !out->gq_valid)
    return fail(ctx, AVO_E_INVALID, "NULL column");
  DFilterIn I{};
  I.n = rows->n_sites;
  DFilterOut O{};
  const size_t osz[7] = {n, n, n, n * 4, n * 4, n * 4, n};
  void* odst[7] = {out->emit, out->rewritten, out->filters_applied, out->filters_failed, out->n_alt,
                   out->n_ref, out->gq_valid};
  void* odev[7];
  if (rows->mem == AVO_MEM_DEVICE) {
    I.emitted = rows->emitted; I.ref_len = ref_len; I.alt_len = alt_len; I.n_alt = rows->n_alt;
    I.n_ref = rows->n_ref; I.n_other = rows->n_other; I.gq = rows->gq; I.total_cov = rows->total_cov;
    I.allele_cov = rows->allele_cov; I.fisher = rows->fisher; I.rms_mapq = rows->rms_mapq;
    for (int k = 0; k < 7; ++k) odev[k] = odst[k];
  } else {
    // one staging block in, one out
    const void* src[11] = {rows->emitted, ref_len, alt_len, rows->n_alt, rows->n_ref, rows->n_other, rows->gq,
                           rows->total_cov, rows->allele_cov, rows->fisher, rows->rms_mapq};
    const size_t isz[11] = {n, n * 4, alt_len ? n * 4 : 0, n * 4, n * 4, n * 4, n * 4, n * 4, n * 4, n * 4, n * 4};
    size_t off[12]; off[0] = 0;
    for (int k = 0; k < 11; ++k) off[k + 1] = off[k] + align_up(isz[k], 256);
    void* blk;
    CKS(get_buf(ctx, SL_F_IN, off[11], &blk));
    const char* b = (const char*)blk;
    for (int k = 0; k < 11; ++k)
      if (isz[k]) {
        CK(cudaMemcpyAsync((char*)blk + off[k], src[k], isz[k], cudaMemcpyHostToDevice, ctx->stream));
        ctx->timing.h2d_bytes += (int64_t)isz[k];
      }
    I.emitted = (const uint8_t*)(b + off[0]); I.ref_len = (const int32_t*)(b + off[1]);
    I.alt_len = alt_len ? (const int32_t*)(b + off[2]) : nullptr;
    I.n_alt = (const int32_t*)(b + off[3]); I.n_ref = (const int32_t*)(b + off[4]);
    I.n_other = (const int32_t*)(b + off[5]); I.gq = (const int32_t*)(b + off[6]);
    I.total_cov = (const int32_t*)(b + off[7]); I.allele_cov = (const int32_t*)(b + off[8]);
    I.fisher = (const float*)(b + off[9]); I.rms_mapq = (const float*)(b + off[10]);
    size_t oo[8]; oo[0] = 0;
    for (int k = 0; k < 7; ++k) oo[k + 1] = oo[k] + align_up(osz[k], 256);
    void* ob;
    CKS(get_buf(ctx, SL_F_OUT, oo[7], &ob));
    for (int k = 0; k < 7; ++k) odev[k] = (char*)ob + oo[k];
  }
  O.emit = (uint8_t*)odev[0]; O.rewritten = (uint8_t*)odev[1]; O.filters_applied = (uint8_t*)odev[2];
  O.filters_failed = (uint32_t*)odev[3]; O.n_alt = (int32_t*)odev[4]; O.n_ref = (int32_t*)odev[5];
  O.gq_valid = (uint8_t*)odev[6];
  CK(cudaEventRecord(ctx->ev[4], ctx->stream));
  CK(launch_filter_genotypes(I, *params, O, ctx->stream));
  CK(cudaEventRecord(ctx->ev[5], ctx->stream));
  ctx->timing.n_launches += 1;
  if (rows->mem == AVO_MEM_HOST)
    for (int k = 0; k < 7; ++k) {
      CK(cudaMemcpyAsync(odst[k], odev[k], osz[k], cudaMemcpyDeviceToHost, ctx->stream));
      ctx->timing.d2h_bytes += (int64_t)osz[k];
    }
  CK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.ms_total = ev_ms(ctx, 0, 7);
  ctx->timing.ms_observe = ev_ms(ctx, 4, 5);
  return AVO_OK;
}

// ---- joint calling ---------------------------------------------------------------------------------
static int joint_inputs(avo_ctx* ctx, const avo_joint_in* in, DJointIn* J) {
  if (!in) return fail(ctx, AVO_E_INVALID, "joint input is NULL");
  if (in->n_sites < 0 || in->n_samples < 1 || in->n_states < 2 || in->n_states > AVO_MAX_PLOIDY + 1)
    return fail(ctx, AVO_E_INVALID, "joint input: bad n_sites / n_samples / n_states");
  if (in->mem != AVO_MEM_HOST && in->mem != AVO_MEM_DEVICE) return fail(ctx, AVO_E_INVALID, "bad mem");
  if (in->n_sites > 0 && (!in->n_alt || !in->n_ref || !in->n_other || !in->gl))
    return fail(ctx, AVO_E_INVALID, "joint input has NULL columns");
  const size_t sn = (size_t)in->n_samples * (size_t)in->n_sites;
  J->n_sites = in->n_sites; J->n_samples = in->n_samples; J->n_states = in->n_states;
  J->present = nullptr; J->n_nocall = nullptr;
  if (in->present) CKS(stage_in(ctx, SL_J_PRESENT, in->present, sn, in->mem, &J->present));
  CKS(stage_in(ctx, SL_J_NALT, in->n_alt, sn, in->mem, &J->n_alt));
  CKS(stage_in(ctx, SL_J_NREF, in->n_ref, sn, in->mem, &J->n_ref));
  CKS(stage_in(ctx, SL_J_NOTHER, in->n_other, sn, in->mem, &J->n_other));
  if (in->n_nocall) CKS(stage_in(ctx, SL_J_NOCALL, in->n_nocall, sn, in->mem, &J->n_nocall));
  CKS(stage_in(ctx, SL_J_GL, in->gl, sn * (size_t)in->n_states, in->mem, &J->gl));
  return AVO_OK;
}

int avo_joint_counts(avo_ctx* ctx, const avo_joint_in* in, int32_t* alt_alleles, int32_t* called_alleles) {
  CKS(begin_call(ctx));
  DJointIn J;
  CKS(joint_inputs(ctx, in, &J));
  if (!alt_alleles || !called_alleles) return fail(ctx, AVO_E_INVALID, "count outputs are NULL");
  const size_t n = (size_t)in->n_sites;
  int32_t *d_alt = alt_alleles, *d_called = called_alleles;
  if (in->mem == AVO_MEM_HOST) {
    void* q;
    CKS(get_buf(ctx, SL_J_ALT, n * 4, &q)); d_alt = (int32_t*)q;
    CKS(get_buf(ctx, SL_J_CALLED, n * 4, &q)); d_called = (int32_t*)q;
  }
  CK(launch_joint_counts(J, d_alt, d_called, ctx->stream));
  ctx->timing.n_launches += 1;
  if (in->mem == AVO_MEM_HOST && n) {
    CK(cudaMemcpyAsync(alt_alleles, d_alt, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(called_alleles, d_called, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->timing.d2h_bytes += (int64_t)n * 8;
  }
  CK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.ms_total = ev_ms(ctx, 0, 7);
  return AVO_OK;
}

int avo_joint_depths(avo_ctx* ctx, const avo_depths_in* in, avo_site_depths* out) {
  CKS(begin_call(ctx));
  if (!in || !out) return fail(ctx, AVO_E_INVALID, "NULL argument");
  if (in->n_sites < 0 || in->n_samples < 1 || out->n_sites != in->n_sites)
    return fail(ctx, AVO_E_INVALID, "depths: bad n_sites / n_samples");
  if ((in->mem != AVO_MEM_HOST && in->mem != AVO_MEM_DEVICE) || out->mem != in->mem)
    return fail(ctx, AVO_E_INVALID, "depths: input and output must live in the same memory");
  const size_t n = (size_t)in->n_sites, sn = n * (size_t)in->n_samples;
  if (n == 0) return AVO_OK;
  int32_t* ocol[6] = {out->read_depth, out->ref_read_depth, out->fwd_read_depth, out->rev_read_depth,
                      out->ref_fwd_read_depth, out->ref_rev_read_depth};
  if (!out->annotated) return fail(ctx, AVO_E_INVALID, "depths output has NULL columns");
  for (int f = 0; f < 6; ++f)
    if (!ocol[f]) return fail(ctx, AVO_E_INVALID, "depths output has NULL columns");
  if (in->mem == AVO_MEM_DEVICE && in->sb && ((uintptr_t)in->sb & 15u))
    return fail(ctx, AVO_E_INVALID, "depths: sb must be 16-byte aligned");
  DDepthsIn I{};
  I.n_sites = in->n_sites; I.n_samples = in->n_samples;
  if (in->present) CKS(stage_in(ctx, SL_J_PRESENT, in->present, sn, in->mem, &I.present));
  if (in->read_depth) CKS(stage_in(ctx, SL_JD_DP, in->read_depth, sn, in->mem, &I.read_depth));
  if (in->ref_read_depth) CKS(stage_in(ctx, SL_JD_REFDP, in->ref_read_depth, sn, in->mem, &I.ref_read_depth));
  if (in->sb) CKS(stage_in(ctx, SL_JD_SB, in->sb, sn * 4, in->mem, &I.sb));
  DDepthsOut O{};
  O.annotated = out->annotated;
  for (int f = 0; f < 6; ++f) O.col[f] = ocol[f];
  char* d_out = nullptr;
  const size_t col_bytes = (n * 4 + 255) & ~(size_t)255;
  if (in->mem == AVO_MEM_HOST) {
    void* q;
    CKS(get_buf(ctx, SL_JD_OUT, 7 * col_bytes, &q));
    d_out = (char*)q;
    O.annotated = (uint8_t*)d_out;
    for (int f = 0; f < 6; ++f) O.col[f] = (int32_t*)(d_out + (size_t)(f + 1) * col_bytes);
  }
  CK(cudaEventRecord(ctx->ev[4], ctx->stream));
  CK(launch_joint_depths(I, O, ctx->stream));
  CK(cudaEventRecord(ctx->ev[5], ctx->stream));
  ctx->timing.n_launches += 1;
  if (in->mem == AVO_MEM_HOST) {
    CK(cudaMemcpyAsync(out->annotated, O.annotated, n, cudaMemcpyDeviceToHost, ctx->stream));
    for (int f = 0; f < 6; ++f)
      CK(cudaMemcpyAsync(ocol[f], O.col[f], n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->timing.d2h_bytes += (int64_t)(n * 25);
  }
  CK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.ms_total = ev_ms(ctx, 0, 7);
  ctx->timing.ms_observe = ev_ms(ctx, 4, 5);
  return AVO_OK;
}

int avo_joint(avo_ctx* ctx, const avo_joint_in* in, avo_joint_out* out) {
  CKS(begin_call(ctx));
  DJointIn J;
  CKS(joint_inputs(ctx, in, &J));
  if (!out) return fail(ctx, AVO_E_INVALID, "joint output is NULL");
  if (out->n_sites != in->n_sites || out->n_samples != in->n_samples || out->n_states != in->n_states)
    return fail(ctx, AVO_E_INVALID, "joint output shape differs from the input");
  if (out->mem != AVO_MEM_HOST && out->mem != AVO_MEM_DEVICE) return fail(ctx, AVO_E_INVALID, "bad mem");
  if (in->n_sites > 0 && (!out->site_emitted || !out->maf || !out->quality || !out->post0_sum ||
                          !out->flags || !out->posteriors || !out->priors || !out->n_alt ||
                          !out->n_ref || !out->gq))
    return fail(ctx, AVO_E_INVALID, "joint output has NULL columns");
  if ((in->alt_alleles == nullptr) != (in->called_alleles == nullptr))
    return fail(ctx, AVO_E_INVALID, "alt_alleles and called_alleles go together");
  const size_t n = (size_t)in->n_sites, sn = n * (size_t)in->n_samples, ns = (size_t)in->n_states;
  // site totals: given (sharded cohort) or counted here
  const int32_t *d_alt, *d_called;
  if (in->alt_alleles) {
    CKS(stage_in(ctx, SL_J_ALT, in->alt_alleles, n, in->mem, &d_alt));
    CKS(stage_in(ctx, SL_J_CALLED, in->called_alleles, n, in->mem, &d_called));
  } else {
    void* q;
    CKS(get_buf(ctx, SL_J_ALT, n * 4, &q)); d_alt = (int32_t*)q;
    CKS(get_buf(ctx, SL_J_CALLED, n * 4, &q)); d_called = (int32_t*)q;
    CK(launch_joint_counts(J, (int32_t*)d_alt, (int32_t*)d_called, ctx->stream));
    ctx->timing.n_launches += 1;
  }
  // outputs: straight into the caller's arrays when they are on the device, else one staging block
  DJointOut O;
  const size_t sz[10] = {n, n * 8, n * 8, n * 8, sn, sn * ns * 4, sn * ns * 4, sn * 4, sn * 4, sn * 4};
  void* dst[10] = {out->site_emitted, out->maf, out->quality, out->post0_sum, out->flags,
                   out->posteriors, out->priors, out->n_alt, out->n_ref, out->gq};
  void* dev[10];
  if (out->mem == AVO_MEM_DEVICE) {
    for (int k = 0; k < 10; ++k) dev[k] = dst[k];
  } else {
    size_t off[11]; off[0] = 0;
    for (int k = 0; k < 10; ++k) off[k + 1] = off[k] + align_up(sz[k], 256);
    void* blk;
    CKS(get_buf(ctx, SL_J_OUT, off[10], &blk));
    for (int k = 0; k < 10; ++k) dev[k] = (char*)blk + off[k];
  }
  O.site_emitted = (uint8_t*)dev[0]; O.maf = (double*)dev[1]; O.quality = (double*)dev[2];
  O.post0_sum = (double*)dev[3]; O.flags = (uint8_t*)dev[4]; O.posteriors = (float*)dev[5];
  O.priors = (float*)dev[6]; O.n_alt = (int32_t*)dev[7]; O.n_ref = (int32_t*)dev[8]; O.gq = (int32_t*)dev[9];
  if (sn) {
    CK(cudaMemsetAsync(O.posteriors, 0, sz[5], ctx->stream));
    CK(cudaMemsetAsync(O.priors, 0, sz[6], ctx->stream));
  }
  CK(cudaEventRecord(ctx->ev[4], ctx->stream));
  CK(launch_joint_recall(J, d_alt, d_called, O, ctx->d_lgam, ctx->stream));
  CK(cudaEventRecord(ctx->ev[5], ctx->stream));
  ctx->timing.n_launches += 1;
  if (out->mem == AVO_MEM_HOST) {
    for (int k = 0; k < 10; ++k) {
      if (!sz[k]) continue;
      CK(cudaMemcpyAsync(dst[k], dev[k], sz[k], cudaMemcpyDeviceToHost, ctx->stream));
      ctx->timing.d2h_bytes += (int64_t)sz[k];
    }
  }
  CK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.ms_total = ev_ms(ctx, 0, 7);
  ctx->timing.ms_observe = ev_ms(ctx, 4, 5);
  return AVO_OK;
}

}  // extern "C"
