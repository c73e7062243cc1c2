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
#include "avo_synth_core.h"

namespace avo {

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
