// avo_setup.cu -- small set-up kernels around the two tile kernels: batch statistics, the site
// table's derived columns (end, prefix-max end = the Forest fast-path test), the coarse position
// index that replaces per-tile binary searches, and the logFactorial table.
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <algorithm>

#include "avo_device.cuh"

namespace avo {

// ---- batch statistics -------------------------------------------------------------------------
__global__ void k_batch_stats(DReads R, BatchStats* __restrict__ st) {
  int32_t mn = INT32_MAX, mxe = INT32_MIN, mxs = 0, uns = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < R.n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const MetaA a = load_meta_a(R.meta, i);
    const int32_t s = a.start, e = a.end;
    mn = min(mn, s);
    mxe = max(mxe, e);
    mxs = max(mxs, e - s);
    if (i > 0 && meta_start(R.meta, i - 1) > s) uns = 1;
    if (R.coords) R.coords[i] = make_int2(s, e);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o));
    mxe = max(mxe, __shfl_xor_sync(0xFFFFFFFFu, mxe, o));
    mxs = max(mxs, __shfl_xor_sync(0xFFFFFFFFu, mxs, o));
    uns |= __shfl_xor_sync(0xFFFFFFFFu, uns, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&st->min_start, mn);
    atomicMax(&st->max_end, mxe);
    atomicMax(&st->max_span, mxs);
    if (uns) atomicOr(&st->unsorted, 1);
  }
}

__global__ void k_init_stats(BatchStats* st) {
  st->min_start = INT32_MAX; st->max_end = INT32_MIN; st->max_span = 0; st->unsorted = 0;
}

cudaError_t launch_batch_stats(const DReads& R, BatchStats* d_stats, cudaStream_t s) {
  k_init_stats<<<1, 1, 0, s>>>(d_stats);
  if (R.n > 0) {
    const int grid = (int)min((int64_t)2048, (R.n + 255) / 256);
    k_batch_stats<<<grid, 256, 0, s>>>(R, d_stats);
  }
  return cudaGetLastError();
}

// ---- site table -------------------------------------------------------------------------------
__global__ void k_site_end(int32_t n, const int32_t* __restrict__ contig,
                           const int32_t* __restrict__ start, const int32_t* __restrict__ ref_len,
                           int32_t* __restrict__ end, int32_t* __restrict__ unsorted) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t e = start[i] + ref_len[i];
  end[i] = e;
  if (i > 0) {  // ReferenceRegion ordering: (contig, start, end)
    const int32_t c0 = contig ? contig[i - 1] : 0, c1 = contig ? contig[i] : 0;
    const int32_t s0 = start[i - 1], e0 = s0 + ref_len[i - 1];
    const bool le = (c0 != c1) ? (c0 < c1) : (s0 != start[i]) ? (s0 < start[i]) : (e0 <= e);
    if (!le) atomicOr(unsorted, 1);
  }
}

struct MaxOp {
  __device__ __forceinline__ int32_t operator()(int32_t a, int32_t b) const { return a > b ? a : b; }
};

size_t site_setup_temp_bytes(int32_t n) {
  size_t a = 0, b = 0;
  cub::DeviceScan::InclusiveScan((void*)nullptr, a, (const int32_t*)nullptr, (int32_t*)nullptr,
                                 MaxOp(), n);
  cub::DeviceScan::InclusiveScanByKey((void*)nullptr, b, (const int32_t*)nullptr,
                                      (const int32_t*)nullptr, (int32_t*)nullptr, MaxOp(), n);
  return (a > b ? a : b) + 256;
}

cudaError_t launch_site_setup(int32_t n, const int32_t* contig, const int32_t* start,
                              const int32_t* ref_len, int32_t* end, int32_t* pme, int32_t* d_unsorted,
                              void* d_temp, size_t temp_bytes, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(d_unsorted, 0, sizeof(int32_t), s);
  if (e != cudaSuccess || n == 0) return e;
  k_site_end<<<(n + 255) / 256, 256, 0, s>>>(n, contig, start, ref_len, end, d_unsorted);
  if (contig)
    e = cub::DeviceScan::InclusiveScanByKey(d_temp, temp_bytes, contig, (const int32_t*)end, pme,
                                            MaxOp(), n, cub::Equality(), s);
  else
    e = cub::DeviceScan::InclusiveScan(d_temp, temp_bytes, (const int32_t*)end, pme, MaxOp(), n, s);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

// index range [lo, hi) of contig c in a table sorted by contig
__global__ void k_contig_range(int32_t n, const int32_t* __restrict__ contig, int32_t c,
                               int32_t* __restrict__ range) {
  if (!contig) { range[0] = (c == 0) ? 0 : n; range[1] = (c == 0) ? n : n; return; }
  range[0] = lower_bound_i32(contig, 0, n, c);
  range[1] = lower_bound_i32(contig, range[0], n, c + 1);
}

cudaError_t launch_contig_range(int32_t n, const int32_t* contig, int32_t c, int32_t* d_range,
                                cudaStream_t s) {
  k_contig_range<<<1, 1, 0, s>>>(n, contig, c, d_range);
  return cudaGetLastError();
}

// ---- logFactorial table (BG:755-762): lnfact[n] = ln n + (ln(n-1) + ( ... + ln 2)) accumulated
// from n downwards exactly as the reference's tail recursion does, from the host-built ln table.
__global__ void k_lnfact(const double* __restrict__ lnk, double* __restrict__ lnfact, int32_t n) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double f = 0.0;
  for (int32_t k = i; k > 1; --k) f = lnk[k] + f;
  lnfact[i] = f;
}

cudaError_t launch_lnfact_table(const double* lnk, double* lnfact, int32_t n, cudaStream_t s) {
  k_lnfact<<<(n + 127) / 128, 128, 0, s>>>(lnk, lnfact, n);
  return cudaGetLastError();
}

// ---- coarse position index --------------------------------------------------------------------
int32_t index_bins(const BatchStats& h) {
  if (h.max_end <= h.min_start) return 1;
  const int64_t span = (int64_t)h.max_end - (int64_t)h.min_start;
  return (int32_t)((span + IDX_G - 1) / IDX_G) + 1;
}

__global__ void k_build_index(DReads R, DSites S, int32_t win0, int32_t nb, int32_t max_end,
                              int32_t* __restrict__ ridx, int32_t* __restrict__ sidx,
                              int32_t* __restrict__ s_range) {
  const int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k <= nb) {
    const int64_t key64 = (int64_t)win0 + (int64_t)k * IDX_G;
    const int32_t key = (int32_t)min(key64, (int64_t)INT32_MAX);
    if (ridx) ridx[k] = (k == nb) ? (int32_t)R.n : (int32_t)lower_bound_reads(R.meta, 0, R.n, key);
    if (S.n > 0) sidx[k] = lower_bound_i32(S.start, S.c_lo, S.c_hi, key);
  }
  if (k == 0 && S.n > 0) {
    // first site whose prefix-max end reaches past win0 (= first read start): earlier sites
    // cannot overlap any read of the batch
    int32_t lo = S.c_lo, hi = S.c_hi;
    while (lo < hi) {
      const int32_t mid = lo + ((hi - lo) >> 1);
      if (__ldg(S.pme + mid) <= win0) lo = mid + 1; else hi = mid;
    }
    s_range[0] = lo;
    s_range[1] = lower_bound_i32(S.start, S.c_lo, S.c_hi, max_end);
  }
}

cudaError_t launch_build_index(const DReads& R, const DSites& S, const BatchStats& h, int32_t* ridx,
                               int32_t* sidx, int32_t* s_range, cudaStream_t s) {
  const int32_t nb = index_bins(h);
  k_build_index<<<(nb + 1 + 255) / 256, 256, 0, s>>>(R, S, h.min_start, nb, h.max_end, ridx, sidx,
                                                      s_range);
  return cudaGetLastError();
}

// ---- compact wire form -> records: the offsets are prefix sums of the per-read sizes ----------------
struct Off3 { uint32_t b, o, f; };
struct Off3Sum {
  __host__ __device__ Off3 operator()(const Off3& x, const Off3& y) const { return Off3{x.b + y.b, x.o + y.o, x.f + y.f}; }
};
struct WireSizes {
  __host__ __device__ Off3 operator()(const avo_read_wire& w) const {
    return Off3{((uint32_t)w.l_seq + AVO_BLOCK_BASES - 1) / AVO_BLOCK_BASES, w.n_ops, w.n_refb};
  }
};

__global__ void k_wire_expand(const avo_read_wire* __restrict__ wire, const Off3* __restrict__ offs, int64_t n,
                              avo_read_meta* __restrict__ meta) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wire + r));
  const Off3 o = offs[r];
  const int32_t start = (int32_t)wv.x;
  const uint32_t span = wv.y & 0xFFFFu, l_seq = wv.y >> 16;
  const uint32_t n_ops = wv.z & 0xFFu, mapq = (wv.z >> 16) & 0xFFu, flags = wv.z >> 24;
  int4 a, b;
  a.x = start; a.y = start + (int32_t)span; a.z = (int32_t)o.b; a.w = (int32_t)o.o;
  b.x = (int32_t)o.f; b.y = (int32_t)(n_ops | (mapq << 16) | (flags << 24)); b.z = (int32_t)wv.w; b.w = (int32_t)l_seq;
  reinterpret_cast<int4*>(meta + r)[0] = a;
  reinterpret_cast<int4*>(meta + r)[1] = b;
}

__global__ void k_wire_ops(const uint16_t* __restrict__ w, int64_t n, uint32_t* __restrict__ ops) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) ops[i] = (uint32_t)__ldg(w + i);      // (length << 4) | code in both widths
}

size_t wire_temp_bytes(int64_t n) {
  size_t a = 0;
  cub::TransformInputIterator<Off3, WireSizes, const avo_read_wire*> it(nullptr, WireSizes());
  cub::DeviceScan::ExclusiveScan((void*)nullptr, a, it, (Off3*)nullptr, Off3Sum(), Off3{0, 0, 0}, (int)n);
  return a + 256;
}

cudaError_t launch_wire_expand(const avo_read_wire* wire, const uint16_t* wire_ops, int64_t n_reads,
                               int64_t n_ops, void* offs, avo_read_meta* meta, uint32_t* ops,
                               void* d_temp, size_t temp_bytes, cudaStream_t s) {
  if (n_reads > 0) {
    cub::TransformInputIterator<Off3, WireSizes, const avo_read_wire*> it(wire, WireSizes());
    cudaError_t e = cub::DeviceScan::ExclusiveScan(d_temp, temp_bytes, it, (Off3*)offs, Off3Sum(), Off3{0, 0, 0},
                                                   (int)n_reads, s);
    if (e != cudaSuccess) return e;
    k_wire_expand<<<(unsigned)((n_reads + 255) / 256), 256, 0, s>>>(wire, (const Off3*)offs, n_reads, meta);
  }
  if (n_ops > 0) k_wire_ops<<<(unsigned)((n_ops + 255) / 256), 256, 0, s>>>(wire_ops, n_ops, ops);
  return cudaGetLastError();
}

// ---- 6-byte wire form -> records: starts and operator offsets from one scan, then the sizes the
// operators imply (end, l_seq, payload blocks, refb bytes), exceptions, a second scan for the
// payload / refb offsets.
struct W6A { uint32_t o; uint32_t s; };          // operators, start difference
struct W6ASum { __host__ __device__ W6A operator()(const W6A& x, const W6A& y) const { return W6A{x.o + y.o, x.s + y.s}; } };
struct W6AIn {
  __host__ __device__ W6A operator()(const avo_read_wire6& w) const { return W6A{w.n_ops, (uint32_t)(w.dstart_flags >> 2)}; }
};
struct W6B { uint32_t b; uint32_t f; };          // payload blocks, refb bytes
struct W6BSum { __host__ __device__ W6B operator()(const W6B& x, const W6B& y) const { return W6B{x.b + y.b, x.f + y.f}; } };

// operators of the 6-byte form: a length byte and a code nibble each -> (length << 4) | code
__global__ void k_wire6_ops(const uint8_t* __restrict__ oplen, const uint8_t* __restrict__ opcode, int64_t n,
                            uint32_t* __restrict__ ops) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) ops[i] = ((uint32_t)__ldg(oplen + i) << 4) | (((uint32_t)__ldg(opcode + (i >> 1)) >> ((i & 1) * 4)) & 15u);
}

__global__ void k_wire6_records(const avo_read_wire6* __restrict__ wire, const W6A* __restrict__ offs,
                                const uint32_t* __restrict__ wops, int64_t n, int32_t base,
                                avo_read_meta* __restrict__ meta, W6B* __restrict__ sizes) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint16_t* w16 = reinterpret_cast<const uint16_t*>(wire + r);
  const uint32_t h0 = __ldg(w16), h1 = __ldg(w16 + 1), h2 = __ldg(w16 + 2);
  const W6A o = offs[r];
  const uint32_t d = h0 >> 2, flags = h0 & 3u, n_ops = h1 & 0xFFu, mapq = h1 >> 8;
  const int32_t start = base + (int32_t)(o.s + d);
  uint32_t span = 0, lseq = 0, nrefb = 0;
  for (uint32_t k = 0; k < n_ops; ++k) {
    const uint32_t op = __ldg(wops + o.o + k), len = op >> 4, code = op & 15u;
    span += (code <= AVO_OP_MISMATCH || code == AVO_OP_DEL) ? len : 0u;
    lseq += (code <= AVO_OP_MISMATCH || code == AVO_OP_INS || code == AVO_OP_SOFTCLIP) ? len : 0u;
    nrefb += code == AVO_OP_MISMATCH ? len : code == AVO_OP_DEL ? ((len + 1) >> 1) : 0u;
  }
  int4 a, b;
  a.x = start; a.y = start + (int32_t)span; a.z = 0; a.w = (int32_t)o.o;
  b.x = 0; b.y = (int32_t)(n_ops | (mapq << 16) | (flags << 24)); b.z = (int32_t)h2; b.w = (int32_t)lseq;   // qhead: q0, q1 only
  reinterpret_cast<int4*>(meta + r)[0] = a;
  reinterpret_cast<int4*>(meta + r)[1] = b;
  sizes[r] = W6B{(lseq + AVO_BLOCK_BASES - 1) / AVO_BLOCK_BASES, nrefb};
}

__global__ void k_wire6_exceptions(const avo_wire_exception* __restrict__ exc, int64_t n_exc, int64_t n,
                                   avo_read_meta* __restrict__ meta, W6B* __restrict__ sizes) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_exc) return;
  const avo_wire_exception e = exc[i];
  if ((int64_t)e.index >= n) return;
  meta[e.index].end = e.end;
  meta[e.index].l_seq = e.l_seq;
  sizes[e.index].b = (e.l_seq + AVO_BLOCK_BASES - 1) / AVO_BLOCK_BASES;
}

__global__ void k_wire6_offsets(const W6B* __restrict__ offs, int64_t n, avo_read_meta* __restrict__ meta) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  const W6B o = offs[r];
  meta[r].seq_off = o.b;
  meta[r].refb_off = o.f;
}

size_t wire6_temp_bytes(int64_t n) {
  size_t a = 0, b = 0;
  cub::TransformInputIterator<W6A, W6AIn, const avo_read_wire6*> it(nullptr, W6AIn());
  cub::DeviceScan::ExclusiveScan((void*)nullptr, a, it, (W6A*)nullptr, W6ASum(), W6A{0, 0}, (int)n);
  cub::DeviceScan::ExclusiveScan((void*)nullptr, b, (const W6B*)nullptr, (W6B*)nullptr, W6BSum(), W6B{0, 0}, (int)n);
  return (a > b ? a : b) + 256;
}

// offs: n x 8 bytes, sizes: n x 8 bytes (scanned in place)
cudaError_t launch_wire6_expand(const avo_read_wire6* wire, const uint8_t* oplen, const uint8_t* opcode,
                                const avo_wire_exception* exc, int64_t n_exc, int32_t base, int64_t n_reads,
                                int64_t n_ops, void* offs, void* sizes, avo_read_meta* meta, uint32_t* ops,
                                void* d_temp, size_t temp_bytes, cudaStream_t s, int* n_launches) {
  if (n_ops > 0) {
    k_wire6_ops<<<(unsigned)((n_ops + 255) / 256), 256, 0, s>>>(oplen, opcode, n_ops, ops);
    if (n_launches) *n_launches += 1;
  }
  if (n_reads > 0) {
    const unsigned grid = (unsigned)((n_reads + 255) / 256);
    cub::TransformInputIterator<W6A, W6AIn, const avo_read_wire6*> it(wire, W6AIn());
    cudaError_t e = cub::DeviceScan::ExclusiveScan(d_temp, temp_bytes, it, (W6A*)offs, W6ASum(), W6A{0, 0},
                                                   (int)n_reads, s);
    if (e != cudaSuccess) return e;
    k_wire6_records<<<grid, 256, 0, s>>>(wire, (const W6A*)offs, ops, n_reads, base, meta, (W6B*)sizes);
    if (n_exc > 0)
      k_wire6_exceptions<<<(unsigned)((n_exc + 255) / 256), 256, 0, s>>>(exc, n_exc, n_reads, meta, (W6B*)sizes);
    e = cub::DeviceScan::ExclusiveScan(d_temp, temp_bytes, (const W6B*)sizes, (W6B*)sizes, W6BSum(), W6B{0, 0},
                                       (int)n_reads, s);
    if (e != cudaSuccess) return e;
    k_wire6_offsets<<<grid, 256, 0, s>>>((const W6B*)sizes, n_reads, meta);
    if (n_launches) *n_launches += 6 + (n_exc > 0);
  }
  return cudaGetLastError();
}

// ---- union of K sorted batches' records (reads.union of CLI/TrioGenotyper.scala:216): every record finds
// its place in the merged order by counting, in each other part, the records that come before it
// (ties: the earlier part first, so the merge is stable), and is written there with its three offsets
// rebased to the concatenated payload / operator / mismatch columns.
__global__ void __launch_bounds__(256)
k_merge_meta(MergeParts M, avo_read_meta* __restrict__ out) {
  const int k = blockIdx.y;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M.n[k]) return;
  const int4 a = __ldg(reinterpret_cast<const int4*>(M.meta[k] + i));
  int4 b = __ldg(reinterpret_cast<const int4*>(M.meta[k] + i) + 1);
  const int32_t start = a.x;
  int64_t rank = i;
  for (int q = 0; q < M.k; ++q) {
    if (q == k) continue;
    int64_t lo = 0, hi = M.n[q];
    while (lo < hi) {                       // q < k: records with start <= mine; q > k: with start < mine
      const int64_t mid = lo + ((hi - lo) >> 1);
      const int32_t v = __ldg(&M.meta[q][mid].start);
      if (q < k ? v <= start : v < start) lo = mid + 1; else hi = mid;
    }
    rank += lo;
  }
  int4 a2 = a;
  a2.z = (int32_t)((uint32_t)a.z + M.block_base[k]);      // seq_off
  a2.w = (int32_t)((uint32_t)a.w + M.op_base[k]);         // op_off
  b.x = (int32_t)((uint32_t)b.x + M.refb_base[k]);        // refb_off
  reinterpret_cast<int4*>(out + rank)[0] = a2;
  reinterpret_cast<int4*>(out + rank)[1] = b;
}

cudaError_t launch_merge_meta(const MergeParts& M, avo_read_meta* out, cudaStream_t s) {
  int64_t mx = 0;
  for (int k = 0; k < M.k; ++k) mx = std::max(mx, M.n[k]);
  if (mx == 0) return cudaSuccess;
  k_merge_meta<<<dim3((unsigned)((mx + 255) / 256), (unsigned)M.k), 256, 0, s>>>(M, out);
  return cudaGetLastError();
}

// ---- max of an int32 column (device tables whose longest entry the host does not know) -------------------
__global__ void k_max_i32(const int32_t* __restrict__ a, int64_t n, int32_t* __restrict__ out) {
  int32_t m = INT32_MIN;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = max(m, a[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
  if ((threadIdx.x & 31) == 0 && m != INT32_MIN) atomicMax(out, m);
}
cudaError_t launch_max_i32(const int32_t* a, int64_t n, int32_t* d_out /* preset by the caller */, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  k_max_i32<<<(unsigned)std::min<int64_t>(1024, (n + 255) / 256), 256, 0, s>>>(a, n, d_out);
  return cudaGetLastError();
}

// ---- device -> device export of result columns: one launch instead of one copy per column ------
__global__ void k_copy_columns(ColumnCopies cc) {
  for (int c = blockIdx.y; c < cc.n; c += gridDim.y) {
    const size_t bytes = cc.bytes[c];
    const char* src = (const char*)cc.src[c];
    char* dst = (char*)cc.dst[c];
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((((uintptr_t)src | (uintptr_t)dst) & 15u) == 0) {
      const size_t n16 = bytes / 16;
      for (size_t k = i; k < n16; k += stride)
        reinterpret_cast<uint4*>(dst)[k] = reinterpret_cast<const uint4*>(src)[k];
      for (size_t k = n16 * 16 + i; k < bytes; k += stride) dst[k] = src[k];
    } else {
      for (size_t k = i; k < bytes; k += stride) dst[k] = src[k];
    }
  }
}

cudaError_t launch_copy_columns(const ColumnCopies& cc, cudaStream_t s) {
  if (cc.n == 0) return cudaSuccess;
  size_t mx = 0;
  for (int c = 0; c < cc.n; ++c) mx = cc.bytes[c] > mx ? cc.bytes[c] : mx;
  const int gx = (int)std::min<size_t>(64, (mx / 16 + 255) / 256 + 1);
  k_copy_columns<<<dim3(gx, cc.n), 256, 0, s>>>(cc);
  return cudaGetLastError();
}

cudaError_t launch_build_site_index(const DSites& S, const BatchStats& h, int32_t* sidx,
                                    int32_t* s_range, cudaStream_t s) {
  const int32_t nb = index_bins(h);
  DReads none{};
  k_build_index<<<(nb + 1 + 255) / 256, 256, 0, s>>>(none, S, h.min_start, nb, h.max_end, nullptr, sidx,
                                                      s_range);
  return cudaGetLastError();
}

}  // namespace avo
