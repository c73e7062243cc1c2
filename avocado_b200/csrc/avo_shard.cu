// avo_shard.cu -- the device side of region-sharded discoverAndCall (the reference's
// `sortByKey().collect` + `sc.broadcast` of the variants, util/TreeRegionJoin.scala:43-50, 184-185, and
// its only calling-path shuffle, genotyping/BiallelicGenotyper.scala:475-500).
//
// One process per GPU; rank g owns the sites that start in [own_lo, own_hi) and holds every read that
// can overlap one of them or put a discovery candidate on one (read-overlap halo).  Nothing is reduced
// across ranks; two all-gathers move (1) the owned slice of every rank's table, so that every rank joins
// its reads against the GLOBAL sorted table (the Forest look-up depends on global array order), and
// (2) the result rows of the owned sites.  The kernels here build and take apart the two fixed-capacity
// blocks the collectives move, entirely from device-side sizes: the host never waits between discovery,
// the first collective, the call stage and the second collective.
#include "avo_device.cuh"

namespace avo {

// ---- table block: ShardHdr + six int32 columns of `cap` rows + `acap` allele bytes ------------------
__host__ __device__ inline size_t shard_table_bytes(int64_t cap, int64_t acap) {
  return sizeof(ShardHdr) + (size_t)cap * 4 * 6 + (size_t)acap;
}

// first index in [0, n) with a[i] >= v
__device__ __forceinline__ int32_t lb_plain(const int32_t* __restrict__ a, int32_t n, int32_t v) {
  int32_t lo = 0, hi = n;
  while (lo < hi) {
    const int32_t mid = lo + ((hi - lo) >> 1);
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
k_shard_owned(ShardTableIn T, int32_t own_lo, int32_t own_hi, const int32_t* __restrict__ counters, ShardLocalCaps C,
              unsigned char* __restrict__ block, int32_t cap, int32_t acap) {
  const int32_t n_all = T.hdr->n;
  const int32_t n = min(n_all, T.cap);
  const int32_t a = lb_plain(T.start, n, own_lo), b = lb_plain(T.start, n, own_hi);
  const int32_t rows = b - a;
  const uint32_t nib_a = rows > 0 ? T.allele_off[a] : 0u;
  const uint32_t nib_b = rows > 0 ? (b < n ? T.allele_off[b] : min(T.hdr->n_nibbles, T.nib_cap)) : 0u;
  const int32_t bytes = (int32_t)((nib_b - nib_a) >> 1);
  ShardHdr* H = reinterpret_cast<ShardHdr*>(block);
  int32_t* col = reinterpret_cast<int32_t*>(block + sizeof(ShardHdr));
  unsigned char* alle = block + sizeof(ShardHdr) + (size_t)cap * 4 * 6;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int32_t flags = 0;
    // a capacity of this rank's own discovery was too small (or it failed): every rank must hear of it
    const int32_t dflags = counters[C.w_flags], n_out = counters[C.w_nout], n_indel = counters[C.w_nindel];
    if (n_out > C.ucap || n_indel > C.lcap || (dflags & DISC_FLAG_TABLE_FULL) || n_all > T.cap || T.hdr->n_nibbles > T.nib_cap)
      flags |= SHARD_RETRY;
    if (dflags & DISC_FLAG_UNSORTED) flags |= SHARD_UNSORTED;
    if (dflags & DISC_FLAG_BAD_STATS) flags |= SHARD_BAD_STATS;
    if (rows > cap || bytes > acap) flags |= SHARD_BLOCK_FULL;
    // furthest end of any site up to this rank's last owned one (sites the rank found below own_lo are
    // sites of earlier ranks too, so the value is exact as a carry for the ranks after it)
    H->n_rows = rows; H->n_bytes = bytes; H->flags = flags; H->max_end = rows > 0 ? T.pme[b - 1] : INT32_MIN;
    H->reserved[0] = H->reserved[1] = H->reserved[2] = H->reserved[3] = 0;
  }
  const int32_t rcopy = min(rows, cap), bcopy = min(bytes, acap);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rcopy; i += stride) {
    const int32_t j = a + (int32_t)i;
    col[0 * (int64_t)cap + i] = T.start[j];
    col[1 * (int64_t)cap + i] = T.ref_len[j];
    col[2 * (int64_t)cap + i] = T.alt_len[j];
    col[3 * (int64_t)cap + i] = T.count[j];
    col[4 * (int64_t)cap + i] = (int32_t)(T.allele_off[j] - nib_a);
    col[5 * (int64_t)cap + i] = T.pme[j];
  }
  const unsigned char* src = T.alleles4 + (nib_a >> 1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < bcopy; i += stride) alle[i] = src[i];
}

cudaError_t launch_shard_owned(const ShardTableIn& T, int32_t own_lo, int32_t own_hi, const int32_t* counters,
                               const ShardLocalCaps& C, void* block, int32_t cap, int32_t acap, int num_sms, cudaStream_t s) {
  k_shard_owned<<<num_sms, 256, 0, s>>>(T, own_lo, own_hi, counters, C, (unsigned char*)block, cap, acap);
  return cudaGetLastError();
}

// ---- the global table from the gathered blocks (rank order = position order) -----------------------
__global__ void __launch_bounds__(256)
k_shard_concat(const unsigned char* __restrict__ gathered, int32_t world, int32_t cap, int32_t acap, DSiteTableOut G) {
  const size_t bb = shard_table_bytes(cap, acap);
  // every thread adds up the (few) headers
  int32_t row0[AVO_MAX_SHARDS + 1], carry[AVO_MAX_SHARDS + 1];
  uint32_t byte0[AVO_MAX_SHARDS + 1];
  int32_t flags = 0;
  row0[0] = 0; byte0[0] = 0; carry[0] = INT32_MIN;
  for (int r = 0; r < world; ++r) {
    const ShardHdr* H = reinterpret_cast<const ShardHdr*>(gathered + (size_t)r * bb);
    row0[r + 1] = row0[r] + min(H->n_rows, cap);
    byte0[r + 1] = byte0[r] + (uint32_t)min(H->n_bytes, acap);
    carry[r + 1] = max(carry[r], H->max_end);
    flags |= H->flags;
  }
  const int32_t n = row0[world];
  const uint32_t nbytes = byte0[world];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = tid; i < n; i += stride) {
    int r = 0;
    while (i >= row0[r + 1]) ++r;
    const int32_t* col = reinterpret_cast<const int32_t*>(gathered + (size_t)r * bb + sizeof(ShardHdr));
    const int64_t k = i - row0[r];
    if (i < G.cap) {
      const int32_t st = col[0 * (int64_t)cap + k], rl = col[1 * (int64_t)cap + k];
      G.start[i] = st;
      G.ref_len[i] = rl;
      G.alt_len[i] = col[2 * (int64_t)cap + k];
      if (G.count) G.count[i] = col[3 * (int64_t)cap + k];
      G.allele_off[i] = (uint32_t)col[4 * (int64_t)cap + k] + 2u * byte0[r];
      G.end[i] = st + rl;
      // prefix-max end over the GLOBAL table: the furthest end among the earlier ranks' sites, or the
      // rank's own running maximum (what it found below its interval are sites of earlier ranks too)
      G.pme[i] = max(carry[r], col[5 * (int64_t)cap + k]);
    }
  }
  for (int64_t i = tid; i < nbytes; i += stride) {
    int r = 0;
    while (i >= byte0[r + 1]) ++r;
    if (2 * (uint64_t)i < G.nib_cap)
      G.alleles4[i] = gathered[(size_t)r * bb + sizeof(ShardHdr) + (size_t)cap * 4 * 6 + (size_t)(i - byte0[r])];
  }
  if (tid == 0) {
    if (n <= G.cap) G.allele_off[n] = 2u * nbytes;
    DSiteHdr h;
    h.n = n; h.n_nibbles = 2u * nbytes; h.s_begin = 0; h.s_end = n;
    int32_t mp = 1;                                              // TreeRegionJoin.scala:66-72
    while (!(2 * (int64_t)mp >= (int64_t)n)) mp *= 2;
    h.midpoint = mp;
    h.reserved[0] = flags; h.reserved[1] = h.reserved[2] = 0;
    *G.hdr = h;
  }
}

cudaError_t launch_shard_concat(const void* gathered, int32_t world, int32_t cap, int32_t acap, const DSiteTableOut& G,
                                int num_sms, cudaStream_t s) {
  k_shard_concat<<<num_sms, 256, 0, s>>>((const unsigned char*)gathered, world, cap, acap, G);
  return cudaGetLastError();
}

// site index of a header table against one batch's bins (the batch's reads were indexed at discovery)
__global__ void k_shard_site_index(DSites S_in, int32_t win0, int32_t nb, int32_t max_end, int32_t* __restrict__ sidx,
                                   DSiteHdr* __restrict__ hdr) {
  const DSites S = resolve_sites(S_in);
  const int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k <= nb) {
    const int64_t key64 = (int64_t)win0 + (int64_t)k * IDX_G;
    sidx[k] = lower_bound_i32(S.start, 0, S.n, (int32_t)min(key64, (int64_t)INT32_MAX));
  }
  if (k == 0) {
    // first site whose prefix-max end reaches past the batch's first read; first site at or after its last end
    int32_t lo = 0, hi = S.n;
    while (lo < hi) {
      const int32_t mid = lo + ((hi - lo) >> 1);
      if (__ldg(S.pme + mid) <= win0) lo = mid + 1; else hi = mid;
    }
    hdr->s_begin = lo;
    hdr->s_end = lower_bound_i32(S.start, 0, S.n, max_end);
  }
}

cudaError_t launch_shard_site_index(const DSites& S, const BatchStats& h, int32_t nb, int32_t* sidx, DSiteHdr* hdr,
                                    cudaStream_t s) {
  k_shard_site_index<<<(nb + 1 + 255) / 256, 256, 0, s>>>(S, h.min_start, nb, h.max_end, sidx, hdr);
  return cudaGetLastError();
}

// ---- result rows: the owned rows of every column into one block, and back -------------------------------
__host__ __device__ inline void shard_row_columns(const DResults& D, int ns, const void** src, int* width) {
  const void* p[AVO_RESULT_COLUMNS] = {D.emitted, D.allele_fwd, D.other_fwd, D.allele_cov, D.other_cov, D.total_cov, D.square_mapq,
                                       D.ref_ll, D.allele_ll, D.other_ll, D.nonref_ll, D.first_is_ref, D.copy_number, D.n_alt,
                                       D.n_ref, D.n_other, D.qual, D.gq, D.sb, D.fisher, D.rms_mapq, D.gl, D.nonref_gl};
  const int w[AVO_RESULT_COLUMNS] = {1, 4, 4, 4, 4, 4, 8, 8 * ns, 8 * ns, 8 * ns, 8 * ns, 1, 4, 4, 4, 4, 8, 4, 16, 4, 4, 8 * ns, 8 * ns};
  for (int c = 0; c < AVO_RESULT_COLUMNS; ++c) { src[c] = p[c]; width[c] = w[c]; }
}
__host__ __device__ inline size_t shard_rows_bytes(int64_t cap, int ns) {
  const int w[AVO_RESULT_COLUMNS] = {1, 4, 4, 4, 4, 4, 8, 8 * ns, 8 * ns, 8 * ns, 8 * ns, 1, 4, 4, 4, 4, 8, 4, 16, 4, 4, 8 * ns, 8 * ns};
  size_t off = sizeof(ShardHdr);
  for (int c = 0; c < AVO_RESULT_COLUMNS; ++c) off += ((size_t)cap * w[c] + 15) & ~(size_t)15;
  return off;
}
size_t shard_rows_block_bytes(int64_t cap, int ns) { return shard_rows_bytes(cap, ns); }
size_t shard_table_block_bytes(int64_t cap, int64_t acap) { return shard_table_bytes(cap, acap); }

// bytes [0, n) from src to dst by the whole grid; 16-byte accesses where the two share their alignment (the
// columns of a result set all start 256-byte aligned, so a row range has the same offset in both)
__device__ __forceinline__ void grid_copy(unsigned char* __restrict__ dst, const unsigned char* __restrict__ src, size_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((((uintptr_t)dst ^ (uintptr_t)src) & 15u) == 0) {
    const size_t head = min(n, (size_t)((16u - ((uintptr_t)dst & 15u)) & 15u));
    for (int64_t i = t0; i < (int64_t)head; i += stride) dst[i] = src[i];
    const size_t body = (n - head) >> 4;
    const uint4* s16 = reinterpret_cast<const uint4*>(src + head);
    uint4* d16 = reinterpret_cast<uint4*>(dst + head);
    for (int64_t i = t0; i < (int64_t)body; i += stride) d16[i] = s16[i];
    for (int64_t i = (int64_t)(head + (body << 4)) + t0; i < (int64_t)n; i += stride) dst[i] = src[i];
  } else if (((((uintptr_t)dst) | (uintptr_t)src) & 3u) == 0 && (n & 3u) == 0) {
    const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d4 = reinterpret_cast<uint32_t*>(dst);
    for (int64_t i = t0; i < (int64_t)(n >> 2); i += stride) d4[i] = s4[i];
  } else {
    for (int64_t i = t0; i < (int64_t)n; i += stride) dst[i] = src[i];
  }
}

// rows [A, A + n) of this rank (A, n from the gathered table headers) -> rows block; or (DIRECT) straight
// into the same rows of `dst`, the result columns of the rank that collects them -- its own arrays there,
// a peer mapping elsewhere: contiguous 16-byte stores over NVLink, one range per column
template <bool DIRECT>
__global__ void __launch_bounds__(256)
k_shard_pack_rows(DResults D, int ns, const unsigned char* __restrict__ gathered_tables, int32_t world, int32_t rank,
                  int32_t tcap, int32_t tacap, const uint64_t* __restrict__ slots_used, uint64_t slots_cap,
                  unsigned char* __restrict__ block, int32_t rcap, DResults dst, int64_t dst_rows) {
  const size_t bb = shard_table_bytes(tcap, tacap);
  int32_t A = 0, n = 0;
  for (int r = 0; r <= rank; ++r) {
    const ShardHdr* H = reinterpret_cast<const ShardHdr*>(gathered_tables + (size_t)r * bb);
    const int32_t k = min(H->n_rows, tcap);
    if (r < rank) A += k; else n = k;
  }
  ShardHdr* H = reinterpret_cast<ShardHdr*>(block);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    H->n_rows = n; H->n_bytes = 0; H->max_end = 0;
    H->flags = (n > rcap ? SHARD_BLOCK_FULL : 0) | ((slots_used && *slots_used > slots_cap) ? SHARD_RETRY : 0) |
               (DIRECT ? SHARD_ROWS_DIRECT : 0);
    H->reserved[0] = H->reserved[1] = H->reserved[2] = H->reserved[3] = 0;
  }
  const void* src[AVO_RESULT_COLUMNS];
  int width[AVO_RESULT_COLUMNS];
  shard_row_columns(D, ns, src, width);
  if (DIRECT) {
    const void* to[AVO_RESULT_COLUMNS];
    shard_row_columns(dst, ns, to, width);
    const int64_t rows = max((int64_t)0, min((int64_t)n, dst_rows - A));    // (a table past the arrays: the host reruns)
#pragma unroll 1
    for (int c = 0; c < AVO_RESULT_COLUMNS; ++c)
      grid_copy((unsigned char*)to[c] + (size_t)A * width[c], (const unsigned char*)src[c] + (size_t)A * width[c],
                (size_t)rows * width[c]);
    __threadfence_system();             // on their way to the peer before the kernel ends
    return;
  }
  const int32_t rows = min(n, rcap);
  size_t off = sizeof(ShardHdr);
#pragma unroll 1
  for (int c = 0; c < AVO_RESULT_COLUMNS; ++c) {
    grid_copy(block + off, (const unsigned char*)src[c] + (size_t)A * width[c], (size_t)rows * width[c]);
    off += ((size_t)rcap * width[c] + 15) & ~(size_t)15;
  }
}

cudaError_t launch_shard_pack_rows(const DResults& D, int ns, const void* gathered_tables, int32_t world, int32_t rank,
                                   int32_t tcap, int32_t tacap, const uint64_t* slots_used, uint64_t slots_cap, void* block,
                                   int32_t rcap, const DResults* direct_to, int64_t direct_rows, int num_sms, cudaStream_t s) {
  if (direct_to)
    k_shard_pack_rows<true><<<2 * num_sms, 256, 0, s>>>(D, ns, (const unsigned char*)gathered_tables, world, rank, tcap, tacap,
                                                       slots_used, slots_cap, (unsigned char*)block, rcap, *direct_to, direct_rows);
  else
    k_shard_pack_rows<false><<<2 * num_sms, 256, 0, s>>>(D, ns, (const unsigned char*)gathered_tables, world, rank, tcap, tacap,
                                                        slots_used, slots_cap, (unsigned char*)block, rcap, D, 0);
  return cudaGetLastError();
}

// gathered rows blocks -> the global result columns.  The rows of rank r start where its table rows start
// (gathered table headers); a block whose header says 0 rows is skipped (a rank that is not the root of a
// gather-to-root step holds only its own block).  blockIdx.y = rank; the block headers of both gathers are
// collected into hdrs_out ([0, world) tables, [world, 2 world) rows) for the host's one copy.
__global__ void __launch_bounds__(256)
k_shard_unpack_rows(const unsigned char* __restrict__ gathered_rows, const unsigned char* __restrict__ gathered_tables,
                    int32_t tcap, int32_t tacap, int32_t world, int32_t rcap, int ns, DResults O, int64_t out_rows,
                    ShardHdr* __restrict__ hdrs_out) {
  const size_t bb = shard_rows_bytes(rcap, ns), tbb = shard_table_bytes(tcap, tacap);
  const int r = blockIdx.y;
  const unsigned char* blk = gathered_rows + (size_t)r * bb;
  if (blockIdx.x == 0 && threadIdx.x < 16) {
    const int32_t* src = reinterpret_cast<const int32_t*>(threadIdx.x < 8 ? gathered_tables + (size_t)r * tbb : blk);
    reinterpret_cast<int32_t*>(hdrs_out + (threadIdx.x < 8 ? r : world + r))[threadIdx.x & 7] = src[threadIdx.x & 7];
  }
  int64_t A = 0;
  int32_t n_table = 0;
  for (int q = 0; q <= r; ++q) {
    A += n_table;
    n_table = min(reinterpret_cast<const ShardHdr*>(gathered_tables + (size_t)q * tbb)->n_rows, tcap);
  }
  const int32_t n = min(min(reinterpret_cast<const ShardHdr*>(blk)->n_rows, rcap), n_table);
  const int64_t rows = max((int64_t)0, min((int64_t)n, out_rows - A));
  if (rows == 0 || (reinterpret_cast<const ShardHdr*>(blk)->flags & SHARD_ROWS_DIRECT)) return;
  const void* dstp[AVO_RESULT_COLUMNS];
  int width[AVO_RESULT_COLUMNS];
  shard_row_columns(O, ns, dstp, width);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t off = sizeof(ShardHdr);
#pragma unroll 1
  for (int c = 0; c < AVO_RESULT_COLUMNS; ++c) {
    const size_t bytes = (size_t)rows * width[c];
    const unsigned char* s8 = blk + off;
    unsigned char* d8 = (unsigned char*)dstp[c] + (size_t)A * width[c];
    if (((((uintptr_t)s8) | (uintptr_t)d8) & 3u) == 0 && (width[c] & 3) == 0) {
      const uint32_t* s4 = (const uint32_t*)s8;
      uint32_t* d4 = (uint32_t*)d8;
      for (int64_t i = t0; i < (int64_t)(bytes >> 2); i += stride) d4[i] = s4[i];
    } else {
      for (int64_t i = t0; i < (int64_t)bytes; i += stride) d8[i] = s8[i];
    }
    off += ((size_t)rcap * width[c] + 15) & ~(size_t)15;
  }
}

cudaError_t launch_shard_unpack_rows(const void* gathered_rows, const void* gathered_tables, int32_t tcap, int32_t tacap,
                                     int32_t world, int32_t rcap, int ns, const DResults& O, int64_t out_rows, ShardHdr* hdrs_out,
                                     int num_sms, cudaStream_t s) {
  const int bx = max(1, min(2 * num_sms / max(world, 1), (rcap + 255) / 256));
  k_shard_unpack_rows<<<dim3(bx, world), 256, 0, s>>>((const unsigned char*)gathered_rows, (const unsigned char*)gathered_tables,
                                                      tcap, tacap, world, rcap, ns, O, out_rows, hdrs_out);
  return cudaGetLastError();
}

}  // namespace avo
