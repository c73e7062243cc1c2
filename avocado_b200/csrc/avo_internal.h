// avo_internal.h -- device-side views and launcher prototypes shared by the .cu files.
// Not part of the public ABI (that is include/avocado_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/avocado_b200.h"

namespace avo {

// ---- device views (all pointers are device pointers) ----------------------------------------
struct DReads {
  int64_t n;
  int32_t contig;
  const avo_read_meta* meta;   // 32-byte record per read (include/avocado_b200.h)
  int2* coords;                // (start, end) per read: dense copy written by the first kernel that
                               // streams the records (k_batch_stats / k_discover_tiles), read by the join
  const uint8_t* payload;      // 16-byte blocks: 10 bases (4 bit) + their 10 qualities
  const uint32_t* ops;
  const uint8_t* refb;
  int32_t qhead_short;         // 1: only the first two bytes of meta.qhead are valid (6-byte wire form)
  int64_t n_ops, n_refb;       // entries of ops / bytes of refb (bounds of the staged copies)
  // candidate events (avo_read_events), when the batch carries them: discovery counts these instead of walking
  int64_t n_snv, n_indel;
  const int32_t* snv_pos;
  const uint16_t* snv_pq;
  const uint32_t* snv_first;
  const int32_t* indel_read;
  const uint32_t* indel_op;
};
static_assert(sizeof(avo_read_meta) == 32, "avo_read_meta must be one 32-byte sector");

// What the device knows about a site table that the host has not seen yet (a table assembled by
// k_assemble_sites in the same launch chain): kernels read their sizes from here.
struct DSiteHdr {
  int32_t n;           // rows (may exceed the capacity the table was assembled into: the host reruns)
  int32_t midpoint;    // Forest.midpoint over those rows
  uint32_t n_nibbles;  // allele nibbles (same remark)
  int32_t s_begin, s_end;   // DIndex.s_range of the table against its own batch: {0, n}
  int32_t reserved[3];
};

struct DSites {
  int32_t n;          // total entries in the table (all contigs)
  int32_t c_lo, c_hi; // index range of the batch's contig
  int32_t midpoint;   // Forest.midpoint (TreeRegionJoin.scala:66-72) over the whole table
  const DSiteHdr* hdr; // not null: n / c_hi / midpoint above are capacities, the real ones are here
  int32_t cap;         // rows the arrays hold (hdr tables)
  const int32_t* contig; // may be null (all zero)
  const int32_t* start;
  const int32_t* end;    // start + ref_len (built by k_site_setup)
  const int32_t* ref_len;
  const int32_t* alt_len;
  const uint32_t* allele_off;
  const uint8_t* alleles4;
  const int32_t* pme;    // prefix max of `end` within the contig (inclusive)
};

// Coarse position index over one batch (bins of IDX_G bases from win0): built by k_build_index.
//   ridx[k] = first read with start >= win0 + k*IDX_G      (k = 0..nb, ridx[nb] = n_reads)
//   sidx[k] = first site of the batch contig with start >= win0 + k*IDX_G
//   s_begin = first site of the contig whose prefix-max end reaches past the first read's start
//   s_end   = first site of the contig with start >= max read end
constexpr int IDX_G = 256;
struct DIndex {
  int32_t win0;
  int32_t nb;
  const int32_t* ridx;
  const int32_t* sidx;
  const int32_t* s_range;  // [2] = {s_begin, s_end}
};

struct DCnv {   // CNVs of the batch's contig only, sorted by (start, end)
  int32_t n;
  const int32_t* start;
  const int32_t* end;
  const int32_t* cn;
};

struct DResults {
  int32_t ns; // n_states (row stride)
  uint8_t* emitted;
  int32_t *allele_fwd, *other_fwd, *allele_cov, *other_cov, *total_cov;
  double* square_mapq;
  double *ref_ll, *allele_ll, *other_ll, *nonref_ll;
  uint8_t* first_is_ref;
  int32_t* copy_number;
  int32_t *n_alt, *n_ref, *n_other;
  double* qual;
  int32_t* gq;
  int32_t* sb;
  float *fisher, *rms_mapq;
  double *gl, *nonref_gl;
};

struct DCallParams {
  int32_t base_ploidy, min_ploidy, max_ploidy, ns;
  int32_t max_quality, max_mapq;
  // lut[(((cn - min_ploidy) * 128 + mq') * 128 + q') * ns + g]   (S6 of SURVEY.md)
  const double* lut;
  const double* lnk;     // lnk[k] = log(k), k < lnk_n (host libm)
  const double* lnfact;  // lnfact[n] = logFactorial(n) (BG:755-762), n < lnk_n, same summation order
  int32_t lnk_n;
  double ten_div_log10;   // 10 / ln 10   (BG:572)
  double log10;           // ln 10        (LogPhred.scala)
};

// ---- joint calling (avo_joint.cu): S samples x N sites, sample-major -------------------------
struct DJointIn {
  int64_t n_sites;
  int32_t n_samples, n_states;
  const uint8_t* present;   // may be null (all present)
  const int32_t *n_alt, *n_ref, *n_other;
  const int32_t* n_nocall;  // may be null
  const double* gl;
};
struct DJointOut {
  uint8_t* site_emitted;
  double *maf, *quality, *post0_sum;
  uint8_t* flags;
  float *posteriors, *priors;
  int32_t *n_alt, *n_ref, *gq;
};
cudaError_t launch_joint_counts(const DJointIn& J, int32_t* alt, int32_t* called, cudaStream_t s);
struct DDepthsIn {
  int64_t n_sites; int32_t n_samples;
  const uint8_t* present; const int32_t *read_depth, *ref_read_depth, *sb;    // each may be null
};
struct DDepthsOut { uint8_t* annotated; int32_t* col[6]; };
cudaError_t launch_joint_depths(const DDepthsIn& I, const DDepthsOut& O, cudaStream_t s);
cudaError_t launch_joint_recall(const DJointIn& J, const int32_t* alt, const int32_t* called,
                                const DJointOut& O, const double* lgam, cudaStream_t s);

// ---- RewriteHets + HardFilterGenotypes (avo_filter.cu) ------------------------------------------
struct DFilterIn {
  int64_t n;
  const uint8_t* emitted;
  const int32_t *ref_len, *alt_len;     // alt_len may be null (no alternate allele anywhere)
  const int32_t *n_alt, *n_ref, *n_other, *gq, *total_cov, *allele_cov;
  const float *fisher, *rms_mapq;
};
struct DFilterOut {
  uint8_t *emit, *rewritten, *filters_applied;
  uint32_t* filters_failed;
  int32_t *n_alt, *n_ref;
  uint8_t* gq_valid;
};
cudaError_t launch_filter_genotypes(const DFilterIn& I, const avo_filter_params& A, const DFilterOut& O,
                                    cudaStream_t s);

// ---- SquareOffReferenceModel (avo_squareoff.cu) ---------------------------------------------------
struct DSquareIn {
  int32_t n_sites, n_v, n_r, ns, v_max_ref_len, r_max_len;
  const int32_t *c_start, *c_ref_len, *c_alt_len; const uint32_t* c_allele_off; const uint8_t* c_alleles4;
  const int32_t *v_start, *v_ref_len, *v_alt_len; const uint32_t* v_allele_off; const uint8_t* v_alleles4;
  const uint8_t* v_emitted; const int32_t *v_n_alt, *v_n_ref, *v_n_other; const double *v_gl, *v_nonref_gl;
  const int32_t *r_start, *r_len /* NULL: 1 */; const uint8_t* r_emitted; const int32_t *r_n_alt, *r_n_ref, *r_n_other;
  const double* r_nonref_gl;
};
struct DSquareOut {
  uint8_t* kind; int32_t* src; int32_t *n_alt, *n_ref, *n_other; double* gl;
};
cudaError_t launch_square_off(const DSquareIn& I, const DSquareOut& O, cudaStream_t s);

// ---- TrioCaller (avo_trio.cu) ------------------------------------------------------------------------
struct DTrioIn { int64_t n_sites; const uint8_t* present; const int32_t *n_alt, *n_ref, *n_other, *n_nocall; };
struct DTrioOut { uint8_t *kept, *child_action, *filled; };
cudaError_t launch_trio(const DTrioIn& I, const DTrioOut& O, cudaStream_t s);

// batch statistics produced by k_batch_stats
struct BatchStats {
  int32_t min_start;
  int32_t max_end;
  int32_t max_span;   // max(end - start)
  int32_t unsorted;   // != 0 if start[] is not non-decreasing
};

size_t wire6_temp_bytes(int64_t n_reads);
cudaError_t launch_wire6_expand(const avo_read_wire6* wire, const uint8_t* oplen, const uint8_t* opcode,
                                const avo_wire_exception* exc,
                                int64_t n_exc, int32_t base, int64_t n_reads, int64_t n_ops, void* offs /* 8 B x n */,
                                void* sizes /* 8 B x n */, avo_read_meta* meta, uint32_t* ops, void* d_temp,
                                size_t temp_bytes, cudaStream_t s, int* n_launches);
// compact wire columns -> full records / operators on the device (avo_setup.cu)
size_t wire_temp_bytes(int64_t n_reads);
cudaError_t launch_wire_expand(const avo_read_wire* wire, const uint16_t* wire_ops, int64_t n_reads,
                               int64_t n_ops, void* offs /* 12 B x n_reads */, avo_read_meta* meta,
                               uint32_t* ops, void* d_temp, size_t temp_bytes, cudaStream_t s);

// device -> device column copies done by one launch
// ---- the packer on the device (avo_packdev.cu) ------------------------------------------------------
struct DText {                   // avo_text_reads with device pointers
  int64_t n;
  const int64_t *start, *end;
  const int32_t* mapq;
  const uint8_t* rec_flags;
  const char* cigar; const int64_t* cigar_off;
  const char* md;    const int64_t* md_off;
  const char* seq;   const int64_t* seq_off;
  const char* qual;  const int64_t* qual_off;
  const uint8_t* keep;           // may be NULL
  int32_t prefilter;             // != 0: pf holds PrefilterReads' read predicates
  avo_prefilter pf;
  int32_t narrow;                // != 0: start / end are int32, the offsets uint32, mapq uint8 (avo_text_reads.narrow)
};
// column accessors of both packers (T: DText or avo_text_reads)
#define AVO_T_I64(T, col, i) ((T).narrow ? (int64_t)reinterpret_cast<const int32_t*>((T).col)[i] : (T).col[i])
#define AVO_T_OFF(T, col, i) ((T).narrow ? (int64_t)reinterpret_cast<const uint32_t*>((T).col)[i] : (T).col[i])
#define AVO_T_MAPQ(T, i) ((T).narrow ? (int32_t)reinterpret_cast<const uint8_t*>((T).mapq)[i] : (T).mapq[i])
struct PackOffsets { uint64_t k, b, o, f; };   // packed record, payload block, operator, mismatch byte
struct DPacked {
  avo_read_meta* meta;
  uint8_t* payload;
  uint32_t* ops;
  uint8_t* refb;
  int64_t* source_index;         // may be NULL
};
size_t pack_scan_temp_bytes(int64_t n);
// counts[n + 1] and offsets[n + 1]: offsets[n] holds the totals
cudaError_t launch_pack_count(const DText& T, uint4* counts, PackOffsets* offsets, void* temp, size_t temp_bytes,
                              cudaStream_t s, int* n_launches);
cudaError_t launch_pack_write(const DText& T, const PackOffsets* offsets, const DPacked& out, cudaStream_t s,
                              int* n_launches);

struct ColumnCopies {
  int n;
  void* dst[32];
  const void* src[32];
  size_t bytes[32];
};
cudaError_t launch_copy_columns(const ColumnCopies& cc, cudaStream_t s);
constexpr int AVO_MAX_MERGE = 8;
struct MergeParts {
  int k;
  const avo_read_meta* meta[AVO_MAX_MERGE];
  int64_t n[AVO_MAX_MERGE];
  uint32_t block_base[AVO_MAX_MERGE], op_base[AVO_MAX_MERGE], refb_base[AVO_MAX_MERGE];
};
cudaError_t launch_merge_meta(const MergeParts& M, avo_read_meta* out, cudaStream_t s);
cudaError_t launch_max_i32(const int32_t* a, int64_t n, int32_t* d_out, cudaStream_t s);

// ---- region-sharded discoverAndCall (avo_shard.cu) ---------------------------------------------------
constexpr int AVO_MAX_SHARDS = 64;
constexpr int AVO_RESULT_COLUMNS = 23;
enum : int32_t { SHARD_RETRY = 1, SHARD_UNSORTED = 2, SHARD_BAD_STATS = 4, SHARD_BLOCK_FULL = 8, SHARD_ROWS_DIRECT = 16 };
struct ShardHdr {          // first 32 bytes of a table block and of a rows block
  int32_t n_rows;          // rows this rank owns (may exceed the block's capacity: SHARD_BLOCK_FULL)
  int32_t n_bytes;         // their allele bytes
  int32_t flags;           // SHARD_*
  int32_t max_end;         // table block: furthest site end up to the rank's last owned row
  int32_t reserved[4];
};
static_assert(sizeof(ShardHdr) == 32, "block header");
struct ShardTableIn {      // a rank's own table (assembled by k_assemble_sites)
  const DSiteHdr* hdr; int32_t cap; uint32_t nib_cap;
  const int32_t *start, *ref_len, *alt_len, *count, *pme; const uint32_t* allele_off; const uint8_t* alleles4;
};
struct ShardLocalCaps { int32_t w_flags, w_nout, w_nindel; int32_t ucap, lcap; };   // counter words, capacities
size_t shard_table_block_bytes(int64_t cap, int64_t acap);
size_t shard_rows_block_bytes(int64_t cap, int ns);
cudaError_t launch_shard_owned(const ShardTableIn& T, int32_t own_lo, int32_t own_hi, const int32_t* counters,
                               const ShardLocalCaps& C, void* block, int32_t cap, int32_t acap, int num_sms, cudaStream_t s);
struct DSiteTableOut;
cudaError_t launch_shard_concat(const void* gathered, int32_t world, int32_t cap, int32_t acap, const DSiteTableOut& G,
                                int num_sms, cudaStream_t s);
cudaError_t launch_shard_site_index(const DSites& S, const BatchStats& h, int32_t nb, int32_t* sidx, DSiteHdr* hdr,
                                    cudaStream_t s);
cudaError_t launch_shard_pack_rows(const DResults& D, int ns, const void* gathered_tables, int32_t world, int32_t rank,
                                   int32_t tcap, int32_t tacap, const uint64_t* slots_used, uint64_t slots_cap, void* block,
                                   int32_t rcap, const DResults* direct_to, int64_t direct_rows, int num_sms, cudaStream_t s);
cudaError_t launch_shard_unpack_rows(const void* gathered_rows, const void* gathered_tables, int32_t tcap, int32_t tacap,
                                     int32_t world, int32_t rcap, int ns, const DResults& O, int64_t out_rows, ShardHdr* hdrs_out,
                                     int num_sms, cudaStream_t s);

// ---- launchers (each returns the cudaError_t of its launches) --------------------------------
cudaError_t launch_batch_stats(const DReads& R, BatchStats* d_stats, cudaStream_t s);
cudaError_t launch_site_setup(int32_t n, const int32_t* contig, const int32_t* start,
                              const int32_t* ref_len, int32_t* end, int32_t* pme, int32_t* d_unsorted,
                              void* d_temp, size_t temp_bytes, cudaStream_t s);
size_t site_setup_temp_bytes(int32_t n);

cudaError_t launch_lnfact_table(const double* lnk, double* lnfact, int32_t n, cudaStream_t s);
// index over reads (and, when S.n > 0, sites); ridx/sidx need nb+1 entries, nb = index_bins(stats)
int32_t index_bins(const BatchStats& hstats);
cudaError_t launch_build_index(const DReads& R, const DSites& S, const BatchStats& hstats,
                               int32_t* ridx, int32_t* sidx, int32_t* s_range, cudaStream_t s);
// sites only (the read index of the same batch already exists)
cudaError_t launch_build_site_index(const DSites& S, const BatchStats& hstats, int32_t* sidx,
                                    int32_t* s_range, cudaStream_t s);

// call: per-site buckets (one slot per read that can overlap the site), observe, reduce
size_t call_scan_temp_bytes(int32_t n);
cudaError_t launch_call_buckets(const DReads& R, const DSites& S, const DIndex& X, const BatchStats& hstats, int32_t* rbase,
                                uint32_t* blen, uint64_t* boff, void* d_temp, size_t temp_bytes,
                                cudaStream_t s, int* n_launches);
// join -> (optional host gather of the payload blocks the joined reads need) -> observe + reduce
cudaError_t launch_call_join(const DReads& R, const DSites& S, const DIndex& X, int32_t max_span /* >= every end - start */,
                             uint32_t* buckets, size_t n_slots, void* joined_list /* 12 B x n_reads */,
                             int32_t* d_counter, int num_sms, cudaStream_t s, int* n_launches);
size_t call_gather_temp_bytes(int64_t n);
cudaError_t launch_call_pairs(const void* joined_list, int32_t n_joined, uint32_t* pbase /* n_joined + 1 */,
                              void* d_temp, size_t temp_bytes, cudaStream_t s, int* n_launches);
cudaError_t launch_call_gather_plan(const DReads& R, const DSites& S, const void* joined_list, int32_t n_joined,
                                    const uint32_t* pbase, int64_t n_pairs, uint32_t* nblk, uint32_t* slot,
                                    uint32_t* clo, void* d_temp, size_t temp_bytes, cudaStream_t s, int* n_launches);
cudaError_t launch_call_blocklist(const DReads& R, const void* joined_list, int32_t n_joined, const uint32_t* pbase,
                                  const uint32_t* nblk, const uint32_t* slot, const uint32_t* clo, uint32_t* idx,
                                  uint32_t* blk0, cudaStream_t s, int* n_launches);
cudaError_t launch_call_pairs_reduce(const DReads& R, const DSites& S, const DIndex& X, const DCnv& C, const DCallParams& P,
                                     const DResults& out, const int32_t* rbase, const uint32_t* blen, const uint64_t* boff,
                                     uint32_t* buckets, size_t n_slots, int num_sms, cudaStream_t s, int* n_launches,
                                     cudaEvent_t before_pairs = nullptr, cudaEvent_t after_pairs = nullptr);
cudaError_t launch_call_observe(const DReads& R, const DSites& S, const DCnv& C, const DCallParams& P,
                                const DResults& out, const int32_t* rbase, const uint32_t* blen,
                                const uint64_t* boff, uint32_t* buckets, size_t n_slots /* capacity of buckets */,
                                const void* joined_list, const int32_t* d_counter, const uint8_t* compact,
                                const uint32_t* pbase, const uint32_t* blk0, int num_sms, cudaStream_t s,
                                int* n_launches);

// scoreAllSites reference-model rows (avo_allsites.cu).  All pointers are device scratch.
struct AllSitesPlan {
  int64_t p0;          // reference position of bit 0 (min read start - 1)
  int64_t nbits;       // positions covered by the bitmaps (multiple of 256)
  int32_t n_win;       // nbits / 256
  int32_t* flag;       // [n_reads + 1] joined flag per read
  int32_t* jpos;       // [n_reads + 1] exclusive scan of flag
  void* jr;            // [n_joined] (read, a, b)
  int2* jcoord;        // [n_joined] (start, end) of the joined reads
  int32_t* jfine;      // [nbits / 32 + 1] first joined read starting at or after p0 + 32 k
  uint32_t* cover;     // positions that may receive a row
  uint32_t* site_cover;// positions covered by a variant site
  int32_t* win_count;  // [n_win + 1]
  int32_t* win_base;   // [n_win + 1] exclusive scan: first row of every 256-position window
};
size_t allsites_scan_temp_bytes(int64_t n);
cudaError_t launch_allsites_plan(const DReads& R, const DSites& S, const DIndex& X, const AllSitesPlan& A,
                                 void* d_temp, size_t temp_bytes, cudaStream_t s, int* n_launches);
cudaError_t launch_allsites_tiles(const DReads& R, const DSites& S, const DCnv& C, const DCallParams& P,
                                  const DResults& O, int32_t* o_start, const DIndex& X,
                                  const AllSitesPlan& A, int32_t n_joined, int32_t max_span,
                                  cudaStream_t s, int* n_launches);

// discovery
// per 2048-base tile of the genome: the survivors that start in it, as a linked list
struct TileBucket {
  int32_t head;         // last survivor linked in + 1 (0: none)
  int32_t cnt;          // survivors
  uint32_t nib;         // their allele nibbles (even-aligned footprints)
  int32_t max_rel_end;  // max(end - win0) over them (0: none)
};
struct DDiscoverOut {
  // unsorted candidate survivors appended by the kernels
  int32_t capacity;
  int32_t* n;           // device counter (may exceed capacity: overflow)
  int32_t* start;
  int32_t* ref_len;
  int32_t* alt_len;
  int32_t* count;
  uint64_t* src;        // where the allele content lives: see avo_discover.cu
  int32_t* next;        // next survivor of the same tile + 1 (0: end of list)
  TileBucket* tiles;    // [n_tiles], zeroed before the tile kernel
  int32_t win0;         // reference position of tile 0
};

struct DIndelList {
  int32_t capacity;
  int32_t* n;           // device counter
  int32_t* pos;
  int32_t* ref_len;     // full 32-bit lengths: operators are 28-bit (a 70 kb deletion is one candidate)
  int32_t* alt_len;
  uint32_t* nibs;       // (lastRef nibble << 4) | anchor (alt[0] / read base) nibble
  uint32_t* src;        // INS: index (within the read) of the first inserted base; DEL: nibble index (2 * byte) into refb
  uint32_t* blk;        // INS: the read's first payload block
  uint32_t* hash;
};

// d_flags: DISC_FLAG_* bits set by the kernels
enum : int32_t { DISC_FLAG_UNSORTED = 1, DISC_FLAG_BAD_STATS = 2, DISC_FLAG_TABLE_FULL = 4 };
// verify != 0: every record is checked against hstats and the sort order while it is walked
// candidate events: count per 32 reads (+ exact bounds, sort order), then write in read order
cudaError_t launch_events_count(const DReads& R, uint32_t* c_snv, uint32_t* c_ind, BatchStats* d_stats, cudaStream_t s);
size_t events_scan_temp_bytes(int64_t n);
cudaError_t launch_events_scan(const uint32_t* counts, uint32_t* first, int64_t n, void* d_temp, size_t temp_bytes, cudaStream_t s);
cudaError_t launch_events_write(const DReads& R, const uint32_t* snv_first, const uint32_t* ind_first, int64_t snv_cap,
                                int64_t ind_cap, int32_t* snv_pos, uint16_t* snv_pq, int32_t* indel_read, uint32_t* indel_op,
                                cudaStream_t s);
cudaError_t launch_discover_events(const DReads& R, int32_t thr, int32_t min_obs, const DIndex& X, const BatchStats& hstats,
                                   const DDiscoverOut& out, const DIndelList& indels, int32_t* d_tile_counter, int32_t* d_flags,
                                   int num_sms, int* blocks_per_sm_cache, cudaStream_t s, int* n_launches);
cudaError_t launch_discover_tiles(const DReads& R, int32_t thr, int32_t min_obs /* <0: distinct */,
                                  const DIndex& X, const BatchStats& hstats, int verify,
                                  const DDiscoverOut& out, const DIndelList& indels,
                                  int32_t* d_tile_counter, int32_t* d_flags, int num_sms,
                                  int* blocks_per_sm, cudaStream_t s, int* n_launches);

// groups the *L.n (device counter) indel candidates; survivors are appended to `out` and linked into
// their tiles.  Sets DISC_FLAG_TABLE_FULL (and does nothing) when the table is too small for *L.n.
cudaError_t launch_indel_group(const DReads& R, const DIndelList& L, unsigned long long* table,
                               int32_t* tcount, uint32_t tsize, int32_t min_obs,
                               const DDiscoverOut& out, int32_t* d_flags, int num_sms, cudaStream_t s,
                               int* n_launches);
int32_t discover_tiles(const BatchStats& hstats);   // tiles k_discover_tiles cuts the batch into
// Sorted site table from the per-tile survivor lists, in one launch: rows in (start, ref_len, alt_len,
// alleles) order, derived columns (end, prefix-max end), the site index of the batch's bins and the
// header.  Rows / nibbles beyond the capacities are counted in the header but not written.
struct DSiteTableOut {
  int32_t cap; uint32_t nib_cap;
  int32_t *start, *ref_len, *alt_len, *count, *end, *pme;
  uint32_t* allele_off;      // [cap + 1]
  uint8_t* alleles4;
  int32_t* sidx;             // [nb + 1]
  DSiteHdr* hdr;
};
cudaError_t launch_assemble_sites(const DReads& R, const DIndelList& L, const DDiscoverOut& U,
                                  const BatchStats& hstats, int32_t nb, const DSiteTableOut& T,
                                  cudaStream_t s, int* n_launches);
cudaError_t launch_contig_range(int32_t n, const int32_t* contig, int32_t c, int32_t* d_range,
                                cudaStream_t s);

}  // namespace avo
