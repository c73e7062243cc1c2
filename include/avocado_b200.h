/* avocado_b200.h -- C ABI of the B200-native replacement for avocado's genotyping hot path.
 *
 * The reference (bigdatagenomics/avocado, Scala on Spark) has no FFI; the seam this library
 * sits behind is the pair of Scala entry points
 *
 *   DiscoverVariants.apply(reads, optPhredThreshold, optMinObservations): VariantRDD
 *       avocado-core/src/main/scala/org/bdgenomics/avocado/genotyping/DiscoverVariants.scala:51-61
 *   BiallelicGenotyper.call(reads, variants, copyNumber, scoreAllSites, ..., maxQuality, maxMapQ): GenotypeRDD
 *       .../genotyping/BiallelicGenotyper.scala:88-135      (discoverAndCall: :156-184)
 *   JointAnnotatorCaller.apply(genotypes): VariantContextRDD
 *       .../genotyping/JointAnnotatorCaller.scala:52-64
 *
 * A JVM shim keeps those signatures and, per partition / region bin, packs reads into the
 * columnar batch below, calls one of the avo_* functions through JNI, and fills Avro builders
 * from the SoA results without arithmetic (INTEGRATION.md shows the binding).
 *
 * Conventions
 *  - Every buffer is caller-owned; the library keeps no pointer after a call returns.
 *  - `mem` says where a struct's buffers live: AVO_MEM_HOST or AVO_MEM_DEVICE (pointers valid on
 *    the context's GPU).  Host batches should be page-locked (cudaHostAlloc / cudaHostRegister,
 *    e.g. the shim's direct ByteBuffers registered once): the per-read records, operators and
 *    mismatch bytes are then copied at full PCIe rate and the large column (payload)
 *    is read in place, only the 32-byte sectors the kernels touch (avo_params.host_payload).
 *  - All functions return 0 (AVO_OK) or a negative avo_status; nothing throws or aborts.
 *    A malformed read is skipped, never an error (BiallelicGenotyper.scala:385-391).
 *  - A context is bound to one GPU and is NOT thread-safe; use one per thread.
 *  - Coordinates are 0-based half-open (ADAM); bases are 4-bit BAM codes (=ACMGRSVTWYHKDBN),
 *    two per byte, even index in the HIGH nibble; qualities are raw phred (no +33).
 *  - Bases and qualities of a read travel together in 16-byte payload blocks (AVO_BLOCK_BASES = 10
 *    bases per block: bytes 0..4 the 4-bit bases, bytes 5..14 the qualities, byte 15 zero), so the
 *    base and the quality of a position come with ONE 16-byte load.
 *  - There is no CPU fallback: without a usable CUDA device avo_ctx_create fails.
 */
#ifndef AVOCADO_B200_H
#define AVOCADO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct avo_ctx avo_ctx;

typedef enum {
  AVO_OK = 0,
  AVO_E_INVALID = -1,      /* bad argument / inconsistent struct */
  AVO_E_CUDA = -2,         /* CUDA runtime error, see avo_last_error */
  AVO_E_CAPACITY = -3,     /* an output buffer is too small; required sizes are written back */
  AVO_E_EMPTY_SITES = -4,  /* call() with no variants: the reference throws (TreeRegionJoin.scala:77) */
  AVO_E_UNSORTED = -5,     /* reads not sorted by start / sites not sorted by (contig,start,end) */
  AVO_E_UNSUPPORTED = -6,  /* parameter outside the supported domain (e.g. ploidy > AVO_MAX_PLOIDY) */
  AVO_E_NOMEM = -7,
  AVO_E_RETRY = -8         /* avo_shard_finish: a rank ran out of an internal capacity (now grown): run the step again */
} avo_status;

enum { AVO_MEM_HOST = 0, AVO_MEM_DEVICE = 1 };
enum { AVO_MAX_PLOIDY = 7 };
enum { AVO_BLOCK_BASES = 10, AVO_BLOCK_BYTES = 16 };

/* alignment operators after merging CIGAR with MD (ObservationOperator.scala:42-171):
 * op word = (length << 4) | code */
enum {
  AVO_OP_MATCH = 0,     /* Match(n)           sequence match   '=' */
  AVO_OP_MISMATCH = 1,  /* Match(n, Some(ref)) sequence mismatch 'X'; n (ref, read) bytes in refb */
  AVO_OP_INS = 2,       /* Insertion(n) */
  AVO_OP_DEL = 3,       /* Deletion(ref); n ref nibbles ((n+1)/2 bytes) in refb */
  AVO_OP_SOFTCLIP = 4,  /* Clipped(n, soft = true) */
  AVO_OP_HARDCLIP = 5   /* Clipped(n, soft = false) */
};

/* flags[] bits */
enum {
  AVO_READ_NEGATIVE_STRAND = 1, /* AlignmentRecord.readNegativeStrand */
  AVO_READ_SKIP_CALL = 2,       /* the reference's observeRead would throw for this read (null mapq,
                                   sequence/qual shorter than the CIGAR): skipped by call(), still
                                   walked by discover() */
  AVO_READ_OPS_DROPPED = 4      /* set by the packers on a read with more than 65535 operators (the
                                   limit of BAM's own CIGAR field and of avo_read_meta.n_ops): the
                                   record is kept without operators, so the read takes no part in
                                   discovery or calling; callers that need to know count this bit */
};

/* ---- lifecycle ---------------------------------------------------------------------------- */
int avo_ctx_create(int device, avo_ctx** out);
void avo_ctx_destroy(avo_ctx* ctx);
const char* avo_last_error(const avo_ctx* ctx); /* owned by ctx; valid until the next call */
const char* avo_version(void);
/* Run on this CUDA stream (a cudaStream_t) instead of the context's own; NULL restores it. */
int avo_ctx_set_stream(avo_ctx* ctx, void* cuda_stream);

/* ---- inputs ------------------------------------------------------------------------------- */
/* Per-read scalars, one 32-byte record per read (one DRAM sector: a read that joins a site costs
 * one sector for all of them).  Replaces the AlignmentRecord fields listed in SURVEY.md 8b. */
typedef struct {
  int32_t start;     /* AlignmentRecord.start */
  int32_t end;       /* AlignmentRecord.end */
  uint32_t seq_off;  /* index of the read's first 16-byte block in payload */
  uint32_t op_off;   /* index of the read's first operator in ops */
  uint32_t refb_off; /* byte index of the read's first entry in refb */
  uint16_t n_ops;    /* number of operators (0: the reference could not extract any) */
  uint8_t mapq;
  uint8_t flags;     /* AVO_READ_* */
  uint32_t qhead;    /* phred of the read's first four bases, byte i = qual(i): DiscoverVariants
                        tests qual(i), i = offset inside the mismatch run (DiscoverVariants.scala:199) */
  uint32_t l_seq;    /* number of bases */
} avo_read_meta;

/* Bounds a packer knows for free.  Optional (valid = 0: the library computes them with one extra
 * pass over the records).  When given they are only bounds, and avo_discover /
 * avo_discover_and_call check them, together with the sort order, against every record on the
 * device while discovering (AVO_E_INVALID / AVO_E_UNSORTED if wrong); avo_call always recomputes. */
typedef struct {
  int32_t min_start; /* <= every start */
  int32_t max_end;   /* >= every end */
  int32_t max_span;  /* >= every end - start */
  int32_t valid;
} avo_batch_stats;

/* Compact wire form of the record (optional, host batches): the three offsets of avo_read_meta are
 * prefix sums of l_seq (in payload blocks), n_ops and n_refb, so they need not cross PCIe; the
 * library rebuilds the 32-byte records on the device (one scan + one expansion kernel).  A batch
 * can use it when every read has end - start and l_seq below 65536 and fewer than 256 operators
 * and refb bytes, and every operator is shorter than 4096 (avo_pack_wire checks). */
typedef struct {
  int32_t start;
  uint16_t span;    /* end - start */
  uint16_t l_seq;
  uint8_t n_ops;
  uint8_t n_refb;   /* refb bytes of the read */
  uint8_t mapq;
  uint8_t flags;
  uint32_t qhead;
} avo_read_wire;    /* 16 bytes */

/* Tighter wire form, 6 bytes per read: the start travels as the difference to the previous read's
 * start (the batch is sorted), end, l_seq and the refb byte count follow from the read's operators
 * (end - start = lengths of = X D, l_seq = lengths of = X I S, refb bytes = X lengths +
 * (D length + 1) / 2) and are recomputed on the device, and of the four head qualities (avo_read_meta.qhead)
 * only the first two travel: discovery fetches the others from the payload in the rare mismatch run of
 * three or more bases.  The few reads whose sizes do not follow from their operators (no operators,
 * strings longer than the CIGAR's read length) are listed as exceptions.  A batch can use it when it
 * fits avo_read_wire and every start difference is below 16384 (avo_pack_wire6 checks). */
typedef struct {
  uint16_t dstart_flags; /* ((start - previous read's start) << 2) | flags; first read: start - wire6_base */
  uint8_t n_ops;
  uint8_t mapq;
  uint8_t q0, q1;        /* qual(0), qual(1) of the read */
} avo_read_wire6;   /* 6 bytes */
/* With it the operators travel as 12 bits each: one length byte (every operator shorter than 256) and
 * one code nibble (two per byte, the even operator in the low nibble). */

typedef struct {
  uint32_t index;  /* read index in the batch, ascending */
  uint32_t l_seq;
  int32_t end;
  uint32_t reserved;
} avo_wire_exception;

/* Discovery's per-read candidate events, a by-product of packing (DiscoverVariants.variantsInRead,
 * DiscoverVariants.scala:112-252, is a walk over the operators extractAlignmentOperators produced): every
 * mismatching base as (reference position, (reference base << 4 | read base) << 8 | the quality the
 * reference tests, qual(i) with i counted inside the mismatch run) and every I / D operator after the read's
 * first aligned base as (read, operator index), in read order.  avo_derive_events fills them from a packed
 * batch (device batches: three kernels; host batches: a loop); a DEVICE batch that carries them
 * (avo_read_batch.events) is discovered by counting events -- a coalesced stream of 6 bytes per mismatching
 * base -- instead of walking 25 M reads (k_discover_events instead of k_discover_tiles).  `snv_first[w]` is
 * the first SNV event of read 32 w (the index the tiles use).  The deriving pass also measures the exact
 * bounds and checks the sort order (AVO_E_UNSORTED), so a batch with events needs no stats hints. */
typedef struct {
  int64_t n_snv;          /* in (avo_derive_events): capacity of snv_pos / snv_pq (n_refb always suffices); out: events */
  int64_t n_indel;        /* in: capacity of indel_read / indel_op (n_ops always suffices); out: events */
  int32_t* snv_pos;
  uint16_t* snv_pq;
  uint32_t* snv_first;    /* [(n_reads + 31) / 32 + 1] */
  int32_t* indel_read;
  uint32_t* indel_op;
  avo_batch_stats stats;  /* out: exact bounds of the batch (valid = 1) */
} avo_read_events;

/* One batch = reads of ONE contig, sorted by start (ties in any order; the order of the batch is
 * the canonical summation order). */
typedef struct {
  int64_t n_reads;
  int64_t n_blocks; /* payload blocks (a read of l_seq bases owns ceil(l_seq / 10) consecutive blocks) */
  int64_t n_ops;
  int64_t n_refb;  /* refb bytes (< 2^31) */
  int32_t contig_id; /* rank of this contig in the site table's contig order */
  int32_t mem;
  avo_batch_stats stats;
  const avo_read_meta* meta; /* [n_reads] */
  const uint8_t* payload;    /* [n_blocks * 16] bases + qualities, 16-byte aligned, see "Conventions" */
  const uint32_t* ops;       /* [n_ops] */
  const uint8_t* refb;       /* [n_refb] in operator order: a MISMATCH op of length n holds n bytes
                                (reference base << 4) | read base, a DEL op its n deleted reference
                                bases as nibbles (high first) padded to (n+1)/2 bytes */
  /* optional compact wire columns (AVO_MEM_HOST only): when both are non-NULL they are what is
   * copied to the device and `meta` / `ops` are ignored (may be NULL) */
  const avo_read_wire* wire_meta; /* [n_reads] */
  const uint16_t* wire_ops;       /* [n_ops] (length << 4) | code */
  /* or the 6-byte form (with wire6_oplen / wire6_opcode): takes precedence over wire_meta when non-NULL */
  const avo_read_wire6* wire6_meta;     /* [n_reads] */
  const uint8_t* wire6_oplen;           /* [n_ops] operator lengths */
  const uint8_t* wire6_opcode;          /* [(n_ops + 1) / 2] operator codes, 4 bits each */
  const avo_wire_exception* wire6_exc;  /* [n_wire6_exc] */
  int64_t n_wire6_exc;
  int32_t wire6_base;                   /* start of the first read minus its stored difference */
  int32_t reserved;
  const avo_read_events* events;        /* optional (AVO_MEM_DEVICE batches; ignored otherwise): avo_derive_events */
} avo_read_batch;

/* Fills `events` (arrays in the batch's memory: device arrays for a device batch, host arrays for a host
 * batch in the full form -- `ctx` may then be NULL) from the packed columns.  Waits for the device (the
 * counts come back). */
int avo_derive_events(avo_ctx* ctx, const avo_read_batch* reads, avo_read_events* events);

/* Variant sites, sorted by (contig, start, start + ref_len); unique.  Replaces the Variant
 * fields read by DiscoveredVariant.apply (DiscoveredVariant.scala:32-37); end == start + ref_len. */
typedef struct {
  int64_t n;
  int64_t n_allele_nibbles; /* allele_off[n] */
  int32_t mem;
  int32_t reserved;
  const int32_t* contig;      /* [n] contig rank, or NULL = all 0 */
  const int32_t* start;       /* [n] */
  const int32_t* ref_len;     /* [n] >= 1 */
  const int32_t* alt_len;     /* [n] >= 0 */
  const uint32_t* allele_off; /* [n+1] nibble index (always even); ref nibbles then alt nibbles */
  const uint8_t* alleles4;
} avo_sites;

/* CopyNumberMap (CopyNumberMap.scala:75-111): base ploidy + per-contig sorted CNV list.  Host memory. */
typedef struct {
  int64_t n;
  const int32_t* contig; /* [n] contig rank; entries sorted by (contig, start, end) */
  const int32_t* start;
  const int32_t* end;
  const int32_t* copy_number;
} avo_cnv;

typedef struct {
  int32_t ploidy;             /* CopyNumberMap.basePloidy */
  int32_t score_all_sites;    /* scoreAllSites (gVCF mode) */
  int32_t max_quality;        /* call(maxQuality = 93) */
  int32_t max_mapq;           /* call(maxMapQ = 93) */
  int32_t min_phred_discover; /* optPhredThreshold.getOrElse(0) */
  int32_t min_obs_discover;   /* optMinObservations; < 0 = None (distinct); keep count > this */
  int32_t host_payload;       /* AVO_MEM_HOST batches: 0 = read the payload in place over PCIe when
                                 it is page-locked and the mode touches it sparsely (variant-only
                                 calling, discovery), 1 = always copy it to the device, 2 = as 0, but the
                                 call stage's blocks are gathered by host threads: the device lists the
                                 16-byte blocks the joined reads need, the library copies them out of the
                                 payload into a dense page-locked array and sends that (one PCIe request
                                 per block becomes one DMA) */
  int32_t gather_threads;     /* host threads for host_payload == 2 (0 = 4) */
} avo_params;

/* ---- outputs ------------------------------------------------------------------------------ */
/* Discovered variants, sorted (start, ref_len, alt_len, alleles).  Two-phase: on AVO_E_CAPACITY
 * n and n_allele_nibbles hold the required capacities. */
typedef struct {
  int64_t capacity;        /* in: entries available */
  int64_t allele_capacity; /* in: nibbles available in alleles4 */
  int64_t n;               /* out */
  int64_t n_allele_nibbles;/* out */
  int32_t mem;
  int32_t reserved;
  int32_t* start;
  int32_t* ref_len;
  int32_t* alt_len;
  uint32_t* allele_off; /* [capacity+1] */
  uint8_t* alleles4;
  int32_t* count;       /* supporting observations per variant (may be NULL) */
} avo_sites_out;

/* Per-site aggregate (Observation.scala:72-83, summed as BiallelicGenotyper.scala:475-500) and the
 * genotype fields of BiallelicGenotyper.scala:677-748, one row per input site, same order.
 * n_states = maxPloidy + 1 is the row stride of the likelihood arrays. */
typedef struct {
  int64_t n_sites;  /* in: == sites->n */
  int32_t n_states; /* in */
  int32_t mem;
  uint8_t* emitted;      /* 1 iff >= 1 observation survived the score join (the reference emits
                            only such sites); all other fields are zero when 0 */
  int32_t* allele_fwd;   /* alleleForwardStrand */
  int32_t* other_fwd;    /* otherForwardStrand */
  int32_t* allele_cov;   /* alleleCoverage  -> alternateReadDepth */
  int32_t* other_cov;    /* otherCoverage   -> referenceReadDepth */
  int32_t* total_cov;    /* totalCoverage   -> readDepth */
  double* square_mapq;
  double* ref_ll;        /* [n * n_states] referenceLogLikelihoods */
  double* allele_ll;     /* alleleLogLikelihoods */
  double* other_ll;      /* otherLogLikelihoods */
  double* nonref_ll;     /* nonRefLogLikelihoods */
  uint8_t* first_is_ref; /* first("isRef") */
  int32_t* copy_number;  /* first("copyNumber") */
  int32_t* n_alt;        /* alleles = ALT x n_alt ++ REF x n_ref ++ OTHER_ALT x n_other */
  int32_t* n_ref;
  int32_t* n_other;
  double* qual;          /* 10/ln10 * (max - second) */
  int32_t* gq;           /* qual.toInt */
  int32_t* sb;           /* [n * 4] strandBiasComponents */
  float* fisher;         /* fisherStrandBiasPValue */
  float* rms_mapq;
  double* gl;            /* [n * n_states] genotypeLikelihoods (entries 0..copy_number) */
  double* nonref_gl;     /* nonReferenceLikelihoods */
} avo_site_results;

/* Device-side time of the last call, measured with CUDA events on the context's stream. */
typedef struct {
  float ms_total;
  float ms_h2d;       /* host -> device copies of inputs */
  float ms_discover;  /* all discovery kernels */
  float ms_observe;   /* observe/classify/reduce/genotype kernel (+ its set-up kernels) */
  float ms_d2h;
  float ms_discover_tiles; /* k_discover_tiles alone */
  float ms_call_tiles;     /* the call stage's kernels: buckets, join, observe, reduce */
  int32_t n_launches;      /* kernels launched by the call */
  int32_t n_tile_launches; /* of which tile kernels (1 per stage unless a buffer overflowed) */
  int64_t h2d_bytes;
  int64_t d2h_bytes;
  int32_t payload_in_place; /* 1: the payload was read from pinned host memory, not copied; 2: and the call
                               stage's blocks were gathered on the host */
  float ms_call_pairs;      /* of ms_call_tiles: k_call_pairs alone (join + observation; 0 on the host-gathered route) */
  int64_t gathered_blocks;  /* host_payload == 2: 16-byte blocks gathered on the host */
} avo_timing;
int avo_get_timing(const avo_ctx* ctx, avo_timing* out);

/* ---- the hot path ------------------------------------------------------------------------- */
/* DiscoverVariants.apply on one batch. */
int avo_discover(avo_ctx* ctx, const avo_read_batch* reads, const avo_params* params,
                 avo_sites_out* out);

/* BiallelicGenotyper.call on one batch against a site table (variant-only mode; use
 * avo_call_all_sites for scoreAllSites). */
int avo_call(avo_ctx* ctx, const avo_read_batch* reads, const avo_sites* sites, const avo_cnv* cnv,
             const avo_params* params, avo_site_results* out);

/* BiallelicGenotyper.discoverAndCall: discovery, then call against the discovered sites, which
 * stay on the device in between.  `results` rows correspond to `sites_out` rows; results->n_sites
 * is an in/out capacity like sites_out->capacity. */
int avo_discover_and_call(avo_ctx* ctx, const avo_read_batch* reads, const avo_cnv* cnv,
                          const avo_params* params, avo_sites_out* sites_out,
                          avo_site_results* results);

/* ---- union of several samples' batches (CLI/TrioGenotyper.scala:216: `reads.union(...)` before discovery) ----
 * Merges the RECORDS of n_parts sorted device batches of one contig into `out_meta` (device, sum of the
 * parts' n_reads records), stably sorted by start, with seq_off / op_off / refb_off rebased by the given
 * bases: where each part's payload blocks / operators / mismatch bytes begin in the concatenated columns
 * the union batch will point at (the caller lays the parts' columns out back to back, or copies them).
 * One launch; does not wait for the device. */
int avo_merge_meta(avo_ctx* ctx, int32_t n_parts, const avo_read_meta* const* part_meta, const int64_t* part_reads,
                   const uint32_t* block_base, const uint32_t* op_base, const uint32_t* refb_base,
                   avo_read_meta* out_meta);

/* ---- region-sharded discoverAndCall: one process per GPU -------------------------------
 * Replaces, for several GPUs, what the reference does with `sortByKey().collect` + `sc.broadcast` of the
 * variants (util/TreeRegionJoin.scala:43-50, 184-185) and with the shuffle behind the per-site sums
 * (genotyping/BiallelicGenotyper.scala:475-500).  Rank g owns the sites that START in [own_lo, own_hi)
 * and is handed every read that can overlap one of them or put a discovery candidate on one: the reads
 * with own_lo - max_span <= start < own_hi + (longest site).  Nothing is reduced across ranks; the host
 * framework runs two all-gathers of fixed-size DEVICE blocks between three calls that never wait for the
 * device (only avo_shard_finish synchronises):
 *
 *   avo_shard_discover(...)            -> table_block            this rank's owned slice of its table
 *   all-gather(table_block)            -> gathered_tables        (world blocks, rank order)
 *   avo_shard_call(...)                -> rows_block             the global table is assembled on the device,
 *                                                                the reads are called against it, owned rows packed
 *   all-gather(rows_block)             -> gathered_rows
 *   avo_shard_finish(...)              -> global table + every row on every rank; one synchronisation
 *
 * A framework that wants the rows on ONE rank only gathers the rows blocks to that rank and all-gathers just
 * their 32-byte headers (every rank needs the flags in them): the other ranks pass `gathered_rows` with
 * their own block in place and, for the other ranks' blocks, the header with its first word (the row count)
 * set to 0 -- avo_shard_finish then leaves those rows untouched.  Fastest is no rows gather at all:
 * avo_shard_call_direct (below) stores the rows into the collecting rank's arrays over NVLink.
 * Block sizes come from the plan's capacities (avo_shard_*_block_bytes).  avo_shard_finish returns
 * AVO_E_CAPACITY with the capacities that are needed (identical on every rank: they follow from the
 * gathered headers), or AVO_E_RETRY when some rank ran out of an internal capacity (every rank sees it in
 * the gathered headers): in both cases EVERY rank runs the three calls again.  Tables and rows are device
 * memory.  All three calls use the context's stream; the collectives must be ordered on it. */
typedef struct {
  int32_t rank, world;
  int32_t own_lo, own_hi;   /* sites with own_lo <= start < own_hi belong to this rank */
  int64_t site_cap;         /* rows one rank's table block holds */
  int64_t allele_cap;       /* allele bytes one rank's table block holds (a multiple of 16) */
  int64_t row_cap;          /* rows one rank's result block holds */
  int32_t n_states;         /* max ploidy + 1, as in avo_site_results */
  int32_t reserved;
} avo_shard_plan;
int64_t avo_shard_table_block_bytes(const avo_shard_plan* plan);
int64_t avo_shard_rows_block_bytes(const avo_shard_plan* plan);
int avo_shard_discover(avo_ctx* ctx, const avo_read_batch* reads, const avo_params* params,
                       const avo_shard_plan* plan, void* table_block);
int avo_shard_call(avo_ctx* ctx, const avo_cnv* cnv, const avo_params* params, const avo_shard_plan* plan,
                   const void* gathered_tables, avo_sites_out* global_sites, void* rows_block);
int avo_shard_finish(avo_ctx* ctx, const avo_shard_plan* plan, const void* gathered_tables,
                     const void* gathered_rows, avo_sites_out* global_sites, avo_site_results* global_results,
                     avo_shard_plan* needed);

/* The same step with the rows written where they are wanted, over NVLink, by the kernel that computes
 * them: `root_rows` are the global result columns of ONE rank (the root), as device pointers valid on
 * THIS rank's GPU -- the root's own arrays on the root, a peer mapping of them elsewhere (avo_peer_open).
 * The step's last kernel stores this rank's OWNED rows straight into them at their global row numbers
 * (known on the device from the gathered table headers), one contiguous range of 16-byte stores per
 * column; no rows travel through rows_block,
 * whose 32-byte header (flags) is all the framework still gathers -- to every rank -- before
 * avo_shard_finish.  That collective, ordered after this call on every rank's stream, is also what tells
 * the root that the peers' stores have landed.  The root must not hand the rows of a step to anyone who
 * still reads them when it enters the next step's first all-gather. */
int avo_shard_call_direct(avo_ctx* ctx, const avo_cnv* cnv, const avo_params* params, const avo_shard_plan* plan,
                          const void* gathered_tables, avo_sites_out* global_sites, void* rows_block,
                          const avo_site_results* root_rows);

/* Device memory another process of the node can map (CUDA IPC; what the root's result columns live in
 * for avo_shard_call_direct).  avo_peer_alloc returns the block and a 64-byte handle to send to the other
 * processes (any channel); avo_peer_open maps the block of a handle on this context's GPU (peer access
 * over NVLink is enabled on first use); avo_peer_close / avo_peer_free undo them -- every peer closes
 * before the owner frees. */
enum { AVO_PEER_HANDLE_BYTES = 64 };
int avo_peer_alloc(avo_ctx* ctx, int64_t bytes, void** ptr, unsigned char* handle);
int avo_peer_open(avo_ctx* ctx, const unsigned char* handle, void** ptr);
int avo_peer_close(avo_ctx* ctx, void* ptr);
int avo_peer_free(avo_ctx* ctx, void* ptr);

/* ---- scoreAllSites (gVCF mode) ---------------------------------------------------------------- */
/* Reference-model rows of BiallelicGenotyper.call(scoreAllSites = true) (BG:223-275, 299-306): one
 * row per reference position that an observation of a joined read reaches without overlapping
 * one of that read's variants: DiscoveredVariant(rr) = (contig, start, "N", no alternate),
 * end = start + 1 (DiscoveredVariant.scala:59-61).  Rows are in ascending position order; like
 * variant rows, a row with emitted == 0 is not a site (every observation there fell out of the
 * score join).  Two-phase: on AVO_E_CAPACITY `n` holds the capacity that is needed. */
typedef struct {
  int64_t capacity;      /* in: rows available in start[] and rows.* (rows.n_sites is ignored) */
  int64_t n;             /* out */
  int32_t* start;        /* [capacity] */
  avo_site_results rows; /* rows.mem: where start[] and the columns live; rows.n_states as for variants */
} avo_ref_model_out;

/* BiallelicGenotyper.call(..., scoreAllSites = true): `out` receives the variant rows exactly as
 * avo_call produces them, `ref_model` the reference-model rows.  params->score_all_sites is implied. */
int avo_call_all_sites(avo_ctx* ctx, const avo_read_batch* reads, const avo_sites* sites,
                       const avo_cnv* cnv, const avo_params* params, avo_site_results* out,
                       avo_ref_model_out* ref_model);

/* ---- RewriteHets + HardFilterGenotypes (the CLI's post passes, CLI/BiallelicGenotyper.scala:279-282) -- */
/* Thresholds of RewriteHetsArgs (RewriteHets.scala:24-49) and HardFilterGenotypesArgs
 * (HardFilterGenotypes.scala:34-166); a hard filter is applied iff its threshold is > 0. */
typedef struct {
  int32_t min_quality;
  int32_t filter_ref_genotypes;  /* HardFilterGenotypes.apply(filterRefGenotypes): the CLI passes !scoreAllSites */
  int32_t emit_all_genotypes;
  int32_t disable_het_snp_rewriting;
  int32_t disable_het_indel_rewriting;
  int32_t min_snp_depth, max_snp_depth, min_indel_depth, max_indel_depth;
  float min_het_snp_quality_by_depth, min_hom_snp_quality_by_depth;
  float min_het_indel_quality_by_depth, min_hom_indel_quality_by_depth;
  float max_snp_phred_strand_bias, max_indel_phred_strand_bias;
  float min_snp_rms_mapping_quality, min_indel_rms_mapping_quality;
  float min_het_snp_alt_allelic_fraction, max_het_snp_alt_allelic_fraction, min_hom_snp_alt_allelic_fraction;
  float min_het_indel_alt_allelic_fraction, max_het_indel_alt_allelic_fraction, min_hom_indel_alt_allelic_fraction;
} avo_filter_params;

/* bits of filters_failed, in the order of the reference's filter lists (:258-344):
 * bit k (SNP) / bit 9 + k (INDEL), k = 0 HET*QD, 1 HOM*QD, 2 *MQ, 3 *FS, 4 *MINDP, 5 *MAXDP,
 * 6 HET*MINAF, 7 HET*MAXAF, 8 HOM*MINAF */
typedef struct {
  int64_t n;
  int32_t mem;
  int32_t reserved;
  uint8_t* emit;            /* [n] the genotype survives filterGenotype (:632-660) */
  uint8_t* rewritten;       /* [n] RewriteHets rewrote it: alleles all ALT, genotypeQuality unset */
  uint8_t* filters_applied; /* [n] VariantCallingAnnotations.filtersApplied */
  uint32_t* filters_failed; /* [n] bit set of failed filters; filtersPassed = applied && failed == 0 */
  int32_t* n_alt;           /* [n] alleles after rewriting: ALT x n_alt ++ REF x n_ref ++ OTHER_ALT x (unchanged) */
  int32_t* n_ref;
  uint8_t* gq_valid;        /* [n] 0: genotypeQuality was unset by the rewrite */
} avo_filter_out;

/* rows = the columns avo_call / avo_call_all_sites produced (rows->mem says where they and the
 * other arrays live).  ref_len / alt_len: allele lengths per row; alt_len NULL or < 0 = no alternate
 * allele (reference-model rows). */
int avo_filter_genotypes(avo_ctx* ctx, const avo_site_results* rows, const int32_t* ref_len,
                         const int32_t* alt_len, const avo_filter_params* params, avo_filter_out* out);

/* ---- SquareOffReferenceModel.squareOffSite (SquareOffReferenceModel.scala:176-244) ---------------- */
/* One sample's gVCF rows (the outputs of avo_call_all_sites for that sample) joined to the cohort's
 * site list: one row of the (samples x sites) matrix avo_joint consumes.  All tables are those of
 * ONE contig.  kind[j]: how the sample is represented at site j. */
enum {
  AVO_SQ_ABSENT = 0,            /* no genotype of the sample overlaps the site */
  AVO_SQ_SCORED = 1,            /* hasScoredVariant: the sample's own row for exactly this variant */
  AVO_SQ_EXCISED_VARIANT = 2,   /* excised from an overlapping variant row: gl := its nonReferenceLikelihoods */
  AVO_SQ_EXCISED_REF_MODEL = 3  /* excised from the reference-model row at the site */
};
typedef struct {
  const avo_sites* variants;          /* the sample's variant table (as given to avo_call_all_sites) */
  const avo_site_results* variant_rows;
  int64_t n_ref_model;                /* the sample's reference-model rows */
  const int32_t* ref_model_start;     /* [n_ref_model] ascending */
  const avo_site_results* ref_model_rows;
  /* Reference-model BLOCKS (banded gVCFs, e.g. rows that did not come from avo_call_all_sites):
   * row k covers [start, start + ref_model_len[k]).  NULL = every row covers one position.
   * ref_model_max_len bounds the lengths (0 with device tables = "unknown": 65536 is assumed). */
  const int32_t* ref_model_len;       /* [n_ref_model] >= 1, or NULL */
  int32_t ref_model_max_len;
  int32_t reserved;
} avo_sample_rows;
typedef struct {
  int64_t n_sites;   /* == cohort->n */
  int32_t n_states;
  int32_t mem;
  uint8_t* kind;     /* [N] AVO_SQ_* */
  int32_t* src;      /* [N] row index in the table `kind` names, or -1 */
  int32_t* n_alt;    /* [N] alleles of the representing genotype */
  int32_t* n_ref;
  int32_t* n_other;
  double* gl;        /* [N * n_states] genotypeLikelihoods of the squared-off genotype */
} avo_square_out;
/* Every table lives where out->mem says (all host or all device). */
int avo_square_off(avo_ctx* ctx, const avo_sites* cohort, const avo_sample_rows* sample, avo_square_out* out);

/* ---- TrioCaller.processVariant (TrioCaller.scala:86-221) ------------------------------------------ */
enum {
  AVO_TRIO_UNCHANGED = 0,      /* child genotype kept as it is */
  AVO_TRIO_PHASED = 1,         /* confirmed: phased = true, alleles as called */
  AVO_TRIO_NO_CALL = 2,        /* inconsistent with the parents: NO_CALL x 2, genotypeQuality unset */
  AVO_TRIO_PHASED_ALT_REF = 3, /* het phased from the first parent: alleles [ALT, REF], phased */
  AVO_TRIO_PHASED_REF_ALT = 4  /* het phased from the second parent: alleles [REF, ALT], phased */
};
/* The three samples (first parent, second parent, child) at N sites: the rows s = 0, 1, 2 of the
 * (samples x sites) matrix, sample-major like avo_joint_in.  Outputs per site: kept (the site has an
 * ALT allele before and after processing, :92-96), child_action (AVO_TRIO_*), filled (bit s set:
 * sample s had no genotype and is emitted as NO_CALL x 2, :209-216). */
typedef struct {
  int64_t n_sites;
  int32_t mem;
  int32_t reserved;
  const uint8_t* present;  /* [3*N] NULL = all present */
  const int32_t* n_alt;    /* [3*N] */
  const int32_t* n_ref;
  const int32_t* n_other;
  const int32_t* n_nocall; /* NULL = none */
  uint8_t* kept;           /* [N] out */
  uint8_t* child_action;   /* [N] out */
  uint8_t* filled;         /* [N] out */
} avo_trio;
int avo_trio_call(avo_ctx* ctx, avo_trio* trio);

/* ---- joint calling: JointAnnotatorCaller.apply (JointAnnotatorCaller.scala:52-64, 74-281) ------ */
/* Genotypes of S samples at the same N (squared-off) sites, sample-major: entry [s * n_sites + i].
 * The per-sample columns are the ones avo_call produces (n_alt, n_ref, n_other, gl). */
typedef struct {
  int64_t n_sites;
  int32_t n_samples;
  int32_t n_states;        /* row stride of gl */
  int32_t mem;
  int32_t reserved;
  const uint8_t* present;  /* [S*N] 1 iff the sample has a genotype at the site; NULL = all */
  const int32_t* n_alt;    /* [S*N] alleles: ALT x n_alt, REF x n_ref, OTHER_ALT x n_other, NO_CALL x n_nocall */
  const int32_t* n_ref;
  const int32_t* n_other;
  const int32_t* n_nocall; /* NULL = none */
  const double* gl;        /* [S*N*n_states] genotypeLikelihoods, entries 0..copies */
  /* site totals over ALL samples of the cohort, when the samples of this call are only a shard
   * (the caller all-reduced the outputs of avo_joint_counts); NULL = count these samples */
  const int32_t* alt_alleles;    /* [N] */
  const int32_t* called_alleles; /* [N] */
} avo_joint_in;

enum {
  AVO_JOINT_POSTERIORS = 1, /* genotypePosteriors set */
  AVO_JOINT_PRIORS = 2,     /* genotypePriors set */
  AVO_JOINT_RECALLED = 4    /* alleles and genotypeQuality rewritten (n_alt, n_ref, gq) */
};

typedef struct {
  int64_t n_sites;
  int32_t n_samples;
  int32_t n_states;
  int32_t mem;
  int32_t reserved;
  uint8_t* site_emitted; /* [N] minorAlleleFrequency > 0 (sites with 0 are discarded, :81-83) */
  double* maf;           /* [N] */
  double* quality;       /* [N] -10/ln10 * post0_sum (final only when all samples are in this call) */
  double* post0_sum;     /* [N] sum over this call's samples of posterior(0), in sample order */
  uint8_t* flags;        /* [S*N] AVO_JOINT_* */
  float* posteriors;     /* [S*N*n_states] */
  float* priors;         /* [S*N*n_states] */
  int32_t* n_alt;        /* [S*N] re-called alleles: ALT x n_alt ++ REF x n_ref (unchanged when not re-called) */
  int32_t* n_ref;
  int32_t* gq;           /* [S*N] valid when AVO_JOINT_RECALLED */
} avo_joint_out;

/* Local allele counts per site (calculateMinorAlleleFrequency's two sums, :117-128). */
int avo_joint_counts(avo_ctx* ctx, const avo_joint_in* in, int32_t* alt_alleles, int32_t* called_alleles);
/* annotateSite for every site. */
int avo_joint(avo_ctx* ctx, const avo_joint_in* in, avo_joint_out* out);

/* The depth roll-up annotateSite attaches to the variant (JointAnnotatorCaller.calculateAnnotations,
 * JointAnnotatorCaller.scala:85-90, 141-147; VariantSummary.scala:39-146): per site, the sums over the
 * site's genotypes of readDepth, referenceReadDepth and of the strand-bias components
 * (forward = sb[0] + sb[2], reverse = sb[1] + sb[3], reference forward = sb[0], reference reverse = sb[1]),
 * each sum over the genotypes that have the field (Option merge: None only if no genotype has it).
 * Same (samples x sites) sample-major layout as avo_joint_in. */
typedef struct {
  int64_t n_sites;
  int32_t n_samples;
  int32_t mem;
  const uint8_t* present;        /* [S*N] 1 iff the sample has a genotype at the site; NULL = all */
  const int32_t* read_depth;     /* [S*N] < 0 = unset; NULL = unset everywhere */
  const int32_t* ref_read_depth; /* [S*N] < 0 = unset; NULL = unset everywhere */
  const int32_t* sb;             /* [S*N*4] strandBiasComponents, sb[4k] < 0 = the genotype has none; NULL = none */
} avo_depths_in;
typedef struct {
  int64_t n_sites;
  int32_t mem;
  int32_t reserved;
  uint8_t* annotated;            /* [N] 1 iff the site has more than one genotype (:85-90: else no annotation) */
  int32_t* read_depth;           /* [N] the six VariantAnnotation fields, -1 = unset */
  int32_t* ref_read_depth;
  int32_t* fwd_read_depth;
  int32_t* rev_read_depth;
  int32_t* ref_fwd_read_depth;
  int32_t* ref_rev_read_depth;
} avo_site_depths;
int avo_joint_depths(avo_ctx* ctx, const avo_depths_in* in, avo_site_depths* out);

/* ---- host-side packer (PrefilterReads + ObservationOperator.extractAlignmentOperators) ------ */
/* PrefilterReads' read predicates (CORE/util/PrefilterReads.scala:142-209), evaluated by the packers --
 * host (avo_pack_reads*) and device (avo_pack_reads_device, inside k_pack_count / k_pack_records) -- from
 * the columns they load anyway: mapped AND (primary OR keep_non_primary) AND (mapq absent OR mapq >
 * min_mapq) AND (NOT duplicate OR keep_duplicates) AND contig_kept.  The contig predicates (:118-134,
 * 216-274) are tests of the contig NAME: the host evaluates them once per batch (one batch = one contig)
 * and passes the answer. */
typedef struct {
  int32_t min_mapq;          /* PrefilterReadsArgs.minMappingQuality (CLI default 10; strict >) */
  int32_t keep_duplicates;   /* keepDuplicates */
  int32_t keep_non_primary;  /* keepNonPrimary */
  int32_t contig_kept;       /* contigFilterFn(name of the batch's contig) */
} avo_prefilter;

/* Text columns of AlignmentRecords (blobs + [n+1] offsets; bits of rec_flags: bit0 readMapped, bit1
 * readNegativeStrand, bit2 cigar present, bit3 MD present, bit4 mapq present, bit5 primaryAlignment,
 * bit6 duplicateRead).  Packs, in input order, the mapped reads that pass `prefilter` (NULL = no
 * predicate) and for which keep[i] != 0 (NULL = all).  Output arrays must be sized for the upper bounds
 * returned by avo_pack_bounds. */
typedef struct {
  int64_t n;
  const int64_t* start;
  const int64_t* end;
  const int32_t* mapq;
  const uint8_t* rec_flags;
  const char* cigar; const int64_t* cigar_off;
  const char* md;    const int64_t* md_off;
  const char* seq;   const int64_t* seq_off;
  const char* qual;  const int64_t* qual_off;
  const uint8_t* keep;
  const avo_prefilter* prefilter;   /* host pointer (also for avo_pack_reads_device); may be NULL */
  /* narrow != 0: the numeric columns are 32-bit or less -- `start` and `end` point at int32_t, the four
   * `*_off` arrays at uint32_t, `mapq` at uint8_t (absence through rec_flags bit4 only) -- which takes 23 of
   * the ~355 bytes per read off the PCIe link when the text is packed on the device.  For batches whose
   * blobs stay below 4 GB (a region bin). */
  int32_t narrow;
  int32_t reserved;
} avo_text_reads;

typedef struct {
  int64_t n_reads, n_blocks, n_ops, n_refb; /* in: capacities; out: used */
  avo_read_meta* meta;
  uint8_t* payload; uint32_t* ops; uint8_t* refb;
  int64_t* source_index; /* [n_reads] index of each packed read in the input (may be NULL) */
} avo_packed_out;

/* Full records / operators (host) -> compact wire columns (host).  AVO_E_UNSUPPORTED when a read or
 * an operator does not fit the compact form (the batch then travels in the full form). */
int avo_pack_wire(const avo_read_batch* full, avo_read_wire* wire_meta, uint16_t* wire_ops);
/* ... -> the 6-byte form.  exc holds exc_capacity entries; *n_exc and *base are outputs (they go into
 * avo_read_batch.n_wire6_exc / wire6_base).  AVO_E_UNSUPPORTED as above, or with too many exceptions. */
int avo_pack_wire6(const avo_read_batch* full, avo_read_wire6* wire6_meta, uint8_t* wire6_oplen,
                   uint8_t* wire6_opcode, avo_wire_exception* exc, int64_t exc_capacity, int64_t* n_exc,
                   int32_t* base);

int avo_pack_bounds(const avo_text_reads* in, int64_t* n_reads, int64_t* n_blocks, int64_t* n_ops,
                    int64_t* n_refb);
int avo_pack_reads(const avo_text_reads* in, avo_packed_out* out);
/* The same over n_threads host threads (contiguous chunks of the input; identical output). */
int avo_pack_reads_mt(const avo_text_reads* in, avo_packed_out* out, int n_threads);
/* The same on the device: the text columns (in_mem: AVO_MEM_HOST, ideally page-locked, or
 * AVO_MEM_DEVICE) are packed into DEVICE buffers `out` (capacities in, used counts out; identical
 * bytes).  The result is an AVO_MEM_DEVICE avo_read_batch for avo_discover / avo_call / ... */
int avo_pack_reads_device(avo_ctx* ctx, const avo_text_reads* in, int32_t in_mem, avo_packed_out* out);

#ifdef __cplusplus
}
#endif
#endif /* AVOCADO_B200_H */
