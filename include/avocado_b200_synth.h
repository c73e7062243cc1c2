/* avocado_b200_synth.h -- seeded synthetic read generator exported by libavocado_b200.so.
 *
 * Bench / test infrastructure, NOT part of the reference's interface: it produces the packed
 * batches of avocado_b200.h directly (SURVEY.md 8(d) shapes) so that full-size benchmark inputs
 * are born in HBM.  The same stateless generator runs on the host for the small cases that are
 * handed to the CPU oracle; both produce identical bytes.
 */
#ifndef AVOCADO_B200_SYNTH_H
#define AVOCADO_B200_SYNTH_H

#include <stdint.h>

#include "avocado_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  uint64_t seed;
  int32_t contig_len;     /* reads start in [first_pos, first_pos + contig_len - read_len - slack) */
  int32_t first_pos;
  int64_t n_reads_total;  /* reads of the whole contig; read i starts near i/n_reads_total of it */
  int32_t read_len;
  int32_t max_indel;      /* <= 20 */
  double snp_rate;        /* per reference base */
  double indel_rate;
  double triallelic_frac; /* fraction of SNV sites carrying two different alts */
  double softclip_frac;   /* fraction of reads soft-clipped 1..20 bases at one end */
  uint64_t sample;        /* 0: the default sample.  Other values: another sample of the same cohort:
                             same reference and truth variant sites / alleles, its own genotypes
                             (carrier with p = 3/4, then het : hom-alt = 2 : 1) and its own reads */
} avo_synth_params;

int avo_synth_default_params(avo_synth_params* p);

/* Host generation of reads [first, first+n).  out == NULL: only count (n_ops / n_refb totals). */
int avo_synth_host(const avo_synth_params* p, int64_t first, int64_t n, int64_t* n_ops,
                   int64_t* n_refb, avo_packed_out* out);

/* Device generation, two phases (see avo_synth.cu).  All array pointers are device pointers. */
int avo_synth_device(const avo_synth_params* p, int64_t first, int64_t n, uint32_t* op_off_dev,
                     uint32_t* refb_off_dev, int64_t* n_ops, int64_t* n_refb, avo_packed_out* out,
                     void* cuda_stream);

/* Text form (CIGAR with M runs, MD tag, sequence, phred+33 qualities) of reads [first, first+n) of a
 * packed HOST batch: the input avo_pack_reads takes.  Call once with cigar == NULL to get the blob
 * sizes (cigar_bytes / md_bytes / seq_bytes are then outputs), then with buffers; the offset arrays
 * hold n + 1 entries and qual shares seq's offsets. */
typedef struct {
  int64_t cigar_bytes, md_bytes, seq_bytes; /* in: capacities; out: used */
  int64_t* start; int64_t* end; int32_t* mapq; uint8_t* rec_flags;
  char* cigar; int64_t* cigar_off;
  char* md;    int64_t* md_off;
  char* seq;   char* qual; int64_t* seq_off;
} avo_text_out;

int avo_synth_text(const avo_read_batch* batch, int64_t first, int64_t n, avo_text_out* out);

#ifdef __cplusplus
}
#endif
#endif
