/* discover_and_call.c -- the hot path through the C ABI alone (no Python, no torch): what a JNI shim
 * does per region bin.  Generates a small synthetic batch on the host, runs avo_discover_and_call
 * with host buffers and prints the first rows.
 *
 *   gcc -O2 -std=c11 -Iinclude examples/discover_and_call.c -o examples/discover_and_call \
 *       -Lavocado_b200 -lavocado_b200 -Wl,-rpath,$PWD/avocado_b200
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "avocado_b200.h"
#include "avocado_b200_synth.h"

#define CHECK(call)                                                                  \
  do {                                                                               \
    int rc__ = (call);                                                               \
    if (rc__ != AVO_OK) {                                                            \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc__, ctx ? avo_last_error(ctx) : ""); \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

int main(int argc, char** argv) {
  avo_ctx* ctx = NULL;
  const int64_t n = argc > 1 ? atoll(argv[1]) : 20000;
  CHECK(avo_ctx_create(0, &ctx));

  /* a packed batch (what avo_pack_reads produces from AlignmentRecords) */
  avo_synth_params sp;
  avo_synth_default_params(&sp);
  sp.n_reads_total = n;
  sp.contig_len = (int32_t)(n * sp.read_len / 40);
  int64_t n_ops = 0, n_refb = 0;
  CHECK(avo_synth_host(&sp, 0, n, &n_ops, &n_refb, NULL));
  const int64_t n_blocks = n * ((sp.read_len + AVO_BLOCK_BASES - 1) / AVO_BLOCK_BASES);
  avo_packed_out pk;
  memset(&pk, 0, sizeof pk);
  pk.n_reads = n; pk.n_blocks = n_blocks; pk.n_ops = n_ops; pk.n_refb = n_refb;
  pk.meta = aligned_alloc(16, (size_t)n * sizeof(avo_read_meta));
  pk.payload = aligned_alloc(16, (size_t)n_blocks * AVO_BLOCK_BYTES);
  pk.ops = malloc((size_t)(n_ops + 1) * sizeof(uint32_t));
  pk.refb = malloc((size_t)n_refb + 1);
  CHECK(avo_synth_host(&sp, 0, n, NULL, NULL, &pk));

  avo_read_batch reads;
  memset(&reads, 0, sizeof reads);
  reads.n_reads = pk.n_reads; reads.n_blocks = pk.n_blocks; reads.n_ops = pk.n_ops; reads.n_refb = pk.n_refb;
  reads.contig_id = 0; reads.mem = AVO_MEM_HOST;
  reads.meta = pk.meta; reads.payload = pk.payload; reads.ops = pk.ops; reads.refb = pk.refb;

  avo_params prm;
  memset(&prm, 0, sizeof prm);
  prm.ploidy = 2; prm.max_quality = 93; prm.max_mapq = 93;
  prm.min_phred_discover = 25; prm.min_obs_discover = 5;

  /* outputs: capacity in, counts out; every column is caller-owned */
  const int64_t cap = 1 << 14, ns = prm.ploidy + 1;
  avo_sites_out so;
  memset(&so, 0, sizeof so);
  so.capacity = cap; so.allele_capacity = 8 * cap; so.mem = AVO_MEM_HOST;
  so.start = malloc(cap * 4); so.ref_len = malloc(cap * 4); so.alt_len = malloc(cap * 4);
  so.allele_off = malloc((cap + 1) * 4); so.alleles4 = malloc(4 * cap + 1); so.count = malloc(cap * 4);
  avo_site_results res;
  memset(&res, 0, sizeof res);
  res.n_sites = cap; res.n_states = (int32_t)ns; res.mem = AVO_MEM_HOST;
  res.emitted = malloc(cap); res.first_is_ref = malloc(cap);
  res.allele_fwd = malloc(cap * 4); res.other_fwd = malloc(cap * 4); res.allele_cov = malloc(cap * 4);
  res.other_cov = malloc(cap * 4); res.total_cov = malloc(cap * 4); res.copy_number = malloc(cap * 4);
  res.n_alt = malloc(cap * 4); res.n_ref = malloc(cap * 4); res.n_other = malloc(cap * 4); res.gq = malloc(cap * 4);
  res.sb = malloc(cap * 16); res.fisher = malloc(cap * 4); res.rms_mapq = malloc(cap * 4);
  res.square_mapq = malloc(cap * 8); res.qual = malloc(cap * 8);
  res.ref_ll = malloc(cap * ns * 8); res.allele_ll = malloc(cap * ns * 8); res.other_ll = malloc(cap * ns * 8);
  res.nonref_ll = malloc(cap * ns * 8); res.gl = malloc(cap * ns * 8); res.nonref_gl = malloc(cap * ns * 8);

  CHECK(avo_discover_and_call(ctx, &reads, NULL, &prm, &so, &res));

  avo_timing t;
  avo_get_timing(ctx, &t);
  printf("%lld reads -> %lld sites in %.3f ms (%d kernels)\n", (long long)n, (long long)so.n, t.ms_total, t.n_launches);
  const char* NIB = "=ACMGRSVTWYHKDBN";
  for (int64_t i = 0; i < so.n && i < 5; ++i) {
    char ref[64] = {0}, alt[64] = {0};
    const uint32_t off = so.allele_off[i];
    for (int k = 0; k < so.ref_len[i] && k < 63; ++k)
      ref[k] = NIB[(so.alleles4[(off + k) >> 1] >> (((off + k) & 1) ? 0 : 4)) & 15];
    for (int k = 0; k < so.alt_len[i] && k < 63; ++k) {
      const uint32_t o = off + so.ref_len[i] + k;
      alt[k] = NIB[(so.alleles4[o >> 1] >> ((o & 1) ? 0 : 4)) & 15];
    }
    printf("  %d %s>%s  depth %d  alt reads %d  genotype %d alt / %d ref  GQ %d\n", so.start[i], ref, alt,
           res.total_cov[i], res.allele_cov[i], res.n_alt[i], res.n_ref[i], res.gq[i]);
  }
  avo_ctx_destroy(ctx);
  return so.n > 0 ? 0 : 2;
}
