// avo_filter.cu -- RewriteHets and HardFilterGenotypes over the per-site result columns, one thread
// per row (the two passes the CLI applies to call()'s output, CLI/BiallelicGenotyper.scala:279-282).
//
// Reference semantics (avocado-core/src/main/scala/org/bdgenomics/avocado/util/):
//   RewriteHets.scala:95-126   shouldRewrite: het call whose alt allelic fraction reaches the
//                              threshold -> alleles all ALT, genotypeQuality unset (:134-143)
//   HardFilterGenotypes.scala:346-371  emission: (an ALT allele unless filterRefGenotypes is off)
//                              and genotypeQuality > minQuality (unset quality passes)
//   :632-660  filterGenotype: a genotype without alternate allele (reference model) always comes
//                              through; others are dropped when not emitted, else take the SNP or
//                              the INDEL filter list
//   :381-580  the nine filters per class (QD het / hom, RMS mapQ, Fisher strand bias, min / max
//                              depth, min AF het, max AF het, min AF hom); a filter is on iff its
//                              threshold is > 0 (:258-344)
#include "avo_device.cuh"

namespace avo {

__global__ void k_filter_genotypes(DFilterIn I, avo_filter_params A, DFilterOut O) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= I.n) return;
  uint8_t emit = 0, rewritten = 0, applied = 0;
  uint32_t failed = 0;
  int32_t n_alt = I.n_alt[i], n_ref = I.n_ref[i], n_other = I.n_other[i];
  if (I.emitted && !I.emitted[i]) {          // not a genotype at all (no observation reached the site)
    O.emit[i] = 0; O.rewritten[i] = 0; O.filters_applied[i] = 0; O.filters_failed[i] = 0;
    O.n_alt[i] = n_alt; O.n_ref[i] = n_ref; O.gq_valid[i] = 0;
    return;
  }
  const int32_t alt_len = I.alt_len ? I.alt_len[i] : -1;
  const bool has_alt = alt_len >= 0;                               // reference-model rows carry no alternate
  const bool snp = has_alt && I.ref_len[i] == 1 && alt_len == 1;
  const int32_t copies = n_alt + n_ref + n_other;
  const float dp = (float)I.total_cov[i], adp = (float)I.allele_cov[i];
  const float af = adp / dp;                                        // alt.toFloat / dp.toFloat (NaN when 0 / 0)
  bool gq_valid = true;
  // ---- RewriteHets -----------------------------------------------------------------------------
  if (has_alt) {
    const bool het = n_alt != 0 && n_alt != copies;
    const bool on = snp ? !A.disable_het_snp_rewriting : !A.disable_het_indel_rewriting;
    const float thr = snp ? A.max_het_snp_alt_allelic_fraction : A.max_het_indel_alt_allelic_fraction;
    if (on && het && af >= thr) {
      n_alt = copies; n_ref = 0; n_other = 0;
      gq_valid = false;
      rewritten = 1;
    }
  }
  // ---- HardFilterGenotypes ---------------------------------------------------------------------
  const int32_t gq = I.gq[i];
  const bool qual_ok = !gq_valid || gq > A.min_quality;
  const bool emit_ok = A.emit_all_genotypes || ((A.filter_ref_genotypes ? n_alt > 0 : true) && qual_ok);
  if (!has_alt) {
    emit = 1;
  } else if (emit_ok) {
    emit = 1;
    const bool all_alt = n_alt == copies;                           // alleles.forall(_ == ALT)
    const bool all_alt_or_other = n_ref == 0;                       // no NO_CALL alleles on this path
    const float qd = (float)gq / dp;
    const float rms = I.rms_mapq[i], fs = I.fisher[i];
    const int32_t depth = I.total_cov[i];
    const float min_qd_het = snp ? A.min_het_snp_quality_by_depth : A.min_het_indel_quality_by_depth;
    const float min_qd_hom = snp ? A.min_hom_snp_quality_by_depth : A.min_hom_indel_quality_by_depth;
    const float min_mq = snp ? A.min_snp_rms_mapping_quality : A.min_indel_rms_mapping_quality;
    const float max_fs = snp ? A.max_snp_phred_strand_bias : A.max_indel_phred_strand_bias;
    const int32_t min_dp = snp ? A.min_snp_depth : A.min_indel_depth;
    const int32_t max_dp = snp ? A.max_snp_depth : A.max_indel_depth;
    const float min_af_het = snp ? A.min_het_snp_alt_allelic_fraction : A.min_het_indel_alt_allelic_fraction;
    const float max_af_het = snp ? A.max_het_snp_alt_allelic_fraction : A.max_het_indel_alt_allelic_fraction;
    const float min_af_hom = snp ? A.min_hom_snp_alt_allelic_fraction : A.min_hom_indel_alt_allelic_fraction;
    uint32_t f = 0;
    if (min_qd_het > 0.0f) { applied = 1; if (!all_alt && gq_valid && qd < min_qd_het) f |= 1u << 0; }
    if (min_qd_hom > 0.0f) { applied = 1; if (all_alt && gq_valid && qd < min_qd_hom) f |= 1u << 1; }
    if (min_mq > 0.0f) { applied = 1; if (rms < min_mq) f |= 1u << 2; }
    if (max_fs > 0.0f) { applied = 1; if (fs > max_fs) f |= 1u << 3; }
    if (min_dp > 0) { applied = 1; if (depth < min_dp) f |= 1u << 4; }
    if (max_dp > 0) { applied = 1; if (depth > max_dp) f |= 1u << 5; }
    if (min_af_het > 0.0f) { applied = 1; if (!all_alt && af <= min_af_het) f |= 1u << 6; }
    if (max_af_het > 0.0f) { applied = 1; if (!all_alt_or_other && af > max_af_het) f |= 1u << 7; }
    if (min_af_hom > 0.0f) { applied = 1; if (all_alt && af <= min_af_hom) f |= 1u << 8; }
    failed = snp ? f : (f << 9);
  }
  O.emit[i] = emit; O.rewritten[i] = rewritten; O.filters_applied[i] = applied; O.filters_failed[i] = failed;
  O.n_alt[i] = n_alt; O.n_ref[i] = n_ref; O.gq_valid[i] = gq_valid ? 1 : 0;
}

cudaError_t launch_filter_genotypes(const DFilterIn& I, const avo_filter_params& A, const DFilterOut& O,
                                    cudaStream_t s) {
  if (I.n == 0) return cudaSuccess;
  k_filter_genotypes<<<(unsigned)((I.n + 255) / 256), 256, 0, s>>>(I, A, O);
  return cudaGetLastError();
}

}  // namespace avo
