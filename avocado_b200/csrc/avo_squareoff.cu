// avo_squareoff.cu -- SquareOffReferenceModel.squareOffSite for one sample against the cohort's
// site list: the merge-join between a sample's gVCF rows and the sites
// (avocado-core/.../genotyping/SquareOffReferenceModel.scala:176-244).
//
//   hasScoredVariant (:198-208)   a genotype of the sample at exactly this variant (start, end,
//                                 alleles) represents it as is;
//   exciseGenotypeFromReferenceModel (:224-244)  else the first genotype of the sample whose region
//                                 overlaps the variant does, with genotypeLikelihoods :=
//                                 nonReferenceLikelihoods.  "First" is Spark's grouping order in the
//                                 reference; the canonical order here and in the oracle is lowest
//                                 (start, end), a genotype with an alternate allele before a
//                                 reference-model one.
// One thread per cohort site; the sample's variant rows and reference-model rows are both sorted
// by start, so both look-ups are binary searches plus a short scan.  Reference-model rows cover one
// position (what avo_call_all_sites emits) or, with r_len, a block of positions (banded gVCFs).
#include "avo_device.cuh"

namespace avo {

__device__ __forceinline__ bool alleles_equal(const uint8_t* __restrict__ a4, uint32_t ao, const uint8_t* __restrict__ b4,
                                              uint32_t bo, int n) {
  for (int t = 0; t < n; ++t)
    if (nib_at(a4, ao + t) != nib_at(b4, bo + t)) return false;
  return true;
}

__global__ void k_square_off(DSquareIn I, DSquareOut O) {
  const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= I.n_sites) return;
  const int ns = I.ns;
  const int32_t vs = __ldg(I.c_start + j), rl = __ldg(I.c_ref_len + j), al = __ldg(I.c_alt_len + j);
  const int32_t ve = vs + rl;
  int32_t kind = AVO_SQ_ABSENT, src = -1;
  // exact variant row
  int32_t lb = lower_bound_i32(I.v_start, 0, I.n_v, vs);
  for (int32_t k = lb; k < I.n_v && __ldg(I.v_start + k) == vs; ++k) {
    if (!I.v_emitted[k] || __ldg(I.v_ref_len + k) != rl || __ldg(I.v_alt_len + k) != al) continue;
    if (alleles_equal(I.c_alleles4, __ldg(I.c_allele_off + j), I.v_alleles4, __ldg(I.v_allele_off + k), rl + al)) {
      kind = AVO_SQ_SCORED; src = k;
      break;
    }
  }
  if (kind == AVO_SQ_ABSENT) {
    // first overlapping genotype: lowest (start, end); a variant row before a reference-model row
    int32_t bs = INT32_MAX, be = INT32_MAX;
    const int64_t lo_key = (int64_t)vs - I.v_max_ref_len + 1;
    for (int32_t k = lower_bound_i32(I.v_start, 0, I.n_v, (int32_t)max(lo_key, (int64_t)INT32_MIN));
         k < I.n_v && __ldg(I.v_start + k) < ve; ++k) {
      const int32_t s = __ldg(I.v_start + k), e = s + __ldg(I.v_ref_len + k);
      if (!I.v_emitted[k] || e <= vs) continue;
      if (s < bs || (s == bs && e < be)) { bs = s; be = e; kind = AVO_SQ_EXCISED_VARIANT; src = k; }
    }
    if (!I.r_len) {
      for (int32_t k = lower_bound_i32(I.r_start, 0, I.n_r, vs); k < I.n_r && __ldg(I.r_start + k) < ve; ++k) {
        if (!I.r_emitted[k]) continue;
        const int32_t s = __ldg(I.r_start + k), e = s + 1;
        if (s < bs || (s == bs && e < be)) { bs = s; be = e; kind = AVO_SQ_EXCISED_REF_MODEL; src = k; }
        break;                                // later reference-model rows start later
      }
    } else {
      // blocks: a row that starts before the site may still reach it
      const int64_t rlo = (int64_t)vs - I.r_max_len + 1;
      for (int32_t k = lower_bound_i32(I.r_start, 0, I.n_r, (int32_t)max(rlo, (int64_t)INT32_MIN));
           k < I.n_r && __ldg(I.r_start + k) < ve; ++k) {
        const int32_t s = __ldg(I.r_start + k), e = s + __ldg(I.r_len + k);
        if (!I.r_emitted[k] || e <= vs) continue;
        if (s < bs || (s == bs && e < be)) { bs = s; be = e; kind = AVO_SQ_EXCISED_REF_MODEL; src = k; }
      }
    }
  }
  O.kind[j] = (uint8_t)kind;
  O.src[j] = src;
  int32_t na = 0, nr = 0, no = 0;
  const double* gl = nullptr;
  if (kind == AVO_SQ_SCORED) { na = I.v_n_alt[src]; nr = I.v_n_ref[src]; no = I.v_n_other[src]; gl = I.v_gl + (size_t)src * ns; }
  else if (kind == AVO_SQ_EXCISED_VARIANT) { na = I.v_n_alt[src]; nr = I.v_n_ref[src]; no = I.v_n_other[src]; gl = I.v_nonref_gl + (size_t)src * ns; }
  else if (kind == AVO_SQ_EXCISED_REF_MODEL) { na = I.r_n_alt[src]; nr = I.r_n_ref[src]; no = I.r_n_other[src]; gl = I.r_nonref_gl + (size_t)src * ns; }
  O.n_alt[j] = na; O.n_ref[j] = nr; O.n_other[j] = no;
  for (int g = 0; g < ns; ++g) O.gl[(size_t)j * ns + g] = gl ? gl[g] : 0.0;
}

cudaError_t launch_square_off(const DSquareIn& I, const DSquareOut& O, cudaStream_t s) {
  if (I.n_sites == 0) return cudaSuccess;
  k_square_off<<<(I.n_sites + 127) / 128, 128, 0, s>>>(I, O);
  return cudaGetLastError();
}

}  // namespace avo
