// avo_joint.cu -- JointAnnotatorCaller on the GPU: cross-sample allele frequency, binomial
// log-prior, log-sum-exp normalisation, re-call and site quality.
//
// Reference semantics (avocado-core/src/main/scala/org/bdgenomics/avocado/):
//   genotyping/JointAnnotatorCaller.scala:74-106  annotateSite      (sites with MAF <= 0 dropped)
//   :117-128  calculateMinorAlleleFrequency  alt alleles / called alleles over the site's genotypes
//   :156-179  recallGenotypes   prior[c] = Binomial(copies, maf).logProbabilityOf(c)   (breeze 0.13.2)
//   :181-262  recallGenotype    posterior = logNormalize(prior + genotypeLikelihoods), new state / GQ
//   :270-281  computeQuality    -10/ln10 * sum over genotypes of posterior(0) (as float)
//   util/LogUtils.scala:39-66   sumLogProbabilities (sorted descending, log1p/exp chain), logNormalize
//
// Layout: S samples x N sites, sample-major ([s * N + i]); one thread per site walks the samples
// in order, so the per-site sums are produced in sample order.  Samples can be sharded over
// processes: avo_joint_counts gives the local allele counts, the caller all-reduces them (NCCL),
// avo_joint then takes the global counts and returns the local sum of posterior(0) per site,
// which the caller all-reduces into the site quality.
#include <cmath>

#include "avo_device.cuh"

namespace avo {

__global__ void k_joint_counts(DJointIn J, int32_t* __restrict__ alt, int32_t* __restrict__ called) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= J.n_sites) return;
  int32_t a = 0, c = 0;
  for (int s = 0; s < J.n_samples; ++s) {
    const int64_t k = (int64_t)s * J.n_sites + i;
    if (J.present && !J.present[k]) continue;
    const int32_t na = J.n_alt[k];
    a += na;                                                     // :124-126
    c += na + J.n_ref[k] + J.n_other[k];                         // :121-123 (everything but NO_CALL)
  }
  alt[i] = a;
  called[i] = c;
}

// calculateAnnotations (JointAnnotatorCaller.scala:141-147): VariantSummary of every genotype of the site,
// merged (VariantSummary.scala:95-119: a field is the sum over the genotypes that have it).  One thread per
// site; the loads of a warp are 32 neighbouring sites of one sample (coalesced).
__global__ void k_joint_depths(DDepthsIn I, DDepthsOut O) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= I.n_sites) return;
  int32_t sum[6] = {0, 0, 0, 0, 0, 0};
  bool has[6] = {false, false, false, false, false, false};
  int32_t n_gt = 0;
  for (int s = 0; s < I.n_samples; ++s) {
    const int64_t k = (int64_t)s * I.n_sites + i;
    if (I.present && !I.present[k]) continue;
    ++n_gt;
    if (I.read_depth) { const int32_t v = __ldg(I.read_depth + k); if (v >= 0) { sum[0] += v; has[0] = true; } }
    if (I.ref_read_depth) { const int32_t v = __ldg(I.ref_read_depth + k); if (v >= 0) { sum[1] += v; has[1] = true; } }
    if (I.sb) {
      const int4 b = __ldg(reinterpret_cast<const int4*>(I.sb) + k);      // ref fwd, ref rev, alt fwd, alt rev
      if (b.x >= 0) {
        sum[2] += b.x + b.z; sum[3] += b.y + b.w; sum[4] += b.x; sum[5] += b.y;   // VariantSummary.scala:52-65
        has[2] = has[3] = has[4] = has[5] = true;
      }
    }
  }
  O.annotated[i] = n_gt > 1;                                                // :85-90
  for (int f = 0; f < 6; ++f) O.col[f][i] = has[f] ? sum[f] : -1;
}

cudaError_t launch_joint_depths(const DDepthsIn& I, const DDepthsOut& O, cudaStream_t s) {
  if (I.n_sites == 0) return cudaSuccess;
  k_joint_depths<<<(unsigned)((I.n_sites + 127) / 128), 128, 0, s>>>(I, O);
  return cudaGetLastError();
}

// LogUtils.sumLogProbabilities over n <= NS values (sorted descending, stable)
template <int NS>
__device__ double sum_log_probabilities(const double* p, int n) {
  double s[NS];
  for (int i = 0; i < n; ++i) s[i] = p[i];
  for (int i = 1; i < n; ++i) {            // stable insertion sort, descending (sortWith(_ > _))
    const double v = s[i];
    int j = i - 1;
    while (j >= 0 && v > s[j]) { s[j + 1] = s[j]; --j; }
    s[j + 1] = v;
  }
  double acc = s[0];
  for (int i = 1; i < n; ++i) acc = __dadd_rn(acc, log1p(exp(__dadd_rn(s[i], -acc))));   // :80-84
  return acc;
}

// BG:622-668 (the same scan as the single-sample genotyper)
__device__ void joint_state_and_quality(const double* ll, int n, double k, int& state, double& qual) {
  int mi; double mx, sec;
  if (ll[0] >= ll[1]) { mi = 0; mx = ll[0]; sec = ll[1]; } else { mi = 1; mx = ll[1]; sec = ll[0]; }
  for (int i = 2; i < n; ++i) {
    const double c = ll[i];
    if (c > mx) { sec = mx; mx = c; mi = i; } else if (c > sec) { sec = c; }
  }
  state = mi;
  qual = k * (mx - sec);
}

__device__ __forceinline__ int32_t joint_to_int(double d) {   // Scala Double.toInt
  if (d != d) return 0;
  if (d >= 2147483647.0) return 2147483647;
  if (d <= -2147483648.0) return (int32_t)0x80000000;
  return (int32_t)d;
}

constexpr int JNS = AVO_MAX_PLOIDY + 1;

__global__ void k_joint_recall(DJointIn J, const int32_t* __restrict__ alt, const int32_t* __restrict__ called,
                               DJointOut O, const double* __restrict__ lgam /* lgamma(k), k <= 9 */,
                               double ten_div_log10) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= J.n_sites) return;
  const int ns = J.n_states;
  const double maf = (double)alt[i] / (double)called[i];         // :128
  const bool emitted = !(maf <= 0.0);                            // :81 (NaN passes, as in Scala)
  O.site_emitted[i] = emitted ? 1 : 0;
  O.maf[i] = maf;
  const bool priors_defined = !(maf >= 1.0 || maf <= 0.0);       // :162-163
  const double lp = log(maf), lq = log(1.0 - maf);
  double post0 = 0.0;
  for (int s = 0; s < J.n_samples; ++s) {
    const int64_t k = (int64_t)s * J.n_sites + i;
    uint8_t flags = 0;
    int32_t o_alt = J.n_alt[k], o_ref = J.n_ref[k], o_gq = 0;
    const bool present = emitted && (!J.present || J.present[k]);
    if (present) {
      const int nocall = J.n_nocall ? J.n_nocall[k] : 0;
      const int others = J.n_other[k] + nocall;                  // :184-187
      const int copies = o_alt + o_ref + others;                 // gt.getAlleles.size
      const int n = min(copies + 1, ns);                         // genotypeLikelihoods.size
      const double* gl = J.gl + k * ns;
      double nn[JNS], prior[JNS];
      // allele counts come from the caller: a negative count is not a genotype (left untouched, flags 0);
      // more copies than the lgamma table holds (AVO_MAX_PLOIDY + 2) fall back to the device lgamma
      const bool sane = o_alt >= 0 && o_ref >= 0 && others >= 0;
      auto lg = [&](int v) { return v <= AVO_MAX_PLOIDY + 2 ? lgam[v] : lgamma((double)v); };
      if (!sane) {
      } else if (priors_defined && others == 0 && n >= 1) {      // :191-234
        for (int c = 0; c < n; ++c) {                            // breeze Binomial.logProbabilityOf
          const double comb = __dadd_rn(__dadd_rn(lg(copies + 1), -lg(c + 1)), -lg(copies - c + 1));
          prior[c] = __dadd_rn(__dadd_rn(comb, __dmul_rn((double)c, lp)), __dmul_rn((double)(copies - c), lq));
          nn[c] = __dadd_rn(prior[c], gl[c]);
        }
        const double sum = sum_log_probabilities<JNS>(nn, n);
        for (int c = 0; c < n; ++c) {
          O.posteriors[k * ns + c] = (float)__dadd_rn(nn[c], -sum);
          O.priors[k * ns + c] = (float)prior[c];
        }
        int state = 0; double q = 0.0;
        if (n >= 2) joint_state_and_quality(nn, n, ten_div_log10, state, q);
        o_alt = state; o_ref = copies - state; o_gq = joint_to_int(q);
        flags = 1 | 2 | 4;
        post0 += (double)O.posteriors[k * ns];                   // :272-279 posteriors.headOption
      } else if (others == 0 && n >= 1) {                        // :235-250
        for (int c = 0; c < n; ++c) nn[c] = gl[c];
        const double sum = sum_log_probabilities<JNS>(nn, n);
        for (int c = 0; c < n; ++c) O.posteriors[k * ns + c] = (float)__dadd_rn(nn[c], -sum);
        flags = 1;
        post0 += (double)O.posteriors[k * ns];
      }
    }
    O.flags[k] = flags;
    O.n_alt[k] = o_alt; O.n_ref[k] = o_ref; O.gq[k] = o_gq;
  }
  O.post0_sum[i] = post0;
  O.quality[i] = __dmul_rn(-ten_div_log10, post0);               // :270
}

cudaError_t launch_joint_counts(const DJointIn& J, int32_t* alt, int32_t* called, cudaStream_t s) {
  if (J.n_sites == 0) return cudaSuccess;
  k_joint_counts<<<(unsigned)((J.n_sites + 255) / 256), 256, 0, s>>>(J, alt, called);
  return cudaGetLastError();
}

cudaError_t launch_joint_recall(const DJointIn& J, const int32_t* alt, const int32_t* called,
                                const DJointOut& O, const double* lgam, cudaStream_t s) {
  if (J.n_sites == 0) return cudaSuccess;
  k_joint_recall<<<(unsigned)((J.n_sites + 127) / 128), 128, 0, s>>>(J, alt, called, O, lgam,
                                                                      10.0 / std::log(10.0));
  return cudaGetLastError();
}

}  // namespace avo
