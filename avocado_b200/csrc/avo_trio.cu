// avo_trio.cu -- TrioCaller.processVariant over the (3 samples x sites) matrix, one thread per site
// (avocado-core/.../genotyping/TrioCaller.scala:86-221): sites without an ALT allele are dropped
// (:104-111), missing samples become NO_CALL genotypes (:209-216), and when all three are diploid
// without NO_CALL / OTHER_ALT alleles the child's call is confirmed and phased, left alone (both
// parents het) or invalidated (:150-199).
#include "avo_device.cuh"

namespace avo {

__global__ void k_trio(DTrioIn I, DTrioOut O) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= I.n_sites) return;
  bool has[3];
  int na[3], nr[3], no[3], nn[3];
  for (int s = 0; s < 3; ++s) {
    const int64_t k = (int64_t)s * I.n_sites + i;
    has[s] = !I.present || I.present[k];
    na[s] = has[s] ? I.n_alt[k] : 0; nr[s] = has[s] ? I.n_ref[k] : 0; no[s] = has[s] ? I.n_other[k] : 0;
    nn[s] = (has[s] && I.n_nocall) ? I.n_nocall[k] : 0;
  }
  uint8_t kept = 0, action = AVO_TRIO_UNCHANGED, filled = 0;
  int child_alt = na[2];
  if (na[0] > 0 || na[1] > 0 || na[2] > 0) {                                   // filterRef :104-111
    if (has[0] && has[1] && has[2]) {
      bool ok = true;
      for (int s = 0; s < 3; ++s) ok = ok && (na[s] + nr[s] + no[s] + nn[s] == 2) && nn[s] == 0 && no[s] == 0;   // :143-148
      if (ok) {
        const int s1 = na[0], s2 = na[1], sc = na[2];
        if (sc == 0) action = (s1 == 2 || s2 == 2) ? AVO_TRIO_NO_CALL : AVO_TRIO_PHASED;             // :158-165
        else if (sc == 2) action = (s1 == 0 || s2 == 0) ? AVO_TRIO_NO_CALL : AVO_TRIO_PHASED;        // :166-173
        else if ((s1 == 0 && s2 == 0) || (s1 == 2 && s2 == 2)) action = AVO_TRIO_NO_CALL;            // :175-177
        else if (s1 == 1 && s2 == 1) action = AVO_TRIO_UNCHANGED;                                    // :178-181
        else action = s1 != 0 ? AVO_TRIO_PHASED_ALT_REF : AVO_TRIO_PHASED_REF_ALT;                   // :182-192
      }
    } else {
      filled = (uint8_t)((has[0] ? 0 : 1) | (has[1] ? 0 : 2) | (has[2] ? 0 : 4));
    }
    if (action == AVO_TRIO_NO_CALL) child_alt = 0;
    kept = (na[0] > 0 || na[1] > 0 || child_alt > 0) ? 1 : 0;                  // second filterRef :95
  }
  O.kept[i] = kept; O.child_action[i] = action; O.filled[i] = filled;
}

cudaError_t launch_trio(const DTrioIn& I, const DTrioOut& O, cudaStream_t s) {
  if (I.n_sites == 0) return cudaSuccess;
  k_trio<<<(unsigned)((I.n_sites + 255) / 256), 256, 0, s>>>(I, O);
  return cudaGetLastError();
}

}  // namespace avo
