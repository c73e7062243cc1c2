// avo_observe.cuh -- per-read observation / site classification (Observer.scala:48-140,
// BiallelicGenotyper.scala:195-393) and the per-site genotype emission (BG:579-797), shared by the
// variant-site kernel (avo_call.cu) and the reference-model kernel of scoreAllSites (avo_allsites.cu).
#pragma once

#include "avo_device.cuh"

namespace avo {

constexpr int NSMAX = AVO_MAX_PLOIDY + 1;

// categories (reachable flag combinations of SummarizedObservation)
enum : uint32_t { CAT_REF = 0, CAT_ALT = 1, CAT_OTHER = 2, CAT_NONREF = 3, CAT_NONE = 4, CAT_EXC = 5 };

struct ReadCtx {
  int32_t start, end;
  uint32_t sb;  // first payload block of the read
  uint32_t ob;  // first op
  uint32_t no;  // number of ops
  uint32_t w0;  // the first op (loaded with the record, ahead of the walk)
  uint32_t mapq;
  bool fwd;
};

// (category, optional quality) of one read at one site; q = -1 means None
struct Cls { uint32_t cat; int32_t q; };

// mean quality of an insertion (Observer.scala:101-103): integer division
__device__ __forceinline__ int32_t ins_quality(const uint8_t* __restrict__ pl, uint32_t blk0, uint32_t b, int len) {
  int sum = 0;
  for (int i = 0; i < len; ++i) sum += (int)pl_qual(pl, blk0, b + (uint32_t)i);
  return sum / len;
}

// The same for every site that is not an insertion variant (SNVs, MNPs, deletions: BG:327-355 only looks at
// the FIRST observation's allele and at counts), in two phases: the operator walk touches no payload and only
// notes where the first observation lies; the payload is read once, after the walk, when the lanes of a
// warp have left their differently long operator loops and issue that load together.
__device__ inline Cls classify_by_head(const DReads& R, const ReadCtx& rc, int32_t vs, int32_t ref_len,
                                       int32_t alt_len, uint32_t aoff, const uint8_t* __restrict__ alleles4) {
  const int32_t ve = vs + ref_len;
  const uint32_t altoff = aoff + (uint32_t)ref_len;
  int32_t pos = rc.start;
  uint32_t ridx = 0;
  int n_obs = 0, n_nonempty = 0, n_del = 0, first_del_w = 0;
  int head_kind = 0;             // 0 none, 1 aligned base, 2 insertion, 3 deletion
  uint32_t head_at = 0;          // read index of the head's first base
  int head_len = 0;
  bool head_isref = false;
  // The loop body has no branch on the operator kind (the lanes of a warp hold different kinds at the same
  // k): every kind's overlap test is evaluated and the state is updated through selects.
  for (uint32_t k = 0; k < rc.no; ++k) {
    const uint32_t op = (k == 0) ? rc.w0 : __ldg(R.ops + rc.ob + k);
    const int len = (int)(op >> 4);
    const uint32_t code = op & 15u;
    const bool isMX = code <= AVO_OP_MISMATCH, isI = code == AVO_OP_INS, isD = code == AVO_OP_DEL;
    const int32_t lo = max(pos, vs), hi = min(pos + len, ve);
    const bool hitMX = isMX && lo < hi;                                 // Observer.scala:72-92
    const bool hitI = isI && pos - 1 < ve && pos > vs;                  // Observer.scala:93-118
    const bool hitD = isD && pos < ve && pos + len > vs;                // Observer.scala:119-138
    const bool hit = hitMX || hitI || hitD;
    if (hit && n_obs == 0) {
      head_kind = hitMX ? 1 : hitI ? 2 : 3;
      head_at = ridx + (hitMX ? (uint32_t)(lo - pos) : 0u);
      head_len = len;
      head_isref = code == AVO_OP_MATCH;
    }
    const int cnt = hitMX ? hi - lo : (int)hit;
    n_obs += cnt;
    n_nonempty += hitD ? 0 : cnt;
    first_del_w = (hitD && n_del == 0) ? len : first_del_w;
    n_del += (int)hitD;
    pos += (isMX || isD) ? len : 0;
    ridx += (isMX || isI || code == AVO_OP_SOFTCLIP) ? (uint32_t)len : 0u;      // Observer.scala:66-71: clips
    if (pos > ve) break;  // nothing later in the read can overlap the site
  }
  Cls out;
  out.q = -1;
  bool head_eq_alt = false;
  if (head_kind == 1) {
    uint32_t b, bq;
    pl_base_qual(R.payload, rc.sb, head_at, b, bq);       // base and quality: one 16-byte load
    out.q = (int32_t)bq;
    head_eq_alt = (alt_len == 1 && b == nib_at(alleles4, altoff));
  } else if (head_kind == 2) {
    out.q = ins_quality(R.payload, rc.sb, head_at, head_len);
    bool eq = (head_len == alt_len);
    for (int i = 0; eq && i < head_len; ++i)
      eq = pl_base(R.payload, rc.sb, head_at + i) == nib_at(alleles4, altoff + i);
    head_eq_alt = eq;
  } else if (head_kind == 3) {
    head_eq_alt = (alt_len == 0);
  }
  if (n_obs == 0) { out.cat = CAT_NONE; return out; }                   // BG:297-298
  if (n_nonempty == 1 && head_eq_alt) {                                 // BG:327-355
    if (ref_len > 1 && alt_len == 1)                                    // isDeletion
      out.cat = (n_del == 1 && n_obs == 2 && first_del_w == ref_len - alt_len) ? CAT_ALT : CAT_NONREF;
    else
      out.cat = head_isref ? CAT_REF : CAT_ALT;
  } else if (!head_isref || n_obs != ref_len) {
    out.cat = CAT_NONREF;
  } else {
    out.cat = CAT_REF;
  }
  return out;
}

// S4 + S5 of SURVEY.md for one (read, site): walk the read's operators, summarise the observations
// that overlap [vs, vs + ref_len) in read order, then apply BG:297-355.
__device__ inline Cls classify(const DReads& R, const ReadCtx& rc, int32_t vs, int32_t ref_len,
                        int32_t alt_len, uint32_t aoff, const uint8_t* __restrict__ alleles4) {
  const int32_t ve = vs + ref_len;
  const bool insVar = (ref_len == 1 && alt_len > 1);                    // BG:409-412
  if (!insVar) return classify_by_head(R, rc, vs, ref_len, alt_len, aoff, alleles4);
  const uint32_t altoff = aoff + (uint32_t)ref_len;
  int32_t pos = rc.start;
  uint32_t ridx = 0;
  int n_obs = 0, n_nonempty = 0, n_del = 0, first_del_w = 0;
  bool head_isref = false, head_eq_alt = false;
  int32_t head_q = -1;
  int n_ins_eq = 0;
  int32_t inseq_q = -1;
  bool inseq_isref = false;
  int lead_kind = 0;  // 0 none, 1 non-empty allele, 2 empty allele (deletion)
  uint32_t lead_first = 0;
  bool all_isref = true;

  for (uint32_t k = 0; k < rc.no; ++k) {
    const uint32_t op = __ldg(R.ops + rc.ob + k);
    const int len = (int)(op >> 4);
    const uint32_t code = op & 15u;
    if (code == AVO_OP_SOFTCLIP) { ridx += len; continue; }             // Observer.scala:66-71
    if (code == AVO_OP_HARDCLIP) continue;
    if (code == AVO_OP_MATCH || code == AVO_OP_MISMATCH) {              // Observer.scala:72-92
      const int32_t lo = max(pos, vs), hi = min(pos + len, ve);
      if (lo < hi) {
        const bool isref = (code == AVO_OP_MATCH);
        for (int32_t p = lo; p < hi; ++p) {
          const uint32_t bi = ridx + (uint32_t)(p - pos);   // index within the read
          const bool first = (n_obs == 0);
          ++n_obs; ++n_nonempty;
          all_isref = all_isref && isref;
          if (first || insVar) {
            uint32_t b, bq;
            pl_base_qual(R.payload, rc.sb, bi, b, bq);       // base and quality: one 16-byte load
            if (first) {
              head_isref = isref;
              head_q = (int32_t)bq;
              head_eq_alt = (alt_len == 1 && b == nib_at(alleles4, altoff));
            }
            if (insVar) {
              const bool eq = (alt_len == 2 && b == nib_at(alleles4, altoff + 1));
              if (eq) {
                if (n_ins_eq == 0) { inseq_q = (int32_t)bq; inseq_isref = isref; }
                ++n_ins_eq;
              } else if (lead_kind == 0) { lead_kind = 1; lead_first = b; }
            }
          }
        }
      }
      pos += len; ridx += len;
    } else if (code == AVO_OP_INS) {                                    // Observer.scala:93-118
      if (pos - 1 < ve && pos > vs) {
        const bool first = (n_obs == 0);
        ++n_obs; ++n_nonempty;
        all_isref = false;
        if (first) {
          head_isref = false;
          head_q = ins_quality(R.payload, rc.sb, ridx, len);
          bool eq = (len == alt_len);
          for (int i = 0; eq && i < len; ++i)
            eq = pl_base(R.payload, rc.sb, ridx + i) == nib_at(alleles4, altoff + i);
          head_eq_alt = eq;
        }
        if (insVar) {
          bool eq = (len == alt_len - 1);
          for (int i = 0; eq && i < len; ++i)
            eq = pl_base(R.payload, rc.sb, ridx + i) == nib_at(alleles4, altoff + 1 + i);
          if (eq) {
            if (n_ins_eq == 0) { inseq_q = ins_quality(R.payload, rc.sb, ridx, len); inseq_isref = false; }
            ++n_ins_eq;
          } else if (lead_kind == 0) { lead_kind = 1; lead_first = pl_base(R.payload, rc.sb, ridx); }
        }
      }
      ridx += len;
    } else if (code == AVO_OP_DEL) {                                    // Observer.scala:119-138
      if (pos < ve && pos + len > vs) {
        const bool first = (n_obs == 0);
        ++n_obs;
        all_isref = false;
        if (n_del == 0) first_del_w = len;
        ++n_del;
        if (first) { head_isref = false; head_q = -1; head_eq_alt = (alt_len == 0); }
        if (insVar && lead_kind == 0) lead_kind = 2;
      }
      pos += len;
    }
    if (pos > ve) break;  // nothing later in the read can overlap the site
  }

  Cls out;
  out.q = head_q;
  if (n_obs == 0) { out.cat = CAT_NONE; return out; }                   // BG:297-298
  if (insVar) {                                                         // BG:307-326
    if (n_obs == 2 && n_ins_eq == 1) {
      if (lead_kind == 2) { out.cat = CAT_EXC; return out; }            // leadBaseObserved(0) on ""
      out.q = inseq_q;
      if (lead_first == nib_at(alleles4, altoff)) out.cat = inseq_isref ? CAT_REF : CAT_ALT;
      else out.cat = CAT_NONREF;
    } else if (all_isref) {
      out.cat = CAT_REF;
    } else {
      out.cat = CAT_NONREF;
    }
  } else {                                                              // BG:327-355
    if (n_nonempty == 1 && head_eq_alt) {
      if (ref_len > 1 && alt_len == 1) {                                // isDeletion
        out.cat = (n_del == 1 && n_obs == 2 && first_del_w == ref_len - alt_len) ? CAT_ALT : CAT_NONREF;
      } else {
        out.cat = head_isref ? CAT_REF : CAT_ALT;
      }
    } else if (!head_isref || n_obs != ref_len) {
      out.cat = CAT_NONREF;
    } else {
      out.cat = CAT_REF;
    }
  }
  return out;
}

__device__ __forceinline__ Cls classify_site(const DReads& R, const ReadCtx& rc, const DSites& S,
                                             int32_t j) {
  return classify(R, rc, __ldg(S.start + j), __ldg(S.ref_len + j), __ldg(S.alt_len + j),
                  __ldg(S.allele_off + j), S.alleles4);
}

// BG:376-383 with CopyNumberMap.overlappingVariants (CopyNumberMap.scala:103-111): the CNVs seen by
// a read are the first contiguous run of list entries overlapping it; the site takes the first of
// those that overlaps it, else the base ploidy.
__device__ inline int32_t copy_number_for(const DCnv& C, int32_t base, int32_t rs, int32_t re, int32_t vs,
                                   int32_t ve) {
  if (C.n == 0) return base;
  int32_t i = 0;
  while (i < C.n && !(__ldg(C.end + i) > rs && __ldg(C.start + i) < re)) ++i;
  while (i < C.n && (__ldg(C.end + i) > rs && __ldg(C.start + i) < re)) {
    if (vs < __ldg(C.end + i) && ve > __ldg(C.start + i)) return __ldg(C.cn + i);
    ++i;
  }
  return base;
}

// ---- per-site accumulators in shared memory ---------------------------------------------------
struct SiteAcc {
  double ll[4][NSMAX];        // [category][state]
  unsigned long long sq;      // sum of mapQ^2 (exact integer)
  int32_t allele_fwd, other_fwd, allele_cov, other_cov, total_cov;
  int32_t first_is_ref, copy_number, seen;
};

// BG:755-762 logFactorial: table built by k_lnfact in the reference's accumulation order; beyond
// the table the same descending recurrence is run directly
__device__ inline double log_factorial(const DCallParams& P, int n) {
  if (n < P.lnk_n) return n > 1 ? __ldg(P.lnfact + n) : 0.0;
  double f = 0.0;
  while (n > 1) {
    const double l = (n < P.lnk_n) ? __ldg(P.lnk + n) : log((double)n);
    f = l + f;
    --n;
  }
  return f;
}

// BG:622-668
__device__ inline void state_and_quality(const double* ll, int n, double k, int& state, double& qual) {
  int mi; double mx, sec;
  if (ll[0] >= ll[1]) { mi = 0; mx = ll[0]; sec = ll[1]; } else { mi = 1; mx = ll[1]; sec = ll[0]; }
  for (int i = 2; i < n; ++i) {
    const double c = ll[i];
    if (c > mx) { sec = mx; mx = c; mi = i; } else if (c > sec) { sec = c; }
  }
  state = mi;
  qual = k * (mx - sec);
}

__device__ __forceinline__ int32_t double_to_int(double d) {   // Scala Double.toInt
  if (d != d) return 0;
  if (d >= 2147483647.0) return 2147483647;
  if (d <= -2147483648.0) return (int32_t)0x80000000;
  return (int32_t)d;
}

// BG:579-616, 677-748 for one reduced site; writes the result row.
// a row that received no observation: every column zero (the caller's arrays need not be cleared)
__device__ inline void zero_site(const DResults& O, int32_t site, int ns) {
  O.emitted[site] = 0;
  O.allele_fwd[site] = 0; O.other_fwd[site] = 0; O.allele_cov[site] = 0; O.other_cov[site] = 0; O.total_cov[site] = 0;
  O.square_mapq[site] = 0.0; O.first_is_ref[site] = 0; O.copy_number[site] = 0;
  O.n_alt[site] = 0; O.n_ref[site] = 0; O.n_other[site] = 0; O.qual[site] = 0.0; O.gq[site] = 0;
  O.fisher[site] = 0.f; O.rms_mapq[site] = 0.f;
  for (int k = 0; k < 4; ++k) O.sb[(int64_t)site * 4 + k] = 0;
  for (int g = 0; g < ns; ++g) {
    const int64_t i = (int64_t)site * ns + g;
    O.ref_ll[i] = 0.0; O.allele_ll[i] = 0.0; O.other_ll[i] = 0.0; O.nonref_ll[i] = 0.0; O.gl[i] = 0.0; O.nonref_gl[i] = 0.0;
  }
}

template <bool ZERO_UNSEEN = false>
__device__ inline void finalize_site(const DCallParams& P, const DResults& O, int32_t site, const SiteAcc& A) {
  const int ns = P.ns;
  if (!A.seen) {
    if (ZERO_UNSEEN) zero_site(O, site, ns); else O.emitted[site] = 0;
    return;
  }
  O.emitted[site] = 1;
  O.allele_fwd[site] = A.allele_fwd; O.other_fwd[site] = A.other_fwd;
  O.allele_cov[site] = A.allele_cov; O.other_cov[site] = A.other_cov; O.total_cov[site] = A.total_cov;
  const double sq = (double)A.sq;
  O.square_mapq[site] = sq;
  O.first_is_ref[site] = (uint8_t)A.first_is_ref;
  O.copy_number[site] = A.copy_number;
  for (int g = 0; g < ns; ++g) {
    O.ref_ll[(int64_t)site * ns + g] = A.ll[CAT_REF][g];
    O.allele_ll[(int64_t)site * ns + g] = A.ll[CAT_ALT][g];
    O.other_ll[(int64_t)site * ns + g] = A.ll[CAT_OTHER][g];
    O.nonref_ll[(int64_t)site * ns + g] = A.ll[CAT_NONREF][g];
  }
  const int n = A.copy_number + 1;
  double sc[NSMAX];
  int s_ar, s_ao, s_or; double q_ar, q_ao, q_or;
  for (int i = 0; i < n; ++i) sc[i] = A.ll[CAT_ALT][i] + A.ll[CAT_REF][n - 1 - i];
  state_and_quality(sc, n, P.ten_div_log10, s_ar, q_ar);
  for (int i = 0; i < n; ++i) {
    O.gl[(int64_t)site * ns + i] = sc[i];                                       // BG:728
    O.nonref_gl[(int64_t)site * ns + i] = A.ll[CAT_NONREF][i] + A.ll[CAT_REF][n - 1 - i];  // BG:729
  }
  for (int i = n; i < ns; ++i) { O.gl[(int64_t)site * ns + i] = 0.0; O.nonref_gl[(int64_t)site * ns + i] = 0.0; }
  for (int i = 0; i < n; ++i) sc[i] = A.ll[CAT_ALT][i] + A.ll[CAT_OTHER][n - 1 - i];
  state_and_quality(sc, n, P.ten_div_log10, s_ao, q_ao);
  for (int i = 0; i < n; ++i) sc[i] = A.ll[CAT_OTHER][i] + A.ll[CAT_REF][n - 1 - i];
  state_and_quality(sc, n, P.ten_div_log10, s_or, q_or);
  int alt, ref, other; double qual;
  if (q_ar >= q_ao && q_ar >= q_or) { alt = s_ar; ref = n - s_ar - 1; other = 0; qual = q_ar; }
  else if (q_ao >= q_or) { alt = s_ao; ref = 0; other = n - s_ao - 1; qual = q_ao; }
  else { alt = 0; ref = n - s_or - 1; other = s_or; qual = q_or; }
  O.n_alt[site] = alt; O.n_ref[site] = ref; O.n_other[site] = other;
  O.qual[site] = qual;
  O.gq[site] = double_to_int(qual);
  const int ofw = A.other_fwd, orv = A.other_cov - A.other_fwd, afw = A.allele_fwd,
            arv = A.allele_cov - A.allele_fwd;
  O.sb[(int64_t)site * 4 + 0] = ofw; O.sb[(int64_t)site * 4 + 1] = orv;
  O.sb[(int64_t)site * 4 + 2] = afw; O.sb[(int64_t)site * 4 + 3] = arv;
  const double num = log_factorial(P, ofw + orv) + log_factorial(P, ofw + afw) +
                     log_factorial(P, afw + arv) + log_factorial(P, orv + arv);
  const double den = log_factorial(P, ofw) + log_factorial(P, orv) + log_factorial(P, afw) +
                     log_factorial(P, arv) + log_factorial(P, ofw + orv + afw + arv);
  O.fisher[site] = (float)(-10.0 * (num - den) / P.log10);                      // LogPhred.scala:38-40
  O.rms_mapq[site] = (float)sqrt(sq / (double)(A.allele_cov + A.other_cov));    // BG:709
}


}  // namespace avo
