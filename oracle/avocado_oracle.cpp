// avocado_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A single-threaded-by-default, deliberately literal CPU restatement of the
// genotyping hot path of bigdatagenomics/avocado (Scala/Spark).  It exists to
// CHECK the CUDA product in avocado_b200/; nothing under avocado_b200/ may
// import, link or execute it.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it.
//
// Parity status: PINNED.  tests/test_oracle_*.py run this file against every
// known-answer test and SAM fixture the reference's own suites hold for the
// path (BiallelicGenotyperSuite, DiscoverVariantsSuite, ObserverSuite,
// ObservationOperatorSuite, TreeRegionJoinSuite, JointAnnotatorCallerSuite,
// LogUtilsSuite, LogPhredSuite, CopyNumberMapSuite, VariantSummarySuite,
// ObservationSuite, and the filter / square-off / trio suites); expectations are transcribed into
// tests/golden/ with file:line citations.  The reference itself (Scala 2.11 /
// Spark 2.2 / ADAM 0.24) cannot be compiled or run in this image (no JVM).
//
// Citations are relative to /root/reference/avocado-core/src/main/scala/org/
// bdgenomics/avocado/ ("CORE/").  Third-party behaviour that is not vendored
// in the reference tree is restated from its published contract and named
// where used (ADAM 0.24.0 MdTag / PhredUtils / ReferenceRegion, htsjdk
// TextCigarCodec, Spark SQL 2.2 least / inner join / sum / first, breeze
// 0.13.2 Binomial.logProbabilityOf).
//
// Every function below follows the control flow of the Scala it cites, with
// strings and lists rather than packed buffers, so that it is independent of
// the product's packed layout and kernels.

#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <optional>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../avocado_b200/csrc/avo_synth_core.h"   // the workload generator's source (bench / test input, not the path)

namespace py = pybind11;

namespace orc {

// ---------------------------------------------------------------------------
// Records (bdg-formats Avro AlignmentRecord / Variant, only the fields the path
// reads: SURVEY.md 8b)
// ---------------------------------------------------------------------------
struct Read {
  bool mapped = false;
  std::string contigName;
  bool hasStart = false;
  long start = 0, end = 0;
  bool hasCigar = false;
  std::string cigar;
  bool hasMd = false;
  std::string md;
  std::string seq, qual;
  bool hasMapq = false;
  int mapq = 0;
  bool negativeStrand = false;
};

struct Variant {
  std::string contigName;
  long start = 0, end = 0;
  std::string ref;
  bool hasAlt = true;
  std::string alt;
};

struct Thrown : std::runtime_error {  // stands in for any JVM Throwable
  explicit Thrown(const std::string& s) : std::runtime_error(s) {}
};

// ---------------------------------------------------------------------------
// ADAM 0.24.0 ReferenceRegion (un-vendored): half-open [start,end) on a named
// contig; overlaps = same name && end > o.start && start < o.end; ordering =
// (referenceName string compare, start, end).  Pinned by
// TEST/util/TreeRegionJoinSuite.scala:26-250.
// ---------------------------------------------------------------------------
struct Region {
  std::string name;
  long start, end;
  bool overlaps(const Region& o) const {
    return name == o.name && end > o.start && start < o.end;
  }
  int compareTo(const Region& o) const {
    int c = name.compare(o.name);
    if (c != 0) return c < 0 ? -1 : 1;
    if (start != o.start) return start < o.start ? -1 : 1;
    if (end != o.end) return end < o.end ? -1 : 1;
    return 0;
  }
  long width() const { return end - start; }
};

// ---------------------------------------------------------------------------
// htsjdk TextCigarCodec.decode (un-vendored): <len><op>... ; ops MIDNSHP=X
// ---------------------------------------------------------------------------
struct CigarElem { int length; char op; };

static std::vector<CigarElem> decodeCigar(const std::string& s) {
  std::vector<CigarElem> out;
  size_t i = 0;
  while (i < s.size()) {
    if (!isdigit((unsigned char)s[i])) throw Thrown("Malformed CIGAR string: " + s);
    long n = 0;
    while (i < s.size() && isdigit((unsigned char)s[i])) { n = n * 10 + (s[i] - '0'); ++i; }
    if (i >= s.size()) throw Thrown("Malformed CIGAR string: " + s);
    char op = s[i++];
    if (std::string("MIDNSHP=X").find(op) == std::string::npos) throw Thrown("Unknown CIGAR op");
    out.push_back({(int)n, op});
  }
  return out;
}

// ---------------------------------------------------------------------------
// ADAM 0.24.0 MdTag(mdTag, referenceStart = 0, cigar) (un-vendored): the SAM
// spec grammar [0-9]+(([A-Z]|\^[A-Z]+)[0-9]+)* parsed into two maps keyed by
// reference offset; mismatchedBase(i) / deletedBase(i) are map lookups.
// ---------------------------------------------------------------------------
struct MdTag {
  std::unordered_map<long, char> mismatches, deletions;
  std::optional<char> mismatchedBase(long i) const {
    auto it = mismatches.find(i);
    if (it == mismatches.end()) return std::nullopt;
    return it->second;
  }
  std::optional<char> deletedBase(long i) const {
    auto it = deletions.find(i);
    if (it == deletions.end()) return std::nullopt;
    return it->second;
  }
};

static bool isMdBase(char c) {
  return std::string("ACGTNUKMRSWBVHDX").find(c) != std::string::npos;
}

static MdTag parseMdTag(const std::string& in, long referenceStart) {
  MdTag t;
  if (in.empty()) return t;
  std::string md = in;
  for (auto& c : md) c = (char)toupper((unsigned char)c);
  size_t offset = 0, end = md.size();
  long referencePos = referenceStart;
  auto readMatches = [&](const char* err) {
    size_t j = offset;
    while (j < end && isdigit((unsigned char)md[j])) ++j;
    if (j == offset) throw Thrown(err);
    long len = std::stol(md.substr(offset, j - offset));
    offset = j;
    referencePos += len;
  };
  readMatches("MD tag must start with a digit");
  while (offset < end) {
    bool del = false;
    if (md[offset] == '^') { del = true; ++offset; }
    size_t j = offset;
    while (j < end && isMdBase(md[j])) ++j;
    if (j == offset) throw Thrown("Failed to find deleted or mismatched bases after a match");
    for (size_t k = offset; k < j; ++k) {
      if (del) t.deletions[referencePos] = md[k]; else t.mismatches[referencePos] = md[k];
      ++referencePos;
    }
    offset = j;
    readMatches("MD tag should have matching bases after mismatched or missing bases");
  }
  return t;
}

// ---------------------------------------------------------------------------
// CORE/models/ObservationOperator.scala:380-475 -- operator types
// ---------------------------------------------------------------------------
enum class OpKind { Match, Insertion, Deletion, Clipped };
struct Op {
  OpKind kind;
  int size;
  bool hasRef = false;      // Match: optMismatchingBases.isDefined
  std::string ref;          // Match mismatching ref bases / Deletion basesDeleted
  bool soft = true;         // Clipped
};

static Op mkMatch(int n) { return Op{OpKind::Match, n, false, "", true}; }
static Op mkMismatch(const std::string& r) { return Op{OpKind::Match, (int)r.size(), true, r, true}; }

// CORE/models/ObservationOperator.scala:42-171 extractAlignmentOperators
static std::vector<Op> extractAlignmentOperators(const Read& read) {
  if (!read.mapped) throw Thrown("assertion failed: read.getReadMapped");      // :45
  if (!read.hasCigar || read.cigar == "*" || !read.hasMd) return {};            // :47-49
  std::vector<CigarElem> cigar = decodeCigar(read.cigar);                       // :53
  MdTag mdTag = parseMdTag(read.md, 0L);                                        // :54-56
  std::vector<Op> out;
  long idx = 0;                                                                 // :59
  for (const CigarElem& elem : cigar) {
    int length = elem.length;
    switch (elem.op) {
      case 'M': {                                                               // :63-127
        std::vector<Op> operators;   // built in order (the Scala conses then reverses)
        int refRun = 0;
        std::string ref;
        for (int opIdx = 0;; ++opIdx) {
          if (opIdx == length) {                                                // :78
            if (refRun > 0) {
              if (!ref.empty()) throw Thrown("assertion failed");
              operators.push_back(mkMatch(refRun));
            } else {
              if (ref.empty()) throw Thrown("assertion failed");               // 0-length M
              operators.push_back(mkMismatch(ref));
            }
            break;
          }
          std::optional<char> base = mdTag.mismatchedBase(idx + opIdx);         // :93
          if (!base) {                                                          // :96-103
            if (!ref.empty()) {
              operators.push_back(mkMismatch(ref));
              ref.clear();
              refRun = 1;
            } else {
              refRun += 1;
            }
          } else {                                                              // :104-115
            if (refRun > 0) {
              if (!ref.empty()) throw Thrown("assertion failed");
              operators.push_back(mkMatch(refRun));
            }
            refRun = 0;
            ref.push_back(*base);
          }
        }
        idx += length;                                                          // :124
        for (auto& o : operators) out.push_back(o);
        break;
      }
      case '=':                                                                 // :128-131
        idx += length;
        out.push_back(mkMatch(length));
        break;
      case 'X': {                                                               // :132-143
        std::string r;
        for (long i = idx; i < idx + length; ++i) {
          auto b = mdTag.mismatchedBase(i);
          if (!b) throw Thrown("Invalid CIGAR/MD tag combo: X without mismatching base");
          r.push_back(*b);
        }
        // Match(length, Some(r)) requires r.length == size (:408-409); holds here
        out.push_back(mkMismatch(r));
        idx += length;
        break;
      }
      case 'I':                                                                 // :144-146
        out.push_back(Op{OpKind::Insertion, length, false, "", true});
        break;
      case 'D': {                                                               // :147-157
        std::string r;
        for (long i = idx; i < idx + length; ++i) {
          auto b = mdTag.deletedBase(i);
          if (!b) throw Thrown("Invalid CIGAR/MD tag combo: D without deleted base");
          r.push_back(*b);
        }
        out.push_back(Op{OpKind::Deletion, (int)r.size(), false, r, true});
        idx += length;
        break;
      }
      case 'S':                                                                 // :158-160
        out.push_back(Op{OpKind::Clipped, length, false, "", true});
        break;
      case 'H':                                                                 // :161-163
        out.push_back(Op{OpKind::Clipped, length, false, "", false});
        break;
      default:                                                                  // :164-167
        throw Thrown("Unsupported CIGAR element");
    }
  }
  return out;
}

static std::string opsToString(const std::vector<Op>& ops) {
  std::string s;
  for (const Op& o : ops) {
    if (!s.empty()) s += ",";
    switch (o.kind) {
      case OpKind::Match:
        s += std::to_string(o.size) + (o.hasRef ? "X:" + o.ref : "=");
        break;
      case OpKind::Insertion: s += std::to_string(o.size) + "I"; break;
      case OpKind::Deletion: s += std::to_string(o.size) + "D:" + o.ref; break;
      case OpKind::Clipped: s += std::to_string(o.size) + (o.soft ? "S" : "H"); break;
    }
  }
  return s;
}

// ---------------------------------------------------------------------------
// CORE/genotyping/DiscoveredVariant.scala:72-108
// ---------------------------------------------------------------------------
struct DiscoveredVariant {
  std::string contigName;
  int start;
  std::string ref;
  bool hasAlt;
  std::string alt;
  int end() const { return start + (int)ref.size(); }                           // :83
  bool isNonRefModel() const { return !hasAlt; }                                // :81
  bool overlaps(const DiscoveredVariant& v) const {                             // :102-104
    return contigName == v.contigName && start < v.end() && end() > v.start;
  }
  bool overlaps(const Region& rr) const {                                       // :106-108
    return contigName == rr.name && start < rr.end && end() > rr.start;
  }
  bool operator<(const DiscoveredVariant& o) const {
    return std::tie(contigName, start, ref, hasAlt, alt) <
           std::tie(o.contigName, o.start, o.ref, o.hasAlt, o.alt);
  }
  bool operator==(const DiscoveredVariant& o) const {
    return contigName == o.contigName && start == o.start && ref == o.ref &&
           hasAlt == o.hasAlt && alt == o.alt;
  }
};

static char at(const std::string& s, long i) {   // String.apply / charAt bounds check
  if (i < 0 || i >= (long)s.size()) throw Thrown("StringIndexOutOfBoundsException");
  return s[(size_t)i];
}
static std::string substring(const std::string& s, long b, long e) {
  if (b < 0 || e > (long)s.size() || b > e) throw Thrown("StringIndexOutOfBoundsException");
  return s.substr((size_t)b, (size_t)(e - b));
}

// CORE/genotyping/DiscoverVariants.scala:112-252 variantsInRead
static std::vector<DiscoveredVariant> variantsInRead(const Read& read, int phredThreshold) {
  std::vector<DiscoveredVariant> variants;
  if (!read.mapped) return variants;                                            // :115-116
  std::vector<Op> ops;
  try {                                                                         // :119-127
    ops = extractAlignmentOperators(read);
  } catch (const Thrown&) {
    ops.clear();
  }
  int pos = (int)read.start;                                                    // :130
  int idx = 0;                                                                  // :131
  const std::string& sequence = read.seq;
  const std::string& qual = read.qual;
  const std::string& contigName = read.contigName;

  // :139-172 fastForward to the first alignment match
  size_t it = 0;
  while (it < ops.size()) {
    const Op& o = ops[it];
    bool stop = false;
    switch (o.kind) {
      case OpKind::Clipped: if (o.soft) idx += o.size; break;
      case OpKind::Insertion: idx += o.size; break;
      case OpKind::Deletion: pos += (int)o.ref.size(); break;
      case OpKind::Match: stop = true; break;
    }
    if (stop) break;
    ++it;
  }

  // :177-249 emitVariants (the Scala prepends; the order is irrelevant to callers)
  std::string lastRef = "";
  for (; it < ops.size(); ++it) {
    const Op& obs = ops[it];
    switch (obs.kind) {
      case OpKind::Clipped: break;                                              // :191-193
      case OpKind::Match: {                                                     // :194-214
        int length = obs.size;
        if (!obs.hasRef) {
          lastRef = std::string(1, at(sequence, idx + length - 1));             // :196
        } else {
          for (int i = 0; i < length; ++i) {
            if ((int)(unsigned char)at(qual, i) - 33 >= phredThreshold) {       // :199  qual(i), not qual(idx+i)
              variants.push_back(DiscoveredVariant{contigName, pos + i,
                                                   std::string(1, at(obs.ref, i)), true,
                                                   std::string(1, at(sequence, idx + i))});
            }
          }
          lastRef = std::string(1, obs.ref.back());                             // :209
        }
        pos += length;
        idx += length;
        break;
      }
      case OpKind::Insertion: {                                                 // :215-228
        int length = obs.size;
        std::string q = substring(qual, idx - 1, idx + length);                 // :216
        int sum = 0;
        for (char c : q) sum += (int)(unsigned char)c - 33;
        if (length == 0) throw Thrown("ArithmeticException: / by zero");
        int insQuals = sum / length;
        if (insQuals >= phredThreshold) {
          variants.push_back(DiscoveredVariant{contigName, pos - 1, lastRef, true,
                                               substring(sequence, idx - 1, idx + length)});
        }
        idx += length;
        break;
      }
      case OpKind::Deletion: {                                                  // :229-242
        int delLength = (int)obs.ref.size();
        if ((int)(unsigned char)at(qual, idx - 1) - 33 >= phredThreshold) {
          variants.push_back(DiscoveredVariant{contigName, pos - 1, lastRef + obs.ref, true,
                                               substring(sequence, idx - 1, idx)});
        }
        pos += delLength;
        if (obs.ref.empty()) throw Thrown("NoSuchElementException: last of empty");
        lastRef = std::string(1, obs.ref.back());
        break;
      }
    }
  }
  return variants;
}

// CORE/genotyping/DiscoverVariants.scala:71-102 variantsInRdd.  Spark SQL
// `distinct` (optMinObservations = None) or groupBy(4 keys).count().where(count > mo).
// Output order in Spark is unspecified; we return keys sorted (contig, start, ref, alt).
static std::vector<std::pair<DiscoveredVariant, long>> variantsInReads(
    const std::vector<Read>& reads, int phredThreshold, bool hasMinObs, int minObs,
    int nThreads) {
  nThreads = std::max(1, nThreads);
  std::vector<std::map<DiscoveredVariant, long>> parts((size_t)nThreads);
  auto work = [&](int t) {
    size_t lo = reads.size() * (size_t)t / (size_t)nThreads;
    size_t hi = reads.size() * (size_t)(t + 1) / (size_t)nThreads;
    for (size_t i = lo; i < hi; ++i)
      for (auto& v : variantsInRead(reads[i], phredThreshold)) parts[(size_t)t][v] += 1;
  };
  if (nThreads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nThreads; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  std::map<DiscoveredVariant, long>& counts = parts[0];
  for (size_t t = 1; t < parts.size(); ++t)
    for (auto& kv : parts[t]) counts[kv.first] += kv.second;
  std::vector<std::pair<DiscoveredVariant, long>> out;
  for (auto& kv : counts) {
    if (!hasMinObs || kv.second > (long)minObs) out.push_back(kv);              // :95  strict >
  }
  return out;
}

// ---------------------------------------------------------------------------
// CORE/util/TreeRegionJoin.scala:61-120  Forest
// ---------------------------------------------------------------------------
struct Forest {
  std::vector<std::pair<Region, int>> array;  // (key, payload index), sorted by key
  int length = 0;
  int midpoint = 1;

  explicit Forest(std::vector<std::pair<Region, int>> a) : array(std::move(a)) {
    // Forest.apply: rdd.sortByKey().collect  (:43-50); ties keep input order here
    std::stable_sort(array.begin(), array.end(),
                     [](const auto& x, const auto& y) { return x.first.compareTo(y.first) < 0; });
    length = (int)array.size();
    int i = 1;                                                                  // :66-72 pow2ceil
    while (!(2 * i >= length)) i *= 2;
    midpoint = i;
  }

  const Region& key(int i) const {
    if (i < 0 || i >= length) throw Thrown("ArrayIndexOutOfBoundsException");
    return array[(size_t)i].first;
  }

  std::optional<int> binarySearch(const Region& rr) const {                     // :74-92
    int idx = 0, step = midpoint;
    for (;;) {
      if (rr.overlaps(key(idx))) return idx;
      if (step == 0) return std::nullopt;
      int stepIdx = idx + step;
      int nextIdx;
      if (stepIdx >= length || (!rr.overlaps(key(stepIdx)) && rr.compareTo(key(stepIdx)) < 0))
        nextIdx = idx;
      else
        nextIdx = stepIdx;
      idx = nextIdx;
      step = step / 2;
    }
  }

  void expand(const Region& rr, int idx, int step, std::vector<int>& list) const {  // :94-105
    while (!(idx < 0 || idx >= length || !rr.overlaps(key(idx)))) {
      list.push_back(array[(size_t)idx].second);
      idx += step;
    }
  }

  // :111-119 get.  The Scala returns expand(idx,-1) ::: expand(idx+1,+1), each list built
  // by prepending; the resulting order is irrelevant to every caller on the path, so we
  // return ascending array order.
  std::vector<int> get(const Region& rr) const {
    std::vector<int> out;
    auto hit = binarySearch(rr);
    if (!hit) return out;
    std::vector<int> left, right;
    expand(rr, *hit, -1, left);
    expand(rr, *hit + 1, 1, right);
    std::reverse(left.begin(), left.end());
    out = left;
    out.insert(out.end(), right.begin(), right.end());
    return out;
  }
};

// ---------------------------------------------------------------------------
// CORE/genotyping/SummarizedObservation.scala:37-95
// ---------------------------------------------------------------------------
struct SummarizedObservation {
  bool isRef;
  bool forwardStrand;
  bool hasQuality;
  int quality;
  int mapQ;
  bool hasCopyNumber = false;
  int copyNumber = 0;
  bool isOther = false;
  bool isNonRef = false;

  SummarizedObservation asRef() const { return {true, forwardStrand, hasQuality, quality, mapQ}; }
  SummarizedObservation asAlt() const { return {false, forwardStrand, hasQuality, quality, mapQ}; }
  SummarizedObservation otherAlt() const {
    SummarizedObservation o{false, forwardStrand, hasQuality, quality, mapQ};
    o.isOther = true;
    return o;
  }
  SummarizedObservation nullOut() const {
    SummarizedObservation o{false, forwardStrand, hasQuality, quality, mapQ};
    o.isNonRef = true;
    return o;
  }
  SummarizedObservation addCopyNumber(int cn) const {
    SummarizedObservation o = *this;
    o.hasCopyNumber = true;
    o.copyNumber = cn;
    return o;
  }
};

struct ReadObs {   // ((ReferenceRegion, allele, sampleId), SummarizedObservation) minus sampleId
  Region rr;
  std::string allele;
  SummarizedObservation obs;
};

// CORE/genotyping/Observer.scala:48-140 observeRead
static std::vector<ReadObs> observeRead(const Read& read) {
  std::vector<Op> alignment = extractAlignmentOperators(read);                  // :51
  const std::string& contigName = read.contigName;
  const std::string& readSequence = read.seq;
  const std::string& readQualities = read.qual;
  bool forwardStrand = !read.negativeStrand;
  auto mapQ = [&]() -> int {   // Integer -> Int unboxing happens when an observation is built
    if (!read.hasMapq) throw Thrown("NullPointerException: mapq");
    return read.mapq;
  };
  std::vector<ReadObs> out;
  int readIdx = 0;                                                              // :63
  long pos = read.start;                                                        // :64
  for (const Op& op : alignment) {
    switch (op.kind) {
      case OpKind::Clipped:                                                     // :66-71
        if (op.soft) readIdx += op.size;
        break;
      case OpKind::Match:                                                       // :72-92
        for (int i = 0; i < op.size; ++i) {
          std::string allele(1, at(readSequence, readIdx));
          SummarizedObservation o{!op.hasRef, forwardStrand, true,
                                  (int)(unsigned char)at(readQualities, readIdx) - 33, mapQ()};
          out.push_back(ReadObs{Region{contigName, pos, pos + 1}, allele, o});
          readIdx += 1;
          pos += 1;
        }
        break;
      case OpKind::Insertion: {                                                 // :93-118
        int oldReadIdx = readIdx;
        readIdx += op.size;
        std::string bases = substring(readSequence, oldReadIdx, readIdx);
        std::string q = substring(readQualities, oldReadIdx, readIdx);
        int sum = 0;
        for (char c : q) sum += (int)(unsigned char)c - 33;
        if (op.size == 0) throw Thrown("ArithmeticException: / by zero");
        int qual = sum / op.size;
        SummarizedObservation o{false, forwardStrand, true, qual, mapQ()};
        out.push_back(ReadObs{Region{contigName, pos - 1, pos}, bases, o});
        break;
      }
      case OpKind::Deletion: {                                                  // :119-138
        long oldPos = pos;
        pos += op.size;
        SummarizedObservation o{false, forwardStrand, false, 0, mapQ()};
        out.push_back(ReadObs{Region{contigName, oldPos, pos}, "", o});
        break;
      }
    }
  }
  return out;
}

// ADAM 0.24.0 PhredUtils.phredToSuccessProbability (un-vendored): 1 - 10^(-q/10)
static double phredToSuccessProbability(int phred) {
  return 1.0 - std::pow(10.0, -(double)phred / 10.0);
}

// CORE/genotyping/Observer.scala:151-185 likelihoods -> (alleleArray, otherArray)
static std::pair<std::vector<double>, std::vector<double>> likelihoods(
    int copyNumber, double mapSuccessProb, bool hasBaseQuality, int baseQuality,
    int maxCopyNumber) {
  std::vector<double> alleleArray((size_t)maxCopyNumber + 1), otherArray((size_t)maxCopyNumber + 1);
  double baseSuccessProb = hasBaseQuality ? phredToSuccessProbability(baseQuality) : 1.0;
  double oneMinusEpsilon = mapSuccessProb * baseSuccessProb;                    // :167
  double epsilon = 1.0 - oneMinusEpsilon;                                       // :168
  int g = 0;
  while (g <= copyNumber) {                                                     // :171-177
    int mMinusG = copyNumber - g;
    alleleArray[(size_t)g] = std::log(mMinusG * epsilon + g * oneMinusEpsilon);
    otherArray[(size_t)g] = std::log(mMinusG * oneMinusEpsilon + g * epsilon);
    g += 1;
  }
  while (g <= maxCopyNumber) {                                                  // :178-182
    alleleArray[(size_t)g] = 0.0;
    otherArray[(size_t)g] = 0.0;
    g += 1;
  }
  return {alleleArray, otherArray};
}

// CORE/genotyping/ScoredObservation.scala:45-90 apply  (one row of the score table)
struct ScoredRow {
  int alleleForwardStrand, otherForwardStrand;
  double squareMapQ;
  std::vector<double> referenceLL, alleleLL, otherLL, nonRefLL;
  int alleleCoverage, otherCoverage, totalCoverage, copyNumber;
};

static ScoredRow scoredObservation(bool isRef, bool isOther, bool isNonRef, bool forwardStrand,
                                   bool hasQuality, int quality, int mapQ, int copyNumber,
                                   int maxCopyNumber) {
  double mapSuccessProbability = phredToSuccessProbability(mapQ);               // :53
  std::vector<double> alt =
      likelihoods(copyNumber, mapSuccessProbability, hasQuality, quality, maxCopyNumber).first;
  std::vector<double> zeros((size_t)maxCopyNumber + 1, 0.0);
  ScoredRow r;
  if (isOther) {                                                                // :60-71
    r.referenceLL = zeros; r.alleleLL = zeros; r.otherLL = alt; r.nonRefLL = zeros;
  } else if (isNonRef) {
    r.referenceLL = zeros; r.alleleLL = zeros; r.otherLL = zeros; r.nonRefLL = alt;
  } else if (isRef) {
    r.referenceLL = alt; r.alleleLL = zeros; r.otherLL = zeros; r.nonRefLL = zeros;
  } else {
    r.referenceLL = zeros; r.alleleLL = alt; r.otherLL = zeros; r.nonRefLL = zeros;
  }
  r.alleleForwardStrand = (!isRef && !isOther && forwardStrand) ? 1 : 0;        // :79
  r.otherForwardStrand = (isRef && forwardStrand) ? 1 : 0;                      // :80
  r.squareMapQ = (double)(mapQ * mapQ);                                         // :81
  r.alleleCoverage = (!isRef && !isOther) ? 1 : 0;                              // :86
  r.otherCoverage = isRef ? 1 : 0;                                              // :87
  r.totalCoverage = 1;                                                          // :88
  r.copyNumber = copyNumber;
  return r;
}

// ---------------------------------------------------------------------------
// CORE/models/CopyNumberMap.scala:75-111
// ---------------------------------------------------------------------------
struct CopyNumberMap {
  int basePloidy = 2;
  std::map<std::string, std::vector<std::pair<Region, int>>> variantsByContig;  // sorted per contig

  int minPloidy() const {
    int m = basePloidy;
    for (auto& kv : variantsByContig) for (auto& p : kv.second) m = std::min(m, p.second);
    return m;
  }
  int maxPloidy() const {
    int m = basePloidy;
    for (auto& kv : variantsByContig) for (auto& p : kv.second) m = std::max(m, p.second);
    return m;
  }
  std::vector<std::pair<Region, int>> overlappingVariants(const Region& rr) const {  // :103-111
    std::vector<std::pair<Region, int>> out;
    auto it = variantsByContig.find(rr.name);
    if (it == variantsByContig.end()) return out;
    size_t i = 0;
    const auto& v = it->second;
    while (i < v.size() && !v[i].first.overlaps(rr)) ++i;      // dropWhile
    while (i < v.size() && v[i].first.overlaps(rr)) out.push_back(v[i++]);  // takeWhile
    return out;
  }
};

static bool isDeletion(const DiscoveredVariant& v) {                            // BG:400-403
  return v.ref.size() > 1 && v.hasAlt && v.alt.size() == 1;
}
static int deletionLength(const DiscoveredVariant& v) {                         // BG:405-407
  return (int)v.ref.size() - (v.hasAlt ? (int)v.alt.size() : 0);
}
static bool isInsertion(const DiscoveredVariant& v) {                           // BG:409-412
  return v.ref.size() == 1 && v.hasAlt && v.alt.size() > 1;
}

using VarObs = std::pair<DiscoveredVariant, SummarizedObservation>;

// BG:195-393 readToObservations
static std::vector<VarObs> readToObservations(const Read& read,
                                              const std::vector<Variant>& variants,
                                              const CopyNumberMap& copyNumber,
                                              bool scoreAllSites) {
  int baseCopyNumber = copyNumber.basePloidy;                                   // :204
  // ReferenceRegion.unstranded(read) = [start, end) on the read's contig (:205-206)
  std::vector<std::pair<Region, int>> copyNumberVariants =
      copyNumber.overlappingVariants(Region{read.contigName, read.start, read.end});

  if (variants.empty() && !scoreAllSites) return {};                            // :208-209
  try {
    std::vector<ReadObs> observations = observeRead(read);                      // :214-218

    // :221-294 intersection
    std::vector<std::pair<DiscoveredVariant, std::vector<ReadObs>>> intersection;
    auto toDv = [](const Variant& v) {
      return DiscoveredVariant{v.contigName, (int)v.start, v.ref, true, v.alt};  // DiscoveredVariant(v)
    };
    auto toRegion = [](const Variant& v) { return Region{v.contigName, v.start, v.end}; };
    auto nonRefDv = [](const Region& rr) {
      return DiscoveredVariant{rr.name, (int)rr.start, "N", false, ""};         // DiscoveredVariant(rr)
    };
    if (scoreAllSites) {
      if (variants.empty()) {                                                   // :224-227
        for (auto& t : observations) intersection.push_back({nonRefDv(t.rr), {t}});
      } else {                                                                  // :228-274
        std::vector<std::vector<ReadObs>> vLists(variants.size());
        std::vector<Region> variantsByRegion;
        for (auto& v : variants) variantsByRegion.push_back(toRegion(v));
        for (auto& t : observations) {
          bool any = false;
          for (size_t i = 0; i < variantsByRegion.size(); ++i) {
            if (variantsByRegion[i].overlaps(t.rr)) { any = true; vLists[i].push_back(t); }
          }
          if (!any) intersection.push_back({nonRefDv(t.rr), {t}});
        }
        for (size_t i = 0; i < variants.size(); ++i)
          intersection.push_back({toDv(variants[i]), vLists[i]});
      }
    } else {                                                                    // :275-285
      for (auto& v : variants) {
        Region rr = toRegion(v);
        std::vector<ReadObs> obs;
        for (auto& t : observations) if (t.rr.overlaps(rr)) obs.push_back(t);
        intersection.push_back({toDv(v), obs});
      }
    }

    // :288-356 ProcessIntersections
    std::vector<VarObs> obsMap;
    for (auto& kv : intersection) {
      const DiscoveredVariant& variant = kv.first;
      const std::vector<ReadObs>& observed = kv.second;
      if (observed.empty()) {                                                   // :297-298
        continue;
      } else if (variant.isNonRefModel()) {                                     // :299-306
        for (auto& o : observed) {
          if (o.obs.isRef) obsMap.push_back({variant, o.obs.asRef()});
          else obsMap.push_back({variant, o.obs});
        }
      } else if (isInsertion(variant)) {                                        // :307-326
        std::string insAllele = variant.alt.substr(1);
        std::vector<const ReadObs*> insObserved;
        for (auto& o : observed) if (o.allele == insAllele) insObserved.push_back(&o);
        if (observed.size() == 2 && insObserved.size() == 1) {
          const ReadObs* lead = nullptr;
          for (auto& o : observed) if (o.allele != insAllele) { lead = &o; break; }
          if (!lead) throw Thrown("NoSuchElementException: head of empty");
          const std::string& leadBaseObserved = lead->allele;
          if (at(leadBaseObserved, 0) == variant.alt[0]) {                      // :314 (throws on "")
            obsMap.push_back({variant, insObserved[0]->obs});
          } else {
            obsMap.push_back({variant, insObserved[0]->obs.nullOut()});
          }
        } else {
          bool allRef = true;
          for (auto& o : observed) allRef = allRef && o.obs.isRef;
          if (allRef) obsMap.push_back({variant, observed[0].obs.asRef()});     // :322-323
          else obsMap.push_back({variant, observed[0].obs.nullOut()});          // :324-325
        }
      } else {                                                                  // :327-355
        const std::string& allele = observed[0].allele;
        const SummarizedObservation& obs = observed[0].obs;
        size_t nonEmpty = 0;
        for (auto& o : observed) if (!o.allele.empty()) ++nonEmpty;
        if (nonEmpty == 1 && allele == variant.alt) {                           // :329-330
          if (isDeletion(variant)) {
            std::vector<long> delsObserved;
            for (auto& o : observed) if (o.allele.empty()) delsObserved.push_back(o.rr.width());
            if (delsObserved.size() == 1 && observed.size() == 2 &&
                delsObserved[0] == (long)deletionLength(variant)) {
              obsMap.push_back({variant, obs.asAlt()});                         // :342
            } else {
              obsMap.push_back({variant, obs.nullOut()});                       // :344
            }
          } else {
            obsMap.push_back({variant, obs});                                   // :347
          }
        } else if (!obs.isRef || observed.size() != variant.ref.size()) {       // :349-351
          obsMap.push_back({variant, obs.nullOut()});
        } else {
          obsMap.push_back({variant, obs.asRef()});                             // :352-353
        }
      }
    }

    // :359-373 other-alt pass
    std::vector<VarObs> afterOtherAlts;
    for (auto& p : obsMap) {
      if (p.second.isNonRef) {
        bool found = false;
        for (auto& kv : obsMap) {
          if (kv.first.overlaps(p.first)) {
            const SummarizedObservation& o = kv.second;
            if (!o.isRef && !o.isOther && !o.isNonRef) { found = true; break; }
          }
        }
        if (found) afterOtherAlts.push_back({p.first, p.second.otherAlt()});
        else afterOtherAlts.push_back(p);
      } else {
        afterOtherAlts.push_back(p);
      }
    }

    // :376-383 copy number
    std::vector<VarObs> out;
    for (auto& p : afterOtherAlts) {
      int cn = baseCopyNumber;
      for (auto& c : copyNumberVariants) {
        if (p.first.overlaps(c.first)) { cn = c.second; break; }
      }
      out.push_back({p.first, p.second.addCopyNumber(cn)});
    }
    return out;
  } catch (const Thrown&) {                                                     // :385-391
    return {};
  }
}

// ---------------------------------------------------------------------------
// CORE/models/Observation.scala:72-83 -- per-site aggregate
// ---------------------------------------------------------------------------
struct Observation {
  long alleleForwardStrand = 0, otherForwardStrand = 0;
  double squareMapQ = 0.0;
  std::vector<double> referenceLL, alleleLL, otherLL, nonRefLL;
  long alleleCoverage = 0, otherCoverage = 0, totalCoverage = 0;
  bool isRef = true;
  int copyNumber = 0;
  bool seen = false;
};

static const double TEN_DIV_LOG10 = 10.0 / std::log(10.0);                      // BG:572

// BG:622-668 genotypeStateAndQuality(Array[Double])
static std::pair<int, double> genotypeStateAndQuality(const std::vector<double>& ll) {
  if (ll.size() < 2) throw Thrown("ArrayIndexOutOfBoundsException");
  int maxIdx; double max, second;
  if (ll[0] >= ll[1]) { maxIdx = 0; max = ll[0]; second = ll[1]; }              // :649-654
  else { maxIdx = 1; max = ll[1]; second = ll[0]; }
  for (size_t idx = 2; idx < ll.size(); ++idx) {                                // :625-646
    double curr = ll[idx];
    if (curr > max) { second = max; max = curr; maxIdx = (int)idx; }
    else if (curr > second) { second = curr; }
  }
  return {maxIdx, TEN_DIV_LOG10 * (max - second)};                              // :665
}

// BG:579-616 genotypeStateAndQuality(Observation) -> (alt, ref, other, qual)
static std::tuple<int, int, int, double> genotypeStateAndQualityObs(const Observation& obs) {
  int states = obs.copyNumber + 1;
  std::vector<double> score((size_t)states);
  auto blend = [&](const std::vector<double>& a1, const std::vector<double>& a2) {
    int i1 = 0, i2 = states - 1;
    while (i1 < states) { score[(size_t)i1] = a1.at((size_t)i1) + a2.at((size_t)i2); ++i1; --i2; }
  };
  blend(obs.alleleLL, obs.referenceLL);
  auto [altRefState, altRefQual] = genotypeStateAndQuality(score);
  blend(obs.alleleLL, obs.otherLL);
  auto [altOtherState, altOtherQual] = genotypeStateAndQuality(score);
  blend(obs.otherLL, obs.referenceLL);
  auto [otherAltState, otherAltQual] = genotypeStateAndQuality(score);
  if (altRefQual >= altOtherQual && altRefQual >= otherAltQual)
    return {altRefState, states - altRefState - 1, 0, altRefQual};
  else if (altOtherQual >= otherAltQual)
    return {altOtherState, 0, states - altOtherState - 1, altOtherQual};
  else
    return {0, states - otherAltState - 1, otherAltState, otherAltQual};
}

// BG:755-762 logFactorial
static double logFactorial(int n) {
  if (n < 0) throw Thrown("assertion failed: n >= 0");
  double factorial = 0.0;
  while (n > 1) { factorial = std::log((double)n) + factorial; n -= 1; }
  return factorial;
}

static const double LOG10 = std::log(10.0);
// CORE/util/LogPhred.scala:38-40
static double logErrorToPhred(double l) { return -10.0 * l / LOG10; }

// BG:778-797 fisher
static float fisher(int otherFwd, int otherRev, int alleleFwd, int alleleRev) {
  double numerator = logFactorial(otherFwd + otherRev) + logFactorial(otherFwd + alleleFwd) +
                     logFactorial(alleleFwd + alleleRev) + logFactorial(otherRev + alleleRev);
  double denominator = logFactorial(otherFwd) + logFactorial(otherRev) + logFactorial(alleleFwd) +
                       logFactorial(alleleRev) +
                       logFactorial(otherFwd + otherRev + alleleFwd + alleleRev);
  return (float)logErrorToPhred(numerator - denominator);
}

// BG:677-748 observationToGenotype -- the fields of the emitted Avro Genotype
struct GenotypeCall {
  int alt, ref, other;          // allele list = ALT x alt ++ REF x ref ++ OTHER_ALT x other
  double qual;
  int genotypeQuality;          // qual.toInt
  int sb[4];
  float fisherPValue, rmsMapQ;
  int readDepth, referenceReadDepth, alternateReadDepth;
  std::vector<double> genotypeLikelihoods, nonReferenceLikelihoods;
};

static int doubleToInt(double d) {   // Scala Double.toInt: truncate, saturate, NaN -> 0
  if (std::isnan(d)) return 0;
  if (d >= 2147483647.0) return 2147483647;
  if (d <= -2147483648.0) return (int)(-2147483647 - 1);
  return (int)d;
}

static GenotypeCall observationToGenotype(const Observation& obs) {
  GenotypeCall g;
  long coverage = obs.alleleCoverage + obs.otherCoverage;                       // :685
  auto [alt, ref, other, qual] = genotypeStateAndQualityObs(obs);               // :688
  g.alt = alt; g.ref = ref; g.other = other; g.qual = qual;
  g.sb[0] = (int)obs.otherForwardStrand;                                        // :700-703
  g.sb[1] = (int)(obs.otherCoverage - obs.otherForwardStrand);
  g.sb[2] = (int)obs.alleleForwardStrand;
  g.sb[3] = (int)(obs.alleleCoverage - obs.alleleForwardStrand);
  g.fisherPValue = fisher(g.sb[0], g.sb[1], g.sb[2], g.sb[3]);                  // :704-705
  g.rmsMapQ = (float)std::sqrt(obs.squareMapQ / (double)coverage);              // :709
  int n = obs.copyNumber + 1;                                                   // :714
  g.genotypeLikelihoods.resize((size_t)n);
  g.nonReferenceLikelihoods.resize((size_t)n);
  for (int i = 0; i < n; ++i) {                                                 // :717-729
    g.genotypeLikelihoods[(size_t)i] = obs.alleleLL.at((size_t)i) + obs.referenceLL.at((size_t)(n - 1 - i));
    g.nonReferenceLikelihoods[(size_t)i] = obs.nonRefLL.at((size_t)i) + obs.referenceLL.at((size_t)(n - 1 - i));
  }
  g.readDepth = (int)obs.totalCoverage;                                         // :740-742
  g.referenceReadDepth = (int)obs.otherCoverage;
  g.alternateReadDepth = (int)obs.alleleCoverage;
  g.genotypeQuality = doubleToInt(qual);                                        // :746
  return g;
}

// ---------------------------------------------------------------------------
// BG:427-556 readsToObservations: clamp (Spark `least`, null-skipping), inner
// join to the score table (rows exist for optQuality in {None,1..maxQ}, mapQ in
// 1..maxMapQ, copyNumber in minPloidy..maxPloidy, ScoredObservation.scala:109-112;
// a null optQuality never equals the None row under SQL equality), then
// groupBy(contig,start,ref,alt).agg(sum..., first(isRef), first(copyNumber)).
// Sums run in read order within each thread; with n_threads > 1 per-thread
// partial aggregates are merged in thread order (Spark's partial aggregation).
// ---------------------------------------------------------------------------
struct SiteAgg {
  std::map<DiscoveredVariant, Observation> m;
  void add(const VarObs& vo, int maxQuality, int maxMapQ, int minPloidy, int maxPloidy) {
    const SummarizedObservation& o = vo.second;
    int q = o.hasQuality ? std::min(maxQuality, o.quality) : maxQuality;        // BG:451 least()
    int mq = std::min(maxMapQ, o.mapQ);                                         // BG:452
    if (q < 1 || q > maxQuality) return;     // no such table row -> dropped by the inner join
    if (mq < 1 || mq > maxMapQ) return;
    if (o.copyNumber < minPloidy || o.copyNumber > maxPloidy) return;
    ScoredRow r = scoredObservation(o.isRef, o.isOther, o.isNonRef, o.forwardStrand, true, q, mq,
                                    o.copyNumber, maxPloidy);
    Observation& a = m[vo.first];
    if (!a.seen) {
      a.seen = true;
      a.referenceLL.assign((size_t)maxPloidy + 1, 0.0);
      a.alleleLL = a.otherLL = a.nonRefLL = a.referenceLL;
      a.isRef = o.isRef;                                                        // first("isRef")
      a.copyNumber = o.copyNumber;                                              // first("copyNumber")
    }
    a.alleleForwardStrand += r.alleleForwardStrand;
    a.otherForwardStrand += r.otherForwardStrand;
    a.squareMapQ += r.squareMapQ;
    for (size_t i = 0; i <= (size_t)maxPloidy; ++i) {
      a.referenceLL[i] += r.referenceLL[i];
      a.alleleLL[i] += r.alleleLL[i];
      a.otherLL[i] += r.otherLL[i];
      a.nonRefLL[i] += r.nonRefLL[i];
    }
    a.alleleCoverage += r.alleleCoverage;
    a.otherCoverage += r.otherCoverage;
    a.totalCoverage += r.totalCoverage;
  }
  void merge(const SiteAgg& o) {
    for (auto& kv : o.m) {
      Observation& a = m[kv.first];
      const Observation& b = kv.second;
      if (!a.seen) { a = b; continue; }
      a.alleleForwardStrand += b.alleleForwardStrand;
      a.otherForwardStrand += b.otherForwardStrand;
      a.squareMapQ += b.squareMapQ;
      for (size_t i = 0; i < a.referenceLL.size(); ++i) {
        a.referenceLL[i] += b.referenceLL[i];
        a.alleleLL[i] += b.alleleLL[i];
        a.otherLL[i] += b.otherLL[i];
        a.nonRefLL[i] += b.nonRefLL[i];
      }
      a.alleleCoverage += b.alleleCoverage;
      a.otherCoverage += b.otherCoverage;
      a.totalCoverage += b.totalCoverage;
    }
  }
};

// BG:88-135 call.  Returns (site, aggregate, genotype) sorted by site key.
struct CallOut {
  std::vector<DiscoveredVariant> sites;
  std::vector<Observation> obs;
  std::vector<GenotypeCall> gts;
  long joinedReads = 0;
};

static CallOut call(const std::vector<Read>& reads, const std::vector<Variant>& variants,
                    const CopyNumberMap& copyNumber, bool scoreAllSites, int maxQuality,
                    int maxMapQ, int nThreads) {
  // BG:108-114: variants keyed by ReferenceRegion(v) = [v.start, v.end); reads keyed by
  // ReferenceRegion.opt(r) (mapped reads only)
  std::vector<std::pair<Region, int>> keyed;
  for (size_t i = 0; i < variants.size(); ++i)
    keyed.push_back({Region{variants[i].contigName, variants[i].start, variants[i].end}, (int)i});
  Forest forest(keyed);
  if (forest.length == 0)
    throw std::runtime_error("ArrayIndexOutOfBoundsException: empty variant set "
                             "(TreeRegionJoin.scala:77)");
  int minP = copyNumber.minPloidy(), maxP = copyNumber.maxPloidy();
  nThreads = std::max(1, nThreads);
  std::vector<SiteAgg> parts((size_t)nThreads);
  std::vector<long> joined((size_t)nThreads, 0);
  auto work = [&](int t) {
    size_t lo = reads.size() * (size_t)t / (size_t)nThreads;
    size_t hi = reads.size() * (size_t)(t + 1) / (size_t)nThreads;
    for (size_t i = lo; i < hi; ++i) {
      const Read& r = reads[i];
      if (!r.mapped || !r.hasStart) continue;
      std::vector<int> hits = forest.get(Region{r.contigName, r.start, r.end});
      if (hits.empty()) continue;                                               // TreeRegionJoin:196-200
      joined[(size_t)t] += 1;
      std::vector<Variant> vs;
      for (int h : hits) vs.push_back(variants[(size_t)h]);
      for (auto& vo : readToObservations(r, vs, copyNumber, scoreAllSites))
        parts[(size_t)t].add(vo, maxQuality, maxMapQ, minP, maxP);
    }
  };
  if (nThreads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nThreads; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  for (size_t t = 1; t < parts.size(); ++t) parts[0].merge(parts[t]);
  CallOut out;
  for (long j : joined) out.joinedReads += j;
  for (auto& kv : parts[0].m) {
    out.sites.push_back(kv.first);
    out.obs.push_back(kv.second);
    out.gts.push_back(observationToGenotype(kv.second));
  }
  return out;
}

// ---------------------------------------------------------------------------
// CORE/util/LogUtils.scala:39-84
// ---------------------------------------------------------------------------
static double internalLogSum(double logP, double logQ) { return logP + std::log1p(std::exp(logQ - logP)); }

static double sumLogProbabilities(const std::vector<double>& p) {               // :39-55
  if (p.empty()) throw Thrown("assertion failed");
  std::vector<double> s = p;
  std::stable_sort(s.begin(), s.end(), [](double a, double b) { return a > b; });
  double acc = s[0];
  for (size_t i = 1; i < s.size(); ++i) acc = internalLogSum(acc, s[i]);
  return acc;
}

static std::vector<double> logNormalize(const std::vector<double>& p) {         // :63-66
  double sum = sumLogProbabilities(p);
  std::vector<double> out;
  for (double v : p) out.push_back(v - sum);
  return out;
}

static double logSum(double logP, double logQ) {                                // :97-104
  if (logP >= logQ && logP != 0.0) return internalLogSum(logP, logQ);
  return internalLogSum(logQ, logP);
}

static double logAdditiveInverse(double logP) {                                 // :135-137
  return 0.0 + std::log(-std::expm1(logP - 0.0));
}

// ---------------------------------------------------------------------------
// CORE/genotyping/JointAnnotatorCaller.scala:74-281 -- one site across samples.
// Alleles are coded 0 = REF, 1 = ALT, 2 = OTHER_ALT, 3 = NO_CALL.
// breeze 0.13.2 Binomial(n,p).logProbabilityOf(k) (un-vendored) =
//   lgamma(n+1) - lgamma(k+1) - lgamma(n-k+1) + k log p + (n-k) log(1-p)   (0 < p < 1)
// ---------------------------------------------------------------------------
struct JointGenotypeIn {
  std::vector<int> alleles;
  std::vector<double> genotypeLikelihoods;
};
struct JointGenotypeOut {
  std::vector<int> alleles;
  bool hasPosteriors = false, hasPriors = false, gqSet = false;
  std::vector<float> posteriors, priors;
  int genotypeQuality = 0;
};
struct JointSiteOut {
  bool emitted = false;
  double quality = 0.0;
  double maf = 0.0;
  std::vector<JointGenotypeOut> genotypes;
};

static double binomialLogProbabilityOf(int n, double p, int k) {
  if (p == 0) return k == 0 ? 0.0 : -INFINITY;
  if (p == 1) return k == n ? 0.0 : -INFINITY;
  return std::lgamma(n + 1.0) - std::lgamma(k + 1.0) - std::lgamma(n - k + 1.0) +
         k * std::log(p) + (n - k) * std::log(1 - p);
}

static JointSiteOut annotateSite(const std::vector<JointGenotypeIn>& gts) {
  JointSiteOut out;
  long called = 0, alts = 0;                                                    // :117-128
  for (auto& g : gts) for (int a : g.alleles) { if (a != 3) ++called; if (a == 1) ++alts; }
  double maf = (double)alts / (double)called;
  out.maf = maf;
  if (!(maf > 0.0)) {   // `minorAlleleFrequency <= 0.0` is false for NaN in Scala -> emitted;
    if (!std::isnan(maf)) return out;                                           // :81-82
  }
  out.emitted = true;
  std::map<int, std::vector<double>> priors;                                    // :162-176
  if (!(maf >= 1.0 || maf <= 0.0)) {
    for (auto& g : gts) {
      int cn = (int)g.alleles.size();
      if (!priors.count(cn)) {
        std::vector<double> d;
        for (int c = 0; c <= cn; ++c) d.push_back(binomialLogProbabilityOf(cn, maf, c));
        priors[cn] = d;
      }
    }
  }
  for (auto& g : gts) {                                                         // :181-262
    JointGenotypeOut o;
    o.alleles = g.alleles;
    int others = 0;
    for (int a : g.alleles) if (a == 2 || a == 3) ++others;
    if (!priors.empty() && others == 0) {
      const std::vector<double>& prior = priors.at((int)g.alleles.size());
      std::vector<double> nn(g.genotypeLikelihoods.size());
      for (size_t i = 0; i < nn.size(); ++i) nn[i] = prior.at(i) + g.genotypeLikelihoods[i];
      std::vector<double> post = logNormalize(nn);
      o.hasPosteriors = true;
      for (double d : post) o.posteriors.push_back((float)d);
      o.hasPriors = true;
      for (double d : prior) o.priors.push_back((float)d);
      auto [state, quality] = genotypeStateAndQuality(nn);
      o.alleles.clear();
      for (int i = 0; i < state; ++i) o.alleles.push_back(1);
      for (int i = 0; i < (int)g.alleles.size() - state; ++i) o.alleles.push_back(0);
      o.gqSet = true;
      o.genotypeQuality = doubleToInt(quality);
    } else if (others == 0) {
      std::vector<double> post = logNormalize(g.genotypeLikelihoods);
      o.hasPosteriors = true;
      for (double d : post) o.posteriors.push_back((float)d);
    }
    out.genotypes.push_back(o);
  }
  double s = 0.0;                                                               // :270-281
  for (auto& o : out.genotypes)
    if (o.hasPosteriors && !o.posteriors.empty()) s += (double)o.posteriors[0];
  out.quality = (-10.0 / std::log(10.0)) * s;
  return out;
}

// ---------------------------------------------------------------------------
// CORE/util/RewriteHets.scala:60-170 and CORE/util/HardFilterGenotypes.scala:176-660 -- the two
// per-genotype passes the CLI applies to call()'s output (CLI/BiallelicGenotyper.scala:279-282).
// Alleles are coded 0 = REF, 1 = ALT, 2 = OTHER_ALT, 3 = NO_CALL.  Optional fields mirror the Avro
// nullables the reference's own unit tests leave unset.
// ---------------------------------------------------------------------------
struct FilterGt {
  bool hasAlt = true;
  int refLen = 1, altLen = 1;
  std::vector<int> alleles;
  std::optional<int> gq, dp, adp;
  std::optional<float> fs, rms;
};

struct FilterArgs {          // HardFilterGenotypesArgs + RewriteHetsArgs; a filter is on iff its value > 0
  int minQuality = 30;
  float minHetSnpQualityByDepth = 2.0f, minHomSnpQualityByDepth = 1.0f;
  float minHetIndelQualityByDepth = 2.0f, minHomIndelQualityByDepth = 1.0f;
  float maxSnpPhredStrandBias = -1.0f, maxIndelPhredStrandBias = -1.0f;
  float minSnpRMSMappingQuality = 30.0f, minIndelRMSMappingQuality = -1.0f;
  int minSnpDepth = 10, maxSnpDepth = 200, minIndelDepth = 10, maxIndelDepth = 200;
  float minHetSnpAltAllelicFraction = 0.333f, maxHetSnpAltAllelicFraction = 0.666f, minHomSnpAltAllelicFraction = 0.666f;
  float minHetIndelAltAllelicFraction = 0.333f, maxHetIndelAltAllelicFraction = 0.666f, minHomIndelAltAllelicFraction = 0.666f;
  bool disableHetSnpRewriting = false, disableHetIndelRewriting = false;
};

static bool shouldRewrite(const FilterGt& g, float maxSnpAf, float maxIndelAf, bool snps, bool indels) {   // RewriteHets.scala:95-126
  if (!g.hasAlt) return false;
  const bool isSnp = g.refLen == 1 && g.altLen == 1;
  int numAlts = 0;
  for (int a : g.alleles) numAlts += a == 1;
  const bool isHet = numAlts != 0 && numAlts != (int)g.alleles.size();
  auto checkAf = [&](float af) { return g.dp && g.adp ? ((float)*g.adp / (float)*g.dp) >= af : false; };
  if (snps && isSnp && isHet) return checkAf(maxSnpAf);
  if (indels && !isSnp && isHet) return checkAf(maxIndelAf);
  return false;
}

static FilterGt rewriteGenotype(FilterGt g) {                                                             // :134-143
  for (int& a : g.alleles) a = 1;
  g.gq.reset();
  return g;
}

struct FilterResult { bool emitted = false, filtersApplied = false, passed = false; std::vector<std::string> failed; };

static FilterResult filterGenotype(const FilterGt& g, const FilterArgs& A, bool filterRefGenotypes, bool emitAll) {   // HardFilterGenotypes.scala:632-660
  FilterResult r;
  auto anyAlt = [&] { for (int a : g.alleles) if (a == 1) return true; return false; };
  const bool qualOk = !g.gq || *g.gq > A.minQuality;                                                     // :346-349
  const bool emit = emitAll || ((filterRefGenotypes ? anyAlt() : true) && qualOk);                        // :363-371
  if (!g.hasAlt) { r.emitted = true; return r; }                                                         // :656-658
  if (!emit) return r;
  r.emitted = true;
  const bool snp = g.refLen == 1 && g.altLen == 1;
  bool allAlt = true, allAltOrOther = true;
  for (int a : g.alleles) { allAlt = allAlt && a == 1; allAltOrOther = allAltOrOther && (a == 1 || a == 2); }
  std::optional<float> af;
  if (g.dp && g.adp) af = (float)*g.adp / (float)*g.dp;                                                  // :512-517
  int nFilters = 0;
  auto qd = [&](float minQd, const char* msg, bool hom) {                                                // :381-402
    if (!(minQd > 0.0f)) return;
    ++nFilters;
    if ((!allAlt && !hom) || (allAlt && hom))
      if (g.dp && g.gq && ((float)*g.gq / (float)*g.dp) < minQd) r.failed.push_back(msg);
  };
  auto mq = [&](float minMq, const char* msg) {                                                          // :488-500
    if (!(minMq > 0.0f)) return;
    ++nFilters;
    if (g.rms && *g.rms < minMq) r.failed.push_back(msg);
  };
  auto fs = [&](float maxFs, const char* msg) {                                                          // :420-432
    if (!(maxFs > 0.0f)) return;
    ++nFilters;
    if (g.fs && *g.fs > maxFs) r.failed.push_back(msg);
  };
  auto minDp = [&](int v, const char* msg) { if (v > 0) { ++nFilters; if (g.dp && *g.dp < v) r.failed.push_back(msg); } };   // :464-475
  auto maxDp = [&](int v, const char* msg) { if (v > 0) { ++nFilters; if (g.dp && *g.dp > v) r.failed.push_back(msg); } };   // :442-453
  auto minAf = [&](float v, const char* msg, bool hom) {                                                 // :531-549
    if (!(v > 0.0f)) return;
    ++nFilters;
    if (((!allAlt && !hom) || (allAlt && hom)) && af && *af <= v) r.failed.push_back(msg);
  };
  auto maxAf = [&](float v, const char* msg) {                                                           // :561-580 (het-alt counts as hom)
    if (!(v > 0.0f)) return;
    ++nFilters;
    if (!allAltOrOther && af && *af > v) r.failed.push_back(msg);
  };
  if (snp) {                                                                                             // :258-298
    qd(A.minHetSnpQualityByDepth, "HETSNPQD", false); qd(A.minHomSnpQualityByDepth, "HOMSNPQD", true);
    mq(A.minSnpRMSMappingQuality, "SNPMQ"); fs(A.maxSnpPhredStrandBias, "SNPFS");
    minDp(A.minSnpDepth, "SNPMINDP"); maxDp(A.maxSnpDepth, "SNPMAXDP");
    minAf(A.minHetSnpAltAllelicFraction, "HETSNPMINAF", false); maxAf(A.maxHetSnpAltAllelicFraction, "HETSNPMAXAF");
    minAf(A.minHomSnpAltAllelicFraction, "HOMSNPMINAF", true);
  } else {                                                                                               // :304-344
    qd(A.minHetIndelQualityByDepth, "HETINDELQD", false); qd(A.minHomIndelQualityByDepth, "HOMINDELQD", true);
    mq(A.minIndelRMSMappingQuality, "INDELMQ"); fs(A.maxIndelPhredStrandBias, "INDELFS");
    minDp(A.minIndelDepth, "INDELMINDP"); maxDp(A.maxIndelDepth, "INDELMAXDP");
    minAf(A.minHetIndelAltAllelicFraction, "HETINDELMINAF", false); maxAf(A.maxHetIndelAltAllelicFraction, "HETINDELMAXAF");
    minAf(A.minHomIndelAltAllelicFraction, "HOMINDELMINAF", true);
  }
  r.filtersApplied = nFilters > 0;                                                                       // :590-617
  r.passed = r.failed.empty();
  return r;
}

// ---------------------------------------------------------------------------
// CORE/genotyping/SquareOffReferenceModel.scala:68-245 -- cohort site extraction from per-sample
// gVCF genotypes and the squaring off of every site against every sample.
// ---------------------------------------------------------------------------
static int trimRight(const std::string& ref, const std::string& alt) {               // :95-117
  if (ref.empty() || alt.empty()) return 0;
  size_t i = ref.size(), j = alt.size();
  int trimmed = 0;
  for (;;) {
    --i; --j;
    if (ref[i] != alt[j]) return trimmed;
    if (!(i > 0 && j > 0)) return trimmed;        // one of the iterators is exhausted
    ++trimmed;
  }
}

struct SqGt {            // what squareOffSite reads of a Genotype
  std::string sample, contig, ref, alt;
  bool hasAlt = false;
  long start = 0, end = 0;
  int row = -1;          // caller's handle
};

// squareOffSite (:176-189) for one sample: index into `gts` of the genotype that represents the
// sample at the variant and whether it was excised from an overlapping genotype (its
// genotypeLikelihoods are then replaced by its nonReferenceLikelihoods, :224-244).
// `find` takes the first overlapping genotype in Spark's (unspecified) grouping order; the canonical
// order fixed here and in the kernel: lowest (start, end), a genotype with an alternate allele
// before one without.
static std::pair<int, bool> squareOffSample(const std::string& contig, long start, long end, const std::string& ref,
                                            const std::string& alt, const std::vector<SqGt>& gts) {
  for (size_t k = 0; k < gts.size(); ++k) {                                           // hasScoredVariant :198-208
    const SqGt& g = gts[k];
    if (g.start == start && g.end == end && g.contig == contig && g.ref == ref && g.hasAlt && g.alt == alt)
      return {(int)k, false};
  }
  int best = -1;
  for (size_t k = 0; k < gts.size(); ++k) {                                           // excise :224-232
    const SqGt& g = gts[k];
    if (!(g.contig == contig && g.start < end && g.end > start)) continue;
    if (best < 0) { best = (int)k; continue; }
    const SqGt& b = gts[best];
    if (std::make_tuple(g.start, g.end, g.hasAlt ? 0 : 1) < std::make_tuple(b.start, b.end, b.hasAlt ? 0 : 1)) best = (int)k;
  }
  return {best, best >= 0};
}

// ---------------------------------------------------------------------------
// CORE/genotyping/VariantSummary.scala:39-146 -- the per-site roll-up of genotype depths
// JointAnnotatorCaller.calculateAnnotations (:141-147) attaches to the variant.  -1 = None.
// ---------------------------------------------------------------------------
struct VariantSummaryO {
  int readDepth = -1, referenceReadDepth = -1, forwardReadDepth = -1, reverseReadDepth = -1,
      forwardReferenceReadDepth = -1, reverseReferenceReadDepth = -1;
};

static VariantSummaryO variantSummaryOf(int readDepth, int referenceReadDepth, const std::vector<int>& sbc) {  // :39-70
  VariantSummaryO v;
  v.readDepth = readDepth;
  v.referenceReadDepth = referenceReadDepth;
  if (!sbc.empty()) {
    if (sbc.size() != 4) throw std::invalid_argument("Genotype has an illegal number of strand bias components.");  // :49-50
    const int rfs = sbc[0], rrs = sbc[1], afs = sbc[2], ars = sbc[3];             // 0,1 ref; 2,3 alt; 0,2 fwd; 1,3 rev
    v.forwardReadDepth = rfs + afs;
    v.reverseReadDepth = rrs + ars;
    v.forwardReferenceReadDepth = rfs;
    v.reverseReferenceReadDepth = rrs;
  }
  return v;
}

static int mergeOptions(int a, int b) {                                              // :95-103
  if (a >= 0 && b >= 0) return a + b;
  return a >= 0 ? a : b;
}

static VariantSummaryO variantSummaryMerge(const VariantSummaryO& a, const VariantSummaryO& b) {   // :111-119
  VariantSummaryO v;
  v.readDepth = mergeOptions(a.readDepth, b.readDepth);
  v.referenceReadDepth = mergeOptions(a.referenceReadDepth, b.referenceReadDepth);
  v.forwardReadDepth = mergeOptions(a.forwardReadDepth, b.forwardReadDepth);
  v.reverseReadDepth = mergeOptions(a.reverseReadDepth, b.reverseReadDepth);
  v.forwardReferenceReadDepth = mergeOptions(a.forwardReferenceReadDepth, b.forwardReferenceReadDepth);
  v.reverseReferenceReadDepth = mergeOptions(a.reverseReferenceReadDepth, b.reverseReferenceReadDepth);
  return v;
}

// ---------------------------------------------------------------------------
// CORE/models/Observation.scala:22-175 -- the per-site aggregate as a record: its invariants
// (:106-112) and the two derived forms.  Neither invert nor nullOut is called on the path (the
// nullOut of BG:320-351 is SummarizedObservation's); they are here because ObservationSuite pins them.
// ---------------------------------------------------------------------------
struct ObservationRec {
  int alleleForwardStrand, otherForwardStrand;
  double squareMapQ;
  std::vector<double> referenceLogLikelihoods, alleleLogLikelihoods, otherLogLikelihoods, nonRefLogLikelihoods;
  int alleleCoverage, otherCoverage, totalCoverage;
  bool isRef;
  int copyNumber;

  void check() const {                                                               // :106-112
    const int coverage = alleleCoverage + otherCoverage;
    if (!(copyNumber <= (int)otherLogLikelihoods.size() - 1 && copyNumber <= (int)referenceLogLikelihoods.size() - 1 &&
          copyNumber > 0))
      throw std::invalid_argument("assertion failed: copy number vs likelihood lengths");
    if (!(squareMapQ >= 0.0)) throw std::invalid_argument("assertion failed: squareMapQ");
    if (!(alleleCoverage >= 0 && otherCoverage >= 0 && coverage >= 0 && totalCoverage > 0))
      throw std::invalid_argument("assertion failed: coverage");
    if (!(alleleForwardStrand >= 0 && alleleCoverage >= alleleForwardStrand && otherForwardStrand >= 0 &&
          otherCoverage >= otherForwardStrand))
      throw std::invalid_argument("assertion failed: forward strand counts");
  }
  ObservationRec invert() const {                                                    // :137-150
    ObservationRec o{otherForwardStrand, alleleForwardStrand, squareMapQ, referenceLogLikelihoods,
                     otherLogLikelihoods, alleleLogLikelihoods, nonRefLogLikelihoods, otherCoverage,
                     alleleCoverage, totalCoverage, !isRef, copyNumber};
    o.check();
    return o;
  }
  ObservationRec nullOut() const {                                                   // :158-171
    ObservationRec o{0, 0, 0.0, referenceLogLikelihoods, std::vector<double>(alleleLogLikelihoods.size(), 0.0),
                     alleleLogLikelihoods, nonRefLogLikelihoods, 0, 0, totalCoverage, false, copyNumber};
    o.check();
    return o;
  }
};

// ---------------------------------------------------------------------------
// CORE/genotyping/TrioCaller.scala:86-221 -- per-site trio consistency / phasing.
// A genotype is its allele list (0 REF, 1 ALT, 2 OTHER_ALT, 3 NO_CALL); absent = has == false.
// ---------------------------------------------------------------------------
struct TrioGt { bool has = false; std::vector<int> alleles; };
struct TrioOut {
  bool kept = false;                 // survives both filterRef passes (:92-96)
  int childAction = 0;               // 0 unchanged, 1 phased as is, 2 replaced by NO_CALL x 2 (GQ unset),
                                     // 3 phased [ALT, REF] (from first parent), 4 phased [REF, ALT]
  bool filled[3] = {false, false, false};   // sample replaced by a NO_CALL genotype (:209-216)
  std::vector<int> childAlleles;
};

static bool trioAnyAlt(const TrioGt& g) { if (!g.has) return false; for (int a : g.alleles) if (a == 1) return true; return false; }

static TrioOut trioProcess(const TrioGt& p1, const TrioGt& p2, const TrioGt& ch) {
  TrioOut o;
  if (!(trioAnyAlt(p1) || trioAnyAlt(p2) || trioAnyAlt(ch))) return o;               // filterRef :104-111
  o.childAlleles = ch.alleles;
  if (p1.has && p2.has && ch.has) {                                                   // :141-206
    auto ok = [](const TrioGt& g) {
      if (g.alleles.size() != 2) return false;
      for (int a : g.alleles) if (a == 3 || a == 2) return false;
      return true;
    };
    if (ok(p1) && ok(p2) && ok(ch)) {
      auto st = [](const TrioGt& g) { int n = 0; for (int a : g.alleles) n += a == 1; return n; };
      const int s1 = st(p1), s2 = st(p2), sc = st(ch);
      if (sc == 0) o.childAction = (s1 == 2 || s2 == 2) ? 2 : 1;
      else if (sc == 2) o.childAction = (s1 == 0 || s2 == 0) ? 2 : 1;
      else if ((s1 == 0 && s2 == 0) || (s1 == 2 && s2 == 2)) o.childAction = 2;
      else if (s1 == 1 && s2 == 1) o.childAction = 0;
      else o.childAction = s1 != 0 ? 3 : 4;
    }
  } else {
    o.filled[0] = !p1.has; o.filled[1] = !p2.has; o.filled[2] = !ch.has;
    if (!ch.has) o.childAlleles = {3, 3};
  }
  if (o.childAction == 2) o.childAlleles = {3, 3};
  else if (o.childAction == 3) o.childAlleles = {1, 0};
  else if (o.childAction == 4) o.childAlleles = {0, 1};
  bool anyAlt = trioAnyAlt(p1) || trioAnyAlt(p2);
  for (int a : o.childAlleles) anyAlt = anyAlt || a == 1;
  o.kept = anyAlt;                                                                    // second filterRef :95
  return o;
}

}  // namespace orc

// ===========================================================================
// Python binding
// ===========================================================================
using namespace orc;

// Columnar holder so that bench.py can hand over millions of reads cheaply.
struct ReadSet {
  std::vector<Read> reads;
};

static std::shared_ptr<ReadSet> makeReads(
    const std::vector<std::string>& contigNames,
    py::array_t<int32_t, py::array::c_style | py::array::forcecast> contig,
    py::array_t<int64_t, py::array::c_style | py::array::forcecast> start,
    py::array_t<int64_t, py::array::c_style | py::array::forcecast> end,
    py::array_t<int32_t, py::array::c_style | py::array::forcecast> mapq,
    py::array_t<uint8_t, py::array::c_style | py::array::forcecast> flags,
    py::bytes cigarBlob, py::array_t<int64_t, py::array::c_style | py::array::forcecast> cigarOff,
    py::bytes mdBlob, py::array_t<int64_t, py::array::c_style | py::array::forcecast> mdOff,
    py::bytes seqBlob, py::array_t<int64_t, py::array::c_style | py::array::forcecast> seqOff,
    py::bytes qualBlob, py::array_t<int64_t, py::array::c_style | py::array::forcecast> qualOff) {
  // flags: bit0 mapped, bit1 negative strand, bit2 cigar present, bit3 MD present
  auto rs = std::make_shared<ReadSet>();
  size_t n = (size_t)contig.size();
  std::string cig = cigarBlob, md = mdBlob, seq = seqBlob, qual = qualBlob;
  auto c = contig.unchecked<1>();
  auto s = start.unchecked<1>();
  auto e = end.unchecked<1>();
  auto m = mapq.unchecked<1>();
  auto f = flags.unchecked<1>();
  auto co = cigarOff.unchecked<1>();
  auto mo = mdOff.unchecked<1>();
  auto so = seqOff.unchecked<1>();
  auto qo = qualOff.unchecked<1>();
  rs->reads.resize(n);
  for (size_t i = 0; i < n; ++i) {
    Read& r = rs->reads[i];
    r.mapped = f(i) & 1;
    r.negativeStrand = f(i) & 2;
    r.hasCigar = f(i) & 4;
    r.hasMd = f(i) & 8;
    r.contigName = c(i) >= 0 ? contigNames.at((size_t)c(i)) : "";
    r.hasStart = s(i) >= 0;
    r.start = s(i);
    r.end = e(i);
    r.hasMapq = m(i) >= 0;
    r.mapq = m(i);
    r.cigar = cig.substr((size_t)co(i), (size_t)(co(i + 1) - co(i)));
    r.md = md.substr((size_t)mo(i), (size_t)(mo(i + 1) - mo(i)));
    r.seq = seq.substr((size_t)so(i), (size_t)(so(i + 1) - so(i)));
    r.qual = qual.substr((size_t)qo(i), (size_t)(qo(i + 1) - qo(i)));
  }
  return rs;
}


// Packed columnar batch (include/avocado_b200.h) -> text AlignmentRecords, so that synthetic
// batches (which are born packed) can be handed to the oracle.  CIGAR uses =/X/I/D/S/H, the MD
// tag is rebuilt from the =/X/D runs.  skip-call reads get a null mapq (what the flag stands for).
// reads [lo, hi) of a packed batch (include/avocado_b200.h layout) as text records
static void readsFromPackedRaw(const std::string& contigName, size_t lo, size_t hi, const int32_t* st, const int32_t* en,
                               const uint8_t* mq, const uint8_t* fl, const uint32_t* so, const uint32_t* ls,
                               const uint8_t* pl, const uint32_t* oo, const uint32_t* no, const uint32_t* op,
                               const uint32_t* fo, const uint8_t* fb, Read* out) {
  static const char* NIBS = "=ACMGRSVTWYHKDBN";
  auto nib = [](const uint8_t* a, uint64_t i) -> int { int b = a[i >> 1]; return (i & 1) ? (b & 15) : (b >> 4); };
  for (size_t i = lo; i < hi; ++i) {
    Read& r = out[i];
    r.mapped = true; r.contigName = contigName; r.hasStart = true;
    r.start = st[i]; r.end = en[i];
    r.negativeStrand = fl[i] & 1;
    r.hasMapq = !(fl[i] & 2); r.mapq = mq[i];
    // payload: 16-byte blocks of 10 bases (bytes 0..4, 4 bit each) + their 10 qualities (bytes 5..14)
    const uint64_t blk0 = so[i];
    auto baseAt = [&](uint64_t k) -> int {
      const uint64_t c = k / 10, w = k % 10;
      const int v = pl[(blk0 + c) * 16 + (w >> 1)];
      return (w & 1) ? (v & 15) : (v >> 4);
    };
    for (uint64_t k = 0; k < ls[i]; ++k) {
      r.seq.push_back(NIBS[baseAt(k)]);
      r.qual.push_back((char)(pl[(blk0 + k / 10) * 16 + 5 + k % 10] + 33));
    }
    uint64_t f = fo[i];   // byte index into refb
    // read index the way DiscoverVariants counts it (DiscoverVariants.scala:139-172, 189-193): soft
    // clips count before the first aligned operator and are ignored after it
    uint64_t ridx = 0;
    bool aligned = false;
    long run = 0;
    std::string md;
    bool any = false;
    for (uint32_t k = oo[i]; k < oo[i] + no[i]; ++k) {
      const uint32_t w = op[k], len = w >> 4, code = w & 15;
      any = true;
      r.cigar += std::to_string(len) + "=XIDSH"[code];
      if (code == 0) { run += len; ridx += len; aligned = true; }
      else if (code == 1) {
        for (uint32_t j = 0; j < len; ++j) {
          const int pr = fb[f++];
          // the low nibble repeats the read base discovery takes as the alternate allele: a packer /
          // generator consistency check (a base past the end of the sequence is stored as N)
          const int want = ridx + j < ls[i] ? baseAt(ridx + j) : 15;
          if ((pr & 15) != want)
            throw std::runtime_error("packed batch: mismatch pair does not repeat the read base");
          md += std::to_string(run); md.push_back(NIBS[pr >> 4]); run = 0;
        }
        ridx += len; aligned = true;
      } else if (code == 3) {
        md += std::to_string(run) + "^";
        for (uint32_t j = 0; j < len; ++j) md.push_back(NIBS[nib(fb, 2 * f + j)]);
        f += (len + 1) / 2;
        run = 0;
      } else if (code == 2 || (code == 4 && !aligned)) ridx += len;
    }
    md += std::to_string(run);
    r.hasCigar = any; r.hasMd = any;
    r.md = md;
  }
}

static std::shared_ptr<ReadSet> readsFromPacked(
    const std::string& contigName,
    py::array_t<int32_t, py::array::c_style | py::array::forcecast> start,
    py::array_t<int32_t, py::array::c_style | py::array::forcecast> end,
    py::array_t<uint8_t, py::array::c_style | py::array::forcecast> mapq,
    py::array_t<uint8_t, py::array::c_style | py::array::forcecast> flags,
    py::array_t<uint32_t, py::array::c_style | py::array::forcecast> seqOff,
    py::array_t<uint32_t, py::array::c_style | py::array::forcecast> lSeq,
    py::array_t<uint8_t, py::array::c_style | py::array::forcecast> payload,
    py::array_t<uint32_t, py::array::c_style | py::array::forcecast> opOff,
    py::array_t<uint32_t, py::array::c_style | py::array::forcecast> nOps,
    py::array_t<uint32_t, py::array::c_style | py::array::forcecast> ops,
    py::array_t<uint32_t, py::array::c_style | py::array::forcecast> refbOff,
    py::array_t<uint8_t, py::array::c_style | py::array::forcecast> refb) {
  auto rs = std::make_shared<ReadSet>();
  const size_t n = (size_t)start.size();
  rs->reads.resize(n);
  readsFromPackedRaw(contigName, 0, n, start.data(), end.data(), mapq.data(), flags.data(), seqOff.data(), lSeq.data(),
                     payload.data(), opOff.data(), nOps.data(), ops.data(), refbOff.data(), refb.data(), rs->reads.data());
  return rs;
}

// The synthetic workload of bench.py / the tests, generated HERE (same generator source as the product's
// avo_synth_host / avo_synth_device: avocado_b200/csrc/avo_synth_core.h), so that bench.py's reference
// arm and cpu_baseline leg build their input without loading libavocado_b200.so.  Reads [first, first + n)
// of the workload `params` = {seed, contig_len, first_pos, n_reads_total, read_len, max_indel, snp_rate,
// indel_rate, triallelic_frac, softclip_frac, sample} (missing keys: avo_synth_default_params' values).
static avo_synth_params synthParams(const py::dict& d) {
  avo_synth_params p;
  p.seed = 0xA70CAD0ull; p.contig_len = 1000000; p.first_pos = 0; p.n_reads_total = 200000; p.read_len = 150;
  p.max_indel = 20; p.snp_rate = 1.0e-3; p.indel_rate = 1.0e-4; p.triallelic_frac = 0.01; p.softclip_frac = 0.02;
  p.sample = 0;
  auto has = [&](const char* k) { return d.contains(k) && !d[k].is_none(); };
  if (has("seed")) p.seed = d["seed"].cast<uint64_t>();
  if (has("contig_len")) p.contig_len = d["contig_len"].cast<int32_t>();
  if (has("first_pos")) p.first_pos = d["first_pos"].cast<int32_t>();
  if (has("n_reads_total")) p.n_reads_total = d["n_reads_total"].cast<int64_t>();
  if (has("read_len")) p.read_len = d["read_len"].cast<int32_t>();
  if (has("max_indel")) p.max_indel = d["max_indel"].cast<int32_t>();
  if (has("snp_rate")) p.snp_rate = d["snp_rate"].cast<double>();
  if (has("indel_rate")) p.indel_rate = d["indel_rate"].cast<double>();
  if (has("triallelic_frac")) p.triallelic_frac = d["triallelic_frac"].cast<double>();
  if (has("softclip_frac")) p.softclip_frac = d["softclip_frac"].cast<double>();
  if (has("sample")) p.sample = d["sample"].cast<uint64_t>();
  return p;
}

struct SynthPacked {
  std::vector<int32_t> start, end;
  std::vector<uint8_t> mapq, flags, payload, refb;
  std::vector<uint32_t> seq_off, l_seq, op_off, n_ops, ops, refb_off, qhead;
};

static void synthPacked(const avo_synth_params& p, int64_t first, int64_t n, int nThreads, SynthPacked& o) {
  if (n < 0 || first < 0 || first + n > p.n_reads_total || p.read_len < 30 || p.read_len > 1000 || p.max_indel < 1 ||
      p.max_indel > 20 || (int64_t)p.contig_len < (int64_t)p.read_len + 2 * p.max_indel + 4)
    throw std::runtime_error("synth: bad parameters");
  const uint32_t lpad = avo::blocks_of((uint32_t)p.read_len);
  o.start.resize(n); o.end.resize(n); o.mapq.resize(n); o.flags.resize(n); o.seq_off.resize(n); o.l_seq.resize(n);
  o.op_off.resize(n + 1); o.n_ops.resize(n); o.refb_off.resize(n + 1); o.qhead.resize(n);
  const int T = std::max(1, std::min<int>(nThreads, (int)std::max<int64_t>(1, n / 4096)));
  auto span = [&](int t, int64_t& lo, int64_t& hi) { lo = n * t / T; hi = n * (t + 1) / T; };
  auto run = [&](auto&& fn) {
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t) pool.emplace_back(fn, t);
    fn(0);
    for (auto& th : pool) th.join();
  };
  std::vector<uint32_t> cf(n);
  run([&](int t) {
    int64_t lo, hi; span(t, lo, hi);
    for (int64_t j = lo; j < hi; ++j) {
      avo::CountSink c;
      int32_t s, e; uint8_t mq, fl;
      avo::synth_read(p, first + j, c, &s, &e, &mq, &fl);
      o.n_ops[j] = c.n_ops; cf[j] = c.n_refb;
    }
  });
  uint64_t to = 0, tf = 0;
  for (int64_t j = 0; j < n; ++j) { o.op_off[j] = (uint32_t)to; o.refb_off[j] = (uint32_t)tf; to += o.n_ops[j]; tf += cf[j]; }
  o.op_off[n] = (uint32_t)to; o.refb_off[n] = (uint32_t)tf;
  o.payload.assign((size_t)n * lpad * 16 + 16, 0); o.ops.assign(to + 4, 0); o.refb.assign(tf + 16, 0);
  run([&](int t) {
    int64_t lo, hi; span(t, lo, hi);
    for (int64_t j = lo; j < hi; ++j) {
      avo::WriteSink w;
      w.payload = o.payload.data(); w.ops = o.ops.data(); w.refb = o.refb.data();
      w.blk0 = (uint64_t)j * lpad; w.oi = o.op_off[j]; w.fi = o.refb_off[j];
      int32_t s, e; uint8_t mq, fl;
      avo::synth_read(p, first + j, w, &s, &e, &mq, &fl);
      o.start[j] = s; o.end[j] = e; o.mapq[j] = mq; o.flags[j] = fl; o.seq_off[j] = (uint32_t)(j * lpad);
      o.l_seq[j] = (uint32_t)p.read_len; o.qhead[j] = w.qhead;
    }
  });
}

static std::shared_ptr<ReadSet> synthReads(const py::dict& params, int64_t first, int64_t n, const std::string& contigName,
                                           int nThreads, int minMapq) {
  const avo_synth_params p = synthParams(params);
  auto rs = std::make_shared<ReadSet>();
  {
    py::gil_scoped_release rel;
    SynthPacked o;
    synthPacked(p, first, n, nThreads, o);
    rs->reads.resize((size_t)n);
    const int T = std::max(1, std::min<int>(nThreads, (int)std::max<int64_t>(1, n / 4096)));
    std::vector<std::thread> pool;
    auto fn = [&](int t) {
      readsFromPackedRaw(contigName, (size_t)(n * t / T), (size_t)(n * (t + 1) / T), o.start.data(), o.end.data(), o.mapq.data(),
                         o.flags.data(), o.seq_off.data(), o.l_seq.data(), o.payload.data(), o.op_off.data(), o.n_ops.data(),
                         o.ops.data(), o.refb_off.data(), o.refb.data(), rs->reads.data());
    };
    for (int t = 1; t < T; ++t) pool.emplace_back(fn, t);
    fn(0);
    for (auto& th : pool) th.join();
    // PrefilterReads.filterMappingQuality (CORE/util/PrefilterReads.scala:201-209): keep mapq > minMapq (the
    // generator's reads are all mapped, primary and not duplicates)
    if (minMapq >= 0) {
      std::vector<Read> kept;
      kept.reserve(rs->reads.size());
      for (auto& r : rs->reads)
        if (!r.hasMapq || r.mapq > minMapq) kept.push_back(std::move(r));
      rs->reads.swap(kept);
    }
  }
  return rs;
}

// the packed columns themselves (tests: byte-identical to the product's avo_synth_host)
static py::dict synthPackedPy(const py::dict& params, int64_t first, int64_t n, int nThreads) {
  SynthPacked o;
  synthPacked(synthParams(params), first, n, nThreads, o);
  auto arr = [](auto& v, size_t cnt) {
    using T = typename std::remove_reference<decltype(v)>::type::value_type;
    py::array_t<T> a(cnt);
    std::copy(v.begin(), v.begin() + cnt, a.mutable_data());
    return a;
  };
  py::dict d;
  const size_t nn = (size_t)n;
  d["start"] = arr(o.start, nn); d["end"] = arr(o.end, nn); d["mapq"] = arr(o.mapq, nn); d["flags"] = arr(o.flags, nn);
  d["seq_off"] = arr(o.seq_off, nn); d["l_seq"] = arr(o.l_seq, nn); d["op_off"] = arr(o.op_off, nn);
  d["n_ops"] = arr(o.n_ops, nn); d["refb_off"] = arr(o.refb_off, nn); d["qhead"] = arr(o.qhead, nn);
  d["payload"] = arr(o.payload, o.payload.size() - 16); d["ops"] = arr(o.ops, (size_t)o.op_off[nn]);
  d["refb"] = arr(o.refb, (size_t)o.refb_off[nn]);
  return d;
}

static Read readFromDict(const py::dict& d) {
  Read r;
  auto has = [&](const char* k) { return d.contains(k) && !d[k].is_none(); };
  r.mapped = has("readMapped") && d["readMapped"].cast<bool>();
  if (has("contigName")) r.contigName = d["contigName"].cast<std::string>();
  if (has("start")) { r.hasStart = true; r.start = d["start"].cast<long>(); }
  if (has("end")) r.end = d["end"].cast<long>();
  if (has("cigar")) { r.hasCigar = true; r.cigar = d["cigar"].cast<std::string>(); }
  if (has("mismatchingPositions")) { r.hasMd = true; r.md = d["mismatchingPositions"].cast<std::string>(); }
  if (has("sequence")) r.seq = d["sequence"].cast<std::string>();
  if (has("qual")) r.qual = d["qual"].cast<std::string>();
  if (has("mapq")) { r.hasMapq = true; r.mapq = d["mapq"].cast<int>(); }
  r.negativeStrand = has("readNegativeStrand") && d["readNegativeStrand"].cast<bool>();
  return r;
}

static std::shared_ptr<ReadSet> readsFromDicts(const std::vector<py::dict>& ds) {
  auto rs = std::make_shared<ReadSet>();
  for (auto& d : ds) rs->reads.push_back(readFromDict(d));
  return rs;
}

// variant tuple: (contigName, start, end, ref, alt-or-None)
using VarTuple = std::tuple<std::string, long, long, std::string, py::object>;
static Variant variantFromTuple(const VarTuple& t) {
  Variant v;
  v.contigName = std::get<0>(t);
  v.start = std::get<1>(t);
  v.end = std::get<2>(t);
  v.ref = std::get<3>(t);
  v.hasAlt = !std::get<4>(t).is_none();
  if (v.hasAlt) v.alt = std::get<4>(t).cast<std::string>();
  return v;
}

static CopyNumberMap makeCnm(int basePloidy,
                             const std::vector<std::tuple<std::string, long, long, int>>& cnvs) {
  // CopyNumberMap.apply: group by contig, sort each list by region (CORE/models/CopyNumberMap.scala:45-64)
  CopyNumberMap m;
  m.basePloidy = basePloidy;
  for (auto& c : cnvs)
    m.variantsByContig[std::get<0>(c)].push_back(
        {Region{std::get<0>(c), std::get<1>(c), std::get<2>(c)}, std::get<3>(c)});
  for (auto& kv : m.variantsByContig)
    std::stable_sort(kv.second.begin(), kv.second.end(),
                     [](const auto& a, const auto& b) { return a.first.compareTo(b.first) < 0; });
  return m;
}

static py::object altObj(const DiscoveredVariant& v) {
  if (!v.hasAlt) return py::none();
  return py::str(v.alt);
}

static py::dict obsToDict(const SummarizedObservation& o) {
  py::dict d;
  d["isRef"] = o.isRef;
  d["forwardStrand"] = o.forwardStrand;
  d["optQuality"] = o.hasQuality ? py::object(py::int_(o.quality)) : py::object(py::none());
  d["mapQ"] = o.mapQ;
  d["copyNumber"] = o.hasCopyNumber ? py::object(py::int_(o.copyNumber)) : py::object(py::none());
  d["isOther"] = o.isOther;
  d["isNonRef"] = o.isNonRef;
  return d;
}

PYBIND11_MODULE(avocado_oracle, m) {
  m.doc() = "CPU oracle (test infrastructure) restating avocado's genotyping hot path";

  py::class_<ReadSet, std::shared_ptr<ReadSet>>(m, "ReadSet")
      .def("__len__", [](const ReadSet& r) { return r.reads.size(); });
  m.def("make_reads", &makeReads);
  m.def("reads_from_dicts", &readsFromDicts);
  m.def("reads_from_packed", &readsFromPacked);
  m.def("synth_reads", &synthReads, py::arg("params"), py::arg("first"), py::arg("n"), py::arg("contig_name") = "ctg",
        py::arg("n_threads") = 1, py::arg("min_mapq") = -1);
  m.def("synth_packed", &synthPackedPy, py::arg("params"), py::arg("first"), py::arg("n"), py::arg("n_threads") = 1);
  m.def("read_as_dict", [](std::shared_ptr<ReadSet> rs, size_t i) {
    const Read& r = rs->reads.at(i);
    py::dict d;
    d["readMapped"] = r.mapped; d["contigName"] = r.contigName; d["start"] = r.start; d["end"] = r.end;
    d["cigar"] = r.hasCigar ? py::object(py::str(r.cigar)) : py::object(py::none());
    d["mismatchingPositions"] = r.hasMd ? py::object(py::str(r.md)) : py::object(py::none());
    d["sequence"] = r.seq; d["qual"] = r.qual;
    d["mapq"] = r.hasMapq ? py::object(py::int_(r.mapq)) : py::object(py::none());
    d["readNegativeStrand"] = r.negativeStrand;
    return d;
  });

  m.def("extract_alignment_operators", [](const py::dict& d) {
    return opsToString(extractAlignmentOperators(readFromDict(d)));
  });

  m.def("variants_in_read", [](const py::dict& d, int thr) {
    py::list out;
    for (auto& v : variantsInRead(readFromDict(d), thr))
      out.append(py::make_tuple(v.contigName, v.start, v.ref, altObj(v)));
    return out;
  });

  m.def("discover",
        [](std::shared_ptr<ReadSet> rs, int phredThreshold, py::object minObs, int nThreads) {
          std::vector<std::pair<DiscoveredVariant, long>> res;
          bool has = !minObs.is_none();
          int mo = has ? minObs.cast<int>() : 0;
          {
            py::gil_scoped_release rel;
            res = variantsInReads(rs->reads, phredThreshold, has, mo, nThreads);
          }
          py::list out;
          for (auto& kv : res)
            out.append(py::make_tuple(kv.first.contigName, kv.first.start, kv.first.end(),
                                      kv.first.ref, altObj(kv.first), kv.second));
          return out;
        },
        py::arg("reads"), py::arg("phred_threshold") = 0, py::arg("min_observations") = py::none(),
        py::arg("n_threads") = 1);

  m.def("forest_get", [](const std::vector<std::tuple<std::string, long, long>>& regions,
                         const std::tuple<std::string, long, long>& q) {
    std::vector<std::pair<Region, int>> a;
    for (size_t i = 0; i < regions.size(); ++i)
      a.push_back({Region{std::get<0>(regions[i]), std::get<1>(regions[i]), std::get<2>(regions[i])}, (int)i});
    Forest f(a);
    return f.get(Region{std::get<0>(q), std::get<1>(q), std::get<2>(q)});
  });

  m.def("observe_read", [](const py::dict& d) {
    py::list out;
    for (auto& o : observeRead(readFromDict(d)))
      out.append(py::make_tuple(o.rr.name, o.rr.start, o.rr.end, o.allele, obsToDict(o.obs)));
    return out;
  });

  m.def("likelihoods", [](int cn, double mapSuccess, py::object q, py::object maxCn) {
    bool has = !q.is_none();
    int mc = maxCn.is_none() ? cn : maxCn.cast<int>();
    return likelihoods(cn, mapSuccess, has, has ? q.cast<int>() : 0, mc);
  }, py::arg("copy_number"), py::arg("map_success_prob"), py::arg("base_quality"),
     py::arg("max_copy_number") = py::none());

  m.def("scored_observation", [](bool isRef, bool isOther, bool isNonRef, bool fwd, py::object q,
                                 int mapQ, int cn, int maxCn) {
    bool has = !q.is_none();
    ScoredRow r = scoredObservation(isRef, isOther, isNonRef, fwd, has, has ? q.cast<int>() : 0,
                                    mapQ, cn, maxCn);
    py::dict d;
    d["alleleForwardStrand"] = r.alleleForwardStrand;
    d["otherForwardStrand"] = r.otherForwardStrand;
    d["squareMapQ"] = r.squareMapQ;
    d["referenceLogLikelihoods"] = r.referenceLL;
    d["alleleLogLikelihoods"] = r.alleleLL;
    d["otherLogLikelihoods"] = r.otherLL;
    d["nonRefLogLikelihoods"] = r.nonRefLL;
    d["alleleCoverage"] = r.alleleCoverage;
    d["otherCoverage"] = r.otherCoverage;
    d["totalCoverage"] = r.totalCoverage;
    d["copyNumber"] = r.copyNumber;
    return d;
  });

  m.def("read_to_observations",
        [](const py::dict& read, const std::vector<VarTuple>& variants, int basePloidy,
           const std::vector<std::tuple<std::string, long, long, int>>& cnvs, bool scoreAllSites) {
          std::vector<Variant> vs;
          for (auto& t : variants) vs.push_back(variantFromTuple(t));
          py::list out;
          for (auto& vo : readToObservations(readFromDict(read), vs, makeCnm(basePloidy, cnvs),
                                             scoreAllSites))
            out.append(py::make_tuple(
                py::make_tuple(vo.first.contigName, vo.first.start, vo.first.ref, altObj(vo.first)),
                obsToDict(vo.second)));
          return out;
        });

  m.def("genotype_state_and_quality", [](const std::vector<double>& ll) {
    return genotypeStateAndQuality(ll);
  });

  auto obsFromArgs = [](int alleleFwd, int otherFwd, double sq, std::vector<double> refLL,
                        std::vector<double> alleleLL, std::vector<double> otherLL,
                        std::vector<double> nonRefLL, int alleleCov, int otherCov, int totalCov,
                        bool isRef, py::object cn) {
    Observation o;
    o.alleleForwardStrand = alleleFwd; o.otherForwardStrand = otherFwd; o.squareMapQ = sq;
    o.referenceLL = refLL; o.alleleLL = alleleLL; o.otherLL = otherLL; o.nonRefLL = nonRefLL;
    o.alleleCoverage = alleleCov; o.otherCoverage = otherCov; o.totalCoverage = totalCov;
    o.isRef = isRef;
    o.copyNumber = cn.is_none() ? (int)alleleLL.size() - 1 : cn.cast<int>();   // Observation.scala:44-57
    o.seen = true;
    return o;
  };

  auto gtToDict = [](const GenotypeCall& g) {
    py::dict d;
    d["alt"] = g.alt; d["ref"] = g.ref; d["other"] = g.other; d["qual"] = g.qual;
    d["genotypeQuality"] = g.genotypeQuality;
    d["strandBiasComponents"] = std::vector<int>(g.sb, g.sb + 4);
    d["fisherStrandBiasPValue"] = g.fisherPValue;
    d["rmsMapQ"] = g.rmsMapQ;
    d["readDepth"] = g.readDepth;
    d["referenceReadDepth"] = g.referenceReadDepth;
    d["alternateReadDepth"] = g.alternateReadDepth;
    d["genotypeLikelihoods"] = g.genotypeLikelihoods;
    d["nonReferenceLikelihoods"] = g.nonReferenceLikelihoods;
    return d;
  };

  m.def("observation_to_genotype",
        [obsFromArgs, gtToDict](int alleleFwd, int otherFwd, double sq, std::vector<double> refLL,
                                std::vector<double> alleleLL, std::vector<double> otherLL,
                                std::vector<double> nonRefLL, int alleleCov, int otherCov,
                                int totalCov, bool isRef, py::object cn) {
          return gtToDict(observationToGenotype(obsFromArgs(alleleFwd, otherFwd, sq, refLL, alleleLL,
                                                            otherLL, nonRefLL, alleleCov, otherCov,
                                                            totalCov, isRef, cn)));
        },
        py::arg("alleleForwardStrand"), py::arg("otherForwardStrand"), py::arg("squareMapQ"),
        py::arg("referenceLogLikelihoods"), py::arg("alleleLogLikelihoods"),
        py::arg("otherLogLikelihoods"), py::arg("nonRefLogLikelihoods"), py::arg("alleleCoverage"),
        py::arg("otherCoverage"), py::arg("totalCoverage") = 1, py::arg("isRef") = true,
        py::arg("copyNumber") = py::none());

  m.def("log_factorial", &logFactorial);
  m.def("fisher", &fisher);
  m.def("log_error_to_phred", &logErrorToPhred);
  m.def("sum_log_probabilities", &sumLogProbabilities);
  m.def("log_normalize", &logNormalize);
  m.def("log_sum", &logSum);
  m.def("log_additive_inverse", &logAdditiveInverse);
  m.def("phred_to_success_probability", &phredToSuccessProbability);

  // call(): returns a dict of parallel lists / numpy arrays, one entry per emitted site,
  // sorted by (contigName, start, ref, alt).
  m.def("call",
        [](std::shared_ptr<ReadSet> rs, const std::vector<VarTuple>& variants, int basePloidy,
           const std::vector<std::tuple<std::string, long, long, int>>& cnvs, bool scoreAllSites,
           int maxQuality, int maxMapQ, int nThreads) {
          std::vector<Variant> vs;
          for (auto& t : variants) vs.push_back(variantFromTuple(t));
          CopyNumberMap cnm = makeCnm(basePloidy, cnvs);
          CallOut co;
          {
            py::gil_scoped_release rel;
            co = call(rs->reads, vs, cnm, scoreAllSites, maxQuality, maxMapQ, nThreads);
          }
          size_t n = co.sites.size();
          size_t w = (size_t)cnm.maxPloidy() + 1;
          py::list contig, ref, alt;
          py::array_t<int32_t> start((py::ssize_t)n), alleleFwd((py::ssize_t)n), otherFwd((py::ssize_t)n),
              alleleCov((py::ssize_t)n), otherCov((py::ssize_t)n), totalCov((py::ssize_t)n),
              copyNumber((py::ssize_t)n), nAlt((py::ssize_t)n), nRef((py::ssize_t)n),
              nOther((py::ssize_t)n), gq((py::ssize_t)n);
          py::array_t<uint8_t> isRef((py::ssize_t)n);
          py::array_t<double> sq((py::ssize_t)n), qual((py::ssize_t)n);
          py::array_t<float> fisherP((py::ssize_t)n), rms((py::ssize_t)n);
          py::array_t<double> refLL({n, w}), alleleLL({n, w}), otherLL({n, w}), nonRefLL({n, w}),
              gl({n, w}), nrl({n, w});
          py::array_t<int32_t> sb({n, (size_t)4});
          auto rl = refLL.mutable_unchecked<2>(); auto al = alleleLL.mutable_unchecked<2>();
          auto ol = otherLL.mutable_unchecked<2>(); auto nl = nonRefLL.mutable_unchecked<2>();
          auto glm = gl.mutable_unchecked<2>(); auto nrm = nrl.mutable_unchecked<2>();
          auto sbm = sb.mutable_unchecked<2>();
          for (size_t i = 0; i < n; ++i) {
            const DiscoveredVariant& v = co.sites[i];
            const Observation& o = co.obs[i];
            const GenotypeCall& g = co.gts[i];
            contig.append(v.contigName);
            ref.append(v.ref);
            alt.append(altObj(v));
            start.mutable_at(i) = v.start;
            alleleFwd.mutable_at(i) = (int32_t)o.alleleForwardStrand;
            otherFwd.mutable_at(i) = (int32_t)o.otherForwardStrand;
            alleleCov.mutable_at(i) = (int32_t)o.alleleCoverage;
            otherCov.mutable_at(i) = (int32_t)o.otherCoverage;
            totalCov.mutable_at(i) = (int32_t)o.totalCoverage;
            copyNumber.mutable_at(i) = o.copyNumber;
            isRef.mutable_at(i) = o.isRef;
            sq.mutable_at(i) = o.squareMapQ;
            nAlt.mutable_at(i) = g.alt; nRef.mutable_at(i) = g.ref; nOther.mutable_at(i) = g.other;
            qual.mutable_at(i) = g.qual; gq.mutable_at(i) = g.genotypeQuality;
            fisherP.mutable_at(i) = g.fisherPValue; rms.mutable_at(i) = g.rmsMapQ;
            for (size_t k = 0; k < w; ++k) {
              rl(i, k) = o.referenceLL[k]; al(i, k) = o.alleleLL[k];
              ol(i, k) = o.otherLL[k]; nl(i, k) = o.nonRefLL[k];
              glm(i, k) = k < g.genotypeLikelihoods.size() ? g.genotypeLikelihoods[k] : 0.0;
              nrm(i, k) = k < g.nonReferenceLikelihoods.size() ? g.nonReferenceLikelihoods[k] : 0.0;
            }
            for (int k = 0; k < 4; ++k) sbm(i, k) = g.sb[k];
          }
          py::dict d;
          d["contigName"] = contig; d["start"] = start; d["referenceAllele"] = ref;
          d["alternateAllele"] = alt;
          d["alleleForwardStrand"] = alleleFwd; d["otherForwardStrand"] = otherFwd;
          d["squareMapQ"] = sq; d["referenceLogLikelihoods"] = refLL;
          d["alleleLogLikelihoods"] = alleleLL; d["otherLogLikelihoods"] = otherLL;
          d["nonRefLogLikelihoods"] = nonRefLL; d["alleleCoverage"] = alleleCov;
          d["otherCoverage"] = otherCov; d["totalCoverage"] = totalCov; d["isRef"] = isRef;
          d["copyNumber"] = copyNumber; d["nAlt"] = nAlt; d["nRef"] = nRef; d["nOther"] = nOther;
          d["qual"] = qual; d["genotypeQuality"] = gq; d["strandBiasComponents"] = sb;
          d["fisherStrandBiasPValue"] = fisherP; d["rmsMapQ"] = rms;
          d["genotypeLikelihoods"] = gl; d["nonReferenceLikelihoods"] = nrl;
          d["joinedReads"] = co.joinedReads;
          return d;
        },
        py::arg("reads"), py::arg("variants"), py::arg("base_ploidy") = 2,
        py::arg("cnvs") = std::vector<std::tuple<std::string, long, long, int>>{},
        py::arg("score_all_sites") = false, py::arg("max_quality") = 93, py::arg("max_mapq") = 93,
        py::arg("n_threads") = 1);

  // annotate_site: genotypes = list of (alleles[int], genotypeLikelihoods[double])
  m.def("annotate_site", [](const std::vector<std::pair<std::vector<int>, std::vector<double>>>& gts) {
    std::vector<JointGenotypeIn> in;
    for (auto& g : gts) in.push_back({g.first, g.second});
    JointSiteOut o = annotateSite(in);
    py::dict d;
    d["emitted"] = o.emitted;
    d["quality"] = o.quality;
    d["maf"] = o.maf;
    py::list gl;
    for (auto& g : o.genotypes) {
      py::dict e;
      e["alleles"] = g.alleles;
      e["posteriors"] = g.hasPosteriors ? py::object(py::cast(g.posteriors)) : py::object(py::none());
      e["priors"] = g.hasPriors ? py::object(py::cast(g.priors)) : py::object(py::none());
      e["genotypeQuality"] = g.gqSet ? py::object(py::int_(g.genotypeQuality)) : py::object(py::none());
      gl.append(e);
    }
    d["genotypes"] = gl;
    return d;
  });
  // ---- RewriteHets / HardFilterGenotypes ----------------------------------------------------
  auto gtFromDict = [](const py::dict& d) {
    FilterGt g;
    auto has = [&](const char* k) { return d.contains(k) && !d[k].is_none(); };
    g.hasAlt = has("alt");
    if (has("ref")) g.refLen = (int)d["ref"].cast<std::string>().size();
    if (g.hasAlt) g.altLen = (int)d["alt"].cast<std::string>().size();
    g.alleles = d["alleles"].cast<std::vector<int>>();
    if (has("gq")) g.gq = d["gq"].cast<int>();
    if (has("dp")) g.dp = d["dp"].cast<int>();
    if (has("adp")) g.adp = d["adp"].cast<int>();
    if (has("fs")) g.fs = d["fs"].cast<float>();
    if (has("rms")) g.rms = d["rms"].cast<float>();
    return g;
  };
  auto argsFromDict = [](const py::dict& d) {
    FilterArgs a;
#define AVO_ARG(name, type) if (d.contains(#name)) a.name = d[#name].cast<type>();
    AVO_ARG(minQuality, int) AVO_ARG(minHetSnpQualityByDepth, float) AVO_ARG(minHomSnpQualityByDepth, float)
    AVO_ARG(minHetIndelQualityByDepth, float) AVO_ARG(minHomIndelQualityByDepth, float)
    AVO_ARG(maxSnpPhredStrandBias, float) AVO_ARG(maxIndelPhredStrandBias, float)
    AVO_ARG(minSnpRMSMappingQuality, float) AVO_ARG(minIndelRMSMappingQuality, float)
    AVO_ARG(minSnpDepth, int) AVO_ARG(maxSnpDepth, int) AVO_ARG(minIndelDepth, int) AVO_ARG(maxIndelDepth, int)
    AVO_ARG(minHetSnpAltAllelicFraction, float) AVO_ARG(maxHetSnpAltAllelicFraction, float)
    AVO_ARG(minHomSnpAltAllelicFraction, float) AVO_ARG(minHetIndelAltAllelicFraction, float)
    AVO_ARG(maxHetIndelAltAllelicFraction, float) AVO_ARG(minHomIndelAltAllelicFraction, float)
    AVO_ARG(disableHetSnpRewriting, bool) AVO_ARG(disableHetIndelRewriting, bool)
#undef AVO_ARG
    return a;
  };
  m.def("should_rewrite", [gtFromDict](const py::dict& g, float maxSnpAf, float maxIndelAf, bool snps, bool indels) {
    return shouldRewrite(gtFromDict(g), maxSnpAf, maxIndelAf, snps, indels);
  });
  // rewrite_and_filter: RewriteHets(args) then HardFilterGenotypes(args, filterRefGenotypes, emitAllGenotypes)
  m.def("rewrite_and_filter", [gtFromDict, argsFromDict](const py::dict& gd, const py::dict& ad, bool filterRef,
                                                         bool emitAll, bool rewrite) {
    FilterGt g = gtFromDict(gd);
    const FilterArgs a = argsFromDict(ad);
    bool rewritten = false;
    if (rewrite && shouldRewrite(g, a.maxHetSnpAltAllelicFraction, a.maxHetIndelAltAllelicFraction,
                                 !a.disableHetSnpRewriting, !a.disableHetIndelRewriting)) {
      g = rewriteGenotype(g);
      rewritten = true;
    }
    const FilterResult r = filterGenotype(g, a, filterRef, emitAll);
    py::dict d;
    d["rewritten"] = rewritten;
    d["alleles"] = g.alleles;
    d["gq"] = g.gq ? py::object(py::int_(*g.gq)) : py::object(py::none());
    d["emitted"] = r.emitted; d["filtersApplied"] = r.filtersApplied; d["filtersPassed"] = r.passed;
    d["filtersFailed"] = r.failed;
    return d;
  });
  // ---- SquareOffReferenceModel ---------------------------------------------------------------
  m.def("trim_right", &trimRight);
  // square_off_sample(variant (contig, start, end, ref, alt), genotypes [(contig, start, end, ref, alt|None)])
  //   -> (index of the genotype used or -1, excised?)
  m.def("square_off_sample", [](const std::tuple<std::string, long, long, std::string, std::string>& v,
                                const std::vector<std::tuple<std::string, long, long, std::string, py::object>>& gts) {
    std::vector<SqGt> g;
    for (auto& t : gts) {
      SqGt x;
      x.contig = std::get<0>(t); x.start = std::get<1>(t); x.end = std::get<2>(t); x.ref = std::get<3>(t);
      x.hasAlt = !std::get<4>(t).is_none();
      if (x.hasAlt) x.alt = std::get<4>(t).cast<std::string>();
      g.push_back(x);
    }
    return squareOffSample(std::get<0>(v), std::get<1>(v), std::get<2>(v), std::get<3>(v), std::get<4>(v), g);
  });
  // ---- TrioCaller ---------------------------------------------------------------------------------
  // trio_process(first parent alleles | None, second parent | None, child | None)
  m.def("trio_process", [](py::object a, py::object b, py::object c) {
    auto mk = [](py::object o) { TrioGt g; if (!o.is_none()) { g.has = true; g.alleles = o.cast<std::vector<int>>(); } return g; };
    const TrioOut r = trioProcess(mk(a), mk(b), mk(c));
    py::dict d;
    d["kept"] = r.kept; d["childAction"] = r.childAction; d["childAlleles"] = r.childAlleles;
    d["filled"] = std::vector<bool>{r.filled[0], r.filled[1], r.filled[2]};
    return d;
  });
  // ---- VariantSummary ------------------------------------------------------------------------------
  // variant_summary([(readDepth | None, referenceReadDepth | None, strandBiasComponents) | 6-tuple of
  // Optional ints]) -> the merged summary (calculateAnnotations: map(VariantSummary(_)).reduce(_.merge(_)))
  m.def("variant_summary", [](const std::vector<py::object>& items) {
    auto opt = [](py::handle h) { return h.is_none() ? -1 : h.cast<int>(); };
    bool first = true;
    VariantSummaryO acc;
    for (const py::object& it : items) {
      py::tuple t = it.cast<py::tuple>();
      VariantSummaryO v;
      if (t.size() == 3) {
        v = variantSummaryOf(opt(t[0]), opt(t[1]), t[2].cast<std::vector<int>>());
      } else {
        v.readDepth = opt(t[0]); v.referenceReadDepth = opt(t[1]); v.forwardReadDepth = opt(t[2]);
        v.reverseReadDepth = opt(t[3]); v.forwardReferenceReadDepth = opt(t[4]); v.reverseReferenceReadDepth = opt(t[5]);
      }
      acc = first ? v : variantSummaryMerge(acc, v);
      first = false;
    }
    auto out = [](int x) { return x < 0 ? py::object(py::none()) : py::object(py::int_(x)); };
    py::dict d;
    d["readDepth"] = out(acc.readDepth); d["referenceReadDepth"] = out(acc.referenceReadDepth);
    d["forwardReadDepth"] = out(acc.forwardReadDepth); d["reverseReadDepth"] = out(acc.reverseReadDepth);
    d["forwardReferenceReadDepth"] = out(acc.forwardReferenceReadDepth);
    d["reverseReferenceReadDepth"] = out(acc.reverseReferenceReadDepth);
    return d;
  });
  // ---- CopyNumberMap -------------------------------------------------------------------------------
  // copy_number_map(basePloidy, [(contig, start, end, featureType)], query (contig, start, end) | None):
  // CopyNumberMap.apply (CORE/models/CopyNumberMap.scala:45-64: DUP -> base + 1, DEL -> base - 1, others
  // dropped; per contig sorted by region), min / maxPloidy (:81-95), overlappingVariants (:103-111)
  m.def("copy_number_map", [](int basePloidy, const std::vector<std::tuple<std::string, long, long, std::string>>& features,
                              py::object query) {
    std::vector<std::tuple<std::string, long, long, int>> cnvs;
    for (auto& f : features) {
      if (std::get<3>(f) == "DUP") cnvs.emplace_back(std::get<0>(f), std::get<1>(f), std::get<2>(f), basePloidy + 1);
      else if (std::get<3>(f) == "DEL") cnvs.emplace_back(std::get<0>(f), std::get<1>(f), std::get<2>(f), basePloidy - 1);
    }
    const CopyNumberMap cm = makeCnm(basePloidy, cnvs);
    py::dict d, by;
    d["basePloidy"] = cm.basePloidy; d["minPloidy"] = cm.minPloidy(); d["maxPloidy"] = cm.maxPloidy();
    for (auto& kv : cm.variantsByContig) {
      py::list l;
      for (auto& p : kv.second) l.append(py::make_tuple(p.first.start, p.first.end, p.second));
      by[py::str(kv.first)] = l;
    }
    d["variantsByContig"] = by;
    if (!query.is_none()) {
      auto q = query.cast<std::tuple<std::string, long, long>>();
      py::list l;
      for (auto& p : cm.overlappingVariants(Region{std::get<0>(q), std::get<1>(q), std::get<2>(q)}))
        l.append(py::make_tuple(p.first.start, p.first.end, p.second));
      d["overlapping"] = l;
    }
    return d;
  }, py::arg("basePloidy"), py::arg("features"), py::arg("query") = py::none());
  // ---- Observation (record) ------------------------------------------------------------------------
  // observation(afs, ofs, squareMapQ, ref, allele, other, nonRef, alleleCov, otherCov, totalCoverage = 1,
  //             isRef = true, op = "check" | "invert" | "nullOut"): the companion apply of
  // CORE/models/Observation.scala:22-46 (copyNumber = len(allele) - 1), the invariants and derived forms
  m.def("observation", [](int afs, int ofs, double sq, std::vector<double> ref, std::vector<double> allele,
                          std::vector<double> other, std::vector<double> nonref, int acov, int ocov, int tcov,
                          bool isRef, const std::string& op) {
    ObservationRec o{afs, ofs, sq, ref, allele, other, nonref, acov, ocov, tcov, isRef, (int)allele.size() - 1};
    o.check();
    if (op == "invert") o = o.invert();
    else if (op == "nullOut") o = o.nullOut();
    py::dict d;
    d["alleleForwardStrand"] = o.alleleForwardStrand; d["otherForwardStrand"] = o.otherForwardStrand;
    d["squareMapQ"] = o.squareMapQ; d["referenceLogLikelihoods"] = o.referenceLogLikelihoods;
    d["alleleLogLikelihoods"] = o.alleleLogLikelihoods; d["otherLogLikelihoods"] = o.otherLogLikelihoods;
    d["nonRefLogLikelihoods"] = o.nonRefLogLikelihoods; d["alleleCoverage"] = o.alleleCoverage;
    d["otherCoverage"] = o.otherCoverage; d["totalCoverage"] = o.totalCoverage; d["isRef"] = o.isRef;
    d["copyNumber"] = o.copyNumber;
    return d;
  }, py::arg("alleleForwardStrand"), py::arg("otherForwardStrand"), py::arg("squareMapQ"),
     py::arg("referenceLogLikelihoods"), py::arg("alleleLogLikelihoods"), py::arg("otherLogLikelihoods"),
     py::arg("nonRefLogLikelihoods"), py::arg("alleleCoverage"), py::arg("otherCoverage"),
     py::arg("totalCoverage") = 1, py::arg("isRef") = true, py::arg("op") = "check");
}
