"""Shared test helpers: golden fixture loading and oracle glue (test infrastructure)."""
import gzip
import json
import os

import numpy as np

from avocado_b200.records import AlignmentRecord

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name):
    """tests/golden/sam/<name>.json.gz -> (list[AlignmentRecord], contigs, samples)"""
    with gzip.open(os.path.join(GOLDEN, "sam", name + ".json.gz"), "rb") as fh:
        doc = json.loads(fh.read().decode())
    cols = doc["columns"]
    recs = []
    for i in range(doc["n"]):
        recs.append(AlignmentRecord(**{k: cols[k][i] for k in cols}))
    return recs, doc["contigs"], doc["samples"]


def oracle_reads(oracle, recs):
    return oracle.reads_from_dicts([r.as_dict() if hasattr(r, "as_dict") else r for r in recs])


def oracle_discover(oracle, recs, phred=None, min_obs=None):
    """DiscoverVariants(reads, optPhredThreshold, optMinObservations) -> list of variant tuples
    (contig, start, end, ref, alt)."""
    rs = oracle_reads(oracle, recs)
    out = oracle.discover(rs, 0 if phred is None else phred, min_obs)
    return [(c, s, e, r, a) for (c, s, e, r, a, _n) in out]


def oracle_call(oracle, recs, variants, ploidy=2, cnvs=(), score_all=False, max_q=93, max_mq=93):
    rs = oracle_reads(oracle, recs)
    return oracle.call(rs, list(variants), ploidy, list(cnvs), score_all, max_q, max_mq, 1)


def oracle_discover_and_call(oracle, recs, ploidy=2, score_all=False, phred=None, min_obs=None):
    variants = oracle_discover(oracle, recs, phred, min_obs)
    return oracle_call(oracle, recs, variants, ploidy=ploidy, score_all=score_all)


def alleles_of(res, i):
    """ALT x nAlt ++ REF x nRef ++ OTHER_ALT x nOther (BG:691-697)"""
    return (["ALT"] * int(res["nAlt"][i]) + ["REF"] * int(res["nRef"][i]) +
            ["OTHER_ALT"] * int(res["nOther"][i]))


def rows(res):
    """Iterate a call() result dict as per-site dicts."""
    n = len(res["start"])
    out = []
    for i in range(n):
        d = {k: (v[i] if not np.isscalar(v) else v) for k, v in res.items() if k != "joinedReads"}
        d["alleles"] = alleles_of(res, i)
        out.append(d)
    return out


# ---- GPU parity glue ----------------------------------------------------------------------------
def oracle_reads_from_batch(oracle, batch, contig="ctg"):
    b = batch.to_host()
    m = b.meta
    return oracle.reads_from_packed(contig, m["start"], m["end"], m["mapq"], m["flags"], m["seq_off"],
                                    m["l_seq"], b.payload, m["op_off"], m["n_ops"], b.ops,
                                    m["refb_off"], b.refb)


ORACLE_OF = {  # product column -> oracle column
    "allele_fwd": "alleleForwardStrand", "other_fwd": "otherForwardStrand", "allele_cov": "alleleCoverage",
    "other_cov": "otherCoverage", "total_cov": "totalCoverage", "square_mapq": "squareMapQ",
    "ref_ll": "referenceLogLikelihoods", "allele_ll": "alleleLogLikelihoods",
    "other_ll": "otherLogLikelihoods", "nonref_ll": "nonRefLogLikelihoods", "first_is_ref": "isRef",
    "copy_number": "copyNumber", "n_alt": "nAlt", "n_ref": "nRef", "n_other": "nOther", "qual": "qual",
    "gq": "genotypeQuality", "sb": "strandBiasComponents", "fisher": "fisherStrandBiasPValue",
    "rms_mapq": "rmsMapQ", "gl": "genotypeLikelihoods", "nonref_gl": "nonReferenceLikelihoods"}
INT_COLS = ["allele_fwd", "other_fwd", "allele_cov", "other_cov", "total_cov", "first_is_ref",
            "copy_number", "n_alt", "n_ref", "n_other", "gq", "sb"]
FP_COLS = ["square_mapq", "ref_ll", "allele_ll", "other_ll", "nonref_ll", "qual", "fisher", "rms_mapq",
           "gl", "nonref_gl"]
REL_TOL = 1e-6   # north_star: floating-point likelihoods and qualities within 1e-6 relative


def compare_call(site_keys, cols, want, exact_fp=True):
    """site_keys[i] = (contig, start, ref, alt) of product row i.  Checks that the product emits
    exactly the oracle's sites and that every column agrees: integers / alleles / genotypes
    bit-exact, floating point within REL_TOL (and, when exact_fp, bit-identical: the product sums
    in the oracle's read order from the same libm-built table)."""
    import numpy as np
    okey = {(c, int(s), r, a): j for j, (c, s, r, a) in enumerate(
        zip(want["contigName"], want["start"], want["referenceAllele"], want["alternateAllele"]))}
    emitted = [i for i in range(len(site_keys)) if cols["emitted"][i]]
    got_keys = {site_keys[i] for i in emitted}
    assert got_keys == set(okey), "emitted sites differ: only GPU %s / only oracle %s" % (
        sorted(got_keys - set(okey))[:5], sorted(set(okey) - got_keys)[:5])
    idx_g = np.array(emitted, dtype=np.int64)
    idx_o = np.array([okey[site_keys[i]] for i in emitted], dtype=np.int64)
    stats = {"sites": len(emitted), "max_rel": 0.0, "fp_bit_identical": True}
    if len(emitted) == 0:
        return stats
    for c in INT_COLS:
        g, o = np.asarray(cols[c])[idx_g], np.asarray(want[ORACLE_OF[c]])[idx_o]
        assert np.array_equal(g.astype(np.int64), o.astype(np.int64)), "column %s differs" % c
    for c in FP_COLS:
        g = np.asarray(cols[c])[idx_g].astype(np.float64)
        o = np.asarray(want[ORACLE_OF[c]])[idx_o].astype(np.float64)
        if g.ndim == 2:
            # entries beyond a site's copy number are unspecified in the reference; compare 0..cn
            cn = np.asarray(cols["copy_number"])[idx_g]
            mask = np.arange(g.shape[1])[None, :] <= cn[:, None]
            if c in ("ref_ll", "allele_ll", "other_ll", "nonref_ll"):
                mask = np.ones_like(mask)
            g, o = g[mask], o[mask]
        same_nan = np.isnan(g) == np.isnan(o)
        assert same_nan.all(), "NaN pattern differs in " + c
        fin = ~np.isnan(g)
        inf = np.isinf(g) | np.isinf(o)
        assert np.array_equal(g[inf], o[inf]), "inf pattern differs in " + c
        fin &= ~inf
        denom = np.maximum(np.abs(o[fin]), 1e-300)
        rel = np.abs(g[fin] - o[fin]) / denom
        rel[(g[fin] == o[fin])] = 0.0
        if rel.size:
            stats["max_rel"] = max(stats["max_rel"], float(rel.max()))
        assert (rel <= REL_TOL).all(), "column %s exceeds %g relative (max %g)" % (c, REL_TOL, rel.max())
        if not np.array_equal(g[fin], o[fin]):
            stats["fp_bit_identical"] = False
    if exact_fp:
        assert stats["fp_bit_identical"], "floating-point columns are within tolerance but not bit-identical"
    return stats


NIB_CHARS = "=ACMGRSVTWYHKDBN"


def ops_string(batch, i):
    """Operators of packed read i in the oracle's notation ("2=,1X:A,1D:C,2I,1S")."""
    m = batch.meta[i]
    rb = int(m["refb_off"])
    out = []
    for k in range(int(m["n_ops"])):
        op = int(batch.ops[int(m["op_off"]) + k])
        ln, code = op >> 4, op & 15
        c = "=XIDSH"[code]
        if c == "X":
            out.append("%dX:%s" % (ln, "".join(NIB_CHARS[int(batch.refb[rb + j]) >> 4] for j in range(ln))))
            rb += ln
        elif c == "D":
            s = "".join(NIB_CHARS[(int(batch.refb[rb + j // 2]) >> (0 if j & 1 else 4)) & 15] for j in range(ln))
            out.append("%dD:%s" % (ln, s))
            rb += (ln + 1) // 2
        else:
            out.append("%d%s" % (ln, c))
    return ",".join(out)


def fuzz_records(seed, n):
    """Reads with well-formed, slightly broken and random CIGAR / MD strings (packer fuzzing)."""
    import random
    from avocado_b200.records import AlignmentRecord
    rng = random.Random(seed)
    recs = []
    for i in range(n):
        # a consistent alignment first
        cig, md, rlen, run = [], [], 0, 0
        last_del = False
        for _ in range(rng.randint(1, 6)):
            kind = rng.choice("MMMMIDSX=")
            ln = rng.randint(1, 12)
            if kind in "M=X":
                if kind == "=":
                    run += ln
                else:
                    for _k in range(ln):
                        if kind == "X" or rng.random() < 0.15:
                            md.append("%d%s" % (run, rng.choice("ACGTN")))
                            run = 0
                        else:
                            run += 1
                rlen += ln
                last_del = False
            elif kind == "D":
                md.append("%d^%s" % (run, "".join(rng.choice("ACGT") for _k in range(ln))))
                run = 0
            else:
                rlen += ln
            cig.append("%d%s" % (ln, kind))
        md.append(str(run))
        cigar, mdtag = "".join(cig), "".join(md)
        mode = rng.random()
        if mode < 0.25:          # break one of the strings
            which = rng.random()
            def mutate(s, alphabet):
                if not s:
                    return rng.choice(alphabet)
                j = rng.randrange(len(s))
                r = rng.random()
                if r < 0.4:
                    return s[:j] + rng.choice(alphabet) + s[j + 1:]
                if r < 0.7:
                    return s[:j] + s[j + 1:]
                return s[:j] + rng.choice(alphabet) + s[j:]
            if which < 0.5:
                cigar = mutate(cigar, "0123456789MIDNSHP=X*")
            else:
                mdtag = mutate(mdtag, "0123456789ACGTNacgtXU^Z")
        elif mode < 0.32:        # random junk
            cigar = "".join(rng.choice("0123456789MIDSH=X") for _ in range(rng.randint(0, 8)))
            mdtag = "".join(rng.choice("0123456789ACGT^") for _ in range(rng.randint(0, 8)))
        slen = rlen if rng.random() < 0.9 else max(0, rlen + rng.randint(-3, 3))
        qlen = slen if rng.random() < 0.9 else max(0, slen + rng.randint(-3, 3))
        start = 100 + i * 3
        # AlignmentRecord.end follows from the CIGAR (ADAM derives it); unparsable CIGARs keep a guess
        import re
        span = sum(int(n) for n, op in re.findall(r"(\d+)([MIDNSHP=X])", cigar) if op in "M=XDN") \
            if re.fullmatch(r"(\d+[MIDNSHP=X])+", cigar) else rlen
        recs.append(AlignmentRecord(
            readMapped=rng.random() > 0.03, contigName="c", start=start, end=start + span,
            cigar=None if rng.random() < 0.03 else cigar,
            mismatchingPositions=None if rng.random() < 0.03 else mdtag,
            sequence="".join(rng.choice("ACGTNacgtRY=") if rng.random() < 0.05 else rng.choice("ACGT") for _ in range(slen)),
            qual="".join(chr(33 + rng.choice([0, 2, 12, 25, 37, 41, 60, 93])) for _ in range(qlen)),
            mapq=None if rng.random() < 0.03 else rng.choice([0, 5, 20, 60, 255, 300]),
            readNegativeStrand=rng.random() < 0.5, recordGroupSample="s", readName="r%d" % i))
    return recs


def short_operators(recs, limit=255):
    """Records whose CIGAR elements are all at most `limit` long (the 6-byte wire form carries one length byte)."""
    import re
    return [r for r in recs if r.cigar is None or all(int(x) <= limit for x in re.findall(r"\d+", r.cigar))]
