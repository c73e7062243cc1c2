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
