"""Minimal SAM text reader producing :class:`AlignmentRecord`s the way ADAM 0.24's
``SAMRecordConverter`` does for the fields the genotyping path reads (host plumbing; the
reference reaches these through ``sc.loadAlignments``, CLI/BiallelicGenotyper.scala:218-222).

start = POS-1; end = POS-1 + reference length of the CIGAR (M, D, N, =, X); ``mapq`` is None
for 255; ``readMapped`` = flag 0x4 unset; contig/start/end/mapq are only set when RNAME != '*'.
"""
from __future__ import annotations

import re
from typing import Dict, Iterable, List, Tuple

from .records import AlignmentRecord

_CIGAR_RE = re.compile(r"(\d+)([MIDNSHP=X])")


def cigar_reference_length(cigar: str) -> int:
    return sum(int(n) for n, op in _CIGAR_RE.findall(cigar) if op in "MDN=X")


def parse_sam_line(line: str, rg_samples: Dict[str, str]) -> AlignmentRecord:
    f = line.rstrip("\n").split("\t")
    flag = int(f[1])
    rec = AlignmentRecord(readName=f[0], sequence=f[9], qual=f[10])
    rec.cigar = f[5]
    tags = {}
    for t in f[11:]:
        parts = t.split(":", 2)
        if len(parts) == 3:
            tags[parts[0]] = parts[2]
    if f[2] != "*":
        rec.contigName = f[2]
        pos = int(f[3])
        rec.start = pos - 1
        reflen = cigar_reference_length(f[5]) if f[5] != "*" else 0
        # htsjdk getAlignmentEnd = start + refLength - 1 (1-based inclusive)
        rec.end = pos - 1 + reflen
        mq = int(f[4])
        rec.mapq = None if mq == 255 else mq
    rec.readMapped = not (flag & 0x4)
    rec.readNegativeStrand = bool(flag & 0x10) and rec.readMapped
    rec.primaryAlignment = not (flag & 0x100)
    rec.duplicateRead = bool(flag & 0x400)
    if "MD" in tags:
        rec.mismatchingPositions = tags["MD"]
    if "RG" in tags:
        rec.recordGroupSample = rg_samples.get(tags["RG"])
    return rec


def load_sam(path: str) -> Tuple[List[AlignmentRecord], List[str], List[str]]:
    """Returns (records, contig names in @SQ order, sample names)."""
    recs: List[AlignmentRecord] = []
    contigs: List[str] = []
    rg_samples: Dict[str, str] = {}
    with open(path) as fh:
        for line in fh:
            if not line.strip():
                continue
            if line.startswith("@"):
                f = line.rstrip("\n").split("\t")
                kv = dict(x.split(":", 1) for x in f[1:] if ":" in x)
                if f[0] == "@SQ":
                    contigs.append(kv["SN"])
                elif f[0] == "@RG" and "SM" in kv:
                    rg_samples[kv["ID"]] = kv["SM"]
                continue
            recs.append(parse_sam_line(line, rg_samples))
    samples = sorted(set(rg_samples.values()))
    return recs, contigs, samples
