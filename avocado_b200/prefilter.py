"""PrefilterReads (CORE/util/PrefilterReads.scala:52-259): the read / contig predicates the CLI
applies before genotyping (CLI/BiallelicGenotyper.scala:218-230); SURVEY.md 8f N1.  The read predicates
run inside the packers (`prefilter_struct` -> avo_text_reads.prefilter: three flag bits and the mapq
byte of columns the packer kernels load anyway); this module holds the contig-name predicates, the
argument record, and the per-record Python form that the tests check the packers against.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

from .records import AlignmentRecord


@dataclass
class PrefilterReadsArgs:
    """PrefilterReadsArgs (PrefilterReads.scala:24-49); defaults of the CLI."""
    autosomalOnly: bool = False
    keepMitochondrialChromosome: bool = False
    keepDuplicates: bool = False
    minMappingQuality: int = 10
    keepNonPrimary: bool = False


def filterUnique(read) -> bool:                                         # :182-184
    return not read.duplicateRead


def filterMapped(read, keepNonPrimary: bool) -> bool:                   # :190-193
    return bool(read.readMapped) and (keepNonPrimary or bool(read.primaryAlignment))


def filterMappingQuality(read, minMappingQuality: int) -> bool:        # :201-209  (mapq not set: kept)
    return True if read.mapq is None else read.mapq > minMappingQuality


def filterGrcAutosome(name: Optional[str]) -> bool:                     # :216-220
    return name is not None and len(name) >= 4 and name.startswith("chr") and all(c.isdigit() for c in name[3:])


def filterGrcSex(name: Optional[str]) -> bool:                          # :227-236
    return name is not None and len(name) == 4 and name.startswith("chr") and name[3] in "XYZW"


def filterGrcMitochondrial(name: Optional[str]) -> bool:                # :243-245
    return name is not None and name == "chrM"


def filterNonGrcAutosome(name: Optional[str]) -> bool:                  # :252-254 (forall on "" is true, as in Scala)
    return name is not None and all(c.isdigit() for c in name)


def filterNonGrcSex(name: Optional[str]) -> bool:                       # :261-265
    return name is not None and name in ("X", "Y", "Z", "W")


def filterNonGrcMitochondrial(name: Optional[str]) -> bool:             # :272-274
    return name is not None and name == "MT"


def contigFilterFn(args: PrefilterReadsArgs) -> Callable[[Optional[str]], bool]:      # :118-134
    fns = [filterNonGrcAutosome, filterNonGrcSex, filterNonGrcMitochondrial,
           filterGrcAutosome, filterGrcSex, filterGrcMitochondrial]
    on = [True, not args.autosomalOnly, args.keepMitochondrialChromosome,
          True, not args.autosomalOnly, args.keepMitochondrialChromosome]
    kept = [f for f, o in zip(fns, on) if o]
    return lambda name: any(f(name) for f in kept)


def readFilterFn(args: PrefilterReadsArgs, contig_fn) -> Callable[[AlignmentRecord], bool]:   # :142-162
    def base(r):
        return filterMapped(r, args.keepNonPrimary) and filterMappingQuality(r, args.minMappingQuality) \
            and contig_fn(r.contigName)
    if args.keepDuplicates:
        return base
    return lambda r: filterUnique(r) and base(r)


def prefilter_struct(args: PrefilterReadsArgs, contig_name: Optional[str]):
    """The avo_prefilter the packers evaluate per read (host: avo_pack_reads*, device: k_pack_count /
    k_pack_records).  The contig predicates are name tests, decided here once for the batch's contig."""
    from . import _lib as L
    f = L.PrefilterStruct()
    f.min_mapq, f.keep_duplicates, f.keep_non_primary = args.minMappingQuality, int(args.keepDuplicates), int(args.keepNonPrimary)
    f.contig_kept = int(contigFilterFn(args)(contig_name))
    return f


def prefilter_mask(reads: Sequence[AlignmentRecord], args: PrefilterReadsArgs) -> List[bool]:
    """keep[i] -- the same predicates read by read in Python: the checker of the packers' own evaluation
    (tests), and a way to hand `pack_reads(records, keep=...)` an arbitrary selection."""
    fn = readFilterFn(args, contigFilterFn(args))
    return [bool(fn(r)) for r in reads]


def PrefilterReads(reads: Sequence[AlignmentRecord], contigs: Sequence[str], args: PrefilterReadsArgs
                   ) -> Tuple[List[AlignmentRecord], List[str]]:
    """PrefilterReads.apply (:66-82): (kept reads, kept contig names).  Mate fields are not part of
    the record type used here (maybeNullifyMate, :95-111, only matters for writing SAM)."""
    cfn = contigFilterFn(args)
    rfn = readFilterFn(args, cfn)
    return [r for r in reads if rfn(r)], [c for c in contigs if cfn(c)]
