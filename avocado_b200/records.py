"""Plain record types mirroring the bdg-formats Avro schemas the hot path reads and writes.

Only the fields that avocado's genotyping path touches are kept:
``AlignmentRecord`` fields read by ``Observer.observeRead`` (CORE/genotyping/Observer.scala:51-60),
``ObservationOperator.extractAlignmentOperators`` (CORE/models/ObservationOperator.scala:45-56) and
``DiscoverVariants.variantsInRead`` (CORE/genotyping/DiscoverVariants.scala:115,130-136);
``Variant`` fields read by ``DiscoveredVariant.apply`` (CORE/genotyping/DiscoveredVariant.scala:32-37);
``Genotype`` fields written by ``BiallelicGenotyper.observationToGenotype`` (BG:731-747).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional


@dataclass
class AlignmentRecord:
    readMapped: bool = False
    contigName: Optional[str] = None
    start: Optional[int] = None          # 0-based inclusive
    end: Optional[int] = None            # 0-based exclusive
    cigar: Optional[str] = None
    mismatchingPositions: Optional[str] = None   # SAM MD tag
    sequence: str = ""
    qual: str = ""                       # phred+33 text
    mapq: Optional[int] = None
    readNegativeStrand: bool = False
    readName: Optional[str] = None
    recordGroupSample: Optional[str] = None
    primaryAlignment: bool = True
    duplicateRead: bool = False

    def as_dict(self) -> dict:
        return dict(self.__dict__)


@dataclass(frozen=True)
class Variant:
    contigName: str
    start: int
    end: int
    referenceAllele: str
    alternateAllele: Optional[str]

    def as_tuple(self):
        return (self.contigName, self.start, self.end, self.referenceAllele, self.alternateAllele)


# GenotypeAllele enum values (bdg-formats)
REF, ALT, OTHER_ALT, NO_CALL = "REF", "ALT", "OTHER_ALT", "NO_CALL"


@dataclass
class Genotype:
    variant: Variant
    contigName: str
    start: int
    end: int
    sampleId: Optional[str]
    alleles: List[str]
    genotypeQuality: Optional[int]
    genotypeLikelihoods: List[float]
    nonReferenceLikelihoods: List[float]
    strandBiasComponents: List[int]
    readDepth: int
    referenceReadDepth: int
    alternateReadDepth: int
    rmsMapQ: float
    fisherStrandBiasPValue: float
    # set by JointAnnotatorCaller
    genotypePosteriors: Optional[List[float]] = None
    genotypePriors: Optional[List[float]] = None
    # set by HardFilterGenotypes (VariantCallingAnnotations.filtersApplied / filtersPassed / filtersFailed)
    filtersApplied: Optional[bool] = None
    filtersPassed: Optional[bool] = None
    filtersFailed: Optional[List[str]] = None
    # set by TrioCaller
    phased: bool = False
    # gVCF blocks read from a file (FORMAT MIN_DP)
    minReadDepth: Optional[int] = None
    # raw per-site aggregate (Observation), kept for parity checks
    observation: dict = field(default_factory=dict)
