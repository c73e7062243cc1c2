"""PrefilterReads mirror against the reference's own suite (PrefilterReadsSuite.scala:44-236)."""
from avocado_b200 import prefilter as PF
from avocado_b200.records import AlignmentRecord

CONTIGS = ["chr1", "1", "chrX", "X", "chrY", "Y", "chrM", "MT"]


def args(**kw):      # TestPrefilterReadsArgs defaults (PrefilterReadsSuite.scala:30-35)
    d = dict(autosomalOnly=False, keepMitochondrialChromosome=False, keepDuplicates=True,
             minMappingQuality=-1, keepNonPrimary=True)
    d.update(kw)
    return PF.PrefilterReadsArgs(**d)


def passing(fn, items):
    return {i for i, x in enumerate(items) if fn(x)}


def test_read_predicates():                                           # :44-75
    assert PF.filterUnique(AlignmentRecord(duplicateRead=False))
    assert not PF.filterUnique(AlignmentRecord(duplicateRead=True))
    mapped = AlignmentRecord(readMapped=True, primaryAlignment=True)
    secondary = AlignmentRecord(readMapped=True, primaryAlignment=False)
    assert PF.filterMapped(mapped, False)
    assert not PF.filterMapped(secondary, False) and PF.filterMapped(secondary, True)
    assert not PF.filterMapped(AlignmentRecord(readMapped=False), True)
    assert PF.filterMappingQuality(AlignmentRecord(mapq=None), 10)    # PrefilterReads.scala:204-205
    assert PF.filterMappingQuality(AlignmentRecord(mapq=11), 10) and not PF.filterMappingQuality(AlignmentRecord(mapq=10), 10)


def test_contig_predicates():                                         # :104-141
    assert passing(PF.filterGrcAutosome, CONTIGS) == {0}
    assert passing(PF.filterGrcSex, CONTIGS) == {2, 4}
    assert passing(PF.filterGrcMitochondrial, CONTIGS) == {6}
    assert passing(PF.filterNonGrcAutosome, CONTIGS) == {1}
    assert passing(PF.filterNonGrcSex, CONTIGS) == {3, 5}
    assert passing(PF.filterNonGrcMitochondrial, CONTIGS) == {7}
    assert passing(PF.contigFilterFn(args(autosomalOnly=True)), CONTIGS) == {0, 1}
    assert passing(PF.contigFilterFn(args()), CONTIGS) == {0, 1, 2, 3, 4, 5}
    assert passing(PF.contigFilterFn(args(keepMitochondrialChromosome=True)), CONTIGS) == set(range(8))
    assert not PF.contigFilterFn(args())(None)


READS = [AlignmentRecord(contigName=c, **kw) for kw in (dict(readMapped=False), dict(readMapped=True, duplicateRead=True),
                                                        dict(readMapped=True, duplicateRead=False)) for c in CONTIGS]


def test_read_filter_generator_and_apply():                           # :160-236
    cases = [(args(autosomalOnly=True), {8, 9, 16, 17}, 2),
             (args(), {8, 9, 10, 11, 12, 13, 16, 17, 18, 19, 20, 21}, 6),
             (args(keepMitochondrialChromosome=True), set(range(8, 24)), 8),
             (args(keepDuplicates=False, autosomalOnly=True), {16, 17}, 2),
             (args(keepDuplicates=False), {16, 17, 18, 19, 20, 21}, 6),
             (args(keepDuplicates=False, keepMitochondrialChromosome=True), set(range(16, 24)), 8)]
    for a, want, n_contigs in cases:
        fn = PF.readFilterFn(a, PF.contigFilterFn(a))
        assert passing(fn, READS) == want
        assert {i for i, k in enumerate(PF.prefilter_mask(READS, a)) if k} == want
        kept, contigs = PF.PrefilterReads(READS, CONTIGS, a)
        assert len(kept) == len(want) and len(contigs) == n_contigs


def test_packer_evaluates_the_read_predicates():
    """The host packer (avo_pack_reads_mt) given avo_prefilter keeps exactly the reads the per-record
    predicates (PrefilterReads.scala:142-209) keep, for every argument combination."""
    import itertools
    import random
    from avocado_b200 import batch as B
    from avocado_b200.prefilter import PrefilterReadsArgs, prefilter_mask, prefilter_struct
    from avocado_b200.records import AlignmentRecord
    rng = random.Random(4)
    recs = []
    for i in range(600):
        recs.append(AlignmentRecord(readMapped=rng.random() > 0.1, contigName="chr20", start=100 + i, end=110 + i, cigar="10M",
                                    mismatchingPositions="10", sequence="ACGTACGTAC", qual="IIIIIIIIII",
                                    mapq=None if rng.random() < 0.1 else rng.choice([0, 5, 10, 11, 30, 60]),
                                    readNegativeStrand=rng.random() < 0.5, readName="r%d" % i,
                                    primaryAlignment=rng.random() > 0.2, duplicateRead=rng.random() < 0.2))
    cols = B.text_columns(recs)
    for dup, nonprim, minq, contig in itertools.product([False, True], [False, True], [0, 10, 30], ["chr20", "chrM", "GL000"]):
        args = PrefilterReadsArgs(keepDuplicates=dup, keepNonPrimary=nonprim, minMappingQuality=minq)
        want = [i for i, k in enumerate(prefilter_mask([r if contig == "chr20" else __import__("dataclasses").replace(r, contigName=contig)
                                                         for r in recs], args)) if k]
        b = B.pack_text_columns(cols, threads=2, prefilter=prefilter_struct(args, contig))
        assert b.source_index.tolist() == want, (dup, nonprim, minq, contig)
