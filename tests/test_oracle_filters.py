"""Oracle restatement of RewriteHets / HardFilterGenotypes against the reference's own suites
(RewriteHetsSuite.scala:34-205, HardFilterGenotypesSuite.scala:33-420).  Alleles: 0 REF, 1 ALT,
2 OTHER_ALT."""
REF, ALT, OTHER = 0, 1, 2

# TestHardFilterGenotypesArgs defaults (HardFilterGenotypesSuite.scala:33-53)
TEST_ARGS = dict(minQuality=30, minHetSnpQualityByDepth=2.0, minHomSnpQualityByDepth=2.0,
                 minHetIndelQualityByDepth=1.0, minHomIndelQualityByDepth=1.0, maxSnpPhredStrandBias=60.0,
                 maxIndelPhredStrandBias=200.0, minSnpRMSMappingQuality=40.0, minIndelRMSMappingQuality=30.0,
                 minSnpDepth=10, maxSnpDepth=150, minIndelDepth=20, maxIndelDepth=300,
                 minHetSnpAltAllelicFraction=0.333, maxHetSnpAltAllelicFraction=0.666,
                 minHomSnpAltAllelicFraction=0.666, minHetIndelAltAllelicFraction=0.333,
                 maxHetIndelAltAllelicFraction=0.666, minHomIndelAltAllelicFraction=0.666)
NO_FILTERS = {k: (-1 if isinstance(v, int) else -1.0) for k, v in TEST_ARGS.items() if k != "minQuality"}


def gt(ref, alt, gq, dp, adp, alleles, **kw):
    d = dict(ref=ref, alt=alt, gq=gq, dp=dp, adp=adp, alleles=alleles)
    d.update(kw)
    return d


GTS = dict(  # RewriteHetsSuite.scala:56-71
    goodHetSnp=gt("A", "T", 30, 30, 15, [REF, ALT]), badHetSnp=gt("A", "T", 30, 30, 25, [REF, ALT]),
    goodHetIndel=gt("AA", "T", 30, 50, 30, [REF, ALT]), badHetIndel=gt("A", "TCG", 30, 20, 20, [REF, ALT, ALT]),
    homRefSnp=gt("A", "T", 30, 30, 3, [REF, REF]), homRefIndel=gt("A", "TCC", 30, 50, 5, [REF]),
    homAltSnp=gt("A", "T", 30, 30, 29, [ALT]), homAltIndel=gt("A", "TCC", 30, 50, 45, [ALT, ALT]))


def test_should_rewrite(oracle):                                       # :79-105
    sr = lambda g, snps=True, indels=True: oracle.should_rewrite(g, 0.75, 0.65, snps, indels)
    assert sr(GTS["badHetSnp"]) and not sr(GTS["badHetSnp"], snps=False)
    assert sr(GTS["badHetIndel"]) and not sr(GTS["badHetIndel"], indels=False)
    for k in ("goodHetSnp", "goodHetIndel", "homRefSnp", "homRefIndel", "homAltSnp", "homAltIndel"):
        assert not sr(GTS[k]), k
    assert not sr(dict(ref="N", alt=None, alleles=[REF, ALT], gq=3, dp=10, adp=9))     # RewriteHets.scala:100-101


def test_rewrite_whole_set(oracle):                                    # :107-205
    args = dict(NO_FILTERS, minQuality=-100, maxHetSnpAltAllelicFraction=0.75, maxHetIndelAltAllelicFraction=0.65)
    out = {k: oracle.rewrite_and_filter(g, args, False, True, True) for k, g in GTS.items()}
    touched = [k for k, r in out.items() if r["rewritten"]]
    assert sorted(touched) == ["badHetIndel", "badHetSnp"]
    for k in touched:
        assert out[k]["gq"] is None and all(a == ALT for a in out[k]["alleles"])
        assert len(out[k]["alleles"]) == len(GTS[k]["alleles"])
    off = dict(args, disableHetSnpRewriting=True, disableHetIndelRewriting=True)
    assert not any(oracle.rewrite_and_filter(g, off, False, True, True)["rewritten"] for g in GTS.values())


def test_emission_filters(oracle):                                     # HardFilterGenotypesSuite.scala:106-127, 276-304
    f = lambda g, minq, filter_ref: oracle.rewrite_and_filter(g, dict(NO_FILTERS, minQuality=minq), filter_ref, False, False)
    ref = dict(ref="A", alt="T", alleles=[REF, REF], gq=99)
    alt = lambda gq: dict(ref="T", alt="A", alleles=[ALT, REF], gq=gq)
    assert not f(ref, 30, True)["emitted"] and f(ref, 30, False)["emitted"]
    assert f(alt(50), 30, True)["emitted"] and not f(alt(10), 30, True)["emitted"]
    assert not f(dict(ref="A", alt="T", alleles=[REF, OTHER], gq=99), 20, True)["emitted"]     # filterRefCalls: no ALT
    assert not f(dict(ref="A", alt="T", alleles=[OTHER, OTHER], gq=99), 20, True)["emitted"]
    assert not f(alt(10), 20, True)["emitted"]
    r = f(alt(50), 20, True)
    assert r["emitted"] and not r["filtersApplied"]                    # no filter configured (:237-245)
    # a reference-model genotype (no alternate allele) always comes through (HardFilterGenotypes.scala:656-658)
    assert f(dict(ref="N", alt=None, alleles=[REF, REF], gq=5), 30, True)["emitted"]


def test_single_filters(oracle):                                       # :129-208, 396-420
    def failed(g, **kw):
        return oracle.rewrite_and_filter(g, dict(NO_FILTERS, minQuality=-1, **kw), False, True, False)["filtersFailed"]
    alt = lambda **kw: dict(dict(ref="T", alt="A", alleles=[ALT, REF], gq=50), **kw)
    hom = lambda **kw: dict(dict(ref="T", alt="A", alleles=[ALT, ALT], gq=50), **kw)
    assert failed(alt(dp=20), minHetSnpQualityByDepth=2.0) == []
    assert failed(alt(dp=50), minHetSnpQualityByDepth=2.0) == ["HETSNPQD"]
    assert failed(alt(dp=50), minHomSnpQualityByDepth=2.0) == []          # het call, hom filter
    assert failed(hom(dp=50), minHetSnpQualityByDepth=2.0) == []          # hom call, het filter
    assert failed(hom(dp=50), minHomSnpQualityByDepth=2.0) == ["HOMSNPQD"]
    assert failed(alt(dp=50), minSnpDepth=10) == [] and failed(alt(dp=5), minSnpDepth=10) == ["SNPMINDP"]
    assert failed(alt(dp=50), maxSnpDepth=10) == ["SNPMAXDP"] and failed(alt(dp=5), maxSnpDepth=10) == []
    assert failed(alt(rms=50.0), minSnpRMSMappingQuality=30.0) == []
    assert failed(alt(rms=10.0), minSnpRMSMappingQuality=30.0) == ["SNPMQ"]
    assert failed(alt(fs=50.0), maxSnpPhredStrandBias=30.0) == ["SNPFS"]
    assert failed(alt(fs=10.0), maxSnpPhredStrandBias=30.0) == []
    assert failed(alt(dp=50, adp=25), minHetSnpAltAllelicFraction=0.333) == []
    assert failed(alt(dp=50, adp=10), minHetSnpAltAllelicFraction=0.333) == ["HETSNPMINAF"]
    assert failed(alt(dp=50, adp=10), minHomSnpAltAllelicFraction=0.333) == []
    assert failed(alt(dp=50, adp=40), maxHetSnpAltAllelicFraction=0.666) == ["HETSNPMAXAF"]
    assert failed(dict(ref="T", alt="A", alleles=[ALT, OTHER], gq=50, dp=50, adp=40), maxHetSnpAltAllelicFraction=0.666) == []
    assert failed(alt(), minSnpDepth=10, minHetSnpQualityByDepth=2.0) == []   # unset depth: nothing to test


def test_build_filters_and_apply(oracle):                              # :306-383
    snp = dict(ref="A", alt="T", alleles=[ALT, REF], gq=30, adp=15, dp=20, rms=35.0, fs=70.0)
    r = oracle.rewrite_and_filter(snp, dict(TEST_ARGS, minQuality=20), True, False, False)
    assert r["emitted"] and r["filtersApplied"] and not r["filtersPassed"]
    assert set(r["filtersFailed"]) == {"SNPMQ", "HETSNPQD", "SNPFS", "HETSNPMAXAF"}
    indel = dict(ref="A", alt="AT", alleles=[ALT, REF], gq=30, dp=20, rms=35.0, fs=70.0)
    r = oracle.rewrite_and_filter(indel, dict(TEST_ARGS, minQuality=20), True, False, False)
    assert r["emitted"] and r["filtersPassed"] and r["filtersFailed"] == []
    bad = dict(ref="A", alt="AT", alleles=[ALT, REF], gq=15, dp=20, rms=25.0, fs=250.0)
    r = oracle.rewrite_and_filter(bad, dict(TEST_ARGS, minQuality=10), True, False, False)
    assert r["emitted"] and set(r["filtersFailed"]) == {"INDELMQ", "HETINDELQD", "INDELFS"}
