"""Pins the CPU oracle against the reference's end-to-end expectations.

Transcribed from TEST/genotyping/BiallelicGenotyperSuite.scala:383-911 (TEST =
/root/reference/avocado-core/src/test/scala/org/bdgenomics/avocado).  Inputs are the reference's
SAM resources, converted by tests/golden/make_golden.py.
"""
from helpers import load_fixture, oracle_call, oracle_discover, oracle_discover_and_call, rows


def mapq_gt(recs, thr):
    return [r for r in recs if r.mapq is not None and r.mapq > thr]


def all_ref(g):
    return all(a == "REF" for a in g["alleles"])


def all_alt(g):
    return all(a == "ALT" for a in g["alleles"])


def test_discover_and_call_simple_snp(oracle):
    # BiallelicGenotyperSuite.scala:383-410
    recs = mapq_gt(load_fixture("NA12878_snp_A2G_chr20_225058")[0], 0)
    gts = [g for g in rows(oracle_discover_and_call(oracle, recs))
           if g["totalCoverage"] > 10 and not all_ref(g) and g["genotypeQuality"] > 30]
    assert len(gts) == 1
    g = gts[0]
    assert (g["start"], g["referenceAllele"], g["alternateAllele"]) == (225057, "A", "G")
    assert g["alleles"].count("REF") == 1 and g["alleles"].count("ALT") == 1


def test_discover_and_call_simple_snp_score_all_sites(oracle):
    # :412-443
    recs = mapq_gt(load_fixture("NA12878_snp_A2G_chr20_225058")[0], 0)
    res = oracle_discover_and_call(oracle, recs, score_all=True, min_obs=6)
    gts = [g for g in rows(res) if 225000 <= g["start"] < 225120]
    assert len(gts) == 120
    assert sum(1 for g in gts if g["alternateAllele"] is None) == 119
    ref_counts = [g["alleles"].count("REF") for g in gts]
    assert ref_counts.count(2) == 119 and ref_counts.count(1) == 1
    g = [g for g in gts if g["start"] == 225057][0]
    assert (g["referenceAllele"], g["alternateAllele"]) == ("A", "G")
    assert g["alleles"].count("REF") == 1 and g["alleles"].count("ALT") == 1


def test_discover_and_call_short_indel(oracle):
    # :445-477.  The reference then applies HardFilterGenotypes (out of scope here):
    # with the indel allelic-fraction filters disabled (-1) what remains for an insertion is
    # quality > 30 by default... we assert the call itself plus that it is the only called
    # insertion with alternate depth.
    recs = mapq_gt(load_fixture("NA12878.chr1.832736")[0], 0)
    gts = [g for g in rows(oracle_discover_and_call(oracle, recs))
           if g["start"] == 832735 and g["alternateAllele"] == "AGTTTT"]
    assert len(gts) == 1
    g = gts[0]
    assert g["referenceAllele"] == "A"
    assert g["alleles"].count("REF") == 1 and g["alleles"].count("ALT") == 1
    assert g["alleleCoverage"] != 0


def test_discover_and_call_het_and_hom_snps(oracle):
    # :479-519 (het assertions are commented out in the reference; only the hom call is pinned)
    recs = mapq_gt(load_fixture("NA12878.chr1.839395")[0], 0)
    gts = [g for g in rows(oracle_discover_and_call(oracle, recs))
           if g["totalCoverage"] > 30 and g["alleles"].count("ALT") == 2]
    hom = [g for g in gts if g["start"] == 839355]
    assert len(hom) == 1
    assert (hom[0]["referenceAllele"], hom[0]["alternateAllele"]) == ("A", "C")


def test_score_single_read_covering_deletion(oracle):
    # :521-545
    recs = [r for r in load_fixture("NA12878.chr1.567239")[0]
            if r.readName == "H06JUADXX130110:1:2102:2756:44620"]
    assert len(recs) == 1
    variants = [v for v in oracle_discover(oracle, recs)
                if v[1] == 567238 and v[2] == 567240 and v[3] == "CG" and v[4] == "C"]
    assert len(variants) == 1
    obs = oracle.read_to_observations(recs[0].as_dict(), variants, 2, [], False)
    assert len(obs) == 1


def test_discover_and_force_call_hom_alt_deletion(oracle):
    # :547-572
    recs = mapq_gt(load_fixture("NA12878.chr1.567239")[0], 0)
    variants = [v for v in oracle_discover(oracle, recs)
                if v[1] == 567238 and v[2] == 567240 and v[3] == "CG" and v[4] == "C"]
    assert len(variants) == 1
    gts = rows(oracle_call(oracle, recs, variants))
    assert len(gts) == 1 and all_alt(gts[0])


def test_call_hom_alt_19bp_deletion(oracle):
    # :574-598
    recs = mapq_gt(load_fixture("NA12878.chr1.875159")[0], 0)
    variants = [v for v in oracle_discover(oracle, recs)
                if v[1] == 875158 and v[2] == 875177 and v[3] == "AGCCAGTGGACGCCGACCT" and v[4] == "A"]
    assert len(variants) == 1
    gts = rows(oracle_call(oracle, recs, variants))
    assert len(gts) == 1 and all_alt(gts[0])


def test_call_hom_alt_30bp_deletion_min_obs_3(oracle):
    # :600-619 (pins strict `count > minObservations`)
    recs = mapq_gt(load_fixture("NA12878.1_1777263")[0], 0)
    gts = rows(oracle_discover_and_call(oracle, recs, min_obs=3))
    assert len(gts) == 1
    g = gts[0]
    assert g["start"] == 1777262
    assert g["referenceAllele"] == "TACACACACACACACACACACACACACACAC" and g["alternateAllele"] == "T"
    assert all_alt(g)


def test_call_hom_alt_cag_c_deletion(oracle):
    # :621-634 (no mapq filter)
    recs = load_fixture("NA12878.1_1067596")[0]
    gts = [g for g in rows(oracle_discover_and_call(oracle, recs)) if g["start"] == 1067595]
    assert len(gts) == 1 and all_alt(gts[0])


def test_call_hom_alt_c_g_snp(oracle):
    # :636-652
    recs = mapq_gt(load_fixture("NA12878.chr1.877715")[0], 0)
    gts = [g for g in rows(oracle_discover_and_call(oracle, recs)) if g["start"] == 877714]
    assert len(gts) == 1 and all_alt(gts[0])


def test_call_hom_alt_acag_a_deletion(oracle):
    # :654-670
    recs = mapq_gt(load_fixture("NA12878.chr1.886049")[0], 0)
    gts = [g for g in rows(oracle_discover_and_call(oracle, recs)) if g["start"] == 886048]
    assert len(gts) == 1 and all_alt(gts[0])


def test_call_hom_alt_ga_cc_mnp(oracle):
    # :672-702
    recs = mapq_gt(load_fixture("NA12878.chr1.889159")[0], 0)
    gts = [g for g in rows(oracle_discover_and_call(oracle, recs, phred=30)) if not all_ref(g)]
    assert len(gts) == 2
    assert sum(1 for g in gts if (g["start"], g["referenceAllele"], g["alternateAllele"]) == (889157, "G", "C")) == 1
    assert sum(1 for g in gts if (g["start"], g["referenceAllele"], g["alternateAllele"]) == (889158, "A", "C")) == 1
    assert all(all_alt(g) for g in gts)


def test_call_het_atg_a_deletion(oracle):
    # :719-744
    recs = mapq_gt(load_fixture("NA12878.chr1.905130")[0], 0)
    gts = [g for g in rows(oracle_discover_and_call(oracle, recs, phred=30))
           if not all_ref(g) and g["start"] == 905129]
    assert len(gts) == 1
    g = gts[0]
    assert (g["contigName"], g["referenceAllele"], g["alternateAllele"]) == ("1", "ATG", "A")
    assert g["alleles"].count("ALT") == 1


def test_call_het_atg_a_deletion_score_all_sites(oracle):
    # :746-778
    recs = mapq_gt(load_fixture("NA12878.chr1.905130")[0], 0)
    res = oracle_discover_and_call(oracle, recs, score_all=True, min_obs=4, phred=30)
    gts = [g for g in rows(res) if 905100 <= g["start"] < 905200]
    assert len(gts) == 98
    assert sum(1 for g in gts if g["alternateAllele"] is None) == 95
    ref_counts = [g["alleles"].count("REF") for g in gts]
    assert ref_counts.count(2) == 95 and ref_counts.count(1) == 1 and ref_counts.count(0) == 2
    g = [g for g in gts if g["start"] == 905129][0]
    assert (g["referenceAllele"], g["alternateAllele"]) == ("ATG", "A")
    assert g["alleles"].count("ALT") == 1


def test_call_het_ag_a_deletion(oracle):
    # :780-802
    recs = mapq_gt(load_fixture("NA12878.chr1.907170")[0], 0)
    gts = [g for g in rows(oracle_discover_and_call(oracle, recs, phred=30)) if not all_ref(g)]
    assert len(gts) == 1
    g = gts[0]
    assert (g["contigName"], g["start"], g["referenceAllele"], g["alternateAllele"]) == ("1", 907169, "AG", "A")
    assert g["alleles"].count("ALT") == 1


def test_call_het_t_g_snp(oracle):
    # :804-826
    recs = mapq_gt(load_fixture("NA12878.chr1.240898")[0], 10)
    gts = [g for g in rows(oracle_discover_and_call(oracle, recs, phred=25)) if not all_ref(g)]
    assert len(gts) == 1
    g = gts[0]
    assert (g["contigName"], g["start"], g["referenceAllele"], g["alternateAllele"]) == ("1", 240897, "T", "G")
    assert g["alleles"].count("ALT") == 1


def test_het_alt_calls_at_biallelic_snp_locus(oracle):
    # :828-863.  NB the reference builds the qual string as the *digits* "8383838383".
    from avocado_b200.records import AlignmentRecord

    def make_read(allele):
        return AlignmentRecord(contigName="ctg", start=10, end=15, sequence="AC%sTG" % allele,
                               cigar="5M", mismatchingPositions="2T2", qual="8383838383", mapq=50,
                               readMapped=True, primaryAlignment=True)
    recs = [make_read("A")] * 4 + [make_read("C")] * 4
    gts = rows(oracle_discover_and_call(oracle, recs))
    assert len(gts) == 2
    assert all(len(g["alleles"]) == 2 for g in gts)
    assert all(g["alleles"].count("ALT") == 1 for g in gts)
    assert all(g["alleles"].count("OTHER_ALT") == 1 for g in gts)
    assert {g["alternateAllele"] for g in gts} == {"A", "C"}
    # GQ is 67 at both sites
    assert all(g["genotypeQuality"] == 67 for g in gts)


def test_call_hom_alt_t_taaa_insertion(oracle):
    # :865-888
    recs = load_fixture("NA12878.1_4120185")[0]
    gts = [g for g in rows(oracle_discover_and_call(oracle, recs, phred=18, min_obs=3))
           if g["start"] == 4120184]
    assert len(gts) == 2
    taaa = [g for g in gts if g["alternateAllele"] == "TAAA"][0]
    assert taaa["contigName"] == "1" and taaa["referenceAllele"] == "T"
    assert taaa["alleles"].count("ALT") == 2
    caaa = [g for g in gts if g["alternateAllele"] == "CAAA"][0]
    assert caaa["referenceAllele"] == "T"
    assert caaa["alleles"].count("OTHER_ALT") == 2


def test_call_het_alt_ttata_tta_t(oracle):
    # :890-911
    recs = load_fixture("NA12878.1_5274547")[0]
    gts = rows(oracle_discover_and_call(oracle, recs, phred=18, min_obs=3))
    assert len(gts) == 2
    assert all(g["start"] == 5274546 for g in gts)
    assert all(g["alternateAllele"] == "T" for g in gts)
    assert sum(1 for g in gts if g["referenceAllele"] == "TTA") == 1
    assert sum(1 for g in gts if g["referenceAllele"] == "TTATA") == 1
    assert all(g["alleles"].count("ALT") == 1 for g in gts)
    assert all(g["alleles"].count("OTHER_ALT") == 1 for g in gts)
