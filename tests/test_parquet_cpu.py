"""Parquet Genotype output (BG:731-747, CLI/BiallelicGenotyper.scala:285): the columnar writer
fed from SoA result columns gives, read back, exactly the records the record-level mirror builds."""
import numpy as np


def _fake_results(n, ns=3, seed=1):
    rng = np.random.default_rng(seed)
    cn = rng.choice([1, 2, 2, 2], n).astype(np.int32)
    n_alt = np.array([rng.integers(0, c + 1) for c in cn], np.int32)
    n_other = np.array([rng.integers(0, c - a + 1) for c, a in zip(cn, n_alt)], np.int32)
    cols = dict(emitted=(rng.random(n) > 0.2).astype(np.uint8), copy_number=cn, n_alt=n_alt, n_other=n_other,
                n_ref=(cn - n_alt - n_other).astype(np.int32), gq=rng.integers(0, 99, n).astype(np.int32),
                total_cov=rng.integers(1, 80, n).astype(np.int32), other_cov=rng.integers(0, 40, n).astype(np.int32),
                allele_cov=rng.integers(0, 40, n).astype(np.int32), sb=rng.integers(0, 30, (n, 4)).astype(np.int32),
                fisher=rng.random(n).astype(np.float32) * 50, rms_mapq=rng.random(n).astype(np.float32) * 60,
                gl=-rng.random((n, ns)) * 100, nonref_gl=-rng.random((n, ns)) * 100,
                allele_fwd=np.zeros(n, np.int32), other_fwd=np.zeros(n, np.int32), square_mapq=np.zeros(n),
                ref_ll=np.zeros((n, ns)), allele_ll=np.zeros((n, ns)), other_ll=np.zeros((n, ns)), nonref_ll=np.zeros((n, ns)),
                first_is_ref=np.zeros(n, np.uint8), qual=np.zeros(n))
    return cols


def test_parquet_round_trip_equals_the_record_level_mirror(tmp_path):
    from avocado_b200 import batch as B
    from avocado_b200.genotyping import _genotypes_from_columns
    from avocado_b200.parquet import read_genotypes_parquet, write_genotypes_parquet
    from avocado_b200.records import Variant
    rng = np.random.default_rng(2)
    variants = []
    for i in range(300):
        kind = rng.integers(0, 3)
        ref = "".join("ACGT"[j] for j in rng.integers(0, 4, 1 if kind != 2 else rng.integers(2, 9)))
        alt = "".join("ACGT"[j] for j in rng.integers(0, 4, 1 if kind != 1 else rng.integers(2, 9)))
        variants.append(Variant("chr20", 1000 + 7 * i, 1000 + 7 * i + len(ref), ref, alt))
    table = B.sites_from_variants(variants, {"chr20": 0})
    ordered = [variants[i] for i in table.order]
    cols = _fake_results(table.n)
    path = tmp_path / "g.parquet"
    write_genotypes_parquet(path, table, cols, "chr20", "NA12878")
    got = read_genotypes_parquet(path)
    want = _genotypes_from_columns(table, cols, ordered, "NA12878")
    assert len(got) == len(want) == int(cols["emitted"].sum())
    for g, w in zip(got, want):
        assert g.variant.as_tuple() == w.variant.as_tuple() and (g.start, g.end, g.sampleId) == (w.start, w.end, w.sampleId)
        assert g.alleles == w.alleles and g.genotypeQuality == w.genotypeQuality
        assert g.genotypeLikelihoods == w.genotypeLikelihoods and g.nonReferenceLikelihoods == w.nonReferenceLikelihoods
        assert g.strandBiasComponents == w.strandBiasComponents
        assert (g.readDepth, g.referenceReadDepth, g.alternateReadDepth) == (w.readDepth, w.referenceReadDepth, w.alternateReadDepth)
        assert np.float32(g.rmsMapQ) == np.float32(w.rmsMapQ) and np.float32(g.fisherStrandBiasPValue) == np.float32(w.fisherStrandBiasPValue)
    # the schema carries the field names of bdg-formats' Genotype
    import pyarrow.parquet as pq
    names = set(pq.read_schema(path).names)
    assert {"variant", "contigName", "start", "end", "sampleId", "alleles", "genotypeQuality", "readDepth", "referenceReadDepth",
            "alternateReadDepth", "strandBiasComponents", "genotypeLikelihoods", "nonReferenceLikelihoods",
            "variantCallingAnnotations"} == names
