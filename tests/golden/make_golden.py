"""Regenerates tests/golden/sam/*.json.gz from the reference's SAM test resources.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py

Each output holds the ADAM-style AlignmentRecord fields the genotyping path reads (see
avocado_b200/sam.py for the SAM -> record conversion), for every read of the fixture, in
file order.  The expectations that go with these inputs are transcribed, with citations,
in tests/test_oracle_e2e.py from
/root/reference/avocado-core/src/test/scala/org/bdgenomics/avocado/genotyping/BiallelicGenotyperSuite.scala.
"""
import gzip
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from avocado_b200.sam import load_sam  # noqa: E402

RES = "/root/reference/avocado-core/src/test/resources"
FIXTURES = [
    "NA12878_snp_A2G_chr20_225058.sam",
    "NA12878.chr1.832736.sam",
    "NA12878.chr1.839395.sam",
    "NA12878.chr1.567239.sam",
    "NA12878.chr1.875159.sam",
    "NA12878.1_1777263.sam",
    "NA12878.1_1067596.sam",
    "NA12878.chr1.877715.sam",
    "NA12878.chr1.886049.sam",
    "NA12878.chr1.889159.sam",
    "NA12878.chr1.866511.sam",
    "NA12878.chr1.905130.sam",
    "NA12878.chr1.907170.sam",
    "NA12878.chr1.240898.sam",
    "NA12878.1_4120185.sam",
    "NA12878.1_5274547.sam",
    "NA12878.chr1.104160.sam",
    "NA12878.chr1.875635.sam",
]
KEEP = ["readMapped", "contigName", "start", "end", "cigar", "mismatchingPositions", "sequence",
        "qual", "mapq", "readNegativeStrand", "readName", "recordGroupSample", "primaryAlignment",
        "duplicateRead"]


def main():
    out_dir = os.path.join(os.path.dirname(__file__), "sam")
    os.makedirs(out_dir, exist_ok=True)
    for name in FIXTURES:
        recs, contigs, samples = load_sam(os.path.join(RES, name))
        cols = {k: [getattr(r, k) for r in recs] for k in KEEP}
        doc = {"source": "avocado-core/src/test/resources/" + name, "contigs": contigs,
               "samples": samples, "n": len(recs), "columns": cols}
        path = os.path.join(out_dir, name[:-4] + ".json.gz")
        with gzip.GzipFile(path, "wb", mtime=0) as fh:
            fh.write(json.dumps(doc, separators=(",", ":")).encode())
        print(name, len(recs), os.path.getsize(path))


VCF_FIXTURES = ["trio.1_837214.vcf", "trio.1_837214.phased.vcf", "gvcf_multiallelic.g.vcf", "random.vcf"]


def copy_vcfs():
    """The reference's VCF test resources (they are data: the input and the expected output of
    TrioCallerSuite.scala:239-251, the inputs of SquareOffReferenceModelSuite.scala:51-76,
    HardFilterGenotypesSuite.scala:368-375 and MergeDiscoveredSuite.scala:25-33), gzip-compressed with a
    fixed time stamp -> tests/golden/vcf/*.vcf.gz (decompressed they are the reference's bytes)."""
    out_dir = os.path.join(os.path.dirname(__file__), "vcf")
    os.makedirs(out_dir, exist_ok=True)
    cli = "/root/reference/avocado-cli/src/test/resources"     # MergeDiscoveredSuite.scala:25-33
    for src_dir, name in [(RES, n) for n in VCF_FIXTURES] + [(cli, "NA12878.1_907170_AG2A.vcf"), (cli, "NA12878.1_907170_A2C.vcf")]:
        with open(os.path.join(src_dir, name), "rb") as src, \
                gzip.GzipFile(os.path.join(out_dir, name + ".gz"), "wb", mtime=0) as dst:
            dst.write(src.read())
        print(name, os.path.getsize(os.path.join(out_dir, name + ".gz")))


if __name__ == "__main__":
    main()
    copy_vcfs()
