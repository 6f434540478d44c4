#!/usr/bin/env python3
"""Generates the PhiX fixtures under tests/golden/ from the reference's own test data (run in the build container, where
/root/reference is mounted; the fixtures travel, the reference does not).

  inputs : /root/reference/src/test/SharedData/Bams/PhiX_S3.bam, .../Genomes/PhiX/WholeGenomeFasta/genome.fa,
           /root/reference/src/test/Pisces.Tests/TestData/PhiX_S3.{noisy,Forced1,Forced2}.vcf, PhiX_S3.forcedGTInput.vcf
           (the full-text goldens of src/test/Pisces.Tests/FunctionalTests/ForcedGTFxnlTest.cs:11-112)
  outputs: PhiX_S3.bam, collapsed.test.stitched.bam (copies), phix_s3_reads.json.gz (decoded alignments), phix_genome.txt, phix_s3_{noisy,forced1,forced2}.records.vcf (record lines only),
           phix_forced_alleles.json, vcfwriter_crushed_padded.records.vcf (record lines of Pisces.IO.Tests/TestData/VcfFileWriterTests_Crushed_Padded_expected.vcf)
"""
import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import bamio  # noqa: E402

REF = "/root/reference/src/test"


def writer_fixture():
    """Record lines of the crushed / padded writer golden (src/test/Pisces.IO.Tests/UnitTests/VcfFileWriterTests.cs:162-275)."""
    src = f"{REF}/Pisces.IO.Tests/TestData/VcfFileWriterTests_Crushed_Padded_expected.vcf"
    lines = [l for l in open(src) if not l.startswith("#")]
    open(os.path.join(HERE, "vcfwriter_crushed_padded.records.vcf"), "w").writelines(lines)


def main():
    writer_fixture()
    # the two BAM files themselves (15.7 KB and 1.3 KB of test data, not source): inputs of the library's own BAM stager (pb2_bam_*)
    import shutil
    shutil.copyfile(f"{REF}/SharedData/Bams/PhiX_S3.bam", os.path.join(HERE, "PhiX_S3.bam"))
    shutil.copyfile(f"{REF}/Pisces.Tests/TestData/collapsed.test.stitched.bam", os.path.join(HERE, "collapsed.test.stitched.bam"))
    os.chmod(os.path.join(HERE, "PhiX_S3.bam"), 0o644)
    os.chmod(os.path.join(HERE, "collapsed.test.stitched.bam"), 0o644)
    _, refs, recs = bamio.read_bam(f"{REF}/SharedData/Bams/PhiX_S3.bam")
    fa = bamio.read_fasta(f"{REF}/SharedData/Genomes/PhiX/WholeGenomeFasta/genome.fa")
    out = [dict(pos0=r["pos0"], flag=r["flag"], mapq=r["mapq"], cigar=r["cigar"], seq=r["seq"], qual=r["qual"], ref_id=r["ref_id"],
                has_tags=bool(r["tags"])) for r in recs]
    with gzip.open(os.path.join(HERE, "phix_s3_reads.json.gz"), "wt") as f:
        json.dump(dict(refs=refs, reads=out), f)
    open(os.path.join(HERE, "phix_genome.txt"), "w").write(fa["phix"])
    for src, dst in (("PhiX_S3.noisy.vcf", "phix_s3_noisy.records.vcf"), ("PhiX_S3.Forced1.vcf", "phix_s3_forced1.records.vcf"),
                     ("PhiX_S3.Forced2.vcf", "phix_s3_forced2.records.vcf")):
        lines = [l for l in open(f"{REF}/Pisces.Tests/TestData/{src}") if not l.startswith("#")]
        open(os.path.join(HERE, dst), "w").writelines(lines)
    forced = []
    for l in open(f"{REF}/Pisces.Tests/TestData/PhiX_S3.forcedGTInput.vcf"):
        if l.startswith("#"):
            continue
        t = l.split("\t")
        for alt in t[4].split(","):
            forced.append([int(t[1]), t[3], alt])
    json.dump(forced, open(os.path.join(HERE, "phix_forced_alleles.json"), "w"))
    print(f"{len(out)} reads, genome {len(fa['phix'])} bp, {len(forced)} forced alleles")
    # collapsed + stitched golden: src/test/Pisces.Tests/FunctionalTests/SomaticVariantCallerFunctionalTests.cs:683-758
    _, _, crecs = bamio.read_bam(f"{REF}/Pisces.Tests/TestData/collapsed.test.stitched.bam")
    cout = [dict(pos0=r["pos0"], flag=r["flag"], mapq=r["mapq"], cigar=r["cigar"], seq=r["seq"], qual=r["qual"], xd=r["tags"].get("XD"),
                 xr=r["tags"].get("XR"), xv=r["tags"].get("XV"), xw=r["tags"].get("XW")) for r in crecs]
    json.dump(cout, open(os.path.join(HERE, "collapsed_stitched_reads.json"), "w"))
    lines = [l for l in open(f"{REF}/Pisces.Tests/TestData/test_truth.stitched.genome.vcf") if not l.startswith("#")]
    open(os.path.join(HERE, "collapsed_stitched.records.vcf"), "w").writelines(lines)
    print(f"{len(cout)} collapsed/stitched reads, {len(lines)} golden records")


if __name__ == "__main__":
    main()
