"""End-to-end pin of the oracle: the reference's full-text golden VCFs for PhiX_S3.bam, line by line.

Source: src/test/Pisces.Tests/FunctionalTests/ForcedGTFxnlTest.cs:11-112 (TestHelper.CompareFiles on PhiX_S3.noisy.vcf, .Forced1.vcf,
.Forced2.vcf). The run exercises read filtering, candidate discovery with MNVs (max length 10, gap 5), open-end collapsing, MNV
reallocation, gapped-MNV reference take-away, coverage, the Poisson q-score through the MathNet boundary, strand bias, filters, somatic
genotype / GQ, forced alleles, and the VCF text rules. Fixtures: tests/golden/make_phix_fixture.py."""
import gzip
import json
import os

import pytest

from oracle import binding as ob
from oracle.vcf_text import VcfText

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load():
    d = json.load(gzip.open(os.path.join(G, "phix_s3_reads.json.gz"), "rt"))
    genome = open(os.path.join(G, "phix_genome.txt")).read().strip()
    return d["reads"], genome


def _run(min_vq, forced):
    # Program.Main args of ForcedGTFxnlTest.cs:29-33: -c 2 -minbq 10 -minvq <q> -minvf 0.00001 -nl 40 -callMNVs TRUE -maxmnvlength 10
    # -maxgapbetweenmnv 5 -ncfilter 1 (-abfilter 0.01 has no effect without XN amplicon tags)
    kw = dict(min_coverage=2, min_base_call_quality=10, min_vq=min_vq, min_frequency=0.00001, forced_noise_level=40, call_mnvs=1, max_size_mnv=10,
              max_gap_mnv=5, no_call_filter=1.0)
    reads, genome = _load()
    c = ob.Caller(ob.default_config(**kw), "phix", genome)
    for pos, ref, alt in forced:
        if all(ch in "ACGT" for ch in alt):      # Factory.cs:80-93: forced alleles with a non-ACGT alternate (".") are dropped
            c.add_forced(pos, ref, alt)
    for r in reads:
        c.add_read(ob.SimpleRead(r["pos0"] + 1, r["seq"], r["cigar"], r["qual"], flag=r["flag"], mapq=r["mapq"], has_tags=r["has_tags"]))
    c.finish()
    vt = VcfText(ob.default_config(**kw), ob.FILTERS, ob.GENOTYPES)
    return [vt.line("phix", r) for r in c.records()]


@pytest.mark.parametrize("min_vq,forced,golden", [(1, False, "phix_s3_noisy.records.vcf"), (1, True, "phix_s3_forced1.records.vcf"),
                                                  (20, True, "phix_s3_forced2.records.vcf")])
def test_phix_full_text_golden(min_vq, forced, golden):
    fa = json.load(open(os.path.join(G, "phix_forced_alleles.json"))) if forced else []
    got = _run(min_vq, fa)
    exp = [l.rstrip("\n") for l in open(os.path.join(G, golden))]
    assert len(got) == len(exp)
    for a, b in zip(got, exp):
        assert a == b


def test_collapsed_stitched_full_text_golden():
    """src/test/Pisces.Tests/FunctionalTests/SomaticVariantCallerFunctionalTests.cs:683-758: collapsed.test.stitched.bam against the mock chr1
    built at :729-738 -> test_truth.stitched.genome.vcf (XD stitched directions, XV/XW/XR collapsed-read categories, the 12-field US tag).
    That test fills PiscesApplicationOptions by hand, so VariantCallingParameters.Validate() never runs (LowDepthFilter stays null)."""
    reads = json.load(open(os.path.join(G, "collapsed_stitched_reads.json")))
    seq = "N" * (9770498 - 1) + ("GAAGTAACAACGCAGGATGCCCCCTGGGGTGGACTGCCCCATGGAATTCTGGACCAAGGAGGAGAATCAGAGCGTTGTGGTTGACTTCCTGCTGCCCACAGGGGTCTACCTGAACTTCCCTGTGTCCCGCAATGCCAACCTC"
                                 "AGCACCATCAAGCAGGTATGGCCTCCATC")
    kw = dict(call_mnvs=1, max_size_mnv=100, max_gap_mnv=10, source_is_stitched=1, source_is_collapsed=1, apply_validation=0, amplicon_bias_filter=0.01)   # :719
    c = ob.Caller(ob.default_config(**kw), "chr1", seq)
    for r in reads:
        c.add_read(ob.SimpleRead(r["pos0"] + 1, r["seq"], r["cigar"], r["qual"], flag=r["flag"], mapq=r["mapq"], has_tags=True, xd=r["xd"], xr=r["xr"],
                                 xv=r["xv"], xw=r["xw"]))
    c.finish()
    vt = VcfText(ob.default_config(**kw), ob.FILTERS, ob.GENOTYPES, debug=True, output_bias_files=True, report_rc_counts=True, report_ts_counts=True)
    got = [vt.line("chr1", r) for r in c.records()]
    exp = [l.rstrip("\n") for l in open(os.path.join(G, "collapsed_stitched.records.vcf"))]
    assert got == exp
