"""The library's VCF writer (pb2_vcf_format) against the reference's own writer tests: the crushed (one line per position)
germline form, the NC tag and RegionMapper's interval padding. The records are the CalledAlleles the reference's unit tests hand to VcfFileWriter
(src/test/Pisces.IO.Tests/UnitTests/VcfFileWriterTests.cs); the expected lines are its golden file and literal asserts.
The writer is host code, but a handle needs a device: -m gpu."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SNV, DELETION = 0, 2
HOM_ALT, HET_ALT12, OTHERS = 3, 0, 12


def _handle(intervals=None, chr_sequence="ACGTACGT"):   # CHROM is the name given with the reference (pb2_set_reference)
    import pisces_b200 as pb
    # VcfWriterConfig of the tests: DepthFilterThreshold 500, VariantQualityFilterThreshold 20, StrandBiasFilterThreshold 0.5,
    # FrequencyFilterThreshold = MinFrequencyThreshold = 0.007, EstimatedBaseCallQuality 23
    cfg = pb.make_config(min_frequency=0.007, min_frequency_filter=0.007, variant_qscore_filter=20, low_depth_filter=500, strand_bias_acceptance=0.5,
                         min_base_call_quality=23)
    return pb.GpuStateManager(cfg, "chr4", chr_sequence, intervals=intervals)


def _allele(position, kind, ref, alt, support, genotype, noise=23):
    from pisces_b200 import _native as N
    r = np.zeros(1, dtype=N.RECORD_DTYPE)[0]
    r["position"], r["type"], r["genotype"], r["noise_level"] = position, kind, genotype, noise
    r["allele_support"], r["total_coverage"], r["reference_support"] = support, 5394, 7
    b = (ref + alt).encode()
    assert len(b) <= 4
    r["allele_bytes"] = int.from_bytes(b.ljust(4, b"\0"), "little")
    r["ref_len"], r["alt_len"] = len(ref), len(alt)
    return r


def _records(rows):
    from pisces_b200 import _native as N
    out = np.zeros(len(rows), dtype=N.RECORD_DTYPE)
    for i, r in enumerate(rows):
        out[i] = r
    return out


def test_crushed_and_padded_golden():
    """VcfFileWriterTests.TestDiploidStyleWithVariantsAndPadding (:162-275) -> VcfFileWriterTests_Crushed_Padded_expected.vcf, line by line."""
    sm = _handle(intervals=[(2, 3), (6, 8), (10, 11)], chr_sequence="C" * 15)
    recs = _records([_allele(7, SNV, "C", "A", 2387, HOM_ALT), _allele(10, SNV, "A", "G", 2387, HET_ALT12), _allele(10, DELETION, "AA", "G", 2000, HET_ALT12)])
    got = sm.FormatVcf(recs, crushed=True, report_no_calls=True, pad_intervals=2)
    sm.close()
    exp = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, "vcfwriter_crushed_padded.records.vcf"))]
    assert got == exp


def test_crushed_line_without_padding():
    """TestDiploidThresholdingStyleWithVariants (:279-359): the literal line the test asserts."""
    sm = _handle()
    recs = _records([_allele(55141055, SNV, "A", "G", 2387, HET_ALT12), _allele(55141055, DELETION, "AA", "G", 2000, HET_ALT12)])
    got = sm.FormatVcf(recs, crushed=True, report_no_calls=True)
    assert got == ["chr4\t55141055\t.\tAA\tGA,G\t0\tPASS\tDP=5394\tGT:GQ:AD:DP:VF:NL:SB:NC\t1/2:0:2387,2000:5394:0.8133:23:0.0000:0.0000"]
    # the same alleles uncrushed (no golden in the reference; expected lines derived by hand from SetUncrushedReferenceAndAlt, VcfFormatter.cs:432-447,
    # GetAlleleCountString :396-420 and GetFrequencyString :329-358): one line each, the other allele unspecified
    got = sm.FormatVcf(recs, report_no_calls=True)
    sm.close()
    assert got == ["chr4\t55141055\t.\tA\t.,G\t0\tPASS\tDP=5394\tGT:GQ:AD:DP:VF:NL:SB:NC\t1/2:0:7,3000,2387:5394:0.4425:23:0.0000:0.0000",
                   "chr4\t55141055\t.\tAA\t.,G\t0\tPASS\tDP=5394\tGT:GQ:AD:DP:VF:NL:SB:NC\t1/2:0:7,3387,2000:5394:0.3708:23:0.0000:0.0000"]


def test_uncrushed_somatic_line():
    """TestSomaticStyleWithVariants (:96-158): the literal line the test asserts (1/1, NC on)."""
    sm = _handle()
    r = _allele(55141055, SNV, "A", "G", 5387, HOM_ALT)
    got = sm.FormatVcf(_records([r]), report_no_calls=True)
    sm.close()
    assert got == ["chr4\t55141055\t.\tA\tG\t0\tPASS\tDP=5394\tGT:GQ:AD:DP:VF:NL:SB:NC\t1/1:0:7,5387:5394:0.9987:23:0.0000:0.0000"]


def test_float_fields_round_from_seven_significant_digits():
    """VF and NC are C# floats formatted from their 7-significant-digit decimal (float.ToString, netcoreapp2.0): 21/2000 prints 0.011, not the 0.010 the
    float widened to double would give. Same cases as tests/test_oracle_vcf_writer.py."""
    import pisces_b200 as pb
    sm = pb.GpuStateManager(pb.make_config(min_frequency=0.01, min_frequency_filter=0.01), "chr4", "ACGTACGT")   # three VF decimals
    for support, coverage, want in ((21, 2000, "0.011"), (1, 400, "0.003"), (29, 2000, "0.015"), (1, 3, "0.333"), (1, 8, "0.125"), (2, 3, "0.667")):
        r = _allele(100, SNV, "A", "G", support, 2)   # HeterozygousAltRef
        r["total_coverage"], r["reference_support"] = coverage, coverage - support
        r["fraction_no_calls"] = np.float32(0.00105)
        sample = sm.FormatVcf(_records([r]), report_no_calls=True)[0].split("\t")[-1].split(":")
        assert sample[4] == want and sample[-1] == "0.0011", (support, coverage, sample)
    sm.close()
