"""The checker's restatement of the crushed (one line per position) writer and of RegionMapper's padding (oracle/vcf_text.py) against the reference's
own writer golden and literal asserts (src/test/Pisces.IO.Tests/UnitTests/VcfFileWriterTests.cs:162-359). CPU only."""
import os
from types import SimpleNamespace

import numpy as np

from oracle import binding as ob
from oracle.vcf_text import VcfText, pad_positions

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SNV, DELETION, REFERENCE = 0, 2, 4


def _writer():
    # VcfWriterConfig of the tests: DepthFilterThreshold 500, VariantQualityFilterThreshold 20, StrandBiasFilterThreshold 0.5,
    # FrequencyFilterThreshold = MinFrequencyThreshold = 0.007, ShouldOutputNoCallFraction
    vt = VcfText(ob.default_config(min_frequency=0.007, min_frequency_filter=0.007, vq_filter=20, low_depth_filter=500, sb_acceptance=0.5), ob.FILTERS,
                 ob.GENOTYPES)
    vt.nc = True
    return vt


def _allele(pos, kind, ref, alt, support, genotype, coverage=5394, ref_support=7, filters=(), noise=23):
    return SimpleNamespace(pos=pos, type=kind, ref=ref, alt=alt, genotype=ob.GENOTYPES.index(genotype), vq=0, gq=0, total_coverage=coverage,
                           allele_support=support, ref_support=ref_support, frequency=(min(support / coverage, 1.0) if coverage else 0.0), noise_level=noise,
                           gatk_bias_score=0.0, fraction_no_calls=0.0, filters=list(filters), n_filters=len(filters), forced=False, phase_set_index=-1)


def test_crushed_and_padded_golden():
    vt = _writer()
    calls = {7: [_allele(7, SNV, "C", "A", 2387, "HomozygousAlt")],
             10: [_allele(10, SNV, "A", "G", 2387, "HeterozygousAlt1Alt2"), _allele(10, DELETION, "AA", "G", 2000, "HeterozygousAlt1Alt2")]}
    lines = []
    for what, p in pad_positions([(2, 3), (6, 8), (10, 11)], sorted(calls)):
        if what == "call":
            lines.append(vt.crushed_line("chr4", calls[p]))
        else:   # RegionMapper.GetMissingReference (:66-82)
            lines.append(vt.crushed_line("chr4", [_allele(p, REFERENCE, "C", "C", 0, "RefLikeNoCall", coverage=0, ref_support=0,
                                                          filters=[ob.FILTERS.index("LowDepth")])]))
    exp = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, "vcfwriter_crushed_padded.records.vcf"))]
    assert lines == exp


def test_crushed_literal_line():
    vt = _writer()
    recs = [_allele(55141055, SNV, "A", "G", 2387, "HeterozygousAlt1Alt2"), _allele(55141055, DELETION, "AA", "G", 2000, "HeterozygousAlt1Alt2")]
    assert vt.crushed_line("chr4", recs) == "chr4\t55141055\t.\tAA\tGA,G\t0\tPASS\tDP=5394\tGT:GQ:AD:DP:VF:NL:SB:NC\t1/2:0:2387,2000:5394:0.8133:23:0.0000:0.0000"
    one = _allele(55141055, SNV, "A", "G", 5387, "HomozygousAlt")
    assert vt.crushed_line("chr4", [one]) == "chr4\t55141055\t.\tA\tG\t0\tPASS\tDP=5394\tGT:GQ:AD:DP:VF:NL:SB:NC\t1/1:0:7,5387:5394:0.9987:23:0.0000:0.0000"


def test_padding_walk():
    # intervals that abut, a call inside, at the edge and outside of them
    assert pad_positions([(2, 3), (4, 4), (9, 10)], [3, 7, 10]) == [("pad", 2), ("call", 3), ("pad", 4), ("call", 7), ("pad", 9), ("call", 10)]
    assert pad_positions([(5, 6)], [], write_remaining=True) == [("pad", 5), ("pad", 6)]
    assert pad_positions([(5, 6)], [1], write_remaining=False) == [("call", 1)]


def test_float_fields_round_from_seven_significant_digits():
    """VF and NC are C# floats: float.ToString("0.000") of netcoreapp2.0 reduces the value to 7 significant digits first (Number.FormatSingle) and rounds
    that decimal half away from zero. 21/2000 = 0.0105f is 0.01049999986 as a double but prints 0.011; 1/400 and 29/2000 are ties of the same kind."""
    vt = VcfText(ob.default_config(min_frequency=0.01, min_frequency_filter=0.01), ob.FILTERS, ob.GENOTYPES)   # three VF decimals
    vt.nc = True
    for support, coverage, want in ((21, 2000, "0.011"), (1, 400, "0.003"), (29, 2000, "0.015"), (1, 3, "0.333"), (1, 8, "0.125"), (2, 3, "0.667")):
        a = _allele(100, SNV, "A", "G", support, "HeterozygousAltRef", coverage=coverage, ref_support=coverage - support)
        a.fraction_no_calls = float(np.float32(support) / np.float32(coverage + support))
        sample = vt.crushed_line("chr4", [a]).split("\t")[-1].split(":")
        assert sample[4] == want, (support, coverage, sample)
        assert vt.line("chr4", a).split("\t")[-1].split(":")[4] == want
    a = _allele(100, SNV, "A", "G", 21, "HeterozygousAltRef", coverage=2000, ref_support=1979)
    a.fraction_no_calls = float(np.float32(0.00105))
    assert vt.crushed_line("chr4", [a]).split("\t")[-1].split(":")[-1] == "0.0011"
