"""Known-answer tests of the reference for the germline genotypers, re-asserted against the oracle (oracle/po_genotype.hpp):
DiploidGenotypeQualityCalculatorTests.cs:15-130, GenotypeCalculatorTest.cs:27-88, HaploidGenotypeCalculatorTests.cs:55-83,
StrandBiasCalculatorTests.cs:93-155,157-174,176-215 (Diploid strand-bias model, MathNet Binomial CDF), GenotypeCreatorTests.cs."""
import ctypes as C
import math

import pytest

from oracle import binding as ob
from oracle.binding import SNV, INSERTION, DELETION, REFERENCE

G = ob.GENOTYPES.index
DIPLOID, HAPLOID = 1, 3
INT_MAX = 2 ** 31 - 1


def test_mathnet_binomial_cdf():
    # StrandBiasCalculatorTests.TestDistributionFxn :157-174
    L = ob.lib()
    for x, want in ((15, 0.129), (20, 0.559), (25, 0.913)):
        assert L.po_mathnet_binomial_cdf(0.2, 100, x) == pytest.approx(want, abs=5e-4)
    from scipy import stats
    for p, n, x in ((0.2, 100, 15.0), (0.2, 10000, 1900.0), (0.4, 37, 36.0), (0.05, 500, 0.0), (0.9, 64, 64.0)):
        assert L.po_mathnet_binomial_cdf(p, n, x) == pytest.approx(stats.binom.cdf(x, n, p), rel=1e-9, abs=1e-300)
        k = int(x)
        assert L.po_mathnet_binomial_probability_ln(p, n, k) == pytest.approx(stats.binom.logpmf(k, n, p), rel=1e-9)


def _gq_row(gt, depth, freqs):
    out = []
    for f in freqs:   # TestCalculation :121-130
        sup = int(depth * (1.0 - f)) if gt == "HomozygousRef" else int(depth * f)
        out.append(ob.lib().po_diploid_gq(G(gt), int(depth), sup, 0, INT_MAX))
    return out


def test_diploid_genotype_quality_table():
    # DiploidGenotypeQualityCalculatorTests.ComputeGenotypeQualityTests :15-101 ('truth data' from Excel)
    het_f = [0.2, 0.21, 0.25, 0.30, 0.35, 0.45, 0.49, 0.50, 0.51, 0.55, 0.59, 0.60, 0.61, 0.68, 0.69]
    het_q = [0, 0, 18, 57, 96, 174, 205, 212, 201, 156, 122, 99, 88, 9, 0]
    assert _gq_row("HomozygousRef", 100, [0, 0.01, 0.05, 0.10, 0.15, 0.19]) == [200, 188, 144, 89, 36, 0]
    assert _gq_row("HeterozygousAltRef", 100, het_f) == het_q
    assert _gq_row("HomozygousAlt", 100, [0.7, 0.71, 0.75, 0.80, 0.85, 0.90, 0.95, 0.99, 1.0]) == [0, 7, 54, 114, 175, 237, 300, 352, 365]
    assert _gq_row("HeterozygousAlt1Alt2", 100, het_f) == het_q
    for gt in ("RefLikeNoCall", "AltLikeNoCall"):
        assert _gq_row(gt, 100, [0, 0.2, 0.5, 1.0]) == [0, 0, 0, 0]
        assert _gq_row(gt, 1000, [0, 0.2, 0.5, 1.0]) == [0, 0, 0, 0]
    assert _gq_row("HomozygousRef", 1000, [0, 0.19]) == [2001, 0]
    assert _gq_row("HeterozygousAltRef", 1000, [0.2, 0.5, 0.69]) == [0, 2129, 0]
    assert _gq_row("HomozygousAlt", 1000, [0.7, 1.0]) == [0, 3653]
    assert _gq_row("HeterozygousAlt1Alt2", 1000, [0.2, 0.5, 0.69]) == [0, 2129, 0]
    # HigFreqInsertionGT_Test :103-118: more insertion calls than coverage
    assert _gq_row("HomozygousAlt", 100, [1.19, 0.00]) == [INT_MAX, 0]


def _alleles(ref_freqs, alt_freqs, coverage):
    # ExecuteDiploidGenotypeTest :95-121 (float arithmetic of the test harness kept)
    import numpy as np
    f32 = np.float32
    out, ref_freq = [], 0.0
    for rf in ref_freqs:
        s = int(f32(rf) * f32(coverage))
        out.append((REFERENCE, s, coverage, s))
        ref_freq = float(f32(rf))
    if ref_freq == 0:
        ref_freq = 1.0 - float(sum((f32(a) for a in alt_freqs), f32(0)))
    for vf in alt_freqs:
        out.append((SNV, int(f32(vf) * f32(coverage)), coverage, int(ref_freq * coverage)))
    return out


DIPLOID_SCENARIOS = [   # GenotypeCalculatorTest.DiploidGenotypeScenarios :27-88
    ("HomozygousRef", 1, [0.80], [0.19]), ("HomozygousRef", 0, [0.80], []),
    ("HeterozygousAltRef", 0, [0.80], [0.20]), ("HeterozygousAltRef", 0, [0.70], [0.30]), ("HeterozygousAltRef", 0, [0.21], [0.69]),
    ("HeterozygousAltRef", 1, [0.69], [0.30, 0.01]), ("HeterozygousAltRef", 0, [], [0.20]), ("HeterozygousAltRef", 0, [], [0.30]),
    ("HeterozygousAltRef", 1, [], [0.30, 0.01]), ("HeterozygousAltRef", 2, [], [0.01, 0.02, 0.30]),
    ("AltAndNoCall", 0, [0.10], [0.70]),
    ("HomozygousAlt", 0, [0.10], [0.71]), ("HomozygousAlt", 0, [0.10], [0.99]), ("HomozygousAlt", 0, [0.10], [1.0]), ("HomozygousAlt", 0, [], [0.71]),
    ("HomozygousAlt", 0, [], [0.99]), ("HomozygousAlt", 0, [], [1.0]), ("HomozygousAlt", 1, [0.10], [0.99, 0.01]), ("HomozygousAlt", 1, [], [0.99, 0.01]),
    ("AltLikeNoCall", 1, [0.20], [0.40, 0.40]), ("AltLikeNoCall", 1, [0.20], [0.20, 0.40]), ("AltLikeNoCall", 2, [0.20], [0.20, 0.40, 0.02]),
    ("Alt12LikeNoCall", 0, [0.01], [0.40, 0.39]), ("Alt12LikeNoCall", 0, [0.0], [0.20, 0.40]), ("AltLikeNoCall", 2, [], [0.20, 0.40, 0.02]),
    ("AltLikeNoCall", 2, [0.20], [0.20, 0.40, 0.20]), ("AltLikeNoCall", 2, [0.30], [0.20, 0.30, 0.30]), ("AltLikeNoCall", 1, [0.80], [0.20, 0.20]),
    ("HeterozygousAltRef", 1, [0.60], [0.40, 0.01]),
    ("HeterozygousAlt1Alt2", 0, [], [0.50, 0.50]), ("HeterozygousAlt1Alt2", 0, [0.01], [0.40, 0.40]), ("HeterozygousAlt1Alt2", 1, [0.01], [0.35, 0.55, 0.01]),
]


@pytest.mark.parametrize("want,n_prune,refs,alts", DIPLOID_SCENARIOS)
def test_diploid_genotype_scenarios(want, n_prune, refs, alts):
    gt, pruned, _, _ = ob.genotype_locus(DIPLOID, _alleles(refs, alts, 1000), min_depth=100)
    assert gt == want and sum(pruned) == n_prune


@pytest.mark.parametrize("want,n_prune,refs,alts,cov", [   # :83-87 (depth less than required -> no call)
    ("RefAndNoCall", 2, [0.20], [0.01, 0.01], 1000), ("AltAndNoCall", 1, [0.10], [0.21, 0.01], 1000),
    ("RefLikeNoCall", 2, [0.20], [0.01, 0.01], 10), ("AltLikeNoCall", 1, [0.10], [0.21, 0.01], 10)])
def test_diploid_genotype_depth(want, n_prune, refs, alts, cov):
    gt, pruned, _, _ = ob.genotype_locus(DIPLOID, _alleles(refs, alts, cov), min_depth=100)
    assert gt == want and sum(pruned) == n_prune


def test_diploid_multiallelic_site_mixed_types():
    # GenotypeCalculatorTest.ExecuteDiploidMultiAllelicSiteGenotypeTest :140-200: SNP + insertion + deletion -> 1/2, the lowest frequency pruned, no filter
    al = [(SNV, 600, 1000, 400, "A>C"), (INSERTION, 400, 1000, 600, "A>AGGG"), (DELETION, 100, 1000, 900, "ACT>A")]
    gt, pruned, multi, _ = ob.genotype_locus(DIPLOID, al, min_depth=100)
    assert gt == "HeterozygousAlt1Alt2" and pruned == [0, 0, 1] and multi == 0
    al = [(INSERTION, 600, 1000, 400, "A>ACCAT"), (SNV, 100, 1000, 200, "A>G"), (SNV, 400, 1000, 200, "A>C")]
    gt, pruned, multi, _ = ob.genotype_locus(DIPLOID, al, min_depth=100)
    assert gt == "HeterozygousAlt1Alt2" and pruned == [0, 1, 0] and multi == 0


@pytest.mark.parametrize("want,n_prune,ref,alts,cov", [   # HaploidGenotypeCalculatorTests.cs:55-83
    ("HemizygousRef", 2, 0.80, [0.01, 0.01], 1000), ("HemizygousNoCall", 2, 0.70, [0.01, 0.01], 1000), ("HemizygousNoCall", 2, 0.22, [0.75, 0.01], 1000),
    ("HemizygousNoCall", 2, 0.80, [0.01, 0.01], 10), ("HemizygousAlt", 1, 0.10, [0.75, 0.01], 1000)])
def test_haploid_genotype_scenarios(want, n_prune, ref, alts, cov):
    gt, pruned, _, _ = ob.genotype_locus(HAPLOID, _alleles([ref], alts, cov), min_depth=100)
    assert gt == want and sum(pruned) == n_prune


def test_ploidy_for_chromosome():
    # GenotypeCreator.GetPloidyForThisChr :39-68 (GenotypeCreatorTests.cs)
    P = lambda s, m, c: ob.lib().po_ploidy_for_chr(s, m, c.encode())
    SOM, DIP, HAP = 0, 1, 3
    assert P(SOM, 1, "chrX") == SOM and P(DIP, -1, "chrM") == SOM and P(DIP, 1, "M") == SOM
    assert P(HAP, -1, "chr1") == HAP
    assert P(DIP, -1, "chrX") == DIP and P(DIP, 1, "chrX") == HAP and P(DIP, 1, "Y") == HAP and P(DIP, 1, "chr2") == DIP
    assert P(DIP, 0, "chrX") == DIP and P(DIP, 0, "chrY") == HAP


def _sb(cov, sup, q, min_vf, acc, model):
    out = (C.c_double * 29)()
    ob.lib().po_strand_bias((C.c_int32 * 3)(*cov), (C.c_int32 * 3)(*sup), q, min_vf, acc, model, out)
    return list(out)


def test_diploid_strand_bias_model():
    # StrandBiasCalculatorTests.TestSBCalculationsForSomaticAndDiploidSettings :93-155 (model 2 = Diploid)
    cov = [10000, 10000, 0]
    r = _sb(cov, [2500, 2500, 0], 20, 0.20, 0.5, 2)
    assert r[0] == 0 and r[1] == -math.inf and r[2] == 1
    r = _sb(cov, [500, 2500, 0], 20, 0.20, 0.5, 2)
    assert math.log10(r[0]) == pytest.approx(74.3, abs=0.05) and r[1] == pytest.approx(743.5, abs=0.05) and r[2] == 0
    r = _sb(cov, [200, 50, 0], 20, 0.20, 0.5, 2)
    assert r[0] == pytest.approx(1.0, abs=5e-4) and r[1] == pytest.approx(0.0, abs=5e-4) and r[2] == 0
    # TestPopulateDiploidStats :176-215: stats of one strand: [FN, FP, VG] = out[5 + 6*k ...]; overall stats first
    r = _sb([100, 100, 0], [15, 15, 0], 20, 0.20, 0.5, 2)
    fwd = r[11:17]
    assert fwd[0] == pytest.approx(0.129, abs=5e-4) and fwd[1] == pytest.approx(0.049, abs=5e-4) and fwd[2] == pytest.approx(0.129, abs=5e-4)
    r = _sb([100, 100, 0], [20, 50, 0], 20, 0.20, 0.5, 2)
    assert r[11:14] == [1, 0, 1] and r[17:20] == [1, 0, 1]
