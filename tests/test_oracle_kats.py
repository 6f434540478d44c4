"""Pins the CPU oracle (oracle/) against the reference's own known-answer tests. No GPU needed.

Every case cites the reference test it was taken from (paths relative to /root/reference/src/test). The reference cannot run in the
build container (no dotnet), so these vectors — together with the PhiX full-text golden in test_oracle_phix_golden.py — are what
anchors the oracle to the reference's behaviour.
"""
import math

import pytest

from oracle import binding as ob
from oracle.binding import A, G, Cc, T, N, DEL, FWD, REV, STITCHED, SNV, INSERTION, DELETION, MNV, REFERENCE, SimpleRead

L = ob.lib()


# ---------------------------------------------------------------------------------------------- variant Q (MathNet boundary)
@pytest.mark.parametrize("cov,support,q", [(100, 0, 0), (100, 1, 2), (100, 5, 24), (200, 10, 43), (500, 25, 98), (5000, 250, 890),
                                           (10000, 250, 356), (10000, 500, 1770), (10000, 9995, 156912)])
def test_assign_poisson_qscore(cov, support, q):
    # Pisces.Calculators.Tests/UnitTests/QualityCalculatorTests.cs:62-95 (Compute): uncapped and capped at 100
    assert L.po_vq(support, cov, 20, 10 ** 9) == q
    assert L.po_vq(support, cov, 20, 100) == min(q, 100)


def test_raw_q_extreme():
    # PoissonTests.cs:94-103,177 — raw 156911.8104 (4 dp) through the pValue<=0 fallback branch
    assert round(L.po_raw_vq(9995, 10000, 20), 4) == 156911.8104


@pytest.mark.parametrize("k,n,p", [(1, 100, 0.6321), (5, 100, 0.003659), (10, 200, 4.65e-5), (25, 500, 1.599e-10)])
def test_assign_pvalue(k, n, p):
    # QualityCalculatorTests.cs:17-57 (Pisces' own Poisson.Cdf)
    assert L.po_pvalue(k, n, 20) == pytest.approx(p, rel=2e-3)


EXCEL = [(5, 0.559506715, 2.5219), (10, 0.031828057, 14.9719), (15, 0.000226254, 36.4540), (20, 3.45214e-07, 64.6191),
         (25, 1.59959e-10, 97.9599), (30, 2.81997e-14, 135.4976)]


@pytest.mark.parametrize("k,p,q", EXCEL)
def test_excel_table_depth_500(k, p, q):
    # PoissonTests.cs:77-91 — exact p and Q at n=500, noise 0.01 (Excel), 4 dp on Q
    assert L.po_pvalue(k, 500, 20) == pytest.approx(p, rel=1e-5)
    assert L.po_raw_vq(k, 500, 20) == pytest.approx(q, abs=6e-5)


def test_variant_caller_mock_q():
    # Pisces.Tests/UnitTests/VariantCalling/VariantCallerTests.cs:65-68 — support 40, cov 1500 → Q 72
    assert L.po_vq(40, 1500, 20, 100) == 72
    # PhiX_S3.noisy.vcf pos 2: support 1, cov 248, NL 40 → 16
    assert L.po_vq(1, 248, 40, 100) == 16


def test_mathnet_restatement_against_scipy():
    # independent cross-check of the IL-derived restatement (scipy is only a checker here)
    sp = pytest.importorskip("scipy.special")
    for a, x in [(1, 0.5), (5, 1.0), (25, 5.0), (3, 10.0), (250, 50.0), (40, 15.0), (1000, 100.0), (2, 2.5), (700, 650.0), (10, 30.0)]:
        assert L.po_mathnet_gamma_lower_regularized(a, x) == pytest.approx(float(sp.gammainc(a, x)), rel=1e-12, abs=1e-300)
    for z in [0.3, 0.5, 1.0, 2.5, 10.0, 171.0, 172.5, 1000.0, 1e5]:
        assert L.po_mathnet_gamma_ln(z) == pytest.approx(float(sp.gammaln(z)), rel=1e-13, abs=1e-13)
    for k, lam in [(0, 1.0), (4, 1.0), (24, 5.0), (10, 20.0), (249, 50.0), (500, 450.0), (3, 0.02)]:
        # Pisces' own Poisson.Cdf(k, λ) = Q(int(k+1), λ)
        assert L.po_poisson_cdf(k, lam) == pytest.approx(float(sp.gammaincc(k + 1, lam)), rel=1e-9)


# ---------------------------------------------------------------------------------------------- strand bias
def test_sb_phix_golden_value():
    # Pisces.Tests/TestData/PhiX_S3.noisy.vcf pos 4 T>G: cov F/R 49/199, support 0/1, NL 40 → SB -16.9682
    r = ob.strand_bias([49, 199, 0], [0, 1, 0], 40)
    assert f"{r['gatk']:.4f}" == "-16.9682"


def test_sb_poisson_high_depth():
    # StrandBiasCalculatorTests.cs:293-306 — cov 70038/65998, support 54/11, Poisson model → bias 1.0, GATK 0
    r = ob.strand_bias([70038, 65998, 0], [54, 11, 0], 20, model=0)
    assert r["bias"] == 1.0 and r["gatk"] == 0.0 and not r["acceptable"]


def test_sb_somatic_scenarios():
    # StrandBiasCalculatorTests.cs:93-155 (Extended model rows)
    r = ob.strand_bias([10000, 10000, 0], [2500, 2500, 0], 20)
    assert r["bias"] == 0 and r["gatk"] == -math.inf and r["acceptable"]
    r = ob.strand_bias([10000, 10000, 0], [500, 2500, 0], 20)
    assert r["bias"] == 0 and r["acceptable"]
    r = ob.strand_bias([10000, 10000, 0], [200, 50, 0], 20)
    assert r["bias"] == pytest.approx(1.0, abs=1e-3) and not r["acceptable"]


def test_sb_single_strand_coverage():
    # StrandBiasCalculator.cs:60-66: coverage on one strand only → (0, -inf), acceptable
    r = ob.strand_bias([100, 0, 0], [10, 0, 0], 20)
    assert r["bias"] == 0 and r["gatk"] == -math.inf and r["acceptable"] and not r["cov_both"]


def test_sb_stitched_integer_halving():
    # StrandBiasCalculator.cs:36-41: stitched/2 is integer division
    r = ob.strand_bias([10, 10, 5], [1, 1, 3], 20)
    assert r["fwd"]["coverage"] == 12 and r["rev"]["coverage"] == 12 and r["fwd"]["support"] == 2 and r["overall"]["support"] == 5


# ---------------------------------------------------------------------------------------------- somatic GQ
GT = {n: i for i, n in enumerate(ob.GENOTYPES)}


@pytest.mark.parametrize("lod,ref_freq,gq", [(0.05, 1.00, 217), (0.05, 0.99, 119), (0.05, 0.95, 0), (0.05, 0.90, 0),
                                             (0.10, 1.00, 434), (0.10, 0.99, 309), (0.10, 0.95, 76), (0.10, 0.90, 0)])
def test_somatic_gq(lod, ref_freq, gq):
    # Pisces.Genotyping.Tests/SomaticGenotypeQualityCalculatorTests.cs:16-85 — VQ 1000, depth 1000, 0/0 genotype
    support = int(round(ref_freq * 1000))
    assert L.po_somatic_gq(REFERENCE, GT["HomozygousRef"], 1000, 1000, support, lod, 0, 10 ** 6) == gq


def test_somatic_gq_nocall_and_het():
    assert L.po_somatic_gq(SNV, GT["HeterozygousAltRef"], 57, 1000, 100, 0.05, 0, 100) == 57  # het → VQ
    assert L.po_somatic_gq(REFERENCE, GT["RefLikeNoCall"], 57, 1000, 100, 0.05, 3, 100) == 3  # no-call → minGQ
    assert L.po_somatic_gq(REFERENCE, GT["HomozygousRef"], 57, 0, 0, 0.05, 3, 100) == 3       # zero coverage → minGQ


def test_somatic_genotype_thresholds():
    # SomaticGenotyper.cs:65-100
    g = lambda *a: ob.GENOTYPES[L.po_somatic_genotype(*a)]
    assert g(REFERENCE, 5, 5, 5, 0.01, 10) == "RefLikeNoCall"
    assert g(SNV, 5, 5, 0, 0.01, 10) == "AltLikeNoCall"
    assert g(SNV, 1000, 100, 900, 0.01, 10) == "HeterozygousAltRef"
    assert g(SNV, 1000, 1000, 0, 0.01, 10) == "HomozygousAlt"
    assert g(SNV, 1000, 900, 5, 0.01, 10) == "AltAndNoCall"
    assert g(REFERENCE, 1000, 1000, 1000, 0.01, 10) == "HomozygousRef"
    assert g(REFERENCE, 1000, 980, 980, 0.01, 10) == "RefAndNoCall"
    assert g(REFERENCE, 1000, 5, 5, 0.01, 10) == "RefLikeNoCall"


# ---------------------------------------------------------------------------------------------- anchor-adjusted counts
def test_anchor_adjusted_counts():
    # Pisces.Processing.Tests/UnitTests/AlleleCountHelperTests.cs:10-71 — bins {0:50, 4:2, 5:5, 6:3, 10:300}, K=5
    import ctypes as C
    bins = [0] * 11
    bins[0], bins[4], bins[5], bins[6], bins[10] = 50, 2, 5, 3, 300
    arr = (C.c_int32 * 11)(*bins)
    f = lambda mn, mx=-1, fe=0, sym=0: L.po_anchor_adjusted_count(arr, 5, mn, mx, fe, sym)
    assert f(5) == 308
    assert f(2) == 310
    assert f(2, sym=1) == 10
    assert f(0) == 360
    assert f(0, mx=4) == 52
    assert f(0, mx=3) == 50
    assert f(5, fe=1) == 57
    assert f(2, fe=1) == 60
    assert f(0, mx=4, fe=1) == 303


# ---------------------------------------------------------------------------------------------- counting (RegionStateManagerTests.cs)
def _state(min_bq=25, **kw):
    return ob.Caller(ob.default_config(min_base_call_quality=min_bq, output_gvcf=0, **kw), "chr1", "A" * 3000)


def test_add_and_get_allele_counts():
    # RegionStateManagerTests.cs:390-475. read2's unmapped index 7 is expressed as an insertion; the stitched read via an XD tag.
    s = _state()
    s.add_read(SimpleRead(1001, "ACTGGCATC", "9M", 25), "counts")
    s.add_read(SimpleRead(1005, "TCTGCCACT", "7M1I1M", 25, flag=0x10), "counts")
    s.add_read(SimpleRead(999, "ACAC", "4M", 25, xd="4S"), "counts")
    s.add_read(SimpleRead(999, "ACAC", "4M", 25), "counts")
    exp = [(1004, G, FWD, 1), (1005, G, FWD, 1), (1005, T, REV, 1), (1006, Cc, FWD, 1), (1006, Cc, REV, 1), (1007, A, FWD, 1), (1007, T, REV, 1),
           (1008, T, FWD, 1), (1008, G, REV, 1), (1009, Cc, FWD, 1), (1009, Cc, REV, 1), (1010, Cc, REV, 1), (1012, Cc, REV, 0),
           (999, A, STITCHED, 1), (1000, Cc, STITCHED, 1), (1001, A, STITCHED, 1), (1002, Cc, STITCHED, 1), (1001, A, FWD, 2), (1002, Cc, FWD, 2)]
    for pos, al, d, n in exp:
        assert s.count(pos, al, d) == n, (pos, al, d)
    with pytest.raises(RuntimeError):
        s.add_read(SimpleRead(0, "A", "1M", 25), "counts")  # position must be > 0 (RegionStateManager.cs:363-364)
    # no calls and low quality bases map to N
    s.add_read(SimpleRead(999, "NNAC", "4M", [25, 25, 24, 24]), "counts")
    assert s.count(999, N, FWD) == 1 and s.count(1000, N, FWD) == 1 and s.count(1001, N, FWD) == 1
    assert s.count(1001, A, FWD) == 2 and s.count(1002, Cc, FWD) == 2 and s.count(999, A, STITCHED) == 1


@pytest.mark.parametrize("low_q", [False, True])
def test_allele_counts_deletions(low_q):
    # RegionStateManagerTests.cs:478-592 (poor-quality variant) and :596-704
    hi, lo = (30, 20) if low_q else (25, 25)
    exp_lo = 0 if low_q else 1
    s = _state()
    s.add_read(SimpleRead(1001, "TTTTTTTTT", "5M4D4M", hi), "counts")
    s.add_read(SimpleRead(1005, "AAAAAAAAA", "1M2D8M", lo, flag=0x10), "counts")
    assert s.count(1000, T, FWD) == 0
    for i in range(1001, 1014):
        assert s.count(i, DEL if 1006 <= i <= 1009 else T, FWD) == 1
    assert s.count(1014, T, FWD) == 0
    assert s.count(1004, A, REV) == 0
    for i in range(1005, 1016):
        assert s.count(i, DEL if 1006 <= i <= 1007 else A, REV) == exp_lo
    # read beginning with a deletion
    s = _state()
    s.add_read(SimpleRead(1001, "NNNNNTTTT", "5S2D4M", lo), "counts")
    s.add_read(SimpleRead(1005, "AAAAAAAAA", "9M", hi), "counts")
    for i in range(1001, 1006):
        assert s.count(i, DEL if i <= 1002 else T, FWD) == exp_lo
    # terminal deletions
    s = _state()
    s.add_read(SimpleRead(1001, "TTTTNNNNN", "4M2D5S", hi), "counts")
    s.add_read(SimpleRead(1015, "AAAAAAAAA", "9M2D", lo, flag=0x10), "counts")
    for i in range(1001, 1007):
        assert s.count(i, DEL if i >= 1005 else T, FWD) == 1
    assert s.count(1007, DEL, FWD) == 0
    for i in range(1015, 1026):
        assert s.count(i, DEL if i >= 1024 else A, REV) == exp_lo
    assert s.count(1026, DEL, REV) == 0


def test_anchor_bins_and_quality_sums():
    # RegionStateManager.cs:83-116 GetAnchorType; :191 float exponent in the base-quality sum
    s = _state(min_bq=20)
    s.add_read(SimpleRead(101, "A" * 20, "20M", 30), "counts")
    d = s.dump_counts(101, 20)
    for i in range(20):
        left, right = i, 19 - i
        b = (5 if right >= 5 else 10 - right) if left >= right else (5 if left >= 5 else left)
        assert d[i, A, FWD, b] == 1 and d[i].sum() == 1
    assert s.qsum(105, A, FWD) == pytest.approx(1e-3, rel=1e-6)


def test_collapsed_counts():
    # CollapsedRegionState.cs:28-44 / Read.cs:17-71: duplex stitched; simplex FR non-stitched bumps the aggregate too
    s = _state(min_bq=20, source_is_stitched=1, source_is_collapsed=1)
    s.add_read(SimpleRead(11, "ACGT", "4M", 30, xd="4S", xv=3, xw=2, xr="FR"), "counts")
    s.add_read(SimpleRead(11, "ACGT", "4M", 30, xv=3, xw=0, xr="FR"), "counts")
    s.add_read(SimpleRead(11, "ACGT", "4M", 30, xv=3, xr="RR"), "counts")
    assert [s.collapsed_count(12, t) for t in range(8)] == [1, 0, 0, 1, 0, 1, 0, 0]


@pytest.mark.parametrize("name", sorted(__import__("tests.util_counts", fromlist=["x"]).COVERAGE_VECTORS))
def test_coverage_calculator_insertion_vectors(name):
    # CoverageCalculatorTests.cs:268-398 (ComputeCoverage_Insertions), CoverageCalculator(considerAnchorInformation: true)
    from tests.util_counts import COVERAGE_VECTORS
    v = COVERAGE_VECTORS[name]
    c = ob.Caller(ob.default_config(), "chr1", "A" * 64)
    cnt = v["counts"]
    for i in range(cnt.shape[0]):
        for a in range(6):
            for d in range(3):
                for an in range(11):
                    if cnt[i, a, d, an]:
                        c.set_count(i + 1, a, d, an, int(cnt[i, a, d, an]))
    r = c.process_allele(v["type"], 1, v["ref"], v["alt"], v["support"], v["well_anchored"], coverage_only=True)
    assert tuple(r.cov) == v["expected_cov"] and r.total_coverage == sum(v["expected_cov"])
