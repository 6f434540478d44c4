"""The checker's restatement of AmpliconBiasCalculator.CalculateAmpliconBias (oracle/po_calc.hpp) against the reference's own unit tests
(src/test/Pisces.Calculators.Tests/UnitTests/AmpliconBiasCalculatorTests.cs). The product does not build the amplicon-bias filter yet (SURVEY 8a row
a18 / 8f rank 4): this pins the oracle for it. CPU only."""
import numpy as np

from oracle import binding as ob

AMP1, AMP2 = 0, 1


def _execute(support, coverage, expect_null=False):
    """ExecuteTest (:203-231): Compute(variant, 100, 0.01F) on the data and on its reverse, which must agree."""
    r1 = ob.amplicon_bias(support, coverage, 0.01, 100)
    rev = lambda p: (None if p[0] is None else list(p[0])[::-1], list(p[1])[::-1])   # noqa: E731
    r2 = ob.amplicon_bias(rev(support), rev(coverage), 0.01, 100)
    if expect_null:
        assert r1 is None and r2 is None
        return None
    assert r1["bias_detected"] == r2["bias_detected"]
    return r1


def _two_amp(fa, da, fb, db, biased):   # ExecuteTwoAmpTest (:172-184): (int)(float * int) in C# is a float product
    sup = ([AMP1, AMP2], [int(np.float32(fa) * np.float32(da)), int(np.float32(fb) * np.float32(db))])
    r = _execute(sup, ([AMP1, AMP2], [da, db]))
    assert r["bias_detected"] == biased


def test_varying_depth_with_bias():   # HappyPath_VaryingDepthWithBias (:15-43)
    for amp2_depth in range(1000):
        r = _execute(([AMP1, AMP2], [int(0.05 * 1000), int(0.0 * amp2_depth)]), ([AMP1, AMP2], [1000, amp2_depth]))
        assert r["bias_detected"] == (amp2_depth >= 100), amp2_depth


def test_varying_depth_with_no_bias():   # HappyPath_VaryingDepthWithNoBias (:45-73)
    amp1_depth = 10
    while amp1_depth < 2000:
        amp1_depth += 100
        r = _execute(([AMP1, AMP2], [int(0.09 * amp1_depth), int(0.09 * 1000)]), ([AMP1, AMP2], [amp1_depth, 1000]))
        f = {int(e["name"]): e["frequency"] for e in r["per_amplicon"]}
        assert r["bias_detected"] == (not abs(f[AMP1] - f[AMP2]) < 0.05)


def test_forced_variants():   # TestAmpliconBiasCalculationsForForcedVariants (:79-83)
    _two_amp(0.0001, 500000, 0.0001, 500000, False)


def test_names_that_do_not_match_up():   # TestAmpBiasWhenAmpNamesDontMatchUp (:86-143); A, B, C, D = 0..3
    assert _execute(([1], [150]), ([0, 1], [100, 300]))["bias_detected"] is True
    _execute(([], []), ([0, 1], [100, 150]), expect_null=True)
    assert _execute(([2, 3], [100, 150]), ([0, 1], [100, 150]))["bias_detected"] is False


def test_present_on_both_amplicons():   # TestPresentOnBothStrands (:146-170; not a [Fact] in the reference, its expectations hold all the same)
    _two_amp(0.1, 500, 0.1, 500, False)
    _two_amp(0.1, 500, 0.0, 0, False)
    _two_amp(0.0, 0, 0.0, 0, False)
    _two_amp(0.0, 100, 0.0, 100, False)
    _two_amp(0.0, 0, 0.2, 500, False)
    _two_amp(0.0, 5000, 0.2, 500, True)
    _two_amp(0.001, 5000, 0.9, 500, True)
    _two_amp(0.1, 500, 0.0, 500, True)


def test_details_of_one_case():
    """0 of 300 on amplicon A, 150 of 300 on B: expected 150 on A, Poisson.Cdf(0, 150) = e^-150 -> bias; q = (int)PtoQ(1 - p) = 0."""
    r = ob.amplicon_bias(([1], [150]), ([0, 1], [300, 300]))
    a, b = r["per_amplicon"]
    assert r["artifact"] == 1 and r["bias_detected"]
    assert (a["expected_support"], a["observed_support"], a["bias_detected"]) == (150.0, 0.0, 1.0)
    assert 0.0 <= a["chance_its_real"] < 1e-60 and a["qscore"] == 0.0
    assert (b["qscore"], b["bias_detected"], b["chance_its_real"]) == (100.0, 0.0, 1.0)
