// ORACLE — TEST INFRASTRUCTURE ONLY. Not shipped, not linked into libpisces_b200.so, never on the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
//
// CPU restatement of the arithmetic the reference's per-locus scoring uses.
//   * Pisces' own incomplete gamma:   /root/reference/src/lib/Pisces.Calculators/stats/Poisson.cs:16-128
//   * MathOperations:                 /root/reference/src/lib/Pisces.Calculators/stats/MathOperations.cs:7-30
//   * MathNet.Numerics 4.5.1 (NuGet dependency pinned at Pisces.Calculators.csproj:24; source NOT in /root/reference).
//     Restated from the IL of the MathNet.Numerics.dll shipped in binaries/5.2.11.163/Pisces_5.2.11.163.tar.gz,
//     disassembled with oracle/tools/il_dump.py (SpecialFunctions::GammaLowerRegularized / GammaLn / FactorialLn,
//     Distributions.Poisson::CumulativeDistribution / ProbabilityLn, Binomial::CDF, SpecialFunctions::BetaRegularized).
//     Operation order, constants (1e-15, 4503599627370496, 2^-52, -709.782712893384, GammaR 10.900511, the 11 GammaDk
//     doubles read from the FieldRVA blob, factorial cache length 171) are those of the IL.
// Parity pin: reference KATs in tests/test_oracle_kats.py (QualityCalculatorTests.cs:62-95, PoissonTests.cs:77-103, ...).
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace po {

// ---------------------------------------------------------------- Pisces.Calculators.Poisson (Poisson.cs)
namespace pisces_poisson {
constexpr double Epsilon = 1.0E-20;      // Poisson.cs:16
constexpr double Fpmin = 1.0E-50;        // :17
constexpr double LanczCutoff = 700.0;    // :18
constexpr int Itmax = 300;               // :19

inline double LanczosApproximation(double p) {   // Poisson.cs:106-120
    double x = p;
    double tmp = x + 5.5;
    tmp = tmp - (x + 0.5) * std::log(tmp);
    double ser = 1.000000000190015 + 76.18009172947146 / (p + 1.0);
    ser -= 86.50532032941678 / (p + 2.0);
    ser += 24.01409824083091 / (p + 3.0);
    ser -= 1.231739572450155 / (p + 4.0);
    ser += 0.001208650973866179 / (p + 5.0);
    ser -= 5.395239384953E-06 / (p + 6.0);
    return (std::log(2.506628274631001 * ser / x) - tmp);
}
inline double StirlingApproximation(double n) {  // Poisson.cs:125-128
    return (0.5 * std::log(2.0 * M_PI) + (0.5 + n) * std::log(n) - n);
}
inline double GammaUsingContinuedFractions(double a, double x, double g) {  // Poisson.cs:49-74
    double b = x + 1.0 - a;
    double c = 1.0 / Fpmin;
    double d = 1.0 / b;
    double h = d;
    int i;
    for (i = 1; i <= Itmax; i++) {
        double an = i * (a - i);
        b += 2.0;
        d = an * d + b;
        if (std::fabs(d) < Fpmin) d = Fpmin;
        c = b + an / c;
        if (std::fabs(c) < Fpmin) c = Fpmin;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (std::fabs(del - 1.0) < Epsilon) break;
    }
    if (i > Itmax) return -1.0;
    return std::exp(a * std::log(x) - x - g) * h;
}
inline double GammaSeries(double a, double x, double g) {  // Poisson.cs:76-101
    double retval = -1.0;
    if (x == 0.0) return 0.0;
    if (x < 0.0) return retval;
    double ap = a;
    double sum = 1.0 / a;
    double del = sum;
    for (int i = 1; i <= Itmax; i++) {
        ap += 1.0;
        del *= x / ap;
        sum += del;
        if (std::fabs(del) < std::fabs(sum) * Epsilon) {
            retval = sum * std::exp(a * std::log(x) - x - g);
            break;
        }
    }
    return retval;
}
inline double IncompleteGammaFunction(double a, double x) {  // Poisson.cs:34-44
    if ((x < 0) || (a <= 0)) return -1.0;
    double g = (a >= LanczCutoff ? StirlingApproximation(a) : LanczosApproximation(a));
    if (x >= a + 1.0) return GammaUsingContinuedFractions(a, x, g);
    if ((g = GammaSeries(a, x, g)) < 0) return g;
    return 1.0 - g;
}
// Poisson.Cdf(numOccurrences, numExpectedOccurrences)  Poisson.cs:26-29
inline double Cdf(double numOccurrences, double numExpectedOccurrences) {
    return IncompleteGammaFunction((double)(int)(numOccurrences + 1.0), numExpectedOccurrences);
}
}  // namespace pisces_poisson

// ---------------------------------------------------------------- MathOperations.cs
inline double QtoP(double q) { return std::pow(10.0, -1 * q / (double)10.0f); }          // :7-10  (double / (double)10f)
inline double PtoQ(double p) { return (-10 * std::log10(p)); }                             // :12-15
inline double PtoGATKBiasScale(double p) { return 10 * std::log10(p); }                   // :25-28

// ---------------------------------------------------------------- MathNet.Numerics 4.5.1 (from IL, see header)
namespace mathnet {
static const double GammaDk[11] = {
    2.4857408913875355e-05, 1.0514237858172197,   -3.4568709722201625, 4.512277094668948,
    -2.9828522532357664,    1.056397115771267,    -0.19542877319164587, 0.01709705434044412,
    -0.0005719261174043057, 4.633994733599057e-06, -2.7199490848860772e-09};
constexpr double GammaR = 10.900511;
constexpr double LogTwoSqrtEOverPi = 0.6207822376352452;
constexpr double LnPi = 1.1447298858494002;

inline double GammaLn(double z) {  // SpecialFunctions::GammaLn
    if (z < 0.5) {
        double s = GammaDk[0];
        for (int i = 1; i <= 10; i++) s += GammaDk[i] / ((double)i - z);
        return LnPi - std::log(std::sin(3.141592653589793 * z)) - std::log(s) - LogTwoSqrtEOverPi -
               ((0.5 - z) * std::log((0.5 - z + GammaR) / 2.718281828459045));
    }
    double s = GammaDk[0];
    for (int i = 1; i <= 10; i++) s += GammaDk[i] / (z + (double)i - 1.0);
    return std::log(s) + LogTwoSqrtEOverPi + ((z - 0.5) * std::log((z - 0.5 + GammaR) / 2.718281828459045));
}

struct FactorialCache {
    double f[171];
    FactorialCache() { f[0] = 1.0; for (int i = 1; i < 171; i++) f[i] = f[i - 1] * (double)i; }
};
inline double FactorialLn(int x) {  // SpecialFunctions::FactorialLn (x < 0 throws in MathNet; callers never pass it)
    static const FactorialCache cache;
    if (x <= 1) return 0.0;
    if (x < 171) return std::log(cache.f[x]);
    return GammaLn((double)x + 1.0);
}

// Precision::AlmostEqual(a, 0.0) for doubles in MathNet 4.x: |a-b| < 10*2^-53 (DefaultDoubleAccuracy) unless inf/nan.
inline bool AlmostEqualZero(double a) {
    if (std::isnan(a) || std::isinf(a)) return false;
    return std::fabs(a) < 10 * 1.1102230246251565e-16;
}

inline double GammaLowerRegularized(double a, double x) {  // SpecialFunctions::GammaLowerRegularized
    const double epsilon = 1e-15, big = 4503599627370496.0, bigInv = 2.220446049250313e-16;
    if (AlmostEqualZero(a)) return 1.0;
    if (AlmostEqualZero(x)) return 0.0;
    double ax = (a * std::log(x)) - x - GammaLn(a);
    if (ax < -709.782712893384) return a < x ? 1.0 : 0.0;
    if (x <= 1 || x <= a) {
        double r2 = a, c2 = 1, ans2 = 1;
        do {
            r2 = r2 + 1;
            c2 = c2 * x / r2;
            ans2 += c2;
        } while ((c2 / ans2) > epsilon);
        return std::exp(ax) * ans2 / a;
    }
    int c = 0;
    double y = 1 - a;
    double z = x + y + 1;
    double p3 = 1, q3 = x, p2 = x + 1, q2 = z * x;
    double ans = p2 / q2;
    double error;
    do {
        c++;
        y += 1;
        z += 2;
        double yc = y * c;
        double p = (p2 * z) - (p3 * yc);
        double q = (q2 * z) - (q3 * yc);
        if (q != 0) {
            double nextans = p / q;
            error = std::fabs((ans - nextans) / nextans);
            ans = nextans;
        } else {
            error = 1;
        }
        p3 = p2; p2 = p; q3 = q2; q2 = q;
        if (std::fabs(p) > big) { p3 *= bigInv; p2 *= bigInv; q3 *= bigInv; q2 *= bigInv; }
    } while (error > epsilon);
    return 1.0 - (std::exp(ax) * ans);
}
// Distributions.Poisson(lambda).CumulativeDistribution(x) and .ProbabilityLn(k)
inline double PoissonCumulativeDistribution(double lambda, double x) { return 1.0 - GammaLowerRegularized(x + 1.0, lambda); }
inline double PoissonProbabilityLn(double lambda, int k) { return -lambda + (double)k * std::log(lambda) - FactorialLn(k); }
}  // namespace mathnet

}  // namespace po
