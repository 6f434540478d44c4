// Device-side scoring arithmetic for the per-locus hot path. Compiled with --fmad=false: the reference's C# evaluates every
// operation separately in IEEE double / single precision, and several steps are deliberately mixed-precision.
//
// What each block reproduces (reference @ /root/reference):
//   incomplete gamma "Poisson.Cdf"                   src/lib/Pisces.Calculators/stats/Poisson.cs:16-128
//   MathNet.Numerics 4.5.1 GammaLowerRegularized/GammaLn/FactorialLn (NuGet dep; arithmetic taken from the IL of the shipped
//   MathNet.Numerics.dll, see oracle/tools/il_dump.py)  call sites src/lib/Pisces.Calculators/VariantQualityCalculator.cs:36-47
//   variant q-score                                  VariantQualityCalculator.cs:11-65
//   strand bias                                      src/lib/Pisces.Calculators/StrandBiasCalculator.cs:21-105,137-231
//   somatic genotype + GQ                            src/lib/Pisces.Genotyping/Somatic/SomaticGenotyper.cs:65-100, SomaticGenotypeQualityCalculator.cs:10-48
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace pb2 {

// ---- enums (values = reference ordinals)
enum : int { AT_A = 0, AT_G = 1, AT_C = 2, AT_T = 3, AT_N = 4, AT_DEL = 5 };
enum : int { DIR_F = 0, DIR_R = 1, DIR_S = 2 };
enum : int { CAT_SNV = 0, CAT_INS = 1, CAT_DEL = 2, CAT_MNV = 3, CAT_REF = 4 };
enum : int { FLT_STRAND_BIAS = 0, FLT_POOL_BIAS, FLT_AMPLICON_BIAS, FLT_LOW_VQ, FLT_LOW_DEPTH, FLT_LOW_VF, FLT_LOW_GQ, FLT_INDEL_REPEAT,
             FLT_MULTI_ALLELIC, FLT_RMXN, FLT_FORCED_REPORT, FLT_OFF_TARGET, FLT_NO_CALL };
enum : int { GT_HET_ALT12 = 0, GT_ALT12_NOCALL, GT_HET_ALT_REF, GT_HOM_ALT, GT_HOM_REF, GT_REF_NOCALL, GT_ALT_NOCALL, GT_REF_AND_NOCALL,
             GT_ALT_AND_NOCALL, GT_HEMI_REF, GT_HEMI_ALT, GT_HEMI_NOCALL, GT_OTHERS };
enum : int { SBM_POISSON = 0, SBM_EXTENDED = 1, SBM_DIPLOID = 2 };

// ---------------------------------------------------------------- Pisces' own regularised upper incomplete gamma (Poisson.cs)
__device__ __forceinline__ double pisces_lngamma(double a) {
    if (a >= 700.0)  // LanczCutoff -> Stirling (:125-128)
        return 0.5 * log(2.0 * 3.141592653589793) + (0.5 + a) * log(a) - a;
    double tmp = a + 5.5;  // Lanczos (:106-120)
    tmp = tmp - (a + 0.5) * log(tmp);
    double ser = 1.000000000190015 + 76.18009172947146 / (a + 1.0);
    ser -= 86.50532032941678 / (a + 2.0);
    ser += 24.01409824083091 / (a + 3.0);
    ser -= 1.231739572450155 / (a + 4.0);
    ser += 0.001208650973866179 / (a + 5.0);
    ser -= 5.395239384953E-06 / (a + 6.0);
    return log(2.506628274631001 * ser / a) - tmp;
}

// Poisson.Cdf(numOccurrences, lambda) = IncompleteGammaFunction((int)(numOccurrences + 1.0), lambda)   (:26-44)
static __device__ __noinline__ double pisces_poisson_cdf(double num_occurrences, double x) {
    const double a = (double)(int)(num_occurrences + 1.0);
    if ((x < 0) || (a <= 0)) return -1.0;
    const double g = pisces_lngamma(a);
    if (x >= a + 1.0) {  // Lentz continued fraction (:49-74), Itmax 300, Epsilon 1e-20, Fpmin 1e-50
        double b = x + 1.0 - a;
        double c = 1.0 / 1.0E-50;
        double d = 1.0 / b;
        double h = d;
        int i;
        for (i = 1; i <= 300; i++) {
            double an = i * (a - i);
            b += 2.0;
            d = an * d + b;
            if (fabs(d) < 1.0E-50) d = 1.0E-50;
            c = b + an / c;
            if (fabs(c) < 1.0E-50) c = 1.0E-50;
            d = 1.0 / d;
            double del = d * c;
            h *= del;
            if (fabs(del - 1.0) < 1.0E-20) break;
        }
        if (i > 300) return -1.0;
        return exp(a * log(x) - x - g) * h;
    }
    // series (:76-101)
    if (x == 0.0) return 1.0 - 0.0;
    double ap = a;
    double sum = 1.0 / a;
    double del = sum;
    double gs = -1.0;
    for (int i = 1; i <= 300; i++) {
        ap += 1.0;
        del *= x / ap;
        sum += del;
        if (fabs(del) < fabs(sum) * 1.0E-20) {
            gs = sum * exp(a * log(x) - x - g);
            break;
        }
    }
    if (gs < 0) return gs;
    return 1.0 - gs;
}

// ---------------------------------------------------------------- MathNet.Numerics 4.5.1 (IL of the shipped dll)
__host__ __device__ __forceinline__ double mathnet_gamma_ln(double z) {  // SpecialFunctions::GammaLn, z >= 0.5 branch only (callers pass z >= 1)
    const double dk[11] = {2.4857408913875355e-05, 1.0514237858172197,   -3.4568709722201625,  4.512277094668948,
                           -2.9828522532357664,    1.056397115771267,    -0.19542877319164587, 0.01709705434044412,
                           -0.0005719261174043057, 4.633994733599057e-06, -2.7199490848860772e-09};
    double s = dk[0];
#pragma unroll
    for (int i = 1; i <= 10; i++) s += dk[i] / (z + (double)i - 1.0);
    return log(s) + 0.6207822376352452 + ((z - 0.5) * log((z - 0.5 + 10.900511) / 2.718281828459045));
}

__host__ __device__ __forceinline__ double mathnet_factorial_ln(int x) {  // SpecialFunctions::FactorialLn; cache f[i] = f[i-1]*i, 171 entries
    if (x <= 1) return 0.0;
    if (x < 171) {
        double f = 1.0;
        for (int i = 2; i <= x; i++) f = f * (double)i;  // same left-to-right product as the cache
        return log(f);
    }
    return mathnet_gamma_ln((double)x + 1.0);
}

static __device__ __noinline__ double mathnet_gamma_lower_regularized(double a, double x) {  // SpecialFunctions::GammaLowerRegularized
    if (fabs(a) < 10 * 1.1102230246251565e-16) return 1.0;
    if (fabs(x) < 10 * 1.1102230246251565e-16) return 0.0;
    const double ax = (a * log(x)) - x - mathnet_gamma_ln(a);
    if (ax < -709.782712893384) return a < x ? 1.0 : 0.0;
    if (x <= 1 || x <= a) {
        double r2 = a, c2 = 1, ans2 = 1;
        do {
            r2 = r2 + 1;
            c2 = c2 * x / r2;
            ans2 += c2;
        } while ((c2 / ans2) > 1e-15);
        return exp(ax) * ans2 / a;
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
            error = fabs((ans - nextans) / nextans);
            ans = nextans;
        } else {
            error = 1;
        }
        p3 = p2; p2 = p; q3 = q2; q2 = q;
        if (fabs(p) > 4503599627370496.0) {
            p3 *= 2.220446049250313e-16; p2 *= 2.220446049250313e-16; q3 *= 2.220446049250313e-16; q2 *= 2.220446049250313e-16;
        }
    } while (error > 1e-15);
    return 1.0 - (exp(ax) * ans);
}

// ---------------------------------------------------------------- VariantQualityCalculator.cs
// MathOperations.QtoP(double q) = Math.Pow(10, -1 * q / 10f): q is double there, so the division is in double (MathOperations.cs:7-10)
__device__ __forceinline__ double q_to_p(double q) { return pow(10.0, -1 * q / 10.0); }

// error_rate = MathOperations.QtoP(noise level): a per-run constant under the Flat noise model, computed once on the host (pb2_api.cu) with the
// same expression; the Window noise model passes q_to_p(NL) per allele.
__device__ __forceinline__ double raw_poisson_qscore(int call_count, int coverage, double error_rate) {  // :27-52
    const double k_minus_one = call_count - 1;
    const double k = call_count;
    const double lambda = error_rate * coverage;
    // Poisson(lambda).CumulativeDistribution(k-1) = 1 - GammaLowerRegularized(k, lambda); keep the double cancellation
    const double cdf = 1.0 - mathnet_gamma_lower_regularized(k_minus_one + 1.0, lambda);
    const double p_value = 1 - cdf;
    if (p_value > 0) return -10 * log10(p_value);
    const double A = -lambda + (double)(int)k_minus_one * log(lambda) - mathnet_factorial_ln((int)k_minus_one);  // ProbabilityLn
    const double correction = (k - lambda) / k;
    return -10.0 * (A - log(2.0 * correction)) / log(10.0);
}
// Rigorous short-circuit of the q-score cap: returns true when min(maxQ, raw) is provably maxQ, without evaluating the incomplete gamma.
//  * p = P(X >= k) <= exp(-lambda + k - k ln(k/lambda)) (Chernoff, k > lambda). The reference's p-value is 1 - (1 - P_lower): p rounded to a
//    multiple of 2^-53; if the bound is below 10^(-maxQ/10) / 4 (and that threshold is far above 2^-53, i.e. maxQ <= 150) the p > 0 branch gives
//    raw >= maxQ;
//  * if the p-value rounds to 0 the reference uses -10 (A - ln(2(k-lambda)/k)) / ln 10 with A = ln PMF(k-1); ln Gamma(k) >= (k-1/2) ln k - k +
//    ln sqrt(2 pi) bounds A from above, hence that branch from below.
// Both branches >= maxQ  =>  (int)Round(Max(0, Min(maxQ, raw))) == maxQ exactly as VariantQualityCalculator.cs:54-65 computes it.
__device__ __forceinline__ bool qscore_is_capped(int k, int n, double error_rate, int max_q) {
    if (max_q > 150 || max_q < 0) return false;
    const double lambda = error_rate * n;
    const double kd = k;
    if (!(kd > lambda) || !(lambda > 0)) return false;
    const double need0 = -(double)max_q * 0.23025850929940458 - 1.0;   // = `need` below
    // log-free pre-test (the common case of a well-supported allele): with k >= 8 lambda, ln(k/lambda) >= ln 8 bounds both expressions below from
    // above: ln_chernoff <= k (1 - ln 8), and a_ub - corr <= k (1 - ln 8) + ln 8 - 0.9189 - ln(2 * 7/8) <= k (1 - ln 8) + 0.61
    if (kd >= 8.0 * lambda && kd * -1.0794415416798357 + 0.61 <= need0 - 1.3862943611198906) return true;
    // single-precision logs are enough for a bound: their error (<= 2e-7 relative) times k <= 65535 is below 0.2; 1.0 of slack is charged
    const double lnk = (double)logf((float)kd), lnl = (double)logf((float)lambda);
    const double need = -(double)max_q * 0.23025850929940458 - 1.0;          // ln(10^(-maxQ/10)) minus the slack
    const double ln_chernoff = -lambda + kd - kd * (lnk - lnl);
    if (!(ln_chernoff <= need - 1.3862943611198906)) return false;           // bound <= T/4
    const double a_ub = -lambda + (kd - 1.0) * lnl - ((kd - 0.5) * lnk - kd + 0.9189385332046727);
    const double corr = (double)logf((float)(2.0 * ((kd - lambda) / kd)));
    return (a_ub - corr) <= need;
}

__device__ __forceinline__ int poisson_qscore(int call_count, int coverage, double error_rate, int max_q) {  // :54-65
    if ((call_count <= 0) || (coverage <= 0)) return 0;
    if (qscore_is_capped(call_count, coverage, error_rate, max_q)) return max_q;
    double q = fmin((double)max_q, raw_poisson_qscore(call_count, coverage, error_rate));
    q = fmax(q, 0.0);
    return (int)rint(q);  // Math.Round: half to even
}

// ---------------------------------------------------------------- StrandBiasCalculator.cs
struct SbStats { double fn, fp, vg, coverage, support; };

// MathNet SpecialFunctions::BetaRegularized (IL of the shipped dll): Lentz continued fraction, eps = 2^-53, fpmin = 4.94e-324 / eps
static __device__ __noinline__ double mathnet_beta_regularized(double a, double b, double x) {
    const double bt = (x == 0.0 || x == 1.0) ? 0.0 : exp(mathnet_gamma_ln(a + b) - mathnet_gamma_ln(a) - mathnet_gamma_ln(b) + a * log(x) + b * log(1.0 - x));
    const bool symmetry = x >= (a + 1.0) / (a + b + 2.0);
    const double eps = 1.1102230246251565e-16;
    const double fpmin = 4.9406564584124654e-324 / eps;
    if (symmetry) { x = 1.0 - x; const double t = a; a = b; b = t; }
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0;
    double d = 1.0 - qab * x / qap;
    if (fabs(d) < fpmin) d = fpmin;
    d = 1.0 / d;
    double h = d;
    for (int m = 1, m2 = 2; m <= 50000; m++, m2 += 2) {
        double aa = (double)m * (b - (double)m) * x / ((qam + (double)m2) * (a + (double)m2));
        d = 1.0 + aa * d;
        if (fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa / c;
        if (fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + (double)m) * (qab + (double)m) * x / ((a + (double)m2) * (qap + (double)m2));
        d = 1.0 + aa * d;
        if (fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa / c;
        if (fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) <= eps) break;
    }
    return symmetry ? 1.0 - bt * h / a : bt * h / a;
}
// MathNet Distributions.Binomial(p, n).CumulativeDistribution(x)
__device__ __forceinline__ double mathnet_binomial_cdf(double p, int n, double x) {
    if (x < 0.0) return 0.0;
    if (x > (double)n) return 1.0;
    const double k = floor(x);
    return mathnet_beta_regularized((double)n - k, k + 1.0, 1.0 - p);
}

// min_vf: _config.MinFrequency as a double (only the Diploid model reads it: CreateStats :137-148)
__device__ __forceinline__ SbStats sb_create_stats(double support, double coverage, double noise, int model, double min_vf) {  // :137-148,150-231
    SbStats s;
    s.support = support;
    s.coverage = coverage;
    const double min_detectable = model == SBM_DIPLOID ? min_vf : noise;  // model != Diploid: minDetectableSNP = noiseFreq
    if (support == 0) {
        if (model == SBM_POISSON) { s.fp = 1; s.vg = 0; s.fn = 0; }
        else { s.vg = pow(1 - min_detectable, coverage); s.fp = 1 - s.vg; s.fn = s.vg; }
    } else if (model == SBM_DIPLOID) {   // PopulateDiploidStats :150-173
        if (support / coverage >= min_detectable) { s.fn = 1; s.fp = 0; s.vg = 1; }
        else {
            s.fn = fmax(mathnet_binomial_cdf(min_detectable, (int)coverage, support), 0.0);
            s.fp = fmax(0.0, 1 - pisces_poisson_cdf(support, coverage * 0.1));
            s.vg = s.fn;
        }
    } else {
        // Rigorous short-circuit: Cdf(support-1, x) = 1 - gs with gs = sum * exp(a ln x - x - lnGamma~(a)), a = support. For x <= a/2 the series
        // converges (ratio <= 1/2), sum <= 2, and lnGamma~(a) >= (a-1/2) ln a - a + ln sqrt(2 pi) - 1e-6 for both of the reference's
        // approximations (Poisson.cs:106-128); when that bounds gs below 2^-54 the double subtraction yields exactly 1.0.
        const double x = coverage * noise;
        bool saturated = false;
        if (x > 0 && x <= 0.125 * support && support >= 45.0) {
            // log-free: ln(support / x) >= ln 8 gives e_ub <= support (1 - ln 8) + ln(support) / 2 - 0.9189, which decreases in support and is -47.6 at 45
            saturated = true;
        } else if (x > 0 && x <= 0.5 * support) {
            // single-precision logs + 1.0 of slack (see qscore_is_capped)
            const double e_ub = support * (double)logf((float)x) - x - ((support - 0.5) * (double)logf((float)support) - support + 0.9189385332046727);
            saturated = e_ub < -41.0;
        }
        s.vg = saturated ? 1.0 : fmax(0.0, pisces_poisson_cdf(support - 1, x));
        s.fp = fmax(0.0, 1 - s.vg);
        s.fn = 0.0;  // ChanceFalseNeg feeds nothing downstream of the record (StrandBiasStats only); computed on demand by the stats API
    }
    return s;
}

struct SbResult { double bias, gatk; bool acceptable, var_both, cov_both; };

// noise = Math.Pow(10, -1*qNoise/10f) (float exponent, :32): a per-run constant, computed once on the host
// The three sets of statistics of StrandBiasCalculator.CalculateStrandBiasResults (:21-72): which = 0 overall, 1 forward, 2 reverse (stitched reads
// count half to each strand, integer halves as the reference's int arithmetic gives them)
__device__ __forceinline__ void sb_inputs_of(int which, const int cov[3], const int sup[3], int& s, int& c) {
    s = which == 0 ? sup[0] + sup[1] + sup[2] : (which == 1 ? sup[0] : sup[1]) + sup[2] / 2;
    c = which == 0 ? cov[0] + cov[1] + cov[2] : (which == 1 ? cov[0] : cov[1]) + cov[2] / 2;
}
__device__ __forceinline__ SbStats sb_stats_of(int which, const int cov[3], const int sup[3], double noise, int model, double min_vf) {
    // one call site: threads that compute different sets run the same instructions
    int s, c;
    sb_inputs_of(which, cov, sup, s, c);
    return sb_create_stats(s, c, noise, model, min_vf);
}
__device__ __forceinline__ SbResult strand_bias_combine(const SbStats& o, const SbStats& f, const SbStats& r, double acceptance) {  // :40-72,89-105
    double fb = (f.vg * r.fp) / o.vg;
    double rb = (r.vg * f.fp) / o.vg;
    if (o.vg == 0) { fb = 1; rb = 1; }
    SbResult res;
    double p = (isnan(fb) || isnan(rb)) ? nan("") : fmax(fb, rb);  // Math.Max propagates NaN
    res.bias = p;
    res.gatk = (p == 0) ? -INFINITY : 10 * log10(p);   // 10*log10(0) is -inf either way
    res.cov_both = (f.coverage > 0) && (r.coverage > 0);
    res.var_both = (f.support > 0) && (r.support > 0);
    if (!res.cov_both) { res.bias = 0; res.gatk = -INFINITY; }
    res.acceptable = res.bias < acceptance;
    return res;
}
// noise = Math.Pow(10, -1*qNoise/10f) (float exponent, :32): a per-run constant, computed once on the host
__device__ __forceinline__ SbResult strand_bias(const int cov[3], const int sup[3], double noise, double acceptance, int model, double min_vf) {  // :21-72,89-105
    const SbStats o = sb_stats_of(0, cov, sup, noise, model, min_vf);
    const SbStats f = sb_stats_of(1, cov, sup, noise, model, min_vf);
    const SbStats r = sb_stats_of(2, cov, sup, noise, model, min_vf);
    return strand_bias_combine(o, f, r, acceptance);
}

// ---------------------------------------------------------------- CalledAllele.Frequency / RefFrequency (CalledAllele.cs:49-52,123-126): float
__device__ __forceinline__ float allele_frequency(int support, int total) {
    if (total == 0) return 0.0f;
    return fminf((float)support / (float)total, 1.0f);
}

// ---------------------------------------------------------------- SomaticGenotyper.cs:65-100
__device__ __forceinline__ int somatic_genotype(bool is_ref, int total_cov, float freq, float ref_freq, float min_freq_filter, int min_depth) {
    if (total_cov < min_depth) return is_ref ? GT_REF_NOCALL : GT_ALT_NOCALL;
    if (!is_ref) {
        if (ref_freq < min_freq_filter) {
            if ((1 - freq) > min_freq_filter) return GT_ALT_AND_NOCALL;
            return GT_HOM_ALT;
        }
        return GT_HET_ALT_REF;
    }
    if (freq < min_freq_filter) return GT_REF_NOCALL;
    if ((1 - freq) > min_freq_filter) return GT_REF_AND_NOCALL;
    return GT_HOM_REF;
}
// SomaticGenotypeQualityCalculator.cs:10-48
// q_to_p_table: optional device table of QtoP(q) for q = 0..table_max (filled on the host with the same expression), nullptr -> pow on the device
// gq_tail_table: optional device table of the Poisson tail below, Cdf((int)(nonAlleleObservations + 1) - 1, targetLOD * coverage), indexed
// [coverage][(int)(nonAlleleObservations + 1)] for coverage < kGqTailMaxCov and a < kGqTailMaxA: the value depends on nothing else (Poisson.Cdf
// truncates its first argument, Poisson.cs:26-44), and gq_tail_fill_kernel fills it with this very function, so a lookup returns the same double.
// Behind the doubles sits a second table of the same shape: the finished GQ (an int32) of a variant q-score equal to the cap, max_variant_qscore -
// what nearly every confidently homozygous locus has - computed by the fill kernel with the statements below.
constexpr int kGqTailMaxCov = 8192, kGqTailMaxA = 32;
__device__ __forceinline__ int somatic_gq(int genotype, int vq, int total_cov, float freq, float target_lod, int min_gq, int max_gq,
                                          const double* __restrict__ q_to_p_table = nullptr, int table_max = -1,
                                          const double* __restrict__ gq_tail_table = nullptr, int capped_vq = -1) {
    double raw = vq;
    const bool nocall = genotype == GT_ALT12_NOCALL || genotype == GT_ALT_NOCALL || genotype == GT_REF_NOCALL;
    if (total_cov == 0 || nocall) return min_gq;
    if (genotype == GT_HOM_REF || genotype == GT_HOM_ALT) {
        const double p1 = (q_to_p_table != nullptr && vq >= 0 && vq <= table_max) ? q_to_p_table[vq] : q_to_p((double)vq);
        const float non_allele_obs = (1.0f - freq) * (float)total_cov;
        const float expected = target_lod * (float)total_cov;
        if (non_allele_obs >= expected) return min_gq;
        const int a_key = (int)((double)non_allele_obs + 1.0);
        if (gq_tail_table != nullptr && vq == capped_vq && total_cov < kGqTailMaxCov && a_key >= 1 && a_key < kGqTailMaxA)
            return reinterpret_cast<const int*>(gq_tail_table + kGqTailMaxCov * kGqTailMaxA)[total_cov * kGqTailMaxA + a_key];
        const double p2 = (gq_tail_table != nullptr && total_cov < kGqTailMaxCov && a_key >= 1 && a_key < kGqTailMaxA)
                              ? gq_tail_table[total_cov * kGqTailMaxA + a_key]
                              : pisces_poisson_cdf((double)non_allele_obs, (double)expected);
        raw = -10 * log10(p1 + p2);
    }
    double q = fmin((double)max_gq, raw);
    q = fmax(q, (double)min_gq);
    return (int)rint(q);
}

// ---------------------------------------------------------------- germline genotypers; shared by the kernels (reference-only loci)
// and the per-locus pass of pb2_flush (loci with variant alleles)
enum : int { PLOIDY_SOMATIC = 0, PLOIDY_DIPLOID = 1, PLOIDY_ADAPTIVE = 2, PLOIDY_HAPLOID = 3 };

// MathNet Distributions.Poisson(lambda).ProbabilityLn(k) and Binomial(p, n).ProbabilityLn(k) (PMFLn, IL of the shipped dll)
__host__ __device__ __forceinline__ double mathnet_poisson_probability_ln(double lambda, int k) { return -lambda + (double)k * log(lambda) - mathnet_factorial_ln(k); }
__host__ __device__ __forceinline__ double mathnet_binomial_probability_ln(double p, int n, int k) {
    if (k < 0 || k > n) return -INFINITY;
    if (p == 0.0) return k == 0 ? 0.0 : -INFINITY;
    if (p == 1.0) return k == n ? 0.0 : -INFINITY;
    const double binomial_ln = mathnet_factorial_ln(n) - mathnet_factorial_ln(k) - mathnet_factorial_ln(n - k);
    return binomial_ln + (double)k * log(p) + (double)(n - k) * log(1.0 - p);
}
// (int) of a double in C# (unchecked, x64): NaN / out of range -> int.MinValue
__host__ __device__ __forceinline__ int cs_int_cast(double v) {
    if (!(v < 2147483648.0) || !(v > -2147483649.0)) return (int)0x80000000;
    return (int)v;
}
// DiploidGenotypeQualityCalculator.Compute (Thresholding/DiploidGenotypeQualityCalculator.cs:17-103) and HaploidGenotypeQualityCalculator.Compute
// (Haploid/HaploidGenotypeQualityCalculator.cs:12-59): float parameters, MathNet log-probabilities in double
__host__ __device__ inline int germline_gq(bool haploid, int genotype, int total_cov, int allele_support, int min_q, int max_q) {
    if (total_cov == 0) return min_q;
    const float noise_hom_ref = 0.05f, noise_hom_alt = 0.075f, noise_het_alt = 0.10f, expected_het = 0.40f;
    const float depth = (float)total_cov;
    const double lam_hom_ref = (double)(noise_hom_ref * depth), lam_hom_alt = (double)(noise_hom_alt * depth);
    const int non_allele = total_cov - allele_support > 0 ? total_cov - allele_support : 0;
    double h0 = 0, h1 = 0;
    if (genotype == (haploid ? GT_HEMI_REF : GT_HOM_REF)) {
        h0 = mathnet_poisson_probability_ln(lam_hom_ref, non_allele);
        h1 = mathnet_binomial_probability_ln((double)expected_het, total_cov, non_allele);
    } else if (genotype == (haploid ? GT_HEMI_ALT : GT_HOM_ALT)) {
        h0 = mathnet_poisson_probability_ln(lam_hom_alt, non_allele);
        h1 = mathnet_binomial_probability_ln((double)expected_het, total_cov, allele_support);
    } else if (!haploid && (genotype == GT_HET_ALT12 || genotype == GT_HET_ALT_REF)) {
        const float freq = total_cov == 0 ? 0.0f : fminf((float)allele_support / (float)total_cov, 1.0f);
        const int k = (int)(depth * freq);
        h0 = mathnet_binomial_probability_ln((double)expected_het, total_cov, k);
        h1 = ((double)freq >= 0.50) ? mathnet_binomial_probability_ln((double)(1 - noise_het_alt), total_cov, k) : mathnet_binomial_probability_ln((double)noise_het_alt, total_cov, k);
    } else {
        return min_q;
    }
    const int q = cs_int_cast(floor(10.0 * 0.4342944819032518 * (h0 - h1)));   // 10.0 * Math.Log10(Math.E) * (...)
    if (!haploid) {
        if ((h1 <= -2147483648.0) && (h0 > h1)) return max_q;
        if ((h0 <= -2147483648.0) && (h0 < h1)) return min_q;
    }
    const int lo = q < max_q ? q : max_q;
    return lo > min_q ? lo : min_q;
}
// The genotype of a locus whose only allele is its reference allele: DiploidThresholdingGenotyper.CalculateDiploidGenotype
// (DiploidThresholdingGenotyper.cs:77-100, GenotypeCalculatorUtilities.ConvertSimpleGenotypeToComplexGenotype :160-190) /
// HaploidGenotyper.CalculateHaploidGenotype (HaploidGenotyper.cs:54-82) with alleles = [reference]
__host__ __device__ inline int germline_reference_only_genotype(bool haploid, int total_cov, int allele_support, int ref_support, float minor_vf, float major_vf, int min_depth) {
    const float ref_frequency = total_cov == 0 ? 0.0f : fminf((float)ref_support / (float)total_cov, 1.0f);   // alleles.First().RefFrequency (:93-94)
    const float frequency = total_cov == 0 ? 0.0f : fminf((float)allele_support / (float)total_cov, 1.0f);
    const bool ref_exists = (double)ref_frequency >= (double)minor_vf;
    const bool depth_issue = total_cov < min_depth;
    if (haploid) return (!depth_issue && ref_exists && (double)ref_frequency > (double)major_vf) ? GT_HEMI_REF : GT_HEMI_NOCALL;
    if (depth_issue || !ref_exists) return GT_REF_NOCALL;
    if ((1 - frequency) > minor_vf) return GT_REF_AND_NOCALL;
    return GT_HOM_REF;
}

}  // namespace pb2
