// VCF record text for the records of a flush: the somatic writer, one line per allele.
// Host code only (no kernels): text emission is the reference's host-side tail of the path.
//   VcfFileWriter.WriteListOfColocatedAlleles   src/lib/Pisces.IO/VcfFileWriter.cs:206-260 (ALT "." rule :233-244; StrandBias output switch :353-356)
//   VcfFormatter.UpdateFrequencyFormat / GetNumSigDigits   src/lib/Pisces.IO/VcfFormatter.cs:52-71
//   VcfFormatter.MapFilters / MapFilter / MapGenotype       :143-215
//   VcfFormatter FORMAT / SAMPLE columns                    :224-251, US tag :283-316, VF :329-358, DP :373-394, AD :396-420
#include <algorithm>
#include <charconv>
#include <cmath>
#include <vector>
#include <cstring>
#include <string>
#include "pb2_internal.hpp"

using namespace pb2;

namespace {

// C# float.ToString(): shortest round-trip digits, scientific ("1E-05") below 1e-4
std::string cs_float_to_string(float x) {
    char buf[64];
    if (x != 0 && std::fabs(x) < 1e-4f) {
        auto r = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);
        std::string s(buf, r.ptr);                      // e.g. "1e-05"
        const size_t e = s.find('e');
        std::string mant = s.substr(0, e), ex = s.substr(e + 1);
        const bool neg = !ex.empty() && ex[0] == '-';
        if (!ex.empty() && (ex[0] == '-' || ex[0] == '+')) ex = ex.substr(1);
        while (ex.size() < 2) ex = "0" + ex;
        return mant + "E" + (neg ? "-" : "+") + ex;
    }
    auto r = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}
int num_sig_digits(const std::string& v) {   // VcfFormatter.GetNumSigDigits (:67-71)
    const size_t e = v.find('E');
    if (e != std::string::npos) return std::abs(std::stoi(v.substr(e + 1)));
    return (int)v.size() - 1;
}
// custom numeric format "0.000...": the shortest decimal of the double, rounded half away from zero at `decimals` places
std::string fixed_half_up(double x, int decimals) {
    const bool neg = std::signbit(x) && x != 0.0;
    char buf[400];
    auto r = std::to_chars(buf, buf + sizeof(buf), std::fabs(x), std::chars_format::fixed);
    std::string s(buf, r.ptr);
    std::string ip = s, fp;
    const size_t dot = s.find('.');
    if (dot != std::string::npos) { ip = s.substr(0, dot); fp = s.substr(dot + 1); }
    while ((int)fp.size() < decimals + 1) fp.push_back('0');
    const bool up = fp[(size_t)decimals] >= '5';
    std::string digits = ip + fp.substr(0, (size_t)decimals);
    if (up) {
        int i = (int)digits.size() - 1;
        while (i >= 0) { if (digits[(size_t)i] == '9') { digits[(size_t)i] = '0'; i--; } else { digits[(size_t)i]++; break; } }
        if (i < 0) digits.insert(digits.begin(), '1');
    }
    const size_t il = digits.size() - (size_t)decimals;
    std::string out = (neg ? "-" : "") + digits.substr(0, il);
    if (decimals > 0) out += "." + digits.substr(il);
    return out;
}
const char* genotype_string(int gt) {   // VcfFormatter.MapGenotype (:184-215)
    switch (gt) {
        case GT_HOM_ALT: return "1/1";
        case GT_HOM_REF: return "0/0";
        case GT_HET_ALT_REF: return "0/1";
        case GT_HET_ALT12: return "1/2";
        case GT_REF_NOCALL: case GT_ALT_NOCALL: return "./.";
        case GT_REF_AND_NOCALL: return "0/.";
        case GT_ALT_AND_NOCALL: return "1/.";
        case GT_HEMI_ALT: return "1";
        case GT_HEMI_NOCALL: return ".";
        case GT_HEMI_REF: return "0";
        case GT_OTHERS: return "2/2";
        default: return "./.";
    }
}

}  // namespace

extern "C" int pb2_vcf_format(pb2_handle* h, const pb2_call_record* recs, const pb2_call_record_ext* ext, int64_t n, const pb2_vcf_options* opt, const char** text,
                              int64_t* len) {
    if (!h || (!recs && n > 0) || n < 0 || !text || !len) return pb2_fail(h, PB2_ERR_ARG, "pb2_vcf_format: bad argument");
    const pb2_config& c = h->cfg;
    const DeviceConfig& d = h->dcfg;
    pb2_vcf_options o;
    memset(&o, 0, sizeof(o));
    if (opt) o = *opt;
    // VF decimals: significant digits of MinimumFrequency (and of MinimumFrequencyFilter when it is the larger one, VcfFileWriter.cs:334-347)
    int vf_decimals = num_sig_digits(cs_float_to_string(c.min_frequency));
    if (c.min_frequency_filter > c.min_frequency) vf_decimals = std::max(vf_decimals, num_sig_digits(cs_float_to_string(c.min_frequency_filter)));
    const bool out_sb = o.debug_mode || o.output_bias_files || c.strand_bias_acceptance < 1;   // :353-356
    // the order AlleleProcessor.ApplyFilters (AlleleProcessor.cs:25-71), AlleleCaller (ForcedReport :108-112, LowGQ :166-170) and the genotyper
    // (MultiAllelicSite) add filters to CalledAllele.Filters
    static const int kFilterOrder[] = {FLT_LOW_DEPTH, FLT_LOW_VQ, FLT_NO_CALL, FLT_STRAND_BIAS, FLT_AMPLICON_BIAS, FLT_INDEL_REPEAT, FLT_RMXN, FLT_LOW_VF,
                                       FLT_FORCED_REPORT, FLT_MULTI_ALLELIC, FLT_LOW_GQ};
    std::string& out = h->vcf_text;
    out.clear();
    for (int64_t i = 0; i < n; i++) {
        const pb2_call_record& r = recs[i];
        const bool is_ref = r.type == CAT_REF;
        const bool forced = (r.sb_flags & 8) != 0;
        std::string ref, alt;
        if (r.ref_len + r.alt_len <= 4) {
            for (int k = 0; k < r.ref_len; k++) ref.push_back((char)((r.allele_bytes >> (8 * k)) & 0xff));
            for (int k = 0; k < r.alt_len; k++) alt.push_back((char)((r.allele_bytes >> (8 * (r.ref_len + k))) & 0xff));
        } else {
            if ((size_t)r.allele_bytes + r.ref_len + r.alt_len > h->arena.size()) return pb2_fail(h, PB2_ERR_ARG, "pb2_vcf_format: record alleles are not in the handle's arena");
            ref.assign((const char*)h->arena.data() + r.allele_bytes, r.ref_len);
            alt.assign((const char*)h->arena.data() + r.allele_bytes + r.ref_len, r.alt_len);
        }
        // GetDepthCountInt (:373-394)
        int depth = is_ref ? r.reference_support : r.reference_support + r.allele_support;
        depth = std::max(depth, r.total_coverage);
        depth = std::max(depth, r.allele_support);
        const int gt = r.genotype;
        const bool ref_like = gt == GT_HOM_REF || gt == GT_REF_NOCALL || gt == GT_REF_AND_NOCALL || gt == GT_HEMI_NOCALL || gt == GT_HEMI_REF;
        const std::string alt_col = (!forced && ref_like) ? "." : alt;   // VcfFileWriter.cs:233-244
        // FILTER (:143-182)
        std::string filter;
        std::vector<std::string> seen;
        for (int f : kFilterOrder) {
            if (!((r.filters >> f) & 1)) continue;
            std::string s;
            switch (f) {
                case FLT_LOW_VQ: s = "q" + std::to_string(d.vq_filter); break;
                case FLT_STRAND_BIAS: s = "SB"; break;
                case FLT_AMPLICON_BIAS: s = "AB"; break;
                case FLT_LOW_DEPTH: s = "LowDP"; break;
                case FLT_LOW_VF: s = "LowVariantFreq"; break;
                case FLT_LOW_GQ: s = "LowGQ"; break;
                case FLT_INDEL_REPEAT: s = "R" + std::to_string(c.indel_repeat_filter); break;
                case FLT_RMXN: s = "R" + std::to_string(c.rmxn_max_repeat_len) + "x" + std::to_string(c.rmxn_min_repetitions); break;
                case FLT_MULTI_ALLELIC: s = "MultiAllelicSite"; break;
                case FLT_FORCED_REPORT: s = "ForcedReport"; break;
                case FLT_NO_CALL: s = "NC"; break;
                default: break;
            }
            if (std::find(seen.begin(), seen.end(), s) == seen.end()) seen.push_back(s);
        }
        for (size_t k = 0; k < seen.size(); k++) filter += (k ? ";" : "") + seen[k];
        if (filter.empty()) filter = "PASS";
        // AD (:396-420), VF (:329-358)
        const std::string ad = is_ref ? std::to_string(r.allele_support) : std::to_string(r.reference_support) + "," + std::to_string(r.allele_support);
        const float freq = r.total_coverage == 0 ? 0.0f : std::min((float)r.allele_support / (float)r.total_coverage, 1.0f);
        const float vf = is_ref ? (r.total_coverage == 0 ? 0.0f : 1.0f - freq) : freq;
        std::string fmt = "GT:GQ:AD:DP:VF";
        std::string sample = std::string(genotype_string(gt)) + ":" + std::to_string(r.genotype_qscore) + ":" + ad + ":" + std::to_string(depth) + ":" +
                             fixed_half_up((double)vf, vf_decimals);
        if (out_sb) {
            double sb = r.gatk_bias_score > -100.0 ? r.gatk_bias_score : -100.0;   // [-100, 0] (VcfWritingParameters.cs:14-15)
            sb = sb < 0.0 ? sb : 0.0;
            fmt += ":NL:SB";
            sample += ":" + std::to_string(r.noise_level) + ":" + fixed_half_up(sb, 4);
        }
        if (o.report_rc_counts) {   // US (:283-316)
            static const int with_ts[6] = {0, 1, 4, 5, 6, 7}, without_ts[4] = {0, 1, 2, 3};
            const int* idx = o.report_ts_counts ? with_ts : without_ts;
            const int ni = o.report_ts_counts ? 6 : 4;
            fmt += ":US";
            sample += ":";
            for (int k = 0; k < ni; k++) sample += (k ? "," : "") + std::to_string(ext ? ext[i].collapsed_mut[idx[k]] : 0);
            for (int k = 0; k < ni; k++) sample += "," + std::to_string(ext ? ext[i].collapsed_total[idx[k]] : 0);
        }
        out += h->chr_name + "\t" + std::to_string(r.position) + "\t.\t" + ref + "\t" + alt_col + "\t" + std::to_string(r.variant_qscore) + "\t" + filter + "\tDP=" +
               std::to_string(depth) + "\t" + fmt + "\t" + sample + "\n";
    }
    *text = out.c_str();
    *len = (int64_t)out.size();
    return PB2_OK;
}
