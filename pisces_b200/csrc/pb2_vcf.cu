// VCF record text for the records of a flush: one line per allele (somatic writer) or, crushed, one line per position (the
// germline writer: VcfWritingParameters.cs:18-40), with the interval padding of RegionMapper.
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
// The same custom format on a C# float (CalledAllele.Frequency, FractionNoCalls): float.ToString of netcoreapp2.0 first reduces the value to 7
// significant digits (Number.FormatSingle: FLOAT_PRECISION) and rounds THAT decimal half away from zero - 21/2000 = 0.0105f is 0.01049999986 as a
// double but "0.0105" in 7 digits, and prints 0.011 at three places.
std::string fixed_half_up_f32(float x, int decimals) {
    if (!std::isfinite(x)) return fixed_half_up((double)x, decimals);
    char e[40];
    snprintf(e, sizeof(e), "%.6e", (double)std::fabs(x));   // d.dddddde+XX: 7 significant digits, correctly rounded
    std::string mant;
    mant.push_back(e[0]);
    mant.append(e + 2, 6);
    const int exp10 = atoi(e + 9);
    // value = 0.mant x 10^(exp10 + 1): spell it as integer part + fraction digits, then round at `decimals`
    std::string ip, fp;
    const int point = exp10 + 1;
    if (point <= 0) { ip = "0"; fp = std::string((size_t)(-point), '0') + mant; }
    else if (point >= (int)mant.size()) { ip = mant + std::string((size_t)(point - (int)mant.size()), '0'); }
    else { ip = mant.substr(0, (size_t)point); fp = mant.substr((size_t)point); }
    while ((int)fp.size() < decimals + 1) fp.push_back('0');
    const bool up = fp[(size_t)decimals] >= '5';
    std::string digits = ip + fp.substr(0, (size_t)decimals);
    if (up) {
        int i = (int)digits.size() - 1;
        while (i >= 0) { if (digits[(size_t)i] == '9') { digits[(size_t)i] = '0'; i--; } else { digits[(size_t)i]++; break; } }
        if (i < 0) digits.insert(digits.begin(), '1');
    }
    const size_t il = digits.size() - (size_t)decimals;
    std::string body = digits.substr(0, il);
    if (decimals > 0) body += "." + digits.substr(il);
    const bool neg = std::signbit(x) && body.find_first_not_of("0.") != std::string::npos;
    return (neg ? "-" : "") + body;
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
    int status = PB2_OK;
    auto alleles_of = [&](const pb2_call_record& r, std::string& ref, std::string& alt) -> bool {
        ref.clear(); alt.clear();
        if (r.ref_len + r.alt_len <= 4) {
            for (int k = 0; k < r.ref_len; k++) ref.push_back((char)((r.allele_bytes >> (8 * k)) & 0xff));
            for (int k = 0; k < r.alt_len; k++) alt.push_back((char)((r.allele_bytes >> (8 * (r.ref_len + k))) & 0xff));
            return true;
        }
        if ((size_t)r.allele_bytes + r.ref_len + r.alt_len > h->arena.size()) return false;
        ref.assign((const char*)h->arena.data() + r.allele_bytes, r.ref_len);
        alt.assign((const char*)h->arena.data() + r.allele_bytes + r.ref_len, r.alt_len);
        return true;
    };
    // VcfFileWriter.WriteListOfColocatedAlleles (:206-260) for one line: a single allele, or (crushed form) every allele of a position
    auto write_line = [&](const pb2_call_record* const* v, const pb2_call_record_ext* const* ve, size_t nv) {
        const pb2_call_record& r = *v[0];
        const bool is_ref = r.type == CAT_REF;
        const bool forced = (r.sb_flags & 8) != 0;
        const int gt = r.genotype;
        const bool gt12 = gt == GT_HET_ALT12 || gt == GT_ALT12_NOCALL || gt == GT_OTHERS;
        const bool phase_first = (ve[0] != nullptr && ve[0]->phase_set_index == 1) || gt == GT_OTHERS;
        std::string ref, alt;
        if (!alleles_of(r, ref, alt)) { status = PB2_ERR_ARG; return; }
        // GetDepthCountInt (:373-394)
        int depth = is_ref ? r.reference_support : r.reference_support + r.allele_support;
        int total_variant_reads = 0, vq = r.variant_qscore, gq = r.genotype_qscore;
        for (size_t k = 0; k < nv; k++) {
            depth = std::max(depth, v[k]->total_coverage);
            total_variant_reads += v[k]->allele_support;
            vq = std::min(vq, v[k]->variant_qscore);       // MergeVariantQScores / MergeGenotypeQScores (:483-491)
            gq = std::min(gq, v[k]->genotype_qscore);
        }
        depth = std::max(depth, total_variant_reads);
        // REF / ALT: SetUncrushedReferenceAndAlt (:432-447) for one allele, MergeCrushedReferenceAndAlt (:449-479) for several
        if (nv == 1) {
            if (gt12) alt = phase_first ? alt + ",." : ".," + alt;
        } else {
            std::string longest;
            std::vector<std::string> refs(nv), alts(nv);
            for (size_t k = 0; k < nv; k++) {
                if (!alleles_of(*v[k], refs[k], alts[k])) { status = PB2_ERR_ARG; return; }
                if (refs[k].size() > longest.size()) longest = refs[k];
            }
            alt.clear();
            for (size_t k = 0; k < nv; k++) {
                if (k) alt += ",";
                alt += alts[k];
                if (longest.size() != refs[k].size()) alt += longest.substr(refs[k].size());
            }
            ref = longest;
        }
        const bool ref_like = gt == GT_HOM_REF || gt == GT_REF_NOCALL || gt == GT_REF_AND_NOCALL || gt == GT_HEMI_NOCALL || gt == GT_HEMI_REF;
        const std::string alt_col = (!forced && ref_like) ? "." : alt;   // VcfFileWriter.cs:233-244
        // FILTER (:143-182; MergeFilters :423-430: the alleles' filter lists concatenated, then Distinct)
        std::string filter;
        std::vector<std::string> seen;
        for (size_t k = 0; k < nv; k++)
            for (int f : kFilterOrder) {
                if (!((v[k]->filters >> f) & 1)) continue;
                std::string fs;
                switch (f) {
                    case FLT_LOW_VQ: fs = "q" + std::to_string(d.vq_filter); break;
                    case FLT_STRAND_BIAS: fs = "SB"; break;
                    case FLT_AMPLICON_BIAS: fs = "AB"; break;
                    case FLT_LOW_DEPTH: fs = "LowDP"; break;
                    case FLT_LOW_VF: fs = "LowVariantFreq"; break;
                    case FLT_LOW_GQ: fs = "LowGQ"; break;
                    case FLT_INDEL_REPEAT: fs = "R" + std::to_string(c.indel_repeat_filter); break;
                    case FLT_RMXN: fs = "R" + std::to_string(c.rmxn_max_repeat_len) + "x" + std::to_string(c.rmxn_min_repetitions); break;
                    case FLT_MULTI_ALLELIC: fs = "MultiAllelicSite"; break;
                    case FLT_FORCED_REPORT: fs = "ForcedReport"; break;
                    case FLT_NO_CALL: fs = "NC"; break;
                    default: break;
                }
                if (std::find(seen.begin(), seen.end(), fs) == seen.end()) seen.push_back(fs);
            }
        for (size_t k = 0; k < seen.size(); k++) filter += (k ? ";" : "") + seen[k];
        if (filter.empty()) filter = "PASS";
        // AD (:396-420)
        std::string ad;
        if (is_ref) ad = std::to_string(r.allele_support);
        else if (gt12) {
            if (nv > 1) { for (size_t k = 0; k < nv; k++) ad += (k ? "," : "") + std::to_string(v[k]->allele_support); }
            else {
                const int other = depth - r.allele_support - r.reference_support;
                ad = phase_first ? std::to_string(r.reference_support) + "," + std::to_string(r.allele_support) + "," + std::to_string(other)
                                 : std::to_string(r.reference_support) + "," + std::to_string(other) + "," + std::to_string(r.allele_support);
            }
        } else ad = std::to_string(r.reference_support) + "," + std::to_string(r.allele_support);
        // VF (:329-358): SumMultipleVF for 1/2 and Alt12LikeNoCall (a double sum), else the first allele's float Frequency
        const float freq = r.total_coverage == 0 ? 0.0f : std::min((float)r.allele_support / (float)r.total_coverage, 1.0f);
        const float vf32 = is_ref ? (r.total_coverage == 0 ? 0.0f : 1.0f - freq) : freq;
        double vf = (double)vf32;
        bool vf_is_double = false;
        if (!is_ref && (gt == GT_HET_ALT12 || gt == GT_ALT12_NOCALL)) {
            vf = 0;
            vf_is_double = true;
            for (size_t k = 0; k < nv; k++) vf += (double)v[k]->allele_support / (double)depth;
        }
        std::string fmt = "GT:GQ:AD:DP:VF";
        std::string sample = std::string(genotype_string(gt)) + ":" + std::to_string(gq) + ":" + ad + ":" + std::to_string(depth) + ":" +
                             (vf_is_double ? fixed_half_up(vf, vf_decimals) : fixed_half_up_f32(vf32, vf_decimals));
        if (out_sb) {
            double sb = r.gatk_bias_score > -100.0 ? r.gatk_bias_score : -100.0;   // [-100, 0] (VcfWritingParameters.cs:14-15)
            sb = sb < 0.0 ? sb : 0.0;
            fmt += ":NL:SB";
            sample += ":" + std::to_string(r.noise_level) + ":" + fixed_half_up(sb, 4);
        }
        if (o.report_no_calls) {   // NC (:257-263)
            fmt += ":NC";
            sample += ":" + fixed_half_up_f32(r.fraction_no_calls, 4);
        }
        if (o.report_rc_counts) {   // US (:283-316)
            static const int with_ts[6] = {0, 1, 4, 5, 6, 7}, without_ts[4] = {0, 1, 2, 3};
            const int* idx = o.report_ts_counts ? with_ts : without_ts;
            const int ni = o.report_ts_counts ? 6 : 4;
            fmt += ":US";
            sample += ":";
            for (int k = 0; k < ni; k++) sample += (k ? "," : "") + std::to_string(ve[0] ? ve[0]->collapsed_mut[idx[k]] : 0);
            for (int k = 0; k < ni; k++) sample += "," + std::to_string(ve[0] ? ve[0]->collapsed_total[idx[k]] : 0);
        }
        out += h->chr_name + "\t" + std::to_string(r.position) + "\t.\t" + ref + "\t" + alt_col + "\t" + std::to_string(vq) + "\t" + filter + "\tDP=" +
               std::to_string(depth) + "\t" + fmt + "\t" + sample + "\n";
    };

    // RegionMapper (src/lib/Pisces.IO/RegionMapper.cs:31-84): pads the positions of the interval set that no allele was written for with empty
    // reference calls (./., LowDP); the state below is the mapper's, the position bookkeeping VcfFileWriter's (PadIfNeeded :124-139, WriteRemaining :148-166)
    int last_padded = 0, last_cleared_interval = -1, last_written = 0;
    const int n_iv = (int)h->iv_start.size();
    int iv_max = 0;
    for (int k = 0; k < n_iv; k++) iv_max = std::max(iv_max, h->iv_end[(size_t)k]);
    auto next_empty_call = [&](int start_position, bool has_max, int max_up_to, pb2_call_record& nocall) -> bool {   // GetNextEmptyCall
        int region = -1;
        for (int k = last_cleared_interval + 1; k < n_iv; k++) {   // GetNextRegion
            if (h->iv_end[(size_t)k] >= start_position) { region = k; break; }
            last_cleared_interval++;
        }
        if (region < 0) return false;
        const int next_position = std::max(h->iv_start[(size_t)region], std::max(last_padded + 1, start_position));
        const int end_position = !has_max ? iv_max : std::min(max_up_to, iv_max);
        if (next_position > end_position) return false;
        if (h->iv_end[(size_t)region] <= next_position) last_cleared_interval++;
        if (!(next_position >= h->iv_start[(size_t)region] && next_position <= h->iv_end[(size_t)region])) return false;
        last_padded = next_position;
        // GetMissingReference (:66-82); ChrReference.GetBase throws outside the sequence (ChrReference.cs:11-17)
        if (next_position < 1 || (int64_t)next_position > (int64_t)h->h_chr.size()) { status = PB2_ERR_STATE; return false; }
        memset(&nocall, 0, sizeof(nocall));
        nocall.position = next_position;
        nocall.type = CAT_REF;
        nocall.genotype = GT_REF_NOCALL;
        nocall.filters = (uint16_t)(1u << FLT_LOW_DEPTH);
        nocall.noise_level = (uint16_t)c.min_base_call_quality;   // Factory.CreateRegionMapper (Factory.cs:248-251) hands the mapper MinimumBaseCallQuality
        const uint32_t b = h->h_chr[(size_t)next_position - 1];
        nocall.allele_bytes = b | (b << 8);
        nocall.ref_len = 1; nocall.alt_len = 1;
        return true;
    };
    auto pad = [&](bool has_max, int up_to) {
        pb2_call_record nocall;
        while (status == PB2_OK && next_empty_call(last_written + 1, has_max, up_to, nocall)) {
            const pb2_call_record* one = &nocall;
            const pb2_call_record_ext* none = nullptr;
            write_line(&one, &none, 1);
            last_written = nocall.position;
        }
    };
    const bool padding = o.pad_intervals != 0 && h->have_intervals;
    if (o.pad_intervals != 0 && h->have_intervals && h->h_chr.empty() && n_iv > 0) return pb2_fail(h, PB2_ERR_STATE, "pb2_vcf_format: interval padding needs the reference (pb2_set_reference)");
    std::vector<const pb2_call_record*> gv;
    std::vector<const pb2_call_record_ext*> ge;
    for (int64_t i = 0; i < n && status == PB2_OK;) {
        int64_t e = i + 1;
        if (o.crushed) while (e < n && recs[e].position == recs[i].position) e++;   // GroupsAllelesThenWrite (:177-204)
        if (padding && (last_written == 0 || last_written + 1 < recs[i].position)) pad(true, recs[i].position - 1);   // PadIfNeeded
        gv.clear(); ge.clear();
        for (int64_t k = i; k < e; k++) { gv.push_back(recs + k); ge.push_back(ext ? ext + k : nullptr); }
        write_line(gv.data(), ge.data(), gv.size());
        last_written = recs[i].position;
        i = e;
    }
    if (padding && o.pad_intervals >= 2 && status == PB2_OK) pad(false, 0);   // WriteRemaining
    if (status == PB2_ERR_ARG) return pb2_fail(h, PB2_ERR_ARG, "pb2_vcf_format: record alleles are not in the handle's arena");
    if (status != PB2_OK) return pb2_fail(h, status, "pb2_vcf_format: an interval position lies outside the reference sequence");
    *text = out.c_str();
    *len = (int64_t)out.size();
    return PB2_OK;
}
