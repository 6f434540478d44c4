"""ORACLE — TEST INFRASTRUCTURE ONLY. VCF record text for one called allele, restating the reference's writer for the somatic
(one line per allele) case so that the reference's full-text goldens can be compared line by line:
  src/lib/Pisces.IO/VcfFileWriter.cs:206-260 (WriteListOfColocatedAlleles), VcfFormatter.cs:52-71 (VF decimals), :143-182 (filter strings),
  :184-215 (genotype strings), :224-251 (FORMAT/SAMPLE), :329-358 (VF), :373-394 (DP), :396-420 (AD).
"""
from decimal import ROUND_HALF_UP, Decimal

GT_STR = {"HomozygousAlt": "1/1", "HomozygousRef": "0/0", "HeterozygousAltRef": "0/1", "HeterozygousAlt1Alt2": "1/2", "RefLikeNoCall": "./.",
          "AltLikeNoCall": "./.", "RefAndNoCall": "0/.", "AltAndNoCall": "1/.", "HemizygousAlt": "1", "HemizygousNoCall": ".", "HemizygousRef": "0",
          "Others": "2/2"}


def _sig_digits(value_str):
    # VcfFormatter.GetNumSigDigits (:67-71) on float.ToString()
    if "E" in value_str:
        return abs(int(value_str.split("E")[1]))
    return len(value_str) - 1


def _float_tostring(x):
    """C# float.ToString() ("G7"-like shortest round trip up to 7 digits; scientific below 1e-4)."""
    import numpy as np
    s = np.format_float_positional(np.float32(x), unique=True, trim="-")
    if abs(x) < 1e-4 and x != 0:
        m = np.format_float_scientific(np.float32(x), unique=True, trim="-", exp_digits=2)
        return m.upper().replace("E-", "E-")
    return s


def _fixed(x, decimals):
    """Custom numeric format "0.000…": .NET rounds the shortest-15-digit decimal half away from zero."""
    d = Decimal(repr(float(x)))
    q = Decimal(1).scaleb(-decimals)
    s = str(d.quantize(q, rounding=ROUND_HALF_UP))
    return s


class VcfText:
    def __init__(self, cfg, filters_enum, genotypes_enum, debug=False, output_bias_files=False, report_rc_counts=False, report_ts_counts=False):
        self.cfg = cfg
        self.FILTERS, self.GENOTYPES = filters_enum, genotypes_enum
        min_freq_filter = cfg.min_frequency_filter if cfg.min_frequency_filter > cfg.min_frequency else None   # VcfFileWriter.cs:334-347
        digits = _sig_digits(_float_tostring(cfg.min_frequency))
        if min_freq_filter is not None:
            digits = max(digits, _sig_digits(_float_tostring(min_freq_filter)))
        self.vf_decimals = digits
        self.out_sb = debug or output_bias_files or cfg.sb_acceptance < 1   # VcfFileWriter.cs:353-356
        self.rc, self.ts = report_rc_counts, report_ts_counts

    def filter_string(self, rec):
        names = []
        for f in list(rec.filters)[: rec.n_filters]:
            n = self.FILTERS[f]
            s = {"LowVariantQscore": f"q{self.cfg.vq_filter}", "StrandBias": "SB", "PoolBias": "PB", "AmpliconBias": "AB", "LowDepth": "LowDP",
                 "LowVariantFrequency": "LowVariantFreq", "LowGenotypeQuality": "LowGQ", "IndelRepeatLength": f"R{self.cfg.indel_repeat_filter}",
                 "RMxN": f"R{self.cfg.rmxn_max_repeat_len}x{self.cfg.rmxn_min_repetitions}", "MultiAllelicSite": "MultiAllelicSite",
                 "ForcedReport": "ForcedReport", "NoCall": "NC", "Unknown": "Other"}.get(n, "")
            if s not in names:
                names.append(s)
        return ";".join(names) if names else "PASS"

    def line(self, chrom, rec):
        import numpy as np
        gt = self.GENOTYPES[rec.genotype]
        is_ref = rec.type == 4
        # GetDepthCountInt (:373-394)
        depth = rec.ref_support if is_ref else rec.ref_support + rec.allele_support
        depth = max(depth, rec.total_coverage)
        depth = max(depth, rec.allele_support)
        # ALT (VcfFileWriter.cs:233-244)
        alt = "." if (not rec.forced and gt in ("HomozygousRef", "RefLikeNoCall", "RefAndNoCall", "HemizygousNoCall", "HemizygousRef")) else rec.alt
        ad = str(rec.allele_support) if is_ref else f"{rec.ref_support},{rec.allele_support}"
        freq = np.float32(rec.frequency)
        if is_ref:
            vf = np.float32(0) if rec.total_coverage == 0 else np.float32(1) - freq
        else:
            vf = freq
        fmt, sample = "GT:GQ:AD:DP:VF", f"{GT_STR[gt]}:{rec.gq}:{ad}:{depth}:{_fixed(float(vf), self.vf_decimals)}"
        if self.out_sb:
            sb = min(max(-100.0, rec.gatk_bias_score), 0.0)
            fmt += ":NL:SB"
            sample += f":{rec.noise_level}:{_fixed(sb, 4)}"
        if self.rc:   # VcfFormatter.cs:283-316
            m, t = list(rec.collapsed_mut), list(rec.collapsed_total)
            idx = [0, 1, 4, 5, 6, 7] if self.ts else [0, 1, 2, 3]
            fmt += ":US"
            sample += ":" + ",".join(str(m[i]) for i in idx) + "," + ",".join(str(t[i]) for i in idx)
        return "\t".join([chrom, str(rec.pos), ".", rec.ref, alt, str(rec.vq), self.filter_string(rec), f"DP={depth}", fmt, sample])
