"""ORACLE — TEST INFRASTRUCTURE ONLY. VCF record text for called alleles, restating the reference's writer — the somatic case (one line per
allele, `line`) and the crushed germline case (one line per position, `crushed_line`; GroupsAllelesThenWrite, VcfFileWriter.cs:177-204,
MergeCrushedReferenceAndAlt, VcfFormatter.cs:449-479) plus RegionMapper's padding (`pad_positions`, RegionMapper.cs:31-84) — so that the
reference's full-text goldens can be compared line by line:
  src/lib/Pisces.IO/VcfFileWriter.cs:206-260 (WriteListOfColocatedAlleles), VcfFormatter.cs:52-71 (VF decimals), :143-182 (filter strings),
  :184-215 (genotype strings), :224-251 (FORMAT/SAMPLE), :329-358 (VF), :373-394 (DP), :396-420 (AD).
"""
from decimal import ROUND_HALF_UP, Decimal

import numpy as np

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


def _fixed(x, decimals, single=False):
    """Custom numeric format "0.000…": .NET rounds the shortest-15-digit decimal half away from zero. single: the value is a C# float, which
    float.ToString (netcoreapp2.0, Number.FormatSingle) first reduces to 7 significant digits."""
    d = Decimal("%.6e" % float(np.float32(x))) if single else Decimal(repr(float(x)))
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
        self.nc = False   # VcfWritingParameters.ReportNoCalls: the NC tag (set by the caller)

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
        fmt, sample = "GT:GQ:AD:DP:VF", f"{GT_STR[gt]}:{rec.gq}:{ad}:{depth}:{_fixed(float(vf), self.vf_decimals, single=True)}"
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

    def crushed_line(self, chrom, recs):
        """WriteListOfColocatedAlleles for the alleles of one position (records need .phase_set_index for a lone allele of a 1/2 genotype)."""
        import numpy as np
        first = recs[0]
        gt = self.GENOTYPES[first.genotype]
        is_ref = first.type == 4
        gt12 = gt in ("HeterozygousAlt1Alt2", "Alt12LikeNoCall", "Others")
        depth = first.ref_support if is_ref else first.ref_support + first.allele_support          # GetDepthCountInt (:373-394)
        for r in recs:
            depth = max(depth, r.total_coverage)
        depth = max(depth, sum(r.allele_support for r in recs))
        vq, gq = min(r.vq for r in recs), min(r.gq for r in recs)                                   # :483-491
        if len(recs) == 1:                                                                          # SetUncrushedReferenceAndAlt (:432-447)
            ref, alt = first.ref, first.alt
            if gt12:
                alt = alt + ",." if (getattr(first, "phase_set_index", -1) == 1 or gt == "Others") else ".," + alt
        else:                                                                                       # MergeCrushedReferenceAndAlt (:449-479)
            ref = ""
            for r in recs:
                if len(r.ref) > len(ref):
                    ref = r.ref
            alt = ",".join(r.alt + (ref[len(r.ref):] if len(ref) != len(r.ref) else "") for r in recs)
        if not first.forced and gt in ("HomozygousRef", "RefLikeNoCall", "RefAndNoCall", "HemizygousNoCall", "HemizygousRef"):
            alt = "."
        names = []
        for r in recs:                                                                              # MergeFilters (:423-430) + MapFilters
            for n in self.filter_string(r).split(";"):
                if n != "PASS" and n not in names:
                    names.append(n)
        if is_ref:                                                                                  # GetAlleleCountString (:396-420)
            ad = str(first.allele_support)
        elif gt12:
            if len(recs) > 1:
                ad = ",".join(str(r.allele_support) for r in recs)
            else:
                other = depth - first.allele_support - first.ref_support
                ad = (f"{first.ref_support},{first.allele_support},{other}" if (getattr(first, "phase_set_index", -1) == 1 or gt == "Others")
                      else f"{first.ref_support},{other},{first.allele_support}")
        else:
            ad = f"{first.ref_support},{first.allele_support}"
        freq = np.float32(first.frequency)                                                          # GetFrequencyString (:329-358)
        if is_ref:
            vf = float(np.float32(0) if first.total_coverage == 0 else np.float32(1) - freq)
        vf_is_double = False
        if is_ref:
            pass
        elif gt in ("HeterozygousAlt1Alt2", "Alt12LikeNoCall"):
            vf = sum(float(r.allele_support) / float(depth) for r in recs)
            vf_is_double = True
        else:
            vf = float(freq)
        fmt, sample = "GT:GQ:AD:DP:VF", f"{GT_STR[gt]}:{gq}:{ad}:{depth}:{_fixed(vf, self.vf_decimals, single=not vf_is_double)}"
        if self.out_sb:
            fmt += ":NL:SB"
            sample += f":{first.noise_level}:{_fixed(min(max(-100.0, first.gatk_bias_score), 0.0), 4)}"
        if self.nc:
            fmt += ":NC"
            sample += ":" + _fixed(float(np.float32(first.fraction_no_calls)), 4, single=True)
        return "\t".join([chrom, str(first.pos), ".", ref, alt, str(vq), ";".join(names) if names else "PASS", f"DP={depth}", fmt, sample])


def pad_positions(intervals, written_positions, write_remaining=True):
    """RegionMapper.GetNextEmptyCall driven by VcfFileWriter.PadIfNeeded / WriteRemaining: yields ('pad', p) and ('call', p) in output order for the
    sorted positions that get real lines. intervals: list of (start, end)."""
    state = dict(last_padded=0, last_cleared=-1)
    iv_max = max((e for _, e in intervals), default=0)

    def next_empty(start, max_up_to):
        region = None
        for i in range(state["last_cleared"] + 1, len(intervals)):
            if intervals[i][1] >= start:
                region = i
                break
            state["last_cleared"] += 1
        if region is None:
            return None
        nxt = max(intervals[region][0], state["last_padded"] + 1, start)
        end = iv_max if max_up_to is None else min(max_up_to, iv_max)
        if nxt > end:
            return None
        if intervals[region][1] <= nxt:
            state["last_cleared"] += 1
        if not (intervals[region][0] <= nxt <= intervals[region][1]):
            return None
        state["last_padded"] = nxt
        return nxt

    out, last_written = [], 0
    for p in written_positions:
        if last_written == 0 or last_written + 1 < p:
            while (q := next_empty(last_written + 1, p - 1)) is not None:
                out.append(("pad", q))
                last_written = q
        out.append(("call", p))
        last_written = p
    if write_remaining:
        while (q := next_empty(last_written + 1, None)) is not None:
            out.append(("pad", q))
            last_written = q
    return out
