#!/usr/bin/env python3
"""Writes tests/golden/exact_coverage_cases.json: the (allele type, expected direction, ClipAdjustedStart/End, CIGAR, direction string, variant position)
sequence of the reference's ExactCoverageCalculatorTests (src/test/Pisces.Calculators.Tests/UnitTests/ExactCoverageCalculatorTests.cs:18-368), obtained by
replaying the assignments of the three test methods in order. Run in the build container, where /root/reference is mounted."""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/test/Pisces.Calculators.Tests/UnitTests/ExactCoverageCalculatorTests.cs"


def main():
    src = open(SRC, encoding="utf-8-sig").read()
    body = src[src.index("public void Insertion()"):src.index("private void ExecuteTest")]
    state, cases = {}, []
    for line in body.splitlines():
        line = line.strip()
        m = re.match(r'(?:readSummary\.)?(ClipAdjustedStartPosition|ClipAdjustedEndPosition)\s*=\s*(\d+)', line)
        if m:
            state[m.group(1)] = int(m.group(2))
            continue
        m = re.match(r'(?:readSummary\.)?(?:CigarString\s*=\s*"([^"]+)"|Cigar\s*=\s*new CigarAlignment\("([^"]+)"\))', line)
        if m:
            state["cigar"] = m.group(1) or m.group(2)
            continue
        m = re.match(r'(?:readSummary\.)?DirectionString\s*=\s*"([^"]+)"', line)
        if m:
            state["dirs"] = m.group(1)
            continue
        m = re.match(r'ExecuteTest\(AlleleCategory\.(\w+),\s*(null|DirectionType\.(\w+)),\s*readSummary(?:,\s*(\d+))?\);', line)
        if m:
            cases.append(dict(type=m.group(1), expected=m.group(3), start=state["ClipAdjustedStartPosition"], end=state["ClipAdjustedEndPosition"],
                              cigar=state["cigar"], directions=state["dirs"], position=int(m.group(4) or 10)))
            continue
        assert "ExecuteTest" not in line, line
    json.dump(cases, open(os.path.join(HERE, "exact_coverage_cases.json"), "w"), indent=1)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
