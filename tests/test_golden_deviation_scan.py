"""The documented deviations of the CUDA path from the reference ("What a maintainer must know") cannot be reached by any of
the reference's own BAM fixtures: this scan proves it read by read (CPU test; the Python BAM decoder of the test suite).

1. A read whose alignment starts with an insertion / deletion exactly at a 1000-bp block boundary, pushed after that block was called: the reference
   re-creates the cleared block (RegionStateManager.GetBlock, RegionStateManager.cs:361-383) and calls the candidate later against the fresh block's counts;
   the library drops a candidate that lands in cleared positions (pb2_explicit.cu:find_candidates_impl).
2. A stitched read whose XD direction changes INSIDE a deleted span: the library takes a deletion's direction from the base after the gap
   (RegionStateManager.cs:170-177 does the same through SequencedBaseDirectionMap, which has no entry for deleted bases), and the candidate finder's
   GetSupportDirection (CandidateVariantFinder.cs:396-445) looks at the read bases on either side of the gap - an XD run boundary that falls on a deleted
   base is not transmitted by pb2_read_batch.base_dirs.
3. SNV support from '=' / 'X' CIGAR operations and pathological open-end groups are answered with PB2_ERR_UNSUPPORTED instead of a guess.
"""
import os

import pytest

from tests import bamio

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BAMS = ["PhiX_S3.bam", "collapsed.test.stitched.bam", "example_S1.mapped.bam"]
REF_BAMS = "/root/reference/src/test/SharedData/Bams"
OPS = "MIDNSHP=X"


def _all_bams():
    out = [os.path.join(G, b) for b in BAMS]
    if os.path.isdir(REF_BAMS):   # in the build container the reference's other BAM fixtures are scanned too
        out += [os.path.join(REF_BAMS, f) for f in sorted(os.listdir(REF_BAMS)) if f.endswith(".bam") and os.path.getsize(os.path.join(REF_BAMS, f)) < 8 << 20]
    return out


def _kept(recs):
    return [r for r in recs if not (r["flag"] & 0x4 or r["flag"] & 0x100 or r["mapq"] < 1 or not r["cigar"] or r["ref_id"] < 0)]


@pytest.mark.parametrize("path", _all_bams(), ids=os.path.basename)
def test_no_fixture_read_reaches_a_documented_deviation(path):
    _, _, recs = bamio.read_bam(path)
    kept = _kept(recs)
    assert kept
    for r in kept:
        ops = [(OPS[c & 15], c >> 4) for c in r["cigar"]]
        # (3) no '=' / 'X' operations anywhere
        assert not any(op in "=X" for op, _ in ops), r["name"]
        # (1) first non-clip operation an indel AND the candidate position (the base before the read's first aligned base) the last of a block
        body = [op for op, _ in ops if op not in "SH"]
        if body and body[0] in "ID":
            assert r["pos0"] % 1000 != 0, r["name"]
        # (2) XD: the directions over every deleted span equal the direction of the base that follows it
        xd = r["tags"].get("XD")
        if xd and any(op == "D" for op, _ in ops):
            expanded, num = [], ""
            for ch in xd:
                if ch.isdigit():
                    num += ch
                else:
                    expanded += [ch] * int(num)
                    num = ""
            ci = 0
            for k, (op, ln) in enumerate(ops):
                if op == "D" and ci + ln < len(expanded):
                    assert set(expanded[ci:ci + ln]) == {expanded[ci + ln]}, (r["name"], xd)
                ci += ln
