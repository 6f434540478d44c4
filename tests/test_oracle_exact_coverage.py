"""The checker's restatement of ExactCoverageCalculator's per-read decision (oracle/po_exact_coverage.hpp) against every case of the reference's
ExactCoverageCalculatorTests (tests/golden/exact_coverage_cases.json, written by tests/golden/make_exact_coverage_fixture.py). The product does not
build this calculator yet: this pins the oracle for it. CPU only."""
import json
import os

import pytest

from oracle import binding as ob

CASES = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "exact_coverage_cases.json")))
TYPES = {"Snv": 0, "Insertion": 1, "Deletion": 2, "Mnv": 3}
DIRECTIONS = {"Forward": 0, "Reverse": 1, "Stitched": 2, None: None}


def test_fixture_is_complete():
    assert len(CASES) == 53 and {c["type"] for c in CASES} == {"Insertion", "Deletion", "Mnv"}


@pytest.mark.parametrize("i", range(len(CASES)))
def test_reference_case(i):
    c = CASES[i]
    # ExecuteTest (:370-400): insertion A>ATTTT, deletion ATTTT>A, MNV AAAA>TTTT at the variant position: BaseAllele.Length is 4 for all three
    got = ob.exact_spanning_read_direction(TYPES[c["type"]], c["start"], c["end"], c["cigar"], c["directions"], position=c["position"], allele_length=4)
    assert got == DIRECTIONS[c["expected"]], c
