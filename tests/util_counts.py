"""Test helpers: pileup entries from a table of allele counts, and the staged-count vectors of the reference's CoverageCalculatorTests
(src/test/Pisces.Calculators.Tests/UnitTests/CoverageCalculatorTests.cs:268-700) used by both the oracle and the CUDA-path tests."""
import numpy as np

A, G, C, T, N, DEL = 0, 1, 2, 3, 4, 5
SNV, INSERTION, DELETION, MNV, REFERENCE = 0, 1, 2, 3, 4


def pileup_from_counts(counts, n_loci, quality=30):
    """counts[n_loci][6][3][11] (RegionState._alleleCounts) -> CSR pileup (offsets, code, qual, anchor) with one entry per count."""
    counts = np.asarray(counts)
    offs, code, anch = [0], [], []
    for i in range(n_loci):
        for a in range(6):
            for d in range(3):
                for an in range(11):
                    k = int(counts[i, a, d, an])
                    code += [a | (d << 3)] * k
                    anch += [an] * k
        offs.append(len(code))
    code = np.array(code, dtype=np.uint8)
    return np.array(offs, dtype=np.int64), code, np.full(len(code), quality, dtype=np.uint8), np.array(anch, dtype=np.uint8)


def _well(f, r, s):
    m = np.zeros((3, 11), dtype=np.int32)
    m[0, 5], m[1, 5], m[2, 5] = f, r, s
    return m


def _table(entries, n_loci=4):
    c = np.zeros((n_loci, 6, 3, 11), dtype=np.int32)
    for coord, allele, m in entries:
        c[coord - 1, allele] += np.asarray(m, dtype=np.int32)
    return c


_FULL = dict(support=(0, 0, 5), well_anchored=(0, 0, 5))
_UNANCH = dict(support=(0, 0, 5), well_anchored=(0, 0, 0))

# name -> counts, allele, expected EstimatedCoverageByDirection (CoverageCalculator(considerAnchorInformation: true), not stitched unless said)
COVERAGE_VECTORS = {
    # :268-296 all well covered: expect min of the redistributed start / end coverage
    "insertion_all_well_covered": dict(counts=_table([(1, T, _well(10, 100, 20)), (2, C, _well(30, 50, 200))]), n_loci=4, type=INSERTION, ref="A", alt="ATCG",
                                       expected_cov=(20, 110, 0), **_FULL),
    # :298-326 right side matches the first base of the insertion, but everything is well anchored
    "insertion_first_base_well_anchored": dict(counts=_table([(1, A, _well(10, 100, 20)), (2, A, _well(20, 30, 100)), (2, T, _well(10, 20, 90))]), n_loci=4,
                                               type=INSERTION, ref="A", alt="ATCG", expected_cov=(20, 110, 0), **_FULL),
    # :328-362 length 3 -> min anchor 3: fully anchored support ignores the 6 suspicious reads (123, 141); fully unanchored support counts them
    "insertion_boundary_anchor_aware": dict(
        counts=_table([(2, A, _well(100, 1000, 200)),
                       (1, A, [[0, 0, 5, 0, 0, 15, 0, 0, 0, 0, 0], [0, 0, 0, 10, 0, 20, 0, 0, 0, 0, 0], [0, 10, 20, 0, 0, 70, 0, 0, 0, 0, 0]]),
                       (1, G, [[0, 0, 2, 0, 3, 5, 0, 0, 0, 0, 0], [0, 4, 0, 0, 6, 10, 0, 0, 0, 0, 0], [0, 0, 0, 10, 20, 60, 0, 0, 0, 0, 0]])]),
        n_loci=4, type=INSERTION, ref="A", alt="ATCG", expected_cov=(123, 141, 0), **_FULL),
    "insertion_boundary_all_unanchored": dict(
        counts=_table([(2, A, _well(100, 1000, 200)),
                       (1, A, [[0, 0, 5, 0, 0, 15, 0, 0, 0, 0, 0], [0, 0, 0, 10, 0, 20, 0, 0, 0, 0, 0], [0, 10, 20, 0, 0, 70, 0, 0, 0, 0, 0]]),
                       (1, G, [[0, 0, 2, 0, 3, 5, 0, 0, 0, 0, 0], [0, 4, 0, 0, 6, 10, 0, 0, 0, 0, 0], [0, 0, 0, 10, 20, 60, 0, 0, 0, 0, 0]])]),
        n_loci=4, type=INSERTION, ref="A", alt="ATCG", expected_cov=(125, 145, 0), **_UNANCH),
    # :364-398 shorter insertion (length 2): more anchor bins are fair game
    "insertion_shorter_anchor_aware": dict(
        counts=_table([(1, A, _well(100, 1000, 200)),
                       (2, A, [[0, 0, 0, 0, 0, 15, 0, 0, 5, 0, 0], [0, 0, 0, 0, 0, 20, 0, 10, 0, 0, 0], [0, 0, 0, 0, 0, 70, 0, 0, 20, 10, 0]]),
                       (2, T, [[0, 0, 0, 0, 0, 5, 3, 0, 2, 0, 0], [0, 0, 0, 0, 0, 10, 6, 0, 0, 4, 0], [0, 0, 0, 0, 0, 60, 20, 10, 0, 0, 0]])]),
        n_loci=4, type=INSERTION, ref="A", alt="ATC", expected_cov=(125, 141, 0), **_FULL),
}
