"""The checker's amplicon-bias path end to end: XN names on reads -> per-position amplicon coverage (RegionState.AddAmpliconCount,
RegionState.cs:269-307, from RegionStateManager.cs:188) + SupportByAmplicon of SNV candidates (CandidateVariantFinder.cs:214-231, merged in
RegionState.AddCandidate :138-170) -> AmpliconBiasCalculator.Compute (AmpliconBiasCalculator.cs:20-31) -> the AmpliconBias filter (AlleleProcessor.cs:49-50).
The per-amplicon tallies of every called SNV are checked against an independent numpy count over the same reads, and the filter against the calculator that
tests/test_oracle_amplicon_bias.py pins on the reference's own unit tests. CPU only."""
import numpy as np

from oracle import binding as ob
from tests.util_reads import amplicon_counts_at, make_amplicon_reads

AB = ob.FILTERS.index("AmpliconBias")

# amplicons 0..3 start near 20, 120, 220, 320 and are 140 long: 0/1 overlap on 121..160, 1/2 on 221..260, 2/3 on 321..360
VARIANTS = [
    (60, {0: 0.5}),                       # one amplicon only: no verdict
    (130, {0: 0.5, 1: 0.0, -1: 0.5}),     # on amplicon 0 only where 1 covers as well: bias
    (145, {0: 0.4, 1: 0.4}),              # on both: no bias
    (230, {1: 0.3, 2: 0.03}),             # far below expectation on amplicon 2 and under the 10 % free pass: bias
    (245, {1: 0.5, 2: 0.15}),             # low on amplicon 2 but above the free pass: no bias
    (330, {2: 0.0, 3: 0.25}),             # absent from amplicon 2: bias
    (350, {-1: 0.9}),                     # untagged reads only: SupportByAmplicon stays null
]


def _run(d, **kw):
    oc = ob.Caller(ob.default_config(amplicon_bias_filter=0.01, **kw), "chr1", bytes(d["ref"]).decode())
    oc.add_reads_soa(d["pos0"], d["flag"], d["cigar_off"], d["cigar"], d["seq_off"], d["bases"], d["quals"], amplicon=d["amplicon"])
    oc.finish()
    return oc.records()


def test_tallies_and_filter_follow_the_reads():
    d = make_amplicon_reads(seed=5, variants=VARIANTS)
    recs = _run(d, output_gvcf=0)
    snvs = [r for r in recs if r.type == ob.SNV]
    assert {60, 130, 145, 230, 245, 330} <= {r.pos for r in snvs}
    verdict = {}
    for r in snvs:
        cov = amplicon_counts_at(d, r.pos)
        sup = amplicon_counts_at(d, r.pos, alt=ord(r.alt))
        assert dict(zip(r.amp_coverage_names[:r.n_amp_coverage], r.amp_coverage_counts[:r.n_amp_coverage])) == cov, r.pos
        if sup:
            assert dict(zip(r.amp_support_names[:r.n_amp_support], r.amp_support_counts[:r.n_amp_support])) == sup, r.pos
        else:
            assert r.n_amp_support == -1, r.pos
        want = ob.amplicon_bias((list(sup), list(sup.values())) if sup else (None, []), (list(cov), list(cov.values())), 0.01, 100)
        assert bool(r.has_amplicon_bias) == (want is not None), r.pos
        assert bool(r.amplicon_bias_detected) == bool(want and want["bias_detected"]), r.pos
        assert bool((r.filter_mask >> AB) & 1) == bool(want and want["bias_detected"]), r.pos
        verdict[r.pos] = bool((r.filter_mask >> AB) & 1)
    assert [verdict[p] for p in (60, 130, 145, 230, 245, 330)] == [False, True, False, True, False, True]
    assert not verdict.get(350, False)


def test_no_threshold_no_tracking():
    d = make_amplicon_reads(seed=5, variants=VARIANTS)
    oc = ob.Caller(ob.default_config(output_gvcf=0), "chr1", bytes(d["ref"]).decode())
    oc.add_reads_soa(d["pos0"], d["flag"], d["cigar_off"], d["cigar"], d["seq_off"], d["bases"], d["quals"], amplicon=d["amplicon"])
    oc.finish()
    for r in oc.records():
        assert not (r.filter_mask >> AB) & 1 and r.n_amp_support == -1 and r.n_amp_coverage == -1


def test_more_than_six_amplicons_at_a_position_throws():
    """Constants.MaxNumOverlappingAmplicons = 6 slots per position: the seventh name indexes slot -1 (RegionState.cs:293-297)."""
    d = make_amplicon_reads(seed=1, n_amp=8, per_amp=20, stride=10, untagged=0)
    try:
        _run(d, output_gvcf=0)
    except RuntimeError as e:
        assert "outside the bounds" in str(e)
    else:
        raise AssertionError("expected the reference's IndexOutOfRangeException")
