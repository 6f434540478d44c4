"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded reads. Needs a B200: -m gpu."""
import math

import numpy as np
import pytest

from oracle import binding as ob
from tests import util_reads as U

pytestmark = pytest.mark.gpu


def _pb():
    import pisces_b200 as pb
    return pb


def _run_both(reads, ref, o_kw, p_kw, intervals=None, want_counts=None):
    pb = _pb()
    oc = ob.Caller(ob.default_config(**o_kw), "chr1", ref, intervals=intervals)
    for rd in reads:
        oc.add_read(U.to_oracle(rd))
    oc.finish()
    orecs = oc.records()
    sm = pb.GpuStateManager(pb.make_config(**p_kw), "chr1", ref, intervals=intervals)
    sm.AddAlleleCounts([U.to_product(rd) for rd in reads])
    counts = sm.GetAlleleCounts(*want_counts) if want_counts else None
    precs = pb.GpuAlleleCaller().Call(sm, raw=True)
    sm.close()
    return orecs, precs, counts


def _compare_records(orecs, precs, check_qsum=False):
    assert len(orecs) == len(precs), (len(orecs), len(precs))
    for o, p in zip(orecs, precs):
        ctx = f"pos {o.pos} {o.ref}>{o.alt}"
        assert o.pos == int(p["position"]) and o.type == int(p["type"]), ctx
        raw = int(p["allele_bytes"]).to_bytes(4, "little")
        assert raw[:1].decode() == o.ref and raw[1:2].decode() == o.alt, ctx
        # integers: bit exact
        assert o.total_coverage == int(p["total_coverage"]), ctx
        assert list(o.cov) == list(p["coverage_by_direction"]), ctx
        assert list(o.support) == list(p["support_by_direction"]), ctx
        assert o.allele_support == int(p["allele_support"]) and o.ref_support == int(p["reference_support"]), ctx
        assert o.num_no_calls == int(p["num_no_calls"]), ctx
        assert o.vq == int(p["variant_qscore"]) and o.gq == int(p["genotype_qscore"]), ctx
        assert o.genotype == int(p["genotype"]), ctx
        assert o.filter_mask == int(p["filters"]), (ctx, o.filter_mask, int(p["filters"]))
        assert o.noise_level == int(p["noise_level"]), ctx
        assert o.fraction_no_calls == float(p["fraction_no_calls"]), ctx
        assert (bool(o.bias_acceptable), bool(o.var_both_strands), bool(o.cov_both_strands)) == \
            (bool(p["sb_flags"] & 1), bool(p["sb_flags"] & 2), bool(p["sb_flags"] & 4)), ctx
        # doubles: 1e-6 (BASELINE.json north_star)
        for a, b in ((o.bias_score, float(p["bias_score"])), (o.gatk_bias_score, float(p["gatk_bias_score"]))):
            if math.isinf(a) or math.isnan(a):
                assert (math.isinf(b) and (a > 0) == (b > 0)) or (math.isnan(a) and math.isnan(b)), ctx
            else:
                assert b == pytest.approx(a, rel=1e-6, abs=1e-9), ctx
        if check_qsum:
            assert float(p["sum_base_quality"]) == pytest.approx(o.sum_base_quality, rel=1e-9), ctx


def _oracle_counts(reads, ref, pos0, n, **o_kw):
    oc = ob.Caller(ob.default_config(**o_kw), "chr1", ref)
    for rd in reads:
        oc.add_read(U.to_oracle(rd), "counts")
    return oc.dump_counts(pos0, n)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_counts_bit_exact(seed):
    rng = np.random.default_rng(seed)
    ref = U.random_reference(rng, 900, n_rate=0.01)
    reads = U.make_reads(rng, ref, 1500, read_len=50, stitched=(seed == 2))
    exp = _oracle_counts(reads, ref, 1, 900, min_base_call_quality=20, output_gvcf=0)
    pb = _pb()
    sm = pb.GpuStateManager(pb.make_config(), "chr1", ref)
    sm.AddAlleleCounts([U.to_product(rd) for rd in reads])
    got = sm.GetAlleleCounts(1, 900)
    sm.close()
    assert got.sum() == exp.sum() > 0
    np.testing.assert_array_equal(got, exp)


def test_reference_count_unit_vectors():
    # RegionStateManagerTests.cs:596-704 (deletions incl. leading/terminal) through the CUDA path
    pb = _pb()
    sm = pb.GpuStateManager(pb.make_config(min_base_call_quality=25), "chr1", "A" * 3000)
    sm.AddAlleleCounts([pb.Read(1001, "TTTTTTTTT", "5M4D4M", 25), pb.Read(1005, "AAAAAAAAA", "1M2D8M", 25, flag=0x10)])
    T, A, DEL = pb.AlleleType.T, pb.AlleleType.A, pb.AlleleType.Deletion
    F, R = pb.DirectionType.Forward, pb.DirectionType.Reverse
    assert sm.GetAlleleCount(1000, T, F) == 0
    for i in range(1001, 1014):
        assert sm.GetAlleleCount(i, DEL if 1006 <= i <= 1009 else T, F) == 1
    for i in range(1005, 1016):
        assert sm.GetAlleleCount(i, DEL if 1006 <= i <= 1007 else A, R) == 1
    sm.close()
    sm = pb.GpuStateManager(pb.make_config(min_base_call_quality=25), "chr1", "A" * 3000)
    sm.AddAlleleCounts([pb.Read(1001, "TTTTNNNNN", "4M2D5S", 25), pb.Read(1015, "AAAAAAAAA", "9M2D", 25, flag=0x10)])
    for i in range(1001, 1007):
        assert sm.GetAlleleCount(i, DEL if i >= 1005 else T, F) == 1
    for i in range(1015, 1026):
        assert sm.GetAlleleCount(i, DEL if i >= 1024 else A, R) == 1
    assert sm.GetAlleleCount(1026, DEL, R) == 0
    sm.close()


@pytest.mark.parametrize("gvcf", [0, 1])
@pytest.mark.parametrize("seed", [11, 12])
def test_calls_match_oracle_snv_only(seed, gvcf):
    rng = np.random.default_rng(seed)
    ref = U.random_reference(rng, 700)
    hot = {int(p): ("ACGT"[int(rng.integers(0, 4))], float(rng.uniform(0.02, 0.6))) for p in rng.integers(60, 600, 25)}
    reads = U.make_reads(rng, ref, 4000, read_len=50, hotspots=hot, del_rate=0.0, ins_rate=0.0, clip_rate=0.1)
    # collapse off: open-ended bookkeeping does not apply; indels absent: SNV + reference alleles only
    orecs, precs, _ = _run_both(reads, ref, dict(output_gvcf=gvcf, collapse=0), dict(output_gvcf=gvcf, collapse=0, want_sum_base_quality=1))
    assert len(orecs) > (300 if gvcf else 5)
    _compare_records(orecs, precs, check_qsum=True)


def test_calls_match_oracle_with_collapse_on():
    rng = np.random.default_rng(21)
    ref = U.random_reference(rng, 500)
    hot = {int(p): ("ACGT"[int(rng.integers(0, 4))], 0.3) for p in rng.integers(60, 400, 12)}
    reads = U.make_reads(rng, ref, 3000, read_len=40, hotspots=hot, del_rate=0.0, ins_rate=0.0, clip_rate=0.0)
    orecs, precs, _ = _run_both(reads, ref, dict(output_gvcf=1, collapse=1), dict(output_gvcf=1, collapse=1))
    _compare_records(orecs, precs)


def test_intervals_and_zero_coverage():
    rng = np.random.default_rng(31)
    ref = U.random_reference(rng, 2500)
    reads = U.make_reads(rng, ref, 800, read_len=50, start=200, span=600, del_rate=0.0, ins_rate=0.0)
    iv = [(150, 260), (500, 520), (900, 950), (2100, 2200)]   # 150-199 and 900-950 have no coverage; 2100+ is in an untouched block
    orecs, precs, _ = _run_both(reads, ref, dict(output_gvcf=1, collapse=0), dict(output_gvcf=1, collapse=0), intervals=iv)
    assert any(r.total_coverage == 0 for r in orecs)
    _compare_records(orecs, precs)


def test_kat_values_through_kernel():
    # PhiX golden (PhiX_S3.noisy.vcf pos 4 T>G): cov F/R 49/199, support 0/1, NL 40 -> Q 16, SB -16.9682; QualityCalculatorTests.cs:62-95 Q values
    pb = _pb()
    ref = "ACGTT" * 20
    reads = [pb.Read(1, "ACGTT", "5M", 30) for _ in range(49)] + [pb.Read(1, "ACGTT", "5M", 30, flag=0x10) for _ in range(198)] + \
        [pb.Read(1, "ACGGT", "5M", 30, flag=0x10)]
    sm = pb.GpuStateManager(pb.make_config(min_base_call_quality=10, forced_noise_level=40, min_variant_qscore=1, min_frequency=0.00001,
                                           min_coverage=2, collapse=0, output_gvcf=0), "phix", ref)
    sm.AddAlleleCounts(reads)
    calls = pb.GpuAlleleCaller().Call(sm)
    sm.close()
    a = calls[4][0]
    assert (a.ReferenceAllele, a.AlternateAllele, a.VariantQscore, a.TotalCoverage, a.AlleleSupport, a.ReferenceSupport) == ("T", "G", 16, 248, 1, 247)
    assert f"{a.GATKBiasScore:.4f}" == "-16.9682"
    assert [f.name for f in a.Filters] == ["LowVariantQscore"]


def test_empty_and_ragged_inputs():
    pb = _pb()
    sm = pb.GpuStateManager(pb.make_config(), "chr1", "ACGT" * 100)
    assert pb.GpuAlleleCaller().Call(sm) == {}            # nothing staged
    # CSR with empty loci, a 1-entry locus and a locus that is not a multiple of 16 long
    depths = [0, 1, 0, 17, 33, 0]
    off = np.concatenate([[0], np.cumsum(depths)]).astype(np.int64)
    n = int(off[-1])
    code = np.full(n, 0 | (0 << 3), dtype=np.uint8)       # allele A, forward
    qual = np.full(n, 30, dtype=np.uint8)
    anch = np.full(n, 5, dtype=np.uint8)
    sm.AddPileup(off, code, qual, anch, first_position=1)
    got = sm.GetAlleleCounts(1, 6)
    assert [int(got[i].sum()) for i in range(6)] == depths
    assert int(got[3, 0, 0, 5]) == 17
    sm.close()


def _pileup_both(d, gvcf, **kw):
    pb = _pb()
    ref = bytes(d["ref_bases"].numpy()).decode()
    off, code, qual, anch = (d[k].numpy() for k in ("offsets", "code", "qual", "anchor"))
    oc = ob.Caller(ob.default_config(output_gvcf=gvcf, **kw), "chr1", ref)
    oc.add_pileup(off, code, qual, anch, 1)
    oc.finish()
    pkw = dict(kw)
    if pkw.get("noise_model") == 1:
        pkw["want_sum_base_quality"] = 1
    sm = pb.GpuStateManager(pb.make_config(output_gvcf=gvcf, **pkw), "chr1", ref)
    sm.AddPileup(off, code, qual, anch, first_position=1)
    precs = pb.GpuAlleleCaller().Call(sm, raw=True)
    counts = sm.GetAlleleCounts(1, len(off) - 1)
    sm.close()
    oc2 = ob.Caller(ob.default_config(output_gvcf=gvcf, **kw), "chr1", ref)
    oc2.add_pileup(off, code, qual, anch, 1, call_every=0)
    return oc.records(), precs, counts, oc2.dump_counts(1, len(off) - 1)


@pytest.mark.parametrize("gvcf", [0, 1])
@pytest.mark.parametrize("depth,n_loci", [(40, 3000), (500, 2500), (2000, 1500)])
def test_locus_major_pileup_matches_oracle(depth, n_loci, gvcf):
    """The bench data path (pb2_push_pileup) against the oracle fed the same entries; depth 2000 wraps the 8-bit histogram counters."""
    from pisces_b200 import synth
    d = synth.make_pileup(n_loci, depth, seed=depth + gvcf, snv_rate=0.05, del_rate=0.01)
    orecs, precs, counts, ocounts = _pileup_both(d, gvcf)
    np.testing.assert_array_equal(counts, ocounts)     # 16-bit variant (full bin dump)
    assert len(orecs) > 50
    _compare_records(orecs, precs)                      # 8-bit variant (scoring path)


def test_window_noise_model_and_qsum():
    from pisces_b200 import synth
    d = synth.make_pileup(1500, 300, seed=5, snv_rate=0.05)
    orecs, precs, _, _ = _pileup_both(d, 1, noise_model=1)
    _compare_records(orecs, precs, check_qsum=True)


def test_packed2_layout_matches_three_planes():
    """PB2_LAYOUT_PACKED2 (two bytes per entry + sparse candidate flags, the e2e / PCIe form of pb2_push_pileup) stages the same pileup as the three
    planes: identical records, field for field, and identical 198-bin counts."""
    import pisces_b200 as pb
    from pisces_b200 import synth
    d = synth.make_pileup(3000, 90, seed=31, snv_rate=0.05)
    ref = bytes(d["ref_bases"].numpy()).decode()
    off, code, qual, anch = (d[k].numpy().copy() for k in ("offsets", "code", "qual", "anchor"))
    rng = np.random.default_rng(5)
    pick = rng.choice(len(code), 400, replace=False)
    code[pick] |= 0x20                                    # some open-left SNV candidates (PB2_ENTRY_OPEN_LEFT) for the sparse flag list
    outs = []
    for packed in (False, True):
        sm = pb.GpuStateManager(pb.make_config(output_gvcf=1), "chr1", ref)
        if packed:
            pc, pq, fi, fb = pb.GpuStateManager.pack_pileup(code, qual, anch, off, d["ref_bases"].numpy())
            assert 0 < len(fi) < len(pb.GpuStateManager.pack_pileup(code, qual, anch)[2]) and pc.nbytes + pq.nbytes == 2 * len(code)
            sm.AddPileupPacked(off, pc, pq, fi, fb, first_position=1)
        else:
            sm.AddPileup(off, code, qual, anch, first_position=1)
        counts = sm.GetAlleleCounts(1, 3000)
        recs = pb.GpuAlleleCaller().Call(sm, raw=True)
        outs.append((np.asarray(counts).copy(), np.array(recs)))
        sm.close()
    assert (outs[0][0] == outs[1][0]).all()
    a, b = outs[0][1], outs[1][1]
    assert len(a) == len(b) and len(a) >= 3000
    for f in a.dtype.names:
        assert np.array_equal(a[f], b[f], equal_nan=a[f].dtype.kind == "f"), f
