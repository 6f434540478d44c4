"""Row a18 end to end on the device: amplicon name ids on the reads -> per-row ids in the PVERT pileup -> AmpliconBiasCalculator.Compute
(src/lib/Pisces.Calculators/AmpliconBiasCalculator.cs:20-133) over the SNV records -> the AmpliconBias filter (AlleleProcessor.cs:49-50) and the VCF's AB,
against the CPU oracle fed the same reads and names (whose tallies tests/test_oracle_amplicon_path.py checks against an independent count). Needs a B200."""
import numpy as np
import pytest

from oracle import binding as ob
from tests.test_gpu_explicit import compare_records
from tests.test_oracle_amplicon_path import AB, VARIANTS
from tests.util_reads import make_amplicon_reads

pytestmark = pytest.mark.gpu


def _oracle(d, intervals=None, **cfg):
    oc = ob.Caller(ob.default_config(**cfg), "chr1", bytes(d["ref"]).decode(), intervals=intervals)
    oc.add_reads_soa(d["pos0"], d["flag"], d["cigar_off"], d["cigar"], d["seq_off"], d["bases"], d["quals"], amplicon=d["amplicon"])
    oc.finish()
    return oc.records()


def _product(d, packed=False, batches=1, resident=False, intervals=None, **cfg):
    import pisces_b200 as pb
    sm = pb.GpuStateManager(pb.make_config(**cfg), "chr1", bytes(d["ref"]).decode(), intervals=intervals)
    n = len(d["pos0"])
    cuts = np.linspace(0, n, batches + 1).astype(int)
    for a, b in zip(cuts[:-1], cuts[1:]):
        part = dict(pos0=d["pos0"][a:b], flag=d["flag"][a:b], cigar_off=d["cigar_off"][a:b + 1], cigar=d["cigar"], seq_off=d["seq_off"][a:b + 1], bases=d["bases"],
                    quals=d["quals"], amplicon=d["amplicon"][a:b])
        if packed:
            part = dict(part, cigar_off=part["cigar_off"] - part["cigar_off"][0], cigar=d["cigar"][a:b], seq_off=part["seq_off"] - part["seq_off"][0],
                        bases=d["bases"][d["seq_off"][a]:d["seq_off"][b]], quals=d["quals"][d["seq_off"][a]:d["seq_off"][b]])
            sm.AddReadsPacked(pb.GpuStateManager.pack_reads(part))
        else:
            sm.AddReadsSoA(part)
    if resident:   # pb2_stage_reads + pb2_call_resident (plan-building call, then graph replays): the same records as the flush, in any order
        from tests.test_gpu_explicit import _resident_records
        sm.StageReads()
        n1 = sm.call_resident()
        assert sm.call_resident() == n1 and sm.call_resident() == n1
        res = _resident_records(sm)
    recs = pb.GpuAlleleCaller().Call(sm, raw=True)
    if resident:
        assert sorted(bytes(r.tobytes()) for r in res) == sorted(bytes(r.tobytes()) for r in recs)
    arena = sm.AlleleArena()
    sm.close()
    return recs, arena


@pytest.mark.parametrize("gvcf", [0, 1])
@pytest.mark.parametrize("mode", ["flush", "flush_batched", "packed", "resident"])
def test_amplicon_bias_filter_matches_the_oracle(gvcf, mode):
    d = make_amplicon_reads(seed=5, variants=VARIANTS)
    cfg = dict(output_gvcf=gvcf, amplicon_bias_filter=0.01)
    orecs = _oracle(d, **cfg)
    if mode == "resident" and gvcf:
        pytest.skip("the resident results of a gVCF run are the dense reference stream; covered by the flush modes")
    precs, arena = _product(d, packed=mode == "packed", batches=3 if mode == "flush_batched" else 1, resident=mode == "resident", **cfg)
    compare_records(orecs, precs, arena)
    flagged = sorted(int(r["position"]) for r in precs if (int(r["filters"]) >> AB) & 1)
    assert flagged == sorted(r.pos for r in orecs if (r.filter_mask >> AB) & 1)
    assert {130, 230, 330} <= set(flagged) and not {60, 145, 245} & set(flagged)


@pytest.mark.parametrize("gvcf", [0, 1])
def test_with_an_interval_file(gvcf):
    """Interval runs stage listed positions (positions[] / index_of_pos): the amplicon kernel finds a record's locus by binary search."""
    d = make_amplicon_reads(seed=5, variants=VARIANTS)
    iv = [(100, 150), (220, 340)]
    cfg = dict(output_gvcf=gvcf, amplicon_bias_filter=0.01)
    orecs = _oracle(d, intervals=iv, **cfg)
    precs, arena = _product(d, intervals=iv, **cfg)
    compare_records(orecs, precs, arena)
    flagged = sorted(int(r["position"]) for r in precs if (int(r["filters"]) >> AB) & 1)
    assert {130, 230, 330} <= set(flagged) and 60 not in {int(r["position"]) for r in precs}


def test_many_seeds_and_shapes():
    rng = np.random.default_rng(9)
    for seed in range(6):
        n_amp = int(rng.integers(2, 6))
        stride = int(rng.integers(40, 120))
        L = 20 + (n_amp - 1) * stride + 140
        variants = [(int(p), {int(a): float(rng.choice([0.0, 0.03, 0.08, 0.15, 0.4])) for a in range(n_amp)}) for p in rng.integers(30, L - 10, 25)]
        d = make_amplicon_reads(seed=100 + seed, n_amp=n_amp, per_amp=int(rng.integers(60, 400)), stride=stride, untagged=int(rng.integers(0, 60)), variants=variants)
        cfg = dict(output_gvcf=0, amplicon_bias_filter=float(rng.choice([0.01, 0.05, 0.5])))
        orecs = _oracle(d, **cfg)
        precs, arena = _product(d, **cfg)
        compare_records(orecs, precs, arena)
        assert sum((r.filter_mask >> AB) & 1 for r in orecs) == sum((int(r["filters"]) >> AB) & 1 for r in precs)


def test_no_threshold_means_no_tracking_and_ids_are_ignored():
    d = make_amplicon_reads(seed=5, variants=VARIANTS)
    orecs = _oracle(d, output_gvcf=0)
    precs, arena = _product(d, output_gvcf=0)
    compare_records(orecs, precs, arena)
    assert not any((int(r["filters"]) >> AB) & 1 for r in precs)


def test_seventh_amplicon_name_at_a_called_snv_is_the_reference_error():
    import pisces_b200 as pb
    d = make_amplicon_reads(seed=1, n_amp=8, per_amp=40, stride=10, untagged=0, variants=[(100, {a: 0.5 for a in range(8)})])
    with pytest.raises(pb.PiscesB200Error, match="outside the bounds"):
        _product(d, output_gvcf=0, amplicon_bias_filter=0.01)


def test_amplicon_names_with_call_mnvs_are_refused():
    import pisces_b200 as pb
    d = make_amplicon_reads(seed=5, variants=VARIANTS)
    with pytest.raises(pb.PiscesB200Error, match="call_mnvs"):
        _product(d, output_gvcf=0, amplicon_bias_filter=0.01, call_mnvs=1)


def test_streamed_batches_with_partial_flushes():
    """Reads pushed batch by batch with pb2_flush(up_to) in between (the reads that end inside the cleared positions leave the store by a device compaction,
    their amplicon ids with them): the concatenated records equal the oracle's, filter included. Twelve amplicons over two 1000-bp blocks."""
    import pisces_b200 as pb
    rng = np.random.default_rng(3)
    variants = [(int(p), {int(a): float(rng.choice([0.0, 0.05, 0.3, 0.5])) for a in range(12)}) for p in rng.integers(40, 1250, 60)]
    d = make_amplicon_reads(seed=21, n_amp=12, per_amp=150, stride=100, untagged=40, variants=variants)
    cfg = dict(output_gvcf=0, amplicon_bias_filter=0.01)
    orecs = _oracle(d, **cfg)
    sm = pb.GpuStateManager(pb.make_config(**cfg), "chr1", bytes(d["ref"]).decode())
    n = len(d["pos0"])
    cuts = [0, n // 4, n // 2, 3 * n // 4, n]
    got = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        sm.AddReadsSoA(dict(pos0=d["pos0"][a:b], flag=d["flag"][a:b], cigar_off=d["cigar_off"][a:b + 1], cigar=d["cigar"], seq_off=d["seq_off"][a:b + 1], bases=d["bases"],
                            quals=d["quals"], amplicon=d["amplicon"][a:b]))
        up_to = None if b == n else int(d["pos0"][b])        # Call(read.Position - 1) of the next read (0-based pos0 = Position - 1)
        got.append(pb.GpuAlleleCaller().Call(sm, upToPosition=up_to, raw=True))
    arena = sm.AlleleArena()
    sm.close()
    precs = np.concatenate(got)
    assert max(o.pos for o in orecs) > 1000 and sum((o.filter_mask >> AB) & 1 for o in orecs) >= 3
    compare_records(orecs, precs, arena)
