"""The reads path end to end on synthetic read sets of the BASELINE.json shapes: struct-of-arrays reads -> device read store -> PVERT pileup (built on the
device) -> hot kernel + explicit candidates, against the CPU oracle fed the same reads through the reference's per-read loop. Needs a B200: -m gpu."""
import numpy as np
import pytest

from oracle import binding as ob
from pisces_b200 import synth
from tests.test_gpu_explicit import _resident_records, compare_records

pytestmark = pytest.mark.gpu


def _pb():
    import pisces_b200 as pb
    return pb


def _oracle(d, counts_only=False, **kw):
    okw = dict(kw)
    if "expect_stitched" in okw:
        okw["source_is_stitched"] = okw.pop("expect_stitched")
    if "expect_collapsed" in okw:
        okw["source_is_collapsed"] = okw.pop("expect_collapsed")
    oc = ob.Caller(ob.default_config(**okw), "chr1", bytes(d["ref"]).decode())
    oc.add_reads_soa(d["pos0"], d["flag"], d["cigar_off"], d["cigar"], d["seq_off"], d["bases"], d["quals"], d.get("collapsed"), d.get("xd_runs"), counts_only=counts_only)
    return oc


CONFIGS = {
    # BASELINE.json configs[1]: SNV + indel, Poisson model, VCF only
    "c2": (dict(indel_rate=0.002), dict(output_gvcf=0, collapse=1)),
    # configs[2]: gVCF mode
    "c3": (dict(indel_rate=0.0, snv_rate=0.005), dict(output_gvcf=1, collapse=1)),
    # configs[3]: MNV phasing + strand-bias filter
    "c4": (dict(indel_rate=0.0, mnv_pair_rate=0.002, strand_skew_frac=0.1), dict(output_gvcf=0, collapse=1, call_mnvs=1, max_size_mnv=3, max_gap_mnv=1)),
    # configs[4]: SNV / MNV / indel + collapsed-read model, stitched reads
    "c5": (dict(indel_rate=0.001, mnv_pair_rate=0.001, collapsed_frac=1.0, stitched_frac=0.5),
           dict(output_gvcf=0, collapse=1, call_mnvs=1, expect_collapsed=1, expect_stitched=1)),
}


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_synthetic_reads_records_match_the_oracle(name):
    pb = _pb()
    gen, cfg = CONFIGS[name]
    d = synth.make_reads(12000, 120, seed=11, **gen)
    oc = _oracle(d, **cfg)
    oc.finish()
    sm = pb.GpuStateManager(pb.make_config(**cfg), "chr1", bytes(d["ref"]).decode())
    sm.AddReadsSoA(d)
    precs = pb.GpuAlleleCaller().Call(sm, raw=True)
    arena = sm.AlleleArena()
    ext = sm.AlleleExt()
    sm.close()
    orecs = oc.records()
    assert len(orecs) > 50
    compare_records(orecs, precs, arena)
    if cfg.get("expect_collapsed"):
        for o, e in zip(orecs, ext):
            assert list(o.collapsed_total) == list(e["collapsed_total"]) and list(o.collapsed_mut) == list(e["collapsed_mut"]), o.pos


@pytest.mark.parametrize("name", ["c2", "c5"])
def test_counts_of_synthetic_reads_bit_exact(name):
    pb = _pb()
    gen, cfg = CONFIGS[name]
    d = synth.make_reads(3000, 80, seed=5, **gen)
    oc = _oracle(d, counts_only=True, **cfg)
    sm = pb.GpuStateManager(pb.make_config(**cfg), "chr1", bytes(d["ref"]).decode())
    sm.AddReadsSoA(d)
    got = sm.GetAlleleCounts(1, 3000)
    sm.close()
    np.testing.assert_array_equal(got, oc.dump_counts(1, 3000))


@pytest.mark.parametrize("batch", [1, 7, 1000])
def test_batched_pushes_equal_one_push(batch):
    """The store appends batch after batch (offsets rebased on the device); a batch may be a window of larger arrays."""
    pb = _pb()
    gen, cfg = CONFIGS["c2"]
    d = synth.make_reads(4000, 60, seed=9, **gen)
    ref = bytes(d["ref"]).decode()
    sm = pb.GpuStateManager(pb.make_config(**cfg), "chr1", ref)
    sm.AddReadsSoA(d)
    want = pb.GpuAlleleCaller().Call(sm, raw=True)
    sm.close()
    sm = pb.GpuStateManager(pb.make_config(**cfg), "chr1", ref)
    n = d["n_reads"]
    for r0 in range(0, n, batch):
        r1 = min(n, r0 + batch)
        sm.AddReadsSoA(dict(pos0=d["pos0"][r0:r1], flag=d["flag"][r0:r1], cigar_off=d["cigar_off"][r0:r1 + 1], cigar=d["cigar"], seq_off=d["seq_off"][r0:r1 + 1],
                            bases=d["bases"], quals=d["quals"]))
    got = pb.GpuAlleleCaller().Call(sm, raw=True)
    sm.close()
    assert len(want) > 20 and want.tobytes() == got.tobytes()


def test_packed_reads_equal_plain_reads():
    """pb2_push_reads_packed (one byte per base + exceptions) fills the same read store: N bases, qualities 0 and > 63 go through the exception list."""
    pb = _pb()
    gen, cfg = CONFIGS["c2"]
    d = synth.make_reads(6000, 80, seed=13, **gen)
    rng = np.random.default_rng(3)
    bases, quals = d["bases"].copy(), d["quals"].copy()
    bases[rng.random(len(bases)) < 0.004] = ord("N")
    quals[rng.random(len(quals)) < 0.002] = 0
    quals[rng.random(len(quals)) < 0.002] = 70
    d = dict(d, bases=bases, quals=quals)
    ref = bytes(d["ref"]).decode()
    sm = pb.GpuStateManager(pb.make_config(**cfg), "chr1", ref)
    sm.AddReadsSoA(d)
    want = pb.GpuAlleleCaller().Call(sm, raw=True)
    sm.close()
    packed = pb.GpuStateManager.pack_reads(d)
    assert len(packed["exc_index"]) > 100
    sm = pb.GpuStateManager(pb.make_config(**cfg), "chr1", ref)
    sm.AddReadsPacked(packed)
    got = pb.GpuAlleleCaller().Call(sm, raw=True)
    arena = sm.AlleleArena()
    sm.close()
    assert len(want) > 50 and want.tobytes() == got.tobytes()
    oc = _oracle(d, **cfg)
    oc.finish()
    compare_records(oc.records(), got, arena)


def test_staged_reads_resident_step_equals_flush():
    """pb2_stage_reads + pb2_call_resident (the bench's device-resident step) emits the records pb2_flush returns."""
    pb = _pb()
    gen, cfg = CONFIGS["c2"]
    d = synth.make_reads(20000, 100, seed=3, **gen)
    sm = pb.GpuStateManager(pb.make_config(**cfg), "chr1", bytes(d["ref"]).decode())
    sm.AddReadsSoA(d)
    sm.StageReads()
    n1 = sm.call_resident()
    assert sm.call_resident() == n1 and sm.call_resident() == n1   # plan-building call, then replays
    res = _resident_records(sm)
    precs = pb.GpuAlleleCaller().Call(sm, raw=True)
    sm.close()
    short = lambda recs: sorted(bytes(r.tobytes()) for r in recs if int(r["ref_len"]) + int(r["alt_len"]) <= 4)   # longer alleles point into per-call arenas
    assert len(precs) > 100 and len(res) == len(precs)
    assert short(res) == short(precs)


@pytest.mark.parametrize("name,n_shards", [("c2", 2), ("c2", 4), ("c3", 2), ("c4", 3), ("c5", 2)])
def test_interval_shards_concatenate_to_the_unsharded_records(name, n_shards):
    """pb2_shard_plan + pb2_set_owned_range: every shard sees its reads (own positions + halo) and emits only what it owns; the shards' records in shard
    order are the unsharded chromosome's records, byte for byte - with reads spanning every cut (depth 100 everywhere), indels, MNVs and collapsing."""
    from pisces_b200 import sharding
    pb = _pb()
    gen, cfg = CONFIGS[name]
    d = synth.make_reads(9000, 100, seed=21, **gen)
    ref = bytes(d["ref"]).decode()
    sm = pb.GpuStateManager(pb.make_config(**cfg), "chr1", ref)
    sm.AddReadsSoA(d)
    want = pb.GpuAlleleCaller().Call(sm, raw=True)
    sm.close()
    plan = sharding.shard_plan(d["pos0"], 1, len(ref), d["read_len"] + 4, n_shards)
    assert sum(1 for s in plan if s["own_hi"] >= s["own_lo"]) >= 2
    parts = []
    for s in plan:
        if s["own_hi"] < s["own_lo"]:
            continue
        sm = pb.GpuStateManager(pb.make_config(**cfg), "chr1", ref)
        sm._chk(sm._L.pb2_set_owned_range(sm._h, s["own_lo"], s["own_hi"]))
        sm.AddReadsSoA(sharding.shard_reads(d, s))
        got = pb.GpuAlleleCaller().Call(sm, raw=True)
        assert len(got) == 0 or (int(got["position"].min()) >= s["own_lo"] and int(got["position"].max()) <= s["own_hi"])
        parts.append(got)
        sm.close()
    got = np.concatenate(parts)
    short = lambda recs: [bytes(r.tobytes()) for r in recs if int(r["ref_len"]) + int(r["alt_len"]) <= 4]   # longer alleles point into per-call arenas
    key = lambda recs: [(int(r["position"]), int(r["type"]), int(r["ref_len"]), int(r["alt_len"]), int(r["allele_support"]), int(r["total_coverage"]), int(r["variant_qscore"]))
                        for r in recs]
    assert len(want) > 80 and key(got) == key(want)
    assert short(got) == short(want)


def test_example_s1_bam_counts():
    """BASELINE.json configs[0] (C1): the mapped reads of the reference's testdata/example_S1.bam (tests/golden/example_S1.mapped.bam: chr1 and chr12
    amplicon pile-ups, soft clips, insertions, deletions) through the library's BAM stager and the device pileup: the [6][3][11] counts of every covered
    position equal the oracle's, bit for bit. (The hg19 sequence is not available offline - SURVEY 8c - so C1 pins counts, not calls.)"""
    import os
    from tests import bamio
    pb = _pb()
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    path = os.path.join(G, "example_S1.mapped.bam")
    _, refs, recs = bamio.read_bam(path)
    ops = "MIDNSHP=X"
    by_chr = {}
    for r in recs:
        if r["flag"] & 0x4 or r["flag"] & 0x100 or r["flag"] & 0x400 or r["mapq"] < 1 or not r["cigar"] or r["ref_id"] < 0:
            continue
        by_chr.setdefault(r["ref_id"], []).append(r)
    st = pb.BamReadStager(path, max_reads=100)
    sms = {}
    for ref_id, batch, _ in st:
        if ref_id not in sms:
            sms[ref_id] = pb.GpuStateManager(pb.make_config(min_coverage=10), refs[ref_id][0], None)
        sms[ref_id].AddReadBatch(batch)
    st.close()
    assert sorted(sms) == sorted(by_chr) and len(by_chr) == 2
    total = 0
    for ref_id, rr in by_chr.items():
        lo = min(r["pos0"] + 1 for r in rr)
        hi = max(r["pos0"] + sum(c >> 4 for c in r["cigar"] if ops[c & 15] in "MDN=X") for r in rr)
        oc = ob.Caller(ob.default_config(min_coverage=10), refs[ref_id][0], "")
        for r in rr:
            oc.add_read(ob.SimpleRead(r["pos0"] + 1, r["seq"], r["cigar"], r["qual"], flag=r["flag"], mapq=r["mapq"]), mode="counts")
        # the pile-ups are a few hundred positions wide, far apart: compare around every read start
        starts = sorted({(r["pos0"] // 1000) * 1000 + 1 for r in rr} | {((r["pos0"] + 200) // 1000) * 1000 + 1 for r in rr})
        for w in starts:
            if w > hi:
                continue
            got = sms[ref_id].GetAlleleCounts(w, 1000)
            exp = oc.dump_counts(w, 1000)
            np.testing.assert_array_equal(got, exp)
            total += int(exp.sum())
        sms[ref_id].close()
        assert lo >= 1
    assert total > 25_000


def test_invalid_reads_are_rejected_as_the_reference_does():
    pb = _pb()
    sm = pb.GpuStateManager(pb.make_config(), "chr1", "ACGT" * 50)
    with pytest.raises(pb.PiscesB200Error, match="Invalid cigar"):
        sm.AddAlleleCounts(pb.Read(5, "ACGTACGT", "5M", [30] * 8))
    sm.AddAlleleCounts(pb.Read(5, "ACGTACGT", "8M", [30] * 8))   # the handle stays usable
    assert len(pb.GpuAlleleCaller().Call(sm, raw=True)) > 0
    sm.close()


def test_oversized_indel_is_refused():
    pb = _pb()
    sm = pb.GpuStateManager(pb.make_config(), "chr1", "ACGT" * 50)
    with pytest.raises(pb.PiscesB200Error, match="65534"):
        sm.AddAlleleCounts(pb.Read(5, "ACGTACGT", "4M70000D4M", [30] * 8))
    sm.close()


def test_variant_stream_larger_than_one_record_per_locus():
    """Permissive thresholds on deep noisy data: up to three SNV alleles per locus are callable, more records than the one-per-locus stream the segment starts
    with - the flush grows the stream to the counter's exact need and runs the segment again (no PB2_ERR_NOMEM), records equal to the oracle's."""
    pb = _pb()
    cfg = dict(output_gvcf=0, min_frequency=0.0002, min_frequency_filter=0.0002, min_variant_qscore=0, variant_qscore_filter=0, min_coverage=1, forced_noise_level=60)
    d = synth.make_reads(1000, 6000, seed=21, indel_rate=0.0, snv_rate=0.3)
    okw = dict(output_gvcf=0, min_frequency=0.0002, min_frequency_filter=0.0002, min_vq=0, vq_filter=0, min_coverage=1, forced_noise_level=60)
    oc = _oracle(d, **okw)
    oc.finish()
    orecs = oc.records()
    sm = pb.GpuStateManager(pb.make_config(**cfg), "chr1", bytes(d["ref"]).decode())
    sm.AddReadsSoA(d)
    precs = pb.GpuAlleleCaller().Call(sm, raw=True)
    arena = sm.AlleleArena()
    sm.close()
    assert len(orecs) > 1024 and len(orecs) > len(d["ref"])   # more than one per staged locus
    compare_records(orecs, precs, arena)


@pytest.mark.parametrize("gvcf", [0, 1])
def test_reads_far_apart_without_intervals_stage_only_touched_blocks(gvcf):
    """Two read clusters 3 Mbp apart in one flush, no interval file: only the positions of the 1000-bp blocks reads touch are staged (the reference only
    creates those blocks, RegionStateManager.cs:361-383), records equal to the oracle's."""
    pb = _pb()
    d = synth.make_reads(3000, 80, seed=31, indel_rate=0.002)
    far = 3_000_000
    n = len(d["pos0"])
    ref = np.frombuffer(b"ACGT", dtype=np.uint8)[np.random.default_rng(1).integers(0, 4, far + len(d["ref"]) + 10)].copy()
    ref[: len(d["ref"])] = d["ref"]
    ref[far: far + len(d["ref"])] = d["ref"]
    both = dict(ref=ref, pos0=np.concatenate([d["pos0"], d["pos0"] + far]), flag=np.concatenate([d["flag"], d["flag"]]),
                cigar_off=np.concatenate([d["cigar_off"], d["cigar_off"][1:] + d["cigar_off"][-1]]), cigar=np.concatenate([d["cigar"], d["cigar"]]),
                seq_off=np.concatenate([d["seq_off"], d["seq_off"][1:] + d["seq_off"][-1]]), bases=np.concatenate([d["bases"], d["bases"]]),
                quals=np.concatenate([d["quals"], d["quals"]]))
    assert len(both["pos0"]) == 2 * n
    cfg = dict(output_gvcf=gvcf, collapse=1)
    oc = _oracle(both, **cfg)
    oc.finish()
    sm = pb.GpuStateManager(pb.make_config(**cfg), "chr1", bytes(ref).decode())
    sm.AddReadsSoA(both)
    sm.StageReads()
    st = sm.stage_stats()
    assert st["staged_bytes"] < st["rows"] * 32 + 1_000_000   # the rows of two clusters + their few tiles' tables, not three million empty loci
    precs = pb.GpuAlleleCaller().Call(sm, raw=True)
    arena = sm.AlleleArena()
    sm.close()
    orecs = oc.records()
    assert len(orecs) > 40 and any(o.pos > far for o in orecs)
    compare_records(orecs, precs, arena)


@pytest.mark.parametrize("n_slots", [3, 4])
def test_resident_sink_holds_every_steps_records_in_position_order(n_slots):
    """pb2_set_resident_sink + pb2_call_resident_async + pb2_sink_sort + pb2_resident_sync (the job buffer of the multi-GPU gather): every slot of
    [n_slots counts | n_slots record blocks] holds one step's variant records, ordered by position, equal to the flush's records. An odd slot count puts
    the record blocks at an address that is only 8-byte aligned."""
    import torch
    from pisces_b200 import _native as N
    pb = _pb()
    gen, cfg = CONFIGS["c2"]
    d = synth.make_reads(20000, 100, seed=5, **gen)
    sm = pb.GpuStateManager(pb.make_config(**cfg), "chr1", bytes(d["ref"]).decode())
    sm.AddReadsSoA(d)
    sm.StageReads()
    n0 = sm.call_resident()
    assert sm.call_resident() == n0
    cap = n0 + 7
    buf = torch.zeros(8 * n_slots + n_slots * cap * 96, dtype=torch.uint8, device="cuda")
    sm.set_resident_sink(buf.data_ptr(), cap, n_slots)
    for _ in range(n_slots + 2):      # wraps around: slot = step % n_slots
        sm.call_resident_async()
    sm.sink_sort()
    assert sm.resident_sync() == n0
    host = buf.cpu().numpy()
    counts = host[: 8 * n_slots].view(np.int64)
    assert list(counts) == [n0] * n_slots
    sm.set_resident_sink(None, 0, 0)
    precs = pb.GpuAlleleCaller().Call(sm, raw=True)
    sm.close()
    want = sorted(bytes(r.tobytes()) for r in precs if int(r["ref_len"]) + int(r["alt_len"]) <= 4)
    for k in range(n_slots):
        recs = np.frombuffer(host[8 * n_slots + k * cap * 96: 8 * n_slots + (k * cap + n0) * 96].tobytes(), dtype=N.RECORD_DTYPE)
        pos = recs["position"].astype(np.int64)
        assert (np.diff(pos) >= 0).all()
        assert sorted(bytes(r.tobytes()) for r in recs if int(r["ref_len"]) + int(r["alt_len"]) <= 4) == want


def test_compact_offsets_equal_plain_push():
    """pb2_packed_read_batch with cigar_ops (one byte of operation count per read, offsets built on the device) stages the same reads as the offset
    arrays: records identical, over several batches cut out of larger arrays; operation counts that do not add up are refused."""
    pb = _pb()
    gen, cfg = CONFIGS["c2"]
    d = synth.make_reads(15000, 90, seed=17, **gen)
    sm = pb.GpuStateManager(pb.make_config(**cfg), "chr1", bytes(d["ref"]).decode())
    sm.AddReadsSoA(d)
    want = pb.GpuAlleleCaller().Call(sm, raw=True)
    n = len(d["pos0"])
    cuts = [0, n // 3, n // 3 + 1, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        part = dict(pos0=d["pos0"][a:b], flag=d["flag"][a:b], cigar_off=d["cigar_off"][a:b + 1], cigar=d["cigar"], seq_off=d["seq_off"][a:b + 1], bases=d["bases"], quals=d["quals"])
        pk = pb.GpuStateManager.pack_reads(part, compact=True)
        assert pk["cigar_off"] is None and len(pk["cigar_ops"]) == b - a
        sm.AddReadsPacked(pk)
    got = pb.GpuAlleleCaller().Call(sm, raw=True)
    assert len(want) > 100 and [bytes(r.tobytes()) for r in got if int(r["ref_len"]) + int(r["alt_len"]) <= 4] == \
        [bytes(r.tobytes()) for r in want if int(r["ref_len"]) + int(r["alt_len"]) <= 4]
    bad = pb.GpuStateManager.pack_reads(dict(pos0=d["pos0"][:50], flag=d["flag"][:50], cigar_off=d["cigar_off"][:51], cigar=d["cigar"], seq_off=d["seq_off"][:51],
                                             bases=d["bases"], quals=d["quals"]), compact=True)
    bad["cigar_ops"] = bad["cigar_ops"].copy()
    bad["cigar_ops"][7] += 1
    with pytest.raises(pb.PiscesB200Error, match="add up"):
        sm.AddReadsPacked(bad)
    sm.close()
