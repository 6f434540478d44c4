"""Host-side helpers of the library that do no device work (callable without a GPU): pb2_pack_pileup and pb2_pack_reads against numpy restatements."""
import numpy as np
import pytest


def _numpy_pack_pileup(code, qual, anchor, offsets=None, ref_bases=None):
    flag_index = np.flatnonzero(code & 0xe0).astype(np.int64)
    if offsets is not None and ref_bases is not None and len(flag_index):
        locus = np.searchsorted(np.asarray(offsets, dtype=np.int64), flag_index, side="right") - 1
        ref_allele = np.full(256, 4, dtype=np.uint8)
        ref_allele[[ord("A"), ord("G"), ord("C"), ord("T")]] = [0, 1, 2, 3]
        flag_index = flag_index[(code[flag_index] & 7) != ref_allele[np.asarray(ref_bases, dtype=np.uint8)[locus]]]
    return (((code & 0x1f) | ((anchor & 7) << 5)).astype(np.uint8), ((qual & 0x7f) | ((anchor >> 3) << 7)).astype(np.uint8), flag_index,
            (code[flag_index] & 0xe0).astype(np.uint8))


def test_pack_pileup_is_the_library_and_equals_numpy():
    import pisces_b200 as pb
    rng = np.random.default_rng(0)
    n_loci = 700
    off = np.concatenate([[0], np.cumsum(rng.integers(0, 60, n_loci))]).astype(np.int64)   # some loci empty
    n = int(off[-1])
    code = (rng.integers(0, 6, n) | (rng.integers(0, 3, n) << 3) | np.where(rng.random(n) < 0.05, rng.integers(1, 8, n) << 5, 0)).astype(np.uint8)
    qual = rng.integers(0, 64, n).astype(np.uint8)
    anchor = rng.integers(0, 11, n).astype(np.uint8)
    refb = np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, n_loci)]
    for args in ((code, qual, anchor), (code, qual, anchor, off, refb)):
        got, want = pb.GpuStateManager.pack_pileup(*args), _numpy_pack_pileup(*args)
        assert len(want[2]) > 100 and all(np.array_equal(x, y) for x, y in zip(got, want))
    with pytest.raises(ValueError, match="collapsed-read type"):
        pb.GpuStateManager.pack_pileup(code, qual, (anchor | 16).astype(np.uint8))


def test_pack_reads_round_trip_and_compact_offsets():
    import pisces_b200 as pb
    rng = np.random.default_rng(1)
    n_reads, L = 300, 50
    bases = np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.choice(5, n_reads * L, p=[0.24, 0.24, 0.24, 0.24, 0.04])].copy()
    quals = rng.integers(0, 80, n_reads * L).astype(np.uint8)
    d = dict(pos0=np.arange(n_reads, dtype=np.int32), flag=np.zeros(n_reads, dtype=np.uint16), cigar_off=np.arange(n_reads + 1, dtype=np.int64),
             cigar=np.full(n_reads, (L << 4), dtype=np.uint32), seq_off=np.arange(n_reads + 1, dtype=np.int64) * L, bases=bases, quals=quals)
    pk = pb.GpuStateManager.pack_reads(d, compact=True)
    lut = np.frombuffer(b"AGCT", dtype=np.uint8)
    ub, uq = lut[pk["seq"] >> 6].copy(), (pk["seq"] & 63).copy()
    ub[pk["exc_index"]], uq[pk["exc_index"]] = pk["exc_base"], pk["exc_qual"]
    assert np.array_equal(ub, bases) and np.array_equal(uq, quals)          # lossless
    exc = (bases == ord("N")) | (quals > 63) | ((bases == ord("A")) & (quals == 0))
    assert np.array_equal(pk["exc_index"], np.flatnonzero(exc))
    assert pk["cigar_off"] is None and pk["seq_off"] is None and np.array_equal(pk["cigar_ops"], np.ones(n_reads, dtype=np.uint8))
    assert (pk["n_cigar_total"], pk["n_seq_total"]) == (n_reads, n_reads * L)
