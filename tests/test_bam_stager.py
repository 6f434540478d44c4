"""The library's BAM stager (pb2_bam_*: BGZF inflate, record decode, AlignmentSource.ShouldSkipRead, XD / XV / XW / XR) against the independent
Python decoder of the test suite (tests/bamio.py) on the reference's own BAM test files (copies under tests/golden/). Host code: runs without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from pisces_b200 import _native as N
from tests import bamio

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _batches(path, **kw):
    import pisces_b200 as pb
    st = pb.BamReadStager(path, **kw)
    out = []
    for ref_id, b, skipped in st:
        n = b.n_reads
        pos0 = np.ctypeslib.as_array(C.cast(b.pos0, C.POINTER(C.c_int32)), (n,)).copy()
        flag = np.ctypeslib.as_array(C.cast(b.flag, C.POINTER(C.c_uint16)), (n,)).copy()
        coff = np.ctypeslib.as_array(C.cast(b.cigar_off, C.POINTER(C.c_int64)), (n + 1,)).copy()
        soff = np.ctypeslib.as_array(C.cast(b.seq_off, C.POINTER(C.c_int64)), (n + 1,)).copy()
        cigar = np.ctypeslib.as_array(C.cast(b.cigar, C.POINTER(C.c_uint32)), (int(coff[-1]),)).copy()
        bases = bytes(np.ctypeslib.as_array(C.cast(b.bases, C.POINTER(C.c_uint8)), (int(soff[-1]),)))
        quals = np.ctypeslib.as_array(C.cast(b.quals, C.POINTER(C.c_uint8)), (int(soff[-1]),)).copy()
        dirs = np.ctypeslib.as_array(C.cast(b.base_dirs, C.POINTER(C.c_uint8)), (int(soff[-1]),)).copy() if b.base_dirs else None
        coll = np.ctypeslib.as_array(C.cast(b.collapsed, C.POINTER(C.c_uint8)), (n,)).copy() if b.collapsed else None
        out.append(dict(ref_id=ref_id, skipped=skipped, pos0=pos0, flag=flag, coff=coff, soff=soff, cigar=cigar, bases=bases, quals=quals, dirs=dirs, coll=coll))
    refs, flags = st.references, (st.is_stitched, st.is_collapsed)
    st.close()
    return out, refs, flags


def _kept(recs):   # AlignmentSource.ShouldSkipRead (AlignmentsSource.cs:84-92) with the default BamFilterParameters
    return [r for r in recs if not (r["flag"] & 0x4) and not (r["flag"] & 0x100) and not (r["flag"] & 0x400) and r["mapq"] >= 1 and r["cigar"] and r["ref_id"] >= 0]


def test_phix_bam_matches_python_decoder():
    text, refs, recs = bamio.read_bam(os.path.join(G, "PhiX_S3.bam"))
    for max_reads in (65536, 37):     # one batch, and many small ones
        batches, brefs, _ = _batches(os.path.join(G, "PhiX_S3.bam"), max_reads=max_reads)
        assert brefs == [(n, l) for n, l in refs]
        kept = _kept(recs)
        assert sum(b["skipped"] for b in batches) == len(recs) - len(kept) and sum(len(b["pos0"]) for b in batches) == len(kept)
        i = 0
        for b in batches:
            # always handed out; without XD / XV / XW tags they hold what the flags imply
            assert b["dirs"] is not None and b["coll"] is not None and not (b["coll"] & 1).any()
            for k in range(len(b["pos0"])):
                assert set(b["dirs"][b["soff"][k]:b["soff"][k + 1]].tolist()) <= {1 if int(b["flag"][k]) & 0x10 else 0}
                r = kept[i]
                assert (int(b["pos0"][k]), int(b["flag"][k])) == (r["pos0"], r["flag"])
                assert list(b["cigar"][b["coff"][k]:b["coff"][k + 1]]) == r["cigar"]
                assert b["bases"][b["soff"][k]:b["soff"][k + 1]].decode() == r["seq"]
                assert list(b["quals"][b["soff"][k]:b["soff"][k + 1]]) == r["qual"]
                i += 1


def test_collapsed_stitched_bam_tags():
    _, _, recs = bamio.read_bam(os.path.join(G, "collapsed.test.stitched.bam"))
    batches, _, _ = _batches(os.path.join(G, "collapsed.test.stitched.bam"))
    kept = _kept(recs)
    assert len(batches) == 1 and len(batches[0]["pos0"]) == len(kept)
    b = batches[0]
    assert b["dirs"] is not None and b["coll"] is not None
    for k, r in enumerate(kept):
        t = r["tags"]
        # Read.SequencedBaseDirectionMap (Read.cs:390-421,664-682): XD runs along the expanded CIGAR, kept where the operation spans the read
        exp, num = [], ""
        for ch in t["XD"]:
            if ch.isdigit():
                num += ch
            else:
                exp += [{"F": 0, "R": 1, "S": 2}[ch]] * int(num)
                num = ""
        dirs, ci = [], 0
        for c in r["cigar"]:
            for _ in range(c >> 4):
                if (c & 15) in (0, 1, 4, 7, 8):
                    dirs.append(exp[ci])
                ci += 1
        assert list(b["dirs"][b["soff"][k]:b["soff"][k + 1]]) == dirs
        has = "XV" in t or "XW" in t
        duplex = bool(t.get("XV")) and bool(t.get("XW"))
        assert int(b["coll"][k]) == (1 if has else 0) | (2 if duplex else 0) | ({"FR": 1, "RF": 2}.get(t.get("XR"), 0) << 2)


def test_amplicon_names_from_the_xn_tag():
    """Read.GetAmpliconNameIfExists (Read.cs:479-486) through pb2_bam_batch_amplicons / pb2_bam_amplicon_names on the mapped reads of the reference's
    testdata/example_S1.bam (tests/golden/example_S1.mapped.bam, written by tests/golden/make_amplicon_fixture.py): per-read names equal the XN tags the
    Python decoder reads, the dictionary is in first-seen order over the kept reads, across batch boundaries."""
    import pisces_b200 as pb
    path = os.path.join(G, "example_S1.mapped.bam")
    _, _, recs = bamio.read_bam(path)
    kept = _kept(recs)
    assert len(kept) > 300 and all("XN" in r["tags"] for r in kept)
    for max_reads in (65536, 50):
        st = pb.BamReadStager(path, max_reads=max_reads)
        got = []
        for _ref_id, b, _skipped in st:
            ids = st.batch_amplicons()
            assert len(ids) == b.n_reads
            names = st.amplicon_names()
            got += [names[i] if i >= 0 else None for i in ids]
        names = st.amplicon_names()
        st.close()
        assert got == [r["tags"]["XN"] for r in kept]
        first_seen = list(dict.fromkeys(got))
        assert names == first_seen and len(names) == 8
    # a file without the tag: every id is -1, the dictionary stays empty
    st = pb.BamReadStager(os.path.join(G, "PhiX_S3.bam"))
    for _ref_id, b, _skipped in st:
        assert set(st.batch_amplicons()) == {-1}
    assert st.amplicon_names() == []
    st.close()


def _bgzf_block(payload, bsize_override=None, isize_override=None):
    import struct
    import zlib
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = co.compress(payload) + co.flush()
    bsize = len(comp) + 25 if bsize_override is None else bsize_override
    isize = len(payload) if isize_override is None else isize_override
    return (b"\x1f\x8b\x08\x04" + b"\0" * 6 + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize) + comp +
            struct.pack("<II", zlib.crc32(payload) & 0xffffffff, isize))


def _bam_bytes(tag_bytes=b"", l_seq=4):
    import struct
    header = b"BAM\1" + struct.pack("<i", 0) + struct.pack("<i", 1) + struct.pack("<i", 5) + b"chr1\0" + struct.pack("<i", 1000)
    name = b"r\0"
    rec = struct.pack("<iiIIiiii", 0, 10, (len(name)) | (60 << 8) | (4680 << 16), 1 | (0 << 16), l_seq, -1, -1, 0) + name + struct.pack("<I", (l_seq << 4) | 0) + \
        bytes([0x12] * ((l_seq + 1) // 2)) + bytes([30] * l_seq) + tag_bytes
    return header + struct.pack("<i", len(rec)) + rec


@pytest.mark.parametrize("case", ["bsize_underflow", "isize_huge", "tag_past_end", "string_unterminated", "b_array_count", "xd_huge_run"])
def test_malformed_bam_is_rejected_not_crashed(case, tmp_path):
    """block sizes, tag values and XD run lengths are validated against the record before anything is allocated or read."""
    import struct
    L = N.load()
    good = _bam_bytes()
    if case == "bsize_underflow":
        data = _bgzf_block(good, bsize_override=5)
    elif case == "isize_huge":
        data = _bgzf_block(good, isize_override=1 << 30)
    elif case == "tag_past_end":
        data = _bgzf_block(_bam_bytes(b"XVi\x01\x00"))            # an int32 with two of its four bytes
    elif case == "string_unterminated":
        data = _bgzf_block(_bam_bytes(b"XDZ4F"))                    # no NUL
    elif case == "b_array_count":
        data = _bgzf_block(_bam_bytes(b"XBBi" + struct.pack("<i", 1 << 28)))
    else:
        data = _bgzf_block(_bam_bytes(b"XDZ4000000000F\0"))         # a direction run of four billion: clamped to the CIGAR length, read kept
    path = tmp_path / "bad.bam"
    path.write_bytes(data + _bgzf_block(b""))
    rd = C.c_void_p()
    rc = L.pb2_bam_open(str(path).encode(), C.byref(rd))
    if rc != 0:
        return   # rejected at open
    b, ref_id, skipped = N.ReadBatch(), C.c_int32(), C.c_int64()
    rc = L.pb2_bam_next_batch(rd, None, 100, C.byref(b), C.byref(ref_id), C.byref(skipped))
    if case == "xd_huge_run":
        assert rc == 0 and b.n_reads == 1
    else:
        assert rc != 0 and L.pb2_bam_last_error(rd)
    L.pb2_bam_close(rd)


def test_packed_batches_hold_the_same_reads():
    """pb2_bam_next_batch_packed against pb2_bam_next_batch on the reference's BAM fixtures (CPU: no device work): the packed bytes + exception list unpack
    to the same bases and qualities, the operation counts are the offsets' differences, direction / collapsed planes come only with tagged reads."""
    import ctypes as C
    import pisces_b200 as pb
    lut = np.frombuffer(b"AGCT", dtype=np.uint8)

    def arr(ptr, n, dt):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(dt)), shape=(n,)).copy() if ptr and n else np.zeros(0, dtype=dt)
    for name in ("PhiX_S3.bam", "collapsed.test.stitched.bam", "example_S1.mapped.bam"):
        path = os.path.join(G, name)
        plain = pb.BamReadStager(path, max_reads=97)
        packed = pb.BamReadStager(path, max_reads=97, packed=True)
        n_batches = 0
        for (ra, a, sa), (rb, b, sb) in zip(plain, packed):
            n = a.n_reads
            assert (ra, sa, n) == (rb, sb, b.n_reads)
            coff, soff = arr(a.cigar_off, n + 1, C.c_int64), arr(a.seq_off, n + 1, C.c_int64)
            nseq, ncig = int(soff[-1]), int(coff[-1])
            assert np.array_equal(arr(a.pos0, n, C.c_int32), arr(b.pos0, n, C.c_int32)) and np.array_equal(arr(a.flag, n, C.c_uint16), arr(b.flag, n, C.c_uint16))
            assert np.array_equal(arr(a.cigar, ncig, C.c_uint32), arr(b.cigar, ncig, C.c_uint32))
            assert not b.cigar_off and not b.seq_off and (b.n_cigar_total, b.n_seq_total) == (ncig, nseq)
            assert np.array_equal(arr(b.cigar_ops, n, C.c_uint8), np.diff(coff).astype(np.uint8))
            seq = arr(b.seq, nseq, C.c_uint8)
            bases, quals = lut[seq >> 6].copy(), (seq & 63).copy()
            ei = arr(b.exc_index, b.n_exceptions, C.c_int64)
            bases[ei], quals[ei] = arr(b.exc_base, b.n_exceptions, C.c_uint8), arr(b.exc_qual, b.n_exceptions, C.c_uint8)
            assert np.array_equal(bases, arr(a.bases, nseq, C.c_uint8)) and np.array_equal(quals, arr(a.quals, nseq, C.c_uint8))
            tagged = name.startswith("collapsed")
            assert bool(b.base_dirs) == tagged and bool(b.collapsed) == tagged
            if tagged:
                assert np.array_equal(arr(a.base_dirs, nseq, C.c_uint8), arr(b.base_dirs, nseq, C.c_uint8))
            n_batches += 1
        assert n_batches >= 1
        plain.close(); packed.close()
