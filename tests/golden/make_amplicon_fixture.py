#!/usr/bin/env python3
"""Writes tests/golden/example_S1.mapped.bam: the header and the 482 mapped records of the reference's testdata/example_S1.bam (138,230 records, 10 MB;
SURVEY 8c), re-blocked into a small BGZF file. Record bytes are copied unchanged, so the XN (amplicon name) tags are the reference's own.
Run in the build container, where /root/reference is mounted."""
import os
import struct
import sys
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import bamio  # noqa: E402

SRC = "/root/reference/testdata/example_S1.bam"


def bgzf_block(data):
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = comp.compress(data) + comp.flush()
    head = struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, ord("B"), ord("C"), 2, len(body) + 25)
    return head + body + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data))


def main():
    raw = bamio._bgzf_blocks(SRC)
    assert raw[:4] == b"BAM\1"
    l_text, = struct.unpack_from("<i", raw, 4)
    o = 8 + l_text
    n_ref, = struct.unpack_from("<i", raw, o)
    o += 4
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", raw, o)
        o += 4 + l_name + 4
    out = [raw[:o]]
    kept = total = 0
    while o < len(raw):
        bs, = struct.unpack_from("<i", raw, o)
        flag = struct.unpack_from("<I", raw, o + 4 + 12)[0] >> 16
        total += 1
        if not (flag & 0x4):
            out.append(raw[o:o + 4 + bs])
            kept += 1
        o += 4 + bs
    data = b"".join(out)
    with open(os.path.join(HERE, "example_S1.mapped.bam"), "wb") as f:
        for i in range(0, len(data), 60000):
            f.write(bgzf_block(data[i:i + 60000]))
        f.write(bgzf_block(b""))
    print(f"{kept} of {total} records kept, {len(data)} bytes inflated")


if __name__ == "__main__":
    main()
