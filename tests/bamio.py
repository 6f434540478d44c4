"""Minimal BAM / FASTA readers for the tests (test infrastructure; the product never parses BAM — SURVEY.md §2 rows 9-10 are out of scope).

BAM: BGZF blocks inflated with zlib, records decoded per the SAM spec the reference follows (src/lib/Alignment.IO/BamReader.cs:137-224:
4-bit bases through "=ACMGRSVTWYHKDBN", CIGAR op count = low 16 bits of flag_nc)."""
import struct
import zlib

_SEQ = "=ACMGRSVTWYHKDBN"


def _bgzf_blocks(path):
    data = open(path, "rb").read()
    out, p = [], 0
    while p < len(data):
        assert data[p:p + 4] == b"\x1f\x8b\x08\x04", "not BGZF"
        xlen = struct.unpack_from("<H", data, p + 10)[0]
        bsize = None
        q = p + 12
        while q < p + 12 + xlen:
            si1, si2, slen = data[q], data[q + 1], struct.unpack_from("<H", data, q + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", data, q + 4)[0]
            q += 4 + slen
        cdata = data[p + 12 + xlen: p + bsize + 1 - 8]
        out.append(zlib.decompress(cdata, -15))
        p += bsize + 1
    return b"".join(out)


def _tags(buf):
    tags, p = {}, 0
    sizes = {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4, "A": 1}
    fmts = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I", "f": "<f"}
    while p + 3 <= len(buf):
        tag, t = buf[p:p + 2].decode(), chr(buf[p + 2])
        p += 3
        if t == "Z" or t == "H":
            e = buf.index(b"\0", p)
            tags[tag] = buf[p:e].decode()
            p = e + 1
        elif t == "A":
            tags[tag] = chr(buf[p])
            p += 1
        elif t == "B":
            st, n = chr(buf[p]), struct.unpack_from("<i", buf, p + 1)[0]
            p += 5 + sizes[st] * n
        else:
            tags[tag] = struct.unpack_from(fmts[t], buf, p)[0]
            p += sizes[t]
    return tags


def read_bam(path):
    """Returns (header_text, [(name, length)], [record dict]) with record keys: ref_id, pos0, mapq, flag, cigar (BAM-encoded uint32 list),
    seq, qual (list of int), tags (dict), name."""
    d = _bgzf_blocks(path)
    assert d[:4] == b"BAM\1"
    l_text = struct.unpack_from("<i", d, 4)[0]
    text = d[8:8 + l_text].decode(errors="replace")
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", d, p)[0]
    p += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", d, p)[0]
        name = d[p + 4:p + 4 + l_name - 1].decode()
        refs.append((name, struct.unpack_from("<i", d, p + 4 + l_name)[0]))
        p += 8 + l_name
    recs = []
    while p + 4 <= len(d):
        bs = struct.unpack_from("<i", d, p)[0]
        r = d[p + 4:p + 4 + bs]
        p += 4 + bs
        ref_id, pos0, bin_mq_nl, flag_nc, l_seq, _mrid, _mpos, _tlen = struct.unpack_from("<iiIIiiii", r, 0)
        l_name, mapq = bin_mq_nl & 0xff, (bin_mq_nl >> 8) & 0xff
        flag, n_cig = flag_nc >> 16, flag_nc & 0xffff
        q = 32
        name = r[q:q + l_name - 1].decode()
        q += l_name
        cigar = list(struct.unpack_from(f"<{n_cig}I", r, q))
        q += 4 * n_cig
        sb = r[q:q + (l_seq + 1) // 2]
        seq = "".join(_SEQ[(sb[i >> 1] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq))
        q += (l_seq + 1) // 2
        qual = list(r[q:q + l_seq])
        q += l_seq
        recs.append(dict(ref_id=ref_id, pos0=pos0, mapq=mapq, flag=flag, cigar=cigar, seq=seq, qual=qual, tags=_tags(r[q:]), name=name))
    return text, refs, recs


def read_fasta(path):
    """{name: upper-case sequence} (Genome.cs:81-96 upper-cases on load)."""
    seqs, name, parts = {}, None, []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            if name is not None:
                seqs[name] = "".join(parts).upper()
            name, parts = line[1:].split()[0], []
        elif line:
            parts.append(line)
    if name is not None:
        seqs[name] = "".join(parts).upper()
    return seqs
