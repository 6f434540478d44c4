"""Seeded synthetic reads shared by the oracle and the CUDA path (test helper)."""
import numpy as np

from oracle import binding as ob

BASES = "ACGT"


def random_reference(rng, n, n_rate=0.0):
    seq = rng.choice(list(BASES), size=n)
    if n_rate > 0:
        seq[rng.random(n) < n_rate] = "N"
    return "".join(seq)


def make_reads(rng, ref, n_reads, read_len=60, start=1, span=None, snv_rate=0.01, lowq_rate=0.08, n_base_rate=0.005, del_rate=0.05, ins_rate=0.05,
               clip_rate=0.15, hotspots=None, stitched=False, collapsed=False, sorted_by_pos=True, indel_sites=0, linked_hotspots=False):
    """Returns a list of dicts: pos (1-based), seq, cigar (string), quals, flag, dirs (or None), coll (or None), xd/xr/xv/xw for the oracle.

    hotspots: {position: (alt_base, fraction)} real variants so that some calls pass the frequency / q-score bars."""
    span = span or (len(ref) - read_len - 12)
    reads = []
    # recurrent indels: site position (base before the event) -> (op, length, inserted bases, fraction of covering reads)
    sites = {}
    for p in rng.integers(start + 10, start + span, indel_sites) if indel_sites else []:
        ln = int(rng.integers(1, 5))
        sites[int(p)] = ("I" if rng.random() < 0.5 else "D", ln, "".join(BASES[int(x)] for x in rng.integers(0, 4, ln)), float(rng.uniform(0.05, 0.5)))
    for _ in range(n_reads):
        pos = int(rng.integers(start, start + span))
        ops = []
        ins_bases = None
        r = rng.random()
        u_link = rng.random()
        body = read_len
        lead_clip = int(rng.integers(1, 6)) if rng.random() < clip_rate else 0
        tail_clip = int(rng.integers(1, 6)) if rng.random() < clip_rate else 0
        body -= lead_clip + tail_clip
        if lead_clip:
            ops.append(("S", lead_clip))
        site = None
        for sp in range(pos, pos + body - 5):      # first recurrent site this read covers (may sit right at the read ends: open-ended candidates)
            if sp in sites and rng.random() < sites[sp][3]:
                site = sp
                break
        edge = rng.random() if site is not None else 1.0
        if site is not None and sites[site][0] == "I" and edge < 0.2 and not lead_clip and not tail_clip:
            # the read starts or ends inside the insertion: an open-ended insertion candidate carrying a suffix / prefix of the inserted bases
            kind, ln, ib, _ = sites[site]
            k = int(rng.integers(1, ln + 1))
            if edge < 0.1:
                pos = site + 1
                ops += [("I", k), ("M", body - k)]
                ins_bases = ib[ln - k:]
            else:
                pos = max(1, site - (body - k) + 1)
                ops += [("M", site - pos + 1), ("I", k)]
                ins_bases = ib[:k]
                body = site - pos + 1 + k
        elif site is not None:
            kind, ln, ib, _ = sites[site]
            a = site - pos + 1
            if kind == "D":
                ops += [("M", a), ("D", ln), ("M", body - a)]
            else:
                ln = min(ln, body - a)
                ops += [("M", a), ("I", ln)] + ([("M", body - a - ln)] if body - a - ln > 0 else [])
                ins_bases = ib[:ln]
        elif r < del_rate and body > 12:
            a = int(rng.integers(3, body - 6))
            ops += [("M", a), ("D", int(rng.integers(1, 4))), ("M", body - a)]
        elif r < del_rate + ins_rate and body > 12:
            a = int(rng.integers(3, body - 8))
            il = int(rng.integers(1, 4))
            ops += [("M", a), ("I", il), ("M", body - a - il)]
        else:
            ops.append(("M", body))
        if tail_clip:
            ops.append(("S", tail_clip))
        seq, rp = [], pos - 1
        for op, ln in ops:
            if op == "M":
                for k in range(ln):
                    b = ref[rp + k] if rp + k < len(ref) else "A"
                    p1 = rp + k + 1
                    if hotspots and p1 in hotspots and (u_link if linked_hotspots else rng.random()) < hotspots[p1][1]:
                        b = hotspots[p1][0]
                    elif rng.random() < snv_rate:
                        b = BASES[int(rng.integers(0, 4))]
                    if rng.random() < n_base_rate:
                        b = "N"
                    seq.append(b)
                rp += ln
            elif op == "D":
                rp += ln
            elif op == "I" and ins_bases is not None:
                seq += list(ins_bases)
            else:  # I, S
                seq += [BASES[int(rng.integers(0, 4))] for _ in range(ln)]
        quals = np.where(rng.random(len(seq)) < lowq_rate, rng.integers(2, 20, len(seq)), rng.integers(20, 41, len(seq))).astype(int).tolist()
        reverse = bool(rng.random() < 0.5)
        flag = (0x10 if reverse else 0) | 0x1 | 0x2 | (0x40 if rng.random() < 0.5 else 0x80)
        rd = dict(pos=pos, seq="".join(seq), cigar="".join(f"{ln}{op}" for op, ln in ops), quals=quals, flag=flag, dirs=None, coll=None,
                  xd=None, xr=None, xv=None, xw=None)
        if stitched and rng.random() < 0.6:
            # XD direction string over the cigar-expanded alignment: F.. S.. R..
            total = sum(ln for _, ln in ops)
            a = int(rng.integers(1, total // 2))
            b = int(rng.integers(1, total - a))
            rd["xd"] = f"{a}F{b}S{total - a - b}R" if total - a - b > 0 else f"{a}F{b}S"
            exp = [0] * a + [2] * b + [1] * (total - a - b)
            dirs, ci = [], 0
            for op, ln in ops:
                for _k in range(ln):
                    if op in "MIS":
                        dirs.append(exp[ci])
                    ci += 1
            rd["dirs"] = dirs
        if collapsed:
            xv, xw = int(rng.integers(0, 4)), int(rng.integers(0, 3))
            xr = ["FR", "RF", "FF"][int(rng.integers(0, 3))]
            rd.update(xv=xv, xw=xw, xr=xr)
            rd["coll"] = 1 | (2 if (xv != 0 and xw != 0) else 0) | ({"FR": 1, "RF": 2}.get(xr, 0) << 2)
        reads.append(rd)
    if sorted_by_pos:
        reads.sort(key=lambda r: r["pos"])
    return reads


def to_oracle(rd):
    return ob.SimpleRead(rd["pos"], rd["seq"], rd["cigar"], rd["quals"], flag=rd["flag"], xd=rd["xd"], xr=rd["xr"], xv=rd["xv"], xw=rd["xw"],
                         has_tags=any(rd[k] is not None for k in ("xd", "xr", "xv", "xw")))


def to_product(rd):
    import pisces_b200 as pb
    return pb.Read(rd["pos"], rd["seq"], rd["cigar"], rd["quals"], flag=rd["flag"], base_directions=rd["dirs"], collapsed=rd["coll"])


def oracle_cfg_from(**kw):
    """Oracle config with the same knobs the product config takes (names of po_config)."""
    return ob.default_config(**kw)


def make_amplicon_reads(seed=0, n_amp=4, per_amp=200, read_len=140, stride=100, untagged=30, jitter=2, variants=None):
    """Amplicon-shaped reads as a struct of arrays (pb2_read_batch layout + "amplicon", the XN name as an id, -1 = no tag): amplicon a covers about
    [20 + a * stride, +read_len), neighbours overlap by read_len - stride. All reads are one 'M' run. variants: list of (position 1-based, {amplicon id: vaf})
    (key -1: the untagged reads); the alt base is the reference base's successor in ACGT. Reads come out sorted by position, amplicons interleaved."""
    rng = np.random.default_rng(seed)
    L = 20 + (n_amp - 1) * stride + read_len + jitter + 40
    ref = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, L)].copy()
    starts, names = [], []
    for a in range(n_amp):
        starts.append(20 + a * stride + rng.integers(0, jitter + 1, per_amp))
        names.append(np.full(per_amp, a))
    if untagged:
        starts.append(rng.integers(20, L - read_len - 1, untagged))
        names.append(np.full(untagged, -1))
    starts = np.concatenate(starts).astype(np.int32)
    names = np.concatenate(names).astype(np.int32)
    order = np.argsort(starts, kind="stable")
    starts, names = starts[order], names[order]
    n = len(starts)
    idx = starts[:, None] + np.arange(read_len)[None, :]          # 0-based reference index of every base
    bases = ref[idx].copy()
    nxt = np.zeros(256, dtype=np.uint8)
    for x, y in zip(b"ACGT", b"CGTA"):
        nxt[x] = y
    for pos, vafs in (variants or []):
        col = pos - 1 - starts
        for name, vaf in vafs.items():
            hit = (names == name) & (col >= 0) & (col < read_len) & (rng.random(n) < vaf)
            rows = np.nonzero(hit)[0]
            bases[rows, col[rows]] = nxt[ref[pos - 1]]
    quals = rng.integers(30, 41, size=bases.shape).astype(np.uint8)
    low = rng.random(bases.shape) < 0.06
    quals[low] = rng.integers(2, 20, size=int(low.sum()))
    err = rng.random(bases.shape) < 0.002
    bases[err] = nxt[bases[err]]
    flag = np.where(rng.random(n) < 0.5, 0x10, 0).astype(np.uint16)
    return dict(ref=ref, pos0=starts, flag=flag, cigar_off=np.arange(n + 1, dtype=np.int64), cigar=np.full(n, (read_len << 4) | 0, dtype=np.uint32),
                seq_off=np.arange(n + 1, dtype=np.int64) * read_len, bases=bases.reshape(-1), quals=quals.reshape(-1), amplicon=names, n_loci=L)


def amplicon_counts_at(d, pos, min_bq=20, alt=None):
    """Per-amplicon (name -> count) of the reads' usable bases at 1-based pos (RegionState.AddAmpliconCount: every base counted as A/C/G/T at quality >=
    min_bq), in first-seen order; with alt: only the bases equal to it (the support of the SNV candidate)."""
    L = int(d["seq_off"][1] - d["seq_off"][0])
    col = pos - 1 - d["pos0"]
    out = {}
    for r in np.nonzero((col >= 0) & (col < L))[0]:
        k = int(d["seq_off"][r]) + int(col[r])
        b, q, a = int(d["bases"][k]), int(d["quals"][k]), int(d["amplicon"][r])
        if a < 0 or q < min_bq or b not in b"ACGT" or (alt is not None and b != alt):
            continue
        out[a] = out.get(a, 0) + 1
    return out
