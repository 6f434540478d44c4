"""Seeded synthetic locus-major pileups of the BASELINE.json shapes, generated with torch on the CPU or straight
into HBM. Data generator only — no part of the calling path.

Entry model: read length 140 (example_S1-shaped), the entry's offset inside its read is uniform in [0,139] which fixes the anchor bin
(RegionStateManager.GetAnchorType) and the open-end flags of the first/last base; direction Bernoulli(0.5); base quality from an
example_S1-like histogram (Q38 60 %, Q34-37 20 %, Q20-33 11 %, Q10-19 9 %); a base is mis-called with probability 10^(-q/10) to a
uniform other base; variant loci carry an alternate base at VAF ~ U[vaf_lo, vaf_hi]; deletion loci carry Deletion entries.
"""
import math

import torch

READ_LEN = 140
_ALLELE_OF_ACGT = torch.tensor([0, 2, 1, 3], dtype=torch.uint8)   # A,C,G,T -> AlleleType A=0,C=2,G=1,T=3
_ASCII_ACGT = torch.tensor([65, 67, 71, 84], dtype=torch.uint8)


def _quality_table():
    p = torch.zeros(64, dtype=torch.float64)
    p[38] = 0.60
    p[34:38] = 0.05
    p[20:34] = 0.11 / 14
    p[10:20] = 0.09 / 10
    return torch.cumsum(p / p.sum(), 0)


def make_pileup(n_loci, mean_depth, seed=1, device="cpu", depth_dist="poisson", snv_rate=0.01, vaf=(0.01, 0.5), del_rate=0.001, stitched_frac=0.0,
                collapsed_frac=0.0, strand_skew_frac=0.0, chunk_entries=1 << 25, flags=True, indel_rate=0.0, min_bq=20):
    """Returns dict(offsets int64[n+1], code/qual/anchor uint8[total], ref_bases uint8[n] ASCII, snv_loci, del_loci) and, when indel_rate > 0,
    candidates (numpy structured array, pb2_candidate layout) + arena (bytes): 1-3 bp insertions and deletions at indel_rate of the loci
    (0.1 %). A deletion candidate at position p puts Deletion entries on loci p+1..p+len at its VAF and takes its support from
    the usable Deletion entries of locus p+1 (one per read carrying it, as CandidateVariantFinder would count them); an insertion has no
    pileup entries (inserted bases are not mapped), its support is Binomial(depth, VAF) split by direction."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    if depth_dist == "poisson":
        depth = torch.poisson(torch.full((n_loci,), float(mean_depth), device=dev), generator=g).to(torch.int64)
    else:
        depth = torch.full((n_loci,), int(mean_depth), device=dev, dtype=torch.int64)
    offsets = torch.zeros(n_loci + 1, dtype=torch.int64, device=dev)
    torch.cumsum(depth, 0, out=offsets[1:])
    total = int(offsets[-1])
    ref_idx = torch.randint(0, 4, (n_loci,), device=dev, generator=g)
    is_snv = torch.rand(n_loci, device=dev, generator=g) < snv_rate
    alt_idx = (ref_idx + torch.randint(1, 4, (n_loci,), device=dev, generator=g)) % 4
    locus_vaf = vaf[0] + (vaf[1] - vaf[0]) * torch.rand(n_loci, device=dev, generator=g)
    locus_vaf = torch.where(is_snv, locus_vaf, torch.zeros_like(locus_vaf))
    is_del = (torch.rand(n_loci, device=dev, generator=g) < del_rate) & ~is_snv
    del_vaf = torch.where(is_del, 0.05 + 0.3 * torch.rand(n_loci, device=dev, generator=g), torch.zeros(n_loci, device=dev))
    indel_pos = indel_len = indel_is_ins = indel_vaf = None
    if indel_rate > 0:
        # candidate position p = the base before the event; keep events apart (>= 8 loci) and inside the staged range
        cand = torch.nonzero(torch.rand(n_loci, device=dev, generator=g) < indel_rate).flatten()
        cand = cand[(cand >= 1) & (cand < n_loci - 8)]
        if cand.numel() > 1:
            keep = torch.ones_like(cand, dtype=torch.bool)
            keep[1:] = (cand[1:] - cand[:-1]) >= 8
            cand = cand[keep]
        k = cand.numel()
        indel_pos = cand
        indel_len = torch.randint(1, 4, (k,), device=dev, generator=g)
        indel_is_ins = torch.rand(k, device=dev, generator=g) < 0.5
        indel_vaf = 0.03 + 0.4 * torch.rand(k, device=dev, generator=g)
        for j in range(1, 4):   # Deletion entries on the deleted positions p+1..p+len
            m = (~indel_is_ins) & (indel_len >= j)
            del_vaf[cand[m] + j] = indel_vaf[m]
    skew = torch.rand(n_loci, device=dev, generator=g) < strand_skew_frac   # variant support 90/10 across strands (SB filter fires)

    code = torch.empty(total, dtype=torch.uint8, device=dev)
    qual = torch.empty(total, dtype=torch.uint8, device=dev)
    anchor = torch.empty(total, dtype=torch.uint8, device=dev)
    qcdf = _quality_table().to(dev)
    allele_of = _ALLELE_OF_ACGT.to(dev)

    # chunk over loci so that temporaries stay bounded
    loci_per_chunk = max(1, int(chunk_entries // max(1, int(mean_depth))))
    for l0 in range(0, n_loci, loci_per_chunk):
        l1 = min(n_loci, l0 + loci_per_chunk)
        e0, e1 = int(offsets[l0]), int(offsets[l1])
        n = e1 - e0
        if n == 0:
            continue
        locus = torch.repeat_interleave(torch.arange(l0, l1, device=dev), depth[l0:l1])
        u = torch.rand(n, device=dev, generator=g)
        q = torch.searchsorted(qcdf, u.to(torch.float64)).clamp_(0, 63).to(torch.int64)
        base = ref_idx[locus]
        # real variant
        is_alt = torch.rand(n, device=dev, generator=g) < locus_vaf[locus]
        base = torch.where(is_alt, alt_idx[locus], base)
        # sequencing error
        perr = torch.pow(10.0, -q.to(torch.float32) / 10.0)
        is_err = torch.rand(n, device=dev, generator=g) < perr
        base = torch.where(is_err, (base + torch.randint(1, 4, (n,), device=dev, generator=g)) % 4, base)
        allele = allele_of[base].to(torch.int64)
        deleted = torch.rand(n, device=dev, generator=g) < del_vaf[locus]
        allele = torch.where(deleted, torch.full_like(allele, 5), allele)
        rnd_dir = torch.rand(n, device=dev, generator=g)
        direction = (rnd_dir < 0.5).to(torch.int64)                      # 0 forward / 1 reverse
        sk = skew[locus] & is_alt
        direction = torch.where(sk, (rnd_dir < 0.1).to(torch.int64), direction)
        if stitched_frac > 0:
            direction = torch.where(torch.rand(n, device=dev, generator=g) < stitched_frac, torch.full_like(direction, 2), direction)
        off_in_read = torch.randint(0, READ_LEN, (n,), device=dev, generator=g)
        left, right = off_in_read, READ_LEN - 1 - off_in_read
        abin = torch.where(left >= right, torch.where(right >= 5, torch.full_like(left, 5), 10 - right), torch.where(left >= 5, torch.full_like(left, 5), left))
        c = allele | (direction << 3)
        if flags:
            c = c | torch.where((left == 0) & ~deleted, 0x20, 0) | torch.where((right == 0) & ~deleted, 0x40, 0)
        ab = abin
        if collapsed_frac > 0:
            is_c = torch.rand(n, device=dev, generator=g) < collapsed_frac
            duplex = torch.rand(n, device=dev, generator=g) < 0.2
            fr = torch.rand(n, device=dev, generator=g) < 0.5
            st = direction == 2
            ctype = torch.where(duplex, torch.where(st, 0, 1), torch.where(fr, torch.where(st, 4, 5), torch.where(st, 6, 7)))
            ab = ab | torch.where(is_c, (ctype + 1) << 4, 0)
        code[e0:e1] = c.to(torch.uint8)
        qual[e0:e1] = q.to(torch.uint8)
        anchor[e0:e1] = ab.to(torch.uint8)
    ref_bases = _ASCII_ACGT.to(dev)[ref_idx]
    out = dict(offsets=offsets, code=code, qual=qual, anchor=anchor, ref_bases=ref_bases, n_entries=total,
               snv_loci=int(is_snv.sum()), del_loci=int(is_del.sum()))
    if indel_pos is not None:
        import numpy as np
        from . import _native as N
        k = int(indel_pos.numel())
        cands = np.zeros(k, dtype=N.CANDIDATE_DTYPE)
        arena = bytearray()
        refb = ref_bases.cpu().numpy()
        pos_h, len_h, ins_h, vaf_h = (t.cpu().numpy() for t in (indel_pos, indel_len, indel_is_ins, indel_vaf))
        depth_h = depth[indel_pos].cpu().numpy()
        off_h = offsets.cpu()
        rng = np.random.default_rng(int(seed) + 7919)
        for i in range(k):
            p, L = int(pos_h[i]), int(len_h[i])          # locus index p  <->  reference position p + 1
            c = cands[i]
            c["position"] = p + 1
            c["allele_offset"] = len(arena)
            if ins_h[i]:
                ins = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L).tolist())
                ref_s, alt_s = bytes([refb[p]]), bytes([refb[p]]) + ins
                c["type"] = 1
                n_sup = int(rng.binomial(int(depth_h[i]), float(vaf_h[i])))
                f = int(rng.binomial(n_sup, 0.5))
                c["support"] = (f, n_sup - f, 0)
                wf = int(rng.binomial(f, 0.9))
                wr = int(rng.binomial(n_sup - f, 0.9))
                c["well_anchored"] = (wf, wr, 0)
            else:
                ref_s, alt_s = bytes(refb[p:p + L + 1].tolist()), bytes([refb[p]])
                c["type"] = 2
                e0, e1 = int(off_h[p + 1]), int(off_h[p + 2])
                cc, qq, aa = code[e0:e1], qual[e0:e1], anchor[e0:e1]
                usable = ((cc & 7) == 5) & (qq >= min_bq)
                d = (cc >> 3) & 3
                sup = [int((usable & (d == j)).sum()) for j in range(3)]
                wa = [int((usable & (d == j) & ((aa & 15) == 5)).sum()) for j in range(3)]
                c["support"], c["well_anchored"] = tuple(sup), tuple(wa)
            c["ref_len"], c["alt_len"] = len(ref_s), len(alt_s)
            arena += ref_s + alt_s
        out["candidates"], out["arena"] = cands, bytes(arena)
        out["indel_loci"] = k
    return out


def algorithmic_bytes(n_loci, n_entries, n_records, third_byte=True):
    """B(D,E) summed over loci: 3 B per entry + 8 B per locus + 96 B per emitted record; 2 B per entry when the
    anchor/collapsed byte is not needed by the configuration (the hot kernel then never reads that plane)."""
    return (3 if third_byte else 2) * n_entries + 8 * n_loci + 96 * n_records


def make_reads(n_loci, mean_depth, seed=2, device="cpu", read_len=READ_LEN, snv_rate=0.01, vaf=(0.01, 0.5), indel_rate=0.001, mnv_pair_rate=0.0,
               strand_skew_frac=0.0, collapsed_frac=0.0, stitched_frac=0.0, chunk_reads=1 << 17):
    """Seeded synthetic READS of the BASELINE.json shapes: n_loci * mean_depth / read_len reads of length read_len whose start is uniform over
    the chromosome (so the depth is ~Poisson(mean_depth) away from the two ends), sorted by position like a BAM, as the struct of arrays pb2_push_reads
    takes. The chromosome is n_loci bases, uniform over ACGT. Base quality from the example_S1-like histogram of make_pileup; a base is mis-called with
    probability 10^(-q/10); SNV loci (snv_rate) carry an alternate base at VAF ~ U[vaf]; 1-3 bp insertions / deletions at indel_rate of the loci
    (>= 12 loci apart), carried at VAF ~ U[0.03, 0.43] by the reads that cover the site with at least 10 aligned bases on either side (CIGAR aM kI bM / aM kD bM).
    mnv_pair_rate: adjacent SNV pairs carried by the same reads (CallMNVs workloads); strand_skew_frac: fraction of the SNV loci whose alternate allele sits
    9:1 on the forward strand; collapsed_frac: reads tagged as collapsed (duplex 20 %, simplex FR / RF 40 % each; `collapsed` summary bytes);
    stitched_frac: reads (without indels) whose bases run Forward / Stitched / Reverse (`base_dirs` + `xd_runs`, the XD tag's three run lengths).
    Returns a dict of CPU numpy arrays (pos0, flag, cigar_off, cigar, seq_off, bases, quals[, collapsed, base_dirs, xd_runs]) + ref (ASCII bytes) + sizes."""
    import numpy as np
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    L = int(read_len)
    n_reads = max(1, int(round(n_loci * float(mean_depth) / L)))
    hi_start = max(1, n_loci - L - 4 + 1)                         # deletions reach up to 3 bases further
    starts = torch.sort(torch.randint(1, hi_start + 1, (n_reads,), device=dev, generator=g, dtype=torch.int64)).values
    ref_idx = torch.randint(0, 4, (n_loci + 8,), device=dev, generator=g)
    is_snv = torch.rand(n_loci + 8, device=dev, generator=g) < snv_rate
    alt_idx = (ref_idx + torch.randint(1, 4, (n_loci + 8,), device=dev, generator=g)) % 4
    locus_vaf = vaf[0] + (vaf[1] - vaf[0]) * torch.rand(n_loci + 8, device=dev, generator=g)
    linked = torch.zeros(n_loci + 8, dtype=torch.bool, device=dev)
    if mnv_pair_rate > 0:
        first = torch.nonzero(torch.rand(n_loci, device=dev, generator=g) < mnv_pair_rate).flatten()
        first = first[first < n_loci - 2]
        is_snv[first] = True
        is_snv[first + 1] = True
        locus_vaf[first + 1] = locus_vaf[first]
        linked[first] = True
        linked[first + 1] = True
    locus_vaf = torch.where(is_snv, locus_vaf, torch.zeros_like(locus_vaf)).to(torch.float32)
    skew = (torch.rand(n_loci + 8, device=dev, generator=g) < strand_skew_frac) & is_snv
    # indel sites: p = the base before the event (1-based position p <-> index p - 1)
    n_sites = 0
    site_pos = torch.zeros(0, dtype=torch.int64, device=dev)
    if indel_rate > 0:
        cand = torch.nonzero(torch.rand(n_loci, device=dev, generator=g) < indel_rate).flatten() + 1
        cand = cand[(cand >= 2) & (cand < n_loci - 12)]
        if cand.numel() > 1:
            keep = torch.ones_like(cand, dtype=torch.bool)
            keep[1:] = (cand[1:] - cand[:-1]) >= 12
            cand = cand[keep]
        site_pos = cand
        n_sites = int(cand.numel())
    site_len = torch.randint(1, 4, (max(n_sites, 1),), device=dev, generator=g)
    site_ins = torch.rand(max(n_sites, 1), device=dev, generator=g) < 0.5
    site_vaf = 0.03 + 0.4 * torch.rand(max(n_sites, 1), device=dev, generator=g)
    site_bases = torch.randint(0, 4, (max(n_sites, 1), 3), device=dev, generator=g)
    qcdf = _quality_table().to(dev)
    ascii_acgt = _ASCII_ACGT.to(dev)
    ar = torch.arange(L, device=dev, dtype=torch.int64)[None, :]

    out_bases, out_quals, out_dirs = [], [], []
    flags = torch.empty(n_reads, dtype=torch.int32, device=dev)
    n_ops = torch.ones(n_reads, dtype=torch.int64, device=dev)
    ev_a = torch.zeros(n_reads, dtype=torch.int64, device=dev)      # aligned bases before the event
    ev_k = torch.zeros(n_reads, dtype=torch.int64, device=dev)      # event length, > 0 insertion, < 0 deletion, 0 none
    collapsed = torch.zeros(n_reads, dtype=torch.uint8, device=dev) if collapsed_frac > 0 else None
    xd_runs = torch.zeros((n_reads, 3), dtype=torch.int32, device=dev) if stitched_frac > 0 else None
    for r0 in range(0, n_reads, chunk_reads):
        r1 = min(n_reads, r0 + chunk_reads)
        n = r1 - r0
        s = starts[r0:r1]
        reverse = torch.rand(n, device=dev, generator=g) < 0.5
        first_mate = torch.rand(n, device=dev, generator=g) < 0.5
        flags[r0:r1] = (0x1 | 0x2 | torch.where(reverse, 0x10, 0x20) | torch.where(first_mate, 0x40, 0x80)).to(torch.int32)
        k = torch.zeros(n, dtype=torch.int64, device=dev)
        a = torch.zeros(n, dtype=torch.int64, device=dev)
        ins_b = torch.zeros((n, 3), dtype=torch.int64, device=dev)
        if n_sites:
            idx = torch.searchsorted(site_pos, s + 10).clamp_(max=n_sites - 1)
            p = site_pos[idx]
            kk = site_len[idx]
            ins = site_ins[idx]
            room = torch.where(ins, p + 1 + 10, p + kk + 1 + 10)       # >= 10 aligned bases after the event
            ok = (p >= s + 10) & (room <= s + L - 1 - torch.where(ins, kk, torch.zeros_like(kk)))
            carry = ok & (torch.rand(n, device=dev, generator=g) < site_vaf[idx])
            k = torch.where(carry, torch.where(ins, kk, -kk), k)
            a = torch.where(carry, p - s + 1, a)
            ins_b = site_bases[idx]
        ev_a[r0:r1] = a
        ev_k[r0:r1] = k
        n_ops[r0:r1] = torch.where(k != 0, 3, 1)
        kpos, kneg = k.clamp(min=0)[:, None], (-k).clamp(min=0)[:, None]
        a2 = a[:, None]
        inserted = (k[:, None] > 0) & (ar >= a2) & (ar < a2 + kpos)
        refoff = torch.where((k[:, None] > 0) & (ar >= a2 + kpos), ar - kpos, torch.where((k[:, None] < 0) & (ar >= a2), ar + kneg, ar.expand(n, L)))
        refpos = (s[:, None] + refoff).clamp_(max=n_loci + 7)          # 1-based; meaningless where inserted
        ri = refpos - 1
        u = torch.rand((n, L), device=dev, generator=g)
        q = torch.searchsorted(qcdf, u.to(torch.float64)).clamp_(0, 63)
        u_base = torch.rand((n, L), device=dev, generator=g)
        u_read = torch.rand((n, 1), device=dev, generator=g).expand(n, L)
        v = locus_vaf[ri]
        if strand_skew_frac > 0:
            v = torch.where(skew[ri], v * torch.where(reverse, 0.2, 1.8)[:, None], v)
        is_alt = torch.where(linked[ri], u_read, u_base) < v
        base = torch.where(is_alt, alt_idx[ri], ref_idx[ri])
        perr = torch.pow(10.0, -q.to(torch.float32) / 10.0)
        is_err = torch.rand((n, L), device=dev, generator=g) < perr
        base = torch.where(is_err, (base + torch.randint(1, 4, (n, L), device=dev, generator=g)) % 4, base)
        if n_sites:
            j = (ar - a2).clamp_(0, 2).expand(n, L)
            base = torch.where(inserted, torch.gather(ins_b, 1, j), base)
        out_bases.append(ascii_acgt[base].reshape(-1).cpu())
        out_quals.append(q.to(torch.uint8).reshape(-1).cpu())
        if collapsed is not None:
            is_c = torch.rand(n, device=dev, generator=g) < collapsed_frac
            duplex = torch.rand(n, device=dev, generator=g) < 0.2
            fr = torch.rand(n, device=dev, generator=g) < 0.5
            collapsed[r0:r1] = torch.where(is_c, 1 | torch.where(duplex, 2, 0) | torch.where(fr, 1 << 2, 2 << 2), 0).to(torch.uint8)
        if xd_runs is not None:
            st = (torch.rand(n, device=dev, generator=g) < stitched_frac) & (k == 0)
            fa = torch.randint(10, L // 2, (n,), device=dev, generator=g)
            fb = torch.randint(10, L // 3, (n,), device=dev, generator=g)
            runs = torch.stack([fa, fb, L - fa - fb], 1).to(torch.int32)
            xd_runs[r0:r1] = torch.where(st[:, None], runs, torch.zeros_like(runs))
            d_plain = torch.where(reverse, 1, 0)[:, None].expand(n, L)
            d_st = torch.where(ar < fa[:, None], 0, torch.where(ar < (fa + fb)[:, None], 2, 1))
            out_dirs.append(torch.where(st[:, None], d_st, d_plain).to(torch.uint8).reshape(-1).cpu())
    # CIGARs: L M | a M, k I, (L - a - k) M | a M, k D, (L - a) M
    cigar_off = torch.zeros(n_reads + 1, dtype=torch.int64, device=dev)
    torch.cumsum(n_ops, 0, out=cigar_off[1:])
    cigar = torch.empty(int(cigar_off[-1]), dtype=torch.int64, device=dev)
    plain = ev_k == 0
    cigar[cigar_off[:-1][plain]] = (L << 4) | 0
    ev = ~plain
    o = cigar_off[:-1][ev]
    ka, kk = ev_a[ev], ev_k[ev]
    cigar[o] = (ka << 4) | 0
    cigar[o + 1] = torch.where(kk > 0, (kk << 4) | 1, ((-kk) << 4) | 2)
    cigar[o + 2] = (torch.where(kk > 0, L - ka - kk, L - ka) << 4) | 0
    seq_off = torch.arange(n_reads + 1, dtype=torch.int64) * L
    ref = ascii_acgt[ref_idx[:n_loci]].cpu().numpy()
    n_entries = int(n_reads) * L - int(ev_k.clamp(min=0).sum()) + int((-ev_k).clamp(min=0).sum())
    d = dict(n_reads=n_reads, read_len=L, n_loci=n_loci, ref=ref, pos0=(starts - 1).to(torch.int32).cpu().numpy(), flag=flags.to(torch.int16).cpu().numpy().view(np.uint16),
             cigar_off=cigar_off.cpu().numpy(), cigar=cigar.to(torch.int32).cpu().numpy().view(np.uint32), seq_off=seq_off.numpy(),
             bases=torch.cat(out_bases).numpy(), quals=torch.cat(out_quals).numpy(), n_entries=n_entries, snv_loci=int(is_snv[:n_loci].sum()), indel_loci=n_sites,
             collapsed=None if collapsed is None else collapsed.cpu().numpy(), xd_runs=None if xd_runs is None else xd_runs.cpu().numpy(),
             base_dirs=torch.cat(out_dirs).numpy() if out_dirs else None)
    return d


def reads_slice(d, lo, hi):
    """The reads of make_reads' dict whose start (0-based) lies in [lo, hi), offsets rebased to the slice."""
    import numpy as np
    r0, r1 = int(np.searchsorted(d["pos0"], lo, "left")), int(np.searchsorted(d["pos0"], hi, "left"))
    out = dict(d)
    c0, c1, s0, s1 = int(d["cigar_off"][r0]), int(d["cigar_off"][r1]), int(d["seq_off"][r0]), int(d["seq_off"][r1])
    out.update(n_reads=r1 - r0, pos0=d["pos0"][r0:r1], flag=d["flag"][r0:r1], cigar_off=d["cigar_off"][r0:r1 + 1] - c0, cigar=d["cigar"][c0:c1],
               seq_off=d["seq_off"][r0:r1 + 1] - s0, bases=d["bases"][s0:s1], quals=d["quals"][s0:s1])
    for k in ("collapsed", "xd_runs"):
        if d.get(k) is not None:
            out[k] = d[k][r0:r1]
    if d.get("base_dirs") is not None:
        out["base_dirs"] = d["base_dirs"][s0:s1]
    return out


def reads_concat(parts, chunk_loci):
    """Concatenates read sets [(dict, position shift)] (in position order) into one; the reference is the shifted concatenation of the parts' chromosomes
    (a part shifted by a negative amount contributes its tail)."""
    import numpy as np
    ref_parts, pos, flag, cig, bases, quals, coll, dirs, xd = [], [], [], [], [], [], [], [], []
    coff, soff = [np.zeros(1, dtype=np.int64)], [np.zeros(1, dtype=np.int64)]
    n_entries = 0
    total_len = 0
    for dct, shift in parts:
        full = dct["ref"]
        ref_parts.append(full[-shift:] if shift < 0 else full)
        total_len = max(total_len, shift + len(full))
        pos.append(dct["pos0"] + shift); flag.append(dct["flag"]); cig.append(dct["cigar"]); bases.append(dct["bases"]); quals.append(dct["quals"])
        coff.append(dct["cigar_off"][1:] + coff[-1][-1]); soff.append(dct["seq_off"][1:] + soff[-1][-1])
        if dct.get("collapsed") is not None:
            coll.append(dct["collapsed"])
        if dct.get("base_dirs") is not None:
            dirs.append(dct["base_dirs"])
        if dct.get("xd_runs") is not None:
            xd.append(dct["xd_runs"])
        n_entries += int(len(dct["bases"]))
    ref = np.concatenate(ref_parts)
    first = parts[0][0]
    return dict(n_reads=int(sum(len(p) for p in pos)), read_len=first["read_len"], n_loci=len(ref), ref=ref, pos0=np.concatenate(pos).astype(np.int32), flag=np.concatenate(flag),
                cigar_off=np.concatenate(coff), cigar=np.concatenate(cig), seq_off=np.concatenate(soff), bases=np.concatenate(bases), quals=np.concatenate(quals),
                n_entries=n_entries, collapsed=np.concatenate(coll) if coll else None, base_dirs=np.concatenate(dirs) if dirs else None,
                xd_runs=np.concatenate(xd) if xd else None, snv_loci=first.get("snv_loci"), indel_loci=first.get("indel_loci"))
