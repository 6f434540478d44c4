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
                collapsed_frac=0.0, strand_skew_frac=0.0, chunk_entries=1 << 25, flags=True):
    """Returns dict(offsets int64[n+1], code/qual/anchor uint8[total], ref_bases uint8[n] ASCII, snv_loci, del_loci)."""
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
    return dict(offsets=offsets, code=code, qual=qual, anchor=anchor, ref_bases=ref_bases, n_entries=total,
                snv_loci=int(is_snv.sum()), del_loci=int(is_del.sum()))


def algorithmic_bytes(n_loci, n_entries, n_records, third_byte=True):
    """B(D,E) summed over loci: 3 B per entry + 8 B per locus + 96 B per emitted record; 2 B per entry when the
    anchor/collapsed byte is not needed by the configuration (the hot kernel then never reads that plane)."""
    return (3 if third_byte else 2) * n_entries + 8 * n_loci + 96 * n_records
