"""BASELINE.json configs[1] at its full size (1 M loci x depth ~Poisson(500)) through the C ABI, checked with size-independent means:
every locus' coverage / no-call / reference-support integers against an independent torch count of the same entries (a checksum per locus, all
million of them), idempotence of the resident step, and full-record parity against the CPU oracle on three slices. Needs a B200: -m gpu."""
import numpy as np
import pytest

from oracle import binding as ob

pytestmark = pytest.mark.gpu
N_LOCI, DEPTH, MIN_BQ = 1_000_000, 500, 20
AT_N, AT_DEL = 4, 5   # AlleleType codes of the entry byte: A 0, G 1, C 2, T 3, N 4, Deletion 5 (include/pisces_b200.h)


def _per_locus(mask, off):
    import torch
    c = torch.zeros(mask.numel() + 1, dtype=torch.int64, device=mask.device)
    torch.cumsum(mask.to(torch.int64), 0, out=c[1:])
    return (c[off[1:]] - c[off[:-1]]).cpu().numpy()


@pytest.mark.parametrize("gvcf", [1, 0])
def test_full_size_pileup(gvcf):
    import torch
    import pisces_b200 as pb
    from pisces_b200 import synth
    from tests.test_gpu_parity import _compare_records

    d = synth.make_pileup(N_LOCI, DEPTH, seed=2, device="cuda:0", depth_dist="poisson", indel_rate=0.0)
    off, code, qual = d["offsets"], d["code"], d["qual"]
    ref_np = d["ref_bases"].cpu().numpy()
    ref = bytes(ref_np).decode()
    sm = pb.GpuStateManager(pb.make_config(device=0, output_gvcf=gvcf), "chr1", ref)
    sm.AddPileup(off, code, qual, d["anchor"], first_position=1, ref_bases=d["ref_bases"], device=True)
    n1 = sm.call_resident()
    n2 = sm.call_resident()
    assert n1 == n2 > 0                                    # idempotent: the resident step leaves the staged pileup untouched
    recs = pb.GpuAlleleCaller().Call(sm, raw=True)
    sm.close()
    pos = recs["position"].astype(np.int64)
    assert np.all(np.diff(pos) >= 0)                       # ordered by position (AlleleCaller.cs:172-176)
    if gvcf:
        assert len(np.unique(pos)) == N_LOCI               # gVCF: a record at every covered position

    # ---- independent count of the same entries (RegionStateManager.cs:176-188: a base below the quality bar counts as N)
    allele = code & 7
    usable = qual >= MIN_BQ
    covered = usable & (allele != AT_N)
    exp_cov = _per_locus(covered, off)
    exp_nc = _per_locus(~covered & (allele != AT_DEL), off)   # a deletion below the quality bar is dropped, not a no-call (:168-176)
    lut = torch.full((256,), AT_N, dtype=torch.uint8, device=code.device)
    for ch, a in (("A", 0), ("G", 1), ("C", 2), ("T", 3)):
        lut[ord(ch)] = a
    ref_allele = lut[d["ref_bases"].long()]
    depth = off[1:] - off[:-1]
    exp_ref = _per_locus(covered & (allele == torch.repeat_interleave(ref_allele, depth)), off)
    idx = pos - 1
    assert np.array_equal(recs["total_coverage"], exp_cov[idx])
    assert np.array_equal(recs["num_no_calls"], exp_nc[idx])
    assert np.array_equal(recs["coverage_by_direction"].sum(axis=1), recs["total_coverage"])
    is_ref = recs["type"] == 4
    assert np.array_equal(recs["allele_support"][is_ref], exp_ref[idx[is_ref]])
    assert np.array_equal(recs["reference_support"][~is_ref], exp_ref[idx[~is_ref]])
    assert int(recs["total_coverage"][np.unique(pos, return_index=True)[1]].sum()) == int(exp_cov[np.unique(idx)].sum())   # the checksum of checksums
    assert (~is_ref).sum() > 5000                          # the planted SNVs are called

    # ---- full records against the oracle on three slices (first, middle, last 1500 loci); point alleles only, so slices are independent
    off_h = off.cpu().numpy()
    for s in (0, N_LOCI // 2, N_LOCI - 1500):
        e0, e1 = int(off_h[s]), int(off_h[s + 1500])
        oc = ob.Caller(ob.default_config(output_gvcf=gvcf), "chr1", ref)
        oc.add_pileup(off_h[s:s + 1501] - e0, code[e0:e1].cpu().numpy(), qual[e0:e1].cpu().numpy(), d["anchor"][e0:e1].cpu().numpy(), s + 1)
        oc.finish()
        sel = (pos > s) & (pos <= s + 1500)
        _compare_records(oc.records(), recs[sel])
