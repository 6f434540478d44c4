"""Parity of the germline genotypers on the CUDA path (DiploidThresholdingGenotyper, HaploidGenotyper, DiploidLocusProcessor,
Diploid strand-bias model) against the CPU oracle. Needs a B200: -m gpu."""
import numpy as np
import pytest

from oracle import binding as ob
from tests import util_reads as U
from tests.test_gpu_explicit import _compare_chunks, _reads_both, compare_records
from tests.util_counts import pileup_from_counts

pytestmark = pytest.mark.gpu
DIPLOID, HAPLOID = 1, 3


def _pb():
    import pisces_b200 as pb
    return pb


def _scenario_counts(rng, n_loci, ref):
    """Loci with 0-3 SNV alleles at fractions around the 0.20 / 0.70 / 0.80 thresholds, some of them shallow (depth issue) or strand-skewed."""
    counts = np.zeros((n_loci, 6, 3, 11), dtype=np.int32)
    idx = {"A": 0, "G": 1, "C": 2, "T": 3}
    for i in range(n_loci):
        depth = int(rng.choice([6, 40, 100, 250]))
        others = [a for a in range(4) if a != idx[ref[i]]]
        k = int(rng.integers(0, 4))
        fr = []
        for _ in range(k):
            fr.append(float(rng.choice([0.02, 0.1, 0.19, 0.2, 0.21, 0.3, 0.4, 0.5, 0.69, 0.7, 0.71, 0.9, 1.0])))
        tot = sum(fr)
        if tot > 1.0:
            fr = [f / tot for f in fr]
        left = depth
        for a, f in zip(others, fr):
            n = min(left, int(round(f * depth)))
            skew = float(rng.choice([0.5, 0.5, 0.5, 0.95]))
            nf = int(round(n * skew))
            counts[i, a, 0, 5] += nf
            counts[i, a, 1, 5] += n - nf
            left -= n
        nn = min(left, int(rng.integers(0, 3)))
        counts[i, 4, 0, 5] += nn                      # a few no-calls
        left -= nn
        counts[i, idx[ref[i]], 0, 5] += left // 2
        counts[i, idx[ref[i]], 1, 5] += left - left // 2
    return counts


@pytest.mark.parametrize("gvcf", [0, 1])
@pytest.mark.parametrize("ploidy,sb_model", [(DIPLOID, 2), (DIPLOID, 1), (HAPLOID, 2)])
def test_germline_snv_scenarios_match_oracle(ploidy, sb_model, gvcf):
    """Per-locus joint genotypes over count-based SNV alleles: 0/0, 0/1, 1/1, 1/2, the no-call flavours, multi-allelic sites (filter + pruning), depth
    issues; GQ from the MathNet Poisson / Binomial log-probabilities; min frequency = MinorVF; the Diploid strand-bias model."""
    pb = _pb()
    rng = np.random.default_rng(100 + ploidy * 10 + sb_model + gvcf)
    n = 400
    ref = U.random_reference(rng, n)
    counts = _scenario_counts(rng, n, ref)
    off, code, qual, anch = pileup_from_counts(counts, n)
    kw = dict(output_gvcf=gvcf, collapse=0, ploidy=ploidy, min_coverage=10, low_genotype_quality_filter=30)
    oc = ob.Caller(ob.default_config(sb_model=sb_model, min_vq=20, low_gq_filter=30, **{k: v for k, v in kw.items() if k != "low_genotype_quality_filter"}), "chr1", ref)
    oc.add_pileup(off, code, qual, anch, 1)
    oc.finish()
    sm = pb.GpuStateManager(pb.make_config(strand_bias_model=sb_model, min_variant_qscore=20, **kw), "chr1", ref)
    sm.AddPileup(off, code, qual, anch, first_position=1)
    precs = pb.GpuAlleleCaller().Call(sm, raw=True)
    arena = sm.AlleleArena()
    sm.close()
    orecs = oc.records()
    gts = {ob.GENOTYPES[o.genotype] for o in orecs}
    if ploidy == DIPLOID:
        assert {"HeterozygousAltRef", "HomozygousAlt", "HeterozygousAlt1Alt2", "AltLikeNoCall"} <= gts
        assert any(o.filter_mask >> 8 & 1 for o in orecs)          # MultiAllelicSite
    else:
        assert {"HemizygousAlt"} <= gts
    compare_records(orecs, precs, arena)


@pytest.mark.parametrize("ploidy,chr_name,is_male", [(DIPLOID, "chr1", -1), (HAPLOID, "chr1", -1), (DIPLOID, "chrX", 1), (DIPLOID, "chrM", -1)])
def test_germline_reads_with_indels_and_forced_alleles(ploidy, chr_name, is_male):
    """Streamed reads: SNVs from the counts next to explicit indel candidates at the same loci, forced alleles (DiploidLocusProcessor), reference
    records (gVCF), and the per-chromosome ploidy of GenotypeCreator.GetPloidyForThisChr (chrX of a male sample -> haploid, chrM -> somatic)."""
    pb = _pb()
    rng = np.random.default_rng(200 + ploidy + len(chr_name) + is_male)
    ref = U.random_reference(rng, 2600)
    hot = {int(p): ("ACGT"[int(rng.integers(0, 4))], float(rng.choice([0.1, 0.25, 0.45, 0.75, 0.95, 1.0]))) for p in rng.integers(60, 2400, 60)}
    reads = U.make_reads(rng, ref, 6000, read_len=60, hotspots=hot, del_rate=0.02, ins_rate=0.02, clip_rate=0.1, indel_sites=40)
    forced = []
    for p, (b, _) in list(hot.items())[:15]:
        if ref[p - 1] != b:
            forced.append((p, ref[p - 1], b))
    for p in rng.integers(60, 2400, 15):
        p = int(p)
        forced.append((p, ref[p - 1], "ACGT"[("ACGT".index(ref[p - 1]) + 1) % 4]))
    forced = sorted(set(forced))
    kw = dict(output_gvcf=1, collapse=0, ploidy=ploidy, is_male=is_male)
    okw = dict(kw, min_vq=20, sb_model=2)
    pkw = dict(kw, min_variant_qscore=20, strand_bias_model=2)
    oc = ob.Caller(ob.default_config(**okw), chr_name, ref)
    for f in forced:
        oc.add_forced(*f)
    for rd in reads:
        oc.add_read(U.to_oracle(rd))
    oc.finish()
    sm = pb.GpuStateManager(pb.make_config(**pkw), chr_name, ref)
    sm.SetForcedAlleles(forced)
    caller = pb.GpuAlleleCaller()
    chunks = []
    for i in range(0, len(reads), 800):
        chunk = reads[i:i + 800]
        sm.AddAlleleCounts([U.to_product(rd) for rd in chunk])
        r = caller.Call(sm, upToPosition=chunk[-1]["pos"] - 1, raw=True)
        chunks.append((r, sm.AlleleArena()))
    r = caller.Call(sm, raw=True)
    chunks.append((r, sm.AlleleArena()))
    sm.close()
    orecs = oc.records()
    assert sum(1 for o in orecs if o.forced) > 5
    if ploidy == DIPLOID and chr_name != "chrX":
        assert sum(1 for o in orecs if o.type in (ob.INSERTION, ob.DELETION)) > 5
    _compare_chunks(orecs, chunks)
