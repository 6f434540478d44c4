#!/usr/bin/env python
"""bench.py — candidate loci scored per second on synthetic READ sets of the BASELINE.json shapes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4|c5] [--loci L] [--depth D]

The input is reads (struct of arrays, as IStateManager.AddAlleleCounts(Read) receives them): the pileup is built on the device.
  value  a step = one pass of the hot path (pileup count + score + explicit candidates + record compaction) over the `loci` of this GPU, the staged
         pileup resident in HBM (pb2_stage_reads once, then pb2_call_resident per step);
  e2e    the same through the C ABI from HOST buffers: pb2_push_reads (pinned H2D of the reads), device staging, pb2_flush (records back on the host).
Default workload = BASELINE.json configs[1] (c2): 1 M loci x depth ~Poisson(500) (3.57 M reads of 140 bases), SNV (1 % of loci) + 1-3 bp insertions /
deletions (0.1 % of loci), flat Poisson noise model NL 20, gVCF off, one B200. N>1 (torchrun): loci are sharded by interval across ranks (weak scaling:
`loci` per GPU), no data-path collective; the job ends with the single all-gather of the per-rank call records (NCCL).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "candidate loci scored/sec"
UNIT = "loci/s"


# BASELINE.json configs[1..4]: generator arguments (pisces_b200.synth.make_reads), caller options, per-GPU size. configs[2..4] name
# whole-job sizes of 10 M / 50 M / 200 M loci on 1 / 4 / 8 GPUs; a bench step covers `loci` per GPU (stated in config.workload), larger jobs are more steps.
CONFIGS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case. The hg19 sequence is not available offline, so the reference bases are N:
    # the step stages the BAM, builds the pileup and scores the reference allele of every interval position (coverage, no-calls, genotype), no variants.
    "c1": dict(loci=3203, depth=250, seed=1, gen={}, cfg=dict(output_gvcf=1, min_coverage=10),
               what="tests/golden/example_S1.mapped.bam (the 482 mapped reads of testdata/example_S1.bam) x Intervals_1.picard, SNV-only, min-depth 10 (BASELINE.json configs[0])"),
    "c2": dict(loci=1_000_000, depth=500, seed=2, gen=dict(indel_rate=0.001), cfg=dict(output_gvcf=0),
               what="SNV 1% + indel 0.1%, Poisson noise model NL20, gvcf=0 (BASELINE.json configs[1] shape and size)"),
    "c3": dict(loci=1_000_000, depth=1000, seed=3, gen=dict(indel_rate=0.0, snv_rate=0.005), cfg=dict(output_gvcf=1),
               what="SNV 0.5%, gVCF mode: every locus emits a reference record (BASELINE.json configs[2] shape; 10 M loci = 10 such steps)"),
    "c4": dict(loci=500_000, depth=2000, seed=4, gen=dict(indel_rate=0.0, mnv_pair_rate=0.002, strand_skew_frac=0.1),
               cfg=dict(output_gvcf=0, call_mnvs=1, max_size_mnv=3, max_gap_mnv=1),
               what="CallMNVs (MaxSizeMNV 3, gap 1), adjacent-SNV pairs 0.2%, strand skew on 10% of variants (BASELINE.json configs[3] shape; 50 M loci / 4 GPUs = 25 such steps per GPU)"),
    "c5": dict(loci=1_000_000, depth=300, seed=5, gen=dict(indel_rate=0.001, mnv_pair_rate=0.001, collapsed_frac=1.0, stitched_frac=0.5),
               cfg=dict(output_gvcf=0, call_mnvs=1, expect_collapsed=1, expect_stitched=1),
               what="SNV/MNV/indel, collapsed reads (duplex 20%, simplex 80%), 50% stitched reads (BASELINE.json configs[4] shape; 200 M loci / 8 GPUs = 25 such steps per GPU)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json configs[1..4] shapes")
    ap.add_argument("--loci", type=int, default=0, help="loci per GPU (0: the config's)")
    ap.add_argument("--depth", type=int, default=0, help="mean depth (0: the config's)")
    ap.add_argument("--gvcf", type=int, default=-1, help="override the config's gVCF mode")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--gen", dest="gen_json", default="", help="JSON overrides of the config's read-generator arguments (tuning experiments), e.g. '{\"snv_rate\": 0}'")
    ap.add_argument("--tune-ctas", type=int, default=0, help="hot-kernel CTAs per SM (tuning experiments)")
    ap.add_argument("--tune-prefetch", type=int, default=0, help="tuning experiments (9: PTILE32 staging instead of PVERT)")
    ap.add_argument("--cpu-sample-loci", type=int, default=100_000)
    ap.add_argument("--cpu-repeats", type=int, default=5, help="cpu_baseline: best of this many runs of the sample (SURVEY 8d)")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--e2e-jobs", type=int, default=0, help="concurrent (BAM x chromosome) jobs of the end-to-end leg: one handle + one host thread each, as Pisces -t N runs them "
                    "(0: host cores / (2 x ranks), between 2 and 6 - measured at N = 1: 2 jobs 57.8, 3 61.2, 4 65.2, 6 68.3 M loci/s)")
    ap.add_argument("--e2e-input", default="packed", choices=["packed", "soa"], help="host form of the reads: one packed byte per base, or bases + qualities")
    ap.add_argument("--gather", default="all", choices=["all", "root"], help="end-of-job exchange of the ranks' call records: all_gather, or gather to rank 0 (the writer)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu_index, [], threading.Event()
        # NVML in-process, initialised here (before the timed region): spawning nvidia-smi from a process with torch loaded, or nvmlInit itself,
        # costs the main thread milliseconds - the size of the whole timed region at small K. nvidia-smi only if NVML cannot be loaded.
        self.nv = self.hd = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.hd = None
            try:   # the CUDA ordinal is not the NVML index when CUDA_VISIBLE_DEVICES re-maps the devices: go through the UUID
                import torch
                uuid = "GPU-" + str(torch.cuda.get_device_properties(self.gpu).uuid)
                try:
                    self.hd = nv.nvmlDeviceGetHandleByUUID(uuid)
                except TypeError:
                    self.hd = nv.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.hd = None
            if self.hd is None:
                self.hd = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.mx = nv.nvmlDeviceGetMaxClockInfo(self.hd, nv.NVML_CLOCK_SM)
            self.get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.nv = nv
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is not None:
            nv = self.nv
            try:
                while True:
                    m = int(self.get_reasons(self.hd))
                    act = lambda bit: "Active" if m & bit else "Not Active"   # noqa: E731
                    self.rows.append([str(self.gpu), str(nv.nvmlDeviceGetClockInfo(self.hd, nv.NVML_CLOCK_SM)), str(self.mx), "0", act(0x8), act(0x40), act(0x20),
                                      act(0x4)])
                    if self.stop_flag.wait(0.005):
                        return
            except Exception:
                pass
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(self.rows)}


def resolve(a):
    c = CONFIGS[a.config]
    a.loci = a.loci or c["loci"]
    a.depth = a.depth or c["depth"]
    a.seed = a.seed or c["seed"]
    a.gen = dict(c["gen"])
    if a.gen_json:
        a.gen.update(json.loads(a.gen_json))
    a.cfg = dict(c["cfg"])
    if a.gvcf >= 0:
        a.cfg["output_gvcf"] = a.gvcf
    return a


def workload_name(a):
    return f"{a.config}: synthetic reads (length 140) over {a.loci} loci x depth ~Poisson({a.depth}) per GPU; {CONFIGS[a.config]['what']}"


def oracle_caller(a, ref):
    from oracle import binding as ob
    kw = dict(a.cfg)
    if "expect_stitched" in kw:
        kw["source_is_stitched"] = kw.pop("expect_stitched")
    if "expect_collapsed" in kw:
        kw["source_is_collapsed"] = kw.pop("expect_collapsed")
    return ob.Caller(ob.default_config(collapse=1, **kw), "chr1", ref)


def reads_window(d, lo, hi):
    """The reads whose start lies in [lo, hi) (0-based) as views, their positions rebased to the window, + the window's reference."""
    import numpy as np
    r0, r1 = int(np.searchsorted(d["pos0"], lo, "left")), int(np.searchsorted(d["pos0"], hi, "left"))
    L = d["read_len"]
    end = min(len(d["ref"]), hi + L + 8)
    out = dict(pos0=d["pos0"][r0:r1] - lo, flag=d["flag"][r0:r1], cigar_off=d["cigar_off"][r0:r1 + 1] - d["cigar_off"][r0], seq_off=d["seq_off"][r0:r1 + 1] - d["seq_off"][r0],
               cigar=d["cigar"][d["cigar_off"][r0]:d["cigar_off"][r1]], bases=d["bases"][d["seq_off"][r0]:d["seq_off"][r1]],
               quals=d["quals"][d["seq_off"][r0]:d["seq_off"][r1]], ref=bytes(d["ref"][lo:end]).decode(),
               collapsed=None if d.get("collapsed") is None else d["collapsed"][r0:r1], xd_runs=None if d.get("xd_runs") is None else d["xd_runs"][r0:r1])
    return out


def run_oracle(a, w):
    """The CPU restatement over one window of reads: SmallVariantCaller.Execute's per-read loop (find candidates, count, call). Returns (seconds, caller)."""
    oc = oracle_caller(a, w["ref"])
    t0 = time.perf_counter()
    oc.add_reads_soa(w["pos0"], w["flag"], w["cigar_off"], w["cigar"], w["seq_off"], w["bases"], w["quals"], w.get("collapsed"), w.get("xd_runs"))
    oc.finish()
    return time.perf_counter() - t0, oc


def main_reference(a):
    """Reference arm: the reference's CPU implementation of the path. The C# cannot run here (no dotnet runtime in the image), so this is
    the oracle port (oracle/) on all host threads, each step a bounded sample of the same workload: every thread runs the reference's per-read loop
    over its own window of the read set."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if a.config == "c1":   # the reference's own CPU-runnable case: one job, one thread (the reference runs a BAM x chromosome job on one thread)
        from oracle import binding as ob
        from tests import bamio
        path, iv, n_loci = c1_inputs()
        _, refs, recs = bamio.read_bam(path)
        kept = [r for r in recs if not (r["flag"] & 0x4 or r["flag"] & 0x100 or r["flag"] & 0x400 or r["mapq"] < 1 or not r["cigar"] or r["ref_id"] < 0)]
        times = []
        for i in range(a.warmup + a.steps):
            t1 = time.perf_counter()
            for ref_id in sorted({r["ref_id"] for r in kept}):
                oc = ob.Caller(ob.default_config(output_gvcf=1, min_coverage=10), refs[ref_id][0], "", intervals=iv.get(refs[ref_id][0], []))
                for r in kept:
                    if r["ref_id"] == ref_id:
                        oc.add_read(ob.SimpleRead(r["pos0"] + 1, r["seq"], r["cigar"], r["qual"], flag=r["flag"], mapq=r["mapq"]))
                oc.finish()
            if i >= a.warmup:
                times.append(time.perf_counter() - t1)
        value = n_loci * a.steps / sum(times)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": 1e3 * sum(times) / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64",
                          "data": "reference test data (tests/golden)", "config": {"workload": f"c1: {CONFIGS['c1']['what']}", "loci_per_gpu": n_loci},
                          "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"the {len(kept)} reads through the oracle's per-read loop"},
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    import torch
    from pisces_b200 import synth
    cores = os.cpu_count() or 1
    per_thread = max(1000, min(8_000, a.loci // cores))
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    d = synth.make_reads(cores * per_thread + 200, a.depth, seed=a.seed, device=dev, **a.gen)
    wins = [reads_window(d, t * per_thread, (t + 1) * per_thread) for t in range(cores)]
    times = []
    for i in range(a.warmup + a.steps):
        ths = [threading.Thread(target=run_oracle, args=(a, w)) for w in wins]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        if i >= a.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = a.steps * cores * per_thread / total
    sample = f"{cores} threads x {per_thread} loci per step ({cores * per_thread} loci, depth {a.depth}) of the same synthetic read set"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "loci_per_gpu": a.loci},
        "note": "oracle port of the C# path (dotnet runtime absent): SmallVariantCaller.Execute's per-read loop, no I/O",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def compare_with_oracle(orecs, precs, arena, limit):
    """Records of the oracle run on a sample window against the product's records of the same positions (<= limit). Returns (compared, mismatches)."""
    mism = 0
    o = [r for r in orecs if r.pos <= limit]
    p = [r for r in precs if int(r["position"]) <= limit]
    if len(o) != len(p):
        return max(len(o), len(p)), abs(len(o) - len(p)) + 1
    for x, y in zip(o, p):
        rl, al, ab = int(y["ref_len"]), int(y["alt_len"]), int(y["allele_bytes"])
        raw = ab.to_bytes(4, "little") if rl + al <= 4 else bytes(arena[ab:ab + rl + al])
        ok = (x.pos == int(y["position"]) and x.type == int(y["type"]) and raw[:rl].decode() == x.ref and raw[rl:rl + al].decode() == x.alt and
              x.total_coverage == int(y["total_coverage"]) and x.allele_support == int(y["allele_support"]) and x.ref_support == int(y["reference_support"]) and
              x.vq == int(y["variant_qscore"]) and x.gq == int(y["genotype_qscore"]) and x.genotype == int(y["genotype"]) and x.filter_mask == int(y["filters"]) and
              list(x.cov) == list(y["coverage_by_direction"]) and list(x.support) == list(y["support_by_direction"]) and x.num_no_calls == int(y["num_no_calls"]))
        mism += 0 if ok else 1
    return len(o), mism


def c1_inputs():
    import collections
    G = os.path.join(ROOT, "tests", "golden")
    iv = collections.OrderedDict()
    for line in open(os.path.join(G, "Intervals_1.picard")):
        if line.startswith("@") or not line.strip():
            continue
        f = line.split("\t")
        iv.setdefault(f[0], []).append((int(f[1]), int(f[2])))
    n_loci = sum(len(set(p for a_, b_ in v for p in range(a_, b_ + 1))) for v in iv.values())
    return os.path.join(G, "example_S1.mapped.bam"), iv, n_loci


def main_c1(a):
    """configs[0]: BAM file -> library stager -> device pileup -> records, one handle per chromosome with its intervals. Host file in, records out: the step
    IS end to end; the CPU baseline is the oracle's per-read loop over the same reads."""
    import torch
    import pisces_b200 as pb
    from oracle import binding as ob
    from tests import bamio
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; pisces_b200 has no CPU path")
    path, iv, n_loci = c1_inputs()
    cfg = pb.make_config(**a.cfg)

    sms = {}   # one handle per chromosome, kept across jobs as a worker thread of the host would (pb2_reset between jobs; intervals stay set)

    def step():
        st = pb.BamReadStager(path, packed=True)   # pb2_bam_next_batch_packed: the stager's batches in the one-byte-per-base form
        names = [n for n, _ in st.references]
        n_rec, used = 0, []
        for ref_id, batch, _ in st:
            if ref_id not in sms:
                sms[ref_id] = pb.GpuStateManager(cfg, names[ref_id], None, intervals=iv.get(names[ref_id], []))
            if ref_id not in used:
                used.append(ref_id)
                sms[ref_id].DoneProcessing()
            sms[ref_id].AddReadBatch(batch)
        for ref_id in used:
            n_rec += len(pb.GpuAlleleCaller().Call(sms[ref_id], raw=True, copy=False))
        st.close()
        return n_rec
    for _ in range(max(3, a.warmup)):
        n_rec = step()
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        n_rec = step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    for sm in sms.values():
        sm.close()
    value = n_loci * a.steps / dt
    # CPU: the oracle over the same reads (counts + reference calls), single thread
    _, refs, recs = bamio.read_bam(path)
    kept = [r for r in recs if not (r["flag"] & 0x4 or r["flag"] & 0x100 or r["flag"] & 0x400 or r["mapq"] < 1 or not r["cigar"] or r["ref_id"] < 0)]
    best = None
    for _ in range(max(1, a.cpu_repeats)):
        t1 = time.perf_counter()
        for ref_id in sorted({r["ref_id"] for r in kept}):
            oc = ob.Caller(ob.default_config(output_gvcf=1, min_coverage=10), refs[ref_id][0], "", intervals=iv.get(refs[ref_id][0], []))
            for r in kept:
                if r["ref_id"] == ref_id:
                    oc.add_read(ob.SimpleRead(r["pos0"] + 1, r["seq"], r["cigar"], r["qual"], flag=r["flag"], mapq=r["mapq"]))
            oc.finish()
        cdt = time.perf_counter() - t1
        best = cdt if best is None else min(best, cdt)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": max(3, a.warmup), "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "reference test data (tests/golden)",
            "config": {"workload": f"c1: {CONFIGS['c1']['what']}", "loci_per_gpu": n_loci},
            "workload_details": {"reads": len(kept), "records_per_step": n_rec,
                                 "note": "a 482-read job is launch- and host-bound by construction: the number is the latency of one whole job, not a roofline point"},
            "gpu_launches": None, "roofline": None, "clocks": sampler.summary(),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": sum(len(r["seq"]) * 2 + 26 for r in kept), "d2h_bytes_per_step": 96 * n_rec,
                    "input": "BAM file on the host (pb2_bam_* stager), records back on the host"},
            "cpu_baseline": {"value": n_loci / best, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": f"the same {len(kept)} reads through the oracle's per-read loop (python per-read calls included), best of {max(1, a.cpu_repeats)}"}}
    print(json.dumps(line))


def bind_to_gpu_numa_node(gpu):
    """Pins this process (and the job threads it starts) to the CPU cores NVML reports as local to the GPU, as a host that runs one worker per GPU would:
    the pinned read buffers are then allocated on that socket and the host-to-device copies do not cross the inter-socket link. Returns the core count."""
    try:
        import pynvml as nv
        import torch
        nv.nvmlInit()
        uuid = "GPU-" + str(torch.cuda.get_device_properties(gpu).uuid)
        try:
            hd = nv.nvmlDeviceGetHandleByUUID(uuid)
        except TypeError:
            hd = nv.nvmlDeviceGetHandleByUUID(uuid.encode())
        words = nv.nvmlDeviceGetCpuAffinity(hd, (os.cpu_count() + 63) // 64)
        cores = [64 * w + b for w, x in enumerate(words) for b in range(64) if (int(x) >> b) & 1]
        cores = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if cores:
            os.sched_setaffinity(0, cores)
        return len(cores)
    except Exception:
        return 0


def main_ours(a):
    if a.config == "c1":
        return main_c1(a)
    import numpy as np
    import torch
    import torch.distributed as dist
    import pisces_b200 as pb
    from pisces_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; pisces_b200 has no CPU path")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)   # before any pinned allocation: first touch puts the job's host buffers next to this rank's GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = f"cuda:{local}"

    # ---- synthetic read set of this rank, generated on the device, kept on the host in pinned memory. N > 1: the job's chromosome is `world` stretches of
    # `loci` positions, cut by pb2_shard_plan into one interval shard per rank; a rank holds its own stretch plus the reads of its neighbours' stretches
    # inside the plan's halo (positions are kept local to the rank: the shard's own range starts after the left halo).
    from pisces_b200 import sharding
    d = synth.make_reads(a.loci, a.depth, seed=a.seed + 1000 * rank, device=dev, **a.gen)
    own_lo, own_hi = 1, a.loci
    if world > 1:
        span = d["read_len"] + 4
        plan = sharding.shard_plan(None, 1, world * a.loci, span, world)
        me = plan[rank]
        assert (me["own_lo"], me["own_hi"]) == (rank * a.loci + 1, (rank + 1) * a.loci), "loci per GPU must be a multiple of 1000"
        halo_l = me["own_lo"] - me["stage_lo"]
        halo_r = min(me["stage_hi"], world * a.loci) - me["own_hi"]
        parts = []
        if rank > 0:
            nb = synth.make_reads(a.loci, a.depth, seed=a.seed + 1000 * (rank - 1), device=dev, **a.gen)
            parts.append((synth.reads_slice(nb, a.loci - halo_l, a.loci), -(a.loci - halo_l)))   # (the reads inside the halo; its first read span is thinner, 3000 positions from anything owned)
        parts.append((d, halo_l))
        if rank < world - 1:
            nb = synth.make_reads(a.loci, a.depth, seed=a.seed + 1000 * (rank + 1), device=dev, **a.gen)
            parts.append((synth.reads_slice(nb, 0, halo_r), halo_l + a.loci))
        d = synth.reads_concat(parts, a.loci)
        own_lo, own_hi = halo_l + 1, halo_l + a.loci
    torch.cuda.empty_cache()
    n_entries = d["n_entries"]
    ref = bytes(d["ref"]).decode()
    keys = ["pos0", "flag", "cigar_off", "cigar", "seq_off", "bases", "quals"] + [k for k in ("base_dirs", "collapsed") if d.get(k) is not None]
    view = {"flag": np.int16, "cigar": np.int32}
    pinned = {k: torch.from_numpy(d[k].view(view[k]) if k in view else d[k]).pin_memory() for k in keys}
    # the e2e input: the same reads packed to one byte per base (pb2_pack_reads: lossless, exceptions listed), as a host behind a PCIe link would hold them
    if a.e2e_input == "packed" and not a.no_e2e:
        # + compact offsets: one byte of CIGAR-operation count per read instead of two 8-byte offsets (built on the device)
        pk = pb.GpuStateManager.pack_reads(d, compact=bool(np.diff(d["cigar_off"]).max() <= 255))
        e2e_in = {k: (v if isinstance(v, int) else torch.from_numpy(np.ascontiguousarray(v).view(view[k]) if k in view else np.ascontiguousarray(v)).pin_memory())
                  for k, v in pk.items() if v is not None}
    else:
        e2e_in = pinned
    h2d_bytes = sum(int(t.numel() * t.element_size()) for t in e2e_in.values() if not isinstance(t, int))
    cfg = pb.make_config(device=local, **a.cfg)
    cfg.reserved[0] = a.tune_ctas
    cfg.reserved[1] = a.tune_prefetch
    sm = pb.GpuStateManager(cfg, "chr1", ref)
    if world > 1:
        sm.SetOwnedRange(own_lo, own_hi)
    sm.AddReadsSoA(pinned)
    sm.StageReads()
    # the whole job from device-resident reads (candidates found again, pileup staged again, called): reported beside the staged-pileup step
    sm.flush_resident(copy=False)
    t0 = time.perf_counter()
    for _ in range(2):
        sm.flush_resident(copy=False)
    from_reads_ms = 1e3 * (time.perf_counter() - t0) / 2
    sm.StageReads()
    torch.cuda.synchronize()
    stage = sm.stage_stats()   # of this warm staging pass (the handle's buffers exist: the first pass also pays the pool's allocations)

    # configurations whose SNV / MNV candidates come from the candidate finder (CallMNVs) need the collapser / MNV reallocator of pb2_flush: their step is
    # the whole job from the device-resident reads (pb2_flush_resident); the others run the resident staged-pileup step (pb2_call_resident)
    resident_ok = not a.cfg.get("call_mnvs")

    # ---- the job's records: every resident step appends its variant records to the rank's job buffer ON THE DEVICE (pb2_set_resident_sink: the library
    # copies them on its own stream, no host synchronisation per step); when the K steps are done the slots are ordered by position on the device
    # (pb2_sink_sort) and the ranks exchange [K counts | K fixed-capacity record blocks] in ONE all_gather, inside the timed region.
    job_buf = gather_out = None
    n_slots = max(1, a.steps)
    if resident_ok:
        n0 = sm.call_resident()          # synchronous: builds the explicit-candidate plan and the CUDA graph of the step
        n0_all = torch.tensor([n0], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(n0_all, op=dist.ReduceOp.MAX)    # the ranks' shards differ: the block capacity (hence the gathered size) must be the same on all
        cap_records = (int(n0_all.item()) * 9 // 8 + 127) // 64 * 64   # fixed capacity of a step's block: the largest shard's record count plus 12 % head-room
        job_buf = torch.zeros(8 * n_slots + n_slots * cap_records * 96, dtype=torch.uint8, device=dev)
        sm.call_resident()               # replay once synchronously (the first call built the plan on the handle stream only)
        sm.set_resident_sink(job_buf.data_ptr(), cap_records, n_slots)
        if world > 1 and (a.gather == "all" or rank == 0):
            gather_out = torch.zeros(world * job_buf.numel(), dtype=torch.uint8, device=dev)

    def step(k=0):
        if not resident_ok:
            return len(sm.flush_resident(copy=False))
        sm.call_resident_async()         # enqueue only: the step's graph + the copy of its records into slot k of the job buffer
        return 0

    # (measured, N = 8: issuing the exchange in four parts as the steps complete - sort, host sync, asynchronous all_gather per part - costs more in
    # pipeline bubbles and SM sharing than the overlap returns: 0.188 against 0.176 ms per step. One gather at the end it is.)
    def finish_job():
        if resident_ok:
            sm.sink_sort()
            n = sm.resident_sync()
            if world > 1 and a.gather == "all":
                dist.all_gather_into_tensor(gather_out, job_buf)
            elif world > 1:
                dist.gather(job_buf, list(gather_out.view(world, -1).unbind(0)) if rank == 0 else None, dst=0)
            return n
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(max(3, a.warmup)):
        n_records = step(k)
    n_fin = finish_job()   # untimed: NCCL sets its channels up on the first collective
    if n_fin is not None:
        n_records = n_fin
    sm.stats()
    sampler = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    ev0.record()
    t0 = time.perf_counter()
    for k in range(a.steps):
        n_records = step(k)
    n_fin = finish_job()
    barrier()
    ev1.record()
    ev1.synchronize()
    dt_wall = time.perf_counter() - t0
    dt = ev0.elapsed_time(ev1) * 1e-3   # device clock between the two synchronised brackets (the wall clock beside it: wall_ms_per_step)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    if n_fin is not None:
        n_records = n_fin
    timed_launches = sm.stats()["total_launches"]
    # the hot kernel's own time (the library's CUDA events around it), read after the timed region from synchronous steps in blocks of 10: shows a
    # machine-state change between blocks
    blocks = []
    if resident_ok:
        sm.set_resident_sink(None, 0, 0)
        for _ in range(3):
            for _ in range(10):
                sm.call_resident()
            blocks.append(sm.stats())
    else:
        blocks.append(dict(hot_launches=a.steps, hot_ms=0.0, total_launches=timed_launches))
        for _ in range(3):
            sm.flush_resident(copy=False)
        st_ = sm.stats()
        blocks[0]["hot_ms"] = st_["hot_ms"] / max(1, st_["hot_launches"]) * a.steps
    st = {k: sum(b[k] for b in blocks) for k in ("hot_launches", "hot_ms", "total_launches")}
    st["total_launches"] = timed_launches
    per_block = sorted(b["hot_ms"] / max(1, b["hot_launches"]) for b in blocks)
    tmax = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dt = float(tmax.item())
    value = world * a.loci * a.steps / dt

    # ---- roofline of the dominant kernel: algorithmic bytes (B = 2 D + 8 + 96 E per locus; 3 D with collapsed-read tracking) / CUDA-event duration
    n_ref_records = a.loci if a.cfg.get("output_gvcf") else 0
    third_byte = bool(a.cfg.get("expect_collapsed"))
    algo_bytes = synth.algorithmic_bytes(a.loci, n_entries, n_records + n_ref_records, third_byte=third_byte)
    hot_ms = st["hot_ms"] / max(1, st["hot_launches"])
    achieved = algo_bytes / (hot_ms * 1e-3) / 1e9
    peak, peak_src = 6650.0, "fallback"
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        peak_src = "measured"
    except Exception:
        pass
    pvert = a.tune_prefetch != 9
    kernel_name = "pileup_pvert_score_kernel" if pvert else "pileup_nib_score_kernel"
    traffic = None
    try:   # dram__bytes_read + dram__bytes_write of one launch from the committed ncu --set full capture of this command
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))[kernel_name + "_" + a.config]["dram_bytes_per_launch"]
        if a.loci != CONFIGS[a.config]["loci"] or a.depth != CONFIGS[a.config]["depth"]:
            traffic = None
    except Exception:
        pass
    staged_bytes = stage["staged_bytes"] + 96 * (n_records + n_ref_records)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup), "ms_per_step": 1e3 * dt / a.steps, "wall_ms_per_step": 1e3 * dt_wall / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            # (the same two keys as the reference arm's line, so that the two arms compare as the same configuration)
            "config": {"workload": workload_name(a), "loci_per_gpu": a.loci},
            "workload_details": {"reads_per_gpu": d["n_reads"], "entries_per_gpu": n_entries, "records_per_step": n_records + n_ref_records,
                                 "l2": "staged input of the hot kernel (%.2f GB per GPU) larger than L2, no flush needed" % (stage["staged_bytes"] / 1e9),
                                 "parallelism": f"interval shards of one chromosome x{world} (pb2_shard_plan)", "host_cores_bound_to_gpu_numa_node": numa},
            "gpu_launches": st["total_launches"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": kernel_name, "algorithmic_bytes_per_launch": algo_bytes, "bytes_per_entry": 3 if third_byte else 2,
                         # what the kernel physically reads: the staged (PVERT) form is one byte per slot, denser than the model's 2-3 bytes per entry, so `frac`
                         # can exceed the DRAM rate; frac_staged = staged bytes / kernel time / peak is the physical fraction of the HBM roofline
                         "staged_bytes_per_launch": staged_bytes, "staged_bytes_per_entry": stage["staged_bytes"] / max(1, n_entries),
                         "frac_staged": staged_bytes / (hot_ms * 1e-3) / 1e9 / peak,
                         "dram_frac": (traffic / (hot_ms * 1e-3) / 1e9 / peak) if traffic else None,
                         "kernel_ms": hot_ms, "staging_ms": stage["stage_ms"],
                         "from_resident_reads": {"ms_per_step": from_reads_ms, "value": a.loci / (from_reads_ms * 1e-3), "unit": UNIT,
                                                 "what": "find candidates + stage + call from reads resident in HBM (pb2_flush_resident), one GPU"},
                         "value_step": "pb2_call_resident (staged pileup resident)" if resident_ok else "pb2_flush_resident (reads resident; CallMNVs needs the host collapser / reallocator)",
                         "kernel_ms_blocks": {"min": per_block[0], "median": per_block[len(per_block) // 2], "max": per_block[-1]}},
            "clocks": sampler.summary()}

    # ---- end to end through the C ABI with HOST buffers: pinned H2D of the step's reads, device staging, call, D2H of the records
    precs = arena = None
    if not a.no_e2e:
        # One handle + one host thread per job, as the reference runs its (BAM x chromosome) jobs (-t N: JobManager, SURVEY 8b "Threading"): while one job's
        # reads cross the PCIe link another job's pileup is staged and called, so the link stays busy. Every job processes the same `loci` per step.
        n_jobs = a.e2e_jobs if a.e2e_jobs > 0 else max(2, min(6, (os.cpu_count() or 8) // (2 * world)))
        e2e_steps = max(2, min(a.steps, a.e2e_steps))
        sms = [pb.GpuStateManager(cfg, "chr1", ref) for _ in range(n_jobs)]
        caller = pb.GpuAlleleCaller()

        def e2e_step(sm2):
            if a.e2e_input == "packed":
                sm2.AddReadsPacked(e2e_in)
            else:
                sm2.AddReadsSoA(pinned)
            return caller.Call(sm2, raw=True, copy=False)   # the records, on the host, in the library's buffer (read in place, as a P/Invoke host would)
        for sm2 in sms:   # warm-up (allocations, block cache), and the records the oracle check below compares
            precs = e2e_step(sm2).copy()
            arena = sm2.AlleleArena()
            e2e_step(sm2)
        counts = [0] * n_jobs

        def job(j):
            for _ in range(e2e_steps):
                counts[j] = len(e2e_step(sms[j]))
        barrier()
        ths = [threading.Thread(target=job, args=(j,)) for j in range(n_jobs)]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        barrier()
        edt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(edt, op=dist.ReduceOp.MAX)
        nrec = counts[0]
        # one job alone (no overlap): the latency of a step
        t1 = time.perf_counter()
        for _ in range(2):
            e2e_step(sms[0])
        single_ms = 1e3 * (time.perf_counter() - t1) / 2
        line["e2e"] = {"value": world * n_jobs * a.loci * e2e_steps / float(edt.item()), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 96 * nrec,
                       "steps": e2e_steps, "jobs": n_jobs, "ms_per_step": 1e3 * float(edt.item()) / (e2e_steps * n_jobs), "single_job_ms_per_step": single_ms,
                       "single_job_value": world * a.loci / (single_ms * 1e-3),
                       "input": ("reads, one packed byte per base (pb2_push_reads_packed)" if a.e2e_input == "packed" else "reads, bases + qualities (pb2_push_reads)")
                                + " + %.0f B of metadata per read, pinned host memory" % ((h2d_bytes - (1 if a.e2e_input == "packed" else 2) * len(d["bases"])) / max(1, d["n_reads"]))}
        for sm2 in sms:
            sm2.close()

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on a bounded sample, single thread like one Pisces (BAM x chr) job; the same run
    # verifies the records of the timed workload on the sample's positions
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        s_loci = min(a.cpu_sample_loci, a.loci)
        w = reads_window(d, 0, s_loci)
        best, oc = None, None
        for _ in range(max(1, a.cpu_repeats)):
            cdt, oc = run_oracle(a, w)
            best = cdt if best is None else min(best, cdt)
        line["cpu_baseline"] = {"value": s_loci / best, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"reads starting in the first {s_loci} loci of the same workload, single thread, best of {max(1, a.cpu_repeats)} "
                                          "(find candidates + count + call, no I/O)"}
        if precs is not None:
            n_cmp, n_bad = compare_with_oracle(oc.records(), precs, arena, s_loci - d["read_len"] - 8)
            line["verified"] = {"records_compared_with_oracle": n_cmp, "mismatches": n_bad, "positions": f"1..{s_loci - d['read_len'] - 8}"}
    sm.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = resolve(parse())
    if args.impl == "reference":
        main_reference(args)
    else:
        main_ours(args)
