#!/usr/bin/env python
"""bench.py — candidate loci scored per second on synthetic pileups of the BASELINE.json shape.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--loci L] [--depth D] [--gvcf 0|1]

A step = one pass of the hot path (pileup count + score + record compaction) over one batch of `loci` synthetic pileup columns per GPU.
Default workload = BASELINE.json configs[1]: 1 M loci x depth ~Poisson(500), SNV (1 % of loci) + 1-3 bp insertions / deletions (0.1 % of loci,
explicit candidates with spanning coverage), flat Poisson noise model NL 20, gVCF off, one B200. N>1 (torchrun): loci are sharded by interval across ranks (weak scaling: `loci` per GPU), no data-path
collective; each step ends with the single all-gather of the per-rank call records (NCCL).
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--loci", type=int, default=1_000_000, help="loci per GPU")
    ap.add_argument("--depth", type=int, default=500)
    ap.add_argument("--depth-dist", default="poisson", choices=["poisson", "fixed"])
    ap.add_argument("--gvcf", type=int, default=0)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--indel-rate", type=float, default=0.001)
    ap.add_argument("--tune-ctas", type=int, default=0, help="hot-kernel CTAs per SM (tuning experiments)")
    ap.add_argument("--tune-prefetch", type=int, default=0, help="hot-kernel L2 prefetch distance (tuning experiments)")
    ap.add_argument("--cpu-sample-loci", type=int, default=200_000)
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


def workload_name(a):
    return (f"synthetic {a.loci} loci x depth {'~Poisson' if a.depth_dist == 'poisson' else '='}({a.depth}) per GPU, SNV 1% + indel {100 * a.indel_rate:g}%, "
            f"Poisson noise model NL20, gvcf={a.gvcf} (BASELINE.json configs[1] shape)")


def oracle_config(a):
    from oracle import binding as ob
    return ob.default_config(output_gvcf=a.gvcf, collapse=1)


def run_oracle_slices(a, d, n_threads, loci_per_thread):
    """Times the CPU restatement (count + call, no I/O) on n_threads disjoint slices, one thread each. Returns (seconds, loci)."""
    import numpy as np
    from oracle import binding as ob
    off = d["offsets"].cpu().numpy()
    slices = []
    for t in range(n_threads):
        l0, l1 = t * loci_per_thread, (t + 1) * loci_per_thread
        e0, e1 = int(off[l0]), int(off[l1])
        ref = bytes(d["ref_bases"][l0:l1].cpu().numpy()).decode()
        cands = []
        if d.get("candidates") is not None:
            arena = d["arena"]
            for c in d["candidates"]:
                p = int(c["position"])
                if l0 + 1 <= p and p + int(c["ref_len"]) + 1 <= l1:
                    o, rl, al = int(c["allele_offset"]), int(c["ref_len"]), int(c["alt_len"])
                    cands.append((int(c["type"]), p - l0, arena[o:o + rl].decode(), arena[o + rl:o + rl + al].decode(), [int(x) for x in c["support"]],
                                  [int(x) for x in c["well_anchored"]]))
        slices.append((ob.Caller(oracle_config(a), "chr1", ref), (off[l0:l1 + 1] - e0).astype(np.int64), d["code"][e0:e1].cpu().numpy(),
                       d["qual"][e0:e1].cpu().numpy(), d["anchor"][e0:e1].cpu().numpy(), cands))

    def work(s):
        c, o, co, q, an, cands = s
        for t, p, r, al, sup, wa in cands:
            c.add_candidate(t, p, r, al, sup, wa)
        c.add_pileup(o, co, q, an, 1, call_every=1)
        c.finish()
    ths = [threading.Thread(target=work, args=(s,)) for s in slices]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    return time.perf_counter() - t0, n_threads * loci_per_thread


def main_reference(a):
    """Reference arm: the reference's CPU implementation of the path. The C# cannot run here (no dotnet runtime in the image), so this is
    the oracle port (oracle/) on all host threads, each step a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from pisces_b200 import synth
    cores = os.cpu_count() or 1
    per_thread = max(1000, min(20_000, a.loci // cores))
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    d = synth.make_pileup(cores * per_thread, a.depth, seed=a.seed, device=dev, depth_dist=a.depth_dist, indel_rate=a.indel_rate)
    d = {k: (v.cpu() if hasattr(v, "cpu") else v) for k, v in d.items()}
    times = []
    for i in range(a.warmup + a.steps):
        dt, loci = run_oracle_slices(a, d, cores, per_thread)
        if i >= a.warmup:
            times.append(dt)
    total = sum(times)
    value = a.steps * cores * per_thread / total
    sample = f"{cores} threads x {per_thread} loci per step ({cores * per_thread} loci, depth {a.depth}) of the same synthetic workload"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "note": "oracle port of the C# path (dotnet runtime absent); count+call only, no I/O"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main_ours(a):
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = f"cuda:{local}"

    # ---- synthetic shard for this rank, resident in HBM (interval shard `rank` of `world`)
    d = synth.make_pileup(a.loci, a.depth, seed=a.seed + 1000 * rank, device=dev, depth_dist=a.depth_dist, indel_rate=a.indel_rate)
    n_entries = d["n_entries"]
    ref = bytes(d["ref_bases"].cpu().numpy())
    cfg = pb.make_config(device=local, output_gvcf=a.gvcf)
    cfg.reserved[0] = a.tune_ctas
    cfg.reserved[1] = a.tune_prefetch
    sm = pb.GpuStateManager(cfg, "chr1", ref)
    sm.AddPileup(d["offsets"], d["code"], d["qual"], d["anchor"], first_position=1, ref_bases=d["ref_bases"], device=True)
    if d.get("candidates") is not None:
        sm.AddCandidates(d["candidates"], d["arena"])
    torch.cuda.synchronize()

    import ctypes as C
    from pisces_b200 import _native as N

    class DevBuf:   # torch view of a device buffer owned by the library
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    # ---- the single all-gather of per-interval call records (NCCL) at the end of the job: a step appends its variant records to the
    # rank's job buffer (one device copy out of the library's variant stream, which lives at a fixed device address, so its torch view is created
    # once); when the K steps are done the ranks exchange [K int64 counts | K x fixed-capacity record blocks] in ONE all_gather, inside the timed region.
    job_buf = gather_out = var_view = counts_pinned = None
    n_slots = max(1, a.steps)
    if world > 1:
        n0 = sm.call_resident()
        n0_all = torch.tensor([n0], dtype=torch.int64, device=dev)
        dist.all_reduce(n0_all, op=dist.ReduceOp.MAX)    # the ranks' shards differ: the block capacity (hence the gathered size) must be the same on all
        n0 = int(n0_all.item())
        cap_records = (n0 + n0 // 8 + 127) // 64 * 64   # fixed capacity of a step's block: the largest shard's record count plus 12 % head-room
        vr, nv = C.c_void_p(), C.c_int64()
        sm._chk(sm._L.pb2_resident_results(sm._h, None, None, None, C.byref(vr), C.byref(nv)))
        var_view = torch.as_tensor(DevBuf(vr.value, cap_records * 96), device=dev)
        job_buf = torch.zeros(8 * n_slots + n_slots * cap_records * 96, dtype=torch.uint8, device=dev)
        gather_out = torch.zeros(world * job_buf.numel(), dtype=torch.uint8, device=dev)
        counts_pinned = torch.zeros(n_slots, dtype=torch.int64).pin_memory()

    def step(k=0):
        n = sm.call_resident()   # returns after the step's counters are on the host: the library's stream is idle, the variant stream complete
        if world > 1:
            slot = k % n_slots
            counts_pinned[slot] = min(n, cap_records)
            o = 8 * n_slots + slot * cap_records * 96
            job_buf[o:o + cap_records * 96].copy_(var_view, non_blocking=True)
        return n

    def gather_job():
        if world > 1:
            job_buf[:8 * n_slots].copy_(counts_pinned.view(torch.uint8), non_blocking=True)
            dist.all_gather_into_tensor(gather_out, job_buf)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(max(3, a.warmup)):
        n_records = step(k)
    gather_job()   # untimed: NCCL sets its channels up on the first collective
    sm.stats()
    sampler = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    ev0.record()
    t0 = time.perf_counter()
    blocks = []   # hot-kernel time of every 10 timed steps (the library's own CUDA events): shows a machine-state change inside the timed region
    for k in range(a.steps):
        n_records = step(k)
        if (k + 1) % 10 == 0 or k + 1 == a.steps:
            blocks.append(sm.stats())
    gather_job()
    barrier()
    ev1.record()
    ev1.synchronize()
    dt_wall = time.perf_counter() - t0
    dt = ev0.elapsed_time(ev1) * 1e-3   # device clock between the two synchronised brackets (the wall clock beside it: wall_ms_per_step)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    st = {k: sum(b[k] for b in blocks) for k in ("hot_launches", "hot_ms", "total_launches")}
    per_block = sorted(b["hot_ms"] / max(1, b["hot_launches"]) for b in blocks)
    tmax = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dt = float(tmax.item())
    value = world * a.loci * a.steps / dt

    # ---- roofline of the dominant (only) kernel: algorithmic bytes / CUDA-event duration on the launching stream
    n_ref_records = a.loci if a.gvcf else 0
    # point alleles (SNV / reference) never read the anchor/collapsed byte: without collapsed-read tracking the hot kernel skips that plane,
    # and the algorithmic bytes are SURVEY 8d's 2*D + 8 + 96*E form
    third_byte = bool(cfg.expect_collapsed)
    algo_bytes = synth.algorithmic_bytes(a.loci, n_entries, n_records + n_ref_records, third_byte=third_byte)
    hot_ms = st["hot_ms"] / max(1, st["hot_launches"])
    achieved = algo_bytes / (hot_ms * 1e-3) / 1e9
    peak, peak_src = 6650.0, "fallback"
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        peak_src = "measured"
    except Exception:
        pass
    # which hot kernel ran: the PNIB16 one (direction-split, nibble-packed pileup: 1.5 staged bytes per entry) unless collapsed-read tracking / quality
    # sums are on or it is switched off for comparison (--tune-prefetch 9 -> the PTILE32 vertical-counter kernel, 2 staged bytes per entry)
    nib = not cfg.expect_collapsed and not cfg.want_sum_base_quality and cfg.noise_model != 1 and a.tune_prefetch != 9
    kernel_name = "pileup_nib_score_kernel" if nib else "pileup_vcount_score_kernel"
    traffic = None
    try:   # dram__bytes_read + dram__bytes_write of one launch from the committed ncu --set full capture of this command (default workload only)
        if a.loci == 1_000_000 and a.depth == 500 and a.indel_rate == 0.001 and (nib or not a.gvcf):
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))[kernel_name + ("_gvcf" if a.gvcf else "")]["dram_bytes_per_launch"]
    except Exception:
        pass

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup), "ms_per_step": 1e3 * dt / a.steps, "wall_ms_per_step": 1e3 * dt_wall / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "loci_per_gpu": a.loci, "entries_per_gpu": n_entries, "records_per_step": n_records + n_ref_records,
                       "l2": "staged input of the hot kernel (%.2f GB per GPU) larger than L2, no flush needed" % ((1.5 if nib else 2) * n_entries / 1e9), "parallelism": f"interval-sharded x{world}"},
            "gpu_launches": st["total_launches"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": kernel_name, "algorithmic_bytes_per_launch": algo_bytes, "bytes_per_entry": 3 if third_byte else 2,
                         "staged_bytes_per_entry": 1.5 if nib else (3 if third_byte else 2),
                         "dram_frac": (traffic / (hot_ms * 1e-3) / 1e9 / peak) if traffic else None,
                         "kernel_ms": hot_ms,
                         # the same kernel over blocks of 10 timed steps: on this pool the kernel runs at one of two plateaus (x1.0 / x1.5) that
                         # switch on a seconds scale with no clock change or throttle reason reported (profiles/r1_summary.md, "machine state")
                         "kernel_ms_blocks": {"min": per_block[0], "median": per_block[len(per_block) // 2], "max": per_block[-1]}},
            "clocks": sampler.summary()}

    # ---- end to end through the C ABI with HOST buffers: pinned H2D of the step's pileup, tile staging, call, D2H of the records
    if not a.no_e2e:
        # host buffers in PB2_LAYOUT_PACKED2 (2 bytes per entry + the sparse candidate flags): what a host behind a PCIe link hands to pb2_push_pileup
        pc, pq, fi, fb = pb.GpuStateManager.pack_pileup(d["code"].cpu().numpy(), d["qual"].cpu().numpy(), d["anchor"].cpu().numpy(), d["offsets"].cpu().numpy(),
                                                        d["ref_bases"].cpu().numpy())
        h = {"offsets": d["offsets"].cpu().pin_memory(), "ref_bases": d["ref_bases"].cpu().pin_memory(), "pcode": torch.from_numpy(pc).pin_memory(),
             "pqual": torch.from_numpy(pq).pin_memory(), "flag_index": torch.from_numpy(fi).pin_memory(), "flag_bits": torch.from_numpy(fb).pin_memory()}
        del pc, pq
        sm2 = pb.GpuStateManager(cfg, "chr1", ref)
        caller = pb.GpuAlleleCaller()
        e2e_steps = max(2, min(a.steps, 4))

        def e2e_step():
            sm2.AddPileupPacked(h["offsets"].numpy(), h["pcode"].numpy(), h["pqual"].numpy(), h["flag_index"].numpy(), h["flag_bits"].numpy(), first_position=1,
                                ref_bases=h["ref_bases"].numpy())
            if d.get("candidates") is not None:
                sm2.AddCandidates(d["candidates"], d["arena"])
            recs = caller.Call(sm2, raw=True)
            sm2.DoneProcessing()
            return len(recs)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            nrec = e2e_step()
        barrier()
        edt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(edt, op=dist.ReduceOp.MAX)
        line["e2e"] = {"value": world * a.loci * e2e_steps / float(edt.item()), "unit": UNIT, "h2d_bytes_per_step": 2 * n_entries + 8 * (a.loci + 1) + a.loci + 9 * int(h["flag_index"].numel()),
                       "d2h_bytes_per_step": 96 * nrec, "steps": e2e_steps}
        sm2.close()

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on a bounded sample, single thread like one Pisces (BAM x chr) job
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        s_loci = min(a.cpu_sample_loci, a.loci)
        dd = {k: (v[: int(d["offsets"][s_loci]) if k in ("code", "qual", "anchor") else s_loci + 1 if k == "offsets" else s_loci].cpu() if hasattr(v, "cpu") else v)
              for k, v in d.items()}
        cdt, cl = run_oracle_slices(a, dd, 1, s_loci)
        line["cpu_baseline"] = {"value": cl / cdt, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"first {s_loci} loci of the same workload, single thread (count + call, no I/O)"}
    sm.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_ours(args)
