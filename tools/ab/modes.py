"""Which state does the bimodal hot-kernel time follow? One process: stage the same pileup repeatedly (same handle after pb2_reset, then new handles),
time 40 resident steps each, print the per-staging kernel time. tools/ab experiment, not part of the product."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pisces_b200 as pb
from pisces_b200 import synth

gv = int(sys.argv[1]) if len(sys.argv) > 1 else 0
d = synth.make_pileup(1_000_000, 500, seed=2, device="cuda:0", depth_dist="poisson", indel_rate=0.001)
ref = bytes(d["ref_bases"].cpu().numpy())
cfg = pb.make_config(device=0, output_gvcf=gv)


def stage(sm):
    sm.AddPileup(d["offsets"], d["code"], d["qual"], d["anchor"], first_position=1, ref_bases=d["ref_bases"], device=True)
    if d.get("candidates") is not None:
        sm.AddCandidates(d["candidates"], d["arena"])
    torch.cuda.synchronize()


def measure(sm, n=40):
    for _ in range(5):
        sm.call_resident()
    sm.stats()
    t0 = time.perf_counter()
    for _ in range(n):
        sm.call_resident()
    dt = time.perf_counter() - t0
    st = sm.stats()
    return round(st["hot_ms"] / st["hot_launches"], 4), round(1e3 * dt / n, 4)


sm = pb.GpuStateManager(cfg, "chr1", ref)
for i in range(4):
    stage(sm)
    print("same handle, staging", i, measure(sm), flush=True)
    sm.DoneProcessing()
sm.close()
keep = []
for i in range(4):
    sm = pb.GpuStateManager(cfg, "chr1", ref)
    stage(sm)
    print("new handle", i, measure(sm), measure(sm), flush=True)
    keep.append(sm)
    if i % 2 == 1:
        junk = torch.empty(300_000_000 * (i + 1), dtype=torch.uint8, device="cuda:0")   # shift later allocations
for sm in keep:
    print("revisit", measure(sm), flush=True)
