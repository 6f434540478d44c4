"""Hot-kernel time against time under sustained load, with NVML clocks next to it. tools/ab experiment, not part of the product."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pynvml as nv
import pisces_b200 as pb
from pisces_b200 import synth

nv.nvmlInit()
hd = nv.nvmlDeviceGetHandleByIndex(0)
d = synth.make_pileup(1_000_000, 500, seed=2, device="cuda:0", depth_dist="poisson", indel_rate=0.001)
ref = bytes(d["ref_bases"].cpu().numpy())
sm = pb.GpuStateManager(pb.make_config(device=0, output_gvcf=0), "chr1", ref)
sm.AddPileup(d["offsets"], d["code"], d["qual"], d["anchor"], first_position=1, ref_bases=d["ref_bases"], device=True)
sm.AddCandidates(d["candidates"], d["arena"])
torch.cuda.synchronize()
time.sleep(float(sys.argv[1]) if len(sys.argv) > 1 else 0)
# probes of the machine state next to the kernel: a 1 GB device copy (memory system) and a 4096^3 bf16 matmul (SM clock / power)
pa = torch.empty(1 << 29, dtype=torch.uint8, device="cuda:0"); pb_ = torch.empty_like(pa)
ma = torch.randn(4096, 4096, dtype=torch.bfloat16, device="cuda:0"); mb = torch.randn(4096, 4096, dtype=torch.bfloat16, device="cuda:0")
def probe():
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); pb_.copy_(pa); e[1].record(); torch.matmul(ma, mb); e[2].record(); torch.cuda.synchronize()
    return 2 * pa.numel() / (e[0].elapsed_time(e[1]) * 1e-3) / 1e9, 2 * 4096 ** 3 / (e[1].elapsed_time(e[2]) * 1e-3) / 1e12
probe()
t_start = time.perf_counter()
nblk = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for blk in range(nblk):
    if blk == nblk // 2 and len(sys.argv) > 3:
        time.sleep(float(sys.argv[3]))
    sm.stats()
    for _ in range(int(sys.argv[4]) if len(sys.argv) > 4 else 100):
        sm.call_resident()
    st = sm.stats()
    cg, mt = probe()
    print(f"t={time.perf_counter() - t_start:6.3f}s kernel_ms={st['hot_ms'] / st['hot_launches']:.4f} sm={nv.nvmlDeviceGetClockInfo(hd, nv.NVML_CLOCK_SM)} "
          f"mem={nv.nvmlDeviceGetClockInfo(hd, nv.NVML_CLOCK_MEM)} power={nv.nvmlDeviceGetPowerUsage(hd) / 1000:.0f}W temp={nv.nvmlDeviceGetTemperature(hd, 0)} "
          f"copy={cg:.0f}GB/s matmul={mt:.0f}TF reasons={nv.nvmlDeviceGetCurrentClocksEventReasons(hd):#x}", flush=True)
