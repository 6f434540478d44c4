import sys, time; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import pisces_b200 as pb
from pisces_b200 import synth
d = synth.make_pileup(1_000_000, 500, seed=2, device='cuda')
h = {k: d[k].cpu().pin_memory() for k in ("offsets","code","qual","anchor","ref_bases")}
ref = bytes(d["ref_bases"].cpu().numpy())
sm = pb.GpuStateManager(pb.make_config(output_gvcf=0), "chr1", ref)
caller = pb.GpuAlleleCaller()
for it in range(4):
    torch.cuda.synchronize(); t0=time.perf_counter()
    sm.AddPileup(h["offsets"].numpy(), h["code"].numpy(), h["qual"].numpy(), h["anchor"].numpy(), first_position=1, ref_bases=h["ref_bases"].numpy())
    t1=time.perf_counter()
    recs = caller.Call(sm, raw=True)
    t2=time.perf_counter()
    sm.DoneProcessing()
    t3=time.perf_counter()
    print(f"push {1e3*(t1-t0):.1f} ms  call+flush {1e3*(t2-t1):.1f} ms  reset {1e3*(t3-t2):.1f} ms  total {1e3*(t3-t0):.1f}")
# raw H2D rate
x = torch.empty(1_500_000_000, dtype=torch.uint8, device='cuda'); hp = torch.empty(1_500_000_000, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize(); t0=time.perf_counter(); x.copy_(hp, non_blocking=True); torch.cuda.synchronize(); print('H2D 1.5GB pinned ms', 1e3*(time.perf_counter()-t0))
