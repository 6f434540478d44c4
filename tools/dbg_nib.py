import sys; sys.path.insert(0, "/root/repo")
import numpy as np
import pisces_b200 as pb
from pisces_b200 import synth
d = synth.make_pileup(4000, 60, seed=60, snv_rate=0.03)
ref = bytes(d["ref_bases"].numpy()).decode()
off, code, qual, anch = (d[k].numpy() for k in ("offsets", "code", "qual", "anchor"))
sm = pb.GpuStateManager(pb.make_config(output_gvcf=0), "chr1", ref)
sm.AddPileup(off, code, qual, anch, first_position=1)
recs = pb.GpuAlleleCaller().Call(sm, raw=True)
print(len(recs))
