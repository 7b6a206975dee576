import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np, scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device
ctx = sb.Context(0)
cfg = SynthConfig(n_cells=1_300_000, n_genes=33538, seed=3)
dm = generate_device(ctx, cfg)
ip, g, c = dm.to_csc(); dm.free()
import torch
def pin(a):
    t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True); v = t.numpy().view(a.dtype); v[:] = a; return v, t
hip, k1 = pin(ip); hg, k2 = pin(g); hc, k3 = pin(c)
for i in range(2):
    print("---- upload", i, file=sys.stderr)
    t0=time.time(); m = sb.AdaptiveMat.from_csc(ctx, 33538, 1_300_000, hip, hg, hc); ctx.sync(); print("total", time.time()-t0, file=sys.stderr); m.free()
