"""Where the end-to-end time goes: host wall clock around upload / normalize / PCA / free, with and without a device
synchronisation after every call (the bench's e2e arm is the unsynchronised sequence)."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np, scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
ctx = sb.Context(0)
dm = generate_device(ctx, SynthConfig(n_cells=n, n_genes=33538, seed=3))
ip, g, c = dm.to_csc()
def pin(a):
    t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True); v = t.numpy().view(a.dtype); v[:] = a; return v, t
g16, c8, bp, bc = sb.AdaptiveMat.compact_csc(g, c)
hip, k1 = pin(ip); hg, k2 = pin(g16); hc, k3 = pin(c8)
del g, c, g16, c8
out = sb.pinned_outputs(33538, n, 10)
for sync in (True, False, True, False, False):
    t = [time.perf_counter()]
    ctx.timer_begin()
    m = sb.AdaptiveMat.from_csc_compact(ctx, 33538, n, hip, hg, hc, bp, bc)
    if sync: ctx.sync()
    t.append(time.perf_counter())
    a = sb.normalize(m, sb.Normalization.CellRanger)
    if sync: ctx.sync()
    t.append(time.perf_counter())
    r = sb.BkSvd().run_pca(a, 10, out=out)
    if sync: ctx.sync()
    t.append(time.perf_counter())
    a.free(); m.free()
    if sync: ctx.sync()
    t.append(time.perf_counter())
    ms = ctx.timer_end()
    d = np.diff(t) * 1e3
    print(f"sync={sync}: upload {d[0]:.1f} normalize {d[1]:.1f} pca {d[2]:.1f} free {d[3]:.1f} | wall {sum(d):.1f} ms, device timer {ms:.1f} ms", flush=True)
