"""Stage trace of a 1.3M-cell upload.  SCANB200_TRACE=1: unpipelined path; SCANB200_TRACE=2: pipelined path (stage
times are synchronising, so the overlap is lost but the per-chunk device work shows).  argv[1] = packed|compact|plain."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np, scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device
mode = sys.argv[1] if len(sys.argv) > 1 else "compact"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_300_000
ctx = sb.Context(0)
cfg = SynthConfig(n_cells=n, n_genes=33538, seed=3)
dm = generate_device(ctx, cfg)
ip, g, c = dm.to_csc(); dm.free()
import torch
def pin(a):
    t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True); v = t.numpy().view(a.dtype); v[:] = a; return v, t
hip, k1 = pin(ip)
if mode == "packed":
    t0 = time.time()
    pk = sb.AdaptiveMat.pack_csc(ip, g, c, pinned=True)
    print(f"pack_csc: {time.time() - t0:.2f} s, bytes per entry {sum(a.nbytes for a in pk) / len(g):.3f}, escapes {len(pk[2])}, big {len(pk[4])}", file=sys.stderr)
elif mode == "compact":
    g16, c8, bp, bc = sb.AdaptiveMat.compact_csc(g, c)
    hg, k2 = pin(g16); hc, k3 = pin(c8)
else:
    hg, k2 = pin(g); hc, k3 = pin(c)
chunk_sweep = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [8]
for i in range(3 * len(chunk_sweep)):
    ctx.set_option("upload_chunks", chunk_sweep[i // 3])
    print("---- upload", i, mode, "chunks", chunk_sweep[i // 3], file=sys.stderr)
    t0 = time.time()
    if mode == "packed":
        m = sb.AdaptiveMat.from_csc_packed(ctx, 33538, n, hip, *pk)
    elif mode == "compact":
        m = sb.AdaptiveMat.from_csc_compact(ctx, 33538, n, hip, hg, hc, bp, bc)
    else:
        m = sb.AdaptiveMat.from_csc(ctx, 33538, n, hip, hg, hc)
    ctx.sync(); print("total", time.time() - t0, file=sys.stderr); m.free()
