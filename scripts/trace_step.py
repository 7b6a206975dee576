import sys, time
sys.path.insert(0, "/root/repo")
import scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
ctx = sb.Context(0)
dm = generate_device(ctx, SynthConfig(n_cells=n, n_genes=33538, seed=3))
out = sb.pinned_outputs(33538, n, 10)
for i in range(3):
    print("---- step", i, file=sys.stderr, flush=True)
    t0 = time.time()
    a = sb.normalize(dm, sb.Normalization.CellRanger); ctx.sync(); t1 = time.time()
    sb.BkSvd().run_pca(a, 10, out=out); t2 = time.time()
    a.free()
    print(f"normalize {1e3*(t1-t0):.1f} ms, pca {1e3*(t2-t1):.1f} ms", file=sys.stderr, flush=True)
