"""A/B of the T-side gather with its run factor L_c(1) deferred to the plane kernels' reduction (option gather_defer).
usage: python scripts/exp_defer.py [n_cells]"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
ctx = sb.Context(0)
rng = np.random.default_rng(1)
y = rng.standard_normal((20, 33538))
y_wide = rng.standard_normal((50, 33538))
ref = ref_w = None
for defer in (0, 1, 0, 1):
    ctx.set_option("gather_defer", defer)
    dm = generate_device(ctx, SynthConfig(n_cells=n, n_genes=33538, seed=3))
    a = sb.normalize(dm, sb.Normalization.CellRanger)
    for _ in range(2):
        t = a.rdot(y)
    ctx.profile_enable(True); ctx.profile_reset()
    for _ in range(5):
        t = a.rdot(y)
    p = ctx.profile(); ctx.profile_enable(False)
    tw = a.rdot(y_wide)  # three column passes, the last one narrow
    if ref is None:
        ref, ref_w = t, tw
    d = float(np.abs(t - ref).max() / np.abs(ref).max())
    dw = float(np.abs(tw - ref_w).max() / np.abs(ref_w).max())
    print(f"n={n} gather_defer={defer}: spmm_t {p['spmm_t_ms'] / p['spmm_t_launches']:.3f} ms/pass, max rel diff vs first run: w=20 {d:.2e}, w=50 {dw:.2e}", flush=True)
    a.free(); dm.free()
