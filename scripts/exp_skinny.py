"""A/B of the row-per-thread kernel for tall products with <= 16 output columns (option gemm_skinny) on whole normalize + PCA steps.
usage: python scripts/exp_skinny.py [n_cells]"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
ctx = sb.Context(0)
dm = generate_device(ctx, SynthConfig(n_cells=n, n_genes=33538, seed=3))
out = sb.pinned_outputs(33538, n, 10)
ref = None
for on in (0, 1, 0, 1):
    ctx.set_option("gemm_skinny", on)
    for _ in range(2):
        a = sb.normalize(dm, sb.Normalization.CellRanger); r = sb.BkSvd().run_pca(a, 10, out=out); a.free()
    ctx.profile_enable(True); ctx.profile_reset(); ctx.sync(); ctx.timer_begin()
    for _ in range(5):
        a = sb.normalize(dm, sb.Normalization.CellRanger); r = sb.BkSvd().run_pca(a, 10, out=out); a.free()
    ms = ctx.timer_end() / 5
    p = ctx.profile(); ctx.profile_enable(False)
    res = [np.array(x) for x in r]
    if ref is None:
        ref = res
    d = [float(np.abs(np.abs(x) - np.abs(y)).max()) for x, y in zip(res, ref)]
    print(f"n={n} gemm_skinny={on}: {ms:.2f} ms/step, dense {p['dense_ms'] / 5:.2f} ms, max |abs diff| vs first: U {d[0]:.1e} sigma {d[1]:.1e} V {d[2]:.1e}", flush=True)
