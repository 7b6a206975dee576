"""A/B of the host top-k eigensolver for the Gram matrix (option eig_host) on whole normalize + PCA steps.
usage: python scripts/exp_eig.py [n_cells]"""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
ctx = sb.Context(0)
dm = generate_device(ctx, SynthConfig(n_cells=n, n_genes=33538, seed=3))
out = sb.pinned_outputs(33538, n, 10)
ref = None
for host in (0, 1, 0, 1):
    ctx.set_option("eig_host", host)
    for _ in range(2):
        a = sb.normalize(dm, sb.Normalization.CellRanger); r = sb.BkSvd().run_pca(a, 10, out=out); a.free()
    ctx.profile_enable(True); ctx.profile_reset(); ctx.sync(); ctx.timer_begin()
    for _ in range(5):
        a = sb.normalize(dm, sb.Normalization.CellRanger); r = sb.BkSvd().run_pca(a, 10, out=out); a.free()
    ms = ctx.timer_end() / 5
    p = ctx.profile(); ctx.profile_enable(False)
    s = np.array(r[1])
    if ref is None:
        ref = s
    print(f"n={n} eig_host={host}: {ms:.2f} ms/step, dense {p['dense_ms'] / 5:.2f} ms, sigma rel diff vs first {np.abs(s - ref).max() / ref.max():.2e}", flush=True)
