import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device
ctx = sb.Context(0)
which = sys.argv[1]
if which == "c2":
    dm = generate_device(ctx, SynthConfig(n_cells=100_000, n_genes=33538, seed=2))
    sel = dm.hvg_select(2000); dm2 = dm.select_rows(sel); dm.free(); dm = dm2; k = 50
else:
    dm = generate_device(ctx, SynthConfig(n_cells=200_000, n_genes=60_000, sigma_g=3.0, seed=5, n_dense=200)); k = 30
ip, g, c = dm.to_csc()
print("shape", dm.shape(), "nnz", dm.nnz(), "max count", c.max(), "frac>15", (c > 15).mean(), flush=True)
ctx.profile_enable(True); ctx.profile_reset()
t0 = time.time()
a = sb.normalize(dm, sb.Normalization.CellRanger); ctx.sync(); t1 = time.time()
print("normalize", t1 - t0, {k_: round(v, 2) for k_, v in ctx.profile().items() if k_.endswith("_ms")}, flush=True)
ctx.profile_reset()
u, s, v = sb.BkSvd().run_pca(a, k); t2 = time.time()
print("pca", t2 - t1, {k_: round(v, 2) for k_, v in ctx.profile().items() if k_.endswith("_ms")}, ctx.profile()["spmm_t_launches"], flush=True)
