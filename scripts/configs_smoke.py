"""Runs the other BASELINE.json configs (scaled to fit a short GPU session) through the public API and
prints timings + sanity checks: C2 (HVG 2000, k=50), C4-like (k=100 on one GPU, reduced cells),
C5 (60k features with dense antibody rows, k=30)."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device

def run(name, cfg, k, hvg=None):
    t0 = time.time()
    dm = generate_device(ctx, cfg)
    t1 = time.time()
    if hvg:
        sel = dm.hvg_select(hvg)
        dm2 = dm.select_rows(sel)
        dm.free(); dm = dm2
    ctx.sync(); t2 = time.time()
    for rep in range(2):  # first call pages in / lazily loads the cuSOLVER kernels for this shape
        ctx.timer_begin()
        a = sb.normalize(dm, sb.Normalization.CellRanger)
        u, s, v = sb.BkSvd().run_pca(a, k)
        ms = ctx.timer_end()
        if rep == 0:
            first_ms = ms
            a.free()
    ok_u = np.abs(u.T @ u - np.eye(k)).max(); ok_v = np.abs(v.T @ v - np.eye(k)).max()
    ident = np.abs(a.rdot(u.T[:3]).T - v[:, :3] * s[:3]).max() / s[0] if dm.rows() < dm.cols() else float("nan")
    print(f"{name}: shape {dm.shape()} nnz {dm.nnz()} k={k}: gen {t1-t0:.1f}s prep {t2-t1:.2f}s normalize+PCA first call {first_ms:.0f} ms, second {ms:.1f} ms "
          f"-> {dm.cols()/ms*1e3:.3g} cells/s; |U'U-I| {ok_u:.1e} |V'V-I| {ok_v:.1e} A'u=vs {ident:.1e} sigma[:3] {s[:3]}", flush=True)
    a.free(); dm.free()

ctx = sb.Context(0)
run("C2 100k x 33538, HVG 2000, k=50", SynthConfig(n_cells=100_000, n_genes=33538, seed=2), 50, hvg=2000)
run("C5 500k x 60k (200 dense antibody features), k=30", SynthConfig(n_cells=500_000, n_genes=60_000, sigma_g=3.0, seed=5, n_dense=200), 30)
run("C4-like 500k x 36601, k=100 (one GPU's share of 4M/8)", SynthConfig(n_cells=500_000, n_genes=36_601, seed=4), 100)
ctx.close()
