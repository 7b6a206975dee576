"""Sweep of the plane selection (plane_min_density, plane_cap) at n_cells: per-pass times of both products.
usage: python scripts/exp_plane_density.py [n_cells]"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
ctx = sb.Context(0)
x = np.random.default_rng(0).standard_normal((n, 20))
y = np.random.default_rng(1).standard_normal((20, 33538))
for dens, cap in ((0.02, 12288), (0.01, 12288), (0.005, 16384), (0.0025, 20480), (0.001, 24576)):
    ctx.set_option("plane_min_density", dens)
    ctx.set_option("plane_cap", cap)
    dm = generate_device(ctx, SynthConfig(n_cells=n, n_genes=33538, seed=3))
    a = sb.normalize(dm, sb.Normalization.CellRanger)
    a.dot(x); a.rdot(y)
    ctx.profile_enable(True); ctx.profile_reset()
    for _ in range(5):
        a.dot(x); a.rdot(y)
    p = ctx.profile(); ctx.profile_enable(False)
    print(f"plane_min_density={dens} cap={cap}: spmm_t {p['spmm_t_ms'] / p['spmm_t_launches']:.3f} ms/pass  spmm_n {p['spmm_n_ms'] / p['spmm_n_launches']:.3f} ms/pass", flush=True)
    a.free(); dm.free()
