"""Sweep of the T-side plane kernel's work items per CTA.  usage: python scripts/exp_items.py [n_cells]"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
ctx = sb.Context(0)
y = np.random.default_rng(1).standard_normal((20, 33538))
for items in (48, 96):
    ctx.set_option("plane_items_per_cta", items)
    dm = generate_device(ctx, SynthConfig(n_cells=n, n_genes=33538, seed=3))
    a = sb.normalize(dm, sb.Normalization.CellRanger)
    a.rdot(y)
    ctx.profile_enable(True); ctx.profile_reset()
    for _ in range(5):
        a.rdot(y)
    p = ctx.profile(); ctx.profile_enable(False)
    print(f"n={n} plane_items_per_cta={items}: spmm_t {p['spmm_t_ms'] / p['spmm_t_launches']:.3f} ms/pass", flush=True)
    a.free(); dm.free()

for fc in (2.0, 5.0, 8.0, 12.0, 20.0, 40.0):
    ctx.set_option("plane_items_per_cta", 48)
    ctx.set_option("gather_flush_cost", fc)
    dm = generate_device(ctx, SynthConfig(n_cells=n, n_genes=33538, seed=3))
    a = sb.normalize(dm, sb.Normalization.CellRanger)
    a.rdot(y)
    ctx.profile_enable(True); ctx.profile_reset()
    for _ in range(5):
        a.rdot(y)
    p = ctx.profile(); ctx.profile_enable(False)
    print(f"n={n} gather_flush_cost={fc}: spmm_t {p['spmm_t_ms'] / p['spmm_t_launches']:.3f} ms/pass", flush=True)
    a.free(); dm.free()
ctx.set_option("gather_flush_cost", 5.0)
