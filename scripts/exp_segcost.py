"""Sweep of the T-side gather's per-segment cost term.  usage: python scripts/exp_segcost.py [n_cells]"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
ctx = sb.Context(0)
y = np.random.default_rng(1).standard_normal((20, 33538))
for sc in (0.0, 200.0, 500.0, 1000.0, 2000.0, 4000.0):
    ctx.set_option("gather_seg_cost", sc)
    dm = generate_device(ctx, SynthConfig(n_cells=n, n_genes=33538, seed=3))
    a = sb.normalize(dm, sb.Normalization.CellRanger)
    a.rdot(y)
    ctx.profile_enable(True); ctx.profile_reset()
    for _ in range(5):
        a.rdot(y)
    p = ctx.profile(); ctx.profile_enable(False)
    print(f"n={n} gather_seg_cost={sc}: spmm_t {p['spmm_t_ms'] / p['spmm_t_launches']:.3f} ms/pass", flush=True)
    a.free(); dm.free()
