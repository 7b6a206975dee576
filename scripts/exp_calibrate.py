"""A/B of the T-side gather's timed re-cut of its static shares (option gather_calibrate = number of rounds).
usage: python scripts/exp_calibrate.py [n_cells]"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
ctx = sb.Context(0)
y = np.random.default_rng(1).standard_normal((20, 33538))
ref = None
for rounds in (0, 1, 2, 3):
    ctx.set_option("gather_calibrate", rounds)
    dm = generate_device(ctx, SynthConfig(n_cells=n, n_genes=33538, seed=3))
    a = sb.normalize(dm, sb.Normalization.CellRanger)
    for _ in range(rounds + 1):
        t = a.rdot(y)
    ctx.profile_enable(True); ctx.profile_reset()
    for _ in range(5):
        t = a.rdot(y)
    p = ctx.profile(); ctx.profile_enable(False)
    if ref is None:
        ref = t
    d = float(np.abs(t - ref).max() / np.abs(ref).max())
    print(f"n={n} gather_calibrate={rounds}: spmm_t {p['spmm_t_ms'] / p['spmm_t_launches']:.3f} ms/pass, max rel diff vs uncalibrated {d:.2e}", flush=True)
    a.free(); dm.free()
