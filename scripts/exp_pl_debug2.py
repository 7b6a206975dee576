"""Decomposition of k_planes_t by switching parts of it off (pl_debug bits: 1 no output stores, 2 no epilogue arithmetic, 4 no tile
expansion; results are wrong on purpose) for both kernel generations.  usage: python scripts/exp_pl_debug2.py [n_cells]"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
ctx = sb.Context(0)
y = np.random.default_rng(1).standard_normal((20, 33538))
dm = generate_device(ctx, SynthConfig(n_cells=n, n_genes=33538, seed=3))
a = sb.normalize(dm, sb.Normalization.CellRanger)
for variant in (0, 3):
    ctx.set_option("pl_variant", variant)
    for dbg in (0, 1, 2, 3, 4, 6, 7):
        ctx.set_option("pl_debug", dbg)
        a.rdot(y)
        ctx.profile_enable(True); ctx.profile_reset()
        for _ in range(4):
            a.rdot(y)
        p = ctx.profile(); ctx.profile_enable(False)
        print(f"pl_variant {variant} pl_debug {dbg}: spmm_t {p['spmm_t_ms'] / p['spmm_t_launches']:.3f} ms/pass", flush=True)
ctx.set_option("pl_debug", 0)
