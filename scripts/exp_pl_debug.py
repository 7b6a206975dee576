import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device
n = 400_000
ctx = sb.Context(0)
dm = generate_device(ctx, SynthConfig(n_cells=n, n_genes=33538, seed=3))
a = sb.normalize(dm, sb.Normalization.CellRanger)
y = np.random.default_rng(1).standard_normal((20, 33538))
for dbg in (0, 1, 2, 3, 4, 7):
    ctx.set_option("pl_debug", dbg)
    a.rdot(y)
    ctx.profile_enable(True); ctx.profile_reset()
    for _ in range(5):
        a.rdot(y)
    p = ctx.profile(); ctx.profile_enable(False)
    print(f"pl_debug={dbg}: spmm_t {p['spmm_t_ms'] / p['spmm_t_launches']:.3f} ms/pass (includes gather ~0.93)", flush=True)
