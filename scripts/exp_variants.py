"""A/B of the kernel generations behind one product: planes.cu variants (pl_variant bit 0 = T side, bit 1 = N side) and the
T-side gather's work items per CTA (1 = static shares).  Event-timed passes at n_cells, results compared with the first setting.
usage: python scripts/exp_variants.py [n_cells]"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
ctx = sb.Context(0)
x = np.random.default_rng(0).standard_normal((n, 20))
y = np.random.default_rng(1).standard_normal((20, 33538))
ref = None
for items in (1,):
    ctx.set_option("gather_items_per_cta", items)
    dm = generate_device(ctx, SynthConfig(n_cells=n, n_genes=33538, seed=3))
    a = sb.normalize(dm, sb.Normalization.CellRanger)
    for variant in (0, 3, 0, 3):
        ctx.set_option("pl_variant", variant)
        pn, pt = a.dot(x), a.rdot(y)
        ctx.profile_enable(True); ctx.profile_reset()
        for _ in range(5):
            a.dot(x); a.rdot(y)
        p = ctx.profile(); ctx.profile_enable(False)
        line = (f"items/cta {items} pl_variant {variant} n={n}: spmm_t {p['spmm_t_ms'] / p['spmm_t_launches']:.3f} ms/pass  "
                f"spmm_n {p['spmm_n_ms'] / p['spmm_n_launches']:.3f} ms/pass")
        if ref is None:
            ref = (pn, pt)
        else:
            line += f"   vs first: N {np.abs(pn - ref[0]).max() / np.abs(ref[0]).max():.1e}  T {np.abs(pt - ref[1]).max() / np.abs(ref[1]).max():.1e}"
        print(line, flush=True)
    a.free(); dm.free()
