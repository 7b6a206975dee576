"""Round-2 starting point: A/B of the experimental tcgen05 int8 T-side panel kernel (csrc/panel_i8.cu, never run on
hardware when it was written) against the FP64 mma.sync panel, on a matrix whose panel keeps counts 1..3 only.
Prints the product difference first (small sizes), then event-timed passes.  usage: python scripts/exp_panel_i8.py [n_cells]"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
ctx = sb.Context(0)
res = {}
for cap in (15, 3):
    ctx.set_option("dense_max_count", cap)
    dm = generate_device(ctx, SynthConfig(n_cells=n, n_genes=33538, seed=3))
    a = sb.normalize(dm, sb.Normalization.CellRanger)
    y = np.random.default_rng(0).standard_normal((20, 33538))
    for i8 in ((0,) if cap == 15 else (0, 1)):
        ctx.set_option("panel_i8", i8)
        t = a.rdot(y)  # warm + result
        ctx.profile_enable(True); ctx.profile_reset()
        for _ in range(5):
            a.rdot(y)
        p = ctx.profile(); ctx.profile_enable(False)
        res[(cap, i8)] = t
        print(f"dense_max_count={cap} panel_i8={i8}: spmm_t {p['spmm_t_ms'] / p['spmm_t_launches']:.3f} ms/pass", flush=True)
    ctx.set_option("panel_i8", 0)
    a.free(); dm.free()
ref = res[(15, 0)]
for k, v in res.items():
    print(k, "max rel diff vs the 15-count FP64 panel:", float(np.abs(v - ref).max() / np.abs(ref).max()))
