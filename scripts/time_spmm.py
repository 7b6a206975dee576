"""Times the two products (device events through the library profile) with the overlap option on and off."""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
ctx = sb.Context(0)
dm = generate_device(ctx, SynthConfig(n_cells=n, n_genes=33538, seed=3))
a = sb.normalize(dm, sb.Normalization.CellRanger)
out = sb.pinned_outputs(33538, n, 10)
ref = None
for ov in (1, 0, 1):
    ctx.set_option("overlap", ov)
    sb.BkSvd().run_pca(a, 10, out=out)   # warm
    ctx.profile_enable(True); ctx.profile_reset()
    ctx.timer_begin()
    u, s, v = sb.BkSvd().run_pca(a, 10, out=out)
    ms = ctx.timer_end()
    p = ctx.profile(); ctx.profile_enable(False)
    if ref is None: ref = s.copy()
    print(f"overlap={ov}: pca {ms:.1f} ms, spmm_t {p['spmm_t_ms']/p['spmm_t_launches']:.2f} ms/pass, spmm_n {p['spmm_n_ms']/p['spmm_n_launches']:.2f} ms/pass, "
          f"sigma rel diff vs first {np.abs(s-ref).max()/ref.max():.1e}", flush=True)
