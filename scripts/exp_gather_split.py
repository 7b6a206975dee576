"""Round-2 starting point: A/B of the experimental split gather streams (csrc/gather_split.cu, option gather_split) against
the combined-stream gather.  usage: python scripts/exp_gather_split.py [n_cells]"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import scan_rs_b200 as sb
from scan_rs_b200.synth import SynthConfig, generate_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
ctx = sb.Context(0)
dm = generate_device(ctx, SynthConfig(n_cells=n, n_genes=33538, seed=3))
a = sb.normalize(dm, sb.Normalization.CellRanger)
rng = np.random.default_rng(0)
x = rng.standard_normal((n, 20))
y = rng.standard_normal((20, 33538))
res = {}
for on in (0, 1, 0, 1):
    ctx.set_option("gather_split", on)
    pn, pt = a.dot(x), a.rdot(y)  # warm (first call with the option on builds the split streams)
    ctx.profile_enable(True); ctx.profile_reset()
    for _ in range(5):
        a.dot(x); a.rdot(y)
    p = ctx.profile(); ctx.profile_enable(False)
    res.setdefault(on, (pn, pt))
    print(f"gather_split={on}: spmm_t {p['spmm_t_ms'] / p['spmm_t_launches']:.3f} ms/pass, spmm_n {p['spmm_n_ms'] / p['spmm_n_launches']:.3f} ms/pass", flush=True)
for name, i in (("A.X", 0), ("A^T.Y", 1)):
    print(name, "max rel diff split vs combined:", float(np.abs(res[0][i] - res[1][i]).max() / np.abs(res[0][i]).max()))
