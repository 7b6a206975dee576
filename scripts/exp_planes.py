"""A/B of the dense half of the hybrid layout: panel_mode 0 (pure sparse), 1 (u8 panel, FP64 mma.sync), 2 (bit planes, int8 tcgen05).
Small case first: products and moments of every mode against the oracle; then event-timed passes at n_cells.
usage: python scripts/exp_planes.py [n_cells] [modes e.g. 012]"""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import scan_rs_b200 as sb
from oracle import oracle as orc
from scan_rs_b200.synth import SynthConfig, generate_device, generate_host

n_big = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
modes = [int(c) for c in (sys.argv[2] if len(sys.argv) > 2 else "012")]
ctx = sb.Context(0)

def small(n_cells, n_genes, seed, **kw):
    cfg = SynthConfig(n_cells=n_cells, n_genes=n_genes, seed=seed, **kw)
    ip, g, c = generate_host(cfg)
    cm = orc.CountMatrix.from_cell_major(n_genes, n_cells, ip, g, c)
    a_o = orc.normalize(cm, orc.CELLRANGER)
    rng = np.random.default_rng(5)
    for mode in modes:
        ctx.set_option("panel_mode", mode)
        dm = sb.AdaptiveMat.from_csc(ctx, n_genes, n_cells, ip, g, c)
        a = sb.normalize(dm, sb.Normalization.CellRanger)
        errs = []
        for p_g, p_o in zip(a.params(), (a_o.mat.spec.col_scale, a_o.mat.spec.row_scale, a_o.u.ravel(), a_o.v.ravel())):
            errs.append(float(np.abs(p_g - p_o).max() / max(1e-300, np.abs(p_o).max())))
        out = [f"mode {mode} {n_cells}x{n_genes}: params {max(errs):.1e}"]
        for w in (20, 7, 45):
            x = rng.standard_normal((n_cells, w)); y = rng.standard_normal((w, n_genes))
            rn, rt = a_o.dot(x), a_o.rdot(y)
            en = float(np.abs(a.dot(x) - rn).max() / np.abs(rn).max())
            et = float(np.abs(a.rdot(y) - rt).max() / np.abs(rt).max())
            out.append(f"w={w}: N {en:.1e} T {et:.1e}")
        print("  ".join(out), flush=True)
        a.free(); dm.free()

small(3000, 2500, 37, n_dense=40, dense_mean=40.0)
small(6000, 33538, 44)
small(1000, 200, 3)

if n_big > 0:
    x = np.random.default_rng(0).standard_normal((n_big, 20))
    y = np.random.default_rng(1).standard_normal((20, 33538))
    ref = None
    for mode in modes:
        ctx.set_option("panel_mode", mode)
        t0 = time.perf_counter()
        dm = generate_device(ctx, SynthConfig(n_cells=n_big, n_genes=33538, seed=3))
        a = sb.normalize(dm, sb.Normalization.CellRanger)
        ctx.sync()
        t1 = time.perf_counter()
        pn, pt = a.dot(x), a.rdot(y)
        ctx.profile_enable(True); ctx.profile_reset()
        for _ in range(5):
            a.dot(x); a.rdot(y)
        p = ctx.profile(); ctx.profile_enable(False)
        print(f"mode {mode} n={n_big}: build+normalize {1e3 * (t1 - t0):.0f} ms  spmm_t {p['spmm_t_ms'] / p['spmm_t_launches']:.3f} ms/pass  "
              f"spmm_n {p['spmm_n_ms'] / p['spmm_n_launches']:.3f} ms/pass", flush=True)
        if ref is None:
            ref = (pn, pt)
        else:
            print(f"   vs first mode: N {np.abs(pn - ref[0]).max() / np.abs(ref[0]).max():.1e}  T {np.abs(pt - ref[1]).max() / np.abs(ref[1]).max():.1e}", flush=True)
        a.free(); dm.free()
