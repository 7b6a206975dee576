"""CPU prototype (numpy) of the round-2 idea in DESIGN.md 7(1): contract the dense hot-gene panel on the INTEGER tensor
cores.  The panel's A operand is one-hot in the count, T[c,:] = sum_k L_c(k) * (M_k . Ys)[c,:] with M_k = [D == k] a 0/1
matrix, so Ys (row-scaled Y rows of the hot genes) can be cut per column into S signed 7-bit digits of a fixed-point
number and every M_k . digit_s is an exact int8 x int8 -> int32 product (what tcgen05.mma.kind::i8 computes).
This script measures, against an 80-bit long-double reference, the error of (a) the f64 contraction the DMMA kernel does
now and (b) the sliced integer scheme, for S = 6, 7, 8 digits and for "counts 1..KMAX on the tensor path, the rest exact".
No GPU: it is an error-model check, not a kernel.  usage: python scripts/proto_int8_panel.py [n_cells]"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
from oracle import oracle as orc
from scan_rs_b200.synth import SynthConfig, generate_host

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
m, gd, w = 33538, 2048, 20
cfg = SynthConfig(n_cells=n, n_genes=m, seed=3)
ip, g, c = generate_host(cfg)
cm = orc.CountMatrix.from_cell_major(m, n, ip, g, c)
a = orc.normalize(cm, orc.CELLRANGER)
cs, rs = a.mat.spec.col_scale, a.mat.spec.row_scale
cell = np.repeat(np.arange(n), np.diff(ip).astype(np.int64))
nnz_g = np.bincount(g, minlength=m)
hot = np.sort(np.argsort(-nnz_g, kind="stable")[:gd])
col_of = np.full(m, -1); col_of[hot] = np.arange(gd)
sel = (col_of[g] >= 0) & (c <= 15)
D = np.zeros((n, gd), dtype=np.uint8)
D[cell[sel], col_of[g[sel]]] = c[sel]
print(f"panel {n} x {gd}: density {np.count_nonzero(D) / D.size:.3f}; count histogram 1..4 "
      f"{[round(float((D == k).sum()) / np.count_nonzero(D), 3) for k in (1, 2, 3, 4)]}")

rng = np.random.default_rng(0)
Y, _ = np.linalg.qr(rng.standard_normal((m, w)))          # an orthonormal Krylov block
Ys = rs[hot, None] * Y[hot]                                # what k_dense_t stages
L = np.log2(cs[:, None] * np.arange(16)[None, :] + 1.0)    # per-cell value table, L[:, 0] = 0

ld = np.longdouble
Lval = np.take_along_axis(L, D.astype(np.int64), axis=1)   # n x gd
truth = (Lval.astype(ld) @ Ys.astype(ld))                  # 80-bit accumulation
scale = np.abs(Lval) @ np.abs(Ys)                          # sum |terms|: the natural error unit of a dot product
eps = np.finfo(np.float64).eps

def report(name, T):
    err = np.abs(T.astype(ld) - truth).astype(np.float64)
    print(f"{name:44s} max err / (eps*sum|terms|) = {np.max(err / (eps * scale)):9.3f}   max rel to |T| = {np.max(err / np.abs(truth).astype(np.float64).max()):.2e}")

report("f64 contraction (numpy dot, as the DMMA kernel)", Lval @ Ys)

def digits(Ys, S):
    """per column: fixed point with 7*S bits below the column's power-of-two bound, balanced base-128 digits in [-64, 63]"""
    e = np.ceil(np.log2(np.abs(Ys).max(axis=0)))             # |Ys[:, j]| <= 2^e_j
    q = np.rint(Ys * 2.0 ** (7 * S - 1 - e)[None, :]).astype(object)   # exact big integers
    digs = []
    for s in range(S):
        d = ((q + 64) % 128) - 64
        digs.append(np.array(d, dtype=np.int64).astype(np.int8))
        q = (q - d) // 128
    assert all(int(x) == 0 for x in np.ravel(q)[:1000])
    return digs, e

for S in (6, 7, 8):
    digs, e = digits(Ys, S)
    for kmax in (3, 15):
        T = np.zeros((n, w))
        for k in range(1, kmax + 1):
            Mk = (D == k).astype(np.int8)
            acc = np.zeros((n, w))
            for s in range(S):
                p = Mk.astype(np.int32) @ digs[s].astype(np.int32)        # exact: |sum| <= 2048 * 64 < 2^31
                assert np.abs(p).max() < 2 ** 31
                acc += p.astype(np.float64) * 2.0 ** (7 * s)              # int32 * power of two: exact; the adds round
            T += L[:, k][:, None] * (acc * 2.0 ** (e - (7 * S - 1))[None, :])
        if kmax < 15:  # the remaining counts stay on an exact f64 path (they join the sparse side)
            rest = np.where(D > kmax, Lval, 0.0)
            T += rest @ Ys
            frac = float((D > kmax).sum()) / np.count_nonzero(D)
        else:
            frac = 0.0
        report(f"int8 slices S={S}, counts 1..{kmax} (rest {frac:.1%} in f64)", T)

# ---------------------------------------------------------------- N side: P[hot_j,:] = sum_c L_c(D[c,j]) * X[c,:]
# The contraction runs over the cells, so the digit planes are those of the B operand L_c(k) * X[c,:] (one fixed-point
# scale per column and count level, taken over all cells) and the 0/1 operand is [D == k]^T.  X is a realistic block:
# A^T . Y of the normalized matrix (what the Krylov loop feeds to A . X).
print("\nN side (contraction over cells)")
X = a.rdot(Y.T[:w]).T if hasattr(a, "rdot") else None      # n x w
Lval_T = Lval.T                                            # gd x n
truth_n = Lval_T.astype(ld) @ X.astype(ld)
scale_n = np.abs(Lval_T) @ np.abs(X)

def report_n(name, P):
    err = np.abs(P.astype(ld) - truth_n).astype(np.float64)
    print(f"{name:44s} max err / (eps*sum|terms|) = {np.max(err / (eps * scale_n)):9.3f}   max rel to |P| = {np.max(err / np.abs(truth_n).astype(np.float64).max()):.2e}")

report_n("f64 contraction (numpy dot, as the DMMA kernel)", Lval_T @ X)
for S in (7, 8):
    P = np.zeros((gd, w))
    for k in range(1, 4):
        Bk = L[:, k][:, None] * X                          # n x w, the level's B operand in f64
        digs, e = digits(Bk, S)
        Mk = (D == k).astype(np.int32).T                   # gd x n
        acc = np.zeros((gd, w))
        for s in range(S):
            p = Mk @ digs[s].astype(np.int32)
            assert np.abs(p).max() < 2 ** 31
            acc += p.astype(np.float64) * 2.0 ** (7 * s)
        P += acc * 2.0 ** (e - (7 * S - 1))[None, :]
    rest = np.where(D > 3, Lval, 0.0).T
    P += rest @ X
    report_n(f"int8 slices S={S}, counts 1..3 (rest in f64)", P)
