"""CPU oracle for the scan-rs normalize -> PCA path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (scan_rs_b200/) never does and fails
loudly when its CUDA library is missing.

What it restates (paths relative to the reference checkout; every function cites lines):
  * the sparse loops and map chain in C (oracle/scan_oracle.c, -ffp-contract=off),
  * the drivers (normalize variants, scale_and_center, LowRankOffset products, svd_bk,
    svd_rand, partition, select) in numpy here,
  * dense QR / SVD through SciPy's LAPACK (OpenBLAS): the reference's MKL dgeqrf/dorgqr/
    dgesdd are third-party (ndarray-linalg 0.17.0 -> lax 0.17.0 -> intel-mkl-src 0.8.1,
    not vendored); agreement is at rounding level, up to column signs,
  * the start block: rand 0.10.1 SmallRng (Xoshiro256++ seeded through SplitMix64) +
    Uniform::<f64>::new(-1, 1).  The stream model cannot be checked against a Rust build
    offline, so for the exact Omega values: PARITY UNPINNED.  Everything else
    is pinned by the reference's inline golden vectors (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Callable, Optional, Sequence

import numpy as np
import scipy.linalg as sla

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile scan_oracle.c -> oracle/_build/liboracle.so (gcc, no fast-math, no FMA)."""
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "liboracle.so")
    src = os.path.join(_HERE, "scan_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
               "-o", so, src, "-lm"]
        subprocess.check_call(cmd)
    return so


class _Map(C.Structure):
    _fields_ = [("kind", C.c_int), ("log_base", C.c_int), ("square", C.c_int),
                ("col_scale", C.c_void_p), ("row_scale", C.c_void_p),
                ("bn", C.c_void_p), ("bpi", C.c_void_p)]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_median_u32.restype = C.c_int
        _LIB.orc_partition_masks.restype = C.c_int
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def set_num_threads(n: int) -> int:
    """Pin the OpenMP thread count of the threaded products (independent of OMP_NUM_THREADS); returns the count in effect."""
    lib().orc_set_num_threads(C.c_int(int(n)))
    return int(lib().orc_num_threads())


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------------------
# RNG: rand 0.10.1 SmallRng on 64-bit = Xoshiro256++; seed_from_u64 = SplitMix64 fill.
# Known answers: state [1,2,3,4] -> 41943041, 58720359, 3588806011781223;
# seed_from_u64(0) -> 5987356902031041503, 7051070477665621255, ...
# --------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


class Xoshiro256PlusPlus:
    def __init__(self, state: Sequence[int]):
        self.s = [int(x) & _M64 for x in state]

    @classmethod
    def seed_from_u64(cls, seed: int) -> "Xoshiro256PlusPlus":
        st, z = [], int(seed) & _M64
        for _ in range(4):
            z = (z + 0x9E3779B97F4A7C15) & _M64
            x = z
            x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
            x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
            st.append(x ^ (x >> 31))
        return cls(st)

    def next_u64(self) -> int:
        s = self.s
        r = (s[0] + s[3]) & _M64
        r = (((r << 23) | (r >> 41)) & _M64) + s[0] & _M64
        t = (s[1] << 17) & _M64
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = ((s[3] << 45) | (s[3] >> 19)) & _M64
        return r

    def fill_u64(self, count: int) -> np.ndarray:
        """Vectorised-enough bulk draw (pure Python loop; fine up to a few 1e6 draws)."""
        out = np.empty(count, dtype=np.uint64)
        nxt = self.next_u64
        for i in range(count):
            out[i] = nxt()
        return out


def uniform_m1_1(rng: Xoshiro256PlusPlus, shape) -> np.ndarray:
    """Uniform::new(-1.0, 1.0) sampled in row-major fill order (bk_svd.rs:83-84, :90, :118):
    value1_2 = bits(u64 >> 12 | exponent 0) in [1,2); (value1_2 - 1.0) * 2.0 + (-1.0)."""
    count = int(np.prod(shape))
    u = rng.fill_u64(count)
    bits = (u >> np.uint64(12)) | np.uint64(0x3FF0000000000000)
    v12 = bits.view(np.float64)
    return ((v12 - 1.0) * 2.0 + (-1.0)).reshape(shape)


def omega(seed: int, shape) -> np.ndarray:
    return uniform_m1_1(Xoshiro256PlusPlus.seed_from_u64(seed), shape)


# --------------------------------------------------------------------------------------
# Matrix types
# --------------------------------------------------------------------------------------
class CountMatrix:
    """genes(rows) x cells(cols) u32 counts, gene-major CSR: the lossless content of an
    AdaptiveMat<u32> in CSR storage (sqz/src/mat.rs:34-42, from_csmat :92-124)."""

    def __init__(self, rows: int, cols: int, indptr, idx, val):
        self.rows, self.cols = int(rows), int(cols)
        self.indptr = np.ascontiguousarray(indptr, dtype=np.uint64)
        self.idx = np.ascontiguousarray(idx, dtype=np.uint32)
        self.val = np.ascontiguousarray(val, dtype=np.uint32)
        assert self.indptr.shape == (self.rows + 1,)
        assert self.idx.shape == self.val.shape == (int(self.indptr[-1]),)

    @classmethod
    def from_dense(cls, dense) -> "CountMatrix":  # mat.rs:586-609
        dense = np.asarray(dense)
        rows, cols = dense.shape
        indptr, idx, val = [0], [], []
        for r in range(rows):
            nz = np.nonzero(dense[r])[0]
            idx.extend(nz.tolist())
            val.extend(dense[r, nz].tolist())
            indptr.append(len(idx))
        return cls(rows, cols, indptr, np.array(idx, dtype=np.uint32), np.array(val, dtype=np.uint32))

    @classmethod
    def from_cell_major(cls, m: int, n: int, indptr, gene, count) -> "CountMatrix":
        """Build the gene-major CSR from cell-major arrays (stable, ascending cells)."""
        import scipy.sparse as sp
        csc = sp.csr_matrix((np.asarray(count), np.asarray(gene).astype(np.int64),
                             np.asarray(indptr).astype(np.int64)), shape=(n, m))
        t = csc.T.tocsr()  # genes x cells
        t.sort_indices()
        return cls(m, n, t.indptr, t.indices, t.data)

    def cell_major(self):
        """(indptr u64[n+1], gene u32, count u32) with ascending genes in each cell."""
        import scipy.sparse as sp
        a = sp.csr_matrix((self.val, self.idx.astype(np.int64), self.indptr.astype(np.int64)),
                          shape=(self.rows, self.cols))
        t = a.T.tocsr()
        t.sort_indices()
        return t.indptr.astype(np.uint64), t.indices.astype(np.uint32), t.data.astype(np.uint32)

    @property
    def shape(self):
        return [self.rows, self.cols]

    @property
    def nnz(self) -> int:
        return int(self.idx.shape[0])

    def _args(self):
        return (C.c_uint64(self.rows), C.c_uint64(self.cols), _p(self.indptr), _p(self.idx), _p(self.val))

    def to_dense(self) -> np.ndarray:
        out = np.zeros((self.rows, self.cols), dtype=np.uint32)
        for r in range(self.rows):
            s, e = int(self.indptr[r]), int(self.indptr[r + 1])
            out[r, self.idx[s:e]] = self.val[s:e]
        return out

    def sum_axis_u32(self, axis: int) -> np.ndarray:  # mat.rs:377-406
        out = np.zeros(self.cols if axis == 0 else self.rows, dtype=np.uint32)
        lib().orc_sum_axis_u32(*self._args(), C.c_int(axis), _p(out))
        return out

    def sum_axis_u64(self, axis: int, square: bool = False) -> np.ndarray:
        out = np.zeros(self.cols if axis == 0 else self.rows, dtype=np.uint64)
        lib().orc_sum_axis_u64(*self._args(), C.c_int(axis), C.c_int(int(square)), _p(out))
        return out

    def select_rows(self, rows: Sequence[int]) -> "CountMatrix":  # mat.rs:1040-1071 (CSR branch)
        indptr, idx, val = [0], [], []
        for r in rows:
            s, e = int(self.indptr[r]), int(self.indptr[r + 1])
            idx.append(self.idx[s:e])
            val.append(self.val[s:e])
            indptr.append(indptr[-1] + (e - s))
        cat = (lambda xs: np.concatenate(xs) if xs else np.zeros(0, dtype=np.uint32))
        return CountMatrix(len(rows), self.cols, indptr, cat(idx), cat(val))

    def select_cols(self, cols: Sequence[int]) -> "CountMatrix":  # mat.rs:1004-1037 (CSR branch)
        cols = np.asarray(cols, dtype=np.int64)
        indptr, idx, val = [0], [], []
        for r in range(self.rows):
            s, e = int(self.indptr[r]), int(self.indptr[r + 1])
            tmp = np.zeros(self.cols, dtype=np.uint32)
            tmp[self.idx[s:e]] = self.val[s:e]
            picked = tmp[cols] if len(cols) else np.zeros(0, dtype=np.uint32)
            nz = np.nonzero(picked)[0]
            idx.append(nz.astype(np.uint32))
            val.append(picked[nz])
            indptr.append(indptr[-1] + len(nz))
        cat = (lambda xs: np.concatenate(xs) if xs else np.zeros(0, dtype=np.uint32))
        return CountMatrix(self.rows, len(cols), indptr, cat(idx), cat(val))

    def partition_on_thresholds(self, row_threshold: Optional[float], col_threshold: Optional[float]):
        """mat.rs:772-889.  Returns (filtered, residual, selected_rows, selected_cols)."""
        er = np.zeros(self.rows, dtype=np.uint8)
        ec = np.zeros(self.cols, dtype=np.uint8)
        mp = _Map(0, 0, 0, None, None, None, None)
        lib().orc_partition_masks(*self._args(), C.byref(mp),
                                  C.c_int(row_threshold is not None), C.c_double(row_threshold or 0.0),
                                  C.c_int(col_threshold is not None), C.c_double(col_threshold or 0.0),
                                  _p(er), _p(ec))
        sel_rows = np.nonzero(er == 0)[0]
        sel_cols = np.nonzero(ec == 0)[0]
        exc_cols = np.nonzero(ec != 0)[0]
        kept = self.select_rows(sel_rows.tolist())
        filtered = kept.select_cols(sel_cols)
        residual = kept.select_cols(exc_cols)
        return filtered, residual, sel_rows, sel_cols

    def partition_on_threshold(self, threshold: float):  # mat.rs:766-768
        return self.partition_on_thresholds(threshold, threshold)


@dataclass
class MapSpec:
    """The composed MatrixMap (sqz/src/matrix_map.rs:145-197) the normalize path builds."""
    kind: int = 0
    log_base: int = 0
    col_scale: Optional[np.ndarray] = None
    row_scale: Optional[np.ndarray] = None
    bn: Optional[np.ndarray] = None
    bpi: Optional[np.ndarray] = None

    def c(self, square: bool = False) -> _Map:
        return _Map(self.kind, self.log_base, int(square), _p(self.col_scale), _p(self.row_scale),
                    _p(self.bn), _p(self.bpi))


class MappedMatrix:
    """AdaptiveMat<f64, D, impl MatrixMap<u32, f64>>: counts + lazily applied map."""

    def __init__(self, counts: CountMatrix, spec: MapSpec):
        self.counts, self.spec = counts, spec

    rows = property(lambda self: self.counts.rows)
    cols = property(lambda self: self.counts.cols)
    shape = property(lambda self: [self.counts.rows, self.counts.cols])

    def sum_axis(self, axis: int, square: bool = False) -> np.ndarray:  # mat.rs:377-406
        out = np.zeros(self.cols if axis == 0 else self.rows, dtype=np.float64)
        mp = self.spec.c(square)
        lib().orc_sum_axis_map(*self.counts._args(), C.byref(mp), C.c_int(axis), _p(out))
        return out

    def mean_axis(self, axis: int, square: bool = False) -> np.ndarray:  # mat.rs:273-276
        m = float(self.shape[axis])
        return self.sum_axis(axis, square) / m

    def dot(self, rhs: np.ndarray, threads: bool = False) -> np.ndarray:
        """A . rhs (mat.rs:1074-1090 -> prod.rs:30-51, 124-148)."""
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        assert rhs.shape[0] == self.cols
        out = np.zeros((self.rows, rhs.shape[1]), dtype=np.float64)
        mp = self.spec.c()
        fn = lib().orc_spmm_gather_omp if threads else lib().orc_spmm_gather
        fn(*self.counts._args(), C.byref(mp), _p(rhs), C.c_uint64(rhs.shape[1]), _p(out))
        return out

    def rdot(self, lhs: np.ndarray, threads: bool = False) -> np.ndarray:
        """lhs . A for lhs (w x rows) (mat.rs:1114-1133 -> prod.rs:56-81, 190-214)."""
        y = np.ascontiguousarray(np.asarray(lhs, dtype=np.float64).T)  # rows x w, y[g,:] = lhs[:,g]
        assert y.shape[0] == self.rows
        out = np.zeros((self.cols, y.shape[1]), dtype=np.float64)
        mp = self.spec.c()
        fn = lib().orc_spmm_scatter_omp if threads else lib().orc_spmm_scatter
        fn(*self.counts._args(), C.byref(mp), _p(y), C.c_uint64(y.shape[1]), _p(out))
        return out.T  # out.reversed_axes() (mat.rs:1131)

    def to_dense(self) -> np.ndarray:  # mat.rs:190-204
        out = np.zeros((self.rows, self.cols), dtype=np.float64)
        mp = self.spec.c()
        lib().orc_to_dense(*self.counts._args(), C.byref(mp), _p(out))
        return out

    # scale / center / scale_and_center (mat.rs:937-1001), Axis(1) = per gene row only.
    def scale_and_center(self, scaling_factors: Optional[np.ndarray] = None) -> "LowRankOffset":
        means = self.mean_axis(1)                                               # :993
        if scaling_factors is None:
            matsq_means = self.mean_axis(1, square=True)                        # :995
            x = matsq_means - means * means                                     # :996  (powi(2) = x*x)
            scaling_factors = np.where(x <= 0.0, 1.0, np.sqrt(np.where(x <= 0.0, 1.0, x)))
        else:
            scaling_factors = np.asarray(scaling_factors, dtype=np.float64)
        means = means / scaling_factors                                         # :998
        inv = 1.0 / scaling_factors                                             # :970
        if self.spec.row_scale is not None:
            raise ValueError("row scale already present")
        spec = MapSpec(self.spec.kind, self.spec.log_base, self.spec.col_scale, np.ascontiguousarray(inv),
                       self.spec.bn, self.spec.bpi)
        neg_means = -means                                                      # :946
        u = neg_means.reshape(self.rows, 1)                                     # :955
        v = np.ones((1, self.cols))                                             # :957
        return LowRankOffset(MappedMatrix(self.counts, spec), u, v)


class LowRankOffset:
    """A = mat + u.v (sqz/src/low_rank_offset.rs:12-16)."""

    def __init__(self, mat: MappedMatrix, u: np.ndarray, v: np.ndarray):
        assert mat.rows == u.shape[0] and mat.cols == v.shape[1] and u.shape[1] == v.shape[0]
        self.mat, self.u, self.v = mat, u, v

    shape = property(lambda self: [self.mat.rows, self.mat.cols])

    def to_dense(self) -> np.ndarray:  # :55-57
        return self.u.dot(self.v) + self.mat.to_dense()

    def dot(self, rhs: np.ndarray, threads: bool = False) -> np.ndarray:  # :76-80
        res = self.mat.dot(rhs, threads)
        res += self.u.dot(self.v.dot(rhs))
        return res

    def rdot(self, lhs: np.ndarray, threads: bool = False) -> np.ndarray:  # :91-95
        res = self.mat.rdot(lhs, threads)
        res = res + lhs.dot(self.u).dot(self.v)
        return res


class DenseOp:
    """Array2<f64> as the PCA operand (dim_red/mod.rs:61-65): used by the dim_red tests."""

    def __init__(self, a: np.ndarray):
        self.a = np.asarray(a, dtype=np.float64)

    shape = property(lambda self: list(self.a.shape))

    def dot(self, rhs, threads=False):
        return self.a.dot(rhs)

    def rdot(self, lhs, threads=False):
        return lhs.dot(self.a)

    def to_dense(self):
        return self.a


# --------------------------------------------------------------------------------------
# normalization.rs
# --------------------------------------------------------------------------------------
CELLRANGER, CELLRANGER8, SEURATLOG, BINOMIAL_DEVIANCE, BINOMIAL_PEARSON, WITH_SIZE_FACTORS, LOG_TRANSFORM = range(7)
_NORM_NAMES = {"cellranger": CELLRANGER, "cellranger8": CELLRANGER8, "seuratlog": SEURATLOG,
               "binomialdeviance": BINOMIAL_DEVIANCE, "binomialpearson": BINOMIAL_PEARSON}
LOG_E, LOG_TWO, LOG_TEN = 1, 2, 10


def normalization_from_str(s: str) -> int:  # normalization.rs:30-43
    if s not in _NORM_NAMES:
        raise ValueError(f"Normalization not recognized: {s}")
    return _NORM_NAMES[s]


def median_mut(xs: np.ndarray):  # stats.rs:13-38
    xs = np.ascontiguousarray(xs, dtype=np.uint32)
    out = C.c_uint32(0)
    ok = lib().orc_median_u32(_p(xs), C.c_uint64(xs.shape[0]), C.byref(out))
    return int(out.value) if ok else None


def log_normalize_with_size_factor(matrix: CountMatrix, umi_count_sum: Optional[float], log_base: int,
                                   size_factors: Optional[np.ndarray]) -> MappedMatrix:
    """normalization.rs:138-178."""
    if size_factors is not None:
        size_factors = np.asarray(size_factors, dtype=np.uint32)
        assert size_factors.shape[0] == matrix.cols, "Size of the size factor and matrix columns dont match."
        normalization_counts = size_factors
    else:
        normalization_counts = matrix.sum_axis_u32(0)                 # :159
    umi_counts = matrix.sum_axis_u32(0)                               # :161
    if umi_count_sum is not None:
        target = float(umi_count_sum)
    else:
        med = median_mut(umi_counts.copy())                           # :166
        target = 1.0 if med is None else max(float(med), 1.0)
    with np.errstate(divide="ignore"):
        col_scales = target / normalization_counts.astype(np.float64)  # :169
    return MappedMatrix(matrix, MapSpec(kind=1, log_base=log_base, col_scale=np.ascontiguousarray(col_scales)))


def log_normalize(matrix, umi_count_sum, log_base):  # normalization.rs:119-129
    return log_normalize_with_size_factor(matrix, umi_count_sum, log_base, None)


def normalize_with_size_factor(mat: CountMatrix, norm: int, size_factors=None) -> LowRankOffset:
    """normalization.rs:72-102."""
    if norm == CELLRANGER:
        return log_normalize_with_size_factor(mat, None, LOG_TWO, None).scale_and_center(None)
    if norm == CELLRANGER8:
        return log_normalize_with_size_factor(mat, None, LOG_TWO, None).scale_and_center(np.ones(mat.rows))
    if norm == SEURATLOG:
        return log_normalize_with_size_factor(mat, 10000.0, LOG_E, None).scale_and_center(None)
    if norm == WITH_SIZE_FACTORS:
        return log_normalize_with_size_factor(mat, None, LOG_TWO, size_factors).scale_and_center(None)
    if norm == LOG_TRANSFORM:
        ones = np.ones(mat.cols, dtype=np.uint32)
        return log_normalize_with_size_factor(mat, 1.0, LOG_TWO, ones).scale_and_center(None)
    raise NotImplementedError("not implemented")  # :100 panic!("not implemented")


def normalize(mat: CountMatrix, norm: int) -> LowRankOffset:
    """normalization.rs:46-69 (+ the CLI's dispatch of the binomial kinds, tools/src/bin/cmd.rs:67-81)."""
    if norm in (CELLRANGER, CELLRANGER8, SEURATLOG):
        return normalize_with_size_factor(mat, norm, None)
    if norm == BINOMIAL_DEVIANCE:
        return binom_deviance_resid(mat)
    if norm == BINOMIAL_PEARSON:
        return binom_pearson_resid(mat)
    raise NotImplementedError("not implemented")


def log1p_normalize_fixed_point(matrix: CountMatrix, log_base: int, base: int, exponent: int) -> LowRankOffset:
    """normalization.rs:191-213."""
    denom = float(np.uint32(base) ** np.uint32(exponent))
    col_scales = np.ones(matrix.cols) / denom
    mm = MappedMatrix(matrix, MapSpec(kind=1, log_base=log_base, col_scale=np.ascontiguousarray(col_scales)))
    return mm.scale_and_center(None)


def fit_multinomial_model(matrix: CountMatrix):  # normalization.rs:218-228
    ident = MappedMatrix(matrix, MapSpec(kind=0))
    n = ident.sum_axis(0)
    total = float(np.sum(n))  # ndarray .sum(): pairwise-ish; exact for integer-valued f64 below 2^53
    pi = ident.sum_axis(1) / total
    return n, pi


def binom_deviance_resid(matrix: CountMatrix) -> LowRankOffset:  # normalization.rs:233-260
    n, pi = fit_multinomial_model(matrix)
    with np.errstate(divide="ignore", invalid="ignore"):
        u = np.sqrt(np.log(1.0 / (1.0 - pi))).reshape(matrix.rows, 1)
        v = (-np.sqrt(2.0 * n)).reshape(1, matrix.cols)
    spec = MapSpec(kind=2, bn=np.ascontiguousarray(n), bpi=np.ascontiguousarray(pi))
    return LowRankOffset(MappedMatrix(matrix, spec), u, v)


def binom_pearson_resid(matrix: CountMatrix) -> LowRankOffset:  # normalization.rs:307-323
    n, pi = fit_multinomial_model(matrix)
    with np.errstate(divide="ignore", invalid="ignore"):
        u = np.sqrt(pi / (1.0 - pi)).reshape(matrix.rows, 1)
        v = (-np.sqrt(n)).reshape(1, matrix.cols)
    spec = MapSpec(kind=3, bn=np.ascontiguousarray(n), bpi=np.ascontiguousarray(pi))
    return LowRankOffset(MappedMatrix(matrix, spec), u, v)


# --------------------------------------------------------------------------------------
# HVG selection: NOT in the reference (builder-defined, parity unpinned):
# exact u64 S1 = sum v, S2 = sum v^2 per gene; dispersion = var/mean in f64; top-N, ties
# -> lower gene index; the result is sorted ascending so select_rows keeps gene order.
# --------------------------------------------------------------------------------------
def hvg_dispersion(s1: np.ndarray, s2: np.ndarray, n_cells: int) -> np.ndarray:
    s1 = s1.astype(np.float64)
    s2 = s2.astype(np.float64)
    mean = s1 / float(n_cells)
    var = s2 / float(n_cells) - mean * mean
    with np.errstate(divide="ignore", invalid="ignore"):
        disp = np.where(mean > 0.0, var / mean, 0.0)
    return disp


def hvg_select(s1: np.ndarray, s2: np.ndarray, n_cells: int, n_top: int) -> np.ndarray:
    disp = hvg_dispersion(s1, s2, n_cells)
    order = np.lexsort((np.arange(disp.shape[0]), -disp))  # primary: -disp, ties: lower index
    return np.sort(order[: min(n_top, disp.shape[0])]).astype(np.uint32)


# --------------------------------------------------------------------------------------
# dim_red
# --------------------------------------------------------------------------------------
class CancellationError(Exception):  # snoop/src/lib.rs:5-18
    pass


def _qr_q(a: np.ndarray) -> np.ndarray:  # ndarray-linalg QR::qr -> dgeqrf + dorgqr
    q, _ = sla.qr(a, mode="economic", check_finite=False)
    return q


def _svddc(a: np.ndarray):  # svddc_into(JobSvd::Some) -> dgesdd jobz='S'
    return sla.svd(a, full_matrices=False, lapack_driver="gesdd", check_finite=False)


def svd_bk(A, k: int, b: int, n_iter: int, seed: int = 0,
           snoop: Optional[Callable[[float], bool]] = None, threads: bool = False,
           omega_block: Optional[np.ndarray] = None):
    """scan-rs/src/dim_red/bk_svd.rs:57-146.  Returns (U m x k, sigma k, Va k x n).
    `snoop(fraction)` returning True cancels (set_progress_check, snoop/src/lib.rs:45-57)."""
    m, n = A.shape
    if m < 2 or n < 2:
        raise ValueError("The input matrix must be at least 2x2.")  # :73-75
    if k > min(m, n):
        raise ValueError("invalid k")                               # :77-79
    b = min(min(m, n), b)                                           # :81

    def check(frac):
        if snoop is not None and snoop(frac):
            raise CancellationError()

    if m >= n:
        B = omega(seed, (n, b)) if omega_block is None else np.array(omega_block, dtype=np.float64)  # :90
        K = np.zeros((n, b * n_iter))
        for i in range(n_iter):
            B = _qr_q(A.rdot(A.dot(B, threads).T, threads).T)       # :94
            K[:, i * b:(i + 1) * b] = B
            check(i / n_iter * 0.8)
        Q = _qr_q(K)                                                # :98
        check(0.82)
        T = A.dot(Q, threads)                                       # :102
        check(0.93)
        Ut, s, Vt = _svddc(T)                                       # :105
        U, sigma, Va = Ut[:, :k].copy(), s[:k].copy(), Vt[:k, :].copy()
        Va = Va.dot(Q.T)                                            # :113
        check(1.0)
        return U, sigma, Va
    else:
        B = omega(seed, (b, m)) if omega_block is None else np.array(omega_block, dtype=np.float64)  # :118
        K = np.zeros((b * n_iter, m))
        for i in range(n_iter):
            T = A.rdot(B, threads).T                                # :122
            B = _qr_q(A.dot(T, threads)).T                          # :123
            K[i * b:(i + 1) * b, :] = B
            check(i / n_iter * 0.8)
        Q = _qr_q(K.T)                                              # :127
        check(0.82)
        T = A.rdot(Q.T, threads)                                    # :131
        check(0.93)
        Ut, s, Vt = _svddc(T)                                       # :134
        U, sigma, Va = Ut[:, :k].copy(), s[:k].copy(), Vt[:k, :].copy()
        U = Q.dot(U)                                                # :142
        check(1.0)
        return U, sigma, Va


def svd_rand(A, k: int, l: int, n_iter: int, seed: int = 0, threads: bool = False,
             omega_block: Optional[np.ndarray] = None):
    """scan-rs/src/dim_red/rand_svd.rs:54-129.  Returns (U, sigma, Va k x n)."""
    m, n = A.shape
    if m < 2 or n < 2:
        raise ValueError("The input matrix must be at least 2x2.")
    if k > min(m, n):
        raise ValueError("invalid k")
    if m >= n:
        om = omega(seed, (n, l)) if omega_block is None else np.asarray(omega_block, dtype=np.float64)
        Q = _qr_q(A.dot(om, threads))                               # :87
        for _ in range(n_iter):
            Q = _qr_q(A.rdot(Q.T, threads).T)                       # :90
            Q = _qr_q(A.dot(Q, threads))                            # :91
        B = A.rdot(Q.T, threads)                                    # :96
        Ut, s, Vt = _svddc(B)
        U, sigma, Va = Ut[:, :k].copy(), s[:k].copy(), Vt[:k, :].copy()
        U = Q.dot(U)                                                # :104
        return U, sigma, Va
    else:
        om = omega(seed, (l, m)) if omega_block is None else np.asarray(omega_block, dtype=np.float64)
        Q = _qr_q(A.rdot(om, threads).T)                            # :109
        for _ in range(n_iter):
            Q = _qr_q(A.dot(Q, threads))                            # :112
            Q = _qr_q(A.rdot(Q.T, threads).T)                       # :113
        B = A.dot(Q, threads)                                       # :118
        Ut, s, Vt = _svddc(B)
        U, sigma, Va = Ut[:, :k].copy(), s[:k].copy(), Vt[:k, :].copy()
        Va = Va.dot(Q.T)                                            # :126
        return U, sigma, Va


class BkSvd:  # bk_svd.rs:16-53
    def __init__(self, k_multiplier: float = 2.0, n_iter: int = 5):
        self.k_multiplier, self.n_iter = k_multiplier, n_iter

    def run_pca_cancellable(self, array, k: int, snoop=None, threads=False, omega_block=None):
        bsize = int(np.ceil(k * self.k_multiplier))                 # :49
        u, s, vt = svd_bk(array, k, bsize, self.n_iter, 0, snoop, threads, omega_block)
        return u, s, vt.T                                           # :51

    def run_pca(self, array, k: int, **kw):                         # dim_red/mod.rs:108-110
        return self.run_pca_cancellable(array, k, None, **kw)


class RandSvd:  # rand_svd.rs:13-50
    def __init__(self, l_multiplier: float = 10.0, n_iter: int = 2):
        self.l_multiplier, self.n_iter = l_multiplier, n_iter

    def run_pca(self, array, k: int, threads=False, omega_block=None):
        l = max(k + 4, int(k * self.l_multiplier))                  # :46
        u, s, vt = svd_rand(array, k, l, self.n_iter, 0, threads, omega_block)
        return u, s, vt.T


# --------------------------------------------------------------------------------------
# IRLBA: scan-rs/src/dim_red/irlba.rs:71-215 (b = 1 Lanczos bidiagonalisation with implicit restarts).
# Two things the reference takes from third-party crates cannot be restated bit for bit offline and are therefore
# PARAMETERS here (parity unpinned for them, as for the Omega stream): the start vector (rand_distr 0.6 `Normal` on
# SmallRng seed 0, a ziggurat sampler whose tables are not on disk) and the signs LAPACK's dgesvd gives the singular vectors
# of the small matrix B.  The signs matter because the reference tests `resid[i] < tol * smax` WITHOUT an absolute value
# (irlba.rs:176-181): a Ritz pair whose last u component is negative counts as converged.  Both sides of the parity tests
# therefore use the same rule: the start vector of irlba_start() and sign-canonical singular vectors (_svd_canonical).
# --------------------------------------------------------------------------------------
def irlba_start(seed: int, n: int) -> np.ndarray:
    """Builder-defined stand-in for `Normal::new(0, 1)` on SmallRng::seed_from_u64(seed) (irlba.rs:106-113): Box-Muller on
    the Xoshiro256++ stream, two uniforms per pair of normals; NOT the reference's ziggurat stream (unpinned)."""
    rng = Xoshiro256PlusPlus.seed_from_u64(seed)
    u = rng.fill_u64(2 * ((n + 1) // 2))
    f = ((u >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)  # (0, 1)
    r = np.sqrt(-2.0 * np.log(f[0::2]))
    t = 2.0 * np.pi * f[1::2]
    out = np.empty(2 * r.shape[0])
    out[0::2] = r * np.cos(t)
    out[1::2] = r * np.sin(t)
    return out[:n].copy()


def _svd_canonical(B: np.ndarray):
    """SVD::svd(&B, true, true) (dgesvd) with a fixed sign rule: the largest-magnitude entry of every right vector is positive
    (first such entry on ties)."""
    u, s, vt = sla.svd(B, full_matrices=True, lapack_driver="gesvd", check_finite=False)
    for i in range(vt.shape[0]):
        j = int(np.argmax(np.abs(vt[i])))
        if vt[i, j] < 0.0:
            vt[i] = -vt[i]
            u[:, i] = -u[:, i]
    return u, s, vt


def _norm(x):  # irlba.rs:12-14
    return float(np.sqrt(np.sum(x * x)))


def _orthog(y, X):  # irlba.rs:19-22
    return y - X.dot(X.T.dot(y))


def _invcheck(x):  # irlba.rs:25-33
    return 1.0 / x if x > 2.0 * np.finfo(np.float64).eps else 0.0


def irlba(A, nu: int, tol: float = 1e-4, maxit: int = 50, v0: Optional[np.ndarray] = None,
          snoop: Optional[Callable[[float], bool]] = None, threads: bool = False):
    """irlba.rs:71-215.  Returns (U m x nu, sigma nu, V n x nu, mprod, iterations)."""
    m, n = A.shape
    if m < 2 or n < 2:
        raise ValueError("The input matrix must be at least 2x2.")    # :84
    if nu > min(m, n):
        raise ValueError("invalid k")                                 # :85
    m_b = min(nu + 20, min(3 * nu, n))                                # :87
    mprod, it, j, k = 0, 0, 0, nu
    smax = -np.finfo(np.float64).max                                  # f64::MIN
    V = np.zeros((n, m_b)); W = np.zeros((m, m_b)); F = np.zeros(n); B = np.zeros((m_b, m_b))
    u = np.zeros((1, 1)); sigma = np.zeros(nu); vt = np.zeros((1, 1))
    r = irlba_start(0, n) if v0 is None else np.array(v0, dtype=np.float64)
    V[:, 0] = r * (1.0 / _norm(r))                                    # :111-113
    dot1 = lambda x: A.dot(x.reshape(-1, 1), threads)[:, 0]
    rdot1 = lambda y: A.rdot(y.reshape(1, -1), threads)[0, :]
    fnorm = 0.0
    while it < maxit:
        if it > 0:
            j = k
        W[:, j] = dot1(V[:, j]); mprod += 1                           # :121
        if it > 0:
            W[:, k] = _orthog(W[:, j], W[:, :j])                      # :125-126
        s = _norm(W[:, j]); sinv = _invcheck(s)
        W[:, j] *= sinv
        fnorm = 0.0
        while j < m_b:                                                # Lanczos process :136-167
            F = rdot1(W[:, j]); mprod += 1
            F = F - V[:, j] * s
            F = _orthog(F, V[:, :j + 1])
            fnorm = _norm(F)
            F = F * _invcheck(fnorm)
            if j == m_b - 1:
                B[j, j] = s
            else:
                V[:, j + 1] = F
                B[j, j] = s
                B[j, j + 1] = fnorm
                mprod += 1                                            # :152-153 (the reference forms A.V[:, j+1] twice: same values)
                nw = dot1(V[:, j + 1])
                nw = nw - W[:, j] * fnorm
                nw = _orthog(nw, W[:, :j + 1])
                s = _norm(nw); sinv = _invcheck(s)
                W[:, j + 1] = nw * sinv
            j += 1
        u, sigma, vt = _svd_canonical(B)                              # :169-172
        resid = fnorm * u[m_b - 1, :]
        smax = sigma[0] if sigma[0] > smax else smax
        num_converged = int(sum(1 for i in range(nu) if resid[i] < tol * smax))   # no abs: as the reference
        if num_converged < nu:
            k = max(num_converged + nu, k)
            k = min(k, m_b - 3)
        else:
            break
        V[:, :k] = V[:, :m_b].dot(vt.T[:, :k])                        # :191-193
        V[:, k] = F
        B = np.zeros((m_b, m_b))
        for l in range(k):
            B[l, l] = sigma[l]
        B[:k, k] = resid[:k]
        W[:, :k] = W[:, :m_b].dot(u[:, :k])                           # :202-203
        it += 1
        if snoop is not None and snoop(it / maxit):
            raise CancellationError()
    U = W[:, :m_b].dot(u[:, :nu])
    Vo = V[:, :m_b].dot(vt.T[:, :nu])
    return U, sigma[:nu].copy(), Vo, mprod, it


class Irlba:  # irlba.rs:36-69
    def __init__(self, tol: float = 0.0001, max_iter: int = 50):
        self.tol, self.max_iter = tol, max_iter

    def run_pca(self, array, k: int, v0=None, threads=False):
        u, s, v, _, _ = irlba(array, k, self.tol, self.max_iter, v0, None, threads)
        return u, s, v


# --------------------------------------------------------------------------------------
# The per-gene moment consumers of diff-exp on size-normalized counts: sqz/src/mat.rs:285-374 (mean_var_axis, mean_var_rows),
# :414-476 (sum_cols, sum_rows), :484-583 (sum_rows_dual); diff-exp/src/diff_exp.rs:314-334 (size_factors), :340-358
# (SizeNormalized: v / size_factor[c]), diff-exp/src/stat.rs:107-163 (median = 50th percentile, linear interpolation).
# Plain numpy over the cell-major arrays (summation in storage order like the reference's CSC branch).
# --------------------------------------------------------------------------------------
def percentile_median(x: np.ndarray) -> float:  # stat.rs:116-118, :140-163
    s = np.sort(np.asarray(x, dtype=np.float64))
    if s.shape[0] == 1:
        return float(s[0])
    rank = 0.5 * (s.shape[0] - 1)
    lo = int(np.floor(rank))
    d = rank - lo
    return float(s[lo] + (s[lo + 1] - s[lo]) * d)


def size_factors(mat: CountMatrix, cell_indices=None, umi_counts=None) -> np.ndarray:  # diff_exp.rs:314-334
    if umi_counts is not None:
        cpc = np.asarray(umi_counts, dtype=np.float64)
    else:
        tot = mat.sum_axis_u64(0).astype(np.float64)
        cpc = tot if cell_indices is None else tot[np.asarray(cell_indices, dtype=np.int64)]
    med = percentile_median(cpc)
    if cell_indices is None:
        return cpc / med
    out = np.zeros(mat.cols)
    out[np.asarray(cell_indices, dtype=np.int64)] = cpc / med
    return out


def _mapped_entries(mat: CountMatrix, sf: Optional[np.ndarray]):
    ip, g, c = mat.cell_major()
    cell = np.repeat(np.arange(mat.cols, dtype=np.int64), np.diff(ip).astype(np.int64))
    val = c.astype(np.float64)
    if sf is not None:
        d = np.where(np.isnan(sf), 0.0, np.asarray(sf, dtype=np.float64))  # SizeNormalized::new, diff_exp.rs:341-344
        with np.errstate(divide="ignore", invalid="ignore"):
            val = val / d[cell]
    return g.astype(np.int64), cell, val


def mean_var_axis(mat: CountMatrix, axis: int, sf: Optional[np.ndarray] = None):  # mat.rs:285-329
    g, cell, val = _mapped_entries(mat, sf)
    key, sz = (cell, mat.cols) if axis == 0 else (g, mat.rows)
    means = np.bincount(key, weights=val, minlength=sz)
    sq = np.bincount(key, weights=val * val, minlength=sz)
    mm = float(mat.shape[axis])
    means = means / mm
    return means, sq / mm - means * means


def mean_var_rows(mat: CountMatrix, cols, sf: Optional[np.ndarray] = None):  # mat.rs:332-374
    g, cell, val = _mapped_entries(mat, sf)
    cols = np.asarray(cols, dtype=np.int64)
    mult = np.bincount(cols, minlength=mat.cols)  # CSC branch: a column listed twice is walked twice
    keep = mult[cell] > 0                          # only the listed columns are walked (their size factors may be the only non-zero ones)
    g, cell, val = g[keep], cell[keep], val[keep]
    w = mult[cell].astype(np.float64)
    means = np.bincount(g, weights=val * w, minlength=mat.rows)
    sq = np.bincount(g, weights=val * val * w, minlength=mat.rows)
    mm = float(cols.shape[0])
    means = means / mm
    return means, sq / mm - means * means


def sum_rows_dual(mat: CountMatrix, cols1, cols2):  # mat.rs:484-583 (u64 sums: exact)
    ip, g, c = mat.cell_major()
    cell = np.repeat(np.arange(mat.cols, dtype=np.int64), np.diff(ip).astype(np.int64))
    out = []
    for cols in (cols1, cols2):
        sel = np.zeros(mat.cols, dtype=bool)
        sel[np.asarray(cols, dtype=np.int64)] = True  # merge_join_by over sorted unique lists: membership
        k = sel[cell]
        out.append(np.bincount(g[k].astype(np.int64), weights=c[k].astype(np.float64), minlength=mat.rows).astype(np.uint64))
    return out[0], out[1]


def frobenius(a: np.ndarray) -> float:  # dim_red/mod.rs:114-122
    return float(np.sqrt(np.sum(a * a)) / (a.shape[0] * a.shape[1]))


def principal_angle_sin(u0: np.ndarray, u1: np.ndarray) -> float:
    """|| (I - U0 U0^T) U1 ||_2 for orthonormal-column U0, U1."""
    q0, _ = np.linalg.qr(u0)
    q1, _ = np.linalg.qr(u1)
    r = q1 - q0.dot(q0.T.dot(q1))
    return float(np.linalg.norm(r, 2))


# --------------------------------------------------------------------------------------
# kNN on the scores: scan-rs/src/nn.rs:38-83.  The reference answers queries through a ball tree (third-party crate
# `ball-tree`, git dependency in Cargo.lock); its result is the exact k nearest neighbours, which is what the reference's own
# tests check against an exhaustive search (nn.rs:104-152).  This restates that exhaustive search with Pt::distance (nn.rs:12-20).
def find_nn(v: np.ndarray, k: int, points: np.ndarray, include_self: bool, self_offset: int = 0) -> np.ndarray:
    v = np.asarray(v, dtype=np.float64)
    points = np.asarray(points, dtype=np.float64)
    out = np.full((v.shape[0], k), 0xFFFFFFFF, dtype=np.uint32)
    for i in range(v.shape[0]):
        d = np.zeros(points.shape[0])
        for j in range(points.shape[1]):  # sum in coordinate order, like the iterator chain in Pt::distance
            d = d + (points[:, j] - v[i, j]) ** 2
        d = np.sqrt(d)
        order = np.argsort(d, kind="stable")  # ties: lower index first
        if not include_self:
            order = order[order != self_offset + i]
        order = order[:k]
        out[i, : len(order)] = order
    return out


def knn(v: np.ndarray, k: int) -> np.ndarray:
    return find_nn(v, k, v, include_self=False)
