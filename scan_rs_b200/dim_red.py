"""Mirror of scan_rs::dim_red (scan-rs/src/dim_red/{mod,bk_svd,rand_svd,irlba}.rs) on the device path."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L
from .snoop import NoOpSnoop
from .sqz import LowRankOffset


def pinned_outputs(m: int, n: int, k: int):
    """(U, S, V) result buffers in page-locked memory, to be passed as `out=` and reused across calls: the V block is
    n x k and a pageable device-to-host copy of it is several times slower (page-locking itself costs ~1 ms/MB, so
    allocate once)."""
    return L.pinned_empty((m, k), np.float64), L.pinned_empty((k,), np.float64), L.pinned_empty((n, k), np.float64)


def _outputs(m, n, k, out=None):
    if out is None:
        return np.zeros((m, k)), np.zeros(k), np.zeros((n, k))
    U, S, V = out
    for a, shape in ((U, (m, k)), (S, (k,)), (V, (n, k))):
        if a.shape != shape or a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
            raise ValueError(f"out buffer must be C-contiguous float64 of shape {shape}")
    return U, S, V


def variance_explained(array: LowRankOffset, s) -> np.ndarray:
    """DERIVED output, not part of the reference's PcaResult (dim_red/mod.rs:47 returns (u, d, v) only): the fraction of the
    normalized matrix's squared Frobenius norm carried by each singular value, sigma_i^2 / ||A||_F^2."""
    return np.asarray(s, dtype=np.float64) ** 2 / array.frobenius_sq()


def omega(seed: int, rows: int, cols: int) -> np.ndarray:
    """The start block: SmallRng::seed_from_u64(seed) + Uniform(-1, 1), row-major (bk_svd.rs:83-84)."""
    out = np.zeros((rows, cols))
    L.check(L.lib().sb_omega(C.c_uint64(seed), C.c_uint64(rows), C.c_uint64(cols), L.vp(out)))
    return out


def _make_cb(snoop):
    # NoOpSnoop never cancels and drops the progress (snoop/src/lib.rs:60-85): no callback at all, so the library neither
    # synchronises nor agrees a cancel flag across ranks at the milestones (pca.cu: progress)
    if type(snoop) is NoOpSnoop:
        return None

    def _cb(frac, _user):
        # set_progress_check (snoop/src/lib.rs:45-57): cancelled -> Err before the progress is stored
        if snoop.is_cancelled():
            return 1
        snoop.set_progress(frac)
        return 0
    return L.PROGRESS_CB(_cb)


def svd_bk(A: LowRankOffset, k: int, b: int, n_iter: int, seed: int = 0, snoop=None, omega_block: Optional[np.ndarray] = None, out=None):
    """bk_svd.rs:57-146 -> (U m x k, sigma k, Va k x n_local)."""
    m, n = A.shape()
    U, S, V = _outputs(m, n, k, out)
    cb = _make_cb(snoop or NoOpSnoop())
    om = None if omega_block is None else np.ascontiguousarray(omega_block, dtype=np.float64)
    L.check(L.lib().sb_bksvd(A._h, C.c_uint32(k), C.c_uint32(b), C.c_uint32(n_iter), C.c_uint64(seed), L.vp(om), cb, None,
                             L.vp(U), L.vp(S), L.vp(V)))
    return U, S, V.T


def svd_rand(A: LowRankOffset, k: int, l: int, n_iter: int, seed: int = 0, omega_block: Optional[np.ndarray] = None, out=None):
    """rand_svd.rs:54-129 -> (U, sigma, Va k x n)."""
    m, n = A.shape()
    U, S, V = _outputs(m, n, k, out)
    om = None if omega_block is None else np.ascontiguousarray(omega_block, dtype=np.float64)
    L.check(L.lib().sb_randsvd(A._h, C.c_uint32(k), C.c_uint32(l), C.c_uint32(n_iter), C.c_uint64(seed), L.vp(om),
                               L.vp(U), L.vp(S), L.vp(V)))
    return U, S, V.T


class BkSvd:
    """bk_svd.rs:16-53: defaults k_multiplier = 2.0, n_iter = 5."""

    def __init__(self, k_multiplier: float = 2.0, n_iter: int = 5):
        self.k_multiplier, self.n_iter = k_multiplier, n_iter

    def run_pca_cancellable(self, array: LowRankOffset, k: int, snoop, out=None):
        """-> (u m x k, s k, v n x k): PcaResult with `vt.reversed_axes()` (bk_svd.rs:48-52)."""
        m, n = array.shape()
        U, S, V = _outputs(m, n, k, out)
        cb = _make_cb(snoop)
        L.check(L.lib().sb_bksvd_run_pca(array._h, C.c_uint32(k), C.c_double(self.k_multiplier), C.c_uint32(self.n_iter), cb, None,
                                         L.vp(U), L.vp(S), L.vp(V)))
        return U, S, V

    def run_pca(self, array: LowRankOffset, k: int, out=None):  # dim_red/mod.rs:108-110
        return self.run_pca_cancellable(array, k, NoOpSnoop(), out)


class RandSvd:
    """rand_svd.rs:13-50: defaults l_multiplier = 10.0, n_iter = 2 (ignores the snoop, :44-45)."""

    def __init__(self, l_multiplier: float = 10.0, n_iter: int = 2):
        self.l_multiplier, self.n_iter = l_multiplier, n_iter

    def run_pca_cancellable(self, array: LowRankOffset, k: int, _snoop=None, out=None):
        m, n = array.shape()
        U, S, V = _outputs(m, n, k, out)
        L.check(L.lib().sb_randsvd_run_pca(array._h, C.c_uint32(k), C.c_double(self.l_multiplier), C.c_uint32(self.n_iter),
                                           L.vp(U), L.vp(S), L.vp(V)))
        return U, S, V

    def run_pca(self, array: LowRankOffset, k: int, out=None):
        return self.run_pca_cancellable(array, k, None, out)


def irlba_start(seed: int, n: int) -> np.ndarray:
    """The builder-defined default start vector of `irlba` (the reference's rand_distr Normal stream cannot be restated offline)."""
    out = np.zeros(n)
    L.check(L.lib().sb_irlba_start(C.c_uint64(seed), C.c_uint64(n), L.vp(out)))
    return out


def irlba(A: LowRankOffset, nu: int, tol: float = 1e-4, maxit: int = 50, v0: Optional[np.ndarray] = None, snoop=None, info: Optional[dict] = None):
    """irlba.rs:71-215 -> (U m x nu, sigma nu, V n_local x nu); `info` receives mprod (matrix products) and iterations."""
    m, n = A.shape()
    U, S, V = np.zeros((m, nu)), np.zeros(nu), np.zeros((n, nu))
    cb = _make_cb(snoop or NoOpSnoop())
    v = None if v0 is None else np.ascontiguousarray(v0, dtype=np.float64)
    mprod, iters = C.c_uint32(), C.c_uint32()
    L.check(L.lib().sb_irlba(A._h, C.c_uint32(nu), C.c_double(tol), C.c_uint32(maxit), L.vp(v), cb, None, L.vp(U), L.vp(S), L.vp(V),
                             C.byref(mprod), C.byref(iters)))
    if info is not None:
        info["mprod"], info["iterations"] = mprod.value, iters.value
    return U, S, V


class Irlba:
    """irlba.rs:36-69: defaults tol = 0.0001, max_iter = 50."""

    def __init__(self, tol: float = 0.0001, max_iter: int = 50):
        self.tol, self.max_iter = tol, max_iter

    def run_pca_cancellable(self, array: LowRankOffset, k: int, snoop, v0=None):
        return irlba(array, k, self.tol, self.max_iter, v0, snoop)

    def run_pca(self, array: LowRankOffset, k: int, v0=None):
        return self.run_pca_cancellable(array, k, NoOpSnoop(), v0)
