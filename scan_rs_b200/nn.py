"""kNN on the PCA scores (scan-rs/src/nn.rs:38-83) through the C ABI's sb_knn (csrc/knn.cu)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def find_nn(ctx, v: np.ndarray, k: int, points: np.ndarray, include_self: bool, self_offset: int = 0) -> np.ndarray:
    """nn.rs:61-83: the k nearest rows of `points` to each row of `v` (indices, nearest first; 0xFFFFFFFF pads short rows)."""
    v = np.ascontiguousarray(v, dtype=np.float64)
    points = np.ascontiguousarray(points, dtype=np.float64)
    if v.ndim != 2 or points.ndim != 2 or v.shape[1] != points.shape[1]:
        raise ValueError("Dimension mismatch")
    out = np.full((v.shape[0], k), 0xFFFFFFFF, dtype=np.uint32)
    L.check(L.lib().sb_knn(ctx._h, L.vp(points), C.c_uint64(points.shape[0]), C.c_uint32(v.shape[1]), L.vp(v), C.c_uint64(v.shape[0]),
                           C.c_uint32(k), C.c_int(1 if include_self else 0), C.c_uint64(self_offset), L.vp(out)))
    return out


def knn(ctx, v: np.ndarray, k: int) -> np.ndarray:
    """nn.rs:38-58: the k nearest other rows of `v` for each row (the reference also returns its ball tree; there is none here)."""
    return find_nn(ctx, v, k, v, include_self=False)
