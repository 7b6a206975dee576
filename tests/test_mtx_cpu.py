"""Host-side MatrixMarket parser (mirror of scan-rs/src/mtx.rs:10-51): comments, 1-based indices, duplicate
summing, row-major output.  The upload half is covered on the GPU (tests/test_gpu_parity.py::test_load_mtx)."""
import gzip

import numpy as np
import pytest

from scan_rs_b200.mtx import parse_mtx

MTX = b"""%%MatrixMarket matrix coordinate integer general
% a comment
4 5 7
1 1 136
2 5 391
1 2 936
4 1 396
2 5 9
3 3 780
2 4 8
"""


def test_parse_mtx(tmp_path):
    p = tmp_path / "m.mtx.gz"
    with gzip.open(p, "wb") as f:
        f.write(MTX)
    nrow, ncol, indptr, idx, val = parse_mtx(str(p))
    assert (nrow, ncol) == (4, 5)
    dense = np.zeros((4, 5), dtype=np.uint32)
    for r in range(4):
        s, e = int(indptr[r]), int(indptr[r + 1])
        assert (np.diff(idx[s:e].astype(np.int64)) > 0).all()
        dense[r, idx[s:e]] = val[s:e]
    want = np.zeros((4, 5), dtype=np.uint32)
    want[0, 0], want[0, 1], want[1, 3], want[1, 4], want[2, 2], want[3, 0] = 136, 936, 8, 400, 780, 396
    np.testing.assert_array_equal(dense, want)


def test_parse_mtx_errors(tmp_path):
    p = tmp_path / "e.mtx"
    p.write_bytes(b"% only comments\n")
    with pytest.raises(ValueError, match="no matrix found"):
        parse_mtx(str(p))
    p.write_bytes(b"2 2 1\n3 1 5\n")
    with pytest.raises(ValueError, match="out of range"):
        parse_mtx(str(p))
