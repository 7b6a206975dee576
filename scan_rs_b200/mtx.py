"""Mirror of scan_rs::mtx (scan-rs/src/mtx.rs:10-51): gzipped MatrixMarket -> device-resident AdaptiveMat.

Host-side parsing only (this is I/O); the triplets go straight into the gene-major
upload of the C ABI.  Like the reference: lines starting with '%' are skipped, the first other line is
`NROW NCOL NNZ`, every further line is `ROW COL VAL` (1-based, VAL parsed as u32), duplicates are summed
(sprs TriMat::to_csr), rows are features and columns are barcodes."""
from __future__ import annotations

import gzip
import io

import numpy as np

from .sqz import AdaptiveMat, Context


def parse_mtx(path: str):
    """-> (rows, cols, indptr u64, idx u32, val u32) gene-major CSR with ascending column indices."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        header = None
        while header is None:
            line = f.readline()
            if not line:
                raise ValueError("no matrix found")            # mtx.rs:48
            if line.startswith(b"%") or not line.strip():
                continue
            header = line.split()
        if len(header) < 3:
            raise ValueError("no NROW" if len(header) < 1 else "no NCOL" if len(header) < 2 else "no NNZ")
        nrow, ncol, nnz = int(header[0]), int(header[1]), int(header[2])
        body = f.read()
    if body.strip():
        lines = [ln for ln in body.split(b"\n") if ln.strip() and not ln.startswith(b"%")]
        trip = np.loadtxt(io.BytesIO(b"\n".join(lines)), dtype=np.int64, ndmin=2)
        if trip.shape[1] < 3:
            raise ValueError("missing VAL")
    else:
        trip = np.zeros((0, 3), dtype=np.int64)
    r, c, v = trip[:, 0] - 1, trip[:, 1] - 1, trip[:, 2]
    if len(r) and (r.min() < 0 or r.max() >= nrow or c.min() < 0 or c.max() >= ncol or v.min() < 0 or v.max() > 0xFFFFFFFF):
        raise ValueError("triplet out of range")
    # TriMat::to_csr: sort by (row, col), sum duplicates
    order = np.lexsort((c, r))
    r, c, v = r[order], c[order], v[order]
    if len(r):
        key = r * ncol + c
        first = np.r_[True, key[1:] != key[:-1]]
        seg = np.cumsum(first) - 1
        vs = np.zeros(int(seg[-1]) + 1, dtype=np.int64)
        np.add.at(vs, seg, v)
        r, c, v = r[first], c[first], vs
        keep = v != 0
        r, c, v = r[keep], c[keep], v[keep]
    indptr = np.zeros(nrow + 1, dtype=np.uint64)
    np.cumsum(np.bincount(r, minlength=nrow), out=indptr[1:])
    return nrow, ncol, indptr, c.astype(np.uint32), (v & 0xFFFFFFFF).astype(np.uint32)


def load_mtx(ctx: Context, path: str) -> AdaptiveMat:
    """load_mtx (mtx.rs:10-51) onto the device."""
    nrow, ncol, indptr, idx, val = parse_mtx(path)
    return AdaptiveMat.from_csr(ctx, nrow, ncol, indptr, idx, val)
