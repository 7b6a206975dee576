"""Host-side mirror of the reference's `sqz` types for the device path.

`AdaptiveMat`  <-> sqz::AdaptiveMat<u32> (sqz/src/mat.rs:34-42): the count matrix, uploaded once and
                   kept device-resident (cell-major + cell-panelled gene-major u32/u32 copies).
`LowRankOffset` <-> sqz::LowRankOffset (sqz/src/low_rank_offset.rs:12-16): the normalized matrix
                   A = map(counts) + u.v, never materialised.
Method names and argument meanings follow the reference; everything calls the C ABI."""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional, Sequence

import numpy as np

from . import _lib as L


_VARIANT_ID = {"D3": 0, "D4": 1, "D8": 2, "D16": 3, "V": 4, "S3": 5, "S4": 6, "S8": 7}


class SbAdaptiveVec(C.Structure):  # include/scanb200.h: sb_adaptive_vec
    _fields_ = [("variant", C.c_uint32), ("reserved", C.c_uint32), ("len", C.c_uint64), ("dense", C.c_void_p), ("fb_idx", C.c_void_p),
                ("fb_val", C.c_void_p), ("fb_len", C.c_uint64), ("index_bytes", C.c_void_p), ("n_index", C.c_uint64),
                ("block_starts", C.c_void_p), ("n_block_starts", C.c_uint64)]


def adaptive_views(vecs):
    """ctypes array of sb_adaptive_vec over the numpy buffers of `vecs` (+ the list that keeps those buffers alive)"""
    arr = (SbAdaptiveVec * max(1, len(vecs)))()
    keep = []
    ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None
    for i, v in enumerate(vecs):
        bufs = [None if b is None else np.ascontiguousarray(b) for b in (v.dense, v.fb_idx, v.fb_val, v.index_bytes, v.block_starts)]
        keep.append(bufs)
        d, fi, fv, ib, bs = bufs
        arr[i] = SbAdaptiveVec(_VARIANT_ID[v.variant], 0, int(v.len), ptr(d), ptr(fi), ptr(fv), 0 if fi is None else int(fi.size), ptr(ib),
                               0 if ib is None else int(ib.size), ptr(bs), 0 if bs is None else int(bs.size))
    return arr, keep


def adaptive_decode(vec):
    """AdaptiveVec::foreach through the C ABI (sb_adaptive_decode_vec): (indexes, values) of one vector"""
    arr, keep = adaptive_views([vec])
    nnz = C.c_uint64()
    L.check(L.lib().sb_adaptive_decode_vec(arr, None, None, C.c_uint64(0), C.byref(nnz)))
    idx, val = np.zeros(nnz.value, dtype=np.uint32), np.zeros(nnz.value, dtype=np.uint32)
    L.check(L.lib().sb_adaptive_decode_vec(arr, L.vp(idx), L.vp(val), C.c_uint64(nnz.value), C.byref(nnz)))
    return idx, val


class Context:
    """One GPU + stream (+ optional NCCL communicator)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        L.check(L.lib().sb_init(C.c_int(device), C.byref(self._h)))
        self.device = device
        self.nranks, self.rank = 1, 0
        self._mats, self._nmats = weakref.WeakSet(), weakref.WeakSet()

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        L.check(L.lib().sb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, nranks: int, rank: int, uid: bytes) -> None:
        L.check(L.lib().sb_comm_init(self._h, C.c_int(nranks), C.c_int(rank), C.c_char_p(uid)))
        self.nranks, self.rank = nranks, rank

    def set_option(self, name: str, value: float):
        L.check(L.lib().sb_set_option(self._h, C.c_char_p(name.encode()), C.c_double(value)))

    def sync(self):
        L.check(L.lib().sb_sync(self._h))

    def timer_begin(self):
        L.check(L.lib().sb_timer_begin(self._h))

    def timer_end(self) -> float:
        ms = C.c_float(0)
        L.check(L.lib().sb_timer_end(self._h, C.byref(ms)))
        return float(ms.value)

    def flush_l2(self):
        L.check(L.lib().sb_flush_l2(self._h))

    def profile_enable(self, on: bool = True):
        L.check(L.lib().sb_profile_enable(self._h, C.c_int(int(on))))

    def profile_reset(self):
        L.check(L.lib().sb_profile_reset(self._h))

    def pca_diagnostics(self) -> dict:
        """Of the last BkSvd on this context: condition estimate of the Krylov basis' R, probe residual of the projection
        identity (units of sigma_1) and the number of fallbacks taken (include/scanb200.h: sb_pca_diagnostics)."""
        cond, resid, fb = C.c_double(), C.c_double(), C.c_int()
        L.check(L.lib().sb_pca_diagnostics(self._h, C.byref(cond), C.byref(resid), C.byref(fb)))
        return {"cond_r": cond.value, "probe_resid": resid.value, "fallbacks": fb.value}

    def profile(self) -> dict:
        p = L.SbProfile()
        L.check(L.lib().sb_profile_get(self._h, C.byref(p)))
        return p.as_dict()

    @classmethod
    def _adopt(cls, handle, device: int, nranks: int, rank: int) -> "Context":
        """A context owned by a MultiContext (multi.py): same interface, shut down by its owner."""
        self = cls.__new__(cls)
        self._h = handle
        self.device = device
        self.nranks, self.rank = nranks, rank
        self._mats, self._nmats = weakref.WeakSet(), weakref.WeakSet()
        self._owned = False
        return self

    def _release_handles(self):
        for a in list(self._nmats):  # handles borrow the context: free them first
            a.free()
        for m in list(self._mats):
            m.free()

    def close(self):
        if self._h:
            self._release_handles()
            if getattr(self, "_owned", True):
                L.lib().sb_shutdown(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class AdaptiveMat:
    """Device-resident genes x cells u32 count matrix (sqz/src/mat.rs:34-42)."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self._h = ctx, handle
        self._views = weakref.WeakSet()  # LowRankOffsets borrowing this matrix: freed before it
        ctx._mats.add(self)

    # ---- constructors
    @classmethod
    def from_csr(cls, ctx: Context, rows: int, cols: int, indptr, idx, val) -> "AdaptiveMat":
        """Gene-major CSR arrays (from_csmat on a CSR sprs matrix, mat.rs:92-124)."""
        return cls._upload(ctx, L.SB_GENE_MAJOR, rows, cols, indptr, idx, val)

    @classmethod
    def from_csc(cls, ctx: Context, rows: int, cols: int, indptr, idx, val) -> "AdaptiveMat":
        """Cell-major (CSC of the genes x cells matrix) arrays."""
        return cls._upload(ctx, L.SB_CELL_MAJOR, rows, cols, indptr, idx, val)

    @classmethod
    def from_dense(cls, ctx: Context, dense) -> "AdaptiveMat":  # mat.rs:586-609
        dense = np.asarray(dense)
        rows, cols = dense.shape
        indptr, idx, val = [0], [], []
        for r in range(rows):
            nz = np.nonzero(dense[r])[0]
            idx.extend(nz.tolist())
            val.extend(dense[r, nz].tolist())
            indptr.append(len(idx))
        return cls.from_csr(ctx, rows, cols, indptr, idx, val)

    @classmethod
    def _upload(cls, ctx, major, rows, cols, indptr, idx, val):
        indptr = np.ascontiguousarray(indptr, dtype=np.uint64)
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        val = np.ascontiguousarray(val, dtype=np.uint32)
        nvec = rows if major == L.SB_GENE_MAJOR else cols
        if indptr.shape != (nvec + 1,) or idx.shape != val.shape or idx.shape[0] != int(indptr[-1]):
            raise ValueError("inconsistent CSR/CSC arrays")
        h = C.c_void_p()
        L.check(L.lib().sb_upload(ctx._h, C.c_int(major), C.c_uint32(rows), C.c_uint64(cols), L.vp(indptr), L.vp(idx),
                                  L.vp(val), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_csc_unsorted(cls, ctx: Context, rows: int, cols: int, indptr, idx, val) -> "AdaptiveMat":
        """Cell-major arrays with the gene indices of a cell in any order (hdf5-io/src/matrix.rs:63-78: the Cell Ranger 3 repair
        `new_from_unsorted_csc`): sorted per cell on the device; duplicates inside a cell are an error."""
        ip = np.ascontiguousarray(indptr, dtype=np.uint64)
        ix = np.ascontiguousarray(idx, dtype=np.uint32)
        vv = np.ascontiguousarray(val, dtype=np.uint32)
        h = C.c_void_p()
        L.check(L.lib().sb_upload_unsorted(ctx._h, C.c_uint32(rows), C.c_uint64(cols), L.vp(ip), L.vp(ix), L.vp(vv), C.byref(h)))
        return cls(ctx, h)

    def filter_genes(self, type_keep=None, min_total: int = 0):
        """compute_genes_filter + the row selection of read_adaptive_csr_matrix (hdf5-io/src/matrix.rs:93-192) -> (matrix of the
        surviving features, their indices in file order).  type_keep: boolean mask of the features whose type is retained."""
        tk = None if type_keep is None else np.ascontiguousarray(type_keep, dtype=np.uint8)
        kept = np.zeros(self.rows(), dtype=np.uint32)
        nk = C.c_uint32()
        h = C.c_void_p()
        L.check(L.lib().sb_filter_genes(self._h, L.vp(tk), C.c_uint64(min_total), L.vp(kept), C.byref(nk), C.byref(h)))
        return AdaptiveMat(self.ctx, h), kept[: nk.value].copy()

    @staticmethod
    def compact_csc(idx, val):
        """Cell-major u32 index / count arrays -> the narrow host form of sb_upload_compact:
        (idx16, cnt8, big_pos, big_cnt); counts >= 255 go to the side list."""
        idx = np.asarray(idx)
        val = np.asarray(val)
        if idx.size and int(idx.max()) > 0xFFFF:
            raise ValueError("compact form needs gene indices below 65536")
        big_pos = np.flatnonzero(val >= 255).astype(np.uint64)
        return (idx.astype(np.uint16), np.minimum(val, 255).astype(np.uint8), big_pos, val[big_pos].astype(np.uint32))

    @classmethod
    def from_csc_compact(cls, ctx: Context, rows: int, cols: int, indptr, idx16, cnt8, big_pos=None, big_cnt=None) -> "AdaptiveMat":
        """Cell-major upload in the narrow host form (3 B per entry over PCIe instead of 8); rows <= 65536."""
        indptr = np.ascontiguousarray(indptr, dtype=np.uint64)
        idx16 = np.ascontiguousarray(idx16, dtype=np.uint16)
        cnt8 = np.ascontiguousarray(cnt8, dtype=np.uint8)
        big_pos = np.ascontiguousarray(big_pos if big_pos is not None else [], dtype=np.uint64)
        big_cnt = np.ascontiguousarray(big_cnt if big_cnt is not None else [], dtype=np.uint32)
        if indptr.shape != (cols + 1,) or idx16.shape != cnt8.shape or idx16.shape[0] != int(indptr[-1]) or big_pos.shape != big_cnt.shape:
            raise ValueError("inconsistent compact CSC arrays")
        h = C.c_void_p()
        L.check(L.lib().sb_upload_compact(ctx._h, C.c_uint32(rows), C.c_uint64(cols), L.vp(indptr), L.vp(idx16), L.vp(cnt8),
                                          C.c_uint64(big_pos.shape[0]), L.vp(big_pos), L.vp(big_cnt), C.byref(h)))
        return cls(ctx, h)

    @staticmethod
    def pack_csc(indptr, idx, val, threads: int = 0, pinned: bool = False):
        """Cell-major u32 index / count arrays -> the packed host form of sb_upload_packed (sb_pack_csc_count / _fill):
        (dgene u8[nnz], cnt4 u8[(nnz+1)/2], esc_pos, esc_gene, big_pos, big_cnt).  pinned: page-locked output arrays."""
        indptr = np.ascontiguousarray(indptr, dtype=np.uint64)
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        val = np.ascontiguousarray(val, dtype=np.uint32)
        n, nnz = indptr.shape[0] - 1, int(indptr[-1])
        if idx.shape != (nnz,) or val.shape != (nnz,):
            raise ValueError("inconsistent CSC arrays")
        ne, nb = C.c_uint64(), C.c_uint64()
        L.check(L.lib().sb_pack_csc_count(C.c_uint64(n), L.vp(indptr), L.vp(idx), L.vp(val), C.c_int(threads), C.byref(ne), C.byref(nb)))
        new = (lambda k, dt: L.pinned_empty((k,), dt)) if pinned else (lambda k, dt: np.empty(k, dtype=dt))
        dgene, cnt4 = new(nnz, np.uint8), new((nnz + 1) // 2, np.uint8)
        esc_pos, esc_gene = new(ne.value, np.uint64), new(ne.value, np.uint32)
        big_pos, big_cnt = new(nb.value, np.uint64), new(nb.value, np.uint32)
        L.check(L.lib().sb_pack_csc_fill(C.c_uint64(n), L.vp(indptr), L.vp(idx), L.vp(val), C.c_int(threads), L.vp(dgene), L.vp(cnt4),
                                         L.vp(esc_pos), L.vp(esc_gene), L.vp(big_pos), L.vp(big_cnt)))
        return dgene, cnt4, esc_pos, esc_gene, big_pos, big_cnt

    @classmethod
    def from_csc_packed(cls, ctx: Context, rows: int, cols: int, indptr, dgene, cnt4, esc_pos, esc_gene, big_pos, big_cnt) -> "AdaptiveMat":
        """Cell-major upload in the packed host form (one byte of gene delta + one nibble of count per entry)."""
        indptr = np.ascontiguousarray(indptr, dtype=np.uint64)
        dgene = np.ascontiguousarray(dgene, dtype=np.uint8)
        cnt4 = np.ascontiguousarray(cnt4, dtype=np.uint8)
        esc_pos = np.ascontiguousarray(esc_pos, dtype=np.uint64)
        esc_gene = np.ascontiguousarray(esc_gene, dtype=np.uint32)
        big_pos = np.ascontiguousarray(big_pos, dtype=np.uint64)
        big_cnt = np.ascontiguousarray(big_cnt, dtype=np.uint32)
        nnz = int(indptr[-1])
        if indptr.shape != (cols + 1,) or dgene.shape != (nnz,) or cnt4.shape != ((nnz + 1) // 2,) or esc_pos.shape != esc_gene.shape or \
                big_pos.shape != big_cnt.shape:
            raise ValueError("inconsistent packed CSC arrays")
        h = C.c_void_p()
        L.check(L.lib().sb_upload_packed(ctx._h, C.c_uint32(rows), C.c_uint64(cols), L.vp(indptr), L.vp(dgene), L.vp(cnt4),
                                         C.c_uint64(esc_pos.shape[0]), L.vp(esc_pos), L.vp(esc_gene),
                                         C.c_uint64(big_pos.shape[0]), L.vp(big_pos), L.vp(big_cnt), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def _synth(cls, ctx, m, n_local, cell_offset, seed, pf, depth, cluster, r):
        h = C.c_void_p()
        L.check(L.lib().sb_synth_generate(ctx._h, C.c_uint32(m), C.c_uint64(n_local), C.c_uint64(cell_offset), C.c_uint64(seed),
                                          C.c_uint32(pf.shape[0]), L.vp(pf), L.vp(depth), L.vp(cluster), C.c_uint32(r),
                                          C.byref(h)))
        return cls(ctx, h)

    # ---- shape
    def _shape4(self):
        m, n, ng, nnz = C.c_uint32(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        L.check(L.lib().sb_mat_shape(self._h, C.byref(m), C.byref(n), C.byref(ng), C.byref(nnz)))
        return m.value, n.value, ng.value, nnz.value

    def rows(self) -> int:
        return self._shape4()[0]

    def cols(self) -> int:
        """Local cell count (== global without a communicator)."""
        return self._shape4()[1]

    def cols_global(self) -> int:
        return self._shape4()[2]

    def shape(self):
        s = self._shape4()
        return [s[0], s[1]]

    def nnz(self) -> int:
        return self._shape4()[3]

    # ---- reductions
    def sum_axis_u32(self, axis: int) -> np.ndarray:
        """sum_axis::<u32> (mat.rs:377-406): axis 0 -> per-cell totals (wrapping u32);
        axis 1 -> per-gene totals (computed exactly in u64 on the device, truncated like u32 `+=`)."""
        m, n, _, _ = self._shape4()
        if axis == 0:
            out = np.zeros(n, dtype=np.uint32)
            L.check(L.lib().sb_cell_totals(self._h, L.vp(out)))
            return out
        return (self.gene_totals() & np.uint64(0xFFFFFFFF)).astype(np.uint32)

    def gene_totals(self, square: bool = False) -> np.ndarray:
        out = np.zeros(self.rows(), dtype=np.uint64)
        L.check(L.lib().sb_gene_totals(self._h, C.c_int(int(square)), L.vp(out)))
        return out

    def gene_nnz(self) -> np.ndarray:
        out = np.zeros(self.rows(), dtype=np.uint64)
        L.check(L.lib().sb_gene_nnz(self._h, L.vp(out)))
        return out

    def median_cell_total(self) -> Optional[int]:
        med, ok = C.c_uint32(), C.c_int()
        L.check(L.lib().sb_median_cell_total(self._h, C.byref(med), C.byref(ok)))
        return int(med.value) if ok.value else None

    # ---- filters
    def partition_on_threshold(self, threshold: float):  # mat.rs:766-768
        return self.partition_on_thresholds(threshold, threshold)

    def partition_on_thresholds(self, row_threshold: Optional[float], col_threshold: Optional[float]):
        """mat.rs:772-889 -> (filtered, residual, selected_rows, selected_cols)."""
        m, n, _, _ = self._shape4()
        rows = np.zeros(m, dtype=np.uint64)
        cols = np.zeros(n, dtype=np.uint64)
        nr, nc = C.c_uint64(), C.c_uint64()
        kept, resid = C.c_void_p(), C.c_void_p()
        L.check(L.lib().sb_partition(self._h, C.c_int(row_threshold is not None), C.c_double(row_threshold or 0.0),
                                     C.c_int(col_threshold is not None), C.c_double(col_threshold or 0.0),
                                     C.byref(kept), C.byref(resid), L.vp(rows), C.byref(nr), L.vp(cols), C.byref(nc)))
        return (AdaptiveMat(self.ctx, kept), AdaptiveMat(self.ctx, resid), rows[: nr.value].astype(np.int64),
                cols[: nc.value].astype(np.int64))

    @classmethod
    def from_adaptive(cls, ctx, m: int, n: int, vecs, major: str = "gene", threads: int = 0) -> "AdaptiveMat":
        """Upload from the reference's own storage: a list of AdaptiveVec raw parts (objects with .variant/.len/.dense/.fb_idx/
        .fb_val/.index_bytes/.block_starts, e.g. oracle.adaptive_vec.Parts), one per gene (`major="gene"`, the reference's CSR
        storage, mtx.rs:49-50) or per cell.  Decoded on the host by the library (csrc/adaptive.cu)."""
        arr, keep = adaptive_views(vecs)
        h = C.c_void_p()
        L.check(L.lib().sb_upload_adaptive(ctx._h, C.c_int(L.SB_GENE_MAJOR if major == "gene" else L.SB_CELL_MAJOR), C.c_uint32(m), C.c_uint64(n),
                                           arr, C.c_int(threads), C.byref(h)))
        return cls(ctx, h)

    def select_rows(self, rows: Sequence[int]) -> "AdaptiveMat":  # mat.rs:1040-1071
        r = np.ascontiguousarray(rows, dtype=np.uint32)
        h = C.c_void_p()
        L.check(L.lib().sb_select_rows(self._h, L.vp(r), C.c_uint32(r.shape[0]), C.byref(h)))
        return AdaptiveMat(self.ctx, h)

    def select_cols(self, cols: Sequence[int]) -> "AdaptiveMat":  # mat.rs:1004-1037
        c = np.ascontiguousarray(cols, dtype=np.uint64)
        h = C.c_void_p()
        L.check(L.lib().sb_select_cols(self._h, L.vp(c), C.c_uint64(c.shape[0]), C.byref(h)))
        return AdaptiveMat(self.ctx, h)

    # ---- moment consumers (diff-exp's reads of the matrix); `size_factors`, when given, is the SizeNormalized view v / sf[c]
    @staticmethod
    def _cells(cols):
        return np.ascontiguousarray(cols, dtype=np.uint64)

    def mean_var_axis(self, axis: int, size_factors=None):  # mat.rs:285-329
        n_out = self.cols() if axis == 0 else self.rows()
        mean, var = np.zeros(n_out), np.zeros(n_out)
        sf = None if size_factors is None else np.ascontiguousarray(size_factors, dtype=np.float64)
        L.check(L.lib().sb_mean_var_axis(self._h, C.c_int(axis), L.vp(sf), L.vp(mean), L.vp(var)))
        return mean, var

    def mean_var_rows(self, cols, size_factors=None):  # mat.rs:332-374
        c = self._cells(cols)
        mean, var = np.zeros(self.rows()), np.zeros(self.rows())
        sf = None if size_factors is None else np.ascontiguousarray(size_factors, dtype=np.float64)
        L.check(L.lib().sb_mean_var_rows(self._h, L.vp(c), C.c_uint64(c.shape[0]), L.vp(sf), L.vp(mean), L.vp(var)))
        return mean, var

    def sum_rows(self, cols) -> np.ndarray:  # mat.rs:449-476
        c = self._cells(cols)
        out = np.zeros(self.rows(), dtype=np.uint64)
        L.check(L.lib().sb_sum_rows(self._h, L.vp(c), C.c_uint64(c.shape[0]), L.vp(out)))
        return out

    def sum_cols(self, cols) -> np.ndarray:  # mat.rs:414-446
        c = self._cells(cols)
        out = np.zeros(c.shape[0], dtype=np.uint64)
        L.check(L.lib().sb_sum_cols(self._h, L.vp(c), C.c_uint64(c.shape[0]), L.vp(out)))
        return out

    def sum_rows_dual(self, cols1, cols2):  # mat.rs:484-583
        c1, c2 = self._cells(cols1), self._cells(cols2)
        o1, o2 = np.zeros(self.rows(), dtype=np.uint64), np.zeros(self.rows(), dtype=np.uint64)
        L.check(L.lib().sb_sum_rows_dual(self._h, L.vp(c1), C.c_uint64(c1.shape[0]), L.vp(c2), C.c_uint64(c2.shape[0]), L.vp(o1), L.vp(o2)))
        return o1, o2

    def size_factors(self, cell_indices=None, umi_counts=None) -> np.ndarray:  # diff-exp/src/diff_exp.rs:314-334
        c = None if cell_indices is None else self._cells(cell_indices)
        u = None if umi_counts is None else np.ascontiguousarray(umi_counts, dtype=np.float64)
        out = np.zeros(self.cols())
        L.check(L.lib().sb_size_factors(self._h, L.vp(c), C.c_uint64(0 if c is None else c.shape[0]), L.vp(u), L.vp(out)))
        return out

    def hvg_select(self, n_top: int) -> np.ndarray:
        """Builder-defined highly-variable-gene selection (not in the reference)."""
        out = np.zeros(self.rows(), dtype=np.uint32)
        cnt = C.c_uint32()
        L.check(L.lib().sb_hvg_select(self._h, C.c_uint32(n_top), L.vp(out), C.byref(cnt)))
        return out[: cnt.value].copy()

    # ---- download
    def to_csr(self):
        """Gene-major (indptr u64, cell idx u32, count u32): to_csmat (mat.rs:207-239)."""
        return self._download(L.SB_GENE_MAJOR)

    def to_csc(self):
        return self._download(L.SB_CELL_MAJOR)

    def _download(self, major):
        m, n, _, nnz = self._shape4()
        nvec = m if major == L.SB_GENE_MAJOR else n
        indptr = np.zeros(nvec + 1, dtype=np.uint64)
        idx = np.zeros(nnz, dtype=np.uint32)
        cnt = np.zeros(nnz, dtype=np.uint32)
        L.check(L.lib().sb_download(self._h, C.c_int(major), L.vp(indptr), L.vp(idx), L.vp(cnt)))
        return indptr, idx, cnt

    def to_dense(self) -> np.ndarray:
        m, n, _, _ = self._shape4()
        indptr, idx, cnt = self.to_csr()
        out = np.zeros((m, n), dtype=np.uint32)
        for r in range(m):
            s, e = int(indptr[r]), int(indptr[r + 1])
            out[r, idx[s:e]] = cnt[s:e]
        return out

    def free(self):
        if self._h:
            for a in list(self._views):
                a.free()
            L.lib().sb_free_mat(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class LowRankOffset:
    """Device-resident normalized matrix A = map(counts) + u.v (sqz/src/low_rank_offset.rs:12-16)."""

    def __init__(self, mat: AdaptiveMat, handle):
        self.mat, self._h = mat, handle
        mat.ctx._nmats.add(self)
        mat._views.add(self)

    def rows(self) -> int:
        return self.mat.rows()

    def cols(self) -> int:
        return self.mat.cols()

    def shape(self):
        return self.mat.shape()

    def params(self):
        """(col_scale[n], row_scale[m], u[m], v[n]) as derived on the device."""
        m, n = self.shape()
        cs, rs, u, v = np.zeros(n), np.zeros(m), np.zeros(m), np.zeros(n)
        L.check(L.lib().sb_nmat_params(self._h, L.vp(cs), L.vp(rs), L.vp(u), L.vp(v)))
        return cs, rs, u, v

    def to_dense(self) -> np.ndarray:  # low_rank_offset.rs:55-57
        m, n = self.shape()
        out = np.zeros((m, n))
        L.check(L.lib().sb_nmat_to_dense(self._h, L.vp(out)))
        return out

    def dot(self, rhs: np.ndarray) -> np.ndarray:  # low_rank_offset.rs:68-81
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        m, n = self.shape()
        if rhs.shape[0] != n:
            raise ValueError("Dimension mismatch")
        out = np.zeros((m, rhs.shape[1]))
        L.check(L.lib().sb_nmat_dot(self._h, L.vp(rhs), C.c_uint32(rhs.shape[1]), L.vp(out)))
        return out

    def frobenius_sq(self) -> float:
        """||A||_F^2 of the normalized matrix, offset included (derived quantity, not in the reference)."""
        out = C.c_double()
        L.check(L.lib().sb_nmat_frobenius_sq(self._h, C.byref(out)))
        return out.value

    def rdot(self, lhs: np.ndarray) -> np.ndarray:  # low_rank_offset.rs:83-96: lhs . A
        lhs = np.ascontiguousarray(lhs, dtype=np.float64)
        m, n = self.shape()
        if lhs.shape[1] != m:
            raise ValueError("Dimension mismatch")
        out = np.zeros((lhs.shape[0], n))
        L.check(L.lib().sb_nmat_rdot(self._h, L.vp(lhs), C.c_uint32(lhs.shape[0]), L.vp(out)))
        return out

    def free(self):
        if self._h:
            L.lib().sb_free_nmat(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
