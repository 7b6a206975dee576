"""The minimal HDF5 container reader / writer behind the 10x matrix loader (scan_rs_b200/h5.py): round trips on the CPU."""
import numpy as np
import pytest

from scan_rs_b200 import h5


def _tree(rng, n_cells=50, n_genes=40):
    nnz_per = rng.integers(0, 12, size=n_cells)
    indptr = np.concatenate([[0], np.cumsum(nnz_per)]).astype(np.int64)
    indices = np.concatenate([rng.permutation(n_genes)[:k] for k in nnz_per] + [np.zeros(0, dtype=np.int64)]).astype(np.int64)
    data = rng.integers(1, 300, size=indices.shape[0]).astype(np.int32)
    return {"matrix": {"shape": np.array([n_genes, n_cells], dtype=np.int32), "indptr": indptr, "indices": indices, "data": data,
                       "barcodes": np.array([f"BC{i:05d}-1".encode() for i in range(n_cells)], dtype="S18"),
                       "features": {"id": np.array([f"ENSG{i:08d}".encode() for i in range(n_genes)], dtype="S16"),
                                    "name": np.array([f"gene{i}".encode() for i in range(n_genes)], dtype="S12"),
                                    "feature_type": np.array([b"Gene Expression" if i % 5 else b"Antibody Capture" for i in range(n_genes)], dtype="S20")}}}


@pytest.mark.parametrize("compress", [True, False])
def test_write_read_round_trip(tmp_path, compress):
    rng = np.random.default_rng(1)
    tree = _tree(rng)
    tree["floats"] = rng.standard_normal(1000)
    tree["f32"] = rng.standard_normal(7).astype(np.float32)
    tree["u8"] = rng.integers(0, 255, size=10000).astype(np.uint8)
    tree["empty"] = np.zeros(0, dtype=np.int64)
    path = str(tmp_path / "m.h5")
    h5.write_h5(path, tree, chunk=300, compress=compress)
    f = h5.H5File(path)
    assert f.keys("/") == sorted(tree) and f.keys("matrix") == sorted(tree["matrix"]) and f.keys("matrix/features") == ["feature_type", "id", "name"]
    assert "matrix/features/id" in f and "matrix/nope" not in f

    def check(prefix, node):
        for k, v in node.items():
            if isinstance(v, dict):
                check(prefix + k + "/", v)
            else:
                got = f[prefix + k]
                assert got.dtype == np.asarray(v).dtype.newbyteorder("<") or got.dtype == np.asarray(v).dtype
                np.testing.assert_array_equal(got, v)
    check("", tree)
    m, n, indptr, indices, data, barcodes, feats = h5.read_csc_arrays(path)
    assert (m, n) == (40, 50) and indptr.dtype == np.uint64 and indices.dtype == np.uint32 and data.dtype == np.uint32
    assert barcodes[3] == b"BC00003-1" and feats["feature_type"][5] == b"Antibody Capture"


def test_rejects_other_formats(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not hdf5" * 100)
    with pytest.raises(h5.H5FormatError, match="signature"):
        h5.H5File(str(p))
    p.write_bytes(h5.SIG + bytes([2]) + b"\0" * 200)
    with pytest.raises(h5.H5FormatError, match="superblock version 2"):
        h5.H5File(str(p))
