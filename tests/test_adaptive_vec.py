"""The C ABI's AdaptiveVec decoder (csrc/adaptive.cu) against the restated reference encoders (oracle/adaptive_vec.py): every
storage variant of sqz/src/vec.rs:1029-1053, forced and auto-chosen, including the fallback codes, stored zeros, empty vectors
and block boundaries; then a whole AdaptiveMat through sb_upload_adaptive."""
import numpy as np
import pytest

from oracle import adaptive_vec as av
from scan_rs_b200 import sqz


def _rand_vec(rng, length, density, vmax, zero_frac=0.0):
    nnz = int(length * density)
    idx = np.sort(rng.choice(length, size=nnz, replace=False)).astype(np.uint32)
    val = rng.integers(1, vmax + 1, size=nnz).astype(np.uint32)
    if zero_frac:
        val[rng.random(nnz) < zero_frac] = 0  # explicit zeros are legal input; iteration skips them (vec.rs:113)
    return idx, val


@pytest.mark.parametrize("variant", av.VARIANTS)
@pytest.mark.parametrize("length,density,vmax", [(1, 1.0, 3), (21, 0.5, 6), (300, 0.3, 20), (1000, 0.9, 300), (5000, 0.01, 70000), (777, 0.0, 5)])
def test_decoder_every_variant(variant, length, density, vmax):
    rng = np.random.default_rng(hash((variant, length)) % 2**32)
    idx, val = _rand_vec(rng, length, density, vmax, zero_frac=0.1)
    parts = av.encode(length, val, idx, variant)
    got_i, got_v = sqz.adaptive_decode(parts)
    keep = val != 0
    np.testing.assert_array_equal(got_i, idx[keep])
    np.testing.assert_array_equal(got_v, val[keep])


def test_choose_storage_and_boundaries():
    # the reference's own size table picks the expected classes on characteristic data
    rng = np.random.default_rng(0)
    assert av.choose_storage(10000, np.ones(9000, np.uint32)) == "D3"
    assert av.choose_storage(10000, rng.integers(8, 15, 9000).astype(np.uint32)) == "D4"
    assert av.choose_storage(10000, rng.integers(100, 250, 9000).astype(np.uint32)) == "D8"
    assert av.choose_storage(10000, rng.integers(1000, 60000, 9000).astype(np.uint32)) == "D16"
    assert av.choose_storage(100000, np.ones(3000, np.uint32)) == "S3"
    assert av.choose_storage(100000, np.full(5, 10**6, np.uint32)) == "V"
    # entries on both sides of every 256-block boundary, empty blocks in between, last position
    idx = np.array([0, 255, 256, 257, 1023, 1024, 4095, 9999], dtype=np.uint32)
    val = np.array([1, 7, 15, 255, 65535, 3, 99999, 2], dtype=np.uint32)
    for variant in av.VARIANTS:
        got_i, got_v = sqz.adaptive_decode(av.encode(10000, val, idx, variant))
        np.testing.assert_array_equal(got_i, idx)
        np.testing.assert_array_equal(got_v, val)


def test_malformed_vectors_are_rejected():
    import scan_rs_b200 as sb
    p = av.encode(1000, np.array([1, 2], np.uint32), np.array([5, 700], np.uint32), "S4")
    p.block_starts = p.block_starts[:1]
    with pytest.raises(sb.ScanB200Error):
        sqz.adaptive_decode(p)
    v = av.encode(100, np.array([1, 2], np.uint32), np.array([50, 10], np.uint32), "V")  # indexes not ascending
    with pytest.raises(sb.ScanB200Error):
        sqz.adaptive_decode(v)


@pytest.mark.gpu
def test_upload_adaptive_matches_plain_upload():
    import scan_rs_b200 as sb
    from scan_rs_b200.synth import SynthConfig, generate_host
    from oracle import oracle as orc
    cfg = SynthConfig(n_cells=1500, n_genes=900, seed=5, n_dense=20, dense_mean=400.0)
    ip, g, c = generate_host(cfg)
    cm = orc.CountMatrix.from_cell_major(cfg.n_genes, cfg.n_cells, ip, g, c)  # gene-major CSR, like the reference stores it
    vecs = [av.encode(cfg.n_cells, cm.val[cm.indptr[r]:cm.indptr[r + 1]], cm.idx[cm.indptr[r]:cm.indptr[r + 1]]) for r in range(cfg.n_genes)]
    assert len({v.variant for v in vecs}) >= 3  # the synthetic matrix exercises several storage classes
    with sb.Context(0) as ctx:
        a = sb.AdaptiveMat.from_adaptive(ctx, cfg.n_genes, cfg.n_cells, vecs, major="gene")
        b = sb.AdaptiveMat.from_csr(ctx, cfg.n_genes, cfg.n_cells, cm.indptr, cm.idx, cm.val)
        for x, y in zip(a.to_csr(), b.to_csr()):
            np.testing.assert_array_equal(x, y)
        np.testing.assert_array_equal(a.sum_axis_u32(0), cm.sum_axis_u32(0))
