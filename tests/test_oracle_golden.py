"""Pin the CPU oracle against the reference's own inline golden vectors."""
import numpy as np
import pytest

from oracle import oracle as orc

DENSE_A = np.array([[136, 936, 0, 0, 264],
                    [134, 682, 417, 8, 391],
                    [0, 133, 780, 0, 0],
                    [396, 76, 96, 198, 0]], dtype=np.uint32)


def test_rng_known_answers():
    # xoshiro256++ reference vector, state [1,2,3,4]
    r = orc.Xoshiro256PlusPlus([1, 2, 3, 4])
    assert [r.next_u64() for _ in range(3)] == [41943041, 58720359, 3588806011781223]
    r = orc.Xoshiro256PlusPlus.seed_from_u64(0)
    assert [r.next_u64() for _ in range(4)] == [5987356902031041503, 7051070477665621255,
                                                6633766593972829180, 211316841551650330]
    om = orc.omega(0, (2, 2)).ravel()
    np.testing.assert_array_equal(om, [-0.35084946393718663, -0.23552140697665314,
                                       -0.2807655847052897, -0.977088982130693])


def test_median_mut():  # scan-rs/src/stats.rs:67-82
    assert orc.median_mut(np.array([1, 10], dtype=np.uint32)) == 5
    assert orc.median_mut(np.array([1, 10, 100, 1000], dtype=np.uint32)) == 55
    assert orc.median_mut(np.array([3, 1, 2], dtype=np.uint32)) == 2
    assert orc.median_mut(np.zeros(0, dtype=np.uint32)) is None


def test_cellranger_normalisation():  # normalization.rs:539-575
    expected = np.array([[0.61392149, 0.95459951, -1.21707302, -1.21707302, 0.86562504],
                         [-0.11878431, 0.54279925, 0.38607315, -1.85660965, 1.04652156],
                         [-0.78758751, 0.76437149, 1.59839105, -0.78758751, -0.78758751],
                         [0.88718256, -0.25584717, -0.01048423, 1.09574143, -1.71659259]])
    m = orc.CountMatrix.from_dense(DENSE_A)
    out = orc.normalize_with_size_factor(m, orc.CELLRANGER, None).to_dense()
    assert np.abs(out - expected).max() < 1e-6
    out2 = orc.normalize(m, orc.CELLRANGER).to_dense()
    np.testing.assert_array_equal(out, out2)


def test_cellranger8_normalisation():  # normalization.rs:577-612
    expected = np.array([[2.37992764, 3.70059981, -4.71810445, -4.71810445, 3.35568145],
                         [-0.15920674, 0.72751443, 0.51745426, -2.48841594, 1.40265399],
                         [-2.85652852, 2.77232551, 5.79726005, -2.85652852, -2.85652852],
                         [2.94151467, -0.84827885, -0.0347612, 3.63300591, -5.69148053]])
    m = orc.CountMatrix.from_dense(DENSE_A)
    out = orc.normalize_with_size_factor(m, orc.CELLRANGER8, None).to_dense()
    assert np.abs(out - expected).max() < 1e-6


def test_log_normalize_with_size_factor():  # normalization.rs:614-650
    expected = np.array([[9.37098961, 9.18882221, 0., 0., 9.37609671],
                         [9.34964848, 8.73300582, 8.4781546, 12.37964912, 9.94202202],
                         [0., 6.3885887, 9.3796973, 0., 0.],
                         [10.91145213, 5.59409085, 6.37267837, 17.00874593, 0.]])
    m = orc.CountMatrix.from_dense(DENSE_A)
    size_factors = 1 + m.select_rows([0, 2]).sum_axis_u32(0)
    out = orc.log_normalize_with_size_factor(m, None, orc.LOG_TWO, size_factors).to_dense()
    assert np.abs(out - expected).max() < 1e-6


def test_vanilla_log_norm():  # normalization.rs:652-685
    expected = np.array([[0.50075509, 1.16407001, -1.1965938, -1.1965938, 0.72836249],
                         [-0.14245194, 0.89844192, 0.58318993, -1.88113806, 0.54195815],
                         [-0.80111703, 0.89623633, 1.50711477, -0.80111703, -0.80111703],
                         [0.92609909, 0.14507504, 0.25503138, 0.59722303, -1.92342854]])
    m = orc.CountMatrix.from_dense(DENSE_A)
    out = orc.normalize_with_size_factor(m, orc.LOG_TRANSFORM, None).to_dense()
    assert np.abs(out - expected).max() < 1e-6
    # fixed point, base 10 exponent 1 (normalization.rs:687-722): same expected matrix
    m10 = orc.CountMatrix.from_dense(DENSE_A * 10)
    out = orc.log1p_normalize_fixed_point(m10, orc.LOG_TWO, 10, 1).to_dense()
    assert np.abs(out - expected).max() < 1e-6


def test_one_dim_no_nan():  # normalization.rs:477-516: 1 x 649 matrix -> zero variance row -> sd := 1
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "one_dim_649.txt")
    row = np.loadtxt(path, dtype=np.uint32).reshape(1, 649)
    m = orc.CountMatrix.from_dense(row)
    out = orc.normalize(m, orc.CELLRANGER).to_dense()
    assert not np.isnan(out).any()


def test_fit_multinomial_model():  # normalization.rs:463-475
    m = orc.CountMatrix.from_dense(np.array([[1, 0, 2], [0, 0, 0], [3, 0, 6]], dtype=np.uint32))
    n, pi = orc.fit_multinomial_model(m)
    np.testing.assert_allclose(n, [4.0, 0.0, 8.0], rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(pi, [0.25, 0.0, 0.75], rtol=1e-7, atol=1e-12)


INPUT_A = np.array([[136, 936, 0, 0, 264],
                    [134, 682, 417, 8, 391],
                    [0, 133, 780, 885, 0],
                    [396, 76, 96, 198, 0]], dtype=np.uint32)


def test_sum_fns():  # sqz/src/mat.rs:1302-1325
    m = orc.CountMatrix.from_dense(INPUT_A)
    np.testing.assert_array_equal(m.sum_axis_u32(0), [666, 1827, 1293, 1091, 655])
    np.testing.assert_array_equal(m.sum_axis_u32(1), [1336, 1632, 1798, 766])
    np.testing.assert_array_equal(m.sum_axis_u64(1), [1336, 1632, 1798, 766])


def test_mean_var_fns():  # sqz/src/mat.rs:1327-1370
    mm = orc.MappedMatrix(orc.CountMatrix.from_dense(INPUT_A), orc.MapSpec(kind=0))
    np.testing.assert_allclose(mm.mean_axis(0), [166.5, 456.75, 323.25, 272.75, 163.75], atol=1e-7)
    np.testing.assert_allclose(mm.mean_axis(1), [267.2, 326.4, 359.6, 153.2], atol=1e-7)
    var0 = mm.mean_axis(0, square=True) - mm.mean_axis(0) ** 2
    var1 = mm.mean_axis(1, square=True) - mm.mean_axis(1) ** 2
    np.testing.assert_allclose(var0, [20594.75, 132550.6875, 93385.6875, 131230.6875, 28830.1875], atol=1e-7)
    np.testing.assert_allclose(var1, [121461.76, 55445.84, 152550.64, 18732.16], atol=1e-7)


def test_center_rows():  # sqz/src/mat.rs:1259-1291 (Axis(1) centring of [[1,2,3],[2,3,4],[3,4,5]])
    m = orc.CountMatrix.from_dense(np.array([[1, 2, 3], [2, 3, 4], [3, 4, 5]], dtype=np.uint32))
    mm = orc.MappedMatrix(m, orc.MapSpec(kind=1))
    out = mm.scale_and_center(np.ones(3)).to_dense()
    np.testing.assert_allclose(out, [[-1, 0, 1]] * 3, rtol=1e-7, atol=1e-12)


def test_spmm_exact_vs_dense():  # recipe of sqz/src/mat.rs:1406-1486: integer data => exact equality
    rng = np.random.default_rng(1)
    for rows, cols, w in [(7, 5, 3), (40, 63, 16), (1, 9, 1), (33, 2, 20), (0, 4, 2)]:
        dense = (rng.integers(0, 50, size=(rows, cols)) * (rng.random((rows, cols)) < 0.3)).astype(np.uint32)
        mm = orc.MappedMatrix(orc.CountMatrix.from_dense(dense), orc.MapSpec(kind=0))
        x = rng.integers(0, 100, size=(cols, w)).astype(np.float64)
        np.testing.assert_array_equal(mm.dot(x), dense.astype(np.float64).dot(x))
        y = rng.integers(0, 100, size=(w, rows)).astype(np.float64)
        np.testing.assert_array_equal(mm.rdot(y), y.dot(dense.astype(np.float64)))
        np.testing.assert_array_equal(mm.dot(x, threads=True), dense.astype(np.float64).dot(x))
        np.testing.assert_array_equal(mm.rdot(y, threads=True), y.dot(dense.astype(np.float64)))


def test_low_rank_offset_dot():  # sqz/src/low_rank_offset.rs:144-173 (rtol 1e-7 / atol 1e-12)
    rng = np.random.default_rng(2)
    dense = (rng.integers(1, 30, size=(30, 17)) * (rng.random((30, 17)) < 0.4)).astype(np.uint32)
    a = orc.normalize(orc.CountMatrix.from_dense(dense), orc.CELLRANGER)
    ad = a.to_dense()
    x = rng.standard_normal((17, 6))
    y = rng.standard_normal((4, 30))
    np.testing.assert_allclose(a.dot(x), ad.dot(x), rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(a.rdot(y), y.dot(ad), rtol=1e-7, atol=1e-12)


def test_partition_matches_dense_fixpoint():  # recipe of sqz/src/mat.rs:1488-1562
    rng = np.random.default_rng(3)
    dense = (rng.integers(1, 4, size=(40, 30)) * (rng.random((40, 30)) < 0.08)).astype(np.uint32)
    m = orc.CountMatrix.from_dense(dense)
    filt, resid, rows, cols = m.partition_on_threshold(3.0)
    d = dense.astype(np.float64)
    er, ec = np.zeros(40, bool), np.zeros(30, bool)
    while True:
        upd = False
        s = d[~er][:, :].sum(axis=0) * (~ec)
        new = (s < 3.0) & ~ec
        upd |= new.any(); ec |= new
        s = (d[:, ~ec].sum(axis=1)) * (~er)
        new = (s < 3.0) & ~er
        upd |= new.any(); er |= new
        if not upd:
            break
    np.testing.assert_array_equal(rows, np.nonzero(~er)[0])
    np.testing.assert_array_equal(cols, np.nonzero(~ec)[0])
    np.testing.assert_array_equal(filt.to_dense(), dense[~er][:, ~ec])
    np.testing.assert_array_equal(resid.to_dense(), dense[~er][:, ec])


def simple_deterministic_ex(m, n):  # scan-rs/src/dim_red/test.rs:168-176
    x = np.arange(m * n, dtype=np.int64)
    return (x % 7 + x % 4 + x % 50 + x % 47 + x % 12).astype(np.float64).reshape(m, n)


@pytest.mark.parametrize("shape", [(100, 1000), (1000, 100)])
@pytest.mark.parametrize("algo", ["bk", "rand"])
def test_svd_vs_full(shape, algo):  # scan-rs/src/dim_red/test.rs:58-110, three metrics < 1e-3
    a = simple_deterministic_ex(*shape)
    op = orc.DenseOp(a)
    pca = orc.BkSvd() if algo == "bk" else orc.RandSvd()
    u, s, v = pca.run_pca(op, 10)
    assert u.shape == (shape[0], 10) and v.shape == (shape[1], 10)
    ut, st, vt = np.linalg.svd(a, full_matrices=False)
    av = a.dot(v)
    assert orc.frobenius(av - u * s) < 1e-3
    assert np.abs((s - st[:10]) / st[:10]).max() < 1e-3
    av_gt = np.abs(a.dot(vt[:10].T))
    assert np.abs((np.abs(av) - av_gt) / av_gt).max() < 1e-3


@pytest.mark.parametrize("shape", [(100, 1000), (1000, 100)])
def test_irlba_vs_full(shape):  # scan-rs/src/dim_red/test.rs:133-139 (Irlba { tol: 1e-5, .. }, nu = 10), metrics of :58-110
    a = simple_deterministic_ex(*shape)
    u, s, v, mprod, it = orc.irlba(orc.DenseOp(a), 10, tol=1e-5, maxit=50)
    assert u.shape == (shape[0], 10) and v.shape == (shape[1], 10) and mprod > 0
    ut, st, vt = np.linalg.svd(a, full_matrices=False)
    av = a.dot(v)
    assert orc.frobenius(av - u * s) < 1e-3
    assert np.abs((s - st[:10]) / st[:10]).max() < 1e-3
    av_gt = np.abs(a.dot(vt[:10].T))
    assert np.abs((np.abs(av) - av_gt) / av_gt).max() < 1e-3
    with pytest.raises(ValueError, match="at least 2x2"):
        orc.irlba(orc.DenseOp(np.ones((1, 5))), 1)
    with pytest.raises(ValueError, match="invalid k"):
        orc.irlba(orc.DenseOp(np.ones((3, 5))), 4)


def test_moment_consumers():  # sqz/src/mat.rs:1302-1370 (input_a), diff-exp/src/stat.rs:172-178 (median 3.5)
    m = orc.CountMatrix.from_dense(INPUT_A)
    mean0, var0 = orc.mean_var_axis(m, 0)
    np.testing.assert_allclose(mean0, [166.5, 456.75, 323.25, 272.75, 163.75], atol=1e-7)
    np.testing.assert_allclose(var0, [20594.75, 132550.6875, 93385.6875, 131230.6875, 28830.1875], atol=1e-7)
    mean1, var1 = orc.mean_var_axis(m, 1)
    np.testing.assert_allclose(mean1, [267.2, 326.4, 359.6, 153.2], atol=1e-7)
    np.testing.assert_allclose(var1, [121461.76, 55445.84, 152550.64, 18732.16], atol=1e-7)
    densef = INPUT_A[:, 1:4].astype(np.float64)
    meanc, varc = orc.mean_var_rows(m, [1, 2, 3])
    np.testing.assert_allclose(meanc, densef.mean(axis=1), atol=1e-7)
    np.testing.assert_allclose(varc, densef.var(axis=1), atol=1e-7)
    s1, s2 = orc.sum_rows_dual(m, [1, 2, 3], [2, 3, 4])
    np.testing.assert_array_equal(s1, INPUT_A[:, 1:4].sum(axis=1))
    np.testing.assert_array_equal(s2, INPUT_A[:, 2:5].sum(axis=1))
    assert orc.percentile_median(np.array([1, 2, 4, 3, 5, 6])) == 3.5
    sf = orc.size_factors(m)
    np.testing.assert_allclose(sf, INPUT_A.sum(axis=0) / 1091.0, rtol=1e-15)
    sf_sub = orc.size_factors(m, cell_indices=[0, 3])
    np.testing.assert_allclose(sf_sub, [666 / 878.5, 0, 0, 1091 / 878.5, 0], rtol=1e-15)


def test_svd_bk_errors_and_cancel():  # bk_svd.rs:73-79, :96
    with pytest.raises(ValueError, match="at least 2x2"):
        orc.svd_bk(orc.DenseOp(np.ones((1, 5))), 1, 2, 5)
    with pytest.raises(ValueError, match="invalid k"):
        orc.svd_bk(orc.DenseOp(np.ones((3, 5))), 4, 8, 5)
    seen = []
    with pytest.raises(orc.CancellationError):
        orc.svd_bk(orc.DenseOp(simple_deterministic_ex(30, 50)), 3, 6, 5, snoop=lambda f: seen.append(f) or len(seen) >= 2)
    assert seen == [0.0, 0.16000000000000003]
    seen = []
    orc.svd_bk(orc.DenseOp(simple_deterministic_ex(30, 50)), 3, 6, 5, snoop=lambda f: seen.append(f) and False)
    assert seen[-3:] == [0.82, 0.93, 1.0] and len(seen) == 8


def test_binomial_maps_dense():  # normalization.rs:233-354: sparse + u.v equals the residual matrix
    rng = np.random.default_rng(4)
    dense = (rng.integers(1, 9, size=(12, 9)) * (rng.random((12, 9)) < 0.5)).astype(np.uint32)
    dense[:, 0] += 1
    dense[0, :] += 1
    y = dense.astype(np.float64)
    n = y.sum(axis=0)
    pi = y.sum(axis=1) / y.sum()
    mu = np.outer(pi, n)
    pear = (y - mu) / np.sqrt(mu * (1 - pi)[:, None])
    out = orc.binom_pearson_resid(orc.CountMatrix.from_dense(dense)).to_dense()
    np.testing.assert_allclose(out, pear, rtol=1e-9, atol=1e-9)
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = np.where(y > 0, y * np.log(y / mu), 0.0)
        t2 = np.where(n - y > 0, (n - y) * np.log((n - y) / (n - mu)), 0.0)
    dev = np.sign(y - mu) * np.sqrt(np.maximum(2 * (t1 + t2), 0))
    out = orc.binom_deviance_resid(orc.CountMatrix.from_dense(dense)).to_dense()
    np.testing.assert_allclose(out, dev, rtol=1e-8, atol=1e-8)
