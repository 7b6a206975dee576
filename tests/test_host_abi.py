"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the header
declares, fails loudly without a GPU, and its host-side pieces (the start-block RNG, the
Normalization parser) give the reference's known answers.  No compute calls (no GPU here)."""
import os
import re
import subprocess

import numpy as np
import pytest

import scan_rs_b200 as sb
from scan_rs_b200 import _lib as L
from tests.conftest import HAVE_GPU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "scanb200.h")).read()
    declared = set(re.findall(r"SB_API\s+[\w\s\*]+?\b(sb_\w+)\s*\(", hdr))
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (sb_\w+)", out))
    assert declared <= exported, declared - exported
    lib = L.lib()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.sb_version() == 100


def test_only_sm100a_code_is_embedded():
    out = subprocess.run(["cuobjdump", "-lelf", L.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(HAVE_GPU, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(sb.ScanB200Error) as ei:
        sb.Context(0)
    assert ei.value.code == L.SB_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_omega_known_answers():
    om = sb.omega(0, 2, 2).ravel()
    np.testing.assert_array_equal(om, [-0.35084946393718663, -0.23552140697665314,
                                       -0.2807655847052897, -0.977088982130693])
    from oracle import oracle as orc
    np.testing.assert_array_equal(sb.omega(7, 13, 5), orc.omega(7, (13, 5)))
    np.testing.assert_array_equal(sb.omega(0, 20, 40), orc.omega(0, (20, 40)))


def test_normalization_from_str():  # normalization.rs:30-43
    assert sb.Normalization.from_str("cellranger") == sb.Normalization.CellRanger
    assert sb.Normalization.from_str("seuratlog") == sb.Normalization.SeuratLog
    assert sb.Normalization.from_str("binomialpearson") == sb.Normalization.BinomialPearson
    with pytest.raises(ValueError, match="Normalization not recognized: foo"):
        sb.Normalization.from_str("foo")


def test_synth_cpu_is_deterministic_and_shardable():
    from scan_rs_b200.synth import SynthConfig, generate_host
    cfg = SynthConfig(n_cells=300, n_genes=2000, seed=3)
    ip, g, c = generate_host(cfg)
    ip2, g2, c2 = generate_host(cfg)
    np.testing.assert_array_equal(g, g2)
    np.testing.assert_array_equal(c, c2)
    # a shard generated on its own equals the slice of the whole
    ipa, ga, ca = generate_host(cfg, 100, 250)
    s, e = int(ip[100]), int(ip[250])
    np.testing.assert_array_equal(ga, g[s:e])
    np.testing.assert_array_equal(ca, c[s:e])
    assert (c > 0).all() and (np.diff(ip.astype(np.int64)) > 0).all()


def test_compact_host_form_round_trip():
    """AdaptiveMat.compact_csc (the narrow host form sb_upload_compact takes): u16 index + u8 count + side list of the
    counts >= 255 reproduces the u32 arrays; indices past 65535 are refused."""
    import numpy as np
    import pytest
    from scan_rs_b200.sqz import AdaptiveMat
    rng = np.random.default_rng(0)
    idx = rng.integers(0, 65536, 5000).astype(np.uint32)
    val = rng.integers(1, 40, 5000).astype(np.uint32)
    val[rng.integers(0, 5000, 60)] = rng.integers(255, 2**31, 60)
    val[7] = 255
    val[8] = 254
    idx16, cnt8, big_pos, big_cnt = AdaptiveMat.compact_csc(idx, val)
    assert idx16.dtype == np.uint16 and cnt8.dtype == np.uint8 and big_pos.dtype == np.uint64 and big_cnt.dtype == np.uint32
    back = cnt8.astype(np.uint32)
    assert (back[big_pos.astype(np.int64)] == 255).all() and (np.diff(big_pos.astype(np.int64)) > 0).all()
    back[big_pos.astype(np.int64)] = big_cnt
    np.testing.assert_array_equal(back, val)
    np.testing.assert_array_equal(idx16.astype(np.uint32), idx)
    assert 7 in big_pos and 8 not in big_pos
    with pytest.raises(ValueError):
        AdaptiveMat.compact_csc(np.array([70000], dtype=np.uint32), np.array([1], dtype=np.uint32))


def _unpack_packed(indptr, dgene, cnt4, esc_pos, esc_gene, big_pos, big_cnt):
    """numpy decoder of the packed host form as include/scanb200.h states it (independent of the library's kernels)."""
    import numpy as np
    nnz = int(indptr[-1])
    val = np.where(np.arange(nnz) % 2 == 0, cnt4[np.arange(nnz) // 2] & 15, cnt4[np.arange(nnz) // 2] >> 4).astype(np.uint32)
    assert (val[big_pos.astype(np.int64)] == 15).all()
    val[big_pos.astype(np.int64)] = big_cnt
    esc = dict(zip(esc_pos.tolist(), esc_gene.tolist()))
    idx = np.zeros(nnz, dtype=np.uint32)
    for c in range(len(indptr) - 1):
        prev = -1
        for k in range(int(indptr[c]), int(indptr[c + 1])):
            prev = esc[k] if dgene[k] == 0 else prev + int(dgene[k])
            idx[k] = prev
    assert set(esc) == set(np.flatnonzero(dgene == 0).tolist())
    return idx, val


def test_packed_host_form_round_trip():
    """sb_pack_csc_count / sb_pack_csc_fill (host threads, no device): gene deltas + count nibbles + the two side lists decode
    back to the u32 arrays -- first genes past 254, gaps past 255, counts of 14 / 15 / large, empty cells, odd thread cuts."""
    import numpy as np
    import pytest
    from scan_rs_b200 import _lib as L
    from scan_rs_b200.sqz import AdaptiveMat
    rng = np.random.default_rng(5)
    m, n = 40000, 300
    ip, idx, val = [0], [], []
    for c in range(n):
        k = 0 if c % 37 == 5 else int(rng.integers(1, 90))
        g = np.sort(rng.choice(m if c % 3 else 600, size=k, replace=False))
        idx.extend(g.tolist())
        val.extend(rng.choice([1, 1, 1, 2, 3, 14, 15, 16, 300, 70000], size=k).tolist())
        ip.append(len(idx))
    ip, idx, val = np.array(ip, dtype=np.uint64), np.array(idx, dtype=np.uint32), np.array(val, dtype=np.uint32)
    for threads in (1, 3, 7):
        dgene, cnt4, esc_pos, esc_gene, big_pos, big_cnt = AdaptiveMat.pack_csc(ip, idx, val, threads=threads)
        assert dgene.shape == idx.shape and cnt4.shape == ((len(idx) + 1) // 2,)
        assert (np.diff(esc_pos.astype(np.int64)) > 0).all() and (np.diff(big_pos.astype(np.int64)) > 0).all()
        assert big_pos.size == int((val >= 15).sum()) and esc_pos.size > n // 2
        bi, bv = _unpack_packed(ip, dgene, cnt4, esc_pos, esc_gene, big_pos, big_cnt)
        np.testing.assert_array_equal(bi, idx)
        np.testing.assert_array_equal(bv, val)
    with pytest.raises(L.ScanB200Error, match="ascending"):
        bad = idx.copy()
        s0 = int(ip[1])
        bad[s0 + 1] = bad[s0]
        AdaptiveMat.pack_csc(ip, bad, val)
    e = AdaptiveMat.pack_csc(np.zeros(4, dtype=np.uint64), np.zeros(0, dtype=np.uint32), np.zeros(0, dtype=np.uint32))
    assert all(a.size == 0 for a in e)
    # more threads than cells, a single cell with an odd number of entries, a lone entry at the last representable delta
    one = AdaptiveMat.pack_csc(np.array([0, 3], dtype=np.uint64), np.array([254, 509, 765], dtype=np.uint32), np.array([14, 15, 1], dtype=np.uint32), threads=64)
    assert one[0].tolist() == [255, 255, 0] and one[1].tolist() == [0xFE, 0x01] and one[2].tolist() == [2] and one[3].tolist() == [765]
    assert one[4].tolist() == [1] and one[5].tolist() == [15]
    bi, bv = _unpack_packed(np.array([0, 3], dtype=np.uint64), *one)
    assert bi.tolist() == [254, 509, 765] and bv.tolist() == [14, 15, 1]


def test_gather_work_units_cover_every_entry_once():
    """scan_rs_b200/csrc/gather_units.h (how the sparse streams are cut into work units and handed to CTAs) fuzzed on the
    CPU: exact tiling of the stream, units inside their panel / segment, sweep order, balanced CTA loads."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "tests", "cpp", "gather_units_test.cpp")
    exe = os.path.join(root, "tests", "cpp", "gather_units_test")
    hdr = os.path.join(root, "scan_rs_b200", "csrc", "gather_units.h")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O1", src, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ALL PASSED" in out.stdout, out.stdout + out.stderr


def test_host_eigensolver_known_spectra():
    """scan_rs_b200/csrc/eig_host.h (k largest eigenpairs of the small Gram matrix: tridiagonalisation, QL, inverse iteration, check
    against the original matrix) on matrices with a known spectrum: separated, decaying over ten decades, repeated / clustered,
    rank-deficient, orders 1-3 and 128, indefinite; non-finite input is declined."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "tests", "cpp", "eig_host_test.cpp")
    exe = os.path.join(root, "tests", "cpp", "eig_host_test")
    hdr = os.path.join(root, "scan_rs_b200", "csrc", "eig_host.h")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", src, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ALL PASSED" in out.stdout, out.stdout + out.stderr


def test_rust_ffi_matches_header():
    """rust/scan-b200/src/ffi.rs is generated from include/scanb200.h (scripts/gen_rust_ffi.py): every SB_API symbol is declared
    there, the committed file is current, and every declared symbol is exported by the library and listed for ctypes."""
    import importlib.util
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(root, "scripts", "gen_rust_ffi.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    decls = gen.declarations(open(os.path.join(root, "include", "scanb200.h")).read())
    committed = open(os.path.join(root, "rust", "scan-b200", "src", "ffi.rs")).read()
    assert committed == gen.render(decls), "run python scripts/gen_rust_ffi.py"
    names = [d[0] for d in decls]
    assert sorted(names) == sorted(set(names)) and sorted(names) == sorted(L.SYMBOLS)
    assert sorted(re.findall(r"pub fn (sb_\w+)\(", committed)) == sorted(names)
    for crate_file in ("Cargo.toml", "build.rs", os.path.join("src", "lib.rs")):
        assert os.path.exists(os.path.join(root, "rust", "scan-b200", crate_file))
