"""World-size-2 gloo tests (CPU) of the N > 1 path's host logic: shard planning, and the sharded
algebra the CUDA path relies on -- A^T.Y local, A.X / moments / Gram all-reduced -- checked with the
oracle as the per-shard operator against the unsharded oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds():
    from scan_rs_b200.dist import shard_bounds, shard_bounds_by_nnz
    spans = [shard_bounds(1_300_000, 8, r) for r in range(8)]
    assert spans[0][0] == 0 and spans[-1][1] == 1_300_000
    assert all(spans[i][1] == spans[i + 1][0] for i in range(7))
    assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    nnz = np.r_[np.full(100, 10), np.full(100, 1000)]
    by = shard_bounds_by_nnz(nnz, 4)
    assert by[0][0] == 0 and by[-1][1] == 200 and all(by[i][1] == by[i + 1][0] for i in range(3))
    loads = [int(nnz[a:b].sum()) for a, b in by]
    assert max(loads) - min(loads) <= 1000


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as orc
    from scan_rs_b200.dist import shard_bounds
    from scan_rs_b200.synth import SynthConfig, generate_host
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    cfg = SynthConfig(n_cells=900, n_genes=400, seed=8)
    ip, g, c = generate_host(cfg)
    full = orc.CountMatrix.from_cell_major(400, 900, ip, g, c)
    a_full = orc.normalize(full, orc.CELLRANGER)
    lo, hi = shard_bounds(900, world, rank)
    s, e = int(ip[lo]), int(ip[hi])
    shard = orc.CountMatrix.from_cell_major(400, hi - lo, ip[lo:hi + 1] - ip[lo], g[s:e], c[s:e])

    def allreduce(x):
        t = torch.from_numpy(np.ascontiguousarray(x))
        dist.all_reduce(t)
        return t.numpy()

    # normalization: local totals, global median, all-reduced moments
    tot = shard.sum_axis_u32(0)
    parts = [None] * world
    dist.all_gather_object(parts, tot)
    med = orc.median_mut(np.concatenate(parts))
    assert med == orc.median_mut(full.sum_axis_u32(0))
    cs = max(float(med), 1.0) / tot.astype(np.float64)
    mm = orc.MappedMatrix(shard, orc.MapSpec(kind=1, log_base=2, col_scale=cs))
    s1 = allreduce(mm.sum_axis(1))
    s2 = allreduce(mm.sum_axis(1, square=True))
    mean, sq = s1 / 900.0, s2 / 900.0
    var = sq - mean * mean
    sd = np.where(var <= 0, 1.0, np.sqrt(np.where(var <= 0, 1.0, var)))
    np.testing.assert_allclose(1.0 / sd, a_full.mat.spec.row_scale, rtol=1e-11)
    a_loc = orc.LowRankOffset(orc.MappedMatrix(shard, orc.MapSpec(kind=1, log_base=2, col_scale=cs, row_scale=1.0 / sd)),
                              (-(mean / sd)).reshape(-1, 1), np.ones((1, hi - lo)))
    rng = np.random.default_rng(0)
    y = rng.standard_normal((6, 400))
    t_loc = a_loc.rdot(y)                                     # shard-local block of B . A
    np.testing.assert_allclose(t_loc, a_full.rdot(y)[:, lo:hi], rtol=1e-9, atol=1e-9)
    p = allreduce(a_loc.mat.dot(t_loc.T) + a_loc.u.dot(a_loc.v.dot(t_loc.T)))   # A . T: partials + offset, all-reduced
    np.testing.assert_allclose(p, a_full.dot(a_full.rdot(y).T), rtol=1e-9, atol=1e-7)
    gram = allreduce(t_loc.dot(t_loc.T))
    np.testing.assert_allclose(gram, a_full.rdot(y).dot(a_full.rdot(y).T), rtol=1e-9, atol=1e-6)
    dist.barrier()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_sharded_algebra_gloo(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
