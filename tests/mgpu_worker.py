"""Worker of the multi-GPU parity test: launched by torchrun with one rank per GPU.  Two cases, both against the CPU oracle run on
the unsharded matrix by rank 0: a small matrix generated on the device per shard, and a larger one (33,538 genes: plane levels,
several cell tiles per rank) uploaded per shard through the narrow host forms (sb_upload_compact, sb_upload_packed: the bench's end-to-end path)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import scan_rs_b200 as sb
    from oracle import oracle as orc
    from scan_rs_b200.dist import gather_rows, init_context, shard_bounds, shard_bounds_by_nnz
    from scan_rs_b200.synth import SynthConfig, generate_device, generate_host
    from tests.util import check_pca_parity
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = init_context(local_rank)
    orc.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    for case, cfg, k, compact in (("device-generated", SynthConfig(n_cells=6000, n_genes=1500, seed=41), 10, False),
                                  ("compact-upload", SynthConfig(n_cells=48000, n_genes=33538, seed=43), 10, True),
                                  ("k30-dense-features-packed-upload", SynthConfig(n_cells=12000, n_genes=2600, seed=45, n_dense=200, sigma_g=3.0), 30, "packed")):
        ip, g, c = generate_host(cfg)  # every rank generates the whole (small) matrix on the host and keeps its shard
        if compact:  # nnz-balanced contiguous shards, as bench.py cuts them
            bounds = shard_bounds_by_nnz(np.diff(ip.astype(np.int64)), world)
            lo, hi = bounds[rank]
            s0, s1 = int(ip[lo]), int(ip[hi])
            ip_loc = (ip[lo:hi + 1] - ip[lo]).astype(np.uint64)
            if compact == "packed":  # sb_upload_packed, the bench's default end-to-end path
                dm = sb.AdaptiveMat.from_csc_packed(ctx, cfg.n_genes, hi - lo, ip_loc, *sb.AdaptiveMat.pack_csc(ip_loc, g[s0:s1], c[s0:s1]))
            else:
                g16, c8, bpos, bcnt = sb.AdaptiveMat.compact_csc(g[s0:s1], c[s0:s1])
                dm = sb.AdaptiveMat.from_csc_compact(ctx, cfg.n_genes, hi - lo, ip_loc, g16, c8, bpos, bcnt)
        else:
            lo, hi = shard_bounds(cfg.n_cells, world, rank)
            dm = generate_device(ctx, cfg, lo, hi)
        assert dm.cols() == hi - lo and dm.cols_global() == cfg.n_cells
        gene_tot = dm.gene_totals()
        med = dm.median_cell_total()
        a = sb.normalize(dm, sb.Normalization.CellRanger)
        u, s, v_loc = sb.BkSvd().run_pca(a, k)
        tot_loc = dm.sum_axis_u32(0)
        v = gather_rows(np.array(v_loc))
        tot = gather_rows(tot_loc)
        if rank == 0:
            cm = orc.CountMatrix.from_cell_major(cfg.n_genes, cfg.n_cells, ip, g, c)
            np.testing.assert_array_equal(tot, cm.sum_axis_u32(0))
            np.testing.assert_array_equal(gene_tot, cm.sum_axis_u64(1))
            assert med == orc.median_mut(cm.sum_axis_u32(0))
            res_o = orc.BkSvd().run_pca(orc.normalize(cm, orc.CELLRANGER), k, threads=True)
            rel, au, av = check_pca_parity((np.array(u), np.array(s), v), res_o)
            print(f"MGPU_PARITY_OK world={world} case={case} shape={cfg.n_genes}x{cfg.n_cells} k={k} sigma_rel={rel:.2e} sinU={au:.2e} sinV={av:.2e}", flush=True)
        dist.barrier()
        a.free()
        dm.free()
    # ---- the paths sharded in round 2: m >= n branch of svd_bk, svd_rand (both shapes), partition_on_thresholds, select_cols,
    # IRLBA and the diff-exp moment consumers -- every one against the oracle on the unsharded matrix
    for case, cfg in (("wide n>m", SynthConfig(n_cells=5000, n_genes=900, seed=51)), ("tall m>=n", SynthConfig(n_cells=1600, n_genes=2400, seed=52))):
        ip, g, c = generate_host(cfg)
        lo, hi = shard_bounds(cfg.n_cells, world, rank)
        dm = generate_device(ctx, cfg, lo, hi)
        a = sb.normalize(dm, sb.Normalization.CellRanger)
        cm = orc.CountMatrix.from_cell_major(cfg.n_genes, cfg.n_cells, ip, g, c)
        a_o = orc.normalize(cm, orc.CELLRANGER)
        k = 8
        res = {}
        res["bk"] = sb.BkSvd().run_pca(a, k)
        res["rand"] = sb.RandSvd().run_pca(a, k)
        v0 = orc.irlba_start(0, cfg.n_cells)
        info = {}
        res["irlba"] = sb.irlba(a, k, tol=1e-5, maxit=50, v0=v0[lo:hi], info=info)
        sf = dm.size_factors()
        sf_all = gather_rows(sf)
        mean_g, var_g = dm.mean_var_axis(1, sf)
        cells_loc = np.arange(0, hi - lo, 3, dtype=np.uint64)
        mean_s, var_s = dm.mean_var_rows(cells_loc)
        s1, s2 = dm.sum_rows_dual(cells_loc, cells_loc[::2])
        cell_tot = np.bincount(np.repeat(np.arange(cfg.n_cells), np.diff(ip.astype(np.int64))), weights=c.astype(np.float64), minlength=cfg.n_cells)
        thr_r, thr_c = 40.0, float(np.median(cell_tot) * 0.9)  # drops rare genes and the shallowest ~40 % of the cells
        kept, resid, rows, cols = dm.partition_on_thresholds(thr_r, thr_c)
        kept_tot = gather_rows(kept.sum_axis_u32(0))
        cols_glob = gather_rows((np.asarray(cols, dtype=np.int64) + lo))
        sub = dm.select_cols(cells_loc)
        sub_tot = gather_rows(sub.sum_axis_u32(0))
        sel_glob = gather_rows(cells_loc.astype(np.int64) + lo)
        gathered = {name: (np.array(u), np.array(s_), gather_rows(np.array(v))) for name, (u, s_, v) in res.items()}
        if rank == 0:
            for name, want in (("bk", orc.BkSvd().run_pca(a_o, k, threads=True)), ("rand", orc.RandSvd().run_pca(a_o, k, threads=True))):
                rel, au, av = check_pca_parity(gathered[name], want)
                print(f"MGPU_PARITY_OK world={world} case={case} algo={name} sigma_rel={rel:.2e} sinU={au:.2e} sinV={av:.2e}", flush=True)
            uo, so, vo, mprod_o, it_o = orc.irlba(a_o, k, tol=1e-5, maxit=50, v0=v0, threads=True)
            rel, au, av = check_pca_parity(gathered["irlba"], (uo, so, vo))
            assert info["mprod"] == mprod_o and info["iterations"] == it_o, (info, mprod_o, it_o)
            print(f"MGPU_PARITY_OK world={world} case={case} algo=irlba sigma_rel={rel:.2e} sinU={au:.2e} sinV={av:.2e} mprod={mprod_o}", flush=True)
            sf_o = orc.size_factors(cm)
            np.testing.assert_array_equal(sf_all, sf_o)
            mo, vo_ = orc.mean_var_axis(cm, 1, sf_o)
            np.testing.assert_allclose(mean_g, mo, rtol=1e-12, atol=1e-300)
            np.testing.assert_allclose(var_g, vo_, rtol=1e-10, atol=1e-12 * np.abs(vo_).max())
            mo, vo_ = orc.mean_var_rows(cm, sel_glob)
            np.testing.assert_allclose(mean_s, mo, rtol=1e-15)
            np.testing.assert_allclose(var_s, vo_, rtol=1e-12, atol=1e-12 * np.abs(vo_).max())
            o1, o2 = orc.sum_rows_dual(cm, sel_glob, np.concatenate([(np.arange(0, b[1] - b[0], 3)[::2] + b[0]) for b in
                                                                     (shard_bounds(cfg.n_cells, world, r) for r in range(world))]))
            np.testing.assert_array_equal(s1, o1)
            np.testing.assert_array_equal(s2, o2)
            f_o, r_o, rows_o, cols_o = cm.partition_on_thresholds(thr_r, thr_c)
            np.testing.assert_array_equal(np.asarray(rows), rows_o)
            np.testing.assert_array_equal(cols_glob, cols_o)
            np.testing.assert_array_equal(kept_tot, f_o.sum_axis_u32(0))
            np.testing.assert_array_equal(sub_tot, cm.sum_axis_u32(0)[sel_glob])
            print(f"MGPU_PARITY_OK world={world} case={case} partition rows={len(rows_o)} cols={len(cols_o)} + select_cols + moment consumers", flush=True)
        dist.barrier()
        for h in (sub, kept, resid, a, dm):
            h.free()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
