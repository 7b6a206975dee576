"""Worker of the multi-GPU parity test: launched by torchrun with one rank per GPU.  Two cases, both against the CPU oracle run on
the unsharded matrix by rank 0: a small matrix generated on the device per shard, and a larger one (33,538 genes: plane levels,
several cell tiles per rank) uploaded per shard through the narrow host form (sb_upload_compact, the bench's end-to-end path)."""
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
                                  ("k30-dense-features", SynthConfig(n_cells=12000, n_genes=2600, seed=45, n_dense=200, sigma_g=3.0), 30, True)):
        ip, g, c = generate_host(cfg)  # every rank generates the whole (small) matrix on the host and keeps its shard
        if compact:  # nnz-balanced contiguous shards, as bench.py cuts them
            bounds = shard_bounds_by_nnz(np.diff(ip.astype(np.int64)), world)
            lo, hi = bounds[rank]
            s0, s1 = int(ip[lo]), int(ip[hi])
            g16, c8, bpos, bcnt = sb.AdaptiveMat.compact_csc(g[s0:s1], c[s0:s1])
            dm = sb.AdaptiveMat.from_csc_compact(ctx, cfg.n_genes, hi - lo, (ip[lo:hi + 1] - ip[lo]).astype(np.uint64), g16, c8, bpos, bcnt)
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
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
