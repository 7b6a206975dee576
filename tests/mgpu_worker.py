"""Worker of the multi-GPU parity test: launched by torchrun with one rank per GPU."""
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
    from scan_rs_b200.dist import gather_rows, init_context, shard_bounds
    from scan_rs_b200.synth import SynthConfig, generate_device, generate_host
    from tests.util import check_pca_parity
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = init_context(local_rank)
    cfg = SynthConfig(n_cells=6000, n_genes=1500, seed=41)
    lo, hi = shard_bounds(cfg.n_cells, world, rank)
    dm = generate_device(ctx, cfg, lo, hi)
    assert dm.cols() == hi - lo and dm.cols_global() == cfg.n_cells
    gene_tot = dm.gene_totals()
    med = dm.median_cell_total()
    a = sb.normalize(dm, sb.Normalization.CellRanger)
    u, s, v_loc = sb.BkSvd().run_pca(a, 10)
    tot_loc = dm.sum_axis_u32(0)
    v = gather_rows(v_loc)
    tot = gather_rows(tot_loc)
    if rank == 0:
        ip, g, c = generate_host(cfg)
        cm = orc.CountMatrix.from_cell_major(cfg.n_genes, cfg.n_cells, ip, g, c)
        np.testing.assert_array_equal(tot, cm.sum_axis_u32(0))
        np.testing.assert_array_equal(gene_tot, cm.sum_axis_u64(1))
        assert med == orc.median_mut(cm.sum_axis_u32(0))
        res_o = orc.BkSvd().run_pca(orc.normalize(cm, orc.CELLRANGER), 10)
        rel, au, av = check_pca_parity((u, s, v), res_o)
        print(f"MGPU_PARITY_OK world={world} sigma_rel={rel:.2e} sinU={au:.2e} sinV={av:.2e}", flush=True)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
