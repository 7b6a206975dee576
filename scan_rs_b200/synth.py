"""Synthetic single-cell count matrices (workload utility for tests and bench.py).

Not part of the reference interface.  Counts are negative binomial with a counter-based hash
keyed by (seed, gene, global cell) so any cell shard can be generated independently, and the
CUDA generator (csrc/synth.cu, through the C ABI's sb_synth_generate) and the CPU generator
(synth/synth_cpu.c) agree bit for bit (shared sampler csrc/synth_nb.h, FMA contraction off).

Model: gene abundance p_g ~ exp(sigma_g * N(0,1)) normalised to sum 1; cell depth
d_c = depth * exp(0.3 * N(0,1)); n_clusters cell programs each multiply 5% of the genes by
exp(N(0, 1)); v_gc ~ NB(mean d_c * p_g * f_{cl(c),g}, size r = 1/phi = 10).  `n_dense` extra
"antibody" features with mean dense_mean * exp(0.5 z) in every cell reproduce BASELINE config 5.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CPU = None


@dataclass
class SynthConfig:
    n_cells: int
    n_genes: int = 33538
    depth: float = 2000.0
    sigma_g: float = 2.5
    n_clusters: int = 16
    r_dispersion: int = 10
    seed: int = 1
    n_dense: int = 0          # antibody-capture style dense features appended after the genes
    dense_mean: float = 300.0


def tables(cfg: SynthConfig):
    """Host tables shared by both generators: pf[n_clusters x m], depth[n], cluster[n] (all cells)."""
    rng = np.random.default_rng(cfg.seed)
    m_g = cfg.n_genes - cfg.n_dense
    p = np.exp(cfg.sigma_g * rng.standard_normal(m_g))
    p /= p.sum()
    pf = np.tile(p, (cfg.n_clusters, 1))
    for cl in range(cfg.n_clusters):
        sel = rng.random(m_g) < 0.05
        pf[cl, sel] *= np.exp(rng.standard_normal(int(sel.sum())))
        pf[cl] /= pf[cl].sum()
    if cfg.n_dense:
        dense = cfg.dense_mean * np.exp(0.5 * rng.standard_normal(cfg.n_dense)) / cfg.depth
        pf = np.concatenate([pf, np.tile(dense, (cfg.n_clusters, 1))], axis=1)
    depth = cfg.depth * np.exp(0.3 * rng.standard_normal(cfg.n_cells))
    weights = rng.dirichlet(np.full(cfg.n_clusters, 2.0))
    cluster = rng.choice(cfg.n_clusters, size=cfg.n_cells, p=weights).astype(np.uint8)
    return np.ascontiguousarray(pf, dtype=np.float64), np.ascontiguousarray(depth), cluster


def _cpu_lib():
    global _CPU
    if _CPU is None:
        src = os.path.join(_ROOT, "synth", "synth_cpu.c")
        out_dir = os.path.join(_ROOT, "synth", "_build")
        os.makedirs(out_dir, exist_ok=True)
        so = os.path.join(out_dir, "libsynth_cpu.so")
        hdr = os.path.join(_ROOT, "scan_rs_b200", "csrc", "synth_nb.h")
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-o", so, src, "-lm"])
        _CPU = C.CDLL(so)
    return _CPU


def build_cpu():
    _cpu_lib()


def generate_host(cfg: SynthConfig, cell_lo: int = 0, cell_hi: int | None = None):
    """Cell-major arrays (indptr u64, gene u32, count u32) of cells [cell_lo, cell_hi) on the CPU."""
    cell_hi = cfg.n_cells if cell_hi is None else cell_hi
    pf, depth, cluster = tables(cfg)
    depth = np.ascontiguousarray(depth[cell_lo:cell_hi])
    cluster = np.ascontiguousarray(cluster[cell_lo:cell_hi])
    n = cell_hi - cell_lo
    lib = _cpu_lib()
    counts = np.zeros(n, dtype=np.uint32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    args = (C.c_uint32(cfg.n_genes), C.c_uint64(n), C.c_uint64(cell_lo), C.c_uint64(cfg.seed), vp(pf), vp(depth), vp(cluster),
            C.c_uint32(cfg.r_dispersion))
    lib.synth_count(*args, vp(counts))
    indptr = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(counts, out=indptr[1:])
    nnz = int(indptr[-1])
    gene = np.zeros(nnz, dtype=np.uint32)
    count = np.zeros(nnz, dtype=np.uint32)
    lib.synth_fill(*args, vp(indptr), vp(gene), vp(count))
    return indptr, gene, count


def generate_device(ctx, cfg: SynthConfig, cell_lo: int = 0, cell_hi: int | None = None):
    """Same matrix generated on the GPU; returns a device-resident AdaptiveMat."""
    from . import sqz
    cell_hi = cfg.n_cells if cell_hi is None else cell_hi
    pf, depth, cluster = tables(cfg)
    return sqz.AdaptiveMat._synth(ctx, cfg.n_genes, cell_hi - cell_lo, cell_lo, cfg.seed, pf,
                                  np.ascontiguousarray(depth[cell_lo:cell_hi]),
                                  np.ascontiguousarray(cluster[cell_lo:cell_hi]), cfg.r_dispersion)
