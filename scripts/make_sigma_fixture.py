"""Generates tests/golden/c3_seed3_k10.json: the CPU oracle's PCA of the full BASELINE workload (1.3M cells x 33,538 genes, seed 3,
CellRanger normalization, BkSvd k=10) -- singular values plus 64-row probes of U and V.  bench.py compares every run against it.
Runs the oracle with OpenMP products on all host cores (~2-4 min on the GPU box's host, ~40 GB of RAM).
usage: python scripts/make_sigma_fixture.py [n_cells] [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as orc
from scan_rs_b200.synth import SynthConfig, generate_host

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "c3_seed3_k10.json")
M, K = 33538, 10
orc.build()
threads = orc.set_num_threads(os.cpu_count() or 1)
t0 = time.perf_counter()
cfg = SynthConfig(n_cells=n, n_genes=M, seed=3)
ip, g, c = generate_host(cfg)
cm = orc.CountMatrix.from_cell_major(M, n, ip, g, c)
del ip, g, c
t1 = time.perf_counter()
a = orc.normalize(cm, orc.CELLRANGER)
u, s, v = orc.BkSvd().run_pca(a, K, threads=True)
t2 = time.perf_counter()
rows_u = np.linspace(0, M - 1, 64).astype(int)
rows_v = np.linspace(0, n - 1, 64).astype(int)
fix = {"n_cells": n, "n_genes": M, "k": K, "seed": 3, "normalization": "CellRanger", "nnz": int(cm.nnz), "threads": threads,
       "generate_s": t1 - t0, "pca_s": t2 - t1, "sigma": [float(x) for x in s],
       "u_rows": [int(x) for x in rows_u], "u_probe": [[float(x) for x in r] for r in u[rows_u]],
       "v_rows": [int(x) for x in rows_v], "v_probe": [[float(x) for x in r] for r in v[rows_v]],
       "how": "scripts/make_sigma_fixture.py: oracle/ (CPU restatement of the reference) with OpenMP products; signs are arbitrary per column"}
json.dump(fix, open(out, "w"), indent=1)
print(f"wrote {out}: generate {t1 - t0:.1f} s, normalize+PCA {t2 - t1:.1f} s on {threads} threads; sigma {s[:3]} ...", flush=True)
