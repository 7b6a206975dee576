/*
 * synth_cpu.c -- CPU twin of the CUDA workload generator (scan_rs_b200/csrc/synth.cu).
 * Test / bench utility: produces the same cell-major count matrix bit for bit (shared sampler
 * in synth_nb.h; compile with -ffp-contract=off).  Two passes, OpenMP over cells.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../scan_rs_b200/csrc/synth_nb.h"

/* pass 1: per-cell nnz -> counts[n] */
void synth_count(uint32_t m, uint64_t n, uint64_t cell_offset, uint64_t seed, const double *pf, const double *depth,
                 const uint8_t *cluster, uint32_t r, uint32_t *counts) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t c = 0; c < (int64_t)n; c++) {
        const double d = depth[c];
        const double *p = pf + (size_t)cluster[c] * m;
        uint32_t total = 0;
        for (uint32_t g = 0; g < m; g++) total += synth_nb_count(seed, g, cell_offset + (uint64_t)c, d * p[g], r) != 0;
        counts[c] = total;
    }
}

/* pass 2: fill gene / count arrays given indptr[n+1] (exclusive scan of counts) */
void synth_fill(uint32_t m, uint64_t n, uint64_t cell_offset, uint64_t seed, const double *pf, const double *depth,
                const uint8_t *cluster, uint32_t r, const uint64_t *indptr, uint32_t *gene, uint32_t *count) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t c = 0; c < (int64_t)n; c++) {
        const double d = depth[c];
        const double *p = pf + (size_t)cluster[c] * m;
        uint64_t k = indptr[c];
        for (uint32_t g = 0; g < m; g++) {
            uint32_t v = synth_nb_count(seed, g, cell_offset + (uint64_t)c, d * p[g], r);
            if (v) {
                gene[k] = g;
                count[k] = v;
                k++;
            }
        }
    }
}
