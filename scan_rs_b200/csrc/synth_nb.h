/*
 * synth_nb.h -- counter-based negative-binomial count sampler shared by the CUDA generator
 * (synth.cu) and the CPU generator (synth/synth_cpu.c).  Test/bench utility; no reference
 * counterpart (the reference's own generators, sqz/src/gen_rand.rs and dim_red/test.rs:227-253,
 * depend on rand's streams which are not reproducible offline).
 *
 * The count of (gene g, global cell c) is a pure function of (seed, g, c, mu): a 64-bit hash
 * gives u in [0,1); the NB(mean mu, size r) CDF is inverted by sequential search using only
 * IEEE-754 correctly rounded +, *, / on doubles.  Compile WITHOUT FMA contraction (nvcc
 * -fmad=false, gcc -ffp-contract=off) and the CPU and GPU produce identical bits.
 */
#ifndef SYNTH_NB_H
#define SYNTH_NB_H

#include <stdint.h>

#ifdef __CUDACC__
#define SYNTH_HD __host__ __device__ __forceinline__
#else
#define SYNTH_HD static inline
#endif

SYNTH_HD uint64_t synth_mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xFF51AFD7ED558CCDULL;
    x ^= x >> 33;
    x *= 0xC4CEB9FE1A85EC53ULL;
    x ^= x >> 33;
    return x;
}

SYNTH_HD uint64_t synth_hash(uint64_t seed, uint64_t gene, uint64_t cell) {
    uint64_t h = synth_mix64(seed * 0x9E3779B97F4A7C15ULL + cell);
    h = synth_mix64(h ^ (gene * 0xD1B54A32D192ED03ULL + 0x2545F4914F6CDD1DULL));
    return h;
}

/* NB with size r (integer >= 1) and mean mu >= 0: P(0) = (r/(r+mu))^r,
 * P(k+1) = P(k) * (k+r)/(k+1) * mu/(r+mu). */
SYNTH_HD uint32_t synth_nb_count(uint64_t seed, uint64_t gene, uint64_t cell, double mu, uint32_t r) {
    if (!(mu > 0.0)) return 0u;
    double u = (double)(synth_hash(seed, gene, cell) >> 11) * (1.0 / 9007199254740992.0);
    /* (1+mu/r)^-r >= 1 - mu: cheap exact-zero shortcut, part of the definition */
    if (u < 1.0 - mu) return 0u;
    double rd = (double)r;
    double q = rd / (rd + mu);
    double p = mu / (rd + mu);
    double p0 = 1.0;
    for (uint32_t i = 0; i < r; i++) p0 = p0 * q;
    double cdf = p0, pk = p0;
    uint32_t k = 0;
    while (u >= cdf && k < 100000u) {
        pk = pk * (((double)k + rd) / ((double)k + 1.0)) * p;
        k++;
        cdf = cdf + pk;
        if (pk == 0.0) break;
    }
    return k;
}

#endif
