// CPU test of the host logic that cuts the gather streams into work units (scan_rs_b200/csrc/gather_units.h):
// every entry covered exactly once, units never leave their panel / segment, CTA loads balanced.  No GPU needed.
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>

#include "../../scan_rs_b200/csrc/gather_units.h"

static int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            if (failures < 20) printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            failures++;                                                    \
        }                                                                  \
    } while (0)

static void check_tiling(std::vector<GUnit> units, uint64_t total) {
    std::sort(units.begin(), units.end(), [](const GUnit &a, const GUnit &b) { return a.begin < b.begin; });
    uint64_t pos = 0;
    for (const GUnit &u : units) {
        CHECK(u.begin == pos);
        CHECK(u.end > u.begin);
        pos = u.end;
    }
    CHECK(pos == total);
}

static void test_n(std::mt19937_64 &rng, uint32_t np, uint32_t G, uint64_t max_len, double p_empty) {
    std::vector<uint64_t> base(np + 1, 0);
    for (uint32_t p = 0; p < np; p++) {
        uint64_t len = (rng() % 1000 < p_empty * 1000) ? 0 : rng() % (max_len + 1);
        base[p + 1] = base[p] + len;
    }
    const uint64_t nnz = base[np];
    std::vector<GUnit> units;
    std::vector<uint32_t> first;
    gather_units_n(base, nnz, G, units, first);
    CHECK(first.size() == (size_t)G + 1 && first[0] == 0 && first[G] == units.size());
    check_tiling(units, nnz);
    uint64_t pos = 0;
    for (uint32_t b = 0; b < G; b++) {
        CHECK(first[b] <= first[b + 1]);
        uint64_t load = 0;
        for (uint32_t i = first[b]; i < first[b + 1]; i++) {
            const GUnit &u = units[i];
            CHECK(u.begin == pos);  // CTA lists are consecutive pieces of the stream
            pos = u.end;
            CHECK(u.panel < np && base[u.panel] <= u.begin && u.end <= base[u.panel + 1]);
            CHECK(u.begin % 8 == 0 || u.begin == base[u.panel]);
            if (i > first[b]) CHECK(units[i - 1].panel < u.panel);
            load += u.end - u.begin;
        }
        const uint64_t ideal = nnz / G;
        CHECK(load + 8 >= ideal || b + 1 == G || nnz < 8ull * G);
        CHECK(load <= ideal + 9 || b + 1 == G);
    }
}

static void test_t(std::mt19937_64 &rng, uint32_t np, uint32_t nblk, uint32_t G, uint64_t max_len, double skew, bool expect_balance) {
    std::vector<uint64_t> seg_len((size_t)np * nblk), seg_runs((size_t)np * nblk);
    for (uint32_t b = 0; b < nblk; b++)
        for (uint32_t p = 0; p < np; p++) {
            // panels of falling weight with ever shorter runs, like the expression-ranked gene panels
            double wgt = std::exp(-skew * p);
            uint64_t len = (uint64_t)(wgt * (double)(rng() % (max_len + 1)));
            uint64_t run = 1 + (uint64_t)(100.0 * wgt);
            seg_len[(size_t)b * np + p] = len;
            seg_runs[(size_t)b * np + p] = len ? (len + run - 1) / run : 0;
        }
    std::vector<uint64_t> seg_pos(seg_len.size() + 1, 0);
    for (size_t k = 0; k < seg_len.size(); k++) seg_pos[k + 1] = seg_pos[k] + seg_len[k];
    const uint64_t total = seg_pos.back();
    std::vector<GUnit> units;
    std::vector<uint32_t> first;
    const double fc = 5.0;
    gather_units_t(seg_len, seg_runs, np, G, fc, units, first);
    CHECK(first.size() >= 2 && first.size() <= (size_t)G + 1 && first.front() == 0 && first.back() == units.size());
    check_tiling(units, total);
    double tot_cost = 0.0;
    for (size_t k = 0; k < seg_len.size(); k++) tot_cost += (double)seg_len[k] + fc * (double)seg_runs[k];
    double max_cost = 0.0;
    for (size_t b = 0; b + 1 < first.size(); b++) {
        CHECK(first[b] <= first[b + 1]);
        double cost = 0.0;
        for (uint32_t i = first[b]; i < first[b + 1]; i++) {
            const GUnit &u = units[i];
            CHECK(u.panel < np);
            // the unit lies inside one (block, panel) segment of its panel
            bool inside = false;
            for (uint32_t blk = 0; blk < nblk && !inside; blk++) {
                const size_t k = (size_t)blk * np + u.panel;
                inside = seg_pos[k] <= u.begin && u.end <= seg_pos[k + 1];
                if (inside) cost += (double)(u.end - u.begin) * (1.0 + fc * (double)seg_runs[k] / (double)seg_len[k]);
            }
            CHECK(inside);
            // a CTA sweeps the cell blocks of a panel in order, panels in order
            if (i > first[b]) CHECK(units[i - 1].panel < u.panel || (units[i - 1].panel == u.panel && units[i - 1].begin < u.begin));
        }
        max_cost = std::max(max_cost, cost);
    }
    if (expect_balance && total > 0) {
        const double mean = tot_cost / G;
        if (!(max_cost <= 1.1 * mean + 64.0 * (1 + fc) * nblk)) printf("imbalance: max %.0f mean %.0f (np %u nblk %u)\n", max_cost, mean, np, nblk);
        CHECK(max_cost <= 1.1 * mean + 64.0 * (1 + fc) * nblk);
    }
}

int main() {
    std::mt19937_64 rng(12345);
    // N side: the bench shape (1270 panels, 148 CTAs), tiny and degenerate cases
    test_n(rng, 1270, 148, 400000, 0.0);
    test_n(rng, 159, 148, 3000000, 0.02);
    for (int rep = 0; rep < 5000; rep++) test_n(rng, 1 + rng() % 40, 1 + rng() % 160, rng() % 2000, 0.3);
    test_n(rng, 5, 148, 3, 0.5);   // fewer entries than CTAs
    test_n(rng, 1, 148, 0, 1.0);   // empty stream
    test_n(rng, 0, 148, 0, 0.0);   // no panels
    // T side: the bench shape (33 gene panels x 80 cell blocks), tiny and degenerate cases
    test_t(rng, 33, 80, 148, 2000000, 0.25, true);
    test_t(rng, 33, 10, 148, 500000, 0.25, true);
    test_t(rng, 59, 31, 148, 800000, 0.15, true);
    for (int rep = 0; rep < 20000; rep++) test_t(rng, 1 + rng() % 12, 1 + rng() % 9, 1 + rng() % 160, rng() % 3000, 0.3, false);
    for (int rep = 0; rep < 200; rep++) test_t(rng, 1 + rng() % 40, 1 + rng() % 90, 148, 1 + rng() % 300000, 0.05 + 0.001 * (rng() % 300), true);
    test_t(rng, 3, 2, 148, 0, 0.1, false);  // empty stream
    test_t(rng, 1, 1, 148, 5, 0.0, false);  // a handful of entries
    if (failures) {
        printf("%d CHECK(s) FAILED\n", failures);
        return 1;
    }
    printf("ALL PASSED\n");
    return 0;
}
