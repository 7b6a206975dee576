// gather_units.h -- how the entry streams of the panelled gather (gather.cu) are cut into work units and handed to CTAs.
// Plain host C++ (no CUDA): shared by the library and by tests/cpp/gather_units_test.cpp, which checks on the CPU that
// every entry is covered exactly once, that a unit never leaves its panel and that the CTA loads are balanced.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

// A work unit: entries [begin, end) of a stream, all of one panel.
struct alignas(16) GUnit {
    uint64_t begin, end;
    uint32_t panel, pad0;
    uint64_t pad1;
};

// N side.  The stream is panel-major (base[p] .. base[p+1] are panel p's entries, base[np] == nnz).  Every one of the G CTAs
// gets a contiguous share of nnz / G entries (shares start on multiples of 8 entries), cut at panel boundaries.
// first[b] .. first[b+1] are CTA b's units.
inline void gather_units_n(const std::vector<uint64_t> &base, uint64_t nnz, uint32_t G, std::vector<GUnit> &units, std::vector<uint32_t> &first) {
    const uint32_t np = base.empty() ? 0u : (uint32_t)(base.size() - 1);
    units.clear();
    first.assign((size_t)G + 1, 0u);
    uint32_t p = 0;
    for (uint32_t b = 0; b < G; b++) {
        first[b] = (uint32_t)units.size();
        uint64_t lo = nnz / G * b + std::min<uint64_t>(b, nnz % G), hi = nnz / G * (b + 1) + std::min<uint64_t>(b + 1, nnz % G);
        lo &= ~(uint64_t)7;
        hi = b + 1 == G ? nnz : hi & ~(uint64_t)7;
        while (lo < hi) {
            while (p + 1 <= np && base[p + 1] <= lo) p++;
            const uint64_t e = std::min(hi, base[p + 1]);
            GUnit u;
            u.begin = lo;
            u.end = e;
            u.panel = p;
            u.pad0 = 0;
            u.pad1 = 0;
            units.push_back(u);
            lo = e;
        }
    }
    first[G] = (uint32_t)units.size();
}

// T side.  The stream is (cell block, gene panel)-major: segment k = block * np + panel has seg_len[k] entries in
// seg_runs[k] runs.  Every CTA owns one gene panel (or a few small ones) and sweeps the cell blocks in order, taking its
// share of each (block, panel) segment: the panel's rows are staged once per CTA, no barrier separates units, and because
// all CTAs move through the cell blocks together the output rows they reduce into stay in L2.  CTAs are handed out along
// the panel-major line of cost = entries + flush_cost * runs, cut into G equal intervals: a CTA gets the fraction
// [f0, f1) of every segment of a panel (and, where an interval crosses a panel boundary, a fraction of the next panel).
// seg_cost: fixed cost of a non-empty (block, panel) segment, in entries (the warps of a CTA split every segment's span, so a short
// segment keeps most of them idle: panels of rarely expressed genes are sequences of short segments).
inline void gather_units_t(const std::vector<uint64_t> &seg_len, const std::vector<uint64_t> &seg_runs, uint32_t np, uint32_t G, double flush_cost,
                           std::vector<GUnit> &units, std::vector<uint32_t> &first, double seg_cost = 0.0, const std::vector<double> *panel_rate = nullptr) {
    units.clear();
    first.clear();
    const size_t nblk = np ? seg_len.size() / np : 0;
    std::vector<uint64_t> seg_pos(seg_len.size() + 1, 0);
    std::vector<double> pn(np, 0.0);  // cost of a panel
    double total = 0.0;
    for (size_t k = 0; k < seg_len.size(); k++) {
        seg_pos[k + 1] = seg_pos[k] + seg_len[k];
        // panel_rate (optional): measured cycles per unit of modelled cost of every panel, from a timed pass (gather_recalibrate_t)
        const double cost = ((double)seg_len[k] + flush_cost * (double)seg_runs[k] + (seg_len[k] ? seg_cost : 0.0)) * (panel_rate ? (*panel_rate)[k % np] : 1.0);
        pn[k % np] += cost;
        total += cost;
    }
    auto add_units = [&](uint32_t p, double f0, double f1) {
        for (size_t b = 0; b < nblk; b++) {
            const size_t k = b * np + p;
            const uint64_t len = seg_len[k];
            if (len == 0) continue;
            GUnit u;
            u.begin = seg_pos[k] + (f0 <= 0.0 ? 0 : std::min<uint64_t>(len, (uint64_t)((double)len * f0)) & ~(uint64_t)7);
            u.end = seg_pos[k] + (f1 >= 1.0 ? len : std::min<uint64_t>(len, (uint64_t)((double)len * f1)) & ~(uint64_t)7);
            u.panel = p;
            u.pad0 = 0;
            u.pad1 = 0;
            if (u.end > u.begin) units.push_back(u);
        }
    };
    if (total > 0.0) {
        const double per_cta = total / G;
        uint32_t p = 0;
        double used = 0.0;  // cost of panel p already handed out
        for (uint32_t b = 0; b < G; b++) {
            first.push_back((uint32_t)units.size());
            double need = per_cta;
            while (p < np && need > 1e-9 * per_cta) {
                if (pn[p] <= 0.0) {  // a panel without entries
                    p++;
                    used = 0.0;
                    continue;
                }
                const double left = pn[p] - used;
                const bool last_cta = b + 1 == G;
                const double take = last_cta ? left : std::min(left, need);
                // A remainder too small to be worth a CTA of its own is finished here: the piece then ends at the segment
                // ends (f1 = 1), so the up-to-7 entries that the 8-entry alignment of an interior cut holds back are not lost.
                const bool finish = left - take <= 1e-9 * per_cta;
                const double f0 = used / pn[p];
                const double f1 = finish ? 1.0 : (used + take) / pn[p];
                add_units(p, f0, f1);
                need -= take;
                if (finish) {
                    p++;
                    used = 0.0;
                } else {
                    used += take;
                }
                if (last_cta) need = per_cta;  // the last CTA sweeps up whatever rounding left behind
            }
        }
    }
    if (first.empty()) first.push_back(0);
    first.push_back((uint32_t)units.size());
}
