// adaptive.cu -- host-side decoder of the reference's compressed vector storage (sqz/src/vec.rs), so an upload can start from
// the very object the reference holds: an AdaptiveMat = Vec<AdaptiveVec> (sqz/src/mat.rs:34-42).  AdaptiveVec has eight storage
// variants (vec.rs:1029-1053); each is described to the C ABI by the raw parts of its Rust struct (sb_adaptive_vec in
// include/scanb200.h; INTEGRATION.md shows the accessor a maintainer adds to sqz).  Decoding restates AdaptiveVec::foreach /
// AbsIter::next (vec.rs:96-117, 1230-1273): positions ascending, stored zeros skipped (:113), the narrow dense codes' maximum
// value redirects to the SimpleSparse fallback (D3 7, D4 15, D8 255, D16 65535).  The encodings are NOT re-implemented on the
// device (north_star item 1: u32 indices and counts there); this runs once per upload, multi-threaded over the vectors.
#include <atomic>
#include <thread>

#include "common.cuh"

namespace {

struct Fallback {  // SimpleSparse<u32> (vec.rs:123-127): ascending indexes; walked with a cursor because positions ascend too
    const u32 *idx, *val;
    u64 n, cur;
    u32 get(u64 i) {
        while (cur < n && idx[cur] < i) cur++;
        return (cur < n && idx[cur] == i) ? val[cur] : 0u;  // SimpleSparse::get: binary search, zero when absent (:141-148)
    }
};

// value at position i of a dense-coded vector (variant 0 D3, 1 D4, 2 D8, 3 D16), before the zero filter
inline u32 dense_get(u32 variant, const void *dense, u64 i, Fallback &fb) {
    switch (variant) {
    case 0: {  // Dense3 (vec.rs:895-926): 21 values of 3 bits per u64 word
        const u64 w = static_cast<const u64 *>(dense)[i / 21];
        const u32 raw = (u32)((w >> ((i % 21) * 3)) & 7u);
        return raw == 7u ? fb.get(i) : raw;
    }
    case 1: {  // Dense4 (vec.rs:761-791): two values per byte, low nibble first
        const unsigned char b = static_cast<const unsigned char *>(dense)[i >> 1];
        const u32 raw = (i & 1) ? (u32)(b >> 4) : (u32)(b & 0xF);
        return raw == 15u ? fb.get(i) : raw;
    }
    case 2: {  // DenseW<u8, u32> (vec.rs:660-683)
        const u32 raw = static_cast<const unsigned char *>(dense)[i];
        return raw == 255u ? fb.get(i) : raw;
    }
    default: {  // DenseW<u16, u32>
        const u32 raw = static_cast<const unsigned short *>(dense)[i];
        return raw == 65535u ? fb.get(i) : raw;
    }
    }
}

// Walks one vector; idx / val may be NULL (count only).  Returns the number of non-zeros, or ~0 on a malformed description.
u64 decode_vec(const sb_adaptive_vec &v, u32 *idx, u32 *val) {
    u64 out = 0;
    Fallback fb{v.fb_idx, v.fb_val, v.fb_len, 0};
    auto emit = [&](u64 pos, u32 x) {
        if (x == 0) return;  // AbsIter::next skips stored zeros (vec.rs:113)
        if (idx) {
            idx[out] = (u32)pos;
            val[out] = x;
        }
        out++;
    };
    if (v.variant <= 3) {  // D3 / D4 / D8 / D16: every position is visited (incr = i + 1, vec.rs:700-707, 806-813)
        if (v.len && !v.dense) return ~0ull;
        for (u64 i = 0; i < v.len; i++) emit(i, dense_get(v.variant, v.dense, i, fb));
    } else if (v.variant == 4) {  // V: SimpleSparse (vec.rs:123-215)
        for (u64 j = 0; j < v.fb_len; j++) {
            if (v.fb_idx[j] >= v.len || (j && v.fb_idx[j] <= v.fb_idx[j - 1])) return ~0ull;
            emit(v.fb_idx[j], v.fb_val[j]);
        }
    } else if (v.variant <= 7) {  // S3 / S4 / S8: CompressedIndexSparse over a dense-coded value vector of length nnz (vec.rs:222-317)
        const u32 inner = v.variant - 5;  // 0 D3, 1 D4, 2 D8
        if (v.n_index && (!v.index_bytes || !v.block_starts || v.n_block_starts < 2)) return ~0ull;
        u64 block = 0;
        for (u64 j = 0; j < v.n_index; j++) {
            // incr (vec.rs:298-317): fast-forward to the block whose [start, end) holds entry j
            while (block + 1 < v.n_block_starts && !(v.block_starts[block] <= j && v.block_starts[block + 1] > j)) block++;
            if (block + 1 >= v.n_block_starts) return ~0ull;
            const u64 pos = (block << 8) | v.index_bytes[j];
            if (pos >= v.len) return ~0ull;
            emit(pos, dense_get(inner, v.dense, j, fb));
        }
    } else {
        return ~0ull;
    }
    return out;
}

}  // namespace

extern "C" int sb_adaptive_decode_vec(const sb_adaptive_vec *vec, uint32_t *idx, uint32_t *val, uint64_t capacity, uint64_t *nnz) {
    if (!vec || !nnz) return sb_fail(SB_ERR_INVALID_ARG, "sb_adaptive_decode_vec: NULL argument");
    const u64 n = decode_vec(*vec, nullptr, nullptr);
    if (n == ~0ull) return sb_fail(SB_ERR_INVALID_ARG, "sb_adaptive_decode_vec: malformed vector (variant %u)", vec->variant);
    *nnz = n;
    if (idx && val) {
        if (capacity < n) return sb_fail(SB_ERR_INVALID_ARG, "sb_adaptive_decode_vec: capacity %llu < nnz %llu", (unsigned long long)capacity, (unsigned long long)n);
        decode_vec(*vec, idx, val);
    }
    return SB_OK;
}

// AdaptiveMat -> device matrix: `major` says what the vectors are (SB_GENE_MAJOR: one per gene, the reference's usual CSR storage,
// mtx.rs:49-50; SB_CELL_MAJOR: one per cell).  Two multi-threaded passes over the vectors (count, fill), then the ordinary upload.
extern "C" int sb_upload_adaptive(sb_ctx *ctx, int major, uint32_t m, uint64_t n_local, const sb_adaptive_vec *vecs, int threads, sb_mat **out) {
    if (!ctx || !out) return sb_fail(SB_ERR_INVALID_ARG, "sb_upload_adaptive: NULL argument");
    if (major != SB_GENE_MAJOR && major != SB_CELL_MAJOR) return sb_fail(SB_ERR_INVALID_ARG, "sb_upload_adaptive: bad major %d", major);
    const u64 nvec = major == SB_GENE_MAJOR ? (u64)m : n_local, veclen = major == SB_GENE_MAJOR ? n_local : (u64)m;
    if (nvec && !vecs) return sb_fail(SB_ERR_INVALID_ARG, "sb_upload_adaptive: NULL vectors");
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = (int)std::min<u64>((u64)threads, std::max<u64>(1, nvec));
    std::vector<u64> indptr(nvec + 1, 0);
    std::atomic<int> bad{0};
    auto run = [&](auto &&fn) {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) pool.emplace_back([&, t]() { for (u64 i = t; i < nvec; i += threads) fn(i); });
        for (auto &th : pool) th.join();
    };
    run([&](u64 i) {
        if (vecs[i].len != veclen) { bad = 1; return; }
        const u64 c = decode_vec(vecs[i], nullptr, nullptr);
        if (c == ~0ull) { bad = 2; return; }
        indptr[i + 1] = c;
    });
    if (bad == 1) return sb_fail(SB_ERR_INVALID_ARG, "sb_upload_adaptive: a vector's length differs from the matrix dimension");
    if (bad == 2) return sb_fail(SB_ERR_INVALID_ARG, "sb_upload_adaptive: malformed vector");
    for (u64 i = 0; i < nvec; i++) indptr[i + 1] += indptr[i];
    std::vector<u32> idx(std::max<u64>(1, indptr[nvec])), val(std::max<u64>(1, indptr[nvec]));
    run([&](u64 i) { decode_vec(vecs[i], idx.data() + indptr[i], val.data() + indptr[i]); });
    return sb_upload(ctx, major, m, n_local, indptr.data(), idx.data(), val.data(), out);
}

// ---------------------------------------------------------------- packed host form (sb_upload_packed)
// Per entry one byte of gene delta (gene - previous gene of the cell, previous = -1 before the first entry; deltas outside 1..255
// are written as 0 and the entry's absolute gene goes to the escape list) and one nibble of count (low nibble = even stream
// position; counts >= 15 are written as 15 and go to the big-count list).  Both side lists are ordered by stream position.  The
// walk is the same AdaptiveVec::foreach order (vec.rs:1230-1273) that fills sb_upload's arrays: ascending genes inside a cell.
namespace {

struct PackRange { u64 c0, c1, n_esc, n_big; int bad; };

// cells cut into `threads` ranges of about equal entry counts
std::vector<PackRange> pack_ranges(u64 n, const u64 *indptr, int threads) {
    std::vector<PackRange> r((size_t)threads);
    const u64 nnz = indptr[n];
    u64 c = 0;
    for (int t = 0; t < threads; t++) {
        r[t].c0 = c;
        const u64 target = nnz / (u64)threads * (u64)(t + 1);
        if (t + 1 == threads) c = n;
        else c = (u64)(std::lower_bound(indptr + c, indptr + n + 1, target) - indptr), c = std::min(c, n);
        r[t].c1 = c;
        r[t].n_esc = r[t].n_big = 0;
        r[t].bad = 0;
    }
    return r;
}

// one range: counts the side-list records; with outputs, also writes deltas and side lists starting at (esc_at, big_at)
void pack_walk(PackRange &r, const u64 *indptr, const u32 *idx, const u32 *cnt, unsigned char *dgene, u64 *esc_pos, u32 *esc_gene, u64 esc_at,
               u64 *big_pos, u32 *big_cnt, u64 big_at) {
    u64 ne = 0, nb = 0;
    for (u64 c = r.c0; c < r.c1; c++) {
        u32 prev = 0xFFFFFFFFu;
        for (u64 k = indptr[c]; k < indptr[c + 1]; k++) {
            const u32 g = idx[k];
            if (k > indptr[c] && g <= prev) r.bad = 1;  // not strictly ascending inside the cell
            const u32 d = g - prev;                      // first entry: g + 1 (mod 2^32)
            const bool esc = d == 0u || d > 255u;
            if (dgene) dgene[k] = esc ? 0 : (unsigned char)d;
            if (esc) {
                if (dgene) {
                    esc_pos[esc_at + ne] = k;
                    esc_gene[esc_at + ne] = g;
                }
                ne++;
            }
            if (cnt[k] >= 15u) {
                if (dgene) {
                    big_pos[big_at + nb] = k;
                    big_cnt[big_at + nb] = cnt[k];
                }
                nb++;
            }
            prev = g;
        }
    }
    r.n_esc = ne;
    r.n_big = nb;
}

int pack_threads(int threads, u64 n) {
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    return (int)std::min<u64>((u64)threads, std::max<u64>(1, n));
}

int pack_count(u64 n, const u64 *indptr, const u32 *idx, const u32 *cnt, int threads, std::vector<PackRange> &ranges) {
    ranges = pack_ranges(n, indptr, threads);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back([&, t]() { pack_walk(ranges[t], indptr, idx, cnt, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0); });
    for (auto &th : pool) th.join();
    for (auto &r : ranges)
        if (r.bad) return sb_fail(SB_ERR_INVALID_ARG, "sb_pack_csc: gene indices must be strictly ascending inside each cell");
    return SB_OK;
}

}  // namespace

extern "C" int sb_pack_csc_count(uint64_t n, const uint64_t *indptr, const uint32_t *idx, const uint32_t *cnt, int threads, uint64_t *n_esc,
                                 uint64_t *n_big) {
    if (!indptr || !n_esc || !n_big || (indptr[n] && (!idx || !cnt))) return sb_fail(SB_ERR_INVALID_ARG, "sb_pack_csc_count: NULL argument");
    threads = pack_threads(threads, n);
    std::vector<PackRange> ranges;
    SB_TRY(pack_count(n, indptr, idx, cnt, threads, ranges));
    *n_esc = *n_big = 0;
    for (auto &r : ranges) {
        *n_esc += r.n_esc;
        *n_big += r.n_big;
    }
    return SB_OK;
}

extern "C" int sb_pack_csc_fill(uint64_t n, const uint64_t *indptr, const uint32_t *idx, const uint32_t *cnt, int threads, uint8_t *dgene,
                                uint8_t *cnt4, uint64_t *esc_pos, uint32_t *esc_gene, uint64_t *big_pos, uint32_t *big_cnt) {
    if (!indptr) return sb_fail(SB_ERR_INVALID_ARG, "sb_pack_csc_fill: NULL argument");
    const u64 nnz = indptr[n];
    if (nnz && (!idx || !cnt || !dgene || !cnt4)) return sb_fail(SB_ERR_INVALID_ARG, "sb_pack_csc_fill: NULL argument");
    threads = pack_threads(threads, n);
    std::vector<PackRange> ranges;
    SB_TRY(pack_count(n, indptr, idx, cnt, threads, ranges));
    std::vector<u64> esc_at((size_t)threads, 0), big_at((size_t)threads, 0);
    u64 te = 0, tb = 0;
    for (int t = 0; t < threads; t++) {
        esc_at[t] = te;
        big_at[t] = tb;
        te += ranges[t].n_esc;
        tb += ranges[t].n_big;
    }
    if ((te && (!esc_pos || !esc_gene)) || (tb && (!big_pos || !big_cnt))) return sb_fail(SB_ERR_INVALID_ARG, "sb_pack_csc_fill: NULL side list");
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t]() {
            pack_walk(ranges[t], indptr, idx, cnt, dgene, esc_pos, esc_gene, esc_at[t], big_pos, big_cnt, big_at[t]);
            // count nibbles: the thread that owns a byte's even entry writes the whole byte (a range may start on an odd entry)
            const u64 e0 = indptr[ranges[t].c0], e1 = indptr[ranges[t].c1];
            for (u64 b = (e0 + 1) / 2; b < (e1 + 1) / 2; b++) {
                const u32 lo = std::min(cnt[2 * b], 15u), hi = 2 * b + 1 < nnz ? std::min(cnt[2 * b + 1], 15u) : 0u;
                cnt4[b] = (unsigned char)(lo | (hi << 4));
            }
        });
    for (auto &th : pool) th.join();
    return SB_OK;
}
