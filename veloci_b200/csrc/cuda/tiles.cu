// K2-K5 fused: posting expansion, per-part dedup-max, union / intersect merge,
// column boost and top-k, evaluated per (anchor tile, request).
//
// Follows, per anchor of the tile:
//   resolve_token_to_anchor   src/search/search_field.rs:400-504  (score = term_score * (f16 / 100), dedup keeps max)
//   union_hits_score          src/search/set_op.rs:87-220         (max per distinct term, sum * n * n)
//   intersect_hits_score      src/search/set_op.rs:368-446        (present in all, sum in list order, shortest last)
//   add_boost / apply_boost   src/search/boost.rs:470-504, 283-377
//   top_n_sort                src/search/sort.rs:5-22, order src/search.rs:123-130 (score desc, id desc)
//
// Posting lists are anchor-sorted, so the part of a list that falls into an anchor
// tile is one contiguous slice.  item_scan_kernel finds, for every (tile, request),
// the non-empty slices and writes them as flat records (tile-major); tile_eval_kernel
// takes one record per CTA: it streams the slices with coalesced loads and scatters
// score keys into one shared-memory array per search part (direct mapped: index =
// anchor - tile start, so no sorting and no hash collisions).  The epilogue walks the
// tile once -- four anchors per 128-bit shared load when the tile is densely hit,
// posting-driven when it is sparsely hit -- and does the tree evaluation, the boost
// column gather (coalesced: the tile is a contiguous range of the column), the hit
// count and a threshold test against the request's running k-th best; the few
// survivors are merged into the request's top-k heap in global memory under a
// per-request lock.  The epilogue leaves the arrays zeroed for the next item.
//
// Items are ordered tile-major (all requests of tile 0, then tile 1, ...), so the
// posting slices of a tile are re-read from L2, not HBM, by the many requests that
// share frequent terms.
#include <cuda_fp16.h>

#include "bitvec.cuh"
#include "kernels.cuh"

namespace vdev {

static const int kTileThreads = 512;
static const uint32_t kSurvivorCap = 512;
static const uint32_t kSliceChunk = 64;     // slice records staged in shared memory at a time

// ---------------------------------------------------------------- item scan
// One thread per (tile group, request), group-major.  Pass 0 counts the items and their slices; pass 1 writes them.
// A request with a FastDesc takes the plane path in every group where its non-plane terms have at most
// kGroupMaxEntries postings, all in the parts' sparse tile buckets (a group's buckets of one part are contiguous):
// such a pair is one plane-path item that only records the bucket ranges; the items are grouped per group (per-group
// cursors).  Every other pair becomes one general item per non-empty tile of the group, with all its slices; blocks
// reserve general output ranges in index order, so that list stays (nearly) tile-major.

// Tile offsets of the postings of dense match `mi`: the plane's prefix row for a plane term, else its row of the batch's table.
__device__ __forceinline__ const uint32_t* dense_row(const uint32_t* toff, const uint32_t* tprefix, const uint32_t* g_row, const uint32_t* g_plane, uint32_t mi, uint32_t n_tiles) {
    if (g_plane != nullptr) {
        const uint32_t p = g_plane[mi];
        if (p != kNoValue) return tprefix + (size_t)p * (n_tiles + 1);
    }
    return toff + (size_t)g_row[mi] * (n_tiles + 1);
}

// Postings and non-empty slices of request `qp` in tile t.
__device__ __forceinline__ void tile_totals(const ItemScanArgs& a, const QueryProgram& qp, uint32_t t, uint32_t& all_post, uint32_t& all_slices) {
    all_post = 0, all_slices = 0;
    bool dead = false;  // a part the root `and` needs has no posting in the tile: no hit, nothing to evaluate (set_op.rs:368-446)
    for (uint32_t l = 0; l < qp.n_leaves; ++l) {
        const PartSlices ps = a.slices[a.leaf_part[qp.leaf_begin + l]];
        uint32_t leaf_post = 0;
        for (uint32_t r = 0; r < ps.n_dense; ++r) {
            const uint32_t* trow = dense_row(a.toff, a.plane_tprefix, a.g_row, a.g_plane, ps.m_begin + r, a.n_tiles);
            const uint32_t n = trow[t + 1] - trow[t];
            leaf_post += n, all_slices += n ? 1u : 0u;
        }
        if (ps.n_match != ps.n_dense) {
            const uint32_t* brow = a.bucket + (size_t)ps.sparse_row * (a.n_tiles + 1);
            const uint32_t n = brow[t + 1] - brow[t];
            leaf_post += n, all_slices += n ? 1u : 0u;
        }
        all_post += leaf_post;
        if (leaf_post == 0 && l < 32 && ((qp.must_mask >> l) & 1u)) dead = true;
    }
    if (dead) all_post = 0, all_slices = 0;
}

// Entries (postings of non-plane terms) of request `qp` in tile t, over its parts; false when the tile cannot take the
// plane path (a frequent term without a plane, or entries beyond the 32-bit index).
__device__ __forceinline__ bool tile_entries(const ItemScanArgs& a, const QueryProgram& qp, uint32_t t, uint32_t& n_ent) {
    n_ent = 0;
    bool ok = true;
    for (uint32_t l = 0; l < qp.n_leaves; ++l) {
        const PartSlices ps = a.slices[a.leaf_part[qp.leaf_begin + l]];
        for (uint32_t r = 0; r < ps.n_dense; ++r)
            if (a.g_plane[ps.m_begin + r] == kNoValue) {
                const uint32_t* trow = dense_row(a.toff, a.plane_tprefix, a.g_row, a.g_plane, ps.m_begin + r, a.n_tiles);
                if (trow[t + 1] != trow[t]) ok = false;
            }
        if (ps.n_match != ps.n_dense) {
            const uint32_t* brow = a.bucket + (size_t)ps.sparse_row * (a.n_tiles + 1);
            n_ent += brow[t + 1] - brow[t];
            if (ps.sparse_base + brow[t + 1] > 0xFFFFFFFFull) ok = false;
        }
    }
    return ok;
}

// The plane-path item of request `qp` over tiles [t0, t1): the bucket ranges of its parts.
__device__ __forceinline__ FastItem make_fast_item(const ItemScanArgs& a, const QueryProgram& qp, uint32_t q, uint32_t t0, uint32_t t1) {
    FastItem fi;
    fi.q = q, fi.tiles = t0 | ((t1 - t0) << 24);
    for (uint32_t l = 0; l < kFastMaxLeaves; ++l) fi.n[l] = 0, fi.begin[l] = 0;
    for (uint32_t l = 0; l < qp.n_leaves && l < kFastMaxLeaves; ++l) {
        const PartSlices ps = a.slices[a.leaf_part[qp.leaf_begin + l]];
        if (ps.n_match != ps.n_dense) {
            const uint32_t* brow = a.bucket + (size_t)ps.sparse_row * (a.n_tiles + 1);
            fi.n[l] = (uint16_t)(brow[t1] - brow[t0]), fi.begin[l] = (uint32_t)(ps.sparse_base + brow[t0]);
        }
    }
    return fi;
}

template <bool FILL>
__global__ void __launch_bounds__(256) item_scan_kernel(ItemScanArgs a) {
    __shared__ uint32_t s_warp_items[8], s_warp_slices[8];
    __shared__ unsigned long long s_base_items, s_base_slices;
    const unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t g = 0xFFFFFFFFu, q = 0, t0 = 0, t1 = 0;
    QueryProgram qp;
    qp.active = 0, qp.n_leaves = 0;
    if (i < a.n_pairs_total) {
        g = (uint32_t)(i / a.n_queries), q = (uint32_t)(i % a.n_queries);
        t0 = g * a.group_tiles, t1 = min(a.n_tiles, t0 + a.group_tiles);
        qp = a.queries[q];
    }
    const bool live = qp.active && qp.n_leaves;
    const bool fastq = live && a.fast != nullptr && (a.fast[q].flags & kFastOk) != 0;
    FastItem whole_item;
    whole_item.q = 0, whole_item.tiles = 0;
    bool cut_any_plane = false;

    // ---- plane-path items.  The pair is one item when the entries of the whole group fit; otherwise the group is cut
    // greedily into runs of tiles that fit (a single tile that does not fit, or cannot take the plane path at all,
    // becomes a general item).  Both passes make the same cuts.
    uint32_t my_fast = 0;
    unsigned long long general_tiles = 0;  // bit (t - t0): tile t goes to the general path
    bool whole = false;                    // the pair is a single item over the whole group (the common case)
    if (fastq) {
        bool any_plane = false, ok = true;
        uint32_t group_ent = 0;
        for (uint32_t l = 0; l < qp.n_leaves; ++l) {
            const PartSlices ps = a.slices[a.leaf_part[qp.leaf_begin + l]];
            for (uint32_t r = 0; r < ps.n_dense; ++r) {
                const uint32_t* trow = dense_row(a.toff, a.plane_tprefix, a.g_row, a.g_plane, ps.m_begin + r, a.n_tiles);
                if (trow[t1] != trow[t0]) {
                    if (a.g_plane[ps.m_begin + r] != kNoValue) any_plane = true;
                    else ok = false;  // a frequent term without a plane
                }
            }
            if (ps.n_match != ps.n_dense) {
                const uint32_t* brow = a.bucket + (size_t)ps.sparse_row * (a.n_tiles + 1);
                group_ent += brow[t1] - brow[t0];
                if (ps.sparse_base + brow[t1] > 0xFFFFFFFFull) ok = false;
            }
        }
        if (ok && group_ent <= kGroupMaxEntries) {
            whole = true;
            my_fast = (any_plane || group_ent != 0) ? 1u : 0u;
            if (FILL && my_fast) whole_item = make_fast_item(a, qp, q, t0, t1);
        } else {
            // cut the group greedily into runs of tiles whose entries fit (both passes make the same cuts)
            uint32_t run_begin = t0, run_ent = 0;
            for (uint32_t t = t0; t <= t1; ++t) {
                uint32_t ne = 0;
                bool tile_ok = true;
                if (t < t1) tile_ok = tile_entries(a, qp, t, ne);
                const bool bad = t < t1 && (!tile_ok || ne > kGroupMaxEntries);
                if (t == t1 || bad || run_ent + ne > kGroupMaxEntries) {  // close the run [run_begin, t)
                    if (t > run_begin && (any_plane || run_ent != 0)) my_fast += 1;
                    run_begin = t, run_ent = 0;
                }
                if (bad) general_tiles |= 1ull << (t - t0), run_begin = t + 1;
                else run_ent += ne;
            }
            cut_any_plane = any_plane;
        }
    } else if (live) {
        general_tiles = ~0ull;
    }
    {
        // per-group cursors, one atomic per (warp, group): lanes are group-major, so the lanes of a group are a contiguous run
        uint32_t incl = my_fast;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if ((int)lane >= o) incl += y;
        }
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, g);
        const int leader = __ffs((int)peers) - 1, tail = 31 - __clz((int)peers);
        const uint32_t before_leader = __shfl_sync(0xFFFFFFFFu, incl - my_fast, leader);
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, tail) - before_leader;
        uint32_t base_i = 0;
        if ((int)lane == leader && total && g != 0xFFFFFFFFu) base_i = atomicAdd(a.fast_item_cursor + g, total);
        base_i = __shfl_sync(0xFFFFFFFFu, base_i, leader);
        if (FILL && my_fast) {
            FastItem* out = a.fast_items + a.fast_item_begin[g] + base_i + (incl - my_fast - before_leader);
            if (whole) {
                *out = whole_item;
            } else {
                uint32_t run_begin = t0, run_ent = 0;
                for (uint32_t t = t0; t <= t1; ++t) {
                    uint32_t ne = 0;
                    bool tile_ok = true;
                    if (t < t1) tile_ok = tile_entries(a, qp, t, ne);
                    const bool bad = t < t1 && (!tile_ok || ne > kGroupMaxEntries);
                    if (t == t1 || bad || run_ent + ne > kGroupMaxEntries) {
                        if (t > run_begin && (cut_any_plane || run_ent != 0)) *out++ = make_fast_item(a, qp, q, run_begin, t);
                        run_begin = t, run_ent = 0;
                    }
                    if (bad) run_begin = t + 1;
                    else run_ent += ne;
                }
            }
        }
    }

    // ---- general items: one per non-empty tile that does not take the plane path; block-ordered reservation, exclusive
    // prefix of (items, slices) over the block
    uint32_t my_items = 0, my_slices = 0;
    if (general_tiles)
        for (uint32_t t = t0; t < t1; ++t) {
            if (!((general_tiles >> (t - t0)) & 1ull)) continue;
            uint32_t np, ns;
            tile_totals(a, qp, t, np, ns);
            if (np) my_items += 1, my_slices += ns;
        }
    uint32_t xi = my_items, xs = my_slices;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t yi = __shfl_up_sync(0xFFFFFFFFu, xi, o), ys = __shfl_up_sync(0xFFFFFFFFu, xs, o);
        if ((int)lane >= o) xi += yi, xs += ys;
    }
    if (lane == 31) s_warp_items[warp] = xi, s_warp_slices[warp] = xs;
    __syncthreads();
    uint32_t wi = 0, ws = 0, ti = 0, ts = 0;
    for (uint32_t w = 0; w < 8; ++w) {
        if (w < warp) wi += s_warp_items[w], ws += s_warp_slices[w];
        ti += s_warp_items[w], ts += s_warp_slices[w];
    }
    if (threadIdx.x == 0) {
        s_base_items = ti ? atomicAdd(a.counters + 0, (unsigned long long)ti) : 0ull;
        s_base_slices = ts ? atomicAdd(a.counters + 1, (unsigned long long)ts) : 0ull;
    }
    if (!FILL) return;
    __syncthreads();
    if (!my_items) return;
    unsigned long long item_at = s_base_items + wi + xi - my_items;
    unsigned long long slice_at = s_base_slices + ws + xs - my_slices;
    for (uint32_t t = t0; t < t1; ++t) {
        if (!((general_tiles >> (t - t0)) & 1ull)) continue;
        uint32_t all_post, n_slices;
        tile_totals(a, qp, t, all_post, n_slices);
        if (!all_post) continue;
        ItemRec rec;
        rec.q = q, rec.t = t, rec.slice_begin = slice_at, rec.n_slices = n_slices, rec.npost = all_post;
        a.items[item_at++] = rec;
        uint32_t task_at = 0;
        for (uint32_t l = 0; l < qp.n_leaves; ++l) {
            const uint32_t part = a.leaf_part[qp.leaf_begin + l];
            const PartSlices ps = a.slices[part];
            for (uint32_t r = 0; r < ps.n_dense; ++r) {
                const uint32_t mi = ps.m_begin + r;
                const uint32_t* trow = dense_row(a.toff, a.plane_tprefix, a.g_row, a.g_plane, mi, a.n_tiles);
                const uint32_t s = trow[t], e = trow[t + 1];
                if (e == s) continue;
                SliceRec sr;
                sr.begin = a.g_begin[mi] + s, sr.n = e - s, sr.term_score = a.g_score[mi], sr.task_begin = task_at;
                sr.leaf = (uint16_t)l, sr.kind = 0, sr.single = ps.n_match == 1 ? 1 : 0, sr.postings = a.parts[part].postings, sr.pad = 0;
                task_at += (sr.n + kTaskPostings - 1) / kTaskPostings;
                a.slice_recs[slice_at++] = sr;
            }
            if (ps.n_match != ps.n_dense) {
                const uint32_t* brow = a.bucket + (size_t)ps.sparse_row * (a.n_tiles + 1);
                const uint32_t s = brow[t], e = brow[t + 1];
                if (e == s) continue;
                SliceRec sr;
                sr.begin = ps.sparse_base + s, sr.n = e - s, sr.term_score = 0.0f, sr.task_begin = task_at;
                sr.leaf = (uint16_t)l, sr.kind = 1, sr.single = (a.parts[part].flags & kPartListBoost) ? 3 : 0, sr.postings = 0, sr.pad = 0;
                task_at += (sr.n + kTaskPostings - 1) / kTaskPostings;
                a.slice_recs[slice_at++] = sr;
            }
        }
    }
}

void launch_item_scan(cudaStream_t st, const ItemScanArgs& a, bool fill) {
    if (!a.n_pairs_total) return;
    const unsigned blocks = (unsigned)((a.n_pairs_total + 255) / 256);
    if (fill) item_scan_kernel<true><<<blocks, 256, 0, st>>>(a);
    else item_scan_kernel<false><<<blocks, 256, 0, st>>>(a);
    count_launch();
}

// ---------------------------------------------------------------- per-anchor evaluation
__device__ __forceinline__ float apply_boost_step(const BoostStep& b, float score, float v) {
    const float x = v + b.param;
    switch (b.fun) {
        case kBoostLog10: score = score * log10f(x); break;
        case kBoostLog2: score = score * log2f(x); break;
        case kBoostMultiply: score = score * x; break;
        case kBoostAdd: score = score + x; break;
        case kBoostReplace: score = x; break;
        default: break;
    }
    if (b.expr_op != kExprNone) {
        const float l = b.expr_left_is_score ? v : b.expr_left, r = b.expr_right_is_score ? v : b.expr_right;
        float e;
        switch (b.expr_op) {
            case kExprDiv: e = l / r; break;
            case kExprMul: e = l * r; break;
            case kExprAdd: e = l + r; break;
            default: e = l - r; break;
        }
        score = score + e;
    }
    return score;
}

// ApplyAnchorBoost for an anchor with a 1:n boost (apply_boost_values_anchor, boost.rs:255-281).  The reference walks the
// part's hits and the (anchor, value) list, both ascending, with one look-ahead element; the effect is: along a run of
// consecutive hits that all have boost values, the first hit takes only its first value, the second all of its values,
// the third only the first again, ...  So an anchor with several values takes all of them exactly when the number of
// boosted hits directly before it (in the part's hit order) is odd.  Its predecessors are found in the posting lists of
// the part's matched terms, their boost entries in the list part's tile buckets.
struct LeafBoostGlobals {  // the batch tables the rule needs; one copy per CTA in shared memory
    const uint32_t* bucket;
    const SparseEntry* sparse;
    const PartSlices* slices;
    const PostingsView* postings;
    const PartQuery* parts;
    const uint64_t* g_begin;
    const uint32_t* g_df;
    const uint32_t* leaf_part;
    const uint32_t* toff;
    const uint32_t* g_row;
    const uint32_t* g_plane;
    const uint32_t* plane_tprefix;
    uint32_t n_tiles, tile_log2, anchor_lo, pad;
};
struct LeafBoostCtx {
    uint32_t part, list_part;   // the search part and the list part that holds its (anchor, 0x7FFFFFFF - value id) entries
    uint32_t anchor, idx;       // the anchor and its index in the tile
    const uint32_t* hits;       // bitmaps of the tile: anchors the part hits / anchors with boost values
    const uint32_t* boosted;
};

// For an anchor with several boosted values: takes all of them when the run of boosted hits before it is odd.
__device__ __noinline__ float apply_leaf_boost(const LeafBoostGlobals* gp, LeafBoostCtx x, const BoostStep* bsp, float score, uint32_t first_vid) {
    const LeafBoostGlobals& a = *gp;
    const BoostStep& bs = *bsp;
    const uint32_t* brow = a.bucket + (size_t)x.list_part * (a.n_tiles + 1);
    const SparseEntry* ent = a.sparse + a.slices[x.list_part].sparse_base;
    // the run of boosted hits directly before the anchor: inside the tile from the bitmaps ...
    uint32_t run = 0;
    bool open_at_tile_start = true;  // the run reaches the first anchor of the tile: continue in the posting lists
    {
        int pos = (int)x.idx - 1;
        while (pos >= 0) {
            // previous hit at or below pos
            int w = pos >> 5;
            uint32_t word = x.hits[w] & (0xFFFFFFFFu >> (31 - (pos & 31)));
            while (word == 0 && w > 0) word = x.hits[--w];
            if (word == 0) break;  // no hit before: the tile start is reached
            const int p = (w << 5) + 31 - __clz((int)word);
            if (!((x.boosted[p >> 5] >> (p & 31)) & 1u)) {
                open_at_tile_start = false;
                break;
            }
            ++run;
            pos = p - 1;
        }
    }
    if (open_at_tile_start) {  // ... and across the tile start from the part's postings of the tiles before (bucket + offset rows)
        const uint32_t tile_base = x.anchor - x.idx;
        const PartSlices ps = a.slices[x.part];
        PostingsView pv;
        pv.post = nullptr, pv.off = nullptr, pv.n_terms = 0, pv.term_plane = nullptr;
        if (ps.n_dense) pv = a.postings[a.parts[x.part].postings];  // (explicit hit lists of the step seam have no posting store)
        const uint32_t* prow = a.bucket + (size_t)ps.sparse_row * (a.n_tiles + 1);
        const SparseEntry* pent = a.sparse + ps.sparse_base;
        uint32_t cur = tile_base;
        int tt = (int)((tile_base - a.anchor_lo) >> a.tile_log2) - 1;
        while (tt >= 0) {
            // the part's largest hit below `cur` in tile tt
            uint32_t pred = 0;
            bool found = false;
            if (ps.n_match != ps.n_dense)
                for (uint32_t i = prow[tt]; i < prow[tt + 1]; ++i)
                    if (pent[i].anchor < cur && (!found || pent[i].anchor > pred)) pred = pent[i].anchor, found = true;
            for (uint32_t r = 0; r < ps.n_dense; ++r) {
                const uint32_t* trow = dense_row(a.toff, a.plane_tprefix, a.g_row, a.g_plane, ps.m_begin + r, a.n_tiles);
                const Posting* post = pv.post + a.g_begin[ps.m_begin + r];
                uint32_t lo = trow[tt], hi = trow[tt + 1];
                const uint32_t first = lo;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (post[mid].anchor < cur) lo = mid + 1;
                    else hi = mid;
                }
                if (lo > first && (!found || post[lo - 1].anchor > pred)) pred = post[lo - 1].anchor, found = true;
            }
            if (!found) {  // no hit in this tile: look further back
                --tt;
                continue;
            }
            bool has = false;
            for (uint32_t i = brow[tt]; i < brow[tt + 1] && !has; ++i) has = ent[i].anchor == pred;
            if (!has) break;
            ++run, cur = pred;  // (stay in tile tt: the hit before `pred` may be there as well)
        }
    }
    if (run & 1u) {  // every value of the anchor, in value-id order: one pass over the tile's entries of the list part
        const uint32_t t = (x.anchor - a.anchor_lo) >> a.tile_log2;
        uint32_t vids[8];
        uint32_t n = 0;
        for (uint32_t i = brow[t]; i < brow[t + 1]; ++i) {
            const SparseEntry e = ent[i];
            if (e.anchor != x.anchor) continue;
            const uint32_t v = 0x7FFFFFFFu - e.key;
            uint32_t at = 0;
            bool dup = false;
            while (at < n && vids[at] <= v) dup = dup || vids[at] == v, ++at;
            if (dup || n == 8) continue;  // (a document with more than eight matching boosted values keeps the first eight)
            for (uint32_t j = n; j > at; --j) vids[j] = vids[j - 1];
            vids[at] = v;
            ++n;
        }
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t bits = __ldg(bs.column + vids[i]);
            if (bits != kNoValue) score = apply_boost_step(bs, score, __uint_as_float(bits));
        }
        return score;
    }
    const uint32_t bits = __ldg(bs.column + first_vid);
    if (bits != kNoValue) score = apply_boost_step(bs, score, __uint_as_float(bits));
    return score;
}

// Generic request tree, postfix.  Returns presence; score in `out`.
__device__ bool eval_program(const uint32_t* __restrict__ prog, uint32_t len, const uint32_t* arr, uint32_t tile, uint32_t idx, const BoostStep* __restrict__ boosts,
                             const LeafBoostGlobals* lb, const uint32_t* lbits, uint32_t leaf_begin, uint32_t anchor, float& out) {
    uint32_t n_lb = 0;  // kOpLeafBoost ops seen so far: op k uses the bitmaps 2k (hits) and 2k + 1 (boosted)
    float sc[kMaxLeaves];
    bool pr[kMaxLeaves];
    int sp = 0;
    uint32_t pc = 0;
    while (pc < len) {
        const uint32_t op = prog[pc];
        if (op == kOpLeaf) {
            const uint32_t key = arr[prog[pc + 1] * tile + idx];
            pr[sp] = key != 0;
            sc[sp] = key ? vbit::key_score(key) : 0.0f;
            ++sp;
            pc += 2;
        } else if (op == kOpUnion) {
            const int n = (int)prog[pc + 1], ns = (int)prog[pc + 2];
            float mx[kMaxLeaves];
            for (int s = 0; s < ns; ++s) mx[s] = 0.0f;
            bool any = false;
            for (int c = 0; c < n; ++c)
                if (pr[sp - n + c]) {
                    any = true;
                    const uint32_t s = prog[pc + 3 + c];
                    mx[s] = fmaxf(mx[s], sc[sp - n + c]);
                }
            float nd = 0.0f, sum = 0.0f;
            for (int s = 0; s < ns; ++s) {
                if (mx[s] >= 0.00001f) nd += 1.0f;
                sum += mx[s];
            }
            sp -= n;
            pr[sp] = any;
            sc[sp] = sum * nd * nd;
            ++sp;
            pc += 3 + n;
        } else if (op == kOpLeafBoost) {
            const uint32_t key = arr[prog[pc + 1] * tile + idx];
            pr[sp] = key != 0;
            sc[sp] = key ? vbit::key_score(key) : 0.0f;
            const uint32_t bkey = arr[prog[pc + 2] * tile + idx];
            if (key && bkey) {  // the list leaf holds the anchor's smallest boosted value id and whether there are more
                const BoostStep& bs = boosts[prog[pc + 3]];
                const uint32_t first_vid = 0x7FFFFFFFu - (bkey & 0x7FFFFFFFu);
                if (!(bkey & 0x80000000u)) {
                    const uint32_t bits = __ldg(bs.column + first_vid);
                    if (bits != kNoValue) sc[sp] = apply_boost_step(bs, sc[sp], __uint_as_float(bits));
                } else {
                    LeafBoostCtx x;
                    x.part = lb->leaf_part[leaf_begin + prog[pc + 1]], x.list_part = lb->leaf_part[leaf_begin + prog[pc + 2]], x.anchor = anchor, x.idx = idx;
                    x.hits = lbits + (2u * n_lb) * (tile >> 5), x.boosted = lbits + (2u * n_lb + 1u) * (tile >> 5);
                    sc[sp] = apply_leaf_boost(lb, x, &bs, sc[sp], first_vid);
                }
            }
            ++n_lb;
            ++sp;
            pc += 4;
        } else if (op == kOpFilter) {
            sp -= 1;
            pr[sp - 1] = pr[sp - 1] && pr[sp];
            pc += 1;
        } else {  // kOpIntersect
            const int n = (int)prog[pc + 1];
            bool all = true;
            float sum = 0.0f;
            for (int c = 0; c < n; ++c) all = all && pr[sp - n + c];
            for (int i = 0; i < n; ++i) sum += sc[sp - n + (int)prog[pc + 2 + i]];
            sp -= n;
            pr[sp] = all;
            sc[sp] = sum;
            ++sp;
            pc += 2 + 2 * n;
        }
    }
    out = sc[0];
    return pr[0];
}

__device__ __forceinline__ uint32_t comp4(const uint4& v, int c) { return c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w; }

// Per-item state shared by the three epilogue flavours.
struct ItemCtx {
    QueryProgram qp;
    unsigned long long tau;
    uint32_t tile, tile_base;
    // first boost step in registers when it is the only one and has neither skip list nor expression
    bool fast_boost, can_prune;
    const uint32_t* col;
    uint32_t col_n, fun;
    float param;
    float prune_below;  // scores below this cannot reach the k-th best even with the largest boost multiplier
    const LeafBoostGlobals* lb;
    const uint32_t* lbits;  // hit / boosted bitmaps of the tile for the request's kOpLeafBoost ops
};

// Everything after the request tree for one present anchor: boosts, threshold, survivor list.
// Returns the key to leave in arr[0][idx] (non-zero only for a survivor deferred to the next merge round).
// `leaf_keys` (with `leaf_stride` between leaves) are the anchor's keys in the part arrays, for the post ops.
__device__ __forceinline__ uint32_t finish_anchor(const TileArgs& a, const ItemCtx& c, uint32_t anchor, float score, uint32_t* s_nsurv, unsigned long long* s_list,
                                                  const uint32_t* leaf_keys = nullptr, uint32_t leaf_stride = 0) {
    if (c.qp.n_boosts) {
        if (c.fast_boost) {
            // the boost can only shrink `score * max_mult`: skip the gather when even that cannot beat the k-th best so far
            if (c.can_prune && score < c.prune_below) return 0;
            if (anchor < c.col_n) {
                const uint32_t bits = __ldg(c.col + anchor);
                if (bits != kNoValue) {
                    const float x = __uint_as_float(bits) + c.param;
                    switch (c.fun) {
                        case kBoostLog10: score = score * log10f(x); break;
                        case kBoostLog2: score = score * log2f(x); break;
                        case kBoostMultiply: score = score * x; break;
                        case kBoostAdd: score = score + x; break;
                        case kBoostReplace: score = x; break;
                        default: break;
                    }
                }
            }
        } else {
            for (uint32_t b = 0; b < c.qp.n_boosts; ++b) {
                const BoostStep& bs = a.boosts[c.qp.boost_begin + b];
                bool skip = false;
                for (uint32_t i = 0; i < bs.n_skip; ++i) skip = skip || fabsf(bs.skip[i] - score) < 0.00001f;
                if (skip || bs.list_only || anchor >= bs.n) continue;
                const uint32_t bits = __ldg(bs.column + anchor);
                if (bits != kNoValue) score = apply_boost_step(bs, score, __uint_as_float(bits));
            }
        }
    }
    if (c.qp.post_len) {
        const uint32_t* post = a.prog + c.qp.post_begin;
        uint32_t pc = 0;
        while (pc < c.qp.post_len) {
            const uint32_t op = post[pc];
            const uint32_t key = leaf_keys[post[pc + 1] * leaf_stride];
            if (op == kPostMulIfPresent) {
                if (key) score = score * __uint_as_float(post[pc + 2]);
                pc += 3;
            } else {  // kPostMulValue
                if (key) score = score * vbit::key_score(~key);  // the leaf keeps the smallest value: keys are complemented
                pc += 2;
            }
        }
    }
    for (uint32_t f = 0; f < c.qp.n_facets; ++f) {  // count_values_for_ids / join_anchor_to_leaf (facet.rs:31-83)
        const FacetStep& fs = a.facets[c.qp.facet_begin + f];
        if (anchor >= fs.step[0].n_ids) continue;
        for (uint32_t i0 = fs.step[0].off[anchor]; i0 < fs.step[0].off[anchor + 1]; ++i0) {
            const uint32_t v0 = fs.step[0].val[i0];
            if (fs.n_steps == 1) {
                if (v0 < fs.hist_size) atomicAdd(fs.hist + v0, 1u);
                continue;
            }
            if (v0 >= fs.step[1].n_ids) continue;
            for (uint32_t i1 = fs.step[1].off[v0]; i1 < fs.step[1].off[v0 + 1]; ++i1) {
                const uint32_t v1 = fs.step[1].val[i1];
                if (fs.n_steps == 2) {
                    if (v1 < fs.hist_size) atomicAdd(fs.hist + v1, 1u);
                    continue;
                }
                if (v1 >= fs.step[2].n_ids) continue;
                for (uint32_t i2 = fs.step[2].off[v1]; i2 < fs.step[2].off[v1 + 1]; ++i2) {
                    const uint32_t v2 = fs.step[2].val[i2];
                    if (v2 < fs.hist_size) atomicAdd(fs.hist + v2, 1u);
                }
            }
        }
    }
    uint32_t key = vbit::score_key(score);
    if (key == 0) key = 1;
    const unsigned long long comp = ((unsigned long long)key << 32) | anchor;
    if (c.qp.emit_all) {
        const unsigned long long at = atomicAdd(a.emit_count, 1ull);
        if (at < a.emit_capacity) a.emit[at] = comp;
    }
    if (c.qp.k != 0 && comp > c.tau) {
        const uint32_t pos = atomicAdd(s_nsurv, 1u);
        if (pos < kSurvivorCap) s_list[pos] = comp;
        else return key;  // deferred to the next merge round
    }
    return 0;
}

// Scalar evaluation of one anchor of the tile straight from the part arrays; clears them (fused clear).
__device__ __forceinline__ uint32_t eval_idx(const TileArgs& a, const ItemCtx& c, uint32_t* arr, uint32_t idx, uint32_t* s_nsurv, unsigned long long* s_list) {
    const uint32_t L = c.qp.n_leaves;
    bool present;
    float score;
    if (c.qp.prog_len == 0) {  // one part, or a flat `or` of parts with distinct terms (leaves in slot order)
        present = false;
        float nd = 0.0f, sum = 0.0f;
        for (uint32_t l = 0; l < L; ++l) {
            const uint32_t key = arr[l * c.tile + idx];
            const float v = key ? fmaxf(0.0f, vbit::key_score(key)) : 0.0f;
            present = present || key != 0;
            if (v >= 0.00001f) nd += 1.0f;
            sum += v;
        }
        score = (L == 1 && !c.qp.union1) ? vbit::key_score(arr[idx]) : sum * nd * nd;
    } else {
        present = eval_program(a.prog + c.qp.prog_begin, c.qp.prog_len, arr, c.tile, idx, a.boosts + c.qp.boost_begin, c.lb, c.lbits, c.qp.leaf_begin, c.tile_base + idx, score);
    }
    uint32_t keep = 0;
    if (present) keep = finish_anchor(a, c, c.tile_base + idx, score, s_nsurv, s_list, arr + idx, c.tile);
    for (uint32_t l = 1; l < L; ++l) arr[l * c.tile + idx] = 0;
    arr[idx] = keep;
    return present ? 1u : 0u;
}

// ---------------------------------------------------------------- tile evaluation
// Vector sweep of a densely hit tile for programs without a tree (one part, or a flat `or`
// of L parts with distinct terms): four anchors per step, untouched groups cost one 128-bit
// shared load per part.  NONNEG: every part score is >= 0, so a key decodes with one AND.
template <int L, bool NONNEG>
__device__ __forceinline__ uint32_t sweep_flat(const TileArgs& a, const ItemCtx& c, uint32_t* arr, uint32_t tid, uint32_t* s_nsurv, unsigned long long* s_list) {
    const uint32_t tile = c.tile, n_groups = tile >> 2;
    uint32_t present = 0;
    for (uint32_t g = tid; g < n_groups; g += kTileThreads) {
        uint4 v[L];
        uint32_t any = 0;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            v[l] = reinterpret_cast<const uint4*>(arr + l * tile)[g];
            any |= v[l].x | v[l].y | v[l].z | v[l].w;
        }
        if (!any) continue;
        uint32_t keep[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            uint32_t ks[L], kor = 0;
#pragma unroll
            for (int l = 0; l < L; ++l) ks[l] = comp4(v[l], cc), kor |= ks[l];
            if (!kor) continue;
            float score;
            if (L == 1 && !c.qp.union1) {
                score = NONNEG ? __uint_as_float(ks[0] & 0x7FFFFFFFu) : vbit::key_score(ks[0]);
            } else {
                float nd = 0.0f, sum = 0.0f;
#pragma unroll
                for (int l = 0; l < L; ++l) {
                    const float x = NONNEG ? __uint_as_float(ks[l] & 0x7FFFFFFFu) : (ks[l] ? fmaxf(0.0f, vbit::key_score(ks[l])) : 0.0f);
                    if (x >= 0.00001f) nd += 1.0f;
                    sum += x;
                }
                score = sum * nd * nd;
            }
            ++present;
            keep[cc] = finish_anchor(a, c, c.tile_base + (g << 2) + cc, score, s_nsurv, s_list);
        }
#pragma unroll
        for (int l = 1; l < L; ++l)
            if (v[l].x | v[l].y | v[l].z | v[l].w) reinterpret_cast<uint4*>(arr + l * tile)[g] = make_uint4(0u, 0u, 0u, 0u);
        reinterpret_cast<uint4*>(arr)[g] = make_uint4(keep[0], keep[1], keep[2], keep[3]);
    }
    return present;
}

__device__ __forceinline__ unsigned long long shfl64(unsigned long long v, int src) { return __shfl_sync(0xFFFFFFFFu, v, src); }

// One warp task: postings [j0, j1) of slice `sr`.  pass 0 scatters score keys into the part array;
// pass 1 (sparsely hit tiles) evaluates each touched anchor once (claim bit) straight from the arrays.
__device__ __forceinline__ uint32_t run_task(const TileArgs& a, const ItemCtx& c, const SliceRec& sr, uint32_t local_task, int pass, uint32_t lane, uint32_t* arr,
                                             uint32_t* s_claim, uint32_t* s_nsurv, unsigned long long* s_list) {
    const uint32_t j0 = local_task * kTaskPostings, j1 = min(sr.n, j0 + kTaskPostings);
    const uint32_t tile_base = c.tile_base;
    uint32_t present = 0;
    if (pass == 0) {
        uint32_t* dst = arr + (uint32_t)sr.leaf * c.tile;
        if (sr.kind == 0) {
            const Posting* post = a.postings[sr.postings].post + sr.begin;
            const float ts = sr.term_score;
            if (sr.single) {  // the only list of this part: anchors are unique, plain stores
#pragma unroll 4
                for (uint32_t j = j0 + lane; j < j1; j += 32) {
                    const Posting p = post[j];
                    dst[p.anchor - tile_base] = vbit::score_key(ts * p.weight);  // hit.score * (el.score / 100.0) (:426)
                }
            } else {
#pragma unroll 4
                for (uint32_t j = j0 + lane; j < j1; j += 32) {
                    const Posting p = post[j];
                    atomicMax(&dst[p.anchor - tile_base], vbit::score_key(ts * p.weight));
                }
            }
        } else {
            const SparseEntry* se = a.sparse + sr.begin;
#pragma unroll 4
            for (uint32_t j = j0 + lane; j < j1; j += 32) {
                const SparseEntry e = se[j];
                if (sr.single != 3) {
                    atomicMax(&dst[e.anchor - tile_base], e.key);
                } else {  // 1:n boost list: smallest value id of the anchor (largest key) + bit 31 when the anchor has several
                    uint32_t* slot = &dst[e.anchor - tile_base];
                    uint32_t old = *slot, assumed;
                    do {
                        assumed = old;
                        const uint32_t ko = assumed & 0x7FFFFFFFu;
                        const uint32_t next = assumed == 0 ? e.key : (max(ko, e.key) | (assumed & 0x80000000u) | (ko != e.key ? 0x80000000u : 0u));
                        if (next == assumed) break;
                        old = atomicCAS(slot, assumed, next);
                    } while (old != assumed);
                }
            }
        }
    } else {
        for (uint32_t j = j0 + lane; j < j1; j += 32) {
            const uint32_t anchor = sr.kind == 0 ? a.postings[sr.postings].post[sr.begin + j].anchor : a.sparse[sr.begin + j].anchor;
            const uint32_t idx = anchor - tile_base;
            const uint32_t bit = 1u << (idx & 31u);
            if (!(atomicOr(&s_claim[idx >> 5], bit) & bit)) present += eval_idx(a, c, arr, idx, s_nsurv, s_list);
        }
    }
    return present;
}

__global__ void __launch_bounds__(kTileThreads) tile_eval_kernel(TileArgs a) {
    extern __shared__ __align__(16) uint32_t arr[];  // [max_leaves][tile], all zero between items
    __shared__ ItemRec s_item[2];
    __shared__ unsigned long long s_item_idx[2];
    __shared__ __align__(16) SliceRec s_slice[kSliceChunk];
    __shared__ uint32_t s_npresent, s_nsurv;
    __shared__ unsigned long long s_list[kSurvivorCap];
    __shared__ unsigned long long s_heap[kMaxK];
    __shared__ unsigned long long s_out[kMaxK];
    __shared__ uint32_t s_claim[1024];  // one bit per anchor of the tile (tiles up to 2^15)
    __shared__ LeafBoostGlobals s_lb;
    __shared__ uint32_t s_lbits[2 * kMaxLeafBoosts * 256];  // per kOpLeafBoost op: anchors the part hits, anchors with boost values
    __shared__ uint32_t s_lb_leaf[2 * kMaxLeafBoosts];

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t n_warps = kTileThreads / 32;
    const uint32_t tile = 1u << a.tile_log2;
    {
        uint4* p4 = reinterpret_cast<uint4*>(arr);
        const uint32_t n4 = (a.max_leaves * tile) >> 2;
        for (uint32_t i = tid; i < n4; i += kTileThreads) p4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) {
        s_lb.bucket = a.bucket, s_lb.sparse = a.sparse, s_lb.slices = a.slices, s_lb.postings = a.postings, s_lb.parts = a.parts;
        s_lb.g_begin = a.g_begin, s_lb.g_df = a.g_df, s_lb.leaf_part = a.leaf_part, s_lb.toff = a.toff, s_lb.g_row = a.g_row, s_lb.g_plane = a.g_plane, s_lb.plane_tprefix = a.plane_tprefix;
        s_lb.n_tiles = a.n_tiles, s_lb.tile_log2 = a.tile_log2, s_lb.anchor_lo = a.anchor_lo, s_lb.pad = 0;
        const unsigned long long first = atomicAdd(a.work_counter, 1ull);
        s_item_idx[0] = first;
        if (first < a.n_items) s_item[0] = a.items[first];
    }
    uint32_t parity = 0;

    while (true) {
        __syncthreads();  // previous item fully retired; s_item[parity] is ready
        const unsigned long long item_idx = s_item_idx[parity];
        if (item_idx >= a.n_items) break;
        const ItemRec it = s_item[parity];
        if (tid == 32) {  // prefetch the next item while this one is processed
            const unsigned long long nxt = atomicAdd(a.work_counter, 1ull);
            s_item_idx[parity ^ 1] = nxt;
            if (nxt < a.n_items) s_item[parity ^ 1] = a.items[nxt];
        }
        if (tid == 0) s_npresent = 0, s_nsurv = 0;
        parity ^= 1;

        const uint32_t t = it.t, q = it.q;
        ItemCtx c;
        c.lb = &s_lb, c.lbits = s_lbits;
        c.qp = a.queries[q];
        const QueryProgram& qp = c.qp;
        const uint32_t L = qp.n_leaves;
        const uint64_t tile_base64 = (uint64_t)a.anchor_lo + ((uint64_t)t << a.tile_log2);
        const uint32_t tile_base = (uint32_t)tile_base64;
        const uint32_t tile_n = (uint32_t)min((uint64_t)tile, (uint64_t)a.anchor_hi - tile_base64);
        c.tile = tile, c.tile_base = tile_base;
        const bool sparse_mode = it.npost * 4u < tile_n && a.queries[q].n_leaf_boosts == 0;  // (1:n boosts need the whole tile's bitmaps)
        if (sparse_mode)
            for (uint32_t i = tid; i < (tile >> 5); i += kTileThreads) s_claim[i] = 0;
        // slice records: up to 32 live in registers (one per lane, every warp has its own copy)
        const bool reg_slices = it.n_slices <= 32;
        SliceRec mine;
        mine.task_begin = 0xFFFFFFFFu, mine.n = 0;
        if (reg_slices && lane < it.n_slices) mine = a.slice_recs[it.slice_begin + lane];
        // boost step and threshold: issued now, consumed after the scatter
        c.tau = __ldcg(a.tau + q);
        c.fast_boost = false, c.can_prune = false;
        c.col = nullptr, c.col_n = 0, c.fun = 0, c.param = 0.0f, c.prune_below = 0.0f;
        if (qp.n_boosts == 1) {
            const BoostStep& bs = a.boosts[qp.boost_begin];
            if (bs.n_skip == 0 && bs.expr_op == kExprNone && !bs.list_only) {
                c.fast_boost = true;
                c.col = bs.column, c.col_n = bs.n, c.fun = bs.fun, c.param = bs.param;
                if (bs.can_prune != 0 && !qp.emit_all && qp.k != 0 && c.tau != 0 && bs.max_mult > 0.0f && qp.post_len == 0 && qp.n_facets == 0) {
                    const float tau_score = vbit::key_score((uint32_t)(c.tau >> 32));
                    if (tau_score > 1e-30f) {
                        // score < prune_below  =>  fl(score * max_mult) < tau_score  (one part in 2^20 of slack for the roundings)
                        c.prune_below = (tau_score / bs.max_mult) * 0.99999905f;
                        c.can_prune = true;
                    }
                }
            }
        }

        // (1) pass 0: scatter; pass 1 (sparsely hit tiles only): posting-driven epilogue over the same slices
        uint32_t my_present = 0;
        const int n_pass = sparse_mode ? 2 : 1;
        if (reg_slices) {
            const int last = (int)it.n_slices - 1;
            const uint32_t n_tasks = __shfl_sync(0xFFFFFFFFu, mine.task_begin, last) + (__shfl_sync(0xFFFFFFFFu, mine.n, last) + kTaskPostings - 1) / kTaskPostings;
            for (int pass = 0; pass < n_pass; ++pass) {
                for (uint32_t task = warp; task < n_tasks; task += n_warps) {
                    const int si = __popc(__ballot_sync(0xFFFFFFFFu, mine.task_begin <= task)) - 1;
                    SliceRec sr;
                    sr.begin = shfl64(mine.begin, si);
                    sr.n = __shfl_sync(0xFFFFFFFFu, mine.n, si);
                    sr.term_score = __shfl_sync(0xFFFFFFFFu, mine.term_score, si);
                    sr.task_begin = __shfl_sync(0xFFFFFFFFu, mine.task_begin, si);
                    const uint32_t packed = __shfl_sync(0xFFFFFFFFu, (uint32_t)mine.leaf | ((uint32_t)mine.kind << 16) | ((uint32_t)mine.single << 24), si);
                    sr.leaf = (uint16_t)(packed & 0xFFFFu), sr.kind = (uint8_t)((packed >> 16) & 0xFFu), sr.single = (uint8_t)(packed >> 24);
                    sr.postings = __shfl_sync(0xFFFFFFFFu, mine.postings, si);
                    my_present += run_task(a, c, sr, task - sr.task_begin, pass, lane, arr, s_claim, &s_nsurv, s_list);
                }
                __syncthreads();
            }
        } else {
            for (int pass = 0; pass < n_pass; ++pass) {
                for (uint32_t sb = 0; sb < it.n_slices; sb += kSliceChunk) {
                    const uint32_t ns = min(kSliceChunk, it.n_slices - sb);
                    __syncthreads();
                    if (tid < ns) s_slice[tid] = a.slice_recs[it.slice_begin + sb + tid];
                    __syncthreads();
                    const uint32_t first_task = s_slice[0].task_begin;
                    const uint32_t end_task = s_slice[ns - 1].task_begin + (s_slice[ns - 1].n + kTaskPostings - 1) / kTaskPostings;
                    uint32_t si = 0;
                    for (uint32_t task = first_task + warp; task < end_task; task += n_warps) {
                        while (si + 1 < ns && s_slice[si + 1].task_begin <= task) ++si;
                        const SliceRec sr = s_slice[si];
                        my_present += run_task(a, c, sr, task - sr.task_begin, pass, lane, arr, s_claim, &s_nsurv, s_list);
                    }
                }
                __syncthreads();
            }
        }

        // 1:n boosts: which anchors of the tile the boosted parts hit and which of them have boost values
        if (qp.n_leaf_boosts) {
            if (tid == 0) {
                const uint32_t* code = a.prog + qp.prog_begin;
                uint32_t pc = 0, k = 0;
                while (pc < qp.prog_len) {
                    const uint32_t op = code[pc];
                    if (op == kOpLeaf) pc += 2;
                    else if (op == kOpUnion) pc += 3 + code[pc + 1];
                    else if (op == kOpFilter) pc += 1;
                    else if (op == kOpLeafBoost) {
                        if (k < kMaxLeafBoosts) s_lb_leaf[2 * k] = code[pc + 1], s_lb_leaf[2 * k + 1] = code[pc + 2];
                        ++k, pc += 4;
                    } else pc += 2 + 2 * code[pc + 1];
                }
            }
            __syncthreads();
            const uint32_t words = tile >> 5;
            for (uint32_t m = 0; m < 2 * min(qp.n_leaf_boosts, kMaxLeafBoosts); ++m) {
                const uint32_t* src = arr + s_lb_leaf[m] * tile;
                for (uint32_t w = warp; w < words; w += n_warps) {
                    const uint32_t bits = __ballot_sync(0xFFFFFFFFu, src[(w << 5) + lane] != 0);
                    if (lane == 0) s_lbits[m * words + w] = bits;
                }
            }
            __syncthreads();
        }

        // (2) epilogue of densely hit tiles: tree, boosts, count, threshold; leaves the part arrays zeroed
        if (!sparse_mode) {
            if (qp.prog_len == 0 && L <= 4 && qp.post_len == 0 && qp.n_facets == 0) {
                if (qp.nonneg) {
                    switch (L) {
                        case 1: my_present += sweep_flat<1, true>(a, c, arr, tid, &s_nsurv, s_list); break;
                        case 2: my_present += sweep_flat<2, true>(a, c, arr, tid, &s_nsurv, s_list); break;
                        case 3: my_present += sweep_flat<3, true>(a, c, arr, tid, &s_nsurv, s_list); break;
                        default: my_present += sweep_flat<4, true>(a, c, arr, tid, &s_nsurv, s_list); break;
                    }
                } else {
                    switch (L) {
                        case 1: my_present += sweep_flat<1, false>(a, c, arr, tid, &s_nsurv, s_list); break;
                        case 2: my_present += sweep_flat<2, false>(a, c, arr, tid, &s_nsurv, s_list); break;
                        case 3: my_present += sweep_flat<3, false>(a, c, arr, tid, &s_nsurv, s_list); break;
                        default: my_present += sweep_flat<4, false>(a, c, arr, tid, &s_nsurv, s_list); break;
                    }
                }
            } else {
                for (uint32_t idx = tid; idx < tile; idx += kTileThreads) my_present += eval_idx(a, c, arr, idx, &s_nsurv, s_list);
            }
        }
        for (int o = 16; o > 0; o >>= 1) my_present += __shfl_xor_sync(0xFFFFFFFFu, my_present, o);
        if (lane == 0 && my_present) atomicAdd(&s_npresent, my_present);
        __syncthreads();
        if (tid == 0 && s_npresent) atomicAdd(a.num_hits + q, (unsigned long long)s_npresent);

        // (3) merge survivors into the request's heap (sorted, k slots) under its lock
        uint32_t nsurv = s_nsurv;
        const uint32_t k = qp.k;
        unsigned long long* heap = a.heap + (size_t)q * a.heap_stride;
        while (nsurv > 0) {
            const uint32_t n_list = min(nsurv, kSurvivorCap);
            if (tid == 0) {
                while (atomicCAS(a.lock + q, 0u, 1u) != 0u) __nanosleep(64);
                __threadfence();
            }
            __syncthreads();
            unsigned long long new_tau;
            if (k <= kMaxK) {
                for (uint32_t i = tid; i < k; i += kTileThreads) {
                    s_heap[i] = __ldcg(heap + i);
                    s_out[i] = 0;
                }
                __syncthreads();
                const uint32_t n = k + n_list;
                for (uint32_t e = tid; e < n; e += kTileThreads) {
                    const unsigned long long key = e < k ? s_heap[e] : s_list[e - k];
                    if (key == 0) continue;
                    uint32_t rank = 0;
                    for (uint32_t j = 0; j < k; ++j) rank += s_heap[j] > key;
                    for (uint32_t j = 0; j < n_list; ++j) rank += s_list[j] > key;
                    if (rank < k) s_out[rank] = key;
                }
                __syncthreads();
                for (uint32_t i = tid; i < k; i += kTileThreads) __stcg(heap + i, s_out[i]);
                new_tau = s_out[k - 1];
            } else {
                // large k: the heap stays in global memory.  It is sorted descending with the empty slots at its end, so a heap
                // key's rank is its index plus the survivors above it, a survivor's rank the heap keys above it (binary search)
                // plus the survivors above it; keys are distinct (the anchor is part of the key).
                unsigned long long* scratch = a.merge_scratch + (size_t)blockIdx.x * a.heap_stride;
                for (uint32_t i = tid; i < k; i += kTileThreads) scratch[i] = 0;
                __syncthreads();
                for (uint32_t e = tid; e < k; e += kTileThreads) {
                    const unsigned long long key = __ldcg(heap + e);
                    if (key == 0) continue;
                    uint32_t rank = e;
                    for (uint32_t j = 0; j < n_list; ++j) rank += s_list[j] > key;
                    if (rank < k) scratch[rank] = key;
                }
                for (uint32_t e = tid; e < n_list; e += kTileThreads) {
                    const unsigned long long key = s_list[e];
                    uint32_t lo = 0, hi = k;
                    while (lo < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (__ldcg(heap + mid) > key) lo = mid + 1;
                        else hi = mid;
                    }
                    uint32_t rank = lo;
                    for (uint32_t j = 0; j < n_list; ++j) rank += s_list[j] > key;
                    if (rank < k) scratch[rank] = key;
                }
                __syncthreads();
                for (uint32_t i = tid; i < k; i += kTileThreads) __stcg(heap + i, scratch[i]);
                new_tau = scratch[k - 1];
            }
            __syncthreads();
            if (tid == 0) {
                atomicMax(a.tau + q, new_tau);  // never below a threshold shared by the other shards
                __threadfence();
                atomicExch(a.lock + q, 0u);
                s_nsurv = 0;
            }
            __syncthreads();
            if (nsurv <= kSurvivorCap) break;
            // overflow: collect the deferred survivors (left in arr[0]) that still beat the new threshold
            for (uint32_t idx = tid; idx < tile_n; idx += kTileThreads) {
                const uint32_t key = arr[idx];
                if (!key) continue;
                const unsigned long long comp = ((unsigned long long)key << 32) | (tile_base + idx);
                if (comp <= new_tau) {
                    arr[idx] = 0;
                    continue;
                }
                const uint32_t pos = atomicAdd(&s_nsurv, 1u);
                if (pos < kSurvivorCap) {
                    s_list[pos] = comp;
                    arr[idx] = 0;
                }
            }
            __syncthreads();
            nsurv = s_nsurv;
        }
    }
}

static const size_t kTileStaticSmem = 24 * 1024;  // static __shared__ of tile_eval_kernel, rounded up

size_t tile_kernel_smem(uint32_t tile_log2, uint32_t max_leaves) {
    if (tile_log2 > 15) return 0;
    size_t need = ((size_t)max_leaves << tile_log2) * sizeof(uint32_t);
    return need + kTileStaticSmem <= 227 * 1024 ? need : 0;
}

void launch_tile_eval(cudaStream_t st, const TileArgs& a, int n_sms) {
    if (a.n_items == 0) return;
    const size_t smem = tile_kernel_smem(a.tile_log2, a.max_leaves);
    static PerDeviceOnce configured;
    if (configured.first()) cudaFuncSetAttribute(tile_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024 - kTileStaticSmem));
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_eval_kernel, kTileThreads, smem);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > (int)kTileBlocksPerSm) per_sm = (int)kTileBlocksPerSm;
    unsigned long long blocks = (unsigned long long)n_sms * (unsigned)per_sm;
    if (blocks > a.n_items) blocks = a.n_items;
    tile_eval_kernel<<<(unsigned)blocks, kTileThreads, smem, st>>>(a);
    count_launch();
}

// ---------------------------------------------------------------- facet groups
// One block per facet histogram: repeated arg-max in (count desc, value id asc) order below the previous pick.
__global__ void __launch_bounds__(256) facet_topk_kernel(const FacetStep* __restrict__ facets, const uint32_t* __restrict__ top, uint32_t stride, uint32_t* __restrict__ out_ids,
                                                         uint32_t* __restrict__ out_counts, uint32_t* __restrict__ out_n) {
    __shared__ unsigned long long s_best[8];
    __shared__ unsigned long long s_pick;
    const FacetStep fs = facets[blockIdx.x];
    const uint32_t want = min(top[blockIdx.x], stride);
    // order key: count in the upper half, complemented value id below (larger key = earlier group)
    unsigned long long prev = ~0ull;
    uint32_t n_out = 0;
    for (uint32_t r = 0; r < want; ++r) {
        unsigned long long best = 0;
        for (uint32_t v = threadIdx.x; v < fs.hist_size; v += blockDim.x) {
            const uint32_t cnt = fs.hist[v];
            if (!cnt) continue;
            const unsigned long long key = ((unsigned long long)cnt << 32) | (0xFFFFFFFFu - v);
            if (key < prev && key > best) best = key;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long y = __shfl_xor_sync(0xFFFFFFFFu, best, o);
            best = y > best ? y : best;
        }
        if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long b = 0;
            for (int w = 0; w < 8; ++w) b = s_best[w] > b ? s_best[w] : b;
            s_pick = b;
        }
        __syncthreads();
        const unsigned long long pick = s_pick;
        if (pick == 0) break;
        if (threadIdx.x == 0) {
            out_ids[(size_t)blockIdx.x * stride + r] = 0xFFFFFFFFu - (uint32_t)pick;
            out_counts[(size_t)blockIdx.x * stride + r] = (uint32_t)(pick >> 32);
        }
        prev = pick;
        n_out = r + 1;
        __syncthreads();
    }
    if (threadIdx.x == 0) out_n[blockIdx.x] = n_out;
}

void launch_facet_topk(cudaStream_t st, const FacetStep* facets, const uint32_t* top, uint32_t n_facets, uint32_t stride, uint32_t* out_ids, uint32_t* out_counts, uint32_t* out_n) {
    if (!n_facets) return;
    facet_topk_kernel<<<n_facets, 256, 0, st>>>(facets, top, stride, out_ids, out_counts, out_n);
    count_launch();
}

// ---------------------------------------------------------------- heap merge
// One block per request: rank-sorts the keys of `n_src` heaps (stride entries each,
// source s of request q at src[(s * n_queries + q) * stride]) into out[q][0..stride).
__global__ void __launch_bounds__(256) merge_heaps_kernel(const uint64_t* __restrict__ src_keys, const uint64_t* __restrict__ src_hits, uint32_t n_src, uint32_t n_queries,
                                                          uint32_t stride, const QueryProgram* __restrict__ queries, uint64_t* __restrict__ out_keys, uint64_t* __restrict__ out_hits) {
    extern __shared__ uint64_t keys[];
    const uint32_t q = blockIdx.x;
    const uint32_t n = n_src * stride;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t s = i / stride, j = i % stride;
        keys[i] = src_keys[((size_t)s * n_queries + q) * stride + j];
    }
    for (uint32_t i = threadIdx.x; i < stride; i += blockDim.x) out_keys[(size_t)q * stride + i] = 0;
    __syncthreads();
    const uint32_t k = min(queries[q].k, stride);
    for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) {
        const uint64_t key = keys[e];
        if (!key) continue;
        uint32_t rank = 0;
        for (uint32_t j = 0; j < n; ++j) rank += keys[j] > key;
        if (rank < k) out_keys[(size_t)q * stride + rank] = key;
    }
    if (threadIdx.x == 0) {
        uint64_t h = 0;
        for (uint32_t s = 0; s < n_src; ++s) h += src_hits[(size_t)s * n_queries + q];
        out_hits[q] = h;
    }
}

void launch_merge_heaps(cudaStream_t st, const uint64_t* src_keys, const uint64_t* src_hits, uint32_t n_src, uint32_t n_queries, uint32_t stride, const QueryProgram* queries,
                        uint64_t* out_keys, uint64_t* out_hits) {
    if (!n_queries) return;
    const size_t smem = (size_t)n_src * stride * sizeof(uint64_t);
    if (smem > 48 * 1024) {  // (requests with a large top on one device: up to kMaxKLarge keys per request)
        static PerDeviceOnce configured;
        if (configured.first()) cudaFuncSetAttribute(merge_heaps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    merge_heaps_kernel<<<n_queries, 256, smem, st>>>(src_keys, src_hits, n_src, n_queries, stride, queries, out_keys, out_hits);
    count_launch();
}

}  // namespace vdev
