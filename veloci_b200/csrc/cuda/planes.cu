// Plane path of the evaluation: requests whose frequent terms have term planes.
//
// Same per-anchor semantics as tiles.cu (resolve_token_to_anchor search_field.rs:400-504,
// union_hits_score set_op.rs:87-220, add_boost boost.rs:470-504, top_n_sort sort.rs:5-22),
// restricted to flat `or` requests of at most kFastMaxLeaves parts with non-negative scores
// and at most one prunable column boost.  What changes is the work per anchor:
//
//   * presence of a plane term in an anchor is one bit of its plane, so the hit count of an
//     anchor range is a popcount over OR-ed plane words (32 anchors per operation);
//   * the score of an anchor whose parts are all present through planes is bounded by
//     bound[n] (n = parts present) times the boost multiplier; the multiplier is bounded per
//     level of the column's nested "value >= threshold" bitmaps.  Only anchors whose bound
//     reaches the request's running k-th best are evaluated exactly (gathering the weights
//     of their planes and the boost value) -- the same arithmetic, in the same order, as the
//     general path, so results are bit-identical;
//   * postings of the request's other (infrequent) terms inside the range are "entries":
//     their anchors are bounded one by one and evaluated exactly when they can matter.
//
// The unit of work is one (tile group, request) item -- `group_tiles` tiles of 2^13 anchors,
// 128 Ki anchors by default -- taken by one warp: the per-request set-up (descriptor, threshold,
// pruning levels) is paid once per group, the plane words are read straight from L2 with
// 128-bit loads (a group's rows of all planes, 16 KB each, stay L2-resident while the SMs work
// through the group's requests: items are group-major), and the warp's own evaluations tighten
// the request's threshold while it sweeps, so a cold threshold costs a few dozen evaluations,
// not one per hit.
#include <cuda_fp16.h>

#include "bitvec.cuh"
#include "kernels.cuh"

namespace vdev {

static const int kPlaneThreads = 256;
static const int kPlaneWarps = kPlaneThreads / 32;
static const uint32_t kQueueCap = 64;      // pending candidates of a warp: evaluated 32 at a time, one anchor per lane
static const uint32_t kMergeGroup = 8;     // keys of one request merged into its top-k under one lock round trip
static const uint32_t kHashSlots = 1024;   // >= 2 * kGroupMaxEntries
static const uint32_t kHashEmpty = 0xFFFFFFFFu;
static const uint32_t kPartShift = 20;     // entry code = anchor index in the group | part << 20
static_assert(kHashSlots >= 2 * kGroupMaxEntries, "hash load factor");

struct alignas(16) WarpScratch {
    float mult[kBoostLevels];              // largest boost multiplier of an anchor below the level's threshold
    uint32_t hkey[kHashSlots];             // entries: index in group | part << 20 -> largest score key
    uint32_t hval[kHashSlots];
    // candidates waiting for their exact evaluation; self-contained, so they outlive the item that produced them
    uint32_t cand_q[kQueueCap];            // request
    uint32_t cand_rel[kQueueCap];          // anchor - anchor_lo
    float cand_e[kFastMaxLeaves][kQueueCap];  // per part: largest entry score of the anchor (0: none)
    unsigned long long merge[kFastMaxK + kMergeGroup];  // a request's top-k plus the keys merged into it in one go
    const uint32_t* term_bits[kFastMaxTerms];  // the current item's plane rows, at the group's first word
    uint32_t term_part[kFastMaxTerms + 4];     // (+ 4: the entry pass reads parts four at a time)
    float term_ub[kFastMaxTerms + 4];          // largest score the term can give its part (term score x largest weight of its plane)
    uint32_t term_xkey[kFastMaxTerms + 4];     // ... as the exact order key of that product, without slack (PruneInputs: tie rule)
    int term_lev[kFastMaxTerms + 4];           // boost level an anchor whose part takes this term must be inside (-1: any, -2: it cannot matter)
    int count_lev[kFastMaxLeaves];             // ... an anchor with exactly n + 1 parts present must be inside
    const uint32_t* term_lw[kFastMaxTerms + 4]; // row the term's words are AND-ed with: its boost level's, all ones (any anchor) or all zeros (cannot matter)
    // which (boost function, param, column) mult[] was computed for
    uint32_t mult_fun;
    float mult_param;
    const ColumnLevels* mult_lev;
};

struct CtaContext {  // kernel-constant values the out-of-line helpers need
    const FastDesc* fast;
    PlaneSetView planes;
    unsigned long long* heap;
    unsigned long long* tau;
    uint32_t* lock;
    uint32_t heap_stride, anchor_lo;
};

// ---------------------------------------------------------------- index build
__global__ void plane_fill_kernel(const Posting* __restrict__ post, uint64_t n, uint32_t* __restrict__ bits_row, uint16_t* __restrict__ score_row, uint32_t* __restrict__ wmax_bits,
                                  uint32_t* __restrict__ bad, uint32_t anchor_lo, uint32_t span) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t wb = 0;
    if (i < n) {
        const Posting p = post[i];
        const uint32_t rel = p.anchor - anchor_lo;
        const __half h = __float2half_rn(__fmul_rn(p.weight, 100.0f));
        // the plane must reproduce the posting's weight exactly and the list must be strictly ascending inside the shard
        if (rel >= span || !(p.weight >= 0.0f) || __fdiv_rn(__half2float(h), 100.0f) != p.weight || (i > 0 && post[i - 1].anchor >= p.anchor)) {
            atomicExch(bad, 1u);
        } else {
            atomicOr(&bits_row[rel >> 5], 1u << (rel & 31u));
            if (score_row) score_row[rel] = __half_as_ushort(h);
            wb = __float_as_uint(p.weight);
        }
    }
    for (int o = 16; o > 0; o >>= 1) wb = max(wb, __shfl_xor_sync(0xFFFFFFFFu, wb, o));
    if ((threadIdx.x & 31) == 0 && wb) atomicMax(wmax_bits, wb);
}

void launch_plane_fill(cudaStream_t st, const Posting* post, uint64_t n, uint32_t* bits_row, uint16_t* score_row, float* wmax_slot, uint32_t* bad, uint32_t anchor_lo, uint32_t span) {
    if (!n) return;
    plane_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(post, n, bits_row, score_row, reinterpret_cast<uint32_t*>(wmax_slot), bad, anchor_lo, span);
    count_launch();
}

// One warp per (plane, tile): anchors of the plane inside the tile.
__global__ void plane_tile_count_kernel(const uint32_t* __restrict__ bits, uint32_t n_planes, uint32_t words, uint32_t* __restrict__ tcount) {
    const uint32_t tiles = words >> (kPlaneTileLog2 - 5);
    const uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (wid >= (uint64_t)n_planes * tiles) return;
    const uint32_t p = (uint32_t)(wid / tiles), t = (uint32_t)(wid % tiles);
    const uint4* row = reinterpret_cast<const uint4*>(bits + (size_t)p * words + ((size_t)t << (kPlaneTileLog2 - 5)));
    uint32_t c = 0;
    for (uint32_t i = lane; i < (1u << (kPlaneTileLog2 - 7)); i += 32) {
        const uint4 v = row[i];
        c += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if (lane == 0) tcount[wid] = c;
}

// One thread per plane: exclusive prefix sums of its tile counts.
__global__ void plane_tile_prefix_kernel(const uint32_t* __restrict__ tcount, uint32_t n_planes, uint32_t tiles, uint32_t* __restrict__ tprefix) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_planes) return;
    uint32_t acc = 0;
    for (uint32_t t = 0; t < tiles; ++t) {
        tprefix[(size_t)p * (tiles + 1) + t] = acc;
        acc += tcount[(size_t)p * tiles + t];
    }
    tprefix[(size_t)p * (tiles + 1) + tiles] = acc;
}

void launch_plane_tile_counts(cudaStream_t st, const uint32_t* bits, uint32_t n_planes, uint32_t words, uint32_t* tcount, uint32_t* tprefix) {
    const uint32_t tiles = words >> (kPlaneTileLog2 - 5);
    const uint64_t warps = (uint64_t)n_planes * tiles;
    if (!warps) return;
    plane_tile_count_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(bits, n_planes, words, tcount);
    count_launch();
    plane_tile_prefix_kernel<<<(n_planes + 63) / 64, 64, 0, st>>>(tcount, n_planes, tiles, tprefix);
    count_launch();
}

struct LevelThresholds {
    float thr[kBoostLevels];
};

// One thread per 32 anchors of the shard: bit set in level j when the anchor has no value or value >= thr[j].
__global__ void level_fill_kernel(const uint32_t* __restrict__ col, uint32_t col_n, uint32_t anchor_lo, uint32_t span, LevelThresholds th, uint32_t* __restrict__ bits, uint32_t words) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= words) return;
    uint32_t out[kBoostLevels];
#pragma unroll
    for (uint32_t j = 0; j < kBoostLevels; ++j) out[j] = 0;
    for (uint32_t b = 0; b < 32; ++b) {
        const uint32_t rel = w * 32u + b;
        if (rel >= span) break;
        const uint64_t a = (uint64_t)anchor_lo + rel;
        uint32_t v = kNoValue;
        if (a < col_n) v = col[a];
        const float f = __uint_as_float(v);
#pragma unroll
        for (uint32_t j = 0; j < kBoostLevels; ++j)
            if (v == kNoValue || f >= th.thr[j]) out[j] |= 1u << b;
    }
#pragma unroll
    for (uint32_t j = 0; j < kBoostLevels; ++j) bits[(size_t)j * words + w] = out[j];
}

void launch_level_fill(cudaStream_t st, const uint32_t* col, uint32_t col_n, uint32_t anchor_lo, uint32_t span, const float* thr, uint32_t* bits, uint32_t words) {
    if (!words) return;
    LevelThresholds th;
    for (uint32_t j = 0; j < kBoostLevels; ++j) th.thr[j] = thr[j];
    level_fill_kernel<<<(words + 127) / 128, 128, 0, st>>>(col, col_n, anchor_lo, span, th, bits, words);
    count_launch();
}

// ---------------------------------------------------------------- request descriptors
__global__ void build_fast_desc_kernel(const QueryProgram* __restrict__ queries, uint32_t n, const uint32_t* __restrict__ leaf_part, const PartPlanes* __restrict__ part_planes,
                                       const float* __restrict__ wmax, FastDesc* __restrict__ out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const QueryProgram qp = queries[q];
    FastDesc d;
    memset(&d, 0, sizeof d);
    bool ok = qp.active && qp.prog_len == 0 && qp.n_leaves >= 1 && qp.n_leaves <= kFastMaxLeaves && qp.nonneg && qp.k >= 1 && qp.k <= kFastMaxK && !qp.emit_all;
    if (qp.n_boosts) {
        ok = ok && qp.n_boosts == 1 && (qp.fb_flags & 1u) && (qp.fb_flags & 2u) && qp.fb_lev != nullptr;
        ok = ok && (qp.fb_fun == kBoostLog10 || qp.fb_fun == kBoostLog2 || qp.fb_fun == kBoostMultiply);
    }
    float u[kFastMaxLeaves] = {0.0f, 0.0f, 0.0f, 0.0f};
    uint32_t nt = 0;
    if (ok) {
        for (uint32_t l = 0; l < qp.n_leaves; ++l) {
            const PartPlanes pp = part_planes[leaf_part[qp.leaf_begin + l]];
            if (pp.n > kPartPlaneSlots || nt + pp.n > kFastMaxTerms) {
                ok = false;
                break;
            }
            d.np[l] = (uint8_t)pp.n;
            for (uint32_t j = 0; j < pp.n; ++j) {
                d.plane[nt] = (uint16_t)pp.plane[j];
                d.part[nt] = (uint8_t)l;
                d.ts[nt] = pp.ts[j];
                ++nt;
                u[l] = fmaxf(u[l], pp.ts[j] * wmax[pp.plane[j]]);
            }
            d.ub[l] = u[l] * 1.00001f;
        }
    }
    if (ok) {
        // bound[n-1]: the n largest part bounds, summed, times n * n (union_hits_score), with slack for the roundings
        for (int i = 0; i < (int)kFastMaxLeaves; ++i)
            for (int j = i + 1; j < (int)kFastMaxLeaves; ++j)
                if (u[j] > u[i]) {
                    const float tmp = u[i];
                    u[i] = u[j], u[j] = tmp;
                }
        float sum = 0.0f;
        for (uint32_t nn = 1; nn <= qp.n_leaves; ++nn) {
            sum += u[nn - 1];
            const float f = qp.n_leaves == 1 ? 1.0f : (float)(nn * nn);
            d.bound[nn - 1] = sum * f * 1.00001f;
        }
        d.flags = kFastOk | (qp.n_boosts ? kFastBoost : 0u) | (qp.union1 ? kFastUnion1 : 0u);
        d.n_leaves = qp.n_leaves, d.k = qp.k, d.n_terms = nt;
        d.fb_fun = qp.fb_fun, d.fb_param = qp.fb_param, d.fb_max_mult = qp.fb_max_mult, d.fb_n = qp.fb_n;
        d.fb_col = qp.fb_col, d.fb_lev = qp.fb_lev;
    } else {
        memset(&d, 0, sizeof d);
    }
    out[q] = d;
}

void launch_build_fast_desc(cudaStream_t st, const QueryProgram* queries, uint32_t n, const uint32_t* leaf_part, const PartPlanes* part_planes, const float* wmax, FastDesc* out) {
    if (!n) return;
    build_fast_desc_kernel<<<(n + 127) / 128, 128, 0, st>>>(queries, n, leaf_part, part_planes, wmax, out);
    count_launch();
}

// ---------------------------------------------------------------- plane evaluation
__device__ __forceinline__ float boost_mult(uint32_t fun, float x) {
    switch (fun) {
        case kBoostLog10: return log10f(x);
        case kBoostLog2: return log2f(x);
        default: return x;  // kBoostMultiply
    }
}

// Deepest boost level whose outside cannot reach the threshold with bound B (mult[] ascends): -1 = every anchor.
__device__ __forceinline__ int deepest_level(const WarpScratch& S, float B, float tau_score) {
    int lo = -1;
#pragma unroll
    for (int step = (int)kBoostLevels / 2; step > 0; step >>= 1)
        if (B * S.mult[lo + step] < tau_score) lo += step;
    if (lo + 1 < (int)kBoostLevels && B * S.mult[lo + 1] < tau_score) lo += 1;
    return lo;
}

// What the sweep needs to know about a request's threshold (warp-uniform).  A sweep anchor has no entries, so only parts
// with plane terms can be present in it; with P such parts its score is at most
// (sum over present parts of the strongest present term's bound) * n^2 * boost multiplier, n <= P.  Hence, against the
// current k-th best:
//   term_lev[t]  boost level an anchor must be inside when term t is the strongest term of its part, even with every
//                other part at its best (-1: any anchor, -2: such an anchor cannot matter) -- a fuzzy neighbour next to
//                the exact term only matters deep inside the boost column, if at all
//   allow        bit t: term_lev[t] != -2
//   optional     bit l: part l may be absent from an anchor that still matters (always set for parts without plane terms)
struct Prune {
    float tau_score;   // score of the request's k-th best so far when pruning is possible, else 0
    uint32_t allow, optional;
    bool possible;     // some sweep anchor may still matter
    bool by_count;     // some part with plane terms may be absent: anchors are also held against the bound of their part count (count_lev)
};

// Static (per item) inputs of the pruning state: lane t holds term t's bound, lane l part l's `opt_need`.
struct PruneInputs {
    float bound;       // lane < nt: best bound (before the multiplier) of an anchor whose part takes this term
    float opt_need;    // lane < 4: best bound of an anchor without this part, times the largest multiplier
    float count_bound; // lane < 4: best bound (before the multiplier) of an anchor with exactly lane + 1 parts present
    float max_mult;    // largest boost multiplier (1 without a boost)
    uint32_t part_terms[kFastMaxLeaves];  // per part: mask of its terms
    uint32_t nt;
    // Requests of one part without a boost: a sweep anchor's score is exactly ts * weight of its strongest plane term, so
    // term t gives at most exact_key = key(ts[t] * wmax[plane t]), the very product eval_candidate forms for an anchor of
    // that weight.  Against a threshold of the same score the anchor id decides (larger wins, search.rs:123-130): when every
    // anchor of the item lies below the threshold's anchor, the term cannot matter in this item.  Items are taken from the
    // highest anchors down, so the ties of a frequent term are settled by its first group instead of evaluated everywhere.
    // (exact_key lives in WarpScratch::term_xkey, the item's last anchor is passed to compute_prune)
    const uint32_t* lev_bits;             // the boost column's level rows at the item's first word (nullptr without a boost)
    uint32_t lev_words;
    const uint32_t* ones_row;
    const uint32_t* zeros_row;
};

__device__ __forceinline__ Prune compute_prune(WarpScratch& S, const PruneInputs& in, uint32_t flags, unsigned long long tau, uint32_t lane, bool tie_rule, uint32_t item_hi) {
    Prune p;
    p.tau_score = 0.0f;
    if (tau != 0) {
        const float ts = vbit::key_score((uint32_t)(tau >> 32));
        if (ts > 1e-30f) p.tau_score = ts;
    }
    int lv = -1;
    if (lane < in.nt && p.tau_score > 0.0f) {
        if (in.bound * in.max_mult < p.tau_score) lv = -2;
        else if (flags & kFastBoost) lv = deepest_level(S, in.bound, p.tau_score);
        else if (tie_rule) {
            const uint32_t tau_key = (uint32_t)(tau >> 32), exact_key = S.term_xkey[lane];
            if (exact_key < tau_key || (exact_key == tau_key && item_hi < (uint32_t)tau)) lv = -2;
        }
    }
    int cl = -1;
    if (lane < kFastMaxLeaves && p.tau_score > 0.0f) {
        if (in.count_bound * in.max_mult < p.tau_score) cl = -2;
        else if (flags & kFastBoost) cl = deepest_level(S, in.count_bound, p.tau_score);
    }
    __syncwarp();
    if (lane < in.nt) {
        S.term_lev[lane] = lv;
        S.term_lw[lane] = lv == -2 ? in.zeros_row : lv == -1 ? in.ones_row : in.lev_bits + (size_t)lv * in.lev_words;
    }
    if (lane < kFastMaxLeaves) S.count_lev[lane] = cl;
    p.allow = __ballot_sync(0xFFFFFFFFu, lane < in.nt && lv != -2);
    p.optional = __ballot_sync(0xFFFFFFFFu, lane < kFastMaxLeaves && (in.part_terms[lane & 3u] == 0u || in.opt_need >= p.tau_score));
    p.possible = p.allow != 0u;
    p.by_count = false;
#pragma unroll
    for (uint32_t l = 0; l < kFastMaxLeaves; ++l)
        if (in.part_terms[l] != 0u) {
            if ((p.optional >> l) & 1u) p.by_count = p.tau_score > 0.0f;
            else if ((p.allow & in.part_terms[l]) == 0u) p.possible = false;  // a part that must be there, and none of its terms can
        }
    __syncwarp();
    return p;
}

// Merges the order keys of the lanes in `group` (at most kMergeGroup, all of request q, each above the threshold it was
// compared with) into the request's top-k under its lock: one lock round trip for all of them.
__device__ __noinline__ void merge_group(const CtaContext* C, WarpScratch* Sp, uint32_t lane, uint32_t q, unsigned long long comp, uint32_t group) {
    WarpScratch& S = *Sp;
    const uint32_t k = C->fast[q].k;
    unsigned long long* heap = C->heap + (size_t)q * C->heap_stride;
    if (lane == 0) {
        while (atomicCAS(C->lock + q, 0u, 1u) != 0u) __nanosleep(32);
        __threadfence();
    }
    __syncwarp();
    for (uint32_t i = lane; i < k; i += 32) S.merge[i] = __ldcg(heap + i);
    __syncwarp();
    // a key that is in the top-k already (its anchor was evaluated before: the seed pass and the sweep may both reach it) is dropped
    if (group & (1u << lane))
        for (uint32_t i = 0; i < k; ++i)
            if (S.merge[i] == comp) comp = 0;
    group = __ballot_sync(0xFFFFFFFFu, (group & (1u << lane)) && comp != 0);
    const uint32_t cnt = __popc(group);
    if (group & (1u << lane)) S.merge[k + __popc(group & ((1u << lane) - 1u))] = comp;
    __syncwarp();
    const uint32_t n = k + cnt;
    const unsigned long long worst = S.merge[k - 1];
    if (__ballot_sync(0xFFFFFFFFu, (group & (1u << lane)) && comp > worst)) {
        for (uint32_t e = lane; e < n; e += 32) {
            const unsigned long long key = S.merge[e];
            if (key == 0) continue;
            uint32_t rank = 0;
            for (uint32_t j = 0; j < n; ++j) rank += S.merge[j] > key;
            if (rank < k) __stcg(heap + rank, key);
            if (rank == k - 1) atomicMax(C->tau + q, key);  // never below a threshold shared by the other shards
        }
        __threadfence();
    }
    __syncwarp();
    if (lane == 0) atomicExch(C->lock + q, 0u);
    __syncwarp();
}

__device__ __forceinline__ uint32_t hash_slot(uint32_t code) { return (code * 0x9E3779B1u) >> 22; }  // 10 bits: kHashSlots

// Files `key` under `code` (keeping the maximum); true when this call created the code's slot.
__device__ __forceinline__ bool hash_insert(WarpScratch& S, uint32_t code, uint32_t key) {
    uint32_t h = hash_slot(code);
    bool first;
    while (true) {
        const uint32_t old = atomicCAS(&S.hkey[h], kHashEmpty, code);
        first = old == kHashEmpty;
        if (first || old == code) break;
        h = (h + 1u) & (kHashSlots - 1u);
    }
    atomicMax(&S.hval[h], key);
    return first;
}

// Largest score key of the entries with this code, 0 when there is none.
__device__ __forceinline__ uint32_t hash_lookup(const WarpScratch& S, uint32_t code) {
    uint32_t h = hash_slot(code);
    while (true) {
        const uint32_t k = S.hkey[h];
        if (k == code) return S.hval[h];
        if (k == kHashEmpty) return 0u;
        h = (h + 1u) & (kHashSlots - 1u);
    }
}

// Weight of `anchor` in the posting list of a mid plane's term (the anchor is in the list: its plane bit is set).  The
// plane's tile prefix row narrows the search to the postings of the anchor's tile.
__device__ __forceinline__ float posting_weight(const PlaneSetView& P, uint32_t plane, uint32_t rel, uint32_t anchor) {
    const uint32_t tiles = P.words >> (kPlaneTileLog2 - 5), t = rel >> kPlaneTileLog2;
    const uint32_t* row = P.tprefix + (size_t)plane * (tiles + 1) + t;
    uint32_t lo = __ldg(row), hi = __ldg(row + 1);
    const Posting* post = P.info[plane].post;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&post[mid].anchor) < anchor) lo = mid + 1;
        else hi = mid;
    }
    return __ldg(&post[lo].weight);
}

// Exact score of candidate `c` of the warp's list: its order key.
__device__ __forceinline__ unsigned long long eval_candidate(const CtaContext& C, const WarpScratch& S, uint32_t c, uint32_t q) {
    const FastDesc* __restrict__ D = C.fast + q;
    const uint32_t rel = S.cand_rel[c];
    const uint32_t w = rel >> 5, bit = 1u << (rel & 31u);
    const uint32_t L = D->n_leaves, flags = D->flags, nt = D->n_terms;
    const uint32_t anchor = C.anchor_lo + rel;
    float sum = 0.0f, nd = 0.0f, v0 = 0.0f;
    uint32_t t = 0;
#pragma unroll 1
    for (uint32_t l = 0; l < L; ++l) {
        float v = S.cand_e[l][c];
#pragma unroll 1
        for (; t < nt && D->part[t] == l; ++t) {
            const uint32_t p = D->plane[t];
            if (__ldg(C.planes.bits + (size_t)p * C.planes.words + w) & bit) {
                float wgt;
                if (p < C.planes.n_head) {
                    const unsigned short h = __ldg(C.planes.score + (size_t)p * ((size_t)C.planes.words * 32u) + rel);
                    wgt = __fdiv_rn(__half2float(__ushort_as_half(h)), 100.0f);  // el.score.to_f32() / 100.0 (search_field.rs:426)
                } else {
                    wgt = posting_weight(C.planes, p, rel, anchor);  // the same value, kept in the posting (checked at index build)
                }
                v = fmaxf(v, D->ts[t] * wgt);
            }
        }
        if (v >= 0.00001f) nd += 1.0f;
        sum += v;
        if (l == 0) v0 = v;
    }
    float score = (L == 1 && !(flags & kFastUnion1)) ? v0 : sum * nd * nd;
    if (flags & kFastBoost) {
        if (anchor < D->fb_n) {
            const uint32_t bits = __ldg(D->fb_col + anchor);
            if (bits != kNoValue) {
                const float x = __uint_as_float(bits) + D->fb_param;
                switch (D->fb_fun) {
                    case kBoostLog10: score = score * log10f(x); break;
                    case kBoostLog2: score = score * log2f(x); break;
                    default: score = score * x; break;
                }
            }
        }
    }
    uint32_t key = vbit::score_key(score);
    if (key == 0) key = 1;
    return ((unsigned long long)key << 32) | anchor;
}

// Evaluates candidates [from, from + count) (count <= 32) of the warp's list, one per lane, and merges the survivors.
__device__ __noinline__ void drain(const CtaContext* C, WarpScratch* Sp, uint32_t lane, uint32_t from, uint32_t count) {
    WarpScratch& S = *Sp;
    unsigned long long comp = 0;
    uint32_t q = 0;
    if (lane < count) {
        q = S.cand_q[from + lane];
        comp = eval_candidate(*C, S, from + lane, q);
        if (comp <= __ldcg(C->tau + q)) comp = 0;
    }
    uint32_t m = __ballot_sync(0xFFFFFFFFu, comp != 0);
    while (m) {  // survivors of one request (an item's candidates are neighbours in the list) merge together
        const int src = __ffs((int)m) - 1;
        const uint32_t sq = __shfl_sync(0xFFFFFFFFu, q, src);
        uint32_t group = __ballot_sync(0xFFFFFFFFu, comp != 0 && q == sq) & m;
        if (__popc(group) > (int)kMergeGroup) {  // the lowest kMergeGroup lanes of the group now, the others next time round
            uint32_t rest = group;
            for (uint32_t i = 0; i < kMergeGroup; ++i) rest &= rest - 1u;
            group &= ~rest;
        }
        m &= ~group;
        merge_group(C, Sp, lane, sq, comp, group);
    }
}

// Appends the flagged lanes' anchors (with the entry scores of their parts) to the warp's candidate list and evaluates
// a full group of 32 when there is one (returns true then: thresholds may have moved).
__device__ __forceinline__ bool enqueue(const CtaContext* C, WarpScratch& S, uint32_t lane, uint32_t q, bool flag, uint32_t rel, float e0, float e1, float e2, float e3, uint32_t& qn,
                                        uint32_t& ncand) {
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, flag);
    if (!m) return false;
    if (flag) {
        const uint32_t at = qn + __popc(m & ((1u << lane) - 1u));
        S.cand_q[at] = q, S.cand_rel[at] = rel;
        S.cand_e[0][at] = e0, S.cand_e[1][at] = e1, S.cand_e[2][at] = e2, S.cand_e[3][at] = e3;
    }
    qn += __popc(m);
    ncand += __popc(m);
    __syncwarp();
    if (qn >= 32) {
        qn -= 32;
        drain(C, &S, lane, qn, 32);
        return true;
    }
    return false;
}

__device__ __forceinline__ void or4(uint4& a, const uint4& b) { a.x |= b.x, a.y |= b.y, a.z |= b.z, a.w |= b.w; }
__device__ __forceinline__ void and4(uint4& a, const uint4& b) { a.x &= b.x, a.y &= b.y, a.z &= b.z, a.w &= b.w; }
__device__ __forceinline__ uint32_t popc4(const uint4& a) { return __popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w); }
__device__ __forceinline__ bool any4(const uint4& a) { return (a.x | a.y | a.z | a.w) != 0u; }

__global__ void __launch_bounds__(kPlaneThreads, 2) plane_eval_kernel(PlaneArgs a) {
    extern __shared__ __align__(16) unsigned char plane_smem[];
    __shared__ CtaContext s_ctx;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    WarpScratch& S = reinterpret_cast<WarpScratch*>(plane_smem)[warp];
    const CtaContext* C = &s_ctx;
    if (tid == 0) {
        s_ctx.fast = a.fast, s_ctx.planes = a.planes;
        s_ctx.heap = a.heap, s_ctx.tau = a.tau, s_ctx.lock = a.lock, s_ctx.heap_stride = a.heap_stride, s_ctx.anchor_lo = a.anchor_lo;
    }
    for (uint32_t i = lane; i < kHashSlots; i += 32) S.hkey[i] = kHashEmpty, S.hval[i] = 0;
    if (lane == 0) S.mult_lev = nullptr, S.mult_fun = 0xFFFFFFFFu, S.mult_param = 0.0f;
    __syncthreads();
    unsigned long long st_cand = 0, st_items = 0, st_general = 0, st_sweepless = 0;  // lane 0
    const uint32_t ibeg = a.group_item_begin[a.group_begin], iend = a.group_item_begin[a.group_end];
    const uint32_t words = a.planes.words;
    uint32_t qn = 0;  // candidates of this warp waiting for their exact evaluation

    while (true) {
        uint32_t first = 0;
        if (lane == 0) first = (uint32_t)atomicAdd(a.work_counter, (unsigned long long)a.item_batch);
        first = ibeg + __shfl_sync(0xFFFFFFFFu, first, 0);
        if (first >= iend) break;
        const uint32_t last = min(iend, first + a.item_batch);
#pragma unroll 1
        for (uint32_t it = first; it < last; ++it) {
            // ---- item (tiles [t0, t0 + n), request q); the list is group-major ascending and is taken from its end: the groups
            // of the highest anchors first (see PruneInputs::tie_rule)
            const uint32_t ii = ibeg + (iend - 1u - it);
            const uint4* rp = reinterpret_cast<const uint4*>(a.items + ii);
            const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
            const uint32_t q = r0.x, t0 = r0.y & 0xFFFFFFu, n_item_tiles = r0.y >> 24;  // the item's tiles: [t0, t0 + n_item_tiles)
            const uint32_t c0 = r0.z & 0xFFFFu, c1 = c0 + (r0.z >> 16), c2 = c1 + (r0.w & 0xFFFFu), n_ent = c2 + (r0.w >> 16);  // entry prefix sums over the parts
            const FastDesc* __restrict__ D = a.fast + q;
            const uint32_t flags = D->flags, L = D->n_leaves, nt = D->n_terms;
            unsigned long long tau = __ldcg(a.tau + q);
            const uint32_t w0 = t0 << (kPlaneTileLog2 - 5);                              // first word of the item in every plane
            const uint32_t nw = min(n_item_tiles << (kPlaneTileLog2 - 5), words - w0);   // its words
            const uint32_t rel0 = w0 << 5;                                               // its first anchor, relative to the shard
            const ColumnLevels* __restrict__ lev = (flags & kFastBoost) ? D->fb_lev : nullptr;
            const uint32_t* lev_bits = nullptr;
            uint32_t lev_words = 0;
            if (lev) {
                lev_bits = lev->bits + w0, lev_words = lev->words;
                const uint32_t fun = D->fb_fun;
                const float param = D->fb_param;
                if (S.mult_lev != lev || S.mult_fun != fun || S.mult_param != param) {
                    __syncwarp();
                    if (lane < kBoostLevels) {
                        const float m = boost_mult(fun, lev->thr[lane] + param);
                        S.mult[lane] = fmaxf(m, 0.0f) * 1.00001f + 1e-6f;
                    }
                    if (lane == 0) S.mult_lev = lev, S.mult_fun = fun, S.mult_param = param;
                }
            }
            __syncwarp();
            if (lane < nt) {
                S.term_bits[lane] = a.planes.bits + (size_t)D->plane[lane] * words + w0;
                S.term_part[lane] = D->part[lane];
                S.term_ub[lane] = D->ts[lane] * __ldg(a.planes.wmax + D->plane[lane]) * 1.00001f;
                const uint32_t xkey = vbit::score_key(D->ts[lane] * __ldg(a.planes.wmax + D->plane[lane]));
                S.term_xkey[lane] = xkey ? xkey : 1u;  // (eval_candidate never returns key 0)
            }
            const bool tie_rule = L == 1 && !(flags & (kFastBoost | kFastUnion1));
            const uint32_t item_hi = a.anchor_lo + rel0 + (nw << 5) - 1u;
            __syncwarp();
            PruneInputs pin;
            {
                // P parts have plane terms; a sweep anchor with n of them present scores at most (sum of their bounds) * n^2
                const float u0 = D->ub[0], u1 = D->ub[1], u2 = D->ub[2], u3 = D->ub[3];
                const uint32_t P = (D->np[0] != 0) + (D->np[1] != 0) + (D->np[2] != 0) + (D->np[3] != 0);
                const float sum_all = u0 + u1 + u2 + u3;
                const float f_all = L == 1 ? 1.0f : (float)(P * P), f_less = L == 1 ? 1.0f : (float)((P - 1u) * (P - 1u));
                pin.nt = nt;
                pin.max_mult = (flags & kFastBoost) ? D->fb_max_mult * 1.00001f : 1.0f;
                const uint32_t my_part = lane < nt ? S.term_part[lane] : 0u;
                const float my_best = my_part == 0 ? u0 : my_part == 1 ? u1 : my_part == 2 ? u2 : u3;
                pin.bound = lane < nt ? (sum_all - my_best + S.term_ub[lane]) * f_all * 1.0001f : 0.0f;
                const float mine = (lane & 3u) == 0 ? u0 : (lane & 3u) == 1 ? u1 : (lane & 3u) == 2 ? u2 : u3;
                pin.opt_need = (sum_all - mine) * f_less * pin.max_mult * 1.0001f;
                pin.count_bound = D->bound[lane & 3u];  // (0 beyond the request's parts: such a count does not occur)
                pin.lev_bits = lev_bits, pin.lev_words = lev_words, pin.ones_row = a.ones_row, pin.zeros_row = a.zeros_row;

#pragma unroll
                for (uint32_t l = 0; l < kFastMaxLeaves; ++l) pin.part_terms[l] = __ballot_sync(0xFFFFFFFFu, lane < nt && my_part == l);
            }
            // terms are grouped by part: bit t = term t is the last term of its part
            const uint32_t last_mask = __ballot_sync(0xFFFFFFFFu, lane < nt && (lane + 1u == nt || S.term_part[lane + 1u] != S.term_part[lane]));
            Prune pr = compute_prune(S, pin, flags, tau, lane, tie_rule, item_hi);
            uint32_t cnt = 0, ncand = 0;
            bool seeding = false;
            // candidates of one step: `cm` = anchors that pass the word-parallel tests
            auto candidates = [&](uint32_t w4, const uint4& cm) {
                if (__ballot_sync(0xFFFFFFFFu, any4(cm)) == 0) return;
#pragma unroll 1
                for (int c4 = 0; c4 < 4; ++c4) {
                    uint32_t m = c4 == 0 ? cm.x : c4 == 1 ? cm.y : c4 == 2 ? cm.z : cm.w;
                    while (__ballot_sync(0xFFFFFFFFu, m != 0)) {
                        bool flag = m != 0;
                        const uint32_t bp = flag ? (uint32_t)__ffs((int)m) - 1u : 0u;
                        const uint32_t idx = (w4 * 4u + (uint32_t)c4) * 32u + bp;
                        m &= m - 1u;
                        // Second stage: this anchor's bound from the terms it really has (their words come from L1 now).
                        if (flag && pr.tau_score > 0.0f) {
                            float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f, v3 = 0.0f;
#pragma unroll 1
                            for (uint32_t t = 0; t < nt; ++t)
                                if ((__ldg(S.term_bits[t] + (idx >> 5)) >> bp) & 1u) {
                                    const uint32_t part = S.term_part[t];
                                    const float ub = S.term_ub[t];
                                    if (part == 0) v0 = fmaxf(v0, ub);
                                    else if (part == 1) v1 = fmaxf(v1, ub);
                                    else if (part == 2) v2 = fmaxf(v2, ub);
                                    else v3 = fmaxf(v3, ub);
                                }
                            const uint32_t np = (v0 > 0.0f) + (v1 > 0.0f) + (v2 > 0.0f) + (v3 > 0.0f);
                            const float sum = v0 + v1 + v2 + v3;
                            const float B = (L == 1 ? sum : sum * (float)(np * np)) * 1.00001f;
                            if (flags & kFastBoost) {
                                if (B * pin.max_mult < pr.tau_score) flag = false;
                                else {
                                    const int lo = deepest_level(S, B, pr.tau_score);
                                    if (lo >= 0) flag = (__ldg(lev_bits + (size_t)lo * lev_words + (idx >> 5)) >> bp) & 1u;
                                }
                            } else if (B * 1.00001f < pr.tau_score) {
                                flag = false;
                            }
                        }
                        // anchors with entries belong to the entry pass above (it knows their entry scores)
                        if (flag && n_ent) {
                            bool has = hash_lookup(S, idx) != 0u;
                            if (L > 1) has = has || hash_lookup(S, idx | (1u << kPartShift)) != 0u;
                            if (L > 2) has = has || hash_lookup(S, idx | (2u << kPartShift)) != 0u;
                            if (L > 3) has = has || hash_lookup(S, idx | (3u << kPartShift)) != 0u;
                            flag = !has;
                        }
                        if (enqueue(C, S, lane, q, flag, rel0 + idx, 0.0f, 0.0f, 0.0f, 0.0f, qn, ncand)) {
                            tau = __ldcg(a.tau + q);
                            if (!seeding) pr = compute_prune(S, pin, flags, tau, lane, tie_rule, item_hi);  // the rest of the sweep prunes against the tightened threshold
                        }
                    }
                }
            };

            // ---- anchors with entries: bounded one by one (entry scores are known exactly, plane parts by their bound)
            uint32_t own = 0;  // bit r: this lane's entry of round r is the first of its (anchor, part)
            if (n_ent) {
#pragma unroll 1
                for (uint32_t r = 0; r * 32u < n_ent; ++r) {
                    const uint32_t j = r * 32u + lane;
                    if (j < n_ent) {
                        const uint32_t l = (j >= c0) + (j >= c1) + (j >= c2);
                        const uint32_t at = l == 0 ? r1.x + j : l == 1 ? r1.y + (j - c0) : l == 2 ? r1.z + (j - c1) : r1.w + (j - c2);
                        const uint2 e = __ldg(reinterpret_cast<const uint2*>(a.sparse + at));
                        if (hash_insert(S, (e.x - a.anchor_lo - rel0) | (l << kPartShift), e.y)) own |= 1u << r;
                    }
                }
                __syncwarp();
            }

            // ---- a request without a threshold yet (its first item): seed pass.  Evaluating anchors in anchor order until the
            // threshold has converged costs hundreds of evaluations; the anchors that have every part with plane terms present
            // (for a single such part: its anchors in the top 1/64 of the boost column) are few and contain good hits, so they
            // go first, until a first batch of them has given the request a threshold.  Nothing is counted here, and the sweep
            // below reaches the same anchors again: merging is idempotent.
            const uint32_t n_planed = (pin.part_terms[0] != 0u) + (pin.part_terms[1] != 0u) + (pin.part_terms[2] != 0u) + (pin.part_terms[3] != 0u);
            if (pr.tau_score == 0.0f && nt > 0 && (n_planed >= 2 || (flags & kFastBoost))) {
                seeding = true;
                const uint4* seed_lw = n_planed >= 2 ? nullptr : reinterpret_cast<const uint4*>(lev_bits + (size_t)5 * lev_words);
                const uint32_t nw4 = nw >> 2;
#pragma unroll 1
                for (uint32_t w4 = lane; w4 < nw4; w4 += 32) {
                    uint4 cm = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), cur = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
                    for (uint32_t t = 0; t < nt; ++t) {
                        or4(cur, __ldg(reinterpret_cast<const uint4*>(S.term_bits[t]) + w4));
                        if ((last_mask >> t) & 1u) and4(cm, cur), cur = make_uint4(0u, 0u, 0u, 0u);
                    }
                    if (seed_lw && any4(cm)) and4(cm, __ldg(seed_lw + w4));
                    candidates(w4, cm);
                    if (tau != 0) break;  // a first batch has been evaluated and the request has a threshold: enough
                }
                if (qn) drain(C, &S, lane, 0, qn), qn = 0;
                seeding = false;
                tau = __ldcg(a.tau + q);
                pr = compute_prune(S, pin, flags, tau, lane, tie_rule, item_hi);
            }

            if (n_ent) {
#pragma unroll 1
                for (uint32_t r = 0; r * 32u < n_ent; ++r) {
                    const uint32_t j = r * 32u + lane;
                    bool cand = false;
                    uint32_t rel = 0;
                    float e0 = 0.0f, e1 = 0.0f, e2 = 0.0f, e3 = 0.0f;
                    if ((own >> r) & 1u) {
                        const uint32_t el = (j >= c0) + (j >= c1) + (j >= c2);
                        const uint32_t at = el == 0 ? r1.x + j : el == 1 ? r1.y + (j - c0) : el == 2 ? r1.z + (j - c1) : r1.w + (j - c2);
                        const uint32_t idx = __ldg(&a.sparse[at].anchor) - a.anchor_lo - rel0;
                        rel = rel0 + idx;
                        // the anchor's entries per part; its representative is the first entry of its lowest part
                        uint32_t ev0 = hash_lookup(S, idx), ev1 = 0u, ev2 = 0u, ev3 = 0u;
                        if (L > 1) ev1 = hash_lookup(S, idx | (1u << kPartShift));
                        if (L > 2) ev2 = hash_lookup(S, idx | (2u << kPartShift));
                        if (L > 3) ev3 = hash_lookup(S, idx | (3u << kPartShift));
                        const bool rep = el == 0 || (el == 1 && !ev0) || (el == 2 && !ev0 && !ev1) || (el == 3 && !ev0 && !ev1 && !ev2);
                        if (rep) {
                            e0 = ev0 ? __uint_as_float(ev0 & 0x7FFFFFFFu) : 0.0f, e1 = ev1 ? __uint_as_float(ev1 & 0x7FFFFFFFu) : 0.0f;
                            e2 = ev2 ? __uint_as_float(ev2 & 0x7FFFFFFFu) : 0.0f, e3 = ev3 ? __uint_as_float(ev3 & 0x7FFFFFFFu) : 0.0f;
                            uint32_t by_plane = 0;  // bit l: part l is present through one of its planes
                            float pub0 = 0.0f, pub1 = 0.0f, pub2 = 0.0f, pub3 = 0.0f;  // ... and the most such a plane can give it
                            auto present = [&](uint32_t t) {
                                const uint32_t l = S.term_part[t];
                                const float ub = S.term_ub[t];
                                by_plane |= 1u << l;
                                if (l == 0) pub0 = fmaxf(pub0, ub);
                                else if (l == 1) pub1 = fmaxf(pub1, ub);
                                else if (l == 2) pub2 = fmaxf(pub2, ub);
                                else pub3 = fmaxf(pub3, ub);
                            };
                            const uint32_t w = idx >> 5, bit = 1u << (idx & 31u);
#pragma unroll 1
                            for (uint32_t t = 0; t < nt; t += 4) {  // four independent loads at a time
                                const uint32_t x0 = __ldg(S.term_bits[t] + w);
                                const uint32_t x1 = t + 1 < nt ? __ldg(S.term_bits[t + 1] + w) : 0u;
                                const uint32_t x2 = t + 2 < nt ? __ldg(S.term_bits[t + 2] + w) : 0u;
                                const uint32_t x3 = t + 3 < nt ? __ldg(S.term_bits[t + 3] + w) : 0u;
                                if (x0 & bit) present(t);
                                if (x1 & bit) present(t + 1);
                                if (x2 & bit) present(t + 2);
                                if (x3 & bit) present(t + 3);
                            }
                            if (!by_plane) cnt += 1;  // a hit the plane sweep does not count
                            float sum_ub = 0.0f;
                            uint32_t n = 0;
                            auto part = [&](uint32_t l, uint32_t ev, float e, float pub) {
                                if (ev || ((by_plane >> l) & 1u)) {
                                    n += 1;
                                    sum_ub += fmaxf(e, pub);
                                }
                            };
                            part(0, ev0, e0, pub0);
                            if (L > 1) part(1, ev1, e1, pub1);
                            if (L > 2) part(2, ev2, e2, pub2);
                            if (L > 3) part(3, ev3, e3, pub3);
                            const float B = (L == 1 ? sum_ub : sum_ub * (float)(n * n)) * 1.00001f;
                            const float tau_score = pr.tau_score;
                            cand = true;
                            if (tau_score > 0.0f) {
                                if (flags & kFastBoost) {
                                    if (B * D->fb_max_mult * 1.00001f < tau_score) cand = false;
                                    else {
                                        const int lo = deepest_level(S, B, tau_score);  // the anchor must be inside that level
                                        if (lo >= 0) cand = (__ldg(lev_bits + (size_t)lo * lev_words + w) & bit) != 0;
                                    }
                                } else if (B * 1.00001f < tau_score) {
                                    cand = false;
                                }
                            }
                        }
                    }
                    if (enqueue(C, S, lane, q, cand, rel, e0, e1, e2, e3, qn, ncand)) {
                        tau = __ldcg(a.tau + q);
                        pr = compute_prune(S, pin, flags, tau, lane, tie_rule, item_hi);
                    }
                }
            }

            // ---- anchors without entries: the plane sweep, four 32-anchor words per lane and step
            bool swept_cold = false;
            const bool count_from_table = nt == 1 && !pr.possible;  // one plane term and no anchor of it can matter any more
            if (nt == 0) {
                // nothing but entries
            } else if (count_from_table) {
                // the hit count is the plane's count of the group's tiles (+ the entry anchors outside it, counted above)
                const uint32_t tiles_total = words >> (kPlaneTileLog2 - 5);
                if (lane < n_item_tiles && t0 + lane < tiles_total) cnt += __ldg(a.planes.tcount + (size_t)D->plane[0] * tiles_total + t0 + lane);
            } else {
                const uint4* tp0 = reinterpret_cast<const uint4*>(S.term_bits[0]);
                const uint4* tp1 = reinterpret_cast<const uint4*>(S.term_bits[nt > 1 ? 1 : 0]);
                const uint4* tp2 = reinterpret_cast<const uint4*>(S.term_bits[nt > 2 ? 2 : 0]);
                const uint4* tp3 = reinterpret_cast<const uint4*>(S.term_bits[nt > 3 ? 3 : 0]);
                const uint4 ones4 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                // One step per iteration: the plane words of up to four terms and the rows they are restricted to (a boost level,
                // all ones, or all zeros for a term that cannot matter: no branches) are requested before any is used, so up to
                // eight independent 128-bit loads are in flight per lane.  Terms are grouped by part: a running OR over a part's
                // restricted words is AND-ed into the candidate word when the part's last term has been seen (if the part must
                // be present), and -- when parts may be absent -- the part's presence word goes into a bit-sliced counter, so
                // that anchors are also held against the boost level their number of present parts needs.
                const uint32_t nw4 = nw >> 2;  // a multiple of 64
                swept_cold = pr.tau_score == 0.0f;
#pragma unroll 1
                for (uint32_t w4 = lane; w4 < nw4; w4 += 32) {
                    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                    const uint4 a0 = __ldg(tp0 + w4);
                    uint4 a1 = z, a2 = z, a3 = z;
                    if (nt > 1) a1 = __ldg(tp1 + w4);
                    if (nt > 2) a2 = __ldg(tp2 + w4);
                    if (nt > 3) a3 = __ldg(tp3 + w4);
                    if (!pr.possible) {  // nothing can matter any more: count only
                        uint4 any = make_uint4(a0.x | a1.x | a2.x | a3.x, a0.y | a1.y | a2.y | a3.y, a0.z | a1.z | a2.z | a3.z, a0.w | a1.w | a2.w | a3.w);
#pragma unroll 1
                        for (uint32_t t = 4; t < nt; ++t) or4(any, __ldg(reinterpret_cast<const uint4*>(S.term_bits[t]) + w4));
                        cnt += popc4(any);
                        continue;
                    }
                    const uint4 l0 = __ldg(reinterpret_cast<const uint4*>(S.term_lw[0]) + w4);
                    uint4 l1 = z, l2 = z, l3 = z;
                    if (nt > 1) l1 = __ldg(reinterpret_cast<const uint4*>(S.term_lw[1]) + w4);
                    if (nt > 2) l2 = __ldg(reinterpret_cast<const uint4*>(S.term_lw[2]) + w4);
                    if (nt > 3) l3 = __ldg(reinterpret_cast<const uint4*>(S.term_lw[3]) + w4);
                    const uint32_t req_mask = (pr.optional & 1u ? 0u : pin.part_terms[0]) | (pr.optional & 2u ? 0u : pin.part_terms[1]) | (pr.optional & 4u ? 0u : pin.part_terms[2]) |
                                              (pr.optional & 8u ? 0u : pin.part_terms[3]);  // bit t: term t's part must be present
                    uint4 any = z, anyx = z, cur = z, craw = z, cm = ones4, ones = z, twos = z, fours = z;
                    auto combine = [&](uint32_t t, const uint4& v, const uint4& lw) {
                        const uint4 x = make_uint4(v.x & lw.x, v.y & lw.y, v.z & lw.z, v.w & lw.w);
                        or4(any, v), or4(cur, x), or4(anyx, x);
                        if (pr.by_count) or4(craw, v);
                        if ((last_mask >> t) & 1u) {
                            if ((req_mask >> t) & 1u) and4(cm, cur);
                            cur = z;
                            if (pr.by_count) {  // one more part present: ones / twos / fours count the parts per anchor
                                uint32_t cy, cy2;
                                cy = ones.x & craw.x, ones.x ^= craw.x, cy2 = twos.x & cy, twos.x ^= cy, fours.x |= cy2;
                                cy = ones.y & craw.y, ones.y ^= craw.y, cy2 = twos.y & cy, twos.y ^= cy, fours.y |= cy2;
                                cy = ones.z & craw.z, ones.z ^= craw.z, cy2 = twos.z & cy, twos.z ^= cy, fours.z |= cy2;
                                cy = ones.w & craw.w, ones.w ^= craw.w, cy2 = twos.w & cy, twos.w ^= cy, fours.w |= cy2;
                                craw = z;
                            }
                        }
                    };
                    combine(0, a0, l0);
                    if (nt > 1) combine(1, a1, l1);
                    if (nt > 2) combine(2, a2, l2);
                    if (nt > 3) combine(3, a3, l3);
#pragma unroll 1
                    for (uint32_t t = 4; t < nt; ++t)  // requests with more than four plane terms
                        combine(t, __ldg(reinterpret_cast<const uint4*>(S.term_bits[t]) + w4), __ldg(reinterpret_cast<const uint4*>(S.term_lw[t]) + w4));
                    cnt += popc4(any);
                    and4(cm, anyx);  // some term that can matter is present, every part that must be present is present through such a term ...
                    if (pr.by_count && any4(cm)) {  // ... and the anchor is inside the level its number of present parts needs
                        uint4 ok = z;
#pragma unroll
                        for (int i2 = 0; i2 < (int)kFastMaxLeaves; ++i2) {
                            const int lv = S.count_lev[i2];
                            if (lv == -2) continue;
                            uint4 ex;  // anchors with exactly i2 + 1 parts present
                            if (i2 == 0) ex = make_uint4(ones.x & ~twos.x & ~fours.x, ones.y & ~twos.y & ~fours.y, ones.z & ~twos.z & ~fours.z, ones.w & ~twos.w & ~fours.w);
                            else if (i2 == 1) ex = make_uint4(twos.x & ~ones.x & ~fours.x, twos.y & ~ones.y & ~fours.y, twos.z & ~ones.z & ~fours.z, twos.w & ~ones.w & ~fours.w);
                            else if (i2 == 2) ex = make_uint4(ones.x & twos.x, ones.y & twos.y, ones.z & twos.z, ones.w & twos.w);
                            else ex = fours;
                            if (lv >= 0 && any4(ex)) and4(ex, __ldg(reinterpret_cast<const uint4*>(lev_bits + (size_t)lv * lev_words) + w4));
                            or4(ok, ex);
                        }
                        and4(cm, ok);
                    }
                    candidates(w4, cm);
                }
            }
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
            if (lane == 0) {
                if (cnt) atomicAdd(a.num_hits + q, (unsigned long long)cnt);
                st_cand += ncand, st_items += 1, st_general += swept_cold ? 1 : 0, st_sweepless += (nt == 0 || count_from_table) ? 1 : 0;
            }
            if (n_ent) {  // leave the hash table empty
                __syncwarp();
                for (uint32_t j = lane; j < kHashSlots / 4; j += 32) {
                    reinterpret_cast<uint4*>(S.hkey)[j] = make_uint4(kHashEmpty, kHashEmpty, kHashEmpty, kHashEmpty);
                    reinterpret_cast<uint4*>(S.hval)[j] = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            __syncwarp();
        }
    }
    if (qn) drain(C, &S, lane, 0, qn);
    if (lane == 0 && st_items) {
        atomicAdd(a.stats + 0, st_items);
        atomicAdd(a.stats + 1, st_cand);
        atomicAdd(a.stats + 5, st_general);    // items swept without a threshold (the request had fewer than k hits so far)
        atomicAdd(a.stats + 6, st_sweepless);  // items answered from the per-plane tile counts
    }
}

// ---------------------------------------------------------------- threshold seeds
__global__ void seed_compact_kernel(const uint32_t* __restrict__ plane_bits, uint32_t n_planes, uint32_t words, const uint32_t* __restrict__ seed_anchor, uint32_t seed_n,
                                    uint32_t seed_words, uint32_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)n_planes * seed_words) return;
    const uint32_t p = (uint32_t)(i / seed_words), cw = (uint32_t)(i % seed_words);
    const uint32_t* row = plane_bits + (size_t)p * words;
    uint32_t v = 0;
    for (uint32_t b = 0; b < 32; ++b) {
        const uint32_t j = cw * 32u + b;
        if (j >= seed_n) break;
        const uint32_t rel = seed_anchor[j];
        v |= ((row[rel >> 5] >> (rel & 31u)) & 1u) << b;
    }
    out[i] = v;
}

void launch_seed_compact(cudaStream_t st, const uint32_t* plane_bits, uint32_t n_planes, uint32_t words, const uint32_t* seed_anchor, uint32_t seed_n, uint32_t seed_words, uint32_t* out) {
    const uint64_t n = (uint64_t)n_planes * seed_words;
    if (!n) return;
    seed_compact_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(plane_bits, n_planes, words, seed_anchor, seed_n, seed_words, out);
    count_launch();
}

// Inserts the keys of the lanes in `mask` (at most kMergeGroup) into the warp's private top-k (S.merge[0, k), descending).
__device__ __forceinline__ void seed_insert(WarpScratch& S, uint32_t lane, uint32_t k, unsigned long long key, uint32_t mask) {
    const uint32_t cnt = __popc(mask);
    if (mask & (1u << lane)) S.merge[k + __popc(mask & ((1u << lane) - 1u))] = key;
    __syncwarp();
    const uint32_t n = k + cnt;
    unsigned long long mine[3];
    uint32_t rank[3];
#pragma unroll
    for (uint32_t i = 0; i < 3; ++i) {
        const uint32_t e = lane + 32u * i;
        mine[i] = e < n ? S.merge[e] : 0ull, rank[i] = 0xFFFFFFFFu;
        if (mine[i] != 0ull) {
            uint32_t r = 0;
            for (uint32_t j = 0; j < n; ++j) r += (S.merge[j] > mine[i]) || (S.merge[j] == mine[i] && j < e);
            rank[i] = r;
        }
    }
    __syncwarp();
    for (uint32_t e = lane; e < k + kMergeGroup; e += 32) S.merge[e] = 0ull;
    __syncwarp();
#pragma unroll
    for (uint32_t i = 0; i < 3; ++i)
        if (rank[i] < k) S.merge[rank[i]] = mine[i];
    __syncwarp();
}

__global__ void __launch_bounds__(kPlaneThreads, 2) plane_seed_kernel(PlaneArgs a, uint32_t n_queries) {
    extern __shared__ __align__(16) unsigned char plane_smem[];
    __shared__ CtaContext s_ctx;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    WarpScratch& S = reinterpret_cast<WarpScratch*>(plane_smem)[warp];
    const CtaContext* C = &s_ctx;
    if (tid == 0) {
        s_ctx.fast = a.fast, s_ctx.planes = a.planes;
        s_ctx.heap = a.heap, s_ctx.tau = a.tau, s_ctx.lock = a.lock, s_ctx.heap_stride = a.heap_stride, s_ctx.anchor_lo = a.anchor_lo;
    }
    __syncthreads();
    for (uint32_t q = blockIdx.x * kPlaneWarps + warp; q < n_queries; q += gridDim.x * kPlaneWarps) {
        const FastDesc* __restrict__ D = a.fast + q;
        const uint32_t flags = D->flags, nt = D->n_terms, k = D->k;
        if (!(flags & kFastOk) || !(flags & kFastBoost) || nt == 0) continue;
        const ColumnLevels* __restrict__ lev = D->fb_lev;
        const uint32_t seed_n = lev->seed_n, seed_words = lev->seed_words;
        if (seed_n == 0) continue;
        __syncwarp();
        for (uint32_t e = lane; e < k + kMergeGroup; e += 32) S.merge[e] = 0ull;
        if (lane < nt) {
            S.term_bits[lane] = lev->seed_bits + (size_t)D->plane[lane] * seed_words;
            S.term_part[lane] = D->part[lane];
        }
        __syncwarp();
        const uint32_t last_mask = __ballot_sync(0xFFFFFFFFu, lane < nt && (lane + 1u == nt || S.term_part[lane + 1u] != S.term_part[lane]));
        uint32_t qn = 0, n_eval = 0;
        const uint32_t max_eval = max(96u, 4u * k);  // the seed set is in descending boost order: the first anchors found are the best boosted
        // evaluates the queued anchors (plane contributions only) and keeps the best k
        auto flush = [&]() {
            n_eval += qn;
            unsigned long long key = 0;
            if (lane < qn) key = eval_candidate(*C, S, lane, q);
            uint32_t m = __ballot_sync(0xFFFFFFFFu, key != 0);
            while (m) {
                uint32_t part = m;
                if (__popc(part) > (int)kMergeGroup) {
                    uint32_t rest = part;
                    for (uint32_t i = 0; i < kMergeGroup; ++i) rest &= rest - 1u;
                    part &= ~rest;
                }
                m &= ~part;
                seed_insert(S, lane, k, key, part);
            }
            qn = 0;
        };
        // the set is in descending boost order: its first seed_sweep_words words (the shard's best-boosted 128 Ki anchors)
        // hold the seeds that matter, and bound the pass however large the shard is
        const uint32_t sweep_w4 = min(seed_words, a.seed_sweep_words) >> 2;
#pragma unroll 1
        for (uint32_t cw4 = lane; cw4 < sweep_w4; cw4 += 32) {
            uint4 cm = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), cur = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
            for (uint32_t t = 0; t < nt; ++t) {
                or4(cur, __ldg(reinterpret_cast<const uint4*>(S.term_bits[t]) + cw4));
                if ((last_mask >> t) & 1u) and4(cm, cur), cur = make_uint4(0u, 0u, 0u, 0u);
            }
            if (__ballot_sync(0xFFFFFFFFu, any4(cm)) == 0) continue;
#pragma unroll 1
            for (int c4 = 0; c4 < 4; ++c4) {
                uint32_t m = c4 == 0 ? cm.x : c4 == 1 ? cm.y : c4 == 2 ? cm.z : cm.w;
                while (__ballot_sync(0xFFFFFFFFu, m != 0)) {
                    bool flag = m != 0;
                    const uint32_t j = (cw4 * 4u + (uint32_t)c4) * 32u + (flag ? (uint32_t)__ffs((int)m) - 1u : 0u);
                    m &= m - 1u;
                    flag = flag && j < seed_n;
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, flag);
                    if (qn + __popc(bal) > 32) flush();
                    if (flag) {
                        const uint32_t at = qn + __popc(bal & ((1u << lane) - 1u));
                        S.cand_q[at] = q, S.cand_rel[at] = __ldg(lev->seed_anchor + j);
                        S.cand_e[0][at] = 0.0f, S.cand_e[1][at] = 0.0f, S.cand_e[2][at] = 0.0f, S.cand_e[3][at] = 0.0f;
                    }
                    qn += __popc(bal);
                    __syncwarp();
                }
            }
            if (n_eval + qn >= max_eval) break;
        }
        if (qn) flush();
        __syncwarp();
        // the k-th best lower bound: every true score is at least its plane-only score, so the request's k-th best is at least
        // this; the key just below it lets that very anchor through when the sweep evaluates it
        if (lane == 0 && S.merge[k - 1] != 0ull) atomicMax(a.tau + q, S.merge[k - 1] - 1ull);
        __syncwarp();
    }
}

size_t plane_kernel_smem() { return sizeof(WarpScratch) * kPlaneWarps; }

void launch_plane_seed(cudaStream_t st, const PlaneArgs& a, uint32_t n_queries, int n_sms) {
    if (!n_queries) return;
    static PerDeviceOnce configured;
    if (configured.first()) cudaFuncSetAttribute(plane_seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane_kernel_smem());
    plane_seed_kernel<<<(unsigned)n_sms * 2u, kPlaneThreads, plane_kernel_smem(), st>>>(a, n_queries);
    count_launch();
}

void launch_plane_eval(cudaStream_t st, const PlaneArgs& a, int n_sms) {
    if (a.group_end <= a.group_begin) return;
    static PerDeviceOnce configured;
    if (configured.first()) cudaFuncSetAttribute(plane_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane_kernel_smem());
    plane_eval_kernel<<<(unsigned)n_sms * 2u, kPlaneThreads, plane_kernel_smem(), st>>>(a);
    count_launch();
}

}  // namespace vdev
