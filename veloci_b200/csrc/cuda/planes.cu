// Plane path of the tile evaluation: requests whose frequent terms have head-term planes.
//
// Same per-anchor semantics as tiles.cu (resolve_token_to_anchor search_field.rs:400-504,
// union_hits_score set_op.rs:87-220, add_boost boost.rs:470-504, top_n_sort sort.rs:5-22),
// restricted to flat `or` requests of at most kFastMaxLeaves parts with non-negative scores
// and at most one prunable column boost.  What changes is the work per anchor:
//
//   * presence of a head term in an anchor is one bit of its plane, so the hit count of a
//     (tile, request) item is a popcount over OR-ed plane words (32 anchors per operation);
//   * the score of an anchor whose parts are all present through planes is bounded by
//     bound[n] (n = parts present) times the boost multiplier; the multiplier is bounded per
//     level of the column's nested "value >= threshold" bitmaps.  Only anchors whose bound
//     reaches the request's running k-th best are evaluated exactly (gathering the f16
//     scores of their planes and the boost value) -- the same arithmetic, in the same
//     order, as the general path, so results are bit-identical;
//   * postings of the request's other (infrequent) terms inside the tile are "entries":
//     their anchors are always evaluated exactly.
//
// A CTA stages the plane bits (and boost level bits) of one anchor tile in shared memory
// and its warps each take one (tile, request) item at a time.
#include <cuda_fp16.h>

#include "bitvec.cuh"
#include "kernels.cuh"

namespace vdev {

static const int kPlaneThreads = 512;
static const int kPlaneWarps = kPlaneThreads / 32;
static const uint32_t kQueueCap = 64;      // pending candidates of a warp: evaluated 32 at a time, one anchor per lane
static const uint32_t kMergeGroup = 8;   // keys of one request merged into its top-k under one lock round trip
static const uint32_t kHashSlots = 256;    // >= 2 * kFastMaxEntries
static const uint32_t kHashEmpty = 0xFFFFFFFFu;
static const uint32_t kEntRegs = kFastMaxEntries / 32;  // entries a lane holds

struct WarpScratch {
    FastDesc desc;                         // 160
    float mult[kBoostLevels];              // largest boost multiplier of an anchor below the level's threshold
    uint32_t ebits[256];                   // anchors of the tile that have entries
    uint32_t mbits[256];                   // ... that have more than one entry
    uint32_t hkey[kHashSlots];             // entries: index in tile | leaf << 13 -> largest score key
    uint32_t hval[kHashSlots];
    // candidates waiting for their exact evaluation; self-contained, so they outlive the item that produced them
    uint32_t cand_q[kQueueCap];            // request
    uint16_t cand_idx[kQueueCap];          // anchor index in the tile (< 8192)
    float cand_e[kFastMaxLeaves][kQueueCap];  // per part: largest entry score of the anchor (0: none)
    uint16_t ent_idx[kFastMaxEntries];     // index in tile | leaf << 13 of entry r * 32 + lane
    uint32_t ent_key[kFastMaxEntries];     // its score key
    unsigned long long merge[kFastMaxK + kMergeGroup];  // a request's top-k plus the keys merged into it in one go
    // state of the item being processed (warp-uniform)
    unsigned long long tau;                // the request's k-th best so far (0: fewer than k hits)
    float tau_score;                       // score of tau when pruning is possible, else 0
    int lev[kFastMaxLeaves];               // per count of present parts: -2 no candidates, -1 every anchor, else boost level
    uint32_t q, n_ent, tile_base_rel;
    uint32_t lev_in_smem;
    uint32_t pl_off[kFastMaxLeaves * kPartPlaneSlots];  // word offset of the request's planes in the staged tile
    // which (boost function, param, column) mult[] was computed for
    uint32_t mult_fun;
    float mult_param;
    const ColumnLevels* mult_lev;
};

struct CtaContext {  // kernel-constant values the out-of-line helpers need
    const FastDesc* fast;
    const uint16_t* score;
    size_t plane_stride;
    unsigned long long* heap;
    unsigned long long* tau;
    uint32_t* lock;
    uint32_t heap_stride, anchor_lo, W, pad;
    const uint32_t* s_bits;
    const uint32_t* s_lev;
};

// ---------------------------------------------------------------- index build
__global__ void plane_fill_kernel(const Posting* __restrict__ post, uint64_t n, uint32_t* __restrict__ bits_row, uint16_t* __restrict__ score_row, uint32_t* __restrict__ wmax_bits,
                                  uint32_t* __restrict__ bad, uint32_t anchor_lo) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t wb = 0;
    if (i < n) {
        const Posting p = post[i];
        const uint32_t rel = p.anchor - anchor_lo;
        const __half h = __float2half_rn(__fmul_rn(p.weight, 100.0f));
        // the plane must reproduce the posting's weight exactly and the list must be strictly ascending
        if (!(p.weight >= 0.0f) || __fdiv_rn(__half2float(h), 100.0f) != p.weight || (i > 0 && post[i - 1].anchor >= p.anchor)) atomicExch(bad, 1u);
        atomicOr(&bits_row[rel >> 5], 1u << (rel & 31u));
        score_row[rel] = __half_as_ushort(h);
        wb = __float_as_uint(p.weight);
    }
    for (int o = 16; o > 0; o >>= 1) wb = max(wb, __shfl_xor_sync(0xFFFFFFFFu, wb, o));
    if ((threadIdx.x & 31) == 0 && wb) atomicMax(wmax_bits, wb);
}

void launch_plane_fill(cudaStream_t st, const Posting* post, uint64_t n, uint32_t* bits_row, uint16_t* score_row, float* wmax_slot, uint32_t* bad, uint32_t anchor_lo) {
    if (!n) return;
    plane_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(post, n, bits_row, score_row, reinterpret_cast<uint32_t*>(wmax_slot), bad, anchor_lo);
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
    if (ok) {
        for (uint32_t l = 0; l < qp.n_leaves; ++l) {
            const PartPlanes pp = part_planes[leaf_part[qp.leaf_begin + l]];
            if (pp.n > kPartPlaneSlots) {
                ok = false;
                break;
            }
            d.n_planes[l] = (uint8_t)pp.n;
            for (uint32_t j = 0; j < pp.n; ++j) {
                d.plane[l][j] = (uint8_t)pp.plane[j];
                d.ts[l][j] = pp.ts[j];
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
        d.n_leaves = qp.n_leaves, d.k = qp.k;
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

// Which anchors still have to be evaluated, per number of parts present (see the file comment): lane n - 1 decides
// for n parts present.
__device__ __forceinline__ void compute_levels(WarpScratch& S, uint32_t lane, uint32_t pass_mode, int seed_level) {
    const FastDesc& D = S.desc;
    float tau_score = 0.0f;
    if (S.tau != 0) {
        const float ts = vbit::key_score((uint32_t)(S.tau >> 32));
        if (ts > 1e-30f) tau_score = ts;
    }
    if (lane < kFastMaxLeaves) {
        int lev = -1;
        if (lane >= D.n_leaves) lev = -2;
        else if (tau_score > 0.0f) {
            const float B = D.bound[lane];
            if (D.flags & kFastBoost) {
                if (B * D.fb_max_mult * 1.00001f < tau_score) lev = -2;
                else {  // deepest level whose outside cannot reach the threshold (mult[] ascends)
                    int lo = -1;
#pragma unroll
                    for (int step = (int)kBoostLevels / 2; step > 0; step >>= 1)
                        if (B * S.mult[lo + step] < tau_score) lo += step;
                    if (lo + 1 < (int)kBoostLevels && B * S.mult[lo + 1] < tau_score) lo += 1;
                    lev = lo;
                }
            } else if (B * 1.00001f < tau_score) {
                lev = -2;
            }
        }
        if (pass_mode == 1 && lev != -2) lev = (D.flags & kFastBoost) ? max(lev, seed_level) : -2;  // seed pass: inside the seed level only
        S.lev[lane] = lev;
    }
    if (lane == 0) S.tau_score = tau_score;
    __syncwarp();
}

// Inserts one hit into the request's heap (sorted, k slots) under its lock, if it still beats the k-th best.
__device__ __noinline__ void merge_group(const CtaContext* C, WarpScratch* Sp, uint32_t lane, uint32_t q, unsigned long long comp, uint32_t group) {
    // Merges the order keys of the lanes in `group` (at most kMergeGroup, all of request q, each above the threshold it was
    // compared with) into the request's top-k under its lock: one lock round trip for all of them.
    WarpScratch& S = *Sp;
    const uint32_t k = C->fast[q].k;
    unsigned long long* heap = C->heap + (size_t)q * C->heap_stride;
    if (lane == 0) {
        while (atomicCAS(C->lock + q, 0u, 1u) != 0u) __nanosleep(32);
        __threadfence();
    }
    __syncwarp();
    for (uint32_t i = lane; i < k; i += 32) S.merge[i] = __ldcg(heap + i);
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

__device__ __forceinline__ uint32_t hash_slot(uint32_t code) { return (code * 0x9E3779B1u) >> 24; }

__device__ __forceinline__ void hash_insert(WarpScratch& S, uint32_t code, uint32_t key) {
    uint32_t h = hash_slot(code);
    while (true) {
        const uint32_t old = atomicCAS(&S.hkey[h], kHashEmpty, code);
        if (old == kHashEmpty || old == code) break;
        h = (h + 1u) & (kHashSlots - 1u);
    }
    atomicMax(&S.hval[h], key);
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

// Exact score of candidate `c` of the warp's list: its order key.  The plane bits of the tile are still staged.
__device__ __forceinline__ unsigned long long eval_candidate(const CtaContext& C, const WarpScratch& S, uint32_t c, uint32_t q) {
    const FastDesc& D = C.fast[q];
    const uint32_t idx = S.cand_idx[c];
    const uint32_t w = idx >> 5, bit = 1u << (idx & 31u);
    const uint32_t rel = S.tile_base_rel + idx;
    const uint32_t L = D.n_leaves, W = C.W;
    float sum = 0.0f, nd = 0.0f, v0 = 0.0f;
#pragma unroll 1
    for (uint32_t l = 0; l < L; ++l) {
        float v = S.cand_e[l][c];
        const uint32_t np = D.n_planes[l];
#pragma unroll 1
        for (uint32_t j = 0; j < np; ++j) {
            const uint32_t p = D.plane[l][j];
            if (C.s_bits[p * W + w] & bit) {
                const unsigned short h = __ldg(C.score + p * C.plane_stride + rel);
                const float wgt = __fdiv_rn(__half2float(__ushort_as_half(h)), 100.0f);  // el.score.to_f32() / 100.0 (search_field.rs:426)
                v = fmaxf(v, D.ts[l][j] * wgt);
            }
        }
        if (v >= 0.00001f) nd += 1.0f;
        sum += v;
        if (l == 0) v0 = v;
    }
    float score = (L == 1 && !(D.flags & kFastUnion1)) ? v0 : sum * nd * nd;
    const uint32_t anchor = C.anchor_lo + rel;
    if (D.flags & kFastBoost) {
        if (anchor < D.fb_n) {
            const uint32_t bits = __ldg(D.fb_col + anchor);
            if (bits != kNoValue) {
                const float x = __uint_as_float(bits) + D.fb_param;
                switch (D.fb_fun) {
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
// a full group of 32 when there is one.
__device__ __forceinline__ void enqueue(const CtaContext* C, WarpScratch& S, uint32_t lane, bool flag, uint32_t idx, float e0, float e1, float e2, float e3, uint32_t& qn,
                                        uint32_t& ncand) {
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, flag);
    if (!m) return;
    if (flag) {
        const uint32_t at = qn + __popc(m & ((1u << lane) - 1u));
        S.cand_q[at] = S.q, S.cand_idx[at] = (uint16_t)idx;
        S.cand_e[0][at] = e0, S.cand_e[1][at] = e1, S.cand_e[2][at] = e2, S.cand_e[3][at] = e3;
    }
    qn += __popc(m);
    ncand += __popc(m);
    __syncwarp();
    if (qn >= 32) {
        qn -= 32;
        drain(C, &S, lane, qn, 32);
    }
}

struct ItemLoad {  // what the pipeline fetches one item ahead (per lane)
    uint4 d;                      // lane < 10: 16-byte word of the request's FastDesc
    unsigned long long tau;
    uint2 ent[kEntRegs];          // entries lane, lane + 32, ... of the item (anchor, key)
    uint32_t ent_leaf;            // 2 bits per held entry
};

__device__ __forceinline__ void load_item(const PlaneArgs& a, const FastItem& it, uint32_t lane, ItemLoad& o) {
    o.d = make_uint4(0u, 0u, 0u, 0u);
    if (lane < 10) o.d = __ldg(reinterpret_cast<const uint4*>(a.fast + it.q) + lane);
    o.tau = __ldcg(a.tau + it.q);
    o.ent_leaf = 0;
    const uint32_t n0 = it.n[0], n1 = n0 + it.n[1], n2 = n1 + it.n[2], n3 = n2 + it.n[3];
#pragma unroll
    for (uint32_t r = 0; r < kEntRegs; ++r) {
        const uint32_t j = r * 32u + lane;
        o.ent[r] = make_uint2(0u, 0u);
        if (j < n3) {
            const uint32_t l = (j >= n0) + (j >= n1) + (j >= n2);
            const uint32_t at = l == 0 ? it.begin[0] + j : l == 1 ? it.begin[1] + (j - n0) : l == 2 ? it.begin[2] + (j - n1) : it.begin[3] + (j - n2);
            o.ent[r] = __ldg(reinterpret_cast<const uint2*>(a.sparse + at));
            o.ent_leaf |= l << (2u * r);
        }
    }
}

__device__ __forceinline__ FastItem load_record(const PlaneArgs& a, uint32_t at, uint32_t iend) {
    FastItem it;
    it.q = 0, it.pad = 0;
#pragma unroll
    for (int l = 0; l < (int)kFastMaxLeaves; ++l) it.n[l] = 0, it.begin[l] = 0;
    if (at < iend) {
        const uint4* p = reinterpret_cast<const uint4*>(a.items + at);
        const uint4 lo = __ldg(p), hi = __ldg(p + 1);
        it.q = lo.x, it.n[0] = (uint16_t)lo.y, it.n[1] = (uint16_t)(lo.y >> 16), it.n[2] = (uint16_t)lo.z, it.n[3] = (uint16_t)(lo.z >> 16);
        it.begin[0] = lo.w, it.begin[1] = hi.x, it.begin[2] = hi.y, it.begin[3] = hi.z;
    }
    return it;
}

__device__ __forceinline__ FastItem shfl_item(const FastItem& mine, int src) {
    FastItem it;
    it.q = __shfl_sync(0xFFFFFFFFu, mine.q, src);
    const uint32_t n01 = __shfl_sync(0xFFFFFFFFu, (uint32_t)mine.n[0] | ((uint32_t)mine.n[1] << 16), src);
    const uint32_t n23 = __shfl_sync(0xFFFFFFFFu, (uint32_t)mine.n[2] | ((uint32_t)mine.n[3] << 16), src);
    it.n[0] = (uint16_t)n01, it.n[1] = (uint16_t)(n01 >> 16), it.n[2] = (uint16_t)n23, it.n[3] = (uint16_t)(n23 >> 16);
#pragma unroll
    for (int l = 0; l < (int)kFastMaxLeaves; ++l) it.begin[l] = __shfl_sync(0xFFFFFFFFu, mine.begin[l], src);
    it.pad = 0;
    return it;
}

__device__ __forceinline__ uint32_t comp4(const uint4& v, int c) { return c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w; }

__global__ void __launch_bounds__(kPlaneThreads, 1) plane_eval_kernel(PlaneArgs a) {
    extern __shared__ __align__(16) uint32_t plane_smem[];
    __shared__ uint32_t s_unit;
    __shared__ uint32_t s_pcount[kMaxPlanes];  // anchors of the staged tile per plane
    __shared__ CtaContext s_ctx;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t W = 1u << (a.tile_log2 - 5);
    uint32_t* s_bits = plane_smem;
    uint32_t* s_lev = s_bits + a.planes.n_planes * W;
    WarpScratch& S = reinterpret_cast<WarpScratch*>(s_lev + kBoostLevels * W)[warp];
    const CtaContext* C = &s_ctx;
    if (tid == 0) {
        s_ctx.fast = a.fast, s_ctx.score = a.planes.score, s_ctx.plane_stride = (size_t)a.planes.words * 32u;
        s_ctx.heap = a.heap, s_ctx.tau = a.tau, s_ctx.lock = a.lock, s_ctx.heap_stride = a.heap_stride;
        s_ctx.anchor_lo = a.anchor_lo, s_ctx.W = W, s_ctx.pad = 0, s_ctx.s_bits = s_bits, s_ctx.s_lev = s_lev;
    }
    for (uint32_t i = lane; i < 256; i += 32) S.ebits[i] = 0, S.mbits[i] = 0;
    for (uint32_t i = lane; i < kHashSlots; i += 32) S.hkey[i] = kHashEmpty, S.hval[i] = 0;
    if (lane == 0) S.mult_lev = nullptr, S.mult_fun = 0xFFFFFFFFu, S.mult_param = 0.0f;
    unsigned long long st_cand = 0, st_items = 0, st_general = 0, st_sweepless = 0;  // lane 0

    while (true) {
        __syncthreads();
        if (tid == 0) s_unit = (uint32_t)atomicAdd(a.work_counter, 1ull);
        __syncthreads();
        const uint32_t unit = s_unit;
        if (unit >= a.n_units) break;
        const uint32_t t = a.tile_begin + unit / a.chunks_per_tile, c = unit % a.chunks_per_tile;
        const uint32_t tb = a.tile_item_begin[t], te = a.tile_item_begin[t + 1];
        if (te - tb <= c * a.unit_items) continue;
        const uint32_t ibeg = tb + c * a.unit_items, iend = min(te, ibeg + a.unit_items);
        {   // stage the tile's plane bits and boost level bits
            const uint32_t w4 = W >> 2;
            const uint32_t n4 = a.planes.n_planes * w4;
            for (uint32_t i = tid; i < n4; i += kPlaneThreads) {
                const uint32_t p = i / w4, j = i % w4;
                reinterpret_cast<uint4*>(s_bits)[i] = __ldg(reinterpret_cast<const uint4*>(a.planes.bits + (size_t)p * a.planes.words + (size_t)t * W) + j);
            }
            if (a.lev_dev != nullptr) {
                const uint32_t l4 = kBoostLevels * w4;
                for (uint32_t i = tid; i < l4; i += kPlaneThreads) {
                    const uint32_t p = i / w4, j = i % w4;
                    reinterpret_cast<uint4*>(s_lev)[i] = __ldg(reinterpret_cast<const uint4*>(a.lev_hdr.bits + (size_t)p * a.lev_hdr.words + (size_t)t * W) + j);
                }
            }
        }
        __syncthreads();
        for (uint32_t p = warp; p < a.planes.n_planes; p += kPlaneWarps) {  // per-plane hit count of the tile
            uint32_t c = 0;
            for (uint32_t w = lane; w < W; w += 32) c += __popc(s_bits[p * W + w]);
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
            if (lane == 0) s_pcount[p] = c;
        }
        __syncthreads();
        const uint32_t tile_base_rel = t << a.tile_log2;
        const uint32_t tile_base = a.anchor_lo + tile_base_rel;

        // The warp takes items ibeg + warp, + kPlaneWarps, ...: records are fetched 32 items at a time (one per lane),
        // descriptor + threshold + entries one item ahead of the evaluation.
        const uint32_t first = ibeg + warp;
        if (first >= iend) continue;
        const uint32_t n_mine = (iend - first + kPlaneWarps - 1) / kPlaneWarps;
        FastItem mine = load_record(a, first + lane * kPlaneWarps, iend);
        FastItem it_cur = shfl_item(mine, 0), it_nxt = it_cur;
        ItemLoad cur, nxt;
        load_item(a, it_cur, lane, cur);
        nxt = cur;
        uint32_t qn = 0;  // candidates of this warp waiting for their exact evaluation
#pragma unroll 1
        for (uint32_t i = 0; i < n_mine; ++i) {
            const bool has_next = i + 1 < n_mine;
            if (has_next) {
                if (((i + 1) & 31u) == 0) mine = load_record(a, first + (i + 1 + lane) * kPlaneWarps, iend);
                it_nxt = shfl_item(mine, (int)((i + 1) & 31u));
                load_item(a, it_nxt, lane, nxt);
            }

            // ---- item (t, it_cur.q)
            __syncwarp();
            if (lane < 10) reinterpret_cast<uint4*>(&S.desc)[lane] = cur.d;
            const uint32_t n_ent = (uint32_t)it_cur.n[0] + it_cur.n[1] + it_cur.n[2] + it_cur.n[3];
            __syncwarp();
            const FastDesc& D = S.desc;
            const uint32_t L = D.n_leaves;
            if (lane == 0) {
                S.q = it_cur.q, S.tau = cur.tau, S.n_ent = n_ent, S.tile_base_rel = tile_base_rel;
                S.lev_in_smem = D.fb_lev == a.lev_dev ? 1u : 0u;
            }
            if (lane < kFastMaxLeaves * kPartPlaneSlots) S.pl_off[lane] = (uint32_t)D.plane[lane / kPartPlaneSlots][lane % kPartPlaneSlots] * W;
            const bool need_mult = (D.flags & kFastBoost) && (S.mult_lev != D.fb_lev || S.mult_fun != D.fb_fun || S.mult_param != D.fb_param);
            __syncwarp();
            if (need_mult) {
                if (lane < kBoostLevels) {
                    const float m = boost_mult(D.fb_fun, __ldg(&D.fb_lev->thr[lane]) + D.fb_param);
                    S.mult[lane] = fmaxf(m, 0.0f) * 1.00001f + 1e-6f;
                }
                if (lane == 0) S.mult_lev = D.fb_lev, S.mult_fun = D.fb_fun, S.mult_param = D.fb_param;
            }
            uint32_t own = 0;  // bit r: this lane's entry r is the first of its anchor
            bool any_multi = false;
            if (n_ent) {
#pragma unroll
                for (uint32_t r = 0; r < kEntRegs; ++r) {
                    if (r * 32u + lane < n_ent) {
                        const uint32_t idx = cur.ent[r].x - tile_base, l = (cur.ent_leaf >> (2u * r)) & 3u;
                        const uint32_t bit = 1u << (idx & 31u);
                        if (!(atomicOr(&S.ebits[idx >> 5], bit) & bit)) own |= 1u << r;
                        else atomicOr(&S.mbits[idx >> 5], bit);
                        S.ent_idx[r * 32u + lane] = (uint16_t)(idx | (l << 13));
                        S.ent_key[r * 32u + lane] = cur.ent[r].y;
                    }
                }
                __syncwarp();
                // only anchors with several entries need the hash (per-part maximum over their entries)
#pragma unroll
                for (uint32_t r = 0; r < kEntRegs; ++r) {
                    if (r * 32u + lane < n_ent && a.pass_mode != 1) {
                        const uint32_t idx = cur.ent[r].x - tile_base, l = (cur.ent_leaf >> (2u * r)) & 3u;
                        if (S.mbits[idx >> 5] & (1u << (idx & 31u))) {
                            hash_insert(S, idx | (l << 13), cur.ent[r].y);
                            any_multi = true;
                        }
                    }
                }
                any_multi = __ballot_sync(0xFFFFFFFFu, any_multi) != 0;
            }
            __syncwarp();
            compute_levels(S, lane, a.pass_mode, (int)a.seed_level);
            const bool seed_pass = a.pass_mode == 1;
            const bool skip_seeded = a.pass_mode == 2 && t < a.seeded_tiles && (D.flags & kFastBoost) != 0;
            const uint32_t np0 = D.n_planes[0], np1 = D.n_planes[1], np2 = D.n_planes[2], np3 = D.n_planes[3];
            // word offset of each part's first plane (most parts match at most one head term)
            const uint32_t po0 = (uint32_t)D.plane[0][0] * W, po1 = (uint32_t)D.plane[1][0] * W, po2 = (uint32_t)D.plane[2][0] * W, po3 = (uint32_t)D.plane[3][0] * W;
            // presence word `w` of part l
            auto part_word = [&](uint32_t l, uint32_t np, uint32_t po, uint32_t w) -> uint32_t {
                if (np == 0) return 0u;
                uint32_t x = s_bits[po + w];
#pragma unroll 1
                for (uint32_t j = 1; j < np; ++j) x |= s_bits[S.pl_off[l * kPartPlaneSlots + j] + w];
                return x;
            };
            auto part_words4 = [&](uint32_t l, uint32_t np, uint32_t po, uint32_t w4) -> uint4 {
                if (np == 0) return make_uint4(0u, 0u, 0u, 0u);
                uint4 x = reinterpret_cast<const uint4*>(s_bits + po)[w4];
#pragma unroll 1
                for (uint32_t j = 1; j < np; ++j) {
                    const uint4 v = reinterpret_cast<const uint4*>(s_bits + S.pl_off[l * kPartPlaneSlots + j])[w4];
                    x.x |= v.x, x.y |= v.y, x.z |= v.z, x.w |= v.w;
                }
                return x;
            };
            const bool lev_in_smem = S.lev_in_smem != 0;
            const uint32_t* lev_glob = (D.flags & kFastBoost) ? D.fb_lev->bits + (size_t)t * W : nullptr;
            const uint32_t lev_words = lev_in_smem ? a.lev_hdr.words : (D.flags & kFastBoost) ? D.fb_lev->words : 0u;
            uint32_t cnt = 0, ncand = 0;
            // A request with at most one head term in total needs no sweep once anchors with a single part present cannot
            // reach the threshold any more: its hit count is the plane's count of the tile plus the entry anchors outside it.
            const uint32_t np_total = np0 + np1 + np2 + np3;
            const bool sweepless = np_total == 0 || (np_total == 1 && S.lev[0] == -2 && !seed_pass);
            const uint32_t only_po = np0 ? po0 : np1 ? po1 : np2 ? po2 : po3;  // the single plane (np_total == 1)
            if (sweepless && np_total == 1 && lane == 0) cnt = s_pcount[only_po / W];
            // boost level of the seed pass, for the candidates of pass 2 to skip
            const uint32_t* seed_bits = lev_in_smem ? s_lev + a.seed_level * W : (skip_seeded ? lev_glob + (size_t)a.seed_level * lev_words : nullptr);

            // anchors with entries: bounded one by one (entry scores are known exactly, plane parts by their bound)
            if (n_ent && !seed_pass) {
#pragma unroll 1
                for (uint32_t r = 0; r < kEntRegs; ++r) {
                    if (r * 32u >= n_ent) break;
                    const bool mine_r = (own >> r) & 1u;
                    const uint32_t code = mine_r ? S.ent_idx[r * 32u + lane] : 0u;
                    const uint32_t idx = code & 0x1FFFu, el = code >> 13;
                    bool cand = mine_r;
                    const float tau_score = S.tau_score;
                    float e0 = 0.0f, e1 = 0.0f, e2 = 0.0f, e3 = 0.0f;
                    if (mine_r) {
                        const uint32_t w = idx >> 5, bit = 1u << (idx & 31u);
                        if (sweepless && !(np_total == 1 && (s_bits[only_po + w] & bit))) cnt += 1;  // a hit the plane count does not include
                        const bool multi = (S.mbits[w] & bit) != 0;
                        float sum_ub = 0.0f;
                        uint32_t n = 0;
                        // the anchor's only entry is this lane's own; anchors with several entries look them up per part
                        const uint32_t own_ev = multi ? hash_lookup(S, code) : S.ent_key[r * 32u + lane];
                        auto part = [&](uint32_t l, uint32_t np, uint32_t po, float& e) {
                            uint32_t ev = l == el ? own_ev : 0u;
                            if (multi && l != el) ev = hash_lookup(S, idx | (l << 13));
                            e = ev ? __uint_as_float(ev & 0x7FFFFFFFu) : 0.0f;
                            const bool by_plane = (part_word(l, np, po, w) & bit) != 0;
                            if (ev || by_plane) {
                                n += 1;
                                sum_ub += fmaxf(e, by_plane ? D.ub[l] : 0.0f);
                            }
                        };
                        part(0, np0, po0, e0);
                        if (L > 1) part(1, np1, po1, e1);
                        if (L > 2) part(2, np2, po2, e2);
                        if (L > 3) part(3, np3, po3, e3);
                        const float B = (L == 1 ? sum_ub : sum_ub * (float)(n * n)) * 1.00001f;
                        if (tau_score > 0.0f) {
                            if (D.flags & kFastBoost) {
                                if (B * D.fb_max_mult * 1.00001f < tau_score) cand = false;
                                else {
                                    // deepest level whose outside cannot reach the threshold (mult[] ascends): the anchor must be inside it
                                    int lo = -1;
#pragma unroll
                                    for (int step = (int)kBoostLevels / 2; step > 0; step >>= 1)
                                        if (B * S.mult[lo + step] < tau_score) lo += step;
                                    if (lo + 1 < (int)kBoostLevels && B * S.mult[lo + 1] < tau_score) lo += 1;
                                    if (lo >= 0) {
                                        const uint32_t lw = lev_in_smem ? s_lev[(uint32_t)lo * W + w] : __ldg(lev_glob + (size_t)lo * lev_words + w);
                                        cand = (lw & bit) != 0;
                                    }
                                }
                            } else if (B * 1.00001f < tau_score) {
                                cand = false;
                            }
                        }
                    }
                    enqueue(C, S, lane, cand, idx, e0, e1, e2, e3, qn, ncand);
                }
            }

            // plane sweep: four 32-anchor words per lane and step
            const int lev_top = S.lev[L - 1];
            bool lower_dead = a.force_general == 0;
#pragma unroll
            for (int i2 = 0; i2 < (int)kFastMaxLeaves - 1; ++i2)
                if ((uint32_t)i2 + 1 < L && S.lev[i2] != -2) lower_dead = false;
#pragma unroll 1
            for (uint32_t w4 = lane; w4 < (sweepless ? 0u : (W >> 2)); w4 += 32) {
                uint4 e = make_uint4(0u, 0u, 0u, 0u);
                if (n_ent) e = reinterpret_cast<const uint4*>(S.ebits)[w4];
                uint32_t cm[4] = {0u, 0u, 0u, 0u};
                if (lower_dead && S.lev[L - 1] == lev_top) {
                    // converged threshold: only anchors with every part present can still matter -> OR for the count, AND for the candidates
                    uint4 any = part_words4(0, np0, po0, w4), all = any;
                    if (L > 1) {
                        const uint4 x = part_words4(1, np1, po1, w4);
                        any.x |= x.x, any.y |= x.y, any.z |= x.z, any.w |= x.w;
                        all.x &= x.x, all.y &= x.y, all.z &= x.z, all.w &= x.w;
                    }
                    if (L > 2) {
                        const uint4 x = part_words4(2, np2, po2, w4);
                        any.x |= x.x, any.y |= x.y, any.z |= x.z, any.w |= x.w;
                        all.x &= x.x, all.y &= x.y, all.z &= x.z, all.w &= x.w;
                    }
                    if (L > 3) {
                        const uint4 x = part_words4(3, np3, po3, w4);
                        any.x |= x.x, any.y |= x.y, any.z |= x.z, any.w |= x.w;
                        all.x &= x.x, all.y &= x.y, all.z &= x.z, all.w &= x.w;
                    }
                    cnt += __popc(any.x | e.x) + __popc(any.y | e.y) + __popc(any.z | e.z) + __popc(any.w | e.w);
                    if (lev_top == -2) continue;
                    if (lev_top >= 0) {
                        uint4 lw;
                        if (lev_in_smem) lw = reinterpret_cast<const uint4*>(s_lev + (uint32_t)lev_top * W)[w4];
                        else lw = __ldg(reinterpret_cast<const uint4*>(lev_glob + (size_t)lev_top * lev_words) + w4);
                        all.x &= lw.x, all.y &= lw.y, all.z &= lw.z, all.w &= lw.w;
                    }
                    cm[0] = all.x & ~e.x, cm[1] = all.y & ~e.y, cm[2] = all.z & ~e.z, cm[3] = all.w & ~e.w;
                    if (skip_seeded) {
                        const uint4 sw = reinterpret_cast<const uint4*>(seed_bits)[w4];
                        cm[0] &= ~sw.x, cm[1] &= ~sw.y, cm[2] &= ~sw.z, cm[3] &= ~sw.w;
                    }
                } else {
                    uint4 ones = make_uint4(0u, 0u, 0u, 0u), twos = ones, fours = ones;
                    auto add_part = [&](uint32_t l, uint32_t np, uint32_t po) {
                        const uint4 x = part_words4(l, np, po, w4);
                        uint32_t cy, cy2;
                        cy = ones.x & x.x, ones.x ^= x.x, cy2 = twos.x & cy, twos.x ^= cy, fours.x |= cy2;
                        cy = ones.y & x.y, ones.y ^= x.y, cy2 = twos.y & cy, twos.y ^= cy, fours.y |= cy2;
                        cy = ones.z & x.z, ones.z ^= x.z, cy2 = twos.z & cy, twos.z ^= cy, fours.z |= cy2;
                        cy = ones.w & x.w, ones.w ^= x.w, cy2 = twos.w & cy, twos.w ^= cy, fours.w |= cy2;
                    };
                    add_part(0, np0, po0);
                    if (L > 1) add_part(1, np1, po1);
                    if (L > 2) add_part(2, np2, po2);
                    if (L > 3) add_part(3, np3, po3);
                    cnt += __popc(ones.x | twos.x | fours.x | e.x) + __popc(ones.y | twos.y | fours.y | e.y) + __popc(ones.z | twos.z | fours.z | e.z) + __popc(ones.w | twos.w | fours.w | e.w);
#pragma unroll
                    for (int i2 = 0; i2 < (int)kFastMaxLeaves; ++i2) {
                        const int lv = S.lev[i2];
                        if (lv == -2) continue;
                        uint4 ex;  // anchors with exactly i2 + 1 parts present
                        if (i2 == 0) ex = make_uint4(ones.x & ~twos.x & ~fours.x, ones.y & ~twos.y & ~fours.y, ones.z & ~twos.z & ~fours.z, ones.w & ~twos.w & ~fours.w);
                        else if (i2 == 1) ex = make_uint4(twos.x & ~ones.x & ~fours.x, twos.y & ~ones.y & ~fours.y, twos.z & ~ones.z & ~fours.z, twos.w & ~ones.w & ~fours.w);
                        else if (i2 == 2) ex = make_uint4(ones.x & twos.x, ones.y & twos.y, ones.z & twos.z, ones.w & twos.w);
                        else ex = fours;
                        if (lv >= 0) {
                            uint4 lw;
                            if (lev_in_smem) lw = reinterpret_cast<const uint4*>(s_lev + (uint32_t)lv * W)[w4];
                            else lw = __ldg(reinterpret_cast<const uint4*>(lev_glob + (size_t)lv * lev_words) + w4);
                            ex.x &= lw.x, ex.y &= lw.y, ex.z &= lw.z, ex.w &= lw.w;
                        }
                        cm[0] |= ex.x, cm[1] |= ex.y, cm[2] |= ex.z, cm[3] |= ex.w;
                    }
                    cm[0] &= ~e.x, cm[1] &= ~e.y, cm[2] &= ~e.z, cm[3] &= ~e.w;  // anchors with entries were handled above
                    if (skip_seeded) {
                        const uint4 sw = reinterpret_cast<const uint4*>(seed_bits)[w4];
                        cm[0] &= ~sw.x, cm[1] &= ~sw.y, cm[2] &= ~sw.z, cm[3] &= ~sw.w;
                    }
                }
                if (__ballot_sync(0xFFFFFFFFu, (cm[0] | cm[1] | cm[2] | cm[3]) != 0) == 0) continue;
#pragma unroll 1
                for (int c4 = 0; c4 < 4; ++c4) {
                    uint32_t m = c4 == 0 ? cm[0] : c4 == 1 ? cm[1] : c4 == 2 ? cm[2] : cm[3];
                    while (__ballot_sync(0xFFFFFFFFu, m != 0)) {
                        const bool flag = m != 0;
                        const uint32_t idx = (w4 * 4u + (uint32_t)c4) * 32u + (uint32_t)__ffs((int)m) - 1u;
                        m &= m - 1u;
                        enqueue(C, S, lane, flag, idx, 0.0f, 0.0f, 0.0f, 0.0f, qn, ncand);
                    }
                }
            }
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
            if (lane == 0) {
                if (cnt && !seed_pass) atomicAdd(a.num_hits + it_cur.q, (unsigned long long)cnt);
                st_cand += ncand, st_items += 1, st_general += (!sweepless && !lower_dead) ? 1 : 0, st_sweepless += sweepless ? 1 : 0;
            }
            if (n_ent) {  // leave the entry bitmap and the hash table empty
                __syncwarp();
#pragma unroll
                for (uint32_t r = 0; r < kEntRegs; ++r)
                    if (r * 32u + lane < n_ent) S.ebits[(cur.ent[r].x - tile_base) >> 5] = 0, S.mbits[(cur.ent[r].x - tile_base) >> 5] = 0;
                if (any_multi)
                    for (uint32_t j = lane; j < kHashSlots / 4; j += 32) {
                        reinterpret_cast<uint4*>(S.hkey)[j] = make_uint4(kHashEmpty, kHashEmpty, kHashEmpty, kHashEmpty);
                        reinterpret_cast<uint4*>(S.hval)[j] = make_uint4(0u, 0u, 0u, 0u);
                    }
            }
            cur = nxt, it_cur = it_nxt;
        }
        if (qn) drain(C, &S, lane, 0, qn);  // the tile's bits are still staged
    }
    if (lane == 0 && st_items) {
        atomicAdd(a.stats + 0, st_items);
        atomicAdd(a.stats + 1, st_cand);
        atomicAdd(a.stats + 5, st_general);    // items swept with the exact per-anchor part count (threshold not converged)
        atomicAdd(a.stats + 6, st_sweepless);  // items answered from the per-plane counts
    }
}

static size_t plane_smem_bytes(uint32_t tile_log2, uint32_t n_planes) {
    const size_t W = (size_t)1 << (tile_log2 - 5);
    return ((size_t)n_planes + kBoostLevels) * W * 4 + sizeof(WarpScratch) * kPlaneWarps;
}

size_t plane_kernel_smem(uint32_t tile_log2, uint32_t n_planes) {
    if (tile_log2 < 12 || tile_log2 > kPlaneTileLog2 || n_planes > kMaxPlanes) return 0;
    const size_t need = plane_smem_bytes(tile_log2, n_planes);
    return need + 1024 <= 227 * 1024 ? need : 0;
}

void launch_plane_eval(cudaStream_t st, const PlaneArgs& a, int n_sms) {
    if (a.n_units == 0) return;
    const size_t smem = plane_kernel_smem(a.tile_log2, a.planes.n_planes);
    static PerDeviceOnce configured;
    if (configured.first()) cudaFuncSetAttribute(plane_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    unsigned blocks = (unsigned)n_sms;
    if (blocks > a.n_units) blocks = a.n_units;
    plane_eval_kernel<<<blocks, kPlaneThreads, smem, st>>>(a);
    count_launch();
}

}  // namespace vdev
