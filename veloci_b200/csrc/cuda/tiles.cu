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
// tile is one contiguous slice: the CTA streams those slices with coalesced loads
// and scatters score keys into one shared-memory array per search part (direct
// mapped: index = anchor - tile start, so no sorting and no hashing collisions).
// A dense list touches every anchor at most once, so its scatter needs no atomics;
// only the pre-bucketed sparse lists use shared-memory atomicMax.  The epilogue
// walks the tile once: tree evaluation, boost column gather (coalesced, the tile
// is a contiguous range of the column), hit count, and a threshold test against
// the request's running k-th best; the few survivors are merged into the
// request's top-k heap in global memory under a per-request lock.
//
// Work items are ordered tile-major (all requests of tile 0, then tile 1, ...), so
// the posting slices of a tile are re-read from L2, not HBM, by the many requests
// that share frequent terms.
#include <cuda_fp16.h>

#include "bitvec.cuh"
#include "kernels.cuh"

namespace vdev {

static const int kTileThreads = 512;
static const uint32_t kSurvivorCap = 1024;

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

// Generic request tree, postfix.  Returns presence; score in `out`.
__device__ bool eval_program(const uint32_t* __restrict__ prog, uint32_t len, const uint32_t* arr, uint32_t tile, uint32_t idx, float& out) {
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
    float param, max_mult;
};

// Everything after the request tree for one present anchor: boosts, threshold, survivor list.
// Returns the key to leave in arr[0][idx] (non-zero only for a survivor deferred to the next merge round).
__device__ __forceinline__ uint32_t finish_anchor(const TileArgs& a, const ItemCtx& c, uint32_t anchor, float score, uint32_t* s_nsurv, unsigned long long* s_list) {
    if (c.qp.n_boosts) {
        if (c.fast_boost) {
            // the boost can only shrink `score * max_mult`: skip the gather when even that cannot beat the k-th best so far
            if (c.can_prune && score >= 0.0f) {
                const float bound = score * c.max_mult;
                if ((((unsigned long long)vbit::score_key(bound) << 32) | 0xFFFFFFFFull) <= c.tau) return 0;
            }
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
                if (skip || anchor >= bs.n) continue;
                const uint32_t bits = __ldg(bs.column + anchor);
                if (bits != kNoValue) score = apply_boost_step(bs, score, __uint_as_float(bits));
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
        score = L == 1 ? vbit::key_score(arr[idx]) : sum * nd * nd;
    } else {
        present = eval_program(a.prog + c.qp.prog_begin, c.qp.prog_len, arr, c.tile, idx, score);
    }
    for (uint32_t l = 0; l < L; ++l) arr[l * c.tile + idx] = 0;
    if (!present) return 0;
    const uint32_t keep = finish_anchor(a, c, c.tile_base + idx, score, s_nsurv, s_list);
    if (keep) arr[idx] = keep;
    return 1;
}

__global__ void __launch_bounds__(kTileThreads) tile_eval_kernel(TileArgs a) {
    extern __shared__ __align__(16) uint32_t arr[];  // [max_leaves][tile], all zero between items
    __shared__ unsigned long long s_item;
    __shared__ uint32_t s_npresent, s_nsurv;
    __shared__ unsigned long long s_list[kSurvivorCap];
    __shared__ unsigned long long s_heap[kMaxK];
    __shared__ unsigned long long s_out[kMaxK];
    __shared__ uint32_t s_claim[1024];  // one bit per anchor of the tile (tiles up to 2^15)

    const uint32_t tid = threadIdx.x;
    const uint32_t tile = 1u << a.tile_log2;
    unsigned long long cta_postings = 0;  // thread 0 only
    {
        uint4* p4 = reinterpret_cast<uint4*>(arr);
        const uint32_t n4 = (a.max_leaves * tile) >> 2;
        for (uint32_t i = tid; i < n4; i += kTileThreads) p4[i] = make_uint4(0u, 0u, 0u, 0u);
    }

    while (true) {
        __syncthreads();  // everyone is done with the previous item's shared state
        if (tid == 0) s_item = atomicAdd(a.work_counter, 1ull), s_npresent = 0, s_nsurv = 0;
        __syncthreads();
        const unsigned long long item = s_item;
        if (item >= a.n_items) break;
        const uint32_t t = (uint32_t)(item / a.n_queries), q = (uint32_t)(item % a.n_queries);
        ItemCtx c;
        c.qp = a.queries[q];
        if (!c.qp.active || c.qp.n_leaves == 0) continue;
        const QueryProgram& qp = c.qp;
        const uint32_t L = qp.n_leaves;
        const uint64_t tile_base64 = (uint64_t)a.anchor_lo + ((uint64_t)t << a.tile_log2);
        const uint32_t tile_base = (uint32_t)tile_base64;
        const uint32_t tile_n = (uint32_t)min((uint64_t)tile, (uint64_t)a.anchor_hi - tile_base64);
        c.tile = tile, c.tile_base = tile_base;

        // (0) how many postings of this request fall into the tile (uniform across the CTA)
        uint32_t max_dense = 0, npost = 0;
        for (uint32_t l = 0; l < L; ++l) {
            const PartSlices ps = a.slices[a.leaf_part[qp.leaf_begin + l]];
            max_dense = max(max_dense, ps.n_dense);
            for (uint32_t r = 0; r < ps.n_dense; ++r) {
                const uint32_t* trow = a.toff + (size_t)a.g_row[ps.m_begin + r] * (a.n_tiles + 1);
                npost += trow[t + 1] - trow[t];
            }
            if (ps.n_match != ps.n_dense) {
                const uint32_t* brow = a.bucket + (size_t)ps.sparse_row * (a.n_tiles + 1);
                npost += brow[t + 1] - brow[t];
            }
        }
        if (npost == 0) continue;  // nothing of this request lives in the tile
        if (tid == 0) cta_postings += npost;
        const bool sparse_mode = npost * 4u < tile_n;
        if (sparse_mode)
            for (uint32_t i = tid; i < (tile >> 5); i += kTileThreads) s_claim[i] = 0;

        // (1) dense slices, one round per rank so that a part array sees one list at a time
        for (uint32_t r = 0; r < max_dense; ++r) {
            for (uint32_t l = 0; l < L; ++l) {
                const uint32_t part = a.leaf_part[qp.leaf_begin + l];
                const PartSlices ps = a.slices[part];
                if (r >= ps.n_dense) continue;
                const uint32_t mi = ps.m_begin + r;
                const uint32_t* trow = a.toff + (size_t)a.g_row[mi] * (a.n_tiles + 1);
                const uint32_t s = trow[t], e = trow[t + 1];
                const PostingsView& pv = a.postings[a.parts[part].postings];
                const uint32_t* anchors = pv.anchors + a.g_begin[mi];
                const uint16_t* scores = pv.scores + a.g_begin[mi];
                const float term_score = a.g_score[mi];
                uint32_t* dst = arr + l * tile;
                const bool single = ps.n_match == 1;
                for (uint32_t j = s + tid; j < e; j += kTileThreads) {
                    const uint32_t idx = anchors[j] - tile_base;
                    const float w = __half2float(__ushort_as_half(scores[j])) / 100.0f;  // el.score.to_f32() / 100.0 (:426)
                    const uint32_t key = vbit::score_key(term_score * w);
                    if (single) dst[idx] = key;
                    else dst[idx] = max(dst[idx], key);
                }
            }
            __syncthreads();
        }
        // (2) sparse buckets (several terms of a part may hit the same anchor: atomicMax)
        for (uint32_t l = 0; l < L; ++l) {
            const uint32_t part = a.leaf_part[qp.leaf_begin + l];
            const PartSlices ps = a.slices[part];
            if (ps.n_match == ps.n_dense) continue;
            const uint32_t* brow = a.bucket + (size_t)ps.sparse_row * (a.n_tiles + 1);
            const uint32_t s = brow[t], e = brow[t + 1];
            uint32_t* dst = arr + l * tile;
            for (uint32_t j = s + tid; j < e; j += kTileThreads) atomicMax(&dst[a.s_anchor[ps.sparse_base + j] - tile_base], a.s_key[ps.sparse_base + j]);
        }
        __syncthreads();

        // (3) epilogue: tree, boosts, count, threshold; leaves the part arrays zeroed
        c.tau = __ldcg(a.tau + q);
        c.fast_boost = false, c.can_prune = false;
        c.col = nullptr, c.col_n = 0, c.fun = 0, c.param = 0.0f, c.max_mult = 0.0f;
        if (qp.n_boosts == 1) {
            const BoostStep& bs = a.boosts[qp.boost_begin];
            if (bs.n_skip == 0 && bs.expr_op == kExprNone) {
                c.fast_boost = true;
                c.col = bs.column, c.col_n = bs.n, c.fun = bs.fun, c.param = bs.param, c.max_mult = bs.max_mult;
                c.can_prune = bs.can_prune != 0 && !qp.emit_all && qp.k != 0 && c.tau != 0;
            }
        }
        uint32_t my_present = 0;
        if (sparse_mode) {
            // posting-driven: visit only the anchors that were touched; the claim bit makes each one count once
            for (uint32_t l = 0; l < L; ++l) {
                const uint32_t part = a.leaf_part[qp.leaf_begin + l];
                const PartSlices ps = a.slices[part];
                for (uint32_t r = 0; r < ps.n_dense; ++r) {
                    const uint32_t mi = ps.m_begin + r;
                    const uint32_t* trow = a.toff + (size_t)a.g_row[mi] * (a.n_tiles + 1);
                    const uint32_t s = trow[t], e = trow[t + 1];
                    const uint32_t* anchors = a.postings[a.parts[part].postings].anchors + a.g_begin[mi];
                    for (uint32_t j = s + tid; j < e; j += kTileThreads) {
                        const uint32_t idx = anchors[j] - tile_base;
                        const uint32_t bit = 1u << (idx & 31u);
                        if (!(atomicOr(&s_claim[idx >> 5], bit) & bit)) my_present += eval_idx(a, c, arr, idx, &s_nsurv, s_list);
                    }
                }
                if (ps.n_match != ps.n_dense) {
                    const uint32_t* brow = a.bucket + (size_t)ps.sparse_row * (a.n_tiles + 1);
                    const uint32_t s = brow[t], e = brow[t + 1];
                    for (uint32_t j = s + tid; j < e; j += kTileThreads) {
                        const uint32_t idx = a.s_anchor[ps.sparse_base + j] - tile_base;
                        const uint32_t bit = 1u << (idx & 31u);
                        if (!(atomicOr(&s_claim[idx >> 5], bit) & bit)) my_present += eval_idx(a, c, arr, idx, &s_nsurv, s_list);
                    }
                }
            }
        } else if (qp.prog_len == 0 && L <= 4) {
            // vector sweep: four anchors per step, untouched groups cost one 128-bit load per part
            const uint32_t n_groups = tile >> 2;
            for (uint32_t g = tid; g < n_groups; g += kTileThreads) {
                uint4 v[4];
                uint32_t any = 0;
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    v[l] = (uint32_t)l < L ? reinterpret_cast<const uint4*>(arr + l * tile)[g] : make_uint4(0u, 0u, 0u, 0u);
                    any |= v[l].x | v[l].y | v[l].z | v[l].w;
                }
                if (!any) continue;
                uint32_t keep[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const uint32_t k0 = comp4(v[0], cc), k1 = comp4(v[1], cc), k2 = comp4(v[2], cc), k3 = comp4(v[3], cc);
                    if (!(k0 | k1 | k2 | k3)) continue;
                    float score;
                    if (L == 1) score = vbit::key_score(k0);
                    else {
                        float nd = 0.0f, sum = 0.0f;
                        const uint32_t ks[4] = {k0, k1, k2, k3};
#pragma unroll
                        for (int l = 0; l < 4; ++l) {
                            const float x = ks[l] ? fmaxf(0.0f, vbit::key_score(ks[l])) : 0.0f;
                            if (x >= 0.00001f) nd += 1.0f;
                            sum += x;  // absent parts add +0.0: the sum over the request's own parts is unchanged
                        }
                        score = sum * nd * nd;
                    }
                    ++my_present;
                    keep[cc] = finish_anchor(a, c, tile_base + (g << 2) + cc, score, &s_nsurv, s_list);
                }
#pragma unroll
                for (int l = 1; l < 4; ++l)
                    if ((uint32_t)l < L && (v[l].x | v[l].y | v[l].z | v[l].w)) reinterpret_cast<uint4*>(arr + l * tile)[g] = make_uint4(0u, 0u, 0u, 0u);
                reinterpret_cast<uint4*>(arr)[g] = make_uint4(keep[0], keep[1], keep[2], keep[3]);
            }
        } else {
            for (uint32_t idx = tid; idx < tile; idx += kTileThreads) my_present += eval_idx(a, c, arr, idx, &s_nsurv, s_list);
        }
        for (int o = 16; o > 0; o >>= 1) my_present += __shfl_xor_sync(0xFFFFFFFFu, my_present, o);
        if ((tid & 31) == 0 && my_present) atomicAdd(&s_npresent, my_present);
        __syncthreads();
        if (tid == 0 && s_npresent) atomicAdd(a.num_hits + q, (unsigned long long)s_npresent);

        // (4) merge survivors into the request's heap (sorted, k slots) under its lock
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
            const unsigned long long new_tau = s_out[k - 1];
            __syncthreads();
            if (tid == 0) {
                __stcg(a.tau + q, new_tau);
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
    if (tid == 0 && cta_postings) atomicAdd(a.stat_postings, cta_postings);
}

size_t tile_kernel_smem(uint32_t tile_log2, uint32_t max_leaves) {
    size_t need = ((size_t)max_leaves << tile_log2) * sizeof(uint32_t);
    return need <= 200 * 1024 ? need : 0;
}

void launch_tile_eval(cudaStream_t st, const TileArgs& a, int n_sms) {
    if (a.n_items == 0) return;
    const size_t smem = tile_kernel_smem(a.tile_log2, a.max_leaves);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(tile_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        configured = true;
    }
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_eval_kernel, kTileThreads, smem);
    if (per_sm < 1) per_sm = 1;
    unsigned long long blocks = (unsigned long long)n_sms * (unsigned)per_sm;
    if (blocks > a.n_items) blocks = a.n_items;
    tile_eval_kernel<<<(unsigned)blocks, kTileThreads, smem, st>>>(a);
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
    merge_heaps_kernel<<<n_queries, 256, smem, st>>>(src_keys, src_hits, n_src, n_queries, stride, queries, out_keys, out_hits);
    count_launch();
}

}  // namespace vdev
