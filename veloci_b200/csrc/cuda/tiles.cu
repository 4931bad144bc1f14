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

__global__ void __launch_bounds__(kTileThreads) tile_eval_kernel(TileArgs a) {
    extern __shared__ __align__(16) uint32_t arr[];  // [max_leaves][tile]
    __shared__ unsigned long long s_item;
    __shared__ uint32_t s_npresent, s_nsurv;
    __shared__ unsigned long long s_list[kSurvivorCap];
    __shared__ unsigned long long s_heap[kMaxK];
    __shared__ unsigned long long s_out[kMaxK];

    const uint32_t tid = threadIdx.x;
    const uint32_t tile = 1u << a.tile_log2;
    unsigned long long cta_postings = 0;  // thread 0 only

    while (true) {
        __syncthreads();  // everyone is done with the previous item's shared state
        if (tid == 0) s_item = atomicAdd(a.work_counter, 1ull);
        __syncthreads();
        const unsigned long long item = s_item;
        if (item >= a.n_items) break;
        const uint32_t t = (uint32_t)(item / a.n_queries), q = (uint32_t)(item % a.n_queries);
        const QueryProgram qp = a.queries[q];
        if (!qp.active || qp.n_leaves == 0) continue;
        const uint32_t L = qp.n_leaves;
        const uint64_t tile_base64 = (uint64_t)a.anchor_lo + ((uint64_t)t << a.tile_log2);
        const uint32_t tile_base = (uint32_t)tile_base64;
        const uint32_t tile_n = (uint32_t)min((uint64_t)tile, (uint64_t)a.anchor_hi - tile_base64);

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

        // (1) clear the part arrays
        {
            uint4* p4 = reinterpret_cast<uint4*>(arr);
            const uint32_t n4 = (L * tile) >> 2;
            for (uint32_t i = tid; i < n4; i += kTileThreads) p4[i] = make_uint4(0u, 0u, 0u, 0u);
            if (tid == 0) s_npresent = 0, s_nsurv = 0;
        }
        __syncthreads();

        // (2) dense slices, one round per rank so that a part array sees one list at a time
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
        // (3) sparse buckets (several terms of a part may hit the same anchor: atomicMax)
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

        // (4) epilogue: tree, boosts, count, threshold
        const unsigned long long tau = __ldcg(a.tau + q);
        const uint32_t* prog = a.prog + qp.prog_begin;
        uint32_t my_present = 0;
        for (uint32_t idx = tid; idx < tile_n; idx += kTileThreads) {
            bool present;
            float score;
            if (qp.prog_len == 0) {  // one part, or a flat `or` of parts with distinct terms (leaves in slot order)
                present = false;
                float nd = 0.0f, sum = 0.0f;
                for (uint32_t l = 0; l < L; ++l) {
                    const uint32_t key = arr[l * tile + idx];
                    const float v = key ? fmaxf(0.0f, vbit::key_score(key)) : 0.0f;
                    present = present || key != 0;
                    if (v >= 0.00001f) nd += 1.0f;
                    sum += v;
                }
                score = L == 1 ? vbit::key_score(arr[idx]) : sum * nd * nd;
            } else {
                present = eval_program(prog, qp.prog_len, arr, tile, idx, score);
            }
            uint32_t keep = 0;
            if (present) {
                const uint32_t anchor = tile_base + idx;
                for (uint32_t b = 0; b < qp.n_boosts; ++b) {
                    const BoostStep& bs = a.boosts[qp.boost_begin + b];
                    bool skip = false;
                    for (uint32_t i = 0; i < bs.n_skip; ++i) skip = skip || fabsf(bs.skip[i] - score) < 0.00001f;
                    if (skip || anchor >= bs.n) continue;
                    const uint32_t bits = __ldg(bs.column + anchor);
                    if (bits != kNoValue) score = apply_boost_step(bs, score, __uint_as_float(bits));
                }
                ++my_present;
                uint32_t key = vbit::score_key(score);
                if (key == 0) key = 1;
                const unsigned long long comp = ((unsigned long long)key << 32) | anchor;
                if (qp.emit_all) {
                    const unsigned long long at = atomicAdd(a.emit_count, 1ull);
                    if (at < a.emit_capacity) a.emit[at] = comp;
                }
                if (qp.k != 0 && comp > tau) {
                    const uint32_t pos = atomicAdd(&s_nsurv, 1u);
                    if (pos < kSurvivorCap) s_list[pos] = comp;
                    else keep = key;  // deferred to the next merge round
                }
            }
            arr[idx] = keep;
        }
        for (int o = 16; o > 0; o >>= 1) my_present += __shfl_xor_sync(0xFFFFFFFFu, my_present, o);
        if ((tid & 31) == 0 && my_present) atomicAdd(&s_npresent, my_present);
        __syncthreads();
        if (tid == 0 && s_npresent) atomicAdd(a.num_hits + q, (unsigned long long)s_npresent);

        // (5) merge survivors into the request's heap (sorted, k slots) under its lock
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
            // overflow: collect the deferred survivors that still beat the new threshold
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
