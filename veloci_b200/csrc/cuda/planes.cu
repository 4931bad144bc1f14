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
static const uint32_t kUnitItems = 1024;   // items of one tile a CTA takes at a time
static const uint32_t kQueueCap = 128;     // pending candidates of a warp (sparse mode)
static const uint32_t kDenseGroup = 96;    // candidates in one 32-word group from which every lane walks its own word
static const uint32_t kSurvCap = 64;

struct WarpScratch {
    FastDesc desc;                         // 144
    uint32_t ebits[256];                   // anchors of the tile that have entries
    uint32_t ent_key[kFastMaxEntries];
    uint16_t ent_code[kFastMaxEntries];    // index in tile | leaf << 13
    uint16_t queue[kQueueCap];
    unsigned long long surv[kSurvCap];
    unsigned long long merge[kFastMaxK + kSurvCap];
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
        d.flags = kFastOk | (qp.n_boosts ? kFastBoost : 0u);
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
struct ItemState {  // per-warp registers of the item being processed (warp-uniform unless noted)
    uint32_t q, t, n_ent;
    uint32_t L, k;
    uint32_t tile_base_rel;  // first anchor of the tile, relative to anchor_lo
    unsigned long long tau;
    int lev[kFastMaxLeaves];  // per count of present parts: -2 no candidates, -1 every anchor, else boost level
    uint32_t ns;              // survivors pending in scratch
    bool lev_in_smem;
};

__device__ __forceinline__ float boost_mult(uint32_t fun, float x) {
    switch (fun) {
        case kBoostLog10: return log10f(x);
        case kBoostLog2: return log2f(x);
        default: return x;  // kBoostMultiply
    }
}

// Which anchors still have to be evaluated, per number of parts present (see the file comment).
__device__ __forceinline__ void compute_levels(const FastDesc& D, ItemState& s, uint32_t lane) {
#pragma unroll
    for (int i = 0; i < (int)kFastMaxLeaves; ++i) s.lev[i] = -1;
    if (s.tau == 0) return;
    const float tau_score = vbit::key_score((uint32_t)(s.tau >> 32));
    if (!(tau_score > 1e-30f)) return;
    if (D.flags & kFastBoost) {
        float m = 0.0f;
        if (lane < kBoostLevels) {
            m = boost_mult(D.fb_fun, __ldg(&D.fb_lev->thr[lane]) + D.fb_param);
            m = fmaxf(m, 0.0f) * 1.00001f + 1e-6f;  // anchors below the level's threshold multiply by at most this
        }
#pragma unroll
        for (int i = 0; i < (int)kFastMaxLeaves; ++i) {
            if ((uint32_t)i >= s.L) break;
            const float B = D.bound[i];
            if (B * D.fb_max_mult * 1.00001f < tau_score) {
                s.lev[i] = -2;
                continue;
            }
            const uint32_t mask = __ballot_sync(0xFFFFFFFFu, lane < kBoostLevels && B * m < tau_score);
            s.lev[i] = mask ? 31 - __clz((int)mask) : -1;
        }
    } else {
#pragma unroll
        for (int i = 0; i < (int)kFastMaxLeaves; ++i) {
            if ((uint32_t)i >= s.L) break;
            if (D.bound[i] * 1.00001f < tau_score) s.lev[i] = -2;
        }
    }
}

// Merges the warp's pending survivors into the request's heap (sorted, k slots) under its lock; returns the new threshold.
__device__ __noinline__ unsigned long long flush_survivors(unsigned long long* heap, unsigned long long* tau_slot, uint32_t* lock_slot, WarpScratch* Sp, uint32_t k, uint32_t ns, uint32_t lane) {
    WarpScratch& S = *Sp;
    if (lane == 0) {
        while (atomicCAS(lock_slot, 0u, 1u) != 0u) __nanosleep(32);
        __threadfence();
    }
    __syncwarp();
    for (uint32_t i = lane; i < k; i += 32) S.merge[i] = __ldcg(heap + i);
    for (uint32_t i = lane; i < ns; i += 32) S.merge[k + i] = S.surv[i];
    __syncwarp();
    const uint32_t n = k + ns;
    for (uint32_t e = lane; e < n; e += 32) {
        const unsigned long long key = S.merge[e];
        if (key == 0) continue;
        uint32_t rank = 0;
        for (uint32_t j = 0; j < n; ++j) rank += S.merge[j] > key;
        if (rank < k) __stcg(heap + rank, key);
        if (rank == k - 1) __stcg(tau_slot, key);
    }
    __threadfence();
    __syncwarp();
    const unsigned long long tau = __ldcg(tau_slot);
    __syncwarp();
    if (lane == 0) atomicExch(lock_slot, 0u);
    return tau;
}

// Exact score of one anchor of the tile (index `idx`); returns its order key when it beats the threshold, else 0.
__device__ __forceinline__ unsigned long long eval_candidate(const PlaneArgs& a, const WarpScratch& S, const ItemState& s, const uint32_t* __restrict__ s_bits, uint32_t W, uint32_t idx) {
    const FastDesc& D = S.desc;
    const uint32_t w = idx >> 5, bit = 1u << (idx & 31u);
    const uint32_t rel = s.tile_base_rel + idx;
    const bool in_e = s.n_ent != 0 && (S.ebits[w] & bit) != 0;
    const size_t plane_stride = (size_t)a.planes.words * 32u;
    float sum = 0.0f, nd = 0.0f, v0 = 0.0f;
#pragma unroll
    for (uint32_t l = 0; l < kFastMaxLeaves; ++l) {
        if (l >= s.L) break;
        float v = 0.0f;
        const uint32_t np = D.n_planes[l];
#pragma unroll
        for (uint32_t j = 0; j < kPartPlaneSlots; ++j) {
            if (j >= np) break;
            const uint32_t p = D.plane[l][j];
            if (s_bits[p * W + w] & bit) {
                const unsigned short h = __ldg(a.planes.score + p * plane_stride + rel);
                const float wgt = __fdiv_rn(__half2float(__ushort_as_half(h)), 100.0f);  // el.score.to_f32() / 100.0 (search_field.rs:426)
                v = fmaxf(v, D.ts[l][j] * wgt);
            }
        }
        if (in_e) {
            const uint32_t code = idx | (l << 13);
            for (uint32_t e = 0; e < s.n_ent; ++e)
                if (S.ent_code[e] == code) v = fmaxf(v, __uint_as_float(S.ent_key[e] & 0x7FFFFFFFu));
        }
        if (v >= 0.00001f) nd += 1.0f;
        sum += v;
        if (l == 0) v0 = v;
    }
    float score = s.L == 1 ? v0 : sum * nd * nd;
    const uint32_t anchor = a.anchor_lo + rel;
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
    const unsigned long long comp = ((unsigned long long)key << 32) | anchor;
    return comp > s.tau ? comp : 0ull;
}

__device__ __forceinline__ void push_survivors(const PlaneArgs& a, WarpScratch& S, ItemState& s, uint32_t lane, unsigned long long comp) {
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, comp != 0);
    if (!m) return;
    if (comp) S.surv[s.ns + __popc(m & ((1u << lane) - 1u))] = comp;
    s.ns += __popc(m);
    __syncwarp();
    if (s.ns > kSurvCap - 32) {
        s.tau = flush_survivors(a.heap + (size_t)s.q * a.heap_stride, a.tau + s.q, a.lock + s.q, &S, s.k, s.ns, lane);
        s.ns = 0;
        compute_levels(S.desc, s, lane);
    }
}

__global__ void __launch_bounds__(kPlaneThreads, 1) plane_eval_kernel(PlaneArgs a) {
    extern __shared__ __align__(16) uint32_t plane_smem[];
    __shared__ uint32_t s_unit, s_next;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t W = 1u << (a.tile_log2 - 5);
    uint32_t* s_bits = plane_smem;
    uint32_t* s_lev = s_bits + a.planes.n_planes * W;
    WarpScratch& S = reinterpret_cast<WarpScratch*>(s_lev + kBoostLevels * W)[warp];
    for (uint32_t i = lane; i < 256; i += 32) S.ebits[i] = 0;
    unsigned long long st_cand = 0, st_items = 0;  // lane 0

    while (true) {
        __syncthreads();
        if (tid == 0) s_unit = (uint32_t)atomicAdd(a.work_counter, 1ull);
        __syncthreads();
        const uint32_t unit = s_unit;
        if (unit >= a.n_units) break;
        const uint32_t t = unit / a.chunks_per_tile, c = unit % a.chunks_per_tile;
        const uint32_t tb = a.tile_item_begin[t], te = a.tile_item_begin[t + 1];
        if (te - tb <= c * kUnitItems) continue;
        const uint32_t ibeg = tb + c * kUnitItems, iend = min(te, ibeg + kUnitItems);
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
            if (tid == 0) s_next = ibeg + 2 * kPlaneWarps;
        }
        __syncthreads();

        // two-deep software pipeline over the warp's items: item record two ahead, descriptor + threshold one ahead
        uint32_t idx0 = ibeg + warp, idx1 = ibeg + kPlaneWarps + warp;
        ItemRec it0, it1;
        it0.q = 0, it1.q = 0;
        if (idx0 < iend) it0 = a.items[idx0];
        if (idx1 < iend) it1 = a.items[idx1];
        uint4 d0 = make_uint4(0u, 0u, 0u, 0u);
        unsigned long long tau0 = 0;
        if (idx0 < iend) {
            if (lane < 9) d0 = __ldg(reinterpret_cast<const uint4*>(a.fast + it0.q) + lane);
            tau0 = __ldcg(a.tau + it0.q);
        }
        while (idx0 < iend) {
            uint32_t idx2 = 0;
            if (lane == 0) idx2 = atomicAdd(&s_next, 1u);
            idx2 = __shfl_sync(0xFFFFFFFFu, idx2, 0);
            ItemRec it2;
            it2.q = 0;
            if (idx2 < iend) it2 = a.items[idx2];
            uint4 d1 = make_uint4(0u, 0u, 0u, 0u);
            unsigned long long tau1 = 0;
            if (idx1 < iend) {
                if (lane < 9) d1 = __ldg(reinterpret_cast<const uint4*>(a.fast + it1.q) + lane);
                tau1 = __ldcg(a.tau + it1.q);
            }

            // ---- item (t, it0.q)
            __syncwarp();
            if (lane < 9) reinterpret_cast<uint4*>(&S.desc)[lane] = d0;
            __syncwarp();
            const FastDesc& D = S.desc;
            ItemState s;
            s.q = it0.q, s.t = t, s.n_ent = it0.npost, s.L = D.n_leaves, s.k = D.k, s.tau = tau0, s.ns = 0;
            s.tile_base_rel = t << a.tile_log2;
            s.lev_in_smem = D.fb_lev == a.lev_dev;
            const uint32_t tile_base = a.anchor_lo + s.tile_base_rel;
            if (s.n_ent) {  // entries: postings of the request's non-plane terms inside the tile
                uint32_t ne = 0;
                for (uint32_t si = 0; si < it0.n_slices; ++si) {
                    const SliceRec sr = a.slice_recs[it0.slice_begin + si];
                    for (uint32_t j = lane; j < sr.n; j += 32) {
                        uint32_t anchor, key;
                        if (sr.kind == 0) {
                            const Posting p = a.postings[sr.postings].post[sr.begin + j];
                            anchor = p.anchor, key = vbit::score_key(sr.term_score * p.weight);  // hit.score * (el.score / 100.0) (:426)
                        } else {
                            const SparseEntry e = a.sparse[sr.begin + j];
                            anchor = e.anchor, key = e.key;
                        }
                        const uint32_t idx = anchor - tile_base;
                        S.ent_code[ne + j] = (uint16_t)(idx | ((uint32_t)sr.leaf << 13));
                        S.ent_key[ne + j] = key;
                        atomicOr(&S.ebits[idx >> 5], 1u << (idx & 31u));
                    }
                    ne += sr.n;
                }
                __syncwarp();
            }
            compute_levels(D, s, lane);
            // word offsets of the request's planes in the staged tile
            uint32_t pb[kFastMaxLeaves][kPartPlaneSlots];
#pragma unroll
            for (uint32_t l = 0; l < kFastMaxLeaves; ++l)
#pragma unroll
                for (uint32_t j = 0; j < kPartPlaneSlots; ++j) pb[l][j] = (uint32_t)D.plane[l][j] * W;
            const uint32_t np0 = D.n_planes[0], np1 = D.n_planes[1], np2 = D.n_planes[2], np3 = D.n_planes[3];
            const uint32_t* lev_glob = (D.flags & kFastBoost) ? D.fb_lev->bits + (size_t)t * W : nullptr;
            const uint32_t lev_words = (D.flags & kFastBoost) ? D.fb_lev->words : 0u;

            uint32_t cnt = 0, qn = 0, ncand = 0;
            for (uint32_t w = lane; w < W; w += 32) {
                uint32_t x[kFastMaxLeaves] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (uint32_t j = 0; j < kPartPlaneSlots; ++j) {
                    if (j < np0) x[0] |= s_bits[pb[0][j] + w];
                    if (j < np1) x[1] |= s_bits[pb[1][j] + w];
                    if (j < np2) x[2] |= s_bits[pb[2][j] + w];
                    if (j < np3) x[3] |= s_bits[pb[3][j] + w];
                }
                const uint32_t e = s.n_ent ? S.ebits[w] : 0u;
                cnt += __popc(x[0] | x[1] | x[2] | x[3] | e);
                // bit-sliced count of the parts present per anchor
                uint32_t ones = x[0], twos = 0, fours = 0, cy, cy2;
                cy = ones & x[1], ones ^= x[1], twos ^= cy;
                cy = ones & x[2], ones ^= x[2], cy2 = twos & cy, twos ^= cy, fours |= cy2;
                cy = ones & x[3], ones ^= x[3], cy2 = twos & cy, twos ^= cy, fours |= cy2;
                const uint32_t ex[kFastMaxLeaves] = {ones & ~twos & ~fours, twos & ~ones & ~fours, ones & twos, fours};
                uint32_t cm = e;
#pragma unroll
                for (int i = 0; i < (int)kFastMaxLeaves; ++i) {
                    if ((uint32_t)i >= s.L) break;
                    const int lv = s.lev[i];
                    if (lv == -2 || ex[i] == 0) continue;
                    uint32_t lw = 0xFFFFFFFFu;
                    if (lv >= 0) lw = s.lev_in_smem ? s_lev[(uint32_t)lv * W + w] : __ldg(lev_glob + (size_t)lv * lev_words + w);
                    cm |= ex[i] & lw;
                }
                if (__ballot_sync(0xFFFFFFFFu, cm != 0) == 0) continue;
                const uint32_t pc = __popc(cm);
                uint32_t incl = pc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if ((int)lane >= o) incl += y;
                }
                const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                ncand += total;
                if (total >= kDenseGroup) {  // densely hit group (threshold not converged yet): every lane walks its own word
                    while (__ballot_sync(0xFFFFFFFFu, cm != 0)) {
                        unsigned long long comp = 0;
                        if (cm) {
                            const uint32_t idx = w * 32u + (uint32_t)__ffs((int)cm) - 1u;
                            cm &= cm - 1u;
                            comp = eval_candidate(a, S, s, s_bits, W, idx);
                        }
                        push_survivors(a, S, s, lane, comp);
                    }
                } else {
                    uint32_t at = qn + incl - pc;
                    while (cm) {
                        S.queue[at++] = (uint16_t)(w * 32u + (uint32_t)__ffs((int)cm) - 1u);
                        cm &= cm - 1u;
                    }
                    qn += total;
                    __syncwarp();
                    while (qn >= 32) {
                        qn -= 32;
                        const unsigned long long comp = eval_candidate(a, S, s, s_bits, W, S.queue[qn + lane]);
                        push_survivors(a, S, s, lane, comp);
                    }
                    __syncwarp();
                }
            }
            if (qn) {
                unsigned long long comp = 0;
                if (lane < qn) comp = eval_candidate(a, S, s, s_bits, W, S.queue[lane]);
                push_survivors(a, S, s, lane, comp);
            }
            if (s.ns) flush_survivors(a.heap + (size_t)s.q * a.heap_stride, a.tau + s.q, a.lock + s.q, &S, s.k, s.ns, lane);
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
            if (lane == 0) {
                if (cnt) atomicAdd(a.num_hits + s.q, (unsigned long long)cnt);
                st_cand += ncand, st_items += 1;
            }
            if (s.n_ent) {  // leave the entry bitmap zeroed
                __syncwarp();
                for (uint32_t e = lane; e < s.n_ent; e += 32) S.ebits[(S.ent_code[e] & 0x1FFFu) >> 5] = 0;
            }

            idx0 = idx1, it0 = it1, d0 = d1, tau0 = tau1;
            idx1 = idx2, it1 = it2;
        }
    }
    if (lane == 0 && st_items) {
        atomicAdd(a.stats + 0, st_items);
        atomicAdd(a.stats + 1, st_cand);
    }
}

static size_t plane_smem_bytes(uint32_t tile_log2, uint32_t n_planes) {
    const size_t W = (size_t)1 << (tile_log2 - 5);
    return ((size_t)n_planes + kBoostLevels) * W * 4 + sizeof(WarpScratch) * kPlaneWarps;
}

size_t plane_kernel_smem(uint32_t tile_log2, uint32_t n_planes) {
    if (tile_log2 < 10 || tile_log2 > kPlaneTileLog2 || n_planes > kMaxPlanes) return 0;
    const size_t need = plane_smem_bytes(tile_log2, n_planes);
    return need + 1024 <= 227 * 1024 ? need : 0;
}

uint32_t plane_unit_items() { return kUnitItems; }

void launch_plane_eval(cudaStream_t st, const PlaneArgs& a, int n_sms) {
    if (a.n_units == 0) return;
    const size_t smem = plane_kernel_smem(a.tile_log2, a.planes.n_planes);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(plane_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        configured = true;
    }
    unsigned blocks = (unsigned)n_sms;
    if (blocks > a.n_units) blocks = a.n_units;
    plane_eval_kernel<<<blocks, kPlaneThreads, smem, st>>>(a);
    count_launch();
}

}  // namespace vdev
