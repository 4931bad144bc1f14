// K1 fuzzy_match + match grouping/scoring + posting slicing.
//
//   fuzzy_match_kernel      get_term_ids_in_field (src/search/search_field.rs:277-398): which
//                           dictionary terms the Levenshtein automaton of each search part accepts.
//   group_* / score_scatter the (term id, score) hit list per part; score =
//                           get_default_score_for_distance(distance_dfa(..)) (:27-33, :691-732) * boost.
//   dense_tile_offsets /    prepare resolve_token_to_anchor (:400-504) for the tile kernel: where the
//   sparse_*                postings of every matched term cross the anchor-tile boundaries.
#include <cuda_fp16.h>

#include <algorithm>

#include "bitvec.cuh"
#include "kernels.cuh"

namespace vdev {

// ---------------------------------------------------------------- fuzzy_match
// Grid: x = groups of 256 dictionary tiles (8 warps x 32 lanes, one tile per lane),
//       y = chunks of kPartChunk search parts.  Each lane keeps its tile's common
// prefix in registers for the whole part loop, so every dictionary byte is read
// once per part chunk.  Per (part, tile): run the bit-parallel automaton over the
// prefix; if no extension can stay within d the 32 terms are skipped (exact: the
// column minimum is a lower bound for every extension), otherwise the warp
// verifies the tile with one term per lane.
static const int kPartChunk = 64;
static const uint32_t kSparseSegment = 2048;  // postings of one sparse match a warp walks (grid.y = segments of the longest list)
static const int kFuzzyThreads = 256;
static const int kPeqCodes = 128;  // Eq masks of the first 128 alphabet codes come from a shared-memory table

template <class W>
struct PartLite {
    W peq[kPeqCodes];
    uint16_t sym[64];
    uint32_t m, d, flags, id;
};

template <class W>
__device__ __forceinline__ W eq_of(const PartLite<W>& q, uint16_t c) {
    if (c < kPeqCodes) return q.peq[c];
    W eq = 0;
    for (uint32_t j = 0; j < q.m; ++j) eq |= (W)(q.sym[j] == c) << j;
    return eq;
}

template <class W>
__device__ __forceinline__ W shfl_word(W v, int src);
template <>
__device__ __forceinline__ uint32_t shfl_word<uint32_t>(uint32_t v, int src) {
    return __shfl_sync(0xFFFFFFFFu, v, src);
}
template <>
__device__ __forceinline__ uint64_t shfl_word<uint64_t>(uint64_t v, int src) {
    return __shfl_sync(0xFFFFFFFFu, (unsigned long long)v, src);
}

// W = uint32_t when every query of the launch has at most 32 scalars (the usual case), else uint64_t.
template <class W, int CHUNK>
__global__ void __launch_bounds__(kFuzzyThreads) fuzzy_match_kernel(DictView dict, const PartQuery* __restrict__ parts, const uint32_t* __restrict__ part_ids, uint32_t n_parts,
                                                                    MatchRecord* __restrict__ out, uint32_t capacity, unsigned long long* __restrict__ counter) {
    extern __shared__ __align__(16) unsigned char fuzzy_smem[];
    PartLite<W>* sp = reinterpret_cast<PartLite<W>*>(fuzzy_smem);
    const uint32_t chunk_begin = blockIdx.y * CHUNK;
    const uint32_t chunk_n = min((uint32_t)CHUNK, n_parts - chunk_begin);
    for (uint32_t i = threadIdx.x; i < chunk_n * kPeqCodes; i += blockDim.x) sp[i / kPeqCodes].peq[i % kPeqCodes] = 0;
    for (uint32_t i = threadIdx.x; i < chunk_n * 64; i += blockDim.x) {
        uint32_t p = i / 64, j = i % 64;
        sp[p].sym[j] = parts[part_ids[chunk_begin + p]].match_sym[j];
    }
    for (uint32_t p = threadIdx.x; p < chunk_n; p += blockDim.x) {
        const PartQuery& q = parts[part_ids[chunk_begin + p]];
        sp[p].m = q.m;
        sp[p].d = q.d_match;
        sp[p].flags = q.flags;
        sp[p].id = part_ids[chunk_begin + p];
    }
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < chunk_n; p += blockDim.x)  // one thread per part: no atomics needed
        for (uint32_t j = 0; j < sp[p].m; ++j)
            if (sp[p].sym[j] < kPeqCodes) sp[p].peq[sp[p].sym[j]] |= (W)1 << j;
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp_tile0 = (blockIdx.x * (kFuzzyThreads / 32) + (threadIdx.x >> 5)) * 32u;
    if (warp_tile0 >= dict.n_tiles) return;
    const uint32_t my_tile = warp_tile0 + lane;
    const bool have_tile = my_tile < dict.n_tiles;
    TilePrefix pfx0, pfx1;
    pfx0.len = 0, pfx1.len = 0;
    if (have_tile) {
        pfx0 = dict.tiles[0][my_tile];
        pfx1 = dict.tiles[1][my_tile];
    }

    for (uint32_t p = 0; p < chunk_n; ++p) {
        const PartLite<W>& q = sp[p];
        const uint32_t m = q.m, d = q.d;
        const bool prefix_mode = q.flags & kPartPrefix, transposition = q.flags & kPartTransposition;
        const int variant = (q.flags & kPartRawCase) ? 1 : 0;
        bool viable = false;
        vbit::StateT<W> s;
        vbit::init(s, m);
        uint32_t plen = 0;
        if (have_tile) {
            plen = variant ? pfx1.len : pfx0.len;
#pragma unroll
            for (uint32_t i = 0; i < kTilePrefixMax; ++i)
                if (i < plen) vbit::step(s, eq_of(q, variant ? pfx1.sym[i] : pfx0.sym[i]), m, transposition);
            viable = (prefix_mode && s.best <= d) || vbit::column_min(s, m) <= d;
        }
        uint32_t todo = __ballot_sync(0xFFFFFFFFu, viable);
        while (todo) {
            const int b = __ffs(todo) - 1;
            todo &= todo - 1;
            // every term of tile b starts with the tile prefix: continue from lane b's state
            vbit::StateT<W> t;
            t.vp = shfl_word<W>(s.vp, b), t.vn = shfl_word<W>(s.vn, b), t.d0 = shfl_word<W>(s.d0, b), t.eq = shfl_word<W>(s.eq, b);
            t.score = __shfl_sync(0xFFFFFFFFu, s.score, b), t.cols = __shfl_sync(0xFFFFFFFFu, s.cols, b), t.best = __shfl_sync(0xFFFFFFFFu, s.best, b);
            const uint32_t skip = __shfl_sync(0xFFFFFFFFu, plen, b);
            const uint32_t slot = (warp_tile0 + b) * kDictTile + lane;
            bool match = false;
            if (slot < dict.n) {
                const uint32_t o = dict.off[variant][slot];
                const uint16_t* ts = dict.sym[variant] + o;
                const uint32_t n = dict.off[variant][slot + 1] - o;
                for (uint32_t i = skip; i < n; ++i) vbit::step(t, eq_of(q, ts[i]), m, transposition);
                match = prefix_mode ? (t.best <= d) : (t.score <= d);
            }
            const uint32_t mm = __ballot_sync(0xFFFFFFFFu, match);
            if (mm) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(mm));
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (match) {
                    unsigned long long at = base + __popc(mm & ((1u << lane) - 1u));
                    if (at < capacity) out[at] = MatchRecord{q.id, slot};
                }
            }
        }
    }
}

void launch_fuzzy_match(cudaStream_t st, const DictView& dict, const PartQuery* parts, const uint32_t* part_ids, uint32_t n_parts, uint32_t max_m, MatchRecord* out,
                        uint32_t capacity, unsigned long long* counter) {
    if (n_parts == 0 || dict.n == 0) return;
    const unsigned gx = (dict.n_tiles + kFuzzyThreads - 1) / kFuzzyThreads;
    if (max_m <= 32) {
        const size_t smem = sizeof(PartLite<uint32_t>) * kPartChunk;
        static PerDeviceOnce once;
        if (once.first()) cudaFuncSetAttribute(fuzzy_match_kernel<uint32_t, kPartChunk>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dim3 grid(gx, (n_parts + kPartChunk - 1) / kPartChunk);
        fuzzy_match_kernel<uint32_t, kPartChunk><<<grid, kFuzzyThreads, smem, st>>>(dict, parts, part_ids, n_parts, out, capacity, counter);
    } else {
        const size_t smem = sizeof(PartLite<uint64_t>) * (kPartChunk / 2);
        static PerDeviceOnce once;
        if (once.first()) cudaFuncSetAttribute(fuzzy_match_kernel<uint64_t, kPartChunk / 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dim3 grid(gx, (n_parts + kPartChunk / 2 - 1) / (kPartChunk / 2));
        fuzzy_match_kernel<uint64_t, kPartChunk / 2><<<grid, kFuzzyThreads, smem, st>>>(dict, parts, part_ids, n_parts, out, capacity, counter);
    }
    count_launch();
}

// ---------------------------------------------------------------- regex parts
// get_text_lines_from_fst with `is_regex` (search_field.rs:72-83): the reference walks the FST with the pattern's DFA; a
// term matches when the DFA is in a match state after the term's last symbol, or -- starts_with -- after any prefix.
// Here every thread runs the DFA over one term's raw-case symbols (case-insensitivity is in the DFA's classes): two
// dependent loads per symbol (class of the symbol, transition), tables small enough to stay in L1/L2.
__global__ void __launch_bounds__(256) regex_match_kernel(DictView dict, const RegexPartDev* __restrict__ parts, MatchRecord* __restrict__ out, uint32_t capacity,
                                                          unsigned long long* __restrict__ counter) {
    const RegexPartDev p = parts[blockIdx.y];
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    bool match = false;
    if (slot < dict.n) {
        const uint32_t o = dict.off[1][slot];
        const uint16_t* ts = dict.sym[1] + o;
        const uint32_t n = dict.off[1][slot + 1] - o;
        uint32_t st = p.start;
        bool hit = p.sticky && (st & 0x8000u);
        uint32_t i = 0;
        for (; i < n && !hit; ++i) {
            st = __ldg(p.trans + (size_t)(st & 0x7FFFu) * p.n_classes + __ldg(p.class_of_code + ts[i]));
            if ((st & 0x7FFFu) == 0) break;  // dead: no extension matches
            hit = p.sticky && (st & 0x8000u);
        }
        match = hit || (!p.sticky && i == n && (st & 0x8000u));
    }
    const uint32_t mm = __ballot_sync(0xFFFFFFFFu, match);
    if (mm) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(mm));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (match) {
            const unsigned long long at = base + __popc(mm & ((1u << lane) - 1u));
            if (at < capacity) out[at] = MatchRecord{p.part, slot};
        }
    }
}

void launch_regex_match(cudaStream_t st, const DictView& dict, const RegexPartDev* parts, uint32_t n_parts, MatchRecord* out, uint32_t capacity, unsigned long long* counter) {
    if (n_parts == 0 || dict.n == 0) return;
    dim3 grid((dict.n + 255u) / 256u, n_parts);
    regex_match_kernel<<<grid, 256, 0, st>>>(dict, parts, out, capacity, counter);
    count_launch();
}

// ---------------------------------------------------------------- deletion-neighbourhood index
__device__ __forceinline__ uint64_t variant_hash(const uint16_t* __restrict__ sym, uint32_t n, uint32_t skip_a, uint32_t skip_b) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (uint32_t i = 0; i < n; ++i) {
        if (i == skip_a || i == skip_b) continue;
        h = (h ^ sym[i]) * 0xFF51AFD7ED558CCDull;
        h ^= h >> 29;
    }
    h *= 0xC4CEB9FE1A85EC53ull;
    return h ^ (h >> 32);
}

// Variant v of a string of n symbols: 0 = the string, 1..n = one deletion, then the pairs (i < j).
__device__ __forceinline__ void variant_skips(uint32_t v, uint32_t n, uint32_t& a, uint32_t& b) {
    a = b = 0xFFFFFFFFu;
    if (v == 0) return;
    if (v <= n) {
        a = v - 1;
        return;
    }
    uint32_t p = v - 1 - n, i = 0;
    while (p >= n - 1 - i) p -= n - 1 - i, ++i;
    a = i, b = i + 1 + p;
}
__device__ __forceinline__ uint32_t variant_count(uint32_t n, uint32_t max_del) {
    uint32_t c = 1;
    if (max_del >= 1) c += n;
    if (max_del >= 2 && n >= 2) c += n * (n - 1) / 2;
    return c;
}

template <bool FILL>
__global__ void del_index_kernel(DictView dict, uint32_t max_del, uint32_t mask, uint32_t* __restrict__ count_or_cursor, const uint32_t* __restrict__ off, DelEntry* __restrict__ ent) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= dict.n) return;
    const uint32_t o = dict.off[0][slot], n = dict.off[0][slot + 1] - o;
    const uint16_t* sym = dict.sym[0] + o;
    const uint32_t nv = variant_count(n, max_del);
    for (uint32_t v = 0; v < nv; ++v) {
        uint32_t a, b;
        variant_skips(v, n, a, b);
        const uint64_t h = variant_hash(sym, n, a, b);
        const uint32_t bucket = (uint32_t)h & mask;
        const uint32_t at = atomicAdd(&count_or_cursor[bucket], 1u);
        if (FILL) ent[off[bucket] + at] = DelEntry{slot, (uint32_t)(h >> 32)};
    }
}

void launch_del_index_pass(cudaStream_t st, const DictView& dict, uint32_t max_del, uint32_t mask, uint32_t* count_or_cursor, const uint32_t* off, DelEntry* ent) {
    if (!dict.n) return;
    const unsigned blocks = (dict.n + 127) / 128;
    if (ent) del_index_kernel<true><<<blocks, 128, 0, st>>>(dict, max_del, mask, count_or_cursor, off, ent);
    else del_index_kernel<false><<<blocks, 128, 0, st>>>(dict, max_del, mask, count_or_cursor, off, ent);
    count_launch();
}

static const int kProbeWarps = 4;
static const uint32_t kProbeSet = 2048;      // candidate set slots per warp
static const uint32_t kProbeSetMax = 1024;   // candidates above which the part is handed to the scan
static const uint32_t kSetEmpty = 0xFFFFFFFFu;

__global__ void __launch_bounds__(kProbeWarps * 32) fuzzy_probe_kernel(DictView dict, const PartQuery* __restrict__ parts, const uint32_t* __restrict__ part_ids, uint32_t n_parts,
                                                                      MatchRecord* __restrict__ out, uint32_t capacity, unsigned long long* __restrict__ counter,
                                                                      uint32_t* __restrict__ overflow_parts, unsigned long long* __restrict__ overflow_count) {
    __shared__ uint32_t s_set[kProbeWarps][kProbeSet];
    __shared__ uint64_t s_peq[kProbeWarps][kPeqCodes];
    __shared__ uint16_t s_sym[kProbeWarps][64];
    __shared__ uint32_t s_count[kProbeWarps];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t pi = blockIdx.x * kProbeWarps + warp;
    if (pi >= n_parts) return;
    const uint32_t pid = part_ids[pi];
    const PartQuery& q = parts[pid];
    const uint32_t m = q.m, d = q.d_match;
    const bool transposition = q.flags & kPartTransposition;
    uint32_t* set = s_set[warp];
    uint64_t* peq = s_peq[warp];
    uint16_t* qs = s_sym[warp];
    for (uint32_t i = lane; i < kProbeSet; i += 32) set[i] = kSetEmpty;
    for (uint32_t i = lane; i < kPeqCodes; i += 32) peq[i] = 0;
    for (uint32_t i = lane; i < 64; i += 32) qs[i] = q.match_sym[i];
    if (lane == 0) s_count[warp] = 0;
    __syncwarp();
    if (lane == 0)
        for (uint32_t j = 0; j < m; ++j)
            if (qs[j] < kPeqCodes) peq[qs[j]] |= 1ull << j;
    __syncwarp();
    const DelIndexView ix = dict.del[d >= 2 ? 1 : 0];
    const uint32_t nv = variant_count(m, d);
    for (uint32_t v = lane; v < nv; v += 32) {
        if (s_count[warp] > kProbeSetMax) break;
        uint32_t a, b;
        variant_skips(v, m, a, b);
        const uint64_t h = variant_hash(qs, m, a, b);
        const uint32_t bucket = (uint32_t)h & ix.mask, tag = (uint32_t)(h >> 32);
        const uint32_t beg = ix.off[bucket], end = ix.off[bucket + 1];
        for (uint32_t e = beg; e < end; ++e) {
            const DelEntry en = ix.ent[e];
            if (en.tag != tag) continue;
            // new candidate?
            uint32_t hs = (en.slot * 0x9E3779B1u) >> 21;
            bool fresh = false;
            while (true) {
                const uint32_t old = atomicCAS(&set[hs], kSetEmpty, en.slot);
                if (old == kSetEmpty) {
                    fresh = true;
                    break;
                }
                if ((old & 0x7FFFFFFFu) == en.slot) break;
                hs = (hs + 1u) & (kProbeSet - 1u);
            }
            if (!fresh) continue;
            if (atomicAdd(&s_count[warp], 1u) >= kProbeSetMax) break;
            // verify with the automaton (same acceptance as fuzzy_match_kernel)
            const uint32_t o = dict.off[0][en.slot], n = dict.off[0][en.slot + 1] - o;
            const uint16_t* ts = dict.sym[0] + o;
            vbit::State st;
            vbit::init(st, m);
            for (uint32_t i = 0; i < n; ++i) {
                const uint16_t c = ts[i];
                uint64_t eq;
                if (c < kPeqCodes) eq = peq[c];
                else {
                    eq = 0;
                    for (uint32_t j = 0; j < m; ++j) eq |= (uint64_t)(qs[j] == c) << j;
                }
                vbit::step(st, eq, m, transposition);
            }
            if (st.score <= d) atomicOr(&set[hs], 0x80000000u);
        }
    }
    __syncwarp();
    if (s_count[warp] > kProbeSetMax) {  // too many candidates for the table: the scan kernel takes this part
        if (lane == 0) overflow_parts[atomicAdd(overflow_count, 1ull)] = pid;
        return;
    }
    uint32_t mine = 0;
    for (uint32_t i = lane; i < kProbeSet; i += 32) {
        const uint32_t v = set[i];
        mine += (v != kSetEmpty && (v & 0x80000000u)) ? 1u : 0u;
    }
    uint32_t incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((int)lane >= o) incl += y;
    }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (!total) return;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(counter, (unsigned long long)total);
    base = __shfl_sync(0xFFFFFFFFu, base, 0) + incl - mine;
    for (uint32_t i = lane; i < kProbeSet; i += 32) {
        const uint32_t v = set[i];
        if (v != kSetEmpty && (v & 0x80000000u)) {
            if (base < capacity) out[base] = MatchRecord{pid, v & 0x7FFFFFFFu};
            ++base;
        }
    }
}

void launch_fuzzy_probe(cudaStream_t st, const DictView& dict, const PartQuery* parts, const uint32_t* part_ids, uint32_t n_parts, MatchRecord* out, uint32_t capacity,
                        unsigned long long* counter, uint32_t* overflow_parts, unsigned long long* overflow_count) {
    if (n_parts == 0 || dict.n == 0) return;
    fuzzy_probe_kernel<<<(n_parts + kProbeWarps - 1) / kProbeWarps, kProbeWarps * 32, 0, st>>>(dict, parts, part_ids, n_parts, out, capacity, counter, overflow_parts, overflow_count);
    count_launch();
}

// ---------------------------------------------------------------- grouping
__global__ void group_count_kernel(const MatchRecord* __restrict__ rec, uint32_t n, uint32_t* __restrict__ part_count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&part_count[rec[i].part], 1u);
}

// Exclusive scan of `in[0..n)` into `out[0..n]` (out[n] = total) by one block.
__global__ void __launch_bounds__(1024) scan_u32_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < n ? in[i] : 0u, x = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_sums[threadIdx.x], z = w;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xFFFFFFFFu, z, o);
                if (threadIdx.x >= o) z += y;
            }
            warp_sums[threadIdx.x] = z - w;  // exclusive
        }
        __syncthreads();
        uint32_t excl = carry + warp_sums[threadIdx.x >> 5] + x - v;
        if (i < n) out[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

__global__ void __launch_bounds__(1024) scan_u64_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, uint32_t n) {
    __shared__ uint64_t warp_sums[32];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint64_t v = i < n ? in[i] : 0ull, x = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint64_t w = warp_sums[threadIdx.x], z = w;
            for (int o = 1; o < 32; o <<= 1) {
                uint64_t y = __shfl_up_sync(0xFFFFFFFFu, z, o);
                if (threadIdx.x >= o) z += y;
            }
            warp_sums[threadIdx.x] = z - w;
        }
        __syncthreads();
        uint64_t excl = carry + warp_sums[threadIdx.x >> 5] + x - v;
        if (i < n) out[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

// One thread per match record: score it (search_field.rs:304-354), look up its
// posting list, classify it dense/sparse and place it in its part's segment
// (dense matches from the front, sparse ones from the back).
__global__ void score_scatter_kernel(ScoreScatterArgs a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_records) return;
    const MatchRecord r = a.records[i];
    const PartQuery& q = a.parts[r.part];
    float score;
    uint32_t term_id;
    if (a.inj_terms && (a.inj_all || (q.flags & kPartInjected))) {
        term_id = a.inj_terms[r.slot];
        score = a.inj_scores[r.slot];
    } else {
        const DictView& dict = a.dicts[a.part_dict[r.part]];
        const uint16_t* ts = dict.sym[0] + dict.off[0][r.slot];
        uint32_t n = dict.off[0][r.slot + 1] - dict.off[0][r.slot];
        if (dict.n_exc) {  // a term whose exact lower-case text is kept aside (U+0130 / final sigma)
            uint32_t lo = 0, hi = dict.n_exc;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (dict.exc_slot[mid] < r.slot) lo = mid + 1;
                else hi = mid;
            }
            if (lo < dict.n_exc && dict.exc_slot[lo] == r.slot) ts = dict.exc_sym + dict.exc_off[lo], n = dict.exc_off[lo + 1] - dict.exc_off[lo];
        }
        const uint32_t m = q.m_score;
        bool prefix_matches = false;
        if ((q.flags & kPartCheckPrefix) && n >= m) {
            prefix_matches = true;
            for (uint32_t j = 0; j < m; ++j) prefix_matches = prefix_matches && (ts[j] == q.score_sym[j]);
        }
        uint32_t dist = vbit::distance(q.score_sym, m, ts, n, true);  // scoring DFA: transposition on (:298-300)
        if (dist > (q.d_score & 0xFFu)) {                             // Distance::AtLeast -> distance() (:705-732)
            if (q.lower_bytes >= 255u || dict.lower_bytes[r.slot] >= 255u) dist = 255u;
            else dist = vbit::distance(q.score_sym, m, ts, n, false) & 0xFFu;
        }
        score = vbit::default_score(dist, prefix_matches);
        if (q.flags & kPartHasBoost) score = score * q.boost;
        term_id = dict.ids[r.slot];
    }

    uint64_t begin = 0;
    uint32_t df = 0, plane = kNoValue;
    if (q.postings != kNoValue) {
        const PostingsView& pv = a.postings[q.postings];
        if (term_id < pv.n_terms) {
            begin = pv.off[term_id];
            df = (uint32_t)(pv.off[term_id + 1] - begin);
            if (pv.term_plane != nullptr && a.part_planes != nullptr) plane = pv.term_plane[term_id];
        }
    }
    if (plane != kNoValue) {  // head term: register it with its part for the plane path
        PartPlanes& pp = a.part_planes[r.part];
        const uint32_t slot = atomicAdd(&pp.n, 1u);
        if (slot < kPartPlaneSlots) pp.plane[slot] = plane, pp.ts[slot] = score;
    }
    const bool dense = df >= a.dense_min || plane != kNoValue;
    const uint32_t seg = a.part_begin[r.part], cnt = a.part_begin[r.part + 1] - seg;
    uint32_t pos, row = kNoValue;
    if (dense) {
        pos = seg + atomicAdd(&a.dense_cursor[r.part], 1u);
        if (plane != kNoValue) {
            row = kPlaneRow;  // its tile offsets are the plane's tprefix row (index build): nothing to compute per batch
        } else {
            row = atomicAdd(a.n_dense_rows, 1u);
            if (row < a.dense_row_capacity) a.row_match[row] = pos;
        }
    } else {
        pos = seg + cnt - 1u - atomicAdd(&a.sparse_cursor[r.part], 1u);
    }
    a.g_term[pos] = term_id;
    a.g_score[pos] = score;
    a.g_begin[pos] = begin;
    a.g_df[pos] = df;
    a.g_row[pos] = row;
    a.g_part[pos] = r.part;
    if (a.g_plane) a.g_plane[pos] = plane;
    if (df) atomicAdd(&a.part_est[r.part], (unsigned long long)df);
}

// One block per dense match: offset of the first posting >= tile start, for every tile boundary.
__global__ void __launch_bounds__(256) dense_tile_offsets_kernel(DenseOffsetsArgs a) {
    const uint32_t row = blockIdx.x;
    const uint32_t mi = a.row_match[row];
    const PostingsView& pv = a.postings[a.parts[a.g_part[mi]].postings];
    const Posting* post = pv.post + a.g_begin[mi];
    const uint32_t df = a.g_df[mi];
    uint32_t* out = a.toff + (size_t)row * (a.n_tiles + 1);
    for (uint32_t t = threadIdx.x; t <= a.n_tiles; t += blockDim.x) {
        const uint64_t bound64 = (uint64_t)a.anchor_lo + ((uint64_t)t << a.tile_log2);
        uint32_t lo = 0, hi = df;
        if (bound64 > 0xFFFFFFFFull) lo = df;
        else {
            const uint32_t bound = (uint32_t)bound64;
            while (lo < hi) {
                uint32_t mid = (lo + hi) >> 1;
                if (post[mid].anchor < bound) lo = mid + 1;
                else hi = mid;
            }
        }
        out[t] = lo;
    }
}

// One warp per grouped match; sparse ones count their postings per tile.
__global__ void sparse_count_kernel(SparseArgs a) {
    const uint32_t mi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (mi >= a.n_matches || a.g_row[mi] != kNoValue) return;
    const uint32_t df = a.g_df[mi];
    if (df <= blockIdx.y * kSparseSegment) return;
    const uint32_t part = a.g_part[mi];
    const PostingsView& pv = a.postings[a.parts[part].postings];
    const Posting* post = pv.post + a.g_begin[mi];
    uint32_t* row = a.bucket + (size_t)part * (a.n_tiles + 1);
    const uint32_t j0 = blockIdx.y * kSparseSegment, j1 = min(df, j0 + kSparseSegment);
    for (uint32_t j = j0 + lane; j < j1; j += 32) {
        uint32_t t = (post[j].anchor - a.anchor_lo) >> a.tile_log2;
        atomicAdd(&row[t + 1], 1u);
    }
}

// One block per part: row[t+1] = first entry of tile t (exclusive scan of the counts), total -> sparse_total.
__global__ void __launch_bounds__(256) sparse_scan_kernel(uint32_t* __restrict__ bucket, uint32_t n_tiles, uint64_t* __restrict__ sparse_total) {
    __shared__ uint32_t warp_sums[8];
    __shared__ uint32_t carry;
    uint32_t* row = bucket + (size_t)blockIdx.x * (n_tiles + 1);
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += 256) {
        uint32_t t = base + threadIdx.x;
        uint32_t v = t < n_tiles ? row[t + 1] : 0u, x = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = threadIdx.x < 8 ? warp_sums[threadIdx.x] : 0u, z = w;
            for (int o = 1; o < 8; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xFFFFFFFFu, z, o);
                if (threadIdx.x >= o) z += y;
            }
            if (threadIdx.x < 8) warp_sums[threadIdx.x] = z - w;
        }
        __syncthreads();
        uint32_t excl = carry + warp_sums[threadIdx.x >> 5] + x - v;
        if (t < n_tiles) row[t + 1] = excl;
        __syncthreads();
        if (threadIdx.x == 255) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        row[0] = 0;
        sparse_total[blockIdx.x] = carry;
    }
}

// Same walk as sparse_count: writes (anchor, score key) into the tile buckets.  row[t+1]
// is the cursor of tile t; afterwards it is the end of tile t, so bucket t = [row[t], row[t+1]).
__global__ void sparse_fill_kernel(SparseArgs a) {
    const uint32_t mi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (mi >= a.n_matches || a.g_row[mi] != kNoValue) return;
    const uint32_t df = a.g_df[mi];
    if (df <= blockIdx.y * kSparseSegment) return;
    const uint32_t part = a.g_part[mi];
    const PostingsView& pv = a.postings[a.parts[part].postings];
    const Posting* post = pv.post + a.g_begin[mi];
    uint32_t* row = a.bucket + (size_t)part * (a.n_tiles + 1);
    const uint64_t base = a.sparse_base[part];
    const float term_score = a.g_score[mi];
    const uint32_t j0 = blockIdx.y * kSparseSegment, j1 = min(df, j0 + kSparseSegment);
    for (uint32_t j = j0 + lane; j < j1; j += 32) {
        const Posting p = post[j];
        const uint32_t t = (p.anchor - a.anchor_lo) >> a.tile_log2;
        const uint64_t at = base + atomicAdd(&row[t + 1], 1u);
        a.sparse[at] = SparseEntry{p.anchor, vbit::score_key(term_score * p.weight)};  // hit.score * (el.score / 100.0) (:426)
    }
}

__global__ void part_slices_kernel(PartSlices* __restrict__ out, const uint32_t* __restrict__ part_begin, const uint32_t* __restrict__ dense_cursor,
                                   const uint64_t* __restrict__ sparse_base, const PartQuery* __restrict__ parts, uint32_t n_parts) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_parts) return;
    PartSlices s;
    s.m_begin = part_begin[p];
    s.n_match = part_begin[p + 1] - part_begin[p];
    s.n_dense = dense_cursor[p];
    if (parts[p].flags & kPartList) s.n_match = 1, s.n_dense = 0;  // a list part: everything it has is in its tile bucket
    s.sparse_row = p;
    s.sparse_base = sparse_base[p];
    out[p] = s;
}

// ---------------------------------------------------------------- list producers
// Adds one entry to the tile bucket of a list part (count pass: only counts it).
__device__ __forceinline__ void list_emit(const ListArgs& a, uint32_t list_part, uint32_t anchor, uint32_t key) {
    if (anchor < a.anchor_lo || anchor >= a.anchor_hi) return;
    uint32_t* row = a.bucket + (size_t)list_part * (a.n_tiles + 1);
    const uint32_t t = (anchor - a.anchor_lo) >> a.tile_log2;
    const uint32_t at = atomicAdd(&row[t + 1], 1u);
    if (a.sparse) a.sparse[a.sparse_base[list_part] + at] = SparseEntry{anchor, key};
}

// One block per phrase_boosts entry: for every (t1, t2) of the two parts' matched terms, the anchors of the pair.  The
// threads look the pairs up, one each; the anchor lists of the pairs that exist are then emitted by the whole block (a
// pair of two head terms has tens of thousands of anchors: emitted by the thread that found it, it alone set the time
// of the launch -- 48.7 ms of a 74 ms step on the config-3 shape, profiles/r02ac_config3_summary.md).
__global__ void __launch_bounds__(128) phrase_pairs_kernel(const PhraseMember* __restrict__ members, ListArgs a) {
    __shared__ uint32_t s_begin[128], s_end[128];
    __shared__ uint32_t s_found;
    const PhraseMember m = members[blockIdx.x];
    const uint32_t b1 = a.part_begin[m.part1], n1 = a.part_begin[m.part1 + 1] - b1;
    const uint32_t b2 = a.part_begin[m.part2], n2 = a.part_begin[m.part2 + 1] - b2;
    const unsigned long long pairs = (unsigned long long)n1 * n2;
    for (unsigned long long base = 0; base < pairs; base += blockDim.x) {
        if (threadIdx.x == 0) s_found = 0;
        __syncthreads();
        const unsigned long long x = base + threadIdx.x;
        if (x < pairs) {
            const uint64_t key = ((uint64_t)a.g_term[b1 + (uint32_t)(x / n2)] << 32) | a.g_term[b2 + (uint32_t)(x % n2)];
            uint32_t lo = 0, hi = m.store.n;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (m.store.keys[mid] < key) lo = mid + 1;
                else hi = mid;
            }
            if (lo < m.store.n && m.store.keys[lo] == key && m.store.off[lo + 1] > m.store.off[lo]) {
                const uint32_t slot = atomicAdd(&s_found, 1u);
                s_begin[slot] = m.store.off[lo], s_end[slot] = m.store.off[lo + 1];
            }
        }
        __syncthreads();
        const uint32_t found = s_found;
        for (uint32_t r = 0; r < found; ++r)
            for (uint32_t i = s_begin[r] + threadIdx.x; i < s_end[r]; i += blockDim.x) list_emit(a, m.list_part, m.store.anchors[i], 0x80000000u);
        __syncthreads();
    }
}

// The phrase-pair step on its own (PlanStepPhrasePairToAnchorId, plan_steps.rs:279-293): every pair of the two term id lists.
__global__ void __launch_bounds__(128) phrase_lookup_kernel(PhraseView store, const uint32_t* __restrict__ ids1, uint32_t n1, const uint32_t* __restrict__ ids2, uint32_t n2,
                                                            uint32_t* __restrict__ pair_count, const uint32_t* __restrict__ pair_off, uint32_t* __restrict__ out) {
    const unsigned long long pairs = (unsigned long long)n1 * n2;
    for (unsigned long long x = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; x < pairs; x += (unsigned long long)gridDim.x * blockDim.x) {
        const uint64_t key = ((uint64_t)ids1[(uint32_t)(x / n2)] << 32) | ids2[(uint32_t)(x % n2)];
        uint32_t lo = 0, hi = store.n;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (store.keys[mid] < key) lo = mid + 1;
            else hi = mid;
        }
        const bool found = lo < store.n && store.keys[lo] == key;
        const uint32_t begin = found ? store.off[lo] : 0u, n = found ? store.off[lo + 1] - begin : 0u;
        if (out == nullptr) {
            pair_count[x] = n;
        } else {
            for (uint32_t i = 0; i < n; ++i) out[pair_off[x] + i] = store.anchors[begin + i];
        }
    }
}

void launch_phrase_lookup(cudaStream_t st, const PhraseView& store, const uint32_t* ids1, uint32_t n1, const uint32_t* ids2, uint32_t n2, uint32_t* pair_count, const uint32_t* pair_off,
                          uint32_t* out) {
    const unsigned long long pairs = (unsigned long long)n1 * n2;
    if (pairs == 0) return;
    const unsigned blocks = (unsigned)std::min<unsigned long long>((pairs + 127) / 128, 65535ull);
    phrase_lookup_kernel<<<blocks, 128, 0, st>>>(store, ids1, n1, ids2, n2, pair_count, pair_off, out);
    count_launch();
}

// BoostToAnchor on its own (plan_steps.rs:174-196): the joins run one store at a time over explicit id lists.
__global__ void __launch_bounds__(128) csr_expand_kernel(CsrView store, const uint32_t* __restrict__ ids, uint32_t n, uint32_t self_if_empty, uint32_t* __restrict__ count,
                                                         const uint32_t* __restrict__ off, uint32_t* __restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t id = ids[i];
        uint32_t b = 0, e = 0;
        if (id < store.n_ids) b = store.off[id], e = store.off[id + 1];
        const bool self = e == b && self_if_empty;  // a token without an entry is its own text id (search_field.rs:676-678)
        if (out == nullptr) {
            count[i] = self ? 1u : e - b;
        } else if (self) {
            out[off[i]] = id;
        } else {
            for (uint32_t k = b; k < e; ++k) out[off[i] + (k - b)] = store.val[k];
        }
    }
}

void launch_csr_expand(cudaStream_t st, const CsrView& store, const uint32_t* ids, uint32_t n, uint32_t self_if_empty, uint32_t* count, const uint32_t* off, uint32_t* out) {
    if (!n) return;
    csr_expand_kernel<<<std::min<unsigned>((n + 127) / 128, 65535u), 128, 0, st>>>(store, ids, n, self_if_empty, count, off, out);
    count_launch();
}

__global__ void __launch_bounds__(128) boost_values_kernel(const uint32_t* __restrict__ column, uint32_t column_n, CsrView v2a, const uint32_t* __restrict__ value_ids, uint32_t n,
                                                           uint32_t* __restrict__ out_anchor, uint32_t* __restrict__ out_bits) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t v = value_ids[i];
        uint32_t anchor = kNoValue, bits = kNoValue;
        if (v < column_n) bits = column[v];
        if (bits != kNoValue && v < v2a.n_ids && v2a.off[v + 1] > v2a.off[v]) anchor = v2a.val[v2a.off[v]];
        out_anchor[i] = anchor, out_bits[i] = bits;
    }
}

void launch_boost_values(cudaStream_t st, const uint32_t* column, uint32_t column_n, const CsrView& value_id_to_anchor, const uint32_t* value_ids, uint32_t n, uint32_t* out_anchor,
                         uint32_t* out_bits) {
    if (!n) return;
    boost_values_kernel<<<std::min<unsigned>((n + 127) / 128, 65535u), 128, 0, st>>>(column, column_n, value_id_to_anchor, value_ids, n, out_anchor, out_bits);
    count_launch();
}

// One block per member: the matched term ids of the part, as text ids, to their anchors.
__global__ void __launch_bounds__(128) ids_to_anchor_kernel(const IdsMember* __restrict__ members, ListArgs a) {
    const IdsMember m = members[blockIdx.x];
    const uint32_t b = a.part_begin[m.part], n = a.part_begin[m.part + 1] - b;
    for (uint32_t i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
        const uint32_t id = a.g_term[b + i];
        if (m.identity) {
            list_emit(a, m.list_part, id, 0x80000000u);
            continue;
        }
        if (id >= m.text_id_to_anchor.n_ids) continue;
        for (uint32_t j = m.text_id_to_anchor.off[id]; j < m.text_id_to_anchor.off[id + 1]; ++j) list_emit(a, m.list_part, m.text_id_to_anchor.val[j], 0x80000000u);
    }
}

void launch_ids_to_anchor(cudaStream_t st, const IdsMember* members, uint32_t n_members, const ListArgs& a) {
    if (!n_members) return;
    ids_to_anchor_kernel<<<dim3(n_members, 4), 128, 0, st>>>(members, a);
    count_launch();
}

// Text locality.  The lists are the tokens_to_text_id rows of the matched tokens (sorted, unique).  A text id that is in
// two lists is in a list other than the longest one, so the candidates are the elements of all lists but the longest;
// each candidate is emitted from the first list that has it, with c = the number of lists that have it.
__global__ void __launch_bounds__(256) text_locality_kernel(const TlInstance* __restrict__ insts, const uint32_t* __restrict__ term_parts, uint32_t* __restrict__ req_error, ListArgs a) {
    __shared__ uint32_t s_begin[kTlMaxLists], s_len[kTlMaxLists], s_prefix[kTlMaxLists + 1];
    __shared__ uint32_t s_n, s_terms, s_longest;
    const TlInstance in = insts[blockIdx.x];
    if (threadIdx.x == 0) {
        uint32_t n = 0, terms = 0, longest = 0;
        bool overflow = false;
        for (uint32_t t = 0; t < in.n_terms && !overflow; ++t) {
            const uint32_t part = term_parts[in.term_begin + t];
            const uint32_t b = a.part_begin[part], cnt = a.part_begin[part + 1] - b;
            if (cnt) ++terms;  // a term without matches is not in term_id_hits_in_field (search_field.rs:379-383)
            for (uint32_t k = 0; k < cnt; ++k) {
                const uint32_t id = a.g_term[b + k];
                if (id >= in.tokens_to_text_id.n_ids) continue;
                const uint32_t o = in.tokens_to_text_id.off[id], len = in.tokens_to_text_id.off[id + 1] - o;
                if (!len) continue;
                if (n == kTlMaxLists) {
                    overflow = true;
                    break;
                }
                s_begin[n] = o, s_len[n] = len;
                if (n == 0 || len > s_len[longest]) longest = n;
                ++n;
            }
        }
        if (overflow) {
            if (blockIdx.y == 0) req_error[in.request] = 1;
            n = 0;
        }
        // prefix of the candidate counts (the longest list contributes none)
        uint32_t acc = 0;
        for (uint32_t i = 0; i < n; ++i) {
            s_prefix[i] = acc;
            if (i != longest) acc += s_len[i];
        }
        s_prefix[n] = acc;
        s_n = n, s_terms = terms, s_longest = longest;
    }
    __syncthreads();
    const uint32_t n = s_n, longest = s_longest;
    if (s_terms <= 1 || n <= 1) return;  // boost.rs:36-38
    const uint32_t* val = in.tokens_to_text_id.val;
    const uint32_t total = s_prefix[n];
    for (uint32_t x = blockIdx.y * blockDim.x + threadIdx.x; x < total; x += gridDim.y * blockDim.x) {
        uint32_t lo = 0, hi = n;  // list of candidate x: last i with prefix[i] <= x (skipping the longest, whose range is empty)
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_prefix[mid] <= x) lo = mid;
            else hi = mid;
        }
        uint32_t i = lo;
        if (i == longest) continue;  // cannot happen: its range is empty and a later list starts at the same prefix
        const uint32_t e = val[s_begin[i] + (x - s_prefix[i])];
        uint32_t c = 1;
        bool first = true;
        for (uint32_t j = 0; j < n && first; ++j) {
            if (j == i) continue;
            const uint32_t* l = val + s_begin[j];
            uint32_t a0 = 0, a1 = s_len[j];
            while (a0 < a1) {
                const uint32_t mid = (a0 + a1) >> 1;
                if (l[mid] < e) a0 = mid + 1;
                else a1 = mid;
            }
            if (a0 < s_len[j] && l[a0] == e) {
                if (j < i && j != longest) first = false;  // an earlier candidate list emits it
                ++c;
            }
        }
        if (!first || c <= 1) continue;
        const float boost = 2.0f * (float)c * (float)c;
        const uint32_t key = ~vbit::score_key(boost);  // the leaf keeps the maximum key = the minimum boost (boost.rs:23-28)
        if (in.identity) {
            list_emit(a, in.list_part, e, key);
        } else if (e < in.text_id_to_anchor.n_ids) {
            for (uint32_t k = in.text_id_to_anchor.off[e]; k < in.text_id_to_anchor.off[e + 1]; ++k) list_emit(a, in.list_part, in.text_id_to_anchor.val[k], key);
        }
    }
}

void launch_text_locality(cudaStream_t st, const TlInstance* inst, uint32_t n_inst, const uint32_t* term_parts, uint32_t* req_error, const ListArgs& a) {
    if (!n_inst) return;
    text_locality_kernel<<<dim3(n_inst, 16), 256, 0, st>>>(inst, term_parts, req_error, a);
    count_launch();
}

// One block per member: every matched token -> its text ids -> their parent value ids -> those with a boost value, filed
// under their anchor.  apply_boost_values_anchor (boost.rs:255-281) walks the (anchor, value) list, which is in value-id
// order, next to the hits, so the entry carries the value id (as 0x7FFFFFFF - id: the tile scatter keeps the largest key =
// the smallest value id of an anchor, and notes in bit 31 whether the anchor has several).
__global__ void __launch_bounds__(128) boost_to_anchor_kernel(const BoostListMember* __restrict__ members, ListArgs a) {
    const BoostListMember m = members[blockIdx.x];
    if (!m.tokenized && !m.use_ids) return;
    const uint32_t b = a.part_begin[m.part], n = a.part_begin[m.part + 1] - b;
    // The member's 32 warps share the texts of every matched token (32 texts per warp and step); the value ids of a text
    // are walked by its lane, or by the whole warp when there are many (a frequent token is in tens of thousands of texts,
    // a frequent text the value of as many documents).
    const uint32_t lane = threadIdx.x & 31u, warps = (blockDim.x >> 5) * gridDim.y, warp = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (uint32_t i = 0; i < n; ++i) {  // every warp visits every token and takes its share of the token's texts
        const uint32_t token = a.g_term[b + i];
        uint32_t t0 = 0, t1 = 1;  // text ids: the token's row, or the token itself
        bool self = true;
        if (m.tokenized && token < m.tokens_to_text_id.n_ids) {
            const uint32_t o0 = m.tokens_to_text_id.off[token], o1 = m.tokens_to_text_id.off[token + 1];
            if (o1 > o0) t0 = o0, t1 = o1, self = false;
        }
        for (uint32_t tb = t0 + 32u * ((warp + warps - i % warps) % warps); tb < t1; tb += 32u * warps) {
            uint32_t j0 = 0, j1 = 0;  // this lane's text: its value ids
            if (tb + lane < t1) {
                const uint32_t text = self ? token : m.tokens_to_text_id.val[tb + lane];
                if (text < m.value_id_to_parent.n_ids) j0 = m.value_id_to_parent.off[text], j1 = m.value_id_to_parent.off[text + 1];
            }
            auto emit_value = [&](uint32_t j) {
                const uint32_t value_id = m.value_id_to_parent.val[j];
                if (value_id >= m.column_n || value_id >= m.value_id_to_anchor.n_ids) return;
                if (m.column[value_id] == kNoValue) return;
                const uint32_t o = m.value_id_to_anchor.off[value_id];
                if (m.value_id_to_anchor.off[value_id + 1] == o) return;
                if (value_id < 0x7FFFFFFEu) list_emit(a, m.list_part, m.value_id_to_anchor.val[o], 0x7FFFFFFFu - value_id);
            };
            const bool long_range = j1 - j0 > 16;
            if (!long_range)  // few values: the lane walks its own text
                for (uint32_t j = j0; j < j1; ++j) emit_value(j);
            uint32_t todo = __ballot_sync(0xFFFFFFFFu, long_range);
            while (todo) {  // many values: the warp walks the text together
                const int src = __ffs((int)todo) - 1;
                todo &= todo - 1;
                const uint32_t r0 = __shfl_sync(0xFFFFFFFFu, j0, src), r1 = __shfl_sync(0xFFFFFFFFu, j1, src);
                for (uint32_t j = r0 + lane; j < r1; j += 32) emit_value(j);
            }
        }
    }
}

void launch_boost_to_anchor(cudaStream_t st, const BoostListMember* members, uint32_t n_members, const ListArgs& a) {
    if (!n_members) return;
    boost_to_anchor_kernel<<<dim3(n_members, 8), 128, 0, st>>>(members, a);
    count_launch();
}

void launch_phrase_pairs(cudaStream_t st, const PhraseMember* members, uint32_t n_members, const ListArgs& a) {
    if (!n_members) return;
    phrase_pairs_kernel<<<n_members, 128, 0, st>>>(members, a);
    count_launch();
}

// One thread per request: for every `and` node whose inputs are all search parts, the
// input with the fewest (estimated) hits is summed last and the former last input takes
// its place (swap_remove), as in intersect_hits_score (set_op.rs:388-417).
__global__ void finalize_programs_kernel(const QueryProgram* __restrict__ queries, uint32_t n, uint32_t* __restrict__ prog, const uint32_t* __restrict__ leaf_part,
                                         const unsigned long long* __restrict__ part_est, unsigned long long* __restrict__ stat_postings) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const QueryProgram qp = queries[q];
    if (!qp.active) return;
    if (stat_postings) {
        unsigned long long sum = 0;
        for (uint32_t l = 0; l < qp.n_leaves; ++l) sum += part_est[leaf_part[qp.leaf_begin + l]];
        if (sum) atomicAdd(stat_postings, sum);
    }
    uint32_t* code = prog + qp.prog_begin;
    uint32_t pc = 0;
    while (pc < qp.prog_len) {
        const uint32_t op = code[pc];
        if (op == kOpLeaf) pc += 2;
        else if (op == kOpUnion) pc += 3 + code[pc + 1];
        else if (op == kOpFilter) pc += 1;
        else if (op == kOpLeafBoost) pc += 4;
        else {
            const uint32_t cnt = code[pc + 1];
            uint32_t* order = code + pc + 2;
            const uint32_t* child_leaf = code + pc + 2 + cnt;
            uint32_t shortest = 0;
            unsigned long long best = ~0ull;
            bool known = true;
            for (uint32_t i = 0; i < cnt; ++i) {
                if (child_leaf[i] == kNoValue) {
                    known = false;
                    continue;
                }
                const unsigned long long est = part_est[leaf_part[qp.leaf_begin + child_leaf[i]]];
                if (est < best) best = est, shortest = i;
            }
            if (!known) shortest = cnt - 1;  // nested inputs: lengths unknown, keep request order
            uint32_t w = 0;
            for (uint32_t i = 0; i + 1 < cnt; ++i) order[w++] = (i == shortest) ? cnt - 1 : i;
            order[cnt - 1] = shortest;
            pc += 2 + 2 * cnt;
        }
    }
}
void launch_finalize_programs(cudaStream_t st, const QueryProgram* queries, uint32_t n, uint32_t* prog, const uint32_t* leaf_part, const unsigned long long* part_est,
                              unsigned long long* stat_postings) {
    if (!n) return;
    finalize_programs_kernel<<<(n + 127) / 128, 128, 0, st>>>(queries, n, prog, leaf_part, part_est, stat_postings);
    count_launch();
}

void launch_group_count(cudaStream_t st, const MatchRecord* rec, uint32_t n, uint32_t* part_count) {
    if (!n) return;
    group_count_kernel<<<(n + 255) / 256, 256, 0, st>>>(rec, n, part_count);
    count_launch();
}
void launch_scan_u32(cudaStream_t st, const uint32_t* in, uint32_t* out, uint32_t n) {
    scan_u32_kernel<<<1, 1024, 0, st>>>(in, out, n);
    count_launch();
}
void launch_scan_u64(cudaStream_t st, const uint64_t* in, uint64_t* out, uint32_t n) {
    scan_u64_kernel<<<1, 1024, 0, st>>>(in, out, n);
    count_launch();
}
// explain (search_field.rs:429-441): the posting weight of every (matched term, result anchor) pair, by binary search in the
// term's posting list (anchors ascending); -1 where the term has no posting on the anchor.  One thread per pair.
__global__ void posting_lookup_kernel(PostingsView pv, const uint32_t* __restrict__ terms, uint32_t n_terms, const uint32_t* __restrict__ anchors, uint32_t n_anchors,
                                      float* __restrict__ weight) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_terms * n_anchors) return;
    const uint32_t term = terms[i / n_anchors], anchor = anchors[i % n_anchors];
    float w = -1.0f;
    if (term < pv.n_terms) {
        uint64_t lo = pv.off[term], hi = pv.off[term + 1];
        while (lo < hi) {
            const uint64_t mid = lo + ((hi - lo) >> 1);
            if (pv.post[mid].anchor < anchor) lo = mid + 1;
            else hi = mid;
        }
        if (lo < pv.off[term + 1] && pv.post[lo].anchor == anchor) w = pv.post[lo].weight;
    }
    weight[i] = w;
}
void launch_posting_lookup(cudaStream_t st, const PostingsView& pv, const uint32_t* terms, uint32_t n_terms, const uint32_t* anchors, uint32_t n_anchors, float* weight) {
    const uint64_t n = (uint64_t)n_terms * n_anchors;
    if (!n) return;
    posting_lookup_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(pv, terms, n_terms, anchors, n_anchors, weight);
    count_launch();
}
void launch_score_scatter(cudaStream_t st, const ScoreScatterArgs& a) {
    if (!a.n_records) return;
    score_scatter_kernel<<<(a.n_records + 127) / 128, 128, 0, st>>>(a);
    count_launch();
}
void launch_dense_tile_offsets(cudaStream_t st, const DenseOffsetsArgs& a, uint32_t n_rows) {
    if (!n_rows) return;
    dense_tile_offsets_kernel<<<n_rows, 256, 0, st>>>(a);
    count_launch();
}
void launch_sparse_count(cudaStream_t st, const SparseArgs& a) {
    if (!a.n_matches) return;
    const uint64_t threads = (uint64_t)a.n_matches * 32;
    sparse_count_kernel<<<dim3((unsigned)((threads + 255) / 256), std::max<uint32_t>(1, (a.max_df + kSparseSegment - 1) / kSparseSegment)), 256, 0, st>>>(a);
    count_launch();
}
void launch_sparse_scan(cudaStream_t st, uint32_t* bucket, uint32_t n_tiles, uint64_t* sparse_total, uint32_t n_parts) {
    if (!n_parts) return;
    sparse_scan_kernel<<<n_parts, 256, 0, st>>>(bucket, n_tiles, sparse_total);
    count_launch();
}
void launch_sparse_fill(cudaStream_t st, const SparseArgs& a) {
    if (!a.n_matches) return;
    const uint64_t threads = (uint64_t)a.n_matches * 32;
    sparse_fill_kernel<<<dim3((unsigned)((threads + 255) / 256), std::max<uint32_t>(1, (a.max_df + kSparseSegment - 1) / kSparseSegment)), 256, 0, st>>>(a);
    count_launch();
}
void launch_part_slices(cudaStream_t st, PartSlices* out, const uint32_t* part_begin, const uint32_t* dense_cursor, const uint64_t* sparse_base, const PartQuery* parts, uint32_t n_parts) {
    if (!n_parts) return;
    part_slices_kernel<<<(n_parts + 255) / 256, 256, 0, st>>>(out, part_begin, dense_cursor, sparse_base, parts, n_parts);
    count_launch();
}

}  // namespace vdev
