// Bit-parallel (Myers / Hyyro) edit distance over 64-bit words, shared by the
// fuzzy-match kernels (device) and their CPU unit tests (host, via
// host/selftest.cpp).  The query is the pattern (<= 64 symbols, bit j = query
// position j); dictionary-term symbols are streamed one column at a time.
//
// Replaces the Levenshtein DFA of veloci_levenshtein_automata that the reference
// intersects with the FST (src/search/search_field.rs:54-99) and the scoring DFA /
// `distance()` fallback (:691-732): same accepted language, computed per term.
//   * global distance D[m][n] (row 0 = 0..n, so the horizontal carry-in is +1)
//   * optional adjacent transposition at cost one (restricted Damerau, the
//     `transposition_cost_one` flag of LevenshteinAutomatonBuilder::new)
//   * prefix mode (`starts_with`): accept once any prefix of the term is within d
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define VB_HD __host__ __device__ __forceinline__
#else
#define VB_HD inline
#endif

namespace vbit {

static const uint16_t kNoSymbol = 0xFFFFu;  // query scalar absent from the dictionary alphabet
static const int kMaxQuery = 64;

struct Query {
    uint16_t sym[kMaxQuery];
    uint32_t m;
};

// Eq mask of one dictionary symbol against the query symbols.
VB_HD uint64_t eq_mask(const uint16_t* qsym, uint32_t m, uint16_t c) {
    uint64_t eq = 0;
    for (uint32_t j = 0; j < m; ++j) eq |= (uint64_t)(qsym[j] == c) << j;
    return eq;
}

template <class W>
struct StateT {
    W vp, vn;        // vertical +1 / -1 deltas of the current column
    W d0, eq;        // previous column's diagonal-zero vector and Eq (transposition)
    uint32_t score;  // D[m][columns consumed]
    uint32_t cols;   // columns consumed
    uint32_t best;   // min over consumed columns (incl. column 0) of D[m][.]  (prefix mode)
};
typedef StateT<uint64_t> State;

template <class W>
VB_HD W low_mask_t(uint32_t m) {
    return m >= sizeof(W) * 8 ? ~(W)0 : (((W)1 << m) - (W)1);
}
VB_HD uint64_t low_mask(uint32_t m) { return low_mask_t<uint64_t>(m); }

template <class W>
VB_HD void init(StateT<W>& s, uint32_t m) {
    s.vp = low_mask_t<W>(m);
    s.vn = 0;
    s.d0 = 0;
    s.eq = 0;
    s.score = m;
    s.cols = 0;
    s.best = m;
}

// Consumes one dictionary symbol whose Eq mask is `eq`.
template <class W>
VB_HD void step(StateT<W>& s, W eq, uint32_t m, bool transposition) {
    s.cols += 1;
    if (m == 0) {
        s.score = s.cols;
        return;
    }
    const W mask = low_mask_t<W>(m);
    W x = eq | s.vn;
    W d0 = ((((x & s.vp) + s.vp) ^ s.vp) | x) & mask;
    if (transposition) d0 |= (((~s.d0) & eq) << 1) & s.eq;
    d0 &= mask;
    W hn = s.vp & d0;
    W hp = (s.vn | ~(s.vp | d0)) & mask;
    const W top = (W)1 << (m - 1);
    s.score += (hp & top) ? 1u : 0u;
    s.score -= (hn & top) ? 1u : 0u;
    W xh = (hp << 1) | (W)1;  // row 0 grows by one per column (global distance)
    s.vn = xh & d0 & mask;
    s.vp = ((hn << 1) | ~(xh | d0)) & mask;
    s.d0 = d0;
    s.eq = eq;
    if (s.score < s.best) s.best = s.score;
}

// min over the cells of the current column: D[0] = cols, D[j] = D[j-1] + vp_j - vn_j.
// No extension of the consumed prefix can end below this value.
template <class W>
VB_HD uint32_t column_min(const StateT<W>& s, uint32_t m) {
    int32_t v = (int32_t)s.cols, best = (int32_t)s.cols;
    for (uint32_t j = 0; j < m; ++j) {
        v += (int32_t)((s.vp >> j) & 1) - (int32_t)((s.vn >> j) & 1);
        best = v < best ? v : best;
    }
    return (uint32_t)best;
}

// Distance between a whole term and the query.
VB_HD uint32_t distance(const uint16_t* qsym, uint32_t m, const uint16_t* tsym, uint32_t n, bool transposition) {
    State s;
    init(s, m);
    for (uint32_t i = 0; i < n; ++i) step(s, eq_mask(qsym, m, tsym[i]), m, transposition);
    return s.score;
}

// search_field.rs:27-33 get_default_score_for_distance
VB_HD float default_score(uint32_t dist, bool prefix_matches) {
    float d = (float)(dist & 0xFFu);
#if defined(__CUDA_ARCH__)
    if (prefix_matches) return 2.0f / (log2f(d + 1.0f) + 0.2f);
#else
    if (prefix_matches) return 2.0f / (__builtin_log2f(d + 1.0f) + 0.2f);
#endif
    return 2.0f / (d + 0.2f);
}

// Monotone u32 key of an f32 (larger float -> larger key); 0 is below every float.
VB_HD uint32_t score_key(float f) {
    uint32_t b;
#if defined(__CUDA_ARCH__)
    b = __float_as_uint(f);
#else
    __builtin_memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
VB_HD float key_score(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    float f;
#if defined(__CUDA_ARCH__)
    f = __uint_as_float(b);
#else
    __builtin_memcpy(&f, &b, 4);
#endif
    return f;
}

}  // namespace vbit
