// Plain structs shared by the host runtime and the kernels (device pointers inside).
#pragma once
#include <cstdint>

namespace vdev {

static const uint32_t kDictTile = 32;       // terms per dictionary tile: one warp, one term per lane
static const uint32_t kTilePrefixMax = 8;   // symbols of the tile's common prefix that are kept
static const uint32_t kNoValue = 0xFFFFFFFFu;
static const uint32_t kMaxLeaves = 64;      // search parts per request tree on the tile path (one shared-memory array of a tile's anchors each)
static const uint32_t kMaxK = 256;          // top + skip whose heap is merged in shared memory (and the limit on sharded handles)
static const uint32_t kMaxKLarge = 4096;    // top + skip on one device: beyond kMaxK the heap is merged in global memory (tile path only)
static const uint32_t kTileBlocksPerSm = 16;  // upper bound of resident tile_eval CTAs per SM (sizes the large-k merge scratch)

struct TilePrefix {  // 20 bytes
    uint16_t sym[kTilePrefixMax];
    uint16_t len;
    uint16_t pad;
};

// Deletion-neighbourhood index of a dictionary (lower-cased symbols): every term filed under the hash of each
// of its variants with at most max_del symbols deleted.  Two strings within edit distance d (transpositions
// included) share a variant with at most d deletions each, so probing the query's variants yields a superset of
// the matching terms; candidates are verified with the bit-parallel automaton.
struct DelEntry {
    uint32_t slot;  // dictionary slot of the term
    uint32_t tag;   // upper half of the variant hash
};
struct DelIndexView {
    const uint32_t* off;   // [mask + 2] bucket offsets (nullptr: not built)
    const DelEntry* ent;
    uint32_t mask;         // buckets - 1
    uint32_t max_del;
};

struct DictView {
    uint32_t n, n_tiles;
    DelIndexView del[2];   // [0] one deletion, [1] two deletions
    const uint32_t* ids;
    const uint16_t* lower_bytes;
    const uint16_t* sym[2];  // [0] lower-cased, [1] raw
    const uint32_t* off[2];
    const TilePrefix* tiles[2];
    // The few terms whose Rust `to_lowercase` differs from the scalar-by-scalar lowering of sym[0] (U+0130 expands, a
    // word-final capital sigma becomes U+03C2): their exact lower-cased symbols, for scoring (search_field.rs:312).
    const uint32_t* exc_slot;  // sorted dictionary slots
    const uint32_t* exc_off;   // [n_exc + 1] into exc_sym
    const uint16_t* exc_sym;
    uint32_t n_exc, exc_pad;
};

struct Posting {  // 8 B, one 64-bit load
    uint32_t anchor;
    float weight;  // AnchorScore.score.to_f32() / 100.0 (search_field.rs:426), exact: divided once at load
};

struct PostingsView {
    const Posting* post;
    const uint64_t* off;
    uint32_t n_terms;
    const uint32_t* term_plane;  // per term id: head-term plane of the index, kNoValue if the term has none
};

// ---- term planes ----
// The frequent terms of the index also exist anchor-indexed: one presence bit per anchor of the shard
// ("plane").  The plane path (planes.cu) evaluates them with word-wide bit operations instead of
// scattering their postings.  Two classes, by document frequency over the shard's anchors:
//   head planes  df >= span / 128  (at most kMaxHeadPlanes): bits + the f16 score per anchor
//   mid planes   df >= span / 1024 (at most kMaxPlanes in total): bits only; the weight of an anchor is found by
//                binary search in the term's posting list, which only exact evaluations (rare) need
// Planes are ordered by descending df over all fields; plane ids below n_head are head planes.
static const uint32_t kMaxHeadPlanes = 96;
static const uint32_t kMaxPlanes = 2048;
static const uint32_t kPlaneTileLog2 = 13;   // planes are padded to whole tiles of 2^13 anchors
struct PlaneInfo {           // posting list of the plane's term (this shard's part of it)
    const struct Posting* post;
    uint32_t df;
    uint32_t pad;
};
struct PlaneSetView {
    const uint32_t* bits;    // [n_planes][words]
    const uint16_t* score;   // [n_head][words * 32] f16 bits of AnchorScore.score; weight = f16 / 100 (search_field.rs:426)
    const float* wmax;       // [n_planes] largest weight of the plane
    const PlaneInfo* info;   // [n_planes]
    const uint32_t* tcount;  // [n_planes][words / 256] anchors of the plane per tile of 2^13 anchors
    const uint32_t* tprefix; // [n_planes][words / 256 + 1] its exclusive prefix sums: the plane term's posting offset at every tile
                             // boundary (a plane's bits and its term's postings of the shard correspond one to one)
    uint32_t n_planes, n_head;
    uint32_t words;          // 32-anchor words per plane (multiple of 2^kPlaneTileLog2 / 32)
    uint32_t pad;
};
static const uint32_t kPlaneRow = 0xFFFFFFFEu;  // g_row of a matched plane term: its tile offsets are the plane's tprefix row

// Nested "value >= threshold" bitmaps of a boost column (anchors without a value are set in
// every level): lets the plane path restrict the exact evaluation to anchors whose boost
// multiplier can still reach the request's running k-th best.
static const uint32_t kBoostLevels = 16;
struct ColumnLevels {
    const uint32_t* bits;  // [kBoostLevels][words], word 0 = anchors [anchor_lo, anchor_lo + 32)
    uint32_t words;
    uint32_t pad;
    float thr[kBoostLevels];  // ascending
    // Seed set: the shard's anchors whose value is in the top 1/8 of the column (>= thr[kSeedLevel]), by descending value, and every
    // plane restricted to them (bit i of a plane's row = seed_anchor[i] is in the plane).  One pass over these rows -- an
    // eighth of a plane sweep -- finds, for the whole shard, the well-boosted anchors that have every plane part of a
    // request: their scores seed the request's threshold before the sweep starts (plane_seed_kernel).
    const uint32_t* seed_anchor;  // relative to the shard
    const uint32_t* seed_bits;    // [n_planes][seed_words]
    uint32_t seed_n, seed_words;  // seed_words: 32-anchor words per row, a multiple of 128
};
static const uint32_t kSeedLevel = 2;

static const uint32_t kPartPlaneSlots = 8;
struct PartPlanes {  // plane-term matches of one search part (80 B)
    uint32_t n;      // matches registered (more than kPartPlaneSlots: the part cannot take the plane path)
    uint32_t plane[kPartPlaneSlots];
    float ts[kPartPlaneSlots];  // term score of the match
    uint32_t pad[3];
};

// Everything the plane path needs to know about one request.
static const uint32_t kFastMaxLeaves = 4;
static const uint32_t kFastMaxK = 64;
static const uint32_t kFastMaxTerms = 12;      // plane terms of one request, over all its parts
static const uint32_t kGroupMaxEntries = 512;  // postings of non-plane terms in one (tile group, request) item
enum FastFlags : uint32_t { kFastOk = 1u, kFastBoost = 2u, kFastUnion1 = 4u };
struct alignas(16) FastDesc {  // 176 B
    uint32_t flags, n_leaves, k, fb_fun;
    float fb_param, fb_max_mult;
    uint32_t fb_n, n_terms;
    float bound[kFastMaxLeaves];  // bound[n-1]: no anchor with n parts present, all through planes, scores above it (before the boost)
    float ub[kFastMaxLeaves];     // per part: largest score a plane match can contribute
    const uint32_t* fb_col;
    const ColumnLevels* fb_lev;
    float ts[kFastMaxTerms];       // plane terms, grouped by part in part order: term score of the match,
    uint16_t plane[kFastMaxTerms]; // its plane,
    uint8_t part[kFastMaxTerms];   // its part
    uint8_t np[kFastMaxLeaves];    // plane terms per part
    uint8_t pad[8];
};
static_assert(sizeof(FastDesc) == 176, "FastDesc layout");

// One (tile group, request) item of the plane path: per part, the request's entries (postings of non-plane terms
// inside the group's tiles) are one contiguous range of the part's sparse tile buckets.
struct alignas(16) FastItem {  // 32 B
    uint32_t q;
    uint32_t tiles;                  // first tile | number of tiles << 24 (at most the group size; fewer where the entries of a whole group exceed kGroupMaxEntries)
    uint16_t n[kFastMaxLeaves];
    uint32_t begin[kFastMaxLeaves];  // index into the batch's SparseEntry array
};
static_assert(sizeof(FastItem) == 32, "FastItem layout");

struct SparseEntry {  // 8 B: one posting of a sparse (rarely matched) term, already scored
    uint32_t anchor;
    uint32_t key;  // score_key(term_score * weight)
};

struct CsrView {
    const uint32_t* off;
    const uint32_t* val;
    uint32_t n_ids;
};

// ---- fuzzy match (get_term_ids_in_field, search_field.rs:277-398) ----
enum PartFlags : uint32_t {
    kPartPrefix = 1u,         // starts_with
    kPartTransposition = 2u,  // matching automaton built with transposition_cost_one (search_field.rs:87)
    kPartRawCase = 4u,        // ignore_case == false: match on raw scalars
    kPartCheckPrefix = 8u,    // starts_with || levenshtein != 0  (:302)
    kPartHasBoost = 16u,
    kPartInjected = 128u,     // its (term id, score) hits are given by the host (per-part top/skip bound), not matched in this batch
    kPartList = 32u,          // not a search part: its tile bucket is filled by a list producer (phrase pairs, text locality, 1:n boosts)
    kPartListBoost = 64u,     // list part of a 1:n boost: entries carry 0x7FFFFFFF - value id; the tile keeps the smallest value id of an
                              // anchor plus (bit 31) whether the anchor has several
    kPartAnySign = 512u,      // a token_value boost follows the match (search_field.rs:391-395): the given scores may be negative
    kPartRegex = 256u,        // is_regex (search_field.rs:72-83): matched by regex_match_kernel with the part's DFA, scored like any other part
};

// One regex search part on the device: its DFA over the alphabet codes of the part's dictionary (host/regex_dfa.hpp).
struct RegexPartDev {
    const uint16_t* class_of_code;  // [alphabet size] scalar class of every alphabet code
    const uint16_t* trans;          // [n_states * n_classes] next state, bit 15 set when that state is a match state; state 0 is dead
    uint32_t n_classes;
    uint32_t start;                 // start state (bit 15: it matches the empty string)
    uint32_t part;
    uint32_t sticky;                // starts_with: a match after any prefix of the term counts (fst::automaton::StartsWith)
};

struct PartQuery {  // one distinct RequestSearchPart of the batch (272 B)
    uint16_t match_sym[64];  // query scalars as alphabet codes, in the matching case variant
    uint16_t score_sym[64];  // lower-cased query (scoring always runs on lower-cased text, :298-317)
    uint32_t m;              // scalars in the query
    uint32_t d_match;        // min(d, 4)  (:87)
    uint32_t d_score;        // d clamped to chars-1 (:286)
    uint32_t flags;
    float boost;             // per-part scalar boost (:359-364)
    uint32_t lower_bytes;    // byte length of the lower-cased query (distance() 255 rule, :706)
    uint32_t postings;       // index into the batch's PostingsView table, kNoValue if absent
    uint32_t m_score;        // scalars of the lower-cased query (differs from m only where to_lowercase changes the length)
};

struct MatchRecord {  // fuzzy_match output, unordered
    uint32_t part;
    uint32_t slot;  // dictionary slot
};

// ---- per-part posting slices ----
struct PartSlices {
    uint32_t m_begin;     // first grouped match of the part
    uint32_t n_match;     // matches (= matched dictionary terms)
    uint32_t n_dense;     // the first n_dense grouped matches have a tile-offset row
    uint32_t sparse_row;  // row of the part in the sparse bucket table
    uint64_t sparse_base; // first sparse entry of the part
};

// ---- non-empty (tile, request) work items and their posting slices (item_scan_kernel -> tile_eval_kernel) ----
struct ItemRec {  // 24 B
    uint32_t q, t;
    unsigned long long slice_begin;
    uint32_t n_slices, npost;
};
struct SliceRec {  // 32 B
    unsigned long long begin;  // dense: first posting (index into the PostingsView array); sparse: first bucket entry
    uint32_t n;
    float term_score;          // dense: score of the matched term
    uint32_t task_begin;       // warp tasks (runs of kTaskPostings postings) of the item before this slice
    uint16_t leaf;
    uint8_t kind;              // 0 dense posting slice, 1 sparse bucket slice
    uint8_t single;            // the part has exactly one matched term: plain stores suffice
    uint32_t postings;         // PostingsView index (dense)
    uint32_t pad;
};
static const uint32_t kTaskPostings = 128;  // postings one warp takes from a slice at a time

// ---- boosts (boost.rs:283-377, 470-504) ----
enum BoostFunDev : uint32_t { kBoostNone = 0, kBoostLog2 = 1, kBoostLog10 = 2, kBoostMultiply = 3, kBoostAdd = 4, kBoostReplace = 5 };
enum ExprOp : uint32_t { kExprNone = 0, kExprDiv = 1, kExprMul = 2, kExprAdd = 3, kExprSub = 4 };

static const uint32_t kMaxSkipWhenScore = 16;  // values of RequestBoostPart::skip_when_score a step carries
struct BoostStep {  // add_boost on anchor ids (boost.rs:470-504, apply_boost :283-377)
    const uint32_t* column;
    uint32_t n;
    uint32_t fun;
    float param;
    uint32_t n_skip;
    float skip[kMaxSkipWhenScore];
    uint32_t expr_op;       // `x op y` expression, operands: $SCORE (= boost value) or a float
    uint32_t expr_left_is_score, expr_right_is_score;
    float expr_left, expr_right;
    // upper bound of the multiplier this step can apply (Log10/Log2/Multiply over a non-negative column):
    // lets the tile kernel skip the gather for anchors that cannot reach the running k-th best
    uint32_t can_prune;
    float max_mult;
    const ColumnLevels* levels;  // level bitmaps of the column (nullptr: none)
    uint32_t list_only;          // a 1:n boost: applied by kOpLeafBoost with the value id a list part supplies, never per anchor
    uint32_t pad;
};

// ---- list producers: anchors (with a value) computed from the matched terms of search parts ----
struct PhraseView {  // persistence_data_binary_search.rs:126-203, flattened
    const uint64_t* keys;     // sorted (term1 << 32 | term2)
    const uint32_t* off;      // [n + 1]
    const uint32_t* anchors;
    uint32_t n;
};
struct PhraseMember {  // one phrase_boosts entry: every (term of part1, term of part2) pair's anchors go to `list_part`
    uint32_t part1, part2, list_part, pad;
    PhraseView store;
};

struct IdsMember {  // hits_ids of a part resolved to anchors: text_id_to_anchor, or the id itself on an anchor-identity column (search_field.rs:468-498)
    uint32_t part, list_part, identity, pad;
    CsrView text_id_to_anchor;
};

// BoostToAnchor (plan_steps.rs:173-196): matched tokens -> text ids -> parent value ids -> (boost value, anchor)
struct BoostListMember {
    uint32_t part, list_part;
    uint32_t tokenized;        // tokens resolve through tokens_to_text_id (a token without entry is its own text id)
    uint32_t use_ids;          // untokenized field: the matched term ids are the text ids only when the part is also searched for ids
    CsrView tokens_to_text_id, value_id_to_parent, value_id_to_anchor;
    const uint32_t* column;    // boost values by value id
    uint32_t column_n, pad;
};

// boost_text_locality (boost.rs:34-87) of one (request, field): the matched tokens of every query term -> text ids;
// a text id reached c > 1 times boosts its anchors by 2 * c * c.
static const uint32_t kTlMaxLists = 256;
struct TlInstance {
    uint32_t list_part;
    uint32_t term_begin, n_terms;   // into the table of parts (one per distinct query term of the field)
    uint32_t identity;              // anchor-identity column: text id == anchor
    uint32_t request;
    uint32_t pad[3];
    CsrView tokens_to_text_id, text_id_to_anchor;
};

// ---- request programs (plan_creator, execution_plan.rs:132-534) ----
enum ProgOp : uint32_t {
    kOpLeaf = 1,       // [op, leaf index]
    kOpUnion = 2,      // [op, n children, n slots, slot of child 0 .. n-1]      set_op.rs:87-220
    kOpIntersect = 3,  // [op, n children, sum order: child index 0 .. n-1]      set_op.rs:368-446
    kOpFilter = 4,     // [op]: (search result, filter result) -> search result where the filter is present   set_op.rs:311-326
    kOpLeafBoost = 5,  // [op, leaf, list leaf, boost step]: the leaf, boosted with the boost value of the list's value id for the anchor
                       // (BoostToAnchor + ApplyAnchorBoost, plan_steps.rs:173-233)
};

// Steps that run on every hit after the request tree and the column boosts, in the reference's order
// (execution_plan.rs:202-262 phrase boosts, search.rs:176 boost_term, search.rs:180-184 text locality).
enum PostOp : uint32_t {
    kPostMulIfPresent = 1,  // [op, leaf, f32 bits]: score *= value when the leaf has the anchor   boost.rs:380-402
    kPostMulValue = 2,      // [op, leaf]: score *= the leaf's value for the anchor               boost.rs:197-237
};

// One facet of one request (facet.rs:31-73): hit anchor -> value ids through up to three id -> ids joins, counted.
static const uint32_t kMaxFacetSteps = 3;
struct FacetStep {
    CsrView step[kMaxFacetSteps];
    uint32_t n_steps;
    uint32_t hist_size;
    uint32_t* hist;  // [hist_size] counts of this (request, facet)
};


struct alignas(16) QueryProgram {  // 112 B
    uint32_t leaf_begin, n_leaves;  // into the leaf -> part table
    uint32_t prog_begin, prog_len;  // into the program words
    uint32_t boost_begin, n_boosts; // into the BoostStep table
    uint32_t k;                     // top + skip (0: only count)
    uint32_t active;                // 0 = request failed on the host, skip
    uint32_t emit_all;              // step seam: also write every hit to the emit buffer
    uint32_t nonneg;                // every part score of the request is >= 0 (no negative part boost): cheap key decode
    // the request's only boost step when it has neither skip list nor expression ("fast boost"), inlined
    uint32_t fb_flags;              // bit 0: fast boost present, bit 1: max_mult bounds the multiplier (can prune)
    uint32_t fb_n, fb_fun;
    float fb_param, fb_max_mult;
    uint32_t union1;                // one leaf standing for an `or` of identical parts: the union rule applies (score * n * n with n in {0, 1}), not the passthrough
    const uint32_t* fb_col;
    const ColumnLevels* fb_lev;     // level bitmaps of the fast-boost column (nullptr: none)
    uint32_t post_begin, post_len;  // post ops, into the program words
    uint32_t facet_begin, n_facets; // into the FacetStep table
    uint32_t n_leaf_boosts;         // kOpLeafBoost ops in the program (at most kMaxLeafBoosts)
    uint32_t must_mask;             // bit l: leaf l is a direct search-part child of a root `and`: a tile where it has no posting has no hit
    uint32_t pad2[2];
};
static_assert(sizeof(QueryProgram) == 112, "QueryProgram layout");
static const uint32_t kMaxLeafBoosts = 4;

}  // namespace vdev
