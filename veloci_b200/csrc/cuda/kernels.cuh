// Launch wrappers of every kernel (implemented in fuzzy.cu / tiles.cu) and their argument blocks.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>

#include "device_types.cuh"

namespace vdev {

void count_launch();          // bumps the library-wide kernel launch counter
uint64_t launches_so_far();

// Kernel attributes (opt-in dynamic shared memory) are per device: one flag per device ordinal, so a process that
// opens indices on several GPUs configures every one of them.  first() is true exactly once per device.
struct PerDeviceOnce {
    static const int kMaxDevices = 64;
    std::atomic<uint32_t> done[kMaxDevices];
    PerDeviceOnce() {
        for (auto& d : done) d.store(0);
    }
    bool first() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return true;  // unknown device: always configure
        return done[dev].exchange(1u) == 0u;
    }
};

// ---- fuzzy.cu ----
void launch_fuzzy_match(cudaStream_t st, const DictView& dict, const PartQuery* parts, const uint32_t* part_ids, uint32_t n_parts, uint32_t max_m, MatchRecord* out,
                        uint32_t capacity, unsigned long long* counter);
// one launch per dictionary: grid.y = the regex parts on it, one thread per dictionary term
// get_anchor_for_phrases_in_field (search_field.rs:263-275) over explicit term id lists: thread x takes the pair
// (ids1[x / n2], ids2[x % n2]); `out` == nullptr counts the pair's anchors into pair_count[x], else copies them to out + pair_off[x]
void launch_phrase_lookup(cudaStream_t st, const PhraseView& store, const uint32_t* ids1, uint32_t n1, const uint32_t* ids2, uint32_t n2, uint32_t* pair_count, const uint32_t* pair_off,
                          uint32_t* out);
// One step of BoostToAnchor (plan_steps.rs:174-196) over explicit id lists: the values of every id in `store` (the id itself
// when it has none and `self_if_empty`); `out` == nullptr counts into count[i], else copies to out + off[i].
void launch_csr_expand(cudaStream_t st, const CsrView& store, const uint32_t* ids, uint32_t n, uint32_t self_if_empty, uint32_t* count, const uint32_t* off, uint32_t* out);
// get_boost_ids_and_resolve_to_anchor (boost.rs:432-468): per value id its boost value bits and (first) anchor, kNoValue anchor when it has no value
void launch_boost_values(cudaStream_t st, const uint32_t* column, uint32_t column_n, const CsrView& value_id_to_anchor, const uint32_t* value_ids, uint32_t n, uint32_t* out_anchor,
                         uint32_t* out_bits);
void launch_regex_match(cudaStream_t st, const DictView& dict, const RegexPartDev* parts, uint32_t n_parts, MatchRecord* out, uint32_t capacity, unsigned long long* counter);
// Deletion-neighbourhood index build: pass 0 counts the variants per hash bucket, pass 1 (after an exclusive scan of
// the counts into `off`, cursor zeroed) files the terms.
void launch_del_index_pass(cudaStream_t st, const DictView& dict, uint32_t max_del, uint32_t mask, uint32_t* count_or_cursor, const uint32_t* off, DelEntry* ent);
// Fuzzy match by probing the deletion-neighbourhood index: one warp per search part (not starts_with, lower-cased
// matching, distance <= 2).  Parts whose candidate set overflows the warp's table are appended to `overflow_parts`
// (count in *overflow_count) and have to be scanned.
void launch_fuzzy_probe(cudaStream_t st, const DictView& dict, const PartQuery* parts, const uint32_t* part_ids, uint32_t n_parts, MatchRecord* out, uint32_t capacity,
                        unsigned long long* counter, uint32_t* overflow_parts, unsigned long long* overflow_count);
void launch_group_count(cudaStream_t st, const MatchRecord* rec, uint32_t n, uint32_t* part_count);
void launch_scan_u32(cudaStream_t st, const uint32_t* in, uint32_t* out, uint32_t n);
void launch_scan_u64(cudaStream_t st, const uint64_t* in, uint64_t* out, uint32_t n);

struct ScoreScatterArgs {
    const MatchRecord* records;
    uint32_t n_records;
    const PartQuery* parts;
    const uint32_t* part_dict;   // part -> index into dicts
    const DictView* dicts;
    const PostingsView* postings;
    const uint32_t* part_begin;  // n_parts + 1
    uint32_t* dense_cursor;      // per part, zeroed
    uint32_t* sparse_cursor;     // per part, zeroed
    uint32_t* n_dense_rows;      // zeroed
    uint32_t dense_row_capacity;
    uint32_t dense_min;          // df >= dense_min -> tile-offset row
    uint32_t* row_match;         // dense row -> grouped match
    unsigned long long* part_est; // per part: sum of df over its matches (upper bound of its hit count), zeroed
    // grouped match arrays
    uint32_t* g_term;
    float* g_score;
    uint64_t* g_begin;
    uint32_t* g_df;
    uint32_t* g_row;
    uint32_t* g_part;
    // step seam (vgpu_resolve_to_anchor): hits are given, record.slot indexes these arrays
    const uint32_t* inj_terms;  // hits given by the host: records of injected parts index these arrays
    const float* inj_scores;
    uint32_t inj_all;           // every record is injected (vgpu_resolve_to_anchor); else only those of kPartInjected parts
    // head-term planes: matches of plane terms are also registered per part (zeroed before the launch)
    PartPlanes* part_planes;
    uint32_t* g_plane;  // per grouped match: plane id or kNoValue
};
void launch_score_scatter(cudaStream_t st, const ScoreScatterArgs& a);

struct DenseOffsetsArgs {
    const uint32_t* row_match;
    const uint32_t* g_part;
    const uint64_t* g_begin;
    const uint32_t* g_df;
    const PartQuery* parts;
    const PostingsView* postings;
    uint32_t* toff;  // [rows][n_tiles + 1]
    uint32_t n_tiles, tile_log2, anchor_lo;
};
void launch_dense_tile_offsets(cudaStream_t st, const DenseOffsetsArgs& a, uint32_t n_rows);

struct SparseArgs {
    uint32_t n_matches;
    const uint32_t* g_row;
    const uint32_t* g_df;
    const uint32_t* g_part;
    const uint64_t* g_begin;
    const float* g_score;
    const PartQuery* parts;
    const PostingsView* postings;
    uint32_t* bucket;  // [n_parts][n_tiles + 1]
    const uint64_t* sparse_base;
    SparseEntry* sparse;
    uint32_t n_tiles, tile_log2, anchor_lo;
    uint32_t max_df;  // largest posting list a sparse match can have (the lists are walked in segments by several warps)
};
void launch_sparse_count(cudaStream_t st, const SparseArgs& a);
void launch_sparse_scan(cudaStream_t st, uint32_t* bucket, uint32_t n_tiles, uint64_t* sparse_total, uint32_t n_parts);
void launch_sparse_fill(cudaStream_t st, const SparseArgs& a);
void launch_part_slices(cudaStream_t st, PartSlices* out, const uint32_t* part_begin, const uint32_t* dense_cursor, const uint64_t* sparse_base, const PartQuery* parts, uint32_t n_parts);

// List producers write into the sparse tile buckets of their list parts, in the same two passes as the sparse postings:
// count (sparse == nullptr) between launch_sparse_count and launch_sparse_scan, fill after launch_sparse_fill.
struct ListArgs {
    const uint32_t* part_begin;  // grouped matches of every part
    const uint32_t* g_term;      // matched term ids
    uint32_t* bucket;
    const uint64_t* sparse_base;
    SparseEntry* sparse;         // nullptr: count pass
    uint32_t n_tiles, tile_log2, anchor_lo, anchor_hi;
};
// get_anchor_for_phrases_in_field (search_field.rs:263-275): one block per member
void launch_phrase_pairs(cudaStream_t st, const PhraseMember* members, uint32_t n_members, const ListArgs& a);
// boost_text_locality (boost.rs:34-87): gridDim = (instances, chunks).  Requests with more than kTlMaxLists matched
// tokens in one field get req_error[request] = 1 (reported as VGPU_ERR_UNSUPPORTED).
void launch_text_locality(cudaStream_t st, const TlInstance* inst, uint32_t n_inst, const uint32_t* term_parts, uint32_t* req_error, const ListArgs& a);
// BoostToAnchor (plan_steps.rs:173-196): one block per member
void launch_boost_to_anchor(cudaStream_t st, const BoostListMember* members, uint32_t n_members, const ListArgs& a);
// the ids half of resolve_token_to_anchor (search_field.rs:468-498): one block per member
void launch_ids_to_anchor(cudaStream_t st, const IdsMember* members, uint32_t n_members, const ListArgs& a);

// ---- lists.cu: device-resident hit lists of the step seam ----
// per-tile offsets of an anchor-sorted list (row[0 .. n_tiles]), as prepare_lists builds them on the host for host lists
void launch_list_bucket(cudaStream_t st, const SparseEntry* entries, uint32_t n, uint32_t anchor_lo, uint32_t tile_log2, uint32_t n_tiles, uint32_t* row);
// the hits a step emitted ((key << 32) | anchor, unordered) as an anchor-sorted SparseEntry list in `out`; `buf` is overwritten, `alt` scratch
size_t emitted_sort_temp_bytes(uint32_t n);
cudaError_t launch_emitted_to_list(cudaStream_t st, unsigned long long* buf, unsigned long long* alt, unsigned long long* out, uint32_t n, void* temp, size_t temp_bytes);

// explain: weight[t * n_anchors + a] = posting weight of term t on anchor a, -1 without a posting (one thread per pair)
void launch_posting_lookup(cudaStream_t st, const PostingsView& pv, const uint32_t* terms, uint32_t n_terms, const uint32_t* anchors, uint32_t n_anchors, float* weight);

// Patches the sum order of every `and` node (set_op.rs:388-417: the shortest input is
// swap_remove'd and added last); input lengths are estimated by part_est.
// Also sums, over all requests, the postings of their matched terms into *stat_postings (traffic model).
void launch_finalize_programs(cudaStream_t st, const QueryProgram* queries, uint32_t n, uint32_t* prog, const uint32_t* leaf_part, const unsigned long long* part_est,
                              unsigned long long* stat_postings);

// ---- tiles.cu ----
// Classifies every (tile group, request) pair.  A request with a FastDesc takes the plane path in every group of
// `group_tiles` consecutive tiles where its non-plane terms have at most kGroupMaxEntries postings, all of them in
// sparse buckets: one plane-path item per (group, request), evaluated by plane_eval_kernel.  Everything else becomes
// general items, one per non-empty (tile, request), for tile_eval_kernel.  Pass 0 counts, pass 1 writes the records;
// plane-path items are grouped per group, general items form one flat tile-major list with their slice records.
struct ItemScanArgs {
    const QueryProgram* queries;
    uint32_t n_queries;
    const FastDesc* fast;  // nullptr: no plane path in this batch
    const uint32_t* leaf_part;
    const PartSlices* slices;
    const PartQuery* parts;
    const uint32_t* g_row;
    const uint32_t* g_plane;
    const uint64_t* g_begin;
    const float* g_score;
    const uint32_t* toff;
    const uint32_t* bucket;
    const uint32_t* plane_tprefix;     // [n_planes][n_tiles + 1] tile offsets of the plane terms' postings (nullptr without planes)
    uint32_t n_tiles;
    uint32_t group_tiles, n_groups;    // tiles per group (1 when there is no plane path), ceil(n_tiles / group_tiles)
    unsigned long long n_pairs_total;  // n_groups * n_queries
    unsigned long long* counters;      // [0] general items, [1] general slices (zeroed before each pass)
    ItemRec* items;                    // general items
    SliceRec* slice_recs;              // general slices
    // plane-path items, per group: pass 0 accumulates counts in the cursor, pass 1 uses it as cursor (zeroed before each pass)
    uint32_t* fast_item_cursor;        // [n_groups]
    const uint32_t* fast_item_begin;   // [n_groups + 1] (pass 1)
    FastItem* fast_items;
};
void launch_item_scan(cudaStream_t st, const ItemScanArgs& a, bool fill);

struct TileArgs {
    const ItemRec* items;
    const SliceRec* slice_recs;
    // batch programs
    const QueryProgram* queries;
    uint32_t n_queries;
    const uint32_t* leaf_part;
    const uint32_t* prog;
    const BoostStep* boosts;
    const FacetStep* facets;
    // parts and their slices
    const PartQuery* parts;
    const PartSlices* slices;
    const PostingsView* postings;
    const float* g_score;
    const uint64_t* g_begin;
    const uint32_t* g_row;
    const uint32_t* g_df;
    const uint32_t* toff;
    const uint32_t* g_plane;        // per grouped match: plane or kNoValue (nullptr: no planes in this batch)
    const uint32_t* plane_tprefix;  // tile offsets of the plane terms
    const uint32_t* bucket;
    const SparseEntry* sparse;
    // geometry
    uint32_t n_tiles, tile_log2, anchor_lo, anchor_hi;
    uint32_t max_leaves;  // shared-memory arrays per CTA
    // per-query state
    unsigned long long* heap;  // [n_queries][heap_stride] keys, sorted descending, 0 = empty
    uint32_t heap_stride;
    unsigned long long* merge_scratch;  // [n_sms * kTileBlocksPerSm][heap_stride]: merge buffer of requests with k > kMaxK (nullptr when none)
    unsigned long long* tau;   // current k-th best key (0 until k hits were seen)
    uint32_t* lock;
    unsigned long long* num_hits;
    // work queue
    unsigned long long* work_counter;
    unsigned long long n_items;
    // step seam: every hit of requests with emit_all
    unsigned long long* emit;
    unsigned long long* emit_count;
    unsigned long long emit_capacity;
};
// Returns the dynamic shared memory the launch needs (0 = cannot run with these parameters).
size_t tile_kernel_smem(uint32_t tile_log2, uint32_t max_leaves);
void launch_tile_eval(cudaStream_t st, const TileArgs& a, int n_sms);

// ---- planes.cu ----
// Index build: presence bits, f16 scores (score_row == nullptr: a mid plane, bits only) and the largest weight of one
// plane from its posting list (*bad is set when the list cannot be represented: unsorted or out-of-range anchors, or
// a weight that is not f16 / 100).
void launch_plane_fill(cudaStream_t st, const Posting* post, uint64_t n, uint32_t* bits_row, uint16_t* score_row, float* wmax_slot, uint32_t* bad, uint32_t anchor_lo, uint32_t span);
// Index build: anchors per (plane, tile of 2^13 anchors).
void launch_plane_tile_counts(cudaStream_t st, const uint32_t* bits, uint32_t n_planes, uint32_t words, uint32_t* tcount, uint32_t* tprefix);
// Index build: the kBoostLevels nested bitmaps of a boost column for the shard's anchors.
void launch_level_fill(cudaStream_t st, const uint32_t* col, uint32_t col_n, uint32_t anchor_lo, uint32_t span, const float* thr, uint32_t* bits, uint32_t words);

// One thread per request: the request's FastDesc (flags == 0 when it has to take the general path).
void launch_build_fast_desc(cudaStream_t st, const QueryProgram* queries, uint32_t n, const uint32_t* leaf_part, const PartPlanes* part_planes, const float* wmax, FastDesc* out);

struct PlaneArgs {
    const FastItem* items;                // plane-path items, group-major
    const uint32_t* group_item_begin;     // [n_groups + 1]
    const FastDesc* fast;
    const SparseEntry* sparse;
    PlaneSetView planes;
    uint32_t anchor_lo, anchor_hi;
    uint32_t group_tiles;                 // tiles of 2^13 anchors per item
    uint32_t group_begin, group_end;      // this launch evaluates the items of groups [group_begin, group_end)
    const uint32_t* ones_row;             // constant rows of at least group_tiles * 256 words: all ones, all zeros
    const uint32_t* zeros_row;
    uint32_t item_batch;                  // items a warp takes from the work counter at a time
    uint32_t seed_sweep_words;            // 32-anchor words of a seed row the seed pass looks at (a multiple of 128)
    // per-query state (shared with tile_eval_kernel)
    unsigned long long* heap;
    uint32_t heap_stride;
    unsigned long long* tau;
    uint32_t* lock;
    unsigned long long* num_hits;
    unsigned long long* work_counter;     // zeroed before the launch
    unsigned long long* stats;  // [0] items, [1] anchors evaluated exactly, [5] items swept before their threshold converged, [6] items answered from tile counts
};
size_t plane_kernel_smem();
void launch_plane_eval(cudaStream_t st, const PlaneArgs& a, int n_sms);
// Index build: the planes restricted to a boost column's seed set (ColumnLevels::seed_bits).
void launch_seed_compact(cudaStream_t st, const uint32_t* plane_bits, uint32_t n_planes, uint32_t words, const uint32_t* seed_anchor, uint32_t seed_n, uint32_t seed_words, uint32_t* out);
// Threshold seeds: one warp per plane-path request sweeps the seed rows of its planes over the whole shard, scores the
// anchors that have all its plane parts (plane contributions only: a lower bound of their score) and raises the
// request's threshold to just below the k-th best of them.  Uses PlaneArgs' fast / planes / anchor_lo / tau.
void launch_plane_seed(cudaStream_t st, const PlaneArgs& a, uint32_t n_queries, int n_sms);

// Top groups of every facet histogram: block f writes the `top[f]` largest counts of facets[f] (count desc, value id
// asc; zero counts never) to out_ids / out_counts [f * stride ...] and the number written to out_n[f].
void launch_facet_topk(cudaStream_t st, const FacetStep* facets, const uint32_t* top, uint32_t n_facets, uint32_t stride, uint32_t* out_ids, uint32_t* out_counts, uint32_t* out_n);

// Final ordering of each request's heap: merges `n_src` gathered heaps per query
// (n_src = 1: the local one) into `out_keys` [n_queries][stride] sorted by key desc,
// and sums the per-source num_hits.
void launch_merge_heaps(cudaStream_t st, const uint64_t* src_keys, const uint64_t* src_hits, uint32_t n_src, uint32_t n_queries, uint32_t stride, const QueryProgram* queries,
                        uint64_t* out_keys, uint64_t* out_hits);

}  // namespace vdev
