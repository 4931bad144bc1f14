/* veloci_b200 -- C ABI of the B200-native query-time hit pipeline for Veloci indices.
 *
 * This is the drop-in boundary: a Rust `-sys` crate binds exactly these symbols
 * (see INTEGRATION.md).  Plain pointers and sizes only; no C++/torch types.
 * Every entry point cites the reference interface it replaces (paths relative to
 * the veloci repository).
 *
 * Conventions
 *   - every function returns an int32 status (VGPU_OK == 0); the message of the last
 *     failure on the calling thread is read with vgpu_last_error().  Status codes
 *     mirror `VelociError` (src/error.rs:5-43).
 *   - handles are opaque.  The caller owns inputs; the library owns outputs until
 *     the matching *_free / *_close.
 *   - an index handle is immutable after open (like `&Persistence`, Sync); batches
 *     carry all mutable state, one batch per host thread at a time.
 *   - there is no CPU fallback: without a usable CUDA device every compute entry
 *     point fails with VGPU_ERR_CUDA.
 */
#ifndef VELOCI_B200_H
#define VELOCI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    VGPU_OK = 0,
    VGPU_ERR_INVALID_REQUEST = 1, /* VelociError::InvalidRequest (src/search.rs:151-155) */
    VGPU_ERR_FIELD_NOT_FOUND = 2, /* VelociError::FstNotFound: "field does not exist {path} (fst not found)" */
    VGPU_ERR_PATH_NOT_FOUND = 3,  /* VelociError::StringError("Did not found path in indices ...") (src/persistence.rs:454-458) */
    VGPU_ERR_IO = 4,              /* VelociError::Io / metaData.json problems */
    VGPU_ERR_JSON = 5,            /* VelociError::JsonError (serde) */
    VGPU_ERR_CUDA = 6,
    VGPU_ERR_NCCL = 7,
    VGPU_ERR_UNSUPPORTED = 8,     /* valid veloci request outside the accelerated path (see DESIGN.md "Known limits") */
    VGPU_ERR_INTERNAL = 9
};

typedef struct vgpu_index vgpu_index;
typedef struct vgpu_batch vgpu_batch;

/* Hit{id: u32, score: f32} (src/search.rs:53-57) */
typedef struct vgpu_hit {
    uint32_t id;
    float score;
} vgpu_hit;

/* Message of the last failed call on this thread (never NULL). */
const char* vgpu_last_error(void);

/* Number of CUDA devices visible to the library (0 when none / no driver). */
int32_t vgpu_device_count(void);

/* ---- index ----------------------------------------------------------------
 * Persistence::load(dir) (src/persistence.rs:393-409, 206-305): reads metaData.json,
 * every index file it lists and every FST, then builds the device layout in the HBM
 * of `device`.  With n_shards > 1 the handle owns the anchor range
 * [rank*num_docs/n_shards, (rank+1)*num_docs/n_shards): postings, anchor-keyed
 * stores and boost columns are cut to it; term-keyed structures are replicated. */
int32_t vgpu_index_open(const char* dir, int32_t device, uint32_t shard_rank, uint32_t n_shards, vgpu_index** out);
/* The same with build options (diagnostics and A/B tests; results never depend on them): without head-term planes every
 * request takes the posting path, without the deletion-neighbourhood index every fuzzy part scans the dictionary. */
enum { VGPU_OPEN_NO_PLANES = 1, VGPU_OPEN_NO_DELETION_INDEX = 2 };
int32_t vgpu_index_open_ex(const char* dir, int32_t device, uint32_t shard_rank, uint32_t n_shards, uint32_t flags, vgpu_index** out);
void vgpu_index_close(vgpu_index* idx);
/* num_docs of metaData.json and the anchor range this handle owns. */
int32_t vgpu_index_info(const vgpu_index* idx, uint64_t* num_docs, uint64_t* anchor_lo, uint64_t* anchor_hi, uint64_t* device_bytes);

/* ---- whole-query seam (batched) -------------------------------------------
 * search::search(request, &persistence) (src/search.rs:143-228) for `n` requests at
 * once.  `request_json[i]` is the reference's `search::Request` JSON, verbatim.
 *
 *   prepare : parse + plan (plan_creator, src/plan_creator/execution_plan.rs:132) on
 *             the host and upload the batch programs            (host -> device)
 *   execute : run every plan step of every request as batched kernels; results stay
 *             on the device; returns after the stream is idle
 *   fetch_* : copy results to the host                           (device -> host)
 * A request that fails (invalid, unknown field, unsupported) gets its status set and
 * an empty result; the other requests of the batch are unaffected. */
int32_t vgpu_batch_prepare(vgpu_index* idx, const char* const* request_json, uint32_t n, vgpu_batch** out);
/* The same with the requests as JSON lines in one buffer (one `search::Request` per line; empty lines are skipped):
 * what a server hands over when it drains its request queue.  *n_requests receives the number of requests found. */
int32_t vgpu_batch_prepare_jsonl(vgpu_index* idx, const char* text, size_t len, uint32_t* n_requests, vgpu_batch** out);
/* The same for exactly `n` requests joined by single line feeds (an empty line is an empty request, which fails like one):
 * what a binding passes instead of n C strings.  VGPU_ERR_INVALID_REQUEST, before any work, when the buffer does not hold
 * n - 1 line feeds (some request contains one: the caller then uses vgpu_batch_prepare). */
int32_t vgpu_batch_prepare_lines(vgpu_index* idx, const char* text, size_t len, uint32_t n, vgpu_batch** out);
int32_t vgpu_batch_execute(vgpu_batch* batch);
void vgpu_batch_free(vgpu_batch* batch);

/* Number of requests of the batch (a batch made from a plan never saw the request strings). */
int32_t vgpu_batch_size(const vgpu_batch* batch, uint32_t* n);
/* Per-request status (VGPU_OK or the error of that request) and message. */
int32_t vgpu_batch_status(const vgpu_batch* batch, uint32_t q);
const char* vgpu_batch_message(const vgpu_batch* batch, uint32_t q);

/* SearchResult.num_hits and .data (src/search/result/search_result.rs:8-26) of request
 * q after top/skip: hits sorted by (score desc, id desc) (src/search.rs:123-130).
 * `hits` may be NULL to query the count only; at most `cap` hits are written. */
int32_t vgpu_batch_result(const vgpu_batch* batch, uint32_t q, uint64_t* num_hits, vgpu_hit* hits, uint32_t cap, uint32_t* n_hits);

/* All requests at once into row-major [n][k] arrays (rows padded with id 0xFFFFFFFF,
 * score 0).  Any output pointer may be NULL. */
int32_t vgpu_batch_results_flat(const vgpu_batch* batch, uint32_t k, uint32_t* ids, float* scores, uint64_t* num_hits, int32_t* status);

/* facets of request q (src/facet.rs:31-73): number of facet fields, then per field
 * the (value id, count) groups sorted by count desc (ties: value id asc) and the
 * group's text (FST ord_to_term, src/search/search_field.rs:36-51). */
int32_t vgpu_batch_facet_count(const vgpu_batch* batch, uint32_t q, uint32_t* n_fields);
int32_t vgpu_batch_facet(const vgpu_batch* batch, uint32_t q, uint32_t field, const char** field_name, uint32_t* n_groups);
int32_t vgpu_batch_facet_group(const vgpu_batch* batch, uint32_t q, uint32_t field, uint32_t group, uint32_t* value_id, uint32_t* count, const char** text);

/* One-call form: prepare + execute + results_flat + free.  The e2e path of bench.py. */
int32_t vgpu_search_batch(vgpu_index* idx, const char* const* request_json, uint32_t n, uint32_t k, uint32_t* ids, float* scores, uint64_t* num_hits, int32_t* status);

/* ---- multi-GPU (anchor-range shards) ----------------------------
 * One process per GPU.  Process r opens shard r of n (vgpu_index_open with shard_rank = r, n_shards = n), one process
 * calls vgpu_comm_unique_id and hands the 128 bytes to the others (any channel: it is an ncclUniqueId), every process
 * calls vgpu_comm_init.  From then on vgpu_batch_execute on that handle is a collective: every rank executes the same
 * batch (same requests, same order) and, on the batch's stream with no host synchronisation in between,
 *   seed pass -> ncclAllReduce(max) of the per-request thresholds -> bulk pass -> local top-k
 *   -> ncclAllGather of the top-k rows and hit counts (+ ncclAllReduce(sum) of facet histograms) -> final merge,
 * so every rank ends with the complete SearchResult of every request (NCCL over NVLink / NVSwitch; failures are
 * VGPU_ERR_NCCL).  NCCL is bound at run time (libnccl.so.2); without it these calls fail and nothing else is affected. */
#define VGPU_COMM_ID_BYTES 128
int32_t vgpu_comm_unique_id(uint8_t id[VGPU_COMM_ID_BYTES]);
int32_t vgpu_comm_init(vgpu_index* idx, const uint8_t id[VGPU_COMM_ID_BYTES]);
int32_t vgpu_comm_destroy(vgpu_index* idx);

/* Plan once per box.  Parsing and planning a batch is host work that does not depend on the shard: one process prepares
 * the batch and exports its plan (a byte blob without process-local addresses, released with vgpu_free), the other
 * processes of the box import it against their own handle of the same index directory instead of planning again
 * (replaces N x plan_creator, src/plan_creator/execution_plan.rs:132-200, by 1 x).  VGPU_ERR_INVALID_REQUEST when the
 * blob was made for another index. */
int32_t vgpu_batch_export_plan(const vgpu_batch* batch, void** blob, size_t* len);
int32_t vgpu_batch_prepare_from_plan(vgpu_index* idx, const void* blob, size_t len, vgpu_batch** out);

/* The same through shared memory, for the processes of one box: a named channel (POSIX shared memory) that local rank 0
 * publishes plans into and the other local ranks read from.  vgpu_batch_prepare_shared is a collective over the
 * channel's ranks: rank 0 passes the requests, parses, plans and publishes; the others pass NULL / 0, wait for the plan
 * and import it.  Batches are numbered by tickets: every rank calls vgpu_plan_channel_ticket once per batch, in the order
 * the batches will be executed (1, 2, ...), and passes the ticket to the prepare; prepares of consecutive tickets may
 * then run concurrently on different threads (two plans can be in the making or in flight at once).  `capacity`
 * bounds the blob size. */
typedef struct vgpu_plan_channel vgpu_plan_channel;
int32_t vgpu_plan_channel_open(const char* name, uint32_t local_rank, uint32_t local_ranks, size_t capacity, vgpu_plan_channel** out);
void vgpu_plan_channel_close(vgpu_plan_channel* ch);
uint64_t vgpu_plan_channel_ticket(vgpu_plan_channel* ch);
int32_t vgpu_batch_prepare_shared(vgpu_index* idx, vgpu_plan_channel* ch, uint64_t ticket, const char* text, size_t len, uint32_t n, vgpu_batch** out);

/* ---- multi-GPU merge driven by the host (the pieces vgpu_batch_execute runs by itself once vgpu_comm_init was called;
 * kept for hosts with their own transport and for single-device tests of several shards) ----------------------
 * After execute, every shard holds the local top-(top+skip) of each request as
 * device rows of `stride` 64-bit keys ((orderable score << 32) | anchor id, 0 = no
 * hit) plus its local num_hits.  The host all-gathers those buffers (NCCL) and calls
 * vgpu_batch_merge_gathered on every rank: final top-k of n_shards * stride
 * candidates, num_hits summed.  Pointers are device pointers. */
int32_t vgpu_batch_local_topk(const vgpu_batch* batch, uint64_t** keys_dev, uint64_t** num_hits_dev, uint32_t* stride);
/* Optional threshold exchange.  vgpu_batch_execute == execute_begin + execute_finish.  After execute_begin every shard
 * has evaluated its first anchor tiles and holds, per request, the order key of its k-th best hit so far (0: fewer than k
 * hits) in the device array returned by vgpu_batch_thresholds.  The global k-th best is at least the largest of the
 * shards' values, so the host may all-reduce the arrays with MAX (unsigned 64-bit order) before execute_finish: every
 * shard then prunes against the shared threshold.  The merged result is unchanged (hits below it cannot reach the top k). */
int32_t vgpu_batch_execute_begin(vgpu_batch* batch);
int32_t vgpu_batch_thresholds(const vgpu_batch* batch, uint64_t** tau_dev, uint32_t* n);
int32_t vgpu_batch_execute_finish(vgpu_batch* batch);
int32_t vgpu_batch_merge_gathered(vgpu_batch* batch, const uint64_t* gathered_keys_dev, const uint64_t* gathered_num_hits_dev, uint32_t n_shards);
/* Facets on shards: after execute every shard holds the counts of its anchors in one device array of u32 (all facet
 * histograms of the batch, back to back; NULL / 0 when the batch has no facets).  The host all-reduces the arrays with SUM
 * before vgpu_batch_merge_gathered, which then picks the top groups of the summed histograms on every rank. */
int32_t vgpu_batch_facet_histograms(const vgpu_batch* batch, uint32_t** hist_dev, uint64_t* n);

/* ---- suggest ---------------------------------------------------------------
 * suggest_multi (src/search/search_field.rs:178-217): `request_json` is a search::Request with `suggest` (a list of
 * RequestSearchPart), `top` and `skip`.  Every part is matched and scored on the device (get_term_ids_in_field with
 * return_term), the parts' own top/skip bound their hits (:292-294,:322-331,:366-369), equal lower-cased texts merge
 * keeping the largest score, the list is ordered by score and cut by skip/top: SuggestFieldResult = Vec<(String, Score,
 * TermId)>.  Among equal scores the reference's order is whatever its unstable sort leaves; here it is descending text.
 * vgpu_suggest_part is search_field::suggest (:219-228): one part whose top/skip also cut the list.
 * `items[i].text` points into `text_block`; release with vgpu_suggestions_free. */
typedef struct vgpu_suggestion {
    const char* text;  /* lower-cased term, UTF-8, NUL-terminated */
    float score;
    uint32_t id;       /* term id in the part's field */
} vgpu_suggestion;
typedef struct vgpu_suggestions {
    vgpu_suggestion* items;
    uint32_t n;
    char* text_block;
} vgpu_suggestions;
int32_t vgpu_suggest(vgpu_index* idx, const char* request_json, vgpu_suggestions* out);
int32_t vgpu_suggest_part(vgpu_index* idx, const char* part_json, vgpu_suggestions* out);
/* search_field::highlight (src/search/search_field.rs:232-245): the texts of the part's field that it hits, the hit tokens
 * tagged (`snippet: true`; `snippet_info` = {num_words_around_snippet, snippet_start_tag, snippet_end_tag, snippet_connector,
 * max_snippets}, src/search/request/snippet_info.rs), best score first, the part's skip / top applied.  The part's terms are
 * normalised first (util::normalize_text, src/util.rs:11-30).  `out->items[i]`: text = the highlighted text, id = its text id.
 * Released with vgpu_suggestions_free. */
int32_t vgpu_highlight(vgpu_index* idx, const char* part_json, vgpu_suggestions* out);
void vgpu_suggestions_free(vgpu_suggestions* s);

/* ---- request generation (host side, no kernel) ------------------------------
 * query_generator::search_query (src/query_generator.rs:175-257) with the query language of the `query_parser` crate
 * (query_parser/src/lexer.rs:107-195, parser.rs:141-190): `params_json` is a SearchQueryGeneratorParameters
 * (src/query_generator.rs:44-83: search_term, parser_options, top, skip, ignore_case, levenshtein,
 * levenshtein_auto_limit, facetlimit, why_found, text_locality, boost_queries, facets, fields, boost_fields, boost_terms,
 * phrase_pairs, explain, filter, filter_parser_options); `*request_json` receives the search::Request as the reference's
 * serde derive would write it (NUL-terminated, released with vgpu_free), ready for vgpu_batch_prepare.  Field fan-out
 * (query_parser_to_veloci_request.rs:85-114), automatic edit distance (src/query_generator.rs:85-99,129-132), "term*"
 * prefix and "a*b" regex parts (query_parser_to_veloci_request.rs:44-65), phrase pairs (src/query_generator.rs:268-295)
 * and the simplification of nested or/and (src/search/request/search_request.rs:27-76) are the reference's.  Errors: a
 * query that does not parse -> VGPU_ERR_INVALID_REQUEST with the ParseError text; FieldNotFound -> VGPU_ERR_FIELD_NOT_FOUND;
 * AllFieldsFiltered -> VGPU_ERR_INVALID_REQUEST; parameters that do not deserialize -> VGPU_ERR_JSON.
 * vgpu_suggest_query is suggest_query (src/query_generator.rs:297-322); `params_json` = {"request": text, "top", "skip",
 * "levenshtein", "fields", "levenshtein_auto_limit"}; the result goes to vgpu_suggest.
 * vgpu_query_parse is query_parser::parse_with_opt alone: the Debug text of the tree (query_parser/src/ast.rs:51-59);
 * `options` bit 0 no_attributes, bit 1 no_parentheses, bit 2 no_levensthein (query_parser/src/lib.rs:43-54). */
int32_t vgpu_search_query(vgpu_index* idx, const char* params_json, char** request_json);
int32_t vgpu_suggest_query(vgpu_index* idx, const char* params_json, char** request_json);
int32_t vgpu_query_parse(const char* text, uint32_t options, char** tree_debug);

/* ---- documents (host side, no kernel) ----------------------------------------
 * DocLoader::get_doc (doc_store/src/lib.rs:26-62) over the index directory's compressed document store `data`
 * (lz4 blocks + block index, written by src/create/write_docs.rs:11-34): the stored JSON text of document `doc_id`
 * (= the anchor id of a hit), NUL-terminated, released with vgpu_free.  What search::to_documents does per hit when the
 * request has no `select` (src/search.rs:89-98).  VGPU_ERR_IO when the store is missing, damaged, or has no such document. */
int32_t vgpu_get_doc(vgpu_index* idx, uint32_t doc_id, char** doc_json);
/* search::to_search_result (src/search.rs:65-110): after execute, the hits of request q
 * as documents, `{"num_hits": n, "data": [{"doc": {..}, "hit": {"id", "score"}, "why_found": {"<field>": ["..<b>term</b>.."]}}]}`.
 * `why_found` is filled when the request set `"why_found": true`: the texts of the terms its parts matched
 * (SearchResult::why_found_terms, src/search.rs:186) highlight the stored document, field by field
 * (highlight_on_original_document / highlight_text, src/highlight_field.rs:98-186: windows of five words around the hits,
 * " ... " between and around them).  A request with `select` gets its documents from vgpu_read_doc's path instead of the
 * store.  Released with vgpu_free.  Not for sharded batches (a shard holds its own anchors only):
 * fetch the merged hits' documents with vgpu_get_doc there. */
int32_t vgpu_batch_result_docs(vgpu_batch* batch, uint32_t q, char** result_json);
/* SearchResult::explain of the hits request q returns (src/search.rs:174, read per hit by to_documents :86,96;
 * src/search/result/explain.rs:1-21): `{"<anchor id>": [Explain, ...]}` in serde's form of the enum --
 * {"LevenshteinScore": {"score", "text_or_token_id", "term_id"}} (search_field.rs:334-344),
 * {"TermToAnchor": {"term_score", "anchor_score", "final_score", "term_id"}} (:429-441),
 * {"OrSumOverDistinctTerms": f32} (set_op.rs:187-190), {"Boost": f32} (boost.rs:297-300,371-374) -- in the order the
 * reference's plan steps push them.  For requests with `"explain": true` (or a part with `"options": {"explain": true}`);
 * vgpu_batch_result_docs carries the same lists as each hit's "explain".  VGPU_ERR_UNSUPPORTED for the explanation (never for
 * the search) when the request has phrase boosts or 1:n boosts, on sharded handles and on imported plans.  Released with vgpu_free. */
int32_t vgpu_batch_explain(vgpu_batch* batch, uint32_t q, char** explain_json);
/* search::explain_plan (src/search.rs:132-141): the request's execution plan as a Graphviz dot graph -- the steps
 * plan_creator lays out (src/plan_creator/execution_plan.rs:132-200), labelled like the reference's steps
 * (plan_steps.rs:76-135: "search <path> <term>", "token to anchor", "Union", "BoostPlanStepFromBoostRequest", ...), one
 * edge per dependency.  Needs no index.  Released with vgpu_free. */
int32_t vgpu_explain_plan(const char* request_json, char** dot);
/* read_data (src/search/read_document.rs:8-59): document `doc_id` rebuilt from the indices, only the fields of `fields_json`
 * (a JSON list of field paths, the request's `select`): 1:n levels through `<level>.parent_to_value_id`, texts through
 * `<field>.textindex.parent_to_value_id` and the dictionary, long texts from their token ids (src/search.rs:242-269).
 * vgpu_batch_result_docs uses it for requests with `select`, together with the token-id based why_found
 * (src/search/why_found.rs:11-49, highlight_document src/highlight_field.rs:187-271). */
int32_t vgpu_read_doc(vgpu_index* idx, uint32_t doc_id, const char* fields_json, char** doc_json);

/* ---- step seam -------------------------------------------------------------
 * One symbol per PlanStep kind (src/plan_creator/plan_steps.rs:18-74), each over host
 * hit lists; used by the step-level parity tests.  Outputs are malloc'd by the
 * library and released with vgpu_free. */
typedef struct vgpu_hitlist {
    vgpu_hit* hits;   /* hits_scores */
    uint32_t n_hits;
    uint32_t* ids;    /* hits_ids */
    uint32_t n_ids;
} vgpu_hitlist;
void vgpu_free(void* p);
void vgpu_hitlist_free(vgpu_hitlist* l);

/* get_term_ids_in_field (src/search/search_field.rs:277-398): `part_json` is one
 * RequestSearchPart; fills hits_scores (term id, score) and/or hits_ids. */
int32_t vgpu_field_search(vgpu_index* idx, const char* part_json, int32_t get_scores, int32_t get_ids, vgpu_hitlist* out);
/* resolve_token_to_anchor (src/search/search_field.rs:400-504) for the term hits of `in`. */
int32_t vgpu_resolve_to_anchor(vgpu_index* idx, const char* part_json, const vgpu_hitlist* in, vgpu_hitlist* out);
/* union_hits_score (src/search/set_op.rs:87-220); terms[i] = request.terms[0] of input i. */
int32_t vgpu_union_hits_score(vgpu_index* idx, const vgpu_hitlist* inputs, const char* const* terms, uint32_t n, vgpu_hitlist* out);
/* intersect_hits_score (src/search/set_op.rs:368-446). */
int32_t vgpu_intersect_hits_score(vgpu_index* idx, const vgpu_hitlist* inputs, uint32_t n, vgpu_hitlist* out);
/* resolve_token_to_anchor with the FilterResult of the request's filter tree (ResolveTokenIdToAnchor with a filter channel,
 * src/plan_creator/plan_steps.rs:98-131; search_field.rs:423, 540-548): anchors outside `filter_ids` are skipped. */
int32_t vgpu_resolve_to_anchor_filtered(vgpu_index* idx, const char* part_json, const vgpu_hitlist* in, const uint32_t* filter_ids, uint32_t n_filter, vgpu_hitlist* out);
/* union_hits_ids (src/search/set_op.rs:222-258) and intersect_hits_ids (:468-510) over the inputs' hits_ids (the Union /
 * Intersect steps of an ids-only filter sub-plan, plan_steps.rs:295-328). */
int32_t vgpu_union_hits_ids(vgpu_index* idx, const vgpu_hitlist* inputs, uint32_t n, vgpu_hitlist* out);
int32_t vgpu_intersect_hits_ids(vgpu_index* idx, const vgpu_hitlist* inputs, uint32_t n, vgpu_hitlist* out);
/* IntersectScoresWithIds (plan_steps.rs:330-345; intersect_score_hits_with_ids, set_op.rs:311-326): the scored hits of
 * `scores` whose id is among `ids->ids`. */
int32_t vgpu_intersect_scores_with_ids(vgpu_index* idx, const vgpu_hitlist* scores, const vgpu_hitlist* ids, vgpu_hitlist* out);
/* PlanStepPhrasePairToAnchorId (src/plan_creator/plan_steps.rs:279-293) = get_anchor_for_phrases_in_field
 * (src/search/search_field.rs:263-275): the anchors of every (id of `ids1`, id of `ids2`) term pair in the phrase-pair store of
 * `path` ("<field>" or "<field>.textindex[.phrase_pair_to_anchor]"), sorted, duplicates kept, in out->ids. */
int32_t vgpu_phrase_pairs_to_anchor(vgpu_index* idx, const char* path, const uint32_t* ids1, uint32_t n1, const uint32_t* ids2, uint32_t n2, vgpu_hitlist* out);
/* BoostAnchorFromPhraseResults (src/plan_creator/plan_steps.rs:260-277): `phrase_results[i].ids` are the anchors one phrase
 * step produced, `group[i]` numbers the phrase (the (search1.terms[0], search2.terms[0]) pair) it belongs to; the results of
 * one phrase are merged and deduplicated (:230-257), every phrase multiplies the hits it contains by 5.0
 * (boost_hits_ids_vec_multi, src/search/boost.rs:149-195).  out->hits: the hits by ascending id. */
int32_t vgpu_boost_anchor_from_phrase_results(vgpu_index* idx, const vgpu_hitlist* hits, const vgpu_hitlist* phrase_results, const uint32_t* group, uint32_t n, vgpu_hitlist* out);
/* BoostToAnchor (src/plan_creator/plan_steps.rs:174-196): `part_json` is the search part, `in` its term hits (hits: the term
 * ids of a tokenized field; ids: the text ids of an untokenized one), `boost_json` the RequestBoostPart on the part's own 1:n
 * level.  out->hits = SearchFieldResult::boost_ids: (anchor, boost value) in value-id order. */
int32_t vgpu_boost_to_anchor(vgpu_index* idx, const char* part_json, const vgpu_hitlist* in, const char* boost_json, vgpu_hitlist* out);
/* ApplyAnchorBoost (src/plan_creator/plan_steps.rs:198-217) = apply_boost_values_anchor (src/search/boost.rs:255-281): `hits`
 * by ascending anchor, `boost_ids->hits` the (anchor, value) list vgpu_boost_to_anchor returned for the same part.  out->hits:
 * the hits with the boost function / expression of `boost_json` applied the way the reference's merge walk applies it. */
int32_t vgpu_apply_anchor_boost(vgpu_index* idx, const char* boost_json, const vgpu_hitlist* hits, const vgpu_hitlist* boost_ids, vgpu_hitlist* out);
/* boost_text_locality (src/search/boost.rs:34-87) of one field, reduced per anchor like boost_text_locality_all (:11-32):
 * `term_hits[t].ids` = the token ids query term t matched in `path`; out->hits = (anchor, boost 2 c c) by ascending anchor for
 * the texts reached by c > 1 of those tokens (the smallest boost per anchor, as the reference keeps it). */
int32_t vgpu_text_locality(vgpu_index* idx, const char* path, const vgpu_hitlist* term_hits, uint32_t n_terms, vgpu_hitlist* out);
/* get_facet (src/facet.rs:31-73): `facet_json` is one FacetRequest, `ids` the hit ids; groups come back through the
 * suggestion list type: text = the group's text, id = its value id, score = its count (exact below 2^24). */
int32_t vgpu_facet(vgpu_index* idx, const char* facet_json, const uint32_t* ids, uint32_t n_ids, vgpu_suggestions* out);
/* add_boost (src/search/boost.rs:470-504): `boost_json` is one RequestBoostPart. */
int32_t vgpu_add_boost(vgpu_index* idx, const char* boost_json, vgpu_hitlist* inout);
/* top_n_sort + apply_top_skip (src/search/sort.rs:5-22, src/search.rs:230-239). */
int32_t vgpu_top_n(vgpu_index* idx, const vgpu_hitlist* in, uint32_t top, uint32_t skip, vgpu_hitlist* out);

/* ---- step seam, device-resident -------------------------------------------
 * The same steps over hit lists that stay in HBM between them ("device-resident handles between steps"; the
 * reference passes owned SearchFieldResults from step to step over channels, src/plan_creator/plan_steps.rs:357-376).
 * A vgpu_hitlist_dev holds the hits_scores of a SearchFieldResult on the index's device, by ascending anchor id; between
 * two steps only the handles and the lists' lengths cross the bus.  ResolveTokenIdToAnchor starts from the (few) term hits
 * of vgpu_field_search on the host; Union / Intersect / BoostPlanStepFromBoostRequest go from handles to a handle;
 * vgpu_dev_top_n brings the k best hits back.  Handles belong to the index they were made on (steps refuse foreign handles);
 * a handle may be freed after its index was closed, but not used;
 * a sharded handle's lists hold the shard's anchors.  Whole requests should still go through vgpu_batch_execute, which
 * fuses these steps into one pass per tile; the step symbols serve hosts that drive the reference's plan themselves. */
typedef struct vgpu_hitlist_dev vgpu_hitlist_dev;
int32_t vgpu_dev_upload(vgpu_index* idx, const vgpu_hitlist* in, vgpu_hitlist_dev** out);      /* hits of a host list (a repeated anchor keeps its largest score) */
int32_t vgpu_dev_download(const vgpu_hitlist_dev* list, vgpu_hitlist* out);                     /* by ascending anchor id; released with vgpu_hitlist_free */
uint32_t vgpu_dev_len(const vgpu_hitlist_dev* list);
void vgpu_dev_free(vgpu_hitlist_dev* list);
/* ResolveTokenIdToAnchor (src/search/search_field.rs:400-464; plan_steps.rs:155-167) */
int32_t vgpu_dev_resolve_to_anchor(vgpu_index* idx, const char* part_json, const vgpu_hitlist* term_hits, vgpu_hitlist_dev** out);
/* Union (src/search/set_op.rs:87-220): terms[i] = request.terms[0] of input i */
int32_t vgpu_dev_union_hits_score(vgpu_index* idx, const vgpu_hitlist_dev* const* inputs, const char* const* terms, uint32_t n, vgpu_hitlist_dev** out);
/* Intersect (src/search/set_op.rs:368-446) */
int32_t vgpu_dev_intersect_hits_score(vgpu_index* idx, const vgpu_hitlist_dev* const* inputs, uint32_t n, vgpu_hitlist_dev** out);
/* BoostPlanStepFromBoostRequest (add_boost, src/search/boost.rs:470-504) */
int32_t vgpu_dev_add_boost(vgpu_index* idx, const char* boost_json, const vgpu_hitlist_dev* in, vgpu_hitlist_dev** out);
/* top_n_sort + apply_top_skip (src/search/sort.rs:5-22, src/search.rs:230-239) */
int32_t vgpu_dev_top_n(vgpu_index* idx, const vgpu_hitlist_dev* in, uint32_t top, uint32_t skip, vgpu_hitlist* out);

/* ---- instrumentation ---------------------------------------------------------
 * Kernel launches issued by this library since process start, and per-phase device
 * time (CUDA events on the library's stream) of the last vgpu_batch_execute:
 * phase 0 fuzzy match, 1 match grouping + scoring, 2 posting slicing, 3 plane
 * evaluation and 4 tile evaluation (expand + merge + boost + tile top-k of the
 * (tile, request) items on the head-term plane path / the general posting path),
 * 5 final top-k.  Also the algorithmic byte counts of the last execute
 *  and how many items took which path. */
uint64_t vgpu_launch_count(void);
int32_t vgpu_batch_phase_ms(const vgpu_batch* batch, float* ms, uint32_t n_phases);
int32_t vgpu_batch_traffic_model(const vgpu_batch* batch, uint64_t* posting_bytes, uint64_t* boost_bytes, uint64_t* postings, uint64_t* union_hits);
int32_t vgpu_batch_path_stats(const vgpu_batch* batch, uint64_t* plane_items, uint64_t* general_items, uint64_t* plane_evaluated);
/* Per-kernel device time: with profiling on, the next execute brackets every kernel launch with a pair of CUDA events on
 * the batch's stream (a few microseconds each: not for timed runs); vgpu_batch_kernel_times_json returns
 * {"kernel": {"launches": n, "ms": total}, ...} of that execute (owned by the batch). */
int32_t vgpu_batch_set_profiling(vgpu_batch* batch, int32_t on);
const char* vgpu_batch_kernel_times_json(vgpu_batch* batch);
/* Work counters of the last execute, in this order: matched terms, postings of the matched terms (posting-list model),
 * union hits, sparse entries (postings of non-plane terms copied into tile buckets), plane-path items, general items,
 * anchors evaluated exactly on the plane path, plane-path item evaluations (seed + bulk), of which swept before their
 * threshold converged, of which answered from per-plane tile counts, anchor tiles, distinct search parts, planes of the
 * index, 32-anchor words per plane. */
#define VGPU_WORK_STATS 14
int32_t vgpu_batch_work_stats(const vgpu_batch* batch, uint64_t* out, uint32_t n);
/* Bytes copied host->device by prepare (plan tables) and device->host by execute + fetch. */
int32_t vgpu_batch_io_bytes(const vgpu_batch* batch, uint64_t* h2d, uint64_t* d2h);

#ifdef __cplusplus
}
#endif
#endif /* VELOCI_B200_H */
