// Batched executor: runs every plan step of every request of a batch as a fixed
// sequence of kernels (the reference runs one rayon task per step per request,
// src/plan_creator/execution_plan.rs:538-546).
//
//   phase 0  fuzzy_match per dictionary              (PlanStepFieldSearchToTokenIds)
//   phase 1  group matches by part, score them        (  "  , hits_scores of each part)
//   phase 2  dense tile offsets + sparse tile buckets (ResolveTokenIdToAnchor, slicing only)
//   phase 3  plane evaluation  } ResolveTokenIdToAnchor, Union, Intersect, BoostPlanStepFromBoostRequest,
//   phase 4  tile evaluation   } top_n_sort: (tile, request) items on the plane path / the general path
//   phase 5  heap finalisation                        (apply_top_skip happens on the host view)
#pragma once
#include <chrono>
#include <cstdlib>
#include <thread>

#include "../cuda/bitvec.cuh"
#include "../cuda/kernels.cuh"
#include "plan_blob.hpp"
#include "planner.hpp"

namespace vdev {

struct vgpu_hit_pod {  // same layout as vgpu_hit
    uint32_t id;
    float score;
};

struct ExplicitList {  // a caller-provided hit list used as a leaf (step seam)
    std::vector<uint32_t> anchors;
    std::vector<float> scores;
    std::vector<uint32_t> raw_keys;  // when given: the entries' keys as they are (a 1:n boost list: 0x7FFFFFFF - value id) instead of score keys
    uint32_t part_flags = 0;         // PartFlags of the leaf's part (kPartList | kPartListBoost for such a list)
    // a list that is resident on the device (DeviceHitList): anchor-sorted SparseEntry array; the host vectors stay empty
    const SparseEntry* dev = nullptr;
    uint32_t dev_n = 0;
    bool dev_nonneg = true;
};

// Hits of a plan step kept on the device for the next step (vgpu_hitlist_dev): anchor-sorted (anchor, score key) entries,
// the form the tile path reads sparse postings in.
struct DeviceHitList {
    DeviceIndex* ix = nullptr;
    DevBuf<unsigned long long> entries;  // SparseEntry {anchor, key}, anchors ascending and unique
    uint32_t n = 0;
    bool nonneg = true;  // no entry has a negative score (cheap key decode downstream)
    const SparseEntry* data() const { return reinterpret_cast<const SparseEntry*>(entries.p); }
};

static const int kPhases = 6;

// Timing experiments read their settings from the environment only in a library built with -DVELOCI_PROBES
// (tools/*_probe.py); the release library never consults them (one of them, VELOCI_TILE_LIMIT, gives wrong results).
inline const char* probe_env(const char* name) {
#ifdef VELOCI_PROBES
    return getenv(name);
#else
    (void)name;
    return nullptr;
#endif
}

// Set (per thread) around a prepare whose plan is published to the other ranks of the box.
inline bool& plan_for_all_ranks() {
    static thread_local bool flag = false;
    return flag;
}

struct Batch {
    DeviceIndex* ix = nullptr;
    int device = -1;  // ix->device, kept for the batch's destruction
    vplan::BatchPlan plan;
    uint32_t n = 0, n_parts = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[kPhases + 1] = {};
    int n_sms = 148;
    enum Mode { kRequests, kTermHits, kLists } mode = kRequests;

    // device: plan
    DevBuf<PartQuery> d_parts;
    DevBuf<uint32_t> d_part_dict, d_leaf_part, d_prog;
    DevBuf<QueryProgram> d_programs;
    DevBuf<BoostStep> d_boosts;
    DevBuf<DictView> d_dicts;
    DevBuf<PostingsView> d_postings;
    std::vector<std::vector<uint32_t>> parts_of_dict;   // parts the dictionary scan matches (starts_with, raw case, distance > 2)
    std::vector<DevBuf<uint32_t>> d_parts_of_dict;
    std::vector<std::vector<uint32_t>> probe_of_dict;   // parts matched through the deletion-neighbourhood index
    std::vector<DevBuf<uint32_t>> d_probe_of_dict, d_overflow_of_dict;
    DevBuf<unsigned long long> d_overflow_count;
    std::vector<uint32_t> max_m_of_dict;
    std::vector<std::vector<RegexPartDev>> regex_of_dict;  // is_regex parts: their DFAs over each dictionary's alphabet codes
    std::vector<DevBuf<RegexPartDev>> d_regex_of_dict;
    DevBuf<uint16_t> d_regex_tables;                       // class maps and transition tables of all of them, back to back
    // device: match phase
    DevBuf<MatchRecord> d_records;
    DevBuf<unsigned long long> d_counters;  // [0] matches, [1] work counter, [2] stat postings, [3] dense rows, [4] emitted hits
    DevBuf<uint32_t> d_part_count, d_part_begin, d_dense_cursor, d_sparse_cursor, d_row_match;
    DevBuf<unsigned long long> d_part_est;
    DevBuf<uint32_t> d_g_term, d_g_df, d_g_row, d_g_part;
    DevBuf<float> d_g_score;
    DevBuf<uint64_t> d_g_begin;
    DevBuf<uint32_t> d_inj_terms;
    DevBuf<float> d_inj_scores;
    uint32_t n_injected = 0;
    // device: slicing
    DevBuf<uint32_t> d_toff, d_bucket;
    DevBuf<SparseEntry> d_sparse;
    DevBuf<uint64_t> d_sparse_total, d_sparse_base;
    DevBuf<PartSlices> d_slices;
    DevBuf<ItemRec> d_items;
    DevBuf<SliceRec> d_slice_recs;
    // device: plane path
    bool use_planes = false;
    DevBuf<PartPlanes> d_part_planes;
    DevBuf<uint32_t> d_g_plane;
    DevBuf<FastDesc> d_fast;
    DevBuf<uint32_t> d_fast_item_cursor, d_fast_item_begin;
    DevBuf<FastItem> d_fast_items;
    // device: list producers
    DevBuf<PhraseMember> d_phrase_members;
    DevBuf<IdsMember> d_ids_members;
    DevBuf<BoostListMember> d_boost_members;
    DevBuf<TlInstance> d_tl_instances;
    DevBuf<uint32_t> d_tl_term_parts, d_req_error;
    std::vector<uint32_t> h_req_error;
    // device: facets
    DevBuf<FacetStep> d_facets;
    DevBuf<uint32_t> d_facet_top, d_facet_hist, d_facet_ids, d_facet_counts, d_facet_n;
    uint32_t n_facets = 0, facet_stride = 1;
    std::vector<uint32_t> h_facet_ids, h_facet_counts, h_facet_n;
    uint64_t stat_fast_items = 0, stat_general_items = 0, stat_plane_evaluated = 0;
    // between execute_begin and execute_finish
    unsigned long long pending_items = 0;
    uint32_t pending_fast_items = 0;
    bool begun = false;
    // device: per-request state and results
    DevBuf<unsigned long long> d_heap, d_tau, d_num_hits, d_merge_scratch;
    DevBuf<uint32_t> d_lock;
    DevBuf<uint64_t> d_out_keys, d_out_hits;
    DevBuf<unsigned long long> d_emit;  // all hits of request 0 (step seam): (score key << 32) | anchor
    uint64_t emit_capacity = 0;
    uint32_t stride = 1;
    // geometry
    uint32_t tile_log2 = 13, n_tiles = 0, group_tiles = 1, n_groups = 0;
    // host results
    bool matched = false, executed = false, fetched = false;
    std::vector<uint64_t> h_keys, h_hits;
    float phase_ms[kPhases] = {};
    uint64_t stat_postings = 0, stat_matches = 0, stat_union = 0, stat_sparse = 0;
    uint64_t stat_plane_item_evals = 0, stat_plane_unconverged = 0, stat_plane_sweepless = 0;
    uint64_t h2d_bytes = 0, d2h_bytes = 0;

    // Optional per-kernel timing (vgpu_batch_set_profiling): a pair of CUDA events around every launch of the next
    // execute, on the batch's stream; read back by kernel_times_json() after the step.
    bool profiling = false;
    struct KernelSpan {
        const char* name;
        cudaEvent_t a, b;
    };
    std::vector<KernelSpan> spans;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> span_pool;
    size_t spans_used = 0;
    template <class F>
    void timed(const char* name, F&& launch) {
        if (!profiling) return launch();
        if (spans_used == span_pool.size()) {
            cudaEvent_t a, b;
            VDEV_CUDA(cudaEventCreate(&a));
            VDEV_CUDA(cudaEventCreate(&b));
            span_pool.emplace_back(a, b);
        }
        const auto& ev2 = span_pool[spans_used++];
        const uint64_t before = launches_so_far();
        VDEV_CUDA(cudaEventRecord(ev2.first, stream));
        launch();
        if (launches_so_far() == before) {  // the wrapper had nothing to launch
            --spans_used;
            return;
        }
        VDEV_CUDA(cudaEventRecord(ev2.second, stream));
        spans.push_back(KernelSpan{name, ev2.first, ev2.second});
    }
    // {"kernel": {"launches": n, "ms": total}, ...} of the last execute with profiling on
    std::string kernel_times_json() {
        std::vector<std::pair<std::string, std::pair<uint32_t, double>>> agg;
        for (auto& sp : spans) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, sp.a, sp.b) != cudaSuccess) {
                cudaGetLastError();
                continue;
            }
            bool found = false;
            for (auto& kv : agg)
                if (kv.first == sp.name) kv.second.first += 1, kv.second.second += ms, found = true;
            if (!found) agg.push_back({sp.name, {1u, (double)ms}});
        }
        std::string out = "{";
        for (size_t i = 0; i < agg.size(); ++i) {
            char buf[160];
            snprintf(buf, sizeof buf, "%s\"%s\": {\"launches\": %u, \"ms\": %.6f}", i ? ", " : "", agg[i].first.c_str(), agg[i].second.first, agg[i].second.second);
            out += buf;
        }
        return out + "}";
    }

    Batch() = default;
    Batch(const Batch&) = delete;
    Batch& operator=(const Batch&) = delete;
    ~Batch() {
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        for (auto& pr : span_pool) cudaEventDestroy(pr.first), cudaEventDestroy(pr.second);
        if (stream) cudaStreamDestroy(stream);
    }

    // ------------------------------------------------------------ preparation
    // `upload` false: the plan tables stay on the host until upload_plan() (the caller publishes the plan first)
    void prepare(DeviceIndex* index, const char* const* request_json, uint32_t count, bool upload = true) {
        ix = index;
        n = count;
        plan.ix = ix;
        plan.requests.reserve(n);
        const auto t0 = std::chrono::steady_clock::now();
        auto t_joined = t0;
        // every thread parses and plans a contiguous chunk of the requests; the chunk plans are merged in request order
        // Up to 16 planner threads; when several ranks share the host (LOCAL_WORLD_SIZE, set by torchrun) each takes its
        // share of the cores but at least 8: at N = 4 on 32 cores, 8 threads per rank plan a batch in 10.5 ms where 16
        // (128 runnable threads with the ranks' other threads) take 12.4 ms and 4 cannot keep up with the GPU.
        // at most 16 threads and three quarters of the cores: the thread that launches the previous batch's kernels needs one
        // (N = 1 on 16 cores, end to end: 12 threads 656k requests/s, 16 threads 547k-631k; `VELOCI_PLAN_THREADS` overrides)
        unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency() * 3 / 4));
        if (plan_for_all_ranks()) {
            // This process plans for every rank of the box (vgpu_batch_prepare_shared): the other ranks' cores are its to use,
            // but two batches are planned at a time (Index.search_stream, plan channel tickets), so each takes half of them.
            // Measured at N = 8 on 32 cores: 2 x 16 threads 1.13 M requests/s end to end, 2 x 32 threads 0.78 M.
            hw = std::max(4u, std::min(16u, std::thread::hardware_concurrency() / 2));
        } else if (const char* env = getenv("LOCAL_WORLD_SIZE")) {
            const unsigned ranks = (unsigned)std::max(1, atoi(env));
            hw = std::min(hw, std::max(8u, std::thread::hardware_concurrency() / ranks));
        }
        if (const char* env = getenv("VELOCI_PLAN_THREADS")) hw = (unsigned)std::min(64, std::max(1, atoi(env)));
        if (n < 256) hw = 1;
        if (hw > 1) {
            std::vector<vplan::BatchPlan> chunks(hw);
            std::vector<std::thread> pool;
            const uint32_t chunk = (n + hw - 1) / hw;
            for (unsigned t = 0; t < hw; ++t) {
                const uint32_t a = t * chunk, b = std::min(n, a + chunk);
                if (a >= b) break;
                chunks[t].ix = ix;
                pool.emplace_back([&, t, a, b]() {
                    for (uint32_t i = a; i < b; ++i) chunks[t].add_request(request_json[i] ? request_json[i] : "");
                });
            }
            for (auto& th : pool) th.join();
            t_joined = std::chrono::steady_clock::now();
            plan.reserve_for(chunks);
            for (unsigned t = 0; t < hw; ++t) plan.merge(std::move(chunks[t]));
        } else {
            for (uint32_t i = 0; i < n; ++i) plan.add_request(request_json[i] ? request_json[i] : "");
        }
        const auto t1 = std::chrono::steady_clock::now();
        const auto t2 = std::chrono::steady_clock::now();
        if (upload) upload_plan();
        if (getenv("VELOCI_DEBUG")) {
            const auto t3 = std::chrono::steady_clock::now();
            auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
            fprintf(stderr, "[veloci] prepare: parse + plan %.2f ms on %u threads + merge %.2f ms, upload %.2f ms\n", ms(t0, t_joined), hw, ms(t_joined, t1), ms(t2, t3));
        }
    }

    // A plan made by another handle of the same index directory (vgpu_batch_export_plan on the planning process).
    void prepare_from_blob(DeviceIndex* index, const void* blob, size_t len) {
        ix = index;
        vplan::import_plan(ix, blob, len, plan);
        n = (uint32_t)plan.requests.size();
        upload_plan();
    }

    // field search only (vgpu_field_search)
    void prepare_parts(DeviceIndex* index, const std::vector<vhost::SearchPart>& search_parts, std::vector<uint32_t>* part_ids = nullptr) {
        ix = index;
        plan.ix = ix;
        for (auto& p : search_parts) {
            const uint32_t id = plan.add_part(p);  // equal parts share one id
            if (part_ids) part_ids->push_back(id);
        }
        n = 0;
        upload_plan();
    }

    // one request whose single leaf is `part` with the given (term id, score) hits (vgpu_resolve_to_anchor)
    void prepare_term_hits(DeviceIndex* index, const vhost::SearchPart& part, const std::vector<uint32_t>& terms, const std::vector<float>& scores) {
        ix = index;
        plan.ix = ix;
        mode = kTermHits;
        vhost::SearchPart p = part;
        p.top.reset(), p.skip.reset(), p.token_value.reset(), p.is_regex = false;
        if (p.terms.empty()) p.terms.push_back("");
        const uint32_t pid = plan.add_part(p);
        QueryProgram qp;
        memset(&qp, 0, sizeof qp);
        qp.leaf_begin = 0, qp.n_leaves = 1, qp.k = 0, qp.active = 1, qp.emit_all = 1;
        qp.nonneg = 1;
        for (float sc : scores)
            if (!(sc >= 0.0f)) qp.nonneg = 0;
        plan.leaf_part.push_back(pid);
        plan.programs.push_back(qp);
        plan.requests.emplace_back();
        n = 1;
        n_injected = (uint32_t)terms.size();
        d_inj_terms.upload(terms);
        d_inj_scores.upload(scores);
        // every posting may become a hit
        uint64_t cap = 0;
        std::string path = p.path;
        if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
        const vfmt::AnchorScoreView& store = ix->host->get_token_to_anchor(path);
        for (uint32_t t : terms) store.for_each(t, [&](uint32_t, uint32_t) { ++cap; });
        emit_capacity = cap;
        upload_plan();
    }

    // one request over explicit hit lists (union / intersect / add_boost / top_n step entry points)
    // `post`: post ops over the leaves (kPostMulIfPresent / kPostMulValue), applied to the root's score after the boost steps
    void prepare_lists(DeviceIndex* index, const std::vector<ExplicitList>& lists, const std::vector<uint32_t>& code, const std::vector<BoostStep>& boost_steps, uint32_t k, bool all_hits,
                       const vhost::FacetRequest* facet = nullptr, const std::vector<uint32_t>& post = {}) {
        ix = index;
        plan.ix = ix;
        mode = kLists;
        QueryProgram qp;
        memset(&qp, 0, sizeof qp);
        qp.leaf_begin = 0, qp.n_leaves = (uint32_t)lists.size(), qp.k = k, qp.active = 1, qp.emit_all = all_hits ? 1 : 0;
        qp.nonneg = 1;
        bool on_device = false;
        for (auto& l : lists) {
            for (float sc : l.scores)
                if (!(sc >= 0.0f)) qp.nonneg = 0;
            if (l.dev || l.dev_n) on_device = true;
            if (!l.dev_nonneg) qp.nonneg = 0;
        }
        if (on_device)
            for (auto& l : lists)
                if (!l.anchors.empty()) throw std::runtime_error("a step takes its lists either from the host or from the device");
        for (size_t pc = 0; pc < code.size();) {  // (the 1:n boost ops of the program: the tile kernel keeps hit bitmaps for them)
            const uint32_t op = code[pc];
            if (op == kOpLeafBoost) ++qp.n_leaf_boosts;
            pc += op == kOpLeaf ? 2 : op == kOpUnion ? 3 + code[pc + 1] : op == kOpFilter ? 1 : op == kOpLeafBoost ? 4 : 2 + 2 * code[pc + 1];
        }
        const bool trivial = lists.size() == 1 && code.size() == 2 && !facet && post.empty();  // (facets and post ops run on the program path)
        qp.prog_begin = 0, qp.prog_len = trivial ? 0u : (uint32_t)code.size();
        if (!trivial) plan.prog = code;
        qp.post_begin = (uint32_t)plan.prog.size(), qp.post_len = (uint32_t)post.size();
        plan.prog.insert(plan.prog.end(), post.begin(), post.end());
        qp.boost_begin = 0, qp.n_boosts = (uint32_t)boost_steps.size();
        plan.boosts = boost_steps;
        vplan::BatchPlan::set_fast_boost(qp, boost_steps);
        uint64_t total = 0;
        for (size_t i = 0; i < lists.size(); ++i) {
            PartQuery pq;
            memset(&pq, 0, sizeof pq);
            pq.postings = kNoValue;
            pq.flags = lists[i].part_flags;
            plan.parts.push_back(pq);
            plan.part_dict.push_back(0);
            plan.leaf_part.push_back((uint32_t)i);
            total += lists[i].anchors.size() + lists[i].dev_n;
        }
        vplan::RequestPlan rp;
        if (facet) {  // get_facet over the list's ids (facet.rs:31-73)
            plan.add_facet(*facet);
            qp.facet_begin = 0, qp.n_facets = 1;
            rp.facets.push_back(*facet), rp.has_facets = true, rp.facet_begin = 0;
        }
        plan.programs.push_back(qp);
        plan.requests.push_back(std::move(rp));
        plan.max_leaves = std::max<uint32_t>(1, (uint32_t)lists.size());
        plan.max_k = std::max<uint32_t>(1, k);
        n = 1;
        emit_capacity = total;
        upload_plan();
        if (on_device) {
            // device-resident lists: their entries are gathered device to device, their tile offsets found by list_bucket_kernel
            std::vector<PartSlices> slices(lists.size());
            d_bucket.alloc((size_t)lists.size() * (n_tiles + 1));
            d_sparse.alloc((size_t)total + 1);
            uint64_t base = 0;
            for (size_t i = 0; i < lists.size(); ++i) {
                const ExplicitList& l = lists[i];
                slices[i].m_begin = 0, slices[i].n_match = 1, slices[i].n_dense = 0, slices[i].sparse_row = (uint32_t)i, slices[i].sparse_base = base;
                if (l.dev_n) VDEV_CUDA(cudaMemcpyAsync(d_sparse.p + base, l.dev, (size_t)l.dev_n * sizeof(SparseEntry), cudaMemcpyDeviceToDevice, stream));
                launch_list_bucket(stream, d_sparse.p + base, l.dev_n, (uint32_t)ix->anchor_lo, tile_log2, n_tiles, d_bucket.p + i * (n_tiles + 1));
                base += l.dev_n;
            }
            UploadScope staged(stream);
            d_slices.upload(slices);
            return;
        }
        // tile buckets of the lists, built on the host (step seam inputs are small)
        std::vector<uint32_t> bucket((size_t)lists.size() * (n_tiles + 1), 0);
        std::vector<SparseEntry> sparse;
        std::vector<PartSlices> slices(lists.size());
        for (size_t i = 0; i < lists.size(); ++i) {
            const ExplicitList& l = lists[i];
            std::vector<std::pair<uint32_t, uint32_t>> entries;  // dedup keeps the max like resolve_token_to_anchor; inputs are normally unique
            for (size_t j = 0; j < l.anchors.size(); ++j)
                if (l.anchors[j] >= ix->anchor_lo && l.anchors[j] < ix->anchor_hi) entries.emplace_back(l.anchors[j], l.raw_keys.empty() ? vbit::score_key(l.scores[j]) : l.raw_keys[j]);
            std::stable_sort(entries.begin(), entries.end(), [](auto& a, auto& b) { return a.first < b.first; });
            uint32_t* row = &bucket[i * (n_tiles + 1)];
            for (auto& e : entries) row[((e.first - (uint32_t)ix->anchor_lo) >> tile_log2) + 1]++;
            for (uint32_t t = 0; t < n_tiles; ++t) row[t + 1] += row[t];
            slices[i].m_begin = 0, slices[i].n_match = 1, slices[i].n_dense = 0, slices[i].sparse_row = (uint32_t)i, slices[i].sparse_base = sparse.size();
            for (auto& e : entries) sparse.push_back(SparseEntry{e.first, e.second ? e.second : 1u});
        }
        UploadScope staged(stream);
        d_bucket.upload(bucket);
        d_sparse.upload(sparse);
        d_slices.upload(slices);
    }

    void upload_plan() {
        VDEV_CUDA(cudaSetDevice(ix->device));
        device = ix->device;
        n_parts = (uint32_t)plan.parts.size();
        VDEV_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        for (auto& e : ev) VDEV_CUDA(cudaEventCreate(&e));
        UploadScope staged_uploads(stream);  // every table below goes through pinned memory, asynchronously
        n_sms = ix->n_sms;  // (cudaGetDeviceProperties costs milliseconds: asked once, at index open)

        // geometry: the largest tile (<= 8192 anchors) whose part arrays leave room for four CTAs per SM
        const uint32_t L = std::max<uint32_t>(1, plan.max_leaves);
        tile_log2 = 13;
        while (tile_log2 > 10 && ((size_t)L << tile_log2) * 4 > 48 * 1024) --tile_log2;
        while (tile_log2 > 8 && tile_kernel_smem(tile_log2, L) == 0) --tile_log2;  // (requests fanned out over many fields: up to 64 parts at 512 anchors per tile)
        // plane path: requests that are flat `or`s of few parts run on term planes, with plane-sized tiles grouped into items
        use_planes = mode == kRequests && ix->planes.n_planes > 0;
        if (use_planes) {
            uint32_t eligible = 0;
            for (auto& qp : plan.programs)
                if (qp.active && qp.prog_len == 0 && qp.n_leaves <= kFastMaxLeaves && qp.nonneg && qp.k >= 1 && qp.k <= kFastMaxK) ++eligible;
            if (eligible * 2 >= n && eligible > 0 && tile_kernel_smem(kPlaneTileLog2, L) != 0) tile_log2 = kPlaneTileLog2;
            else use_planes = false;
        }
        if (const char* env = probe_env("VELOCI_TILE_LOG2")) {
            int v = atoi(env);
            if (v >= 8 && v <= 13 && tile_kernel_smem((uint32_t)v, L)) tile_log2 = (uint32_t)v;
        }
        const uint64_t span = ix->anchor_hi - ix->anchor_lo;
        n_tiles = (uint32_t)((span + (1ull << tile_log2) - 1) >> tile_log2);
        // plane-path items span `group_tiles` tiles (256 Ki anchors); without a plane path every tile is its own group
        group_tiles = use_planes ? 32u : 1u;
        if (const char* env = probe_env("VELOCI_GROUP_TILES")) group_tiles = use_planes ? (uint32_t)std::min(32, std::max(1, atoi(env))) : 1u;
        n_groups = (n_tiles + group_tiles - 1) / group_tiles;
        stride = std::max<uint32_t>(1, plan.max_k);

        d_parts.upload(plan.parts);
        d_part_dict.upload(plan.part_dict);
        d_leaf_part.upload(plan.leaf_part);
        d_prog.upload(plan.prog);
        d_programs.upload(plan.programs);
        d_boosts.upload(plan.boosts);
        d_boost_members.upload(plan.boost_members);
        h2d_bytes += plan.boost_members.size() * sizeof(BoostListMember);
        d_tl_instances.upload(plan.tl_instances);
        d_tl_term_parts.upload(plan.tl_term_parts);
        d_req_error.alloc(n + 1);
        VDEV_CUDA(cudaMemsetAsync(d_req_error.p, 0, d_req_error.bytes(), stream));
        h2d_bytes += plan.tl_instances.size() * sizeof(TlInstance) + plan.tl_term_parts.size() * 4;
        d_ids_members.upload(plan.ids_members);
        h2d_bytes += plan.ids_members.size() * sizeof(IdsMember);
        d_phrase_members.upload(plan.phrase_members);
        h2d_bytes += plan.phrase_members.size() * sizeof(PhraseMember);
        n_facets = (uint32_t)plan.facets.size();
        if (n_facets) {
            uint64_t total = 0;
            for (auto& f : plan.facets) total += f.hist_size;
            if (total > (1ull << 30)) throw std::runtime_error("facet histograms of this batch exceed 4 GB");
            d_facet_hist.alloc((size_t)total + 1 + (total & 1));  // even word count: the host may reduce it as 64-bit lanes
            uint64_t at = 0;
            facet_stride = 1;
            for (size_t i = 0; i < plan.facets.size(); ++i) {
                plan.facets[i].hist = d_facet_hist.p + at;
                at += plan.facets[i].hist_size;
                facet_stride = std::max(facet_stride, plan.facet_top[i]);
            }
            d_facets.upload(plan.facets);
            d_facet_top.upload(plan.facet_top);
            d_facet_ids.alloc((size_t)n_facets * facet_stride), d_facet_counts.alloc((size_t)n_facets * facet_stride), d_facet_n.alloc(n_facets);
            h2d_bytes += plan.facets.size() * sizeof(FacetStep) + plan.facet_top.size() * 4;
        }
        parts_of_dict.assign(plan.dict_names.size(), {});
        probe_of_dict.assign(plan.dict_names.size(), {});
        max_m_of_dict.assign(plan.dict_names.size(), 0);
        if (mode == kRequests) {
            for (uint32_t p = 0; p < n_parts; ++p) {
                const PartQuery& pq = plan.parts[p];
                if (pq.flags & (kPartList | kPartInjected | kPartRegex)) continue;
                const uint32_t d = plan.part_dict[p];
                bool probe = pq.m >= 1 && pq.d_match <= 2 && !(pq.flags & (kPartPrefix | kPartRawCase));
                if (probe && pq.d_match == 2 && !ix->del_index_built(plan.dict_names[d], 1)) ix->ensure_del_index(plan.dict_names[d], 1);
                probe = probe && ix->del_index_built(plan.dict_names[d], pq.d_match == 2 ? 1 : 0);
                (probe ? probe_of_dict : parts_of_dict)[d].push_back(p);
                max_m_of_dict[d] = std::max(max_m_of_dict[d], pq.m);
            }
        }
        build_regex_tables();
        std::vector<DictView> dv;
        for (auto& name : plan.dict_names) dv.push_back(ix->dict_view(name));
        d_dicts.upload(dv);
        std::vector<PostingsView> pv;
        for (auto& name : plan.postings_names) pv.push_back(ix->postings.at(name).view());
        d_postings.upload(pv);
        const size_t n_dicts = parts_of_dict.size();
        d_parts_of_dict.resize(n_dicts), d_probe_of_dict.resize(n_dicts), d_overflow_of_dict.resize(n_dicts);
        for (size_t d = 0; d < n_dicts; ++d) {
            d_parts_of_dict[d].upload(parts_of_dict[d]), d_probe_of_dict[d].upload(probe_of_dict[d]);
            d_overflow_of_dict[d].alloc(probe_of_dict[d].size() + 1);
            h2d_bytes += (parts_of_dict[d].size() + probe_of_dict[d].size()) * 4;
        }
        d_overflow_count.alloc(n_dicts + 1);
        h2d_bytes += plan.parts.size() * sizeof(PartQuery) + (plan.part_dict.size() + plan.leaf_part.size() + plan.prog.size()) * 4 + plan.programs.size() * sizeof(QueryProgram) +
                     plan.boosts.size() * sizeof(BoostStep) + dv.size() * sizeof(DictView) + pv.size() * sizeof(PostingsView);

        d_counters.alloc(16);  // [0] matches [1] tile work [2] postings [3] dense rows [4] emitted [5,6] item scan [8,9] plane stats [10..12] plane work
        if (use_planes) {
            d_part_planes.alloc(n_parts + 1);
            d_fast.alloc(n + 1);
        }
        d_part_count.alloc(n_parts + 1);
        d_part_begin.alloc(n_parts + 2);
        d_dense_cursor.alloc(n_parts + 1);
        d_sparse_cursor.alloc(n_parts + 1);
        d_part_est.alloc(n_parts + 1);
        d_sparse_total.alloc(n_parts + 1);
        d_sparse_base.alloc(n_parts + 2);
        d_slices.alloc(n_parts + 1);
        d_bucket.alloc((size_t)std::max<uint32_t>(n_parts, 1) * (n_tiles + 1));
        d_heap.alloc((size_t)std::max<uint32_t>(n, 1) * stride);
        d_tau.alloc(n + 1);
        d_num_hits.alloc(n + 1);
        d_lock.alloc(n + 1);
        d_out_keys.alloc((size_t)std::max<uint32_t>(n, 1) * stride);
        d_out_hits.alloc(n + 1);
        d_emit.alloc((size_t)emit_capacity + 1);
        if (mode == kRequests) d_records.reserve(std::max<size_t>(1u << 20, (size_t)n_parts * 64));
        VDEV_CUDA(cudaMemsetAsync(d_counters.p, 0, d_counters.bytes(), stream));
    }

    // The DFA of every regex part (host/regex_dfa.hpp), re-expressed over the alphabet codes of the part's dictionary:
    // per part a class map [alphabet size] and the transition table [states x classes], all in one device array.
    void build_regex_tables() {
        regex_of_dict.assign(plan.dict_names.size(), {});
        d_regex_of_dict.clear();
        d_regex_of_dict.resize(plan.dict_names.size());
        if (plan.regex_parts.empty() || mode != kRequests) return;
        std::vector<uint16_t> tables;
        struct Placed {
            uint32_t dict;
            size_t class_at, trans_at;
            RegexPartDev dev;
        };
        std::vector<Placed> placed;
        for (auto& rp : plan.regex_parts) {
            const vregex::Dfa dfa = vregex::compile(rp.pattern, rp.case_insensitive);  // (validated by the planner: does not throw here)
            const uint32_t d = plan.part_dict[rp.part];
            const DictDev& dict = ix->dicts.at(plan.dict_names[d]);
            Placed pl;
            pl.dict = d, pl.class_at = tables.size();
            for (uint32_t scalar : dict.alphabet) tables.push_back(dfa.class_of(scalar));
            if (dict.alphabet.empty()) tables.push_back(0);
            pl.trans_at = tables.size();
            tables.insert(tables.end(), dfa.trans.begin(), dfa.trans.end());
            pl.dev.class_of_code = nullptr, pl.dev.trans = nullptr;
            pl.dev.n_classes = dfa.n_classes, pl.dev.start = dfa.start, pl.dev.part = rp.part, pl.dev.sticky = rp.starts_with ? 1u : 0u;
            placed.push_back(pl);
        }
        d_regex_tables.upload(tables);
        for (auto& pl : placed) {
            pl.dev.class_of_code = d_regex_tables.p + pl.class_at, pl.dev.trans = d_regex_tables.p + pl.trans_at;
            regex_of_dict[pl.dict].push_back(pl.dev);
        }
        for (size_t d = 0; d < regex_of_dict.size(); ++d) d_regex_of_dict[d].upload(regex_of_dict[d]);
        h2d_bytes += tables.size() * 2 + placed.size() * sizeof(RegexPartDev);
    }

    // Matched terms with at least this many postings get a tile-offset row, the others are copied into their part's
    // sparse tile buckets.  With the plane path every term without a plane goes to the buckets (its postings are the
    // "entries" of the plane-path items); only head terms keep offset rows (for the items that take the general path).
    uint32_t dense_min() const {
        const uint32_t base = std::max<uint32_t>(1, n_tiles / 2);
        if (!use_planes) return base;
        return 0xFFFFFFFFu;
    }

    // count pass (la.sparse == nullptr) or fill pass of every list producer of the batch
    void run_list_producers(const ListArgs& la) {
        timed("phrase_pairs", [&] { launch_phrase_pairs(stream, d_phrase_members.p, (uint32_t)plan.phrase_members.size(), la); });
        timed("ids_to_anchor", [&] { launch_ids_to_anchor(stream, d_ids_members.p, (uint32_t)plan.ids_members.size(), la); });
        timed("boost_to_anchor", [&] { launch_boost_to_anchor(stream, d_boost_members.p, (uint32_t)plan.boost_members.size(), la); });
        timed("text_locality", [&] { launch_text_locality(stream, d_tl_instances.p, (uint32_t)plan.tl_instances.size(), d_tl_term_parts.p, d_req_error.p, la); });
    }

    template <class T>
    T read_back(const T* dev) {
        T v;
        VDEV_CUDA(cudaMemcpyAsync(&v, dev, sizeof(T), cudaMemcpyDeviceToHost, stream));
        VDEV_CUDA(cudaStreamSynchronize(stream));
        d2h_bytes += sizeof(T);
        return v;
    }

    // ------------------------------------------------------------ phases 0-1
    void run_match() {
        VDEV_CUDA(cudaSetDevice(ix->device));
        VDEV_CUDA(cudaEventRecord(ev[0], stream));
        uint64_t n_match = 0;
        if (mode == kRequests) {
            const size_t n_dicts = parts_of_dict.size();
            std::vector<unsigned long long> overflow(n_dicts + 1);
            std::vector<MatchRecord> injected;
            if (!plan.bounded.empty()) match_bounded_parts(injected);
            if (injected.size() > d_records.n) d_records.reserve(injected.size() + 1024);
            for (int attempt = 0; attempt < 2; ++attempt) {
                VDEV_CUDA(cudaMemsetAsync(d_counters.p, 0, d_counters.bytes(), stream));
                VDEV_CUDA(cudaMemsetAsync(d_overflow_count.p, 0, d_overflow_count.bytes(), stream));
                if (!injected.empty()) {  // the given hits take the first records; the match kernels append after them
                    const unsigned long long n_inj = injected.size();
                    VDEV_CUDA(cudaMemcpyAsync(d_records.p, injected.data(), injected.size() * sizeof(MatchRecord), cudaMemcpyHostToDevice, stream));
                    VDEV_CUDA(cudaMemcpyAsync(d_counters.p, &n_inj, 8, cudaMemcpyHostToDevice, stream));
                    VDEV_CUDA(cudaStreamSynchronize(stream));
                }
                const uint32_t capacity = (uint32_t)std::min<size_t>(d_records.n, 0xFFFFFFFFu);
                for (size_t d = 0; d < n_dicts; ++d) {
                    const DictView dict = ix->dict_view(plan.dict_names[d]);
                    timed("fuzzy_probe", [&] { launch_fuzzy_probe(stream, dict, d_parts.p, d_probe_of_dict[d].p, (uint32_t)probe_of_dict[d].size(), d_records.p, capacity, d_counters.p, d_overflow_of_dict[d].p,
                                       d_overflow_count.p + d); });
                    timed("fuzzy_match", [&] { launch_fuzzy_match(stream, dict, d_parts.p, d_parts_of_dict[d].p, (uint32_t)parts_of_dict[d].size(), max_m_of_dict[d], d_records.p, capacity, d_counters.p); });
                    if (!regex_of_dict[d].empty())
                        timed("regex_match", [&] { launch_regex_match(stream, dict, d_regex_of_dict[d].p, (uint32_t)regex_of_dict[d].size(), d_records.p, capacity, d_counters.p); });
                }
                VDEV_CUDA(cudaMemcpyAsync(overflow.data(), d_overflow_count.p, n_dicts * 8, cudaMemcpyDeviceToHost, stream));
                n_match = read_back(d_counters.p);
                d2h_bytes += n_dicts * 8;
                bool rescanned = false;
                for (size_t d = 0; d < n_dicts; ++d)
                    if (overflow[d]) {  // candidate sets too large for the probe: scan the dictionary for those parts
                        const DictView dict = ix->dict_view(plan.dict_names[d]);
                        timed("fuzzy_match", [&] { launch_fuzzy_match(stream, dict, d_parts.p, d_overflow_of_dict[d].p, (uint32_t)overflow[d], max_m_of_dict[d], d_records.p, capacity, d_counters.p); });
                        rescanned = true;
                    }
                if (rescanned) n_match = read_back(d_counters.p);
                if (n_match <= d_records.n) break;
                if (n_match > 0xFFFFFFF0ull) throw std::runtime_error("more than 2^32 term matches in one batch");
                d_records.reserve((size_t)n_match);
            }
        } else {  // kTermHits: the hits are given, one record per hit
            VDEV_CUDA(cudaMemsetAsync(d_counters.p, 0, d_counters.bytes(), stream));
            std::vector<MatchRecord> rec(n_injected);
            for (uint32_t i = 0; i < n_injected; ++i) rec[i] = MatchRecord{0u, i};
            UploadScope staged(stream);
            d_records.upload(rec);
            n_match = n_injected;
        }
        stat_matches = n_match;
        const uint32_t M = (uint32_t)n_match;
        VDEV_CUDA(cudaEventRecord(ev[1], stream));
        d_g_term.reserve(M), d_g_df.reserve(M), d_g_row.reserve(M), d_g_part.reserve(M), d_g_score.reserve(M), d_g_begin.reserve(M), d_row_match.reserve(M);
        if (use_planes) {
            d_g_plane.reserve(M);
            VDEV_CUDA(cudaMemsetAsync(d_part_planes.p, 0, d_part_planes.bytes(), stream));
        }
        VDEV_CUDA(cudaMemsetAsync(d_part_count.p, 0, d_part_count.bytes(), stream));
        VDEV_CUDA(cudaMemsetAsync(d_dense_cursor.p, 0, d_dense_cursor.bytes(), stream));
        VDEV_CUDA(cudaMemsetAsync(d_sparse_cursor.p, 0, d_sparse_cursor.bytes(), stream));
        VDEV_CUDA(cudaMemsetAsync(d_part_est.p, 0, d_part_est.bytes(), stream));
        timed("group_count", [&] { launch_group_count(stream, d_records.p, M, d_part_count.p); });
        timed("scan_u32", [&] { launch_scan_u32(stream, d_part_count.p, d_part_begin.p, n_parts); });
        ScoreScatterArgs a;
        a.records = d_records.p, a.n_records = M, a.parts = d_parts.p, a.part_dict = d_part_dict.p, a.dicts = d_dicts.p, a.postings = d_postings.p;
        a.part_begin = d_part_begin.p, a.dense_cursor = d_dense_cursor.p, a.sparse_cursor = d_sparse_cursor.p, a.n_dense_rows = reinterpret_cast<uint32_t*>(d_counters.p + 3);
        a.dense_row_capacity = M, a.dense_min = dense_min(), a.row_match = d_row_match.p, a.part_est = d_part_est.p;
        a.g_term = d_g_term.p, a.g_score = d_g_score.p, a.g_begin = d_g_begin.p, a.g_df = d_g_df.p, a.g_row = d_g_row.p, a.g_part = d_g_part.p;
        const bool has_injected = mode == kTermHits || !plan.bounded.empty();
        a.inj_terms = has_injected ? d_inj_terms.p : nullptr, a.inj_scores = has_injected ? d_inj_scores.p : nullptr, a.inj_all = mode == kTermHits ? 1u : 0u;
        a.part_planes = use_planes ? d_part_planes.p : nullptr, a.g_plane = use_planes ? d_g_plane.p : nullptr;
        timed("score_scatter", [&] { launch_score_scatter(stream, a); });
        VDEV_CUDA(cudaEventRecord(ev[2], stream));
        VDEV_CUDA(cudaStreamSynchronize(stream));
        VDEV_CUDA(cudaGetLastError());
        matched = true;
    }

    // Parts with a per-part `top` (search_field.rs:292-294) or a `token_value` boost (:391-395): matched and scored on the device in
    // a batch of their own, bounded on the host in FST order (bound_part_hits), boosted by their token values, and handed to
    // this batch as given (term id, score) hits.
    void match_bounded_parts(std::vector<MatchRecord>& records) {
        std::vector<vhost::SearchPart> device_parts;
        for (auto& bp : plan.bounded) {
            vhost::SearchPart p = bp.request;
            p.top.reset(), p.skip.reset(), p.boost.reset(), p.token_value.reset();
            device_parts.push_back(std::move(p));
        }
        Batch own;
        std::vector<uint32_t> ids;
        own.prepare_parts(ix, device_parts, &ids);
        own.run_match();
        std::vector<uint32_t> inj_terms;
        std::vector<float> inj_scores;
        for (size_t i = 0; i < plan.bounded.size(); ++i) {
            std::vector<TermHit> hits;
            own.download_part_hits(ids[i], hits);
            bound_part_hits(plan.bounded[i].request, hits);
            apply_token_value(*ix->host, plan.bounded[i].request, hits);
            for (const TermHit& h : hits) {
                records.push_back(MatchRecord{plan.bounded[i].part, (uint32_t)inj_terms.size()});
                inj_terms.push_back(h.id), inj_scores.push_back(h.score);
            }
        }
        {
            UploadScope staged(stream);  // ordered with the kernels of this batch's stream that read them
            d_inj_terms.upload(inj_terms), d_inj_scores.upload(inj_scores);
        }
        h2d_bytes += inj_terms.size() * 8 + records.size() * sizeof(MatchRecord);
    }

    // (term id, score) hits of one part after run_match, in ascending term id order (the FST stream's)
    void download_part_hits(uint32_t part, std::vector<TermHit>& hits) {
        std::vector<uint32_t> terms;
        std::vector<float> scores;
        download_matches(part, terms, scores);
        hits.resize(terms.size());
        for (size_t j = 0; j < terms.size(); ++j) hits[j] = TermHit{terms[j], scores[j]};
        std::sort(hits.begin(), hits.end(), [](const TermHit& x, const TermHit& y) { return x.id < y.id; });
    }

    // (term id, score) hits of one part after run_match, unordered
    void download_matches(uint32_t part, std::vector<uint32_t>& terms, std::vector<float>& scores) {
        std::vector<uint32_t> begin(n_parts + 1);
        VDEV_CUDA(cudaMemcpy(begin.data(), d_part_begin.p, begin.size() * 4, cudaMemcpyDeviceToHost));
        const uint32_t a = begin[part], cnt = begin[part + 1] - begin[part];
        terms.resize(cnt), scores.resize(cnt);
        if (cnt) {
            VDEV_CUDA(cudaMemcpy(terms.data(), d_g_term.p + a, cnt * 4, cudaMemcpyDeviceToHost));
            VDEV_CUDA(cudaMemcpy(scores.data(), d_g_score.p + a, cnt * 4, cudaMemcpyDeviceToHost));
        }
    }

    // ------------------------------------------------------------ all phases
    void execute() {
        if (ix->comm && ix->comm->n_ranks > 1 && mode == kRequests) return execute_sharded();
        execute_begin(false);  // no host round trip between the seed pass and the bulk pass
        execute_finish();
    }

    // The whole step on an anchor-range shard with a communicator (vgpu_comm_init): seed pass -> all-reduce(max) of the
    // requests' thresholds -> bulk pass -> all-gather of the shard-local top-k rows and hit counts (all-reduce(sum) of
    // the facet histograms) -> final merge, all on the batch's stream, no host synchronisation between them.  Every rank
    // ends up with the complete result.  Collective: every rank of the communicator must execute the same batch.
    DevBuf<uint64_t> d_gather_keys, d_gather_hits;
    void execute_sharded() {
        ShardComm& c = *ix->comm;
        execute_begin(false);
        try {
            if (n) c.all_reduce_max_u64(d_tau.p, n, stream);  // the global k-th best is at least the largest local one
            execute_finish(false);
            const uint32_t R = c.n_ranks;
            d_gather_keys.reserve((size_t)R * std::max<uint32_t>(n, 1) * stride), d_gather_hits.reserve((size_t)R * std::max<uint32_t>(n, 1));
            if (n) {
                c.group_start();
                c.all_gather_u64(d_out_keys.p, d_gather_keys.p, (size_t)n * stride, stream);
                c.all_gather_u64(d_out_hits.p, d_gather_hits.p, n, stream);
                if (n_facets) c.all_reduce_sum_u32(d_facet_hist.p, d_facet_hist.n, stream);
                c.group_end();
                timed("merge_heaps", [&] { launch_merge_heaps(stream, d_gather_keys.p, d_gather_hits.p, R, n, stride, d_programs.p, d_out_keys.p, d_out_hits.p); });
                if (n_facets) timed("facet_topk", [&] { launch_facet_topk(stream, d_facets.p, d_facet_top.p, n_facets, facet_stride, d_facet_ids.p, d_facet_counts.p, d_facet_n.p); });
            }
        } catch (const NcclError&) {
            cudaStreamSynchronize(stream);
            throw;
        }
        VDEV_CUDA(cudaEventRecord(ev[6], stream));  // the final phase now ends after the exchange and the merge
        finish_sync();
    }

    // PlaneArgs of one stage of the plane evaluation: the items of tile groups [g0, g1)
    PlaneArgs plane_stage_args(int stage, uint32_t g0, uint32_t g1) {
        PlaneArgs a;
        memset(&a, 0, sizeof a);
        a.items = d_fast_items.p, a.group_item_begin = d_fast_item_begin.p, a.fast = d_fast.p;
        a.sparse = d_sparse.p, a.planes = ix->planes.view();
        a.anchor_lo = (uint32_t)ix->anchor_lo, a.anchor_hi = (uint32_t)std::min<uint64_t>(ix->anchor_hi, 0xFFFFFFFFull);
        a.group_tiles = group_tiles, a.group_begin = g0, a.group_end = g1;
        a.ones_row = ix->planes.const_rows.p, a.zeros_row = ix->planes.const_rows.p + PlaneSetDev::kConstRowWords;
        // one item per trip to the work counter (batches of four: +0.1 ms at N = 1, +0.2 ms on a 1/8 shard: tail imbalance) and
        // the best-boosted 128 Ki anchors of the seed set (`tools/seed_probe.py`, profiles/r02x_seed_probe.jsonl)
        a.item_batch = 1;
        a.seed_sweep_words = 4096;
        if (const char* env = probe_env("VELOCI_ITEM_BATCH")) a.item_batch = (uint32_t)std::max(1, atoi(env));
        if (const char* env = probe_env("VELOCI_SEED_WORDS")) a.seed_sweep_words = (uint32_t)std::max(128, atoi(env) / 128 * 128);
        a.heap = d_heap.p, a.heap_stride = stride, a.tau = d_tau.p, a.lock = d_lock.p, a.num_hits = d_num_hits.p;
        a.stats = d_counters.p + 8;
        a.work_counter = d_counters.p + 10 + stage;
        return a;
    }

    // Everything up to the point where the requests' thresholds (k-th best so far) are worth sharing between anchor-range
    // shards: match, slicing, item scan and the seed pass of the plane evaluation.  Returns with the stream idle.
    void execute_begin(bool sync_at_end = true) {
        VDEV_CUDA(cudaSetDevice(ix->device));
        executed = false, fetched = false, begun = false;
        d2h_bytes = 0;
        spans.clear(), spans_used = 0;
        uint32_t M = 0;
        if (mode != kLists) {
            run_match();
            M = (uint32_t)stat_matches;
            // ---- phase 2: slicing
            const uint32_t n_rows = read_back(reinterpret_cast<uint32_t*>(d_counters.p + 3));
            d_toff.reserve((size_t)std::max<uint32_t>(n_rows, 1) * (n_tiles + 1));
            DenseOffsetsArgs da;
            da.row_match = d_row_match.p, da.g_part = d_g_part.p, da.g_begin = d_g_begin.p, da.g_df = d_g_df.p, da.parts = d_parts.p, da.postings = d_postings.p;
            da.toff = d_toff.p, da.n_tiles = n_tiles, da.tile_log2 = tile_log2, da.anchor_lo = (uint32_t)ix->anchor_lo;
            timed("dense_tile_offsets", [&] { launch_dense_tile_offsets(stream, da, n_rows); });
            VDEV_CUDA(cudaMemsetAsync(d_bucket.p, 0, d_bucket.bytes(), stream));
            SparseArgs sa;
            sa.n_matches = M, sa.g_row = d_g_row.p, sa.g_df = d_g_df.p, sa.g_part = d_g_part.p, sa.g_begin = d_g_begin.p, sa.g_score = d_g_score.p;
            sa.parts = d_parts.p, sa.postings = d_postings.p, sa.bucket = d_bucket.p, sa.sparse_base = d_sparse_base.p, sa.sparse = nullptr;
            sa.n_tiles = n_tiles, sa.tile_log2 = tile_log2, sa.anchor_lo = (uint32_t)ix->anchor_lo;
            sa.max_df = (uint32_t)std::min<uint64_t>(use_planes ? ix->max_nonplane_list : std::min<uint64_t>(dense_min(), ix->max_posting_list), 0xFFFFFFFFull);
            timed("sparse_count", [&] { launch_sparse_count(stream, sa); });
            ListArgs la;
            la.part_begin = d_part_begin.p, la.g_term = d_g_term.p, la.bucket = d_bucket.p, la.sparse_base = d_sparse_base.p, la.sparse = nullptr;
            la.n_tiles = n_tiles, la.tile_log2 = tile_log2, la.anchor_lo = (uint32_t)ix->anchor_lo, la.anchor_hi = (uint32_t)std::min<uint64_t>(ix->anchor_hi, 0xFFFFFFFFull);
            run_list_producers(la);
            timed("sparse_scan", [&] { launch_sparse_scan(stream, d_bucket.p, n_tiles, d_sparse_total.p, n_parts); });
            timed("scan_u64", [&] { launch_scan_u64(stream, d_sparse_total.p, d_sparse_base.p, n_parts); });
            const uint64_t n_sparse = read_back(d_sparse_base.p + n_parts);
            stat_sparse = n_sparse;
            d_sparse.reserve((size_t)std::max<uint64_t>(n_sparse, 1));
            sa.sparse = d_sparse.p;
            timed("sparse_fill", [&] { launch_sparse_fill(stream, sa); });
            la.sparse = d_sparse.p;
            run_list_producers(la);
            timed("part_slices", [&] { launch_part_slices(stream, d_slices.p, d_part_begin.p, d_dense_cursor.p, d_sparse_base.p, d_parts.p, n_parts); });
            timed("finalize_programs", [&] { launch_finalize_programs(stream, d_programs.p, n, d_prog.p, d_leaf_part.p, d_part_est.p, d_counters.p + 2); });
            if (use_planes) timed("build_fast_desc", [&] { launch_build_fast_desc(stream, d_programs.p, n, d_leaf_part.p, d_part_planes.p, ix->planes.wmax.p, d_fast.p); });
            if (use_planes && getenv("VELOCI_DEBUG")) {
                std::vector<FastDesc> fd(n);
                std::vector<PartPlanes> pp(n_parts);
                VDEV_CUDA(cudaStreamSynchronize(stream));
                VDEV_CUDA(cudaMemcpy(fd.data(), d_fast.p, n * sizeof(FastDesc), cudaMemcpyDeviceToHost));
                VDEV_CUDA(cudaMemcpy(pp.data(), d_part_planes.p, n_parts * sizeof(PartPlanes), cudaMemcpyDeviceToHost));
                uint32_t not_fast = 0, hist[8] = {};
                for (auto& d : fd) not_fast += d.flags == 0;
                for (auto& p : pp) hist[std::min<uint32_t>(p.n, 7)]++;
                for (uint32_t q = 0; q < n; ++q)
                    if (fd[q].flags == 0) {
                        const QueryProgram& qp = plan.programs[q];
                        fprintf(stderr, "[veloci] request %u off the plane path: active %u prog_len %u leaves %u nonneg %u k %u boosts %u fb_flags %u post %u facets %u planes", q, qp.active,
                                qp.prog_len, qp.n_leaves, qp.nonneg, qp.k, qp.n_boosts, qp.fb_flags, qp.post_len, qp.n_facets);
                        for (uint32_t l = 0; l < qp.n_leaves; ++l) fprintf(stderr, " %u", pp[plan.leaf_part[qp.leaf_begin + l]].n);
                        fprintf(stderr, "\n");
                    }
                fprintf(stderr, "[veloci] requests %u not on the plane path %u; parts by plane matches:", n, not_fast);
                for (int i = 0; i < 8; ++i) fprintf(stderr, " %u", hist[i]);
                fprintf(stderr, "\n");
            }
        } else {
            VDEV_CUDA(cudaMemsetAsync(d_counters.p, 0, d_counters.bytes(), stream));
            VDEV_CUDA(cudaEventRecord(ev[0], stream));
            VDEV_CUDA(cudaEventRecord(ev[1], stream));
            VDEV_CUDA(cudaEventRecord(ev[2], stream));
        }
        // (tile group, request) pairs: plane-path items grouped per group, general items (one per non-empty tile) tile-major
        unsigned long long n_items = 0;
        uint32_t n_fast_items = 0;
        const bool planes_on = use_planes && mode == kRequests;
        {
            ItemScanArgs sc;
            memset(&sc, 0, sizeof sc);
            sc.queries = d_programs.p, sc.n_queries = n, sc.leaf_part = d_leaf_part.p, sc.slices = d_slices.p, sc.parts = d_parts.p;
            sc.g_row = d_g_row.p, sc.g_begin = d_g_begin.p, sc.g_score = d_g_score.p, sc.toff = d_toff.p, sc.bucket = d_bucket.p;
            sc.n_tiles = n_tiles, sc.group_tiles = group_tiles, sc.n_groups = n_groups, sc.n_pairs_total = (unsigned long long)n_groups * n;
            sc.counters = d_counters.p + 5, sc.items = nullptr, sc.slice_recs = nullptr;
            d_fast_item_cursor.reserve(n_groups + 1), d_fast_item_begin.reserve(n_groups + 2);
            sc.fast_item_cursor = d_fast_item_cursor.p, sc.fast_item_begin = d_fast_item_begin.p;
            if (planes_on) sc.fast = d_fast.p, sc.g_plane = d_g_plane.p, sc.plane_tprefix = ix->planes.tprefix.p;
            VDEV_CUDA(cudaMemsetAsync(d_fast_item_cursor.p, 0, d_fast_item_cursor.bytes(), stream));
            VDEV_CUDA(cudaMemsetAsync(d_counters.p + 5, 0, 16, stream));
            timed("item_scan", [&] { launch_item_scan(stream, sc, false); });
            unsigned long long counts[2];
            VDEV_CUDA(cudaMemcpyAsync(counts, d_counters.p + 5, 16, cudaMemcpyDeviceToHost, stream));
            if (planes_on) {
                timed("scan_u32", [&] { launch_scan_u32(stream, d_fast_item_cursor.p, d_fast_item_begin.p, n_groups); });
                VDEV_CUDA(cudaMemcpyAsync(&n_fast_items, d_fast_item_begin.p + n_groups, 4, cudaMemcpyDeviceToHost, stream));
                d2h_bytes += 4;
            }
            VDEV_CUDA(cudaStreamSynchronize(stream));
            d2h_bytes += 16;
            n_items = counts[0];
            d_items.reserve((size_t)std::max<unsigned long long>(n_items, 1));
            d_slice_recs.reserve((size_t)std::max<unsigned long long>(counts[1], 1));
            sc.items = d_items.p, sc.slice_recs = d_slice_recs.p;
            if (planes_on) {
                d_fast_items.reserve(std::max<size_t>(n_fast_items, 1));
                sc.fast_items = d_fast_items.p;
            }
            VDEV_CUDA(cudaMemsetAsync(d_fast_item_cursor.p, 0, d_fast_item_cursor.bytes(), stream));
            VDEV_CUDA(cudaMemsetAsync(d_counters.p + 5, 0, 16, stream));
            timed("item_scan", [&] { launch_item_scan(stream, sc, true); });
        }
        stat_fast_items = n_fast_items, stat_general_items = n_items;
        VDEV_CUDA(cudaEventRecord(ev[3], stream));
        VDEV_CUDA(cudaMemsetAsync(d_heap.p, 0, d_heap.bytes(), stream));
        VDEV_CUDA(cudaMemsetAsync(d_tau.p, 0, d_tau.bytes(), stream));
        VDEV_CUDA(cudaMemsetAsync(d_num_hits.p, 0, d_num_hits.bytes(), stream));
        VDEV_CUDA(cudaMemsetAsync(d_lock.p, 0, d_lock.bytes(), stream));
        if (n_facets) VDEV_CUDA(cudaMemsetAsync(d_facet_hist.p, 0, d_facet_hist.bytes(), stream));
        // ---- phase 3: plane evaluation.  First the threshold seeds: one pass over the boost column's seed set (the shard's
        // best-boosted eighth) gives every request a threshold near its final one.  The seeds are what anchor-range shards
        // share (all-reduce MAX) before any of them sweeps.
        pending_items = n_items, pending_fast_items = n_fast_items;
        if (planes_on && n_fast_items) timed("plane_seed", [&] { launch_plane_seed(stream, plane_stage_args(0, 0, 0), n, n_sms); });
        if (sync_at_end) {
            VDEV_CUDA(cudaStreamSynchronize(stream));
            VDEV_CUDA(cudaGetLastError());
        }
        begun = true;
    }

    // The rest of the step: the bulk of the plane evaluation, the general items, the final order.
    void execute_finish(bool sync_at_end = true) {
        if (!begun) throw std::runtime_error("execute_finish without execute_begin");
        VDEV_CUDA(cudaSetDevice(ix->device));
        begun = false;
        const unsigned long long n_items = pending_items;
        const bool planes_on = use_planes && mode == kRequests;
        if (planes_on && pending_fast_items) timed("plane_eval", [&] { launch_plane_eval(stream, plane_stage_args(2, 0, n_groups), n_sms); });
        VDEV_CUDA(cudaEventRecord(ev[4], stream));
        // ---- phase 4: tile evaluation of the general items
        {
            TileArgs a;
            a.items = d_items.p, a.slice_recs = d_slice_recs.p;
            a.queries = d_programs.p, a.n_queries = n, a.leaf_part = d_leaf_part.p, a.prog = d_prog.p, a.boosts = d_boosts.p, a.facets = d_facets.p;
            a.parts = d_parts.p, a.slices = d_slices.p, a.postings = d_postings.p, a.g_score = d_g_score.p, a.g_begin = d_g_begin.p, a.g_row = d_g_row.p, a.g_df = d_g_df.p;
            a.toff = d_toff.p, a.bucket = d_bucket.p, a.sparse = d_sparse.p;
            a.g_plane = use_planes ? d_g_plane.p : nullptr, a.plane_tprefix = use_planes ? ix->planes.tprefix.p : nullptr;
            a.n_tiles = n_tiles, a.tile_log2 = tile_log2, a.anchor_lo = (uint32_t)ix->anchor_lo, a.anchor_hi = (uint32_t)std::min<uint64_t>(ix->anchor_hi, 0xFFFFFFFFull);
            a.max_leaves = std::max<uint32_t>(1, plan.max_leaves);
            a.heap = d_heap.p, a.heap_stride = stride, a.tau = d_tau.p, a.lock = d_lock.p, a.num_hits = d_num_hits.p;
            if (stride > kMaxK) d_merge_scratch.reserve((size_t)n_sms * kTileBlocksPerSm * stride);
            a.merge_scratch = stride > kMaxK ? d_merge_scratch.p : nullptr;
            a.work_counter = d_counters.p + 1, a.n_items = n_items;
            a.emit = d_emit.p, a.emit_count = d_counters.p + 4, a.emit_capacity = emit_capacity;
            timed("tile_eval", [&] { launch_tile_eval(stream, a, n_sms); });
        }
        VDEV_CUDA(cudaEventRecord(ev[5], stream));
        // ---- phase 5: final order of the local heaps, top groups of the facet histograms
        if (n_facets) timed("facet_topk", [&] { launch_facet_topk(stream, d_facets.p, d_facet_top.p, n_facets, facet_stride, d_facet_ids.p, d_facet_counts.p, d_facet_n.p); });
        timed("merge_heaps", [&] { launch_merge_heaps(stream, reinterpret_cast<const uint64_t*>(d_heap.p), reinterpret_cast<const uint64_t*>(d_num_hits.p), 1, n, stride, d_programs.p, d_out_keys.p, d_out_hits.p); });
        VDEV_CUDA(cudaEventRecord(ev[6], stream));
        if (sync_at_end) finish_sync();
    }

    void finish_sync() {
        VDEV_CUDA(cudaStreamSynchronize(stream));
        VDEV_CUDA(cudaGetLastError());
        for (int p = 0; p < kPhases; ++p) VDEV_CUDA(cudaEventElapsedTime(&phase_ms[p], ev[p], ev[p + 1]));
        executed = true;
    }

    void merge_gathered(const uint64_t* keys_dev, const uint64_t* hits_dev, uint32_t n_shards) {
        VDEV_CUDA(cudaSetDevice(ix->device));
        timed("merge_heaps", [&] { launch_merge_heaps(stream, keys_dev, hits_dev, n_shards, n, stride, d_programs.p, d_out_keys.p, d_out_hits.p); });
        // the caller has summed the shards' facet histograms (vgpu_batch_facet_histograms) by now: pick the groups again
        if (n_facets) timed("facet_topk", [&] { launch_facet_topk(stream, d_facets.p, d_facet_top.p, n_facets, facet_stride, d_facet_ids.p, d_facet_counts.p, d_facet_n.p); });
        VDEV_CUDA(cudaStreamSynchronize(stream));
        fetched = false;
    }

    void fetch() {
        if (fetched) return;
        if (!executed) throw std::runtime_error("batch was not executed");
        VDEV_CUDA(cudaSetDevice(ix->device));
        h_keys.resize((size_t)n * stride);
        h_hits.resize(n);
        if (n) {
            VDEV_CUDA(cudaMemcpyAsync(h_keys.data(), d_out_keys.p, h_keys.size() * 8, cudaMemcpyDeviceToHost, stream));
            VDEV_CUDA(cudaMemcpyAsync(h_hits.data(), d_out_hits.p, h_hits.size() * 8, cudaMemcpyDeviceToHost, stream));
            uint64_t stats[16];
            VDEV_CUDA(cudaMemcpyAsync(stats, d_counters.p, sizeof stats, cudaMemcpyDeviceToHost, stream));
            VDEV_CUDA(cudaStreamSynchronize(stream));
            stat_postings = stats[2];
            stat_plane_evaluated = stats[9];
            stat_plane_item_evals = stats[8], stat_plane_unconverged = stats[13], stat_plane_sweepless = stats[14];
            if (getenv("VELOCI_DEBUG") && use_planes)
                fprintf(stderr, "[veloci] plane path: %llu item evaluations (seed + bulk), %llu candidates, %llu items swept before their threshold converged, %llu answered from plane counts\n",
                        (unsigned long long)stats[8], (unsigned long long)stats[9], (unsigned long long)stats[13], (unsigned long long)stats[14]);
            d2h_bytes += h_keys.size() * 8 + h_hits.size() * 8 + sizeof stats;
            stat_union = 0;
            for (uint64_t h : h_hits) stat_union += h;
        }
        if (!plan.tl_instances.empty()) {  // requests the device could not finish (text locality over too many matched tokens)
            h_req_error.resize(n);
            VDEV_CUDA(cudaMemcpy(h_req_error.data(), d_req_error.p, n * 4, cudaMemcpyDeviceToHost));
            d2h_bytes += n * 4;
            for (uint32_t q = 0; q < n; ++q)
                if (h_req_error[q] && plan.requests[q].status == 0) {
                    plan.requests[q].status = 8;
                    plan.requests[q].message = "text_locality over more than 256 matched tokens in one field is outside the accelerated path";
                }
        }
        if (n_facets) {
            h_facet_ids.resize((size_t)n_facets * facet_stride), h_facet_counts.resize((size_t)n_facets * facet_stride), h_facet_n.resize(n_facets);
            VDEV_CUDA(cudaMemcpyAsync(h_facet_ids.data(), d_facet_ids.p, h_facet_ids.size() * 4, cudaMemcpyDeviceToHost, stream));
            VDEV_CUDA(cudaMemcpyAsync(h_facet_counts.data(), d_facet_counts.p, h_facet_counts.size() * 4, cudaMemcpyDeviceToHost, stream));
            VDEV_CUDA(cudaMemcpyAsync(h_facet_n.data(), d_facet_n.p, h_facet_n.size() * 4, cudaMemcpyDeviceToHost, stream));
            VDEV_CUDA(cudaStreamSynchronize(stream));
            d2h_bytes += (h_facet_ids.size() * 2 + h_facet_n.size()) * 4;
        }
        fetched = true;
    }

    // SearchResult.data of request q after apply_top_skip (search.rs:230-239)
    uint32_t result(uint32_t q, uint64_t* num_hits, vgpu_hit_pod* hits, uint32_t cap) {
        fetch();
        const vplan::RequestPlan& rp = plan.requests[q];
        if (num_hits) *num_hits = rp.status == 0 ? h_hits[q] : 0;
        if (rp.status != 0) return 0;
        const uint64_t* row = &h_keys[(size_t)q * stride];
        uint32_t avail = 0;
        while (avail < stride && avail < rp.top + rp.skip && row[avail] != 0) ++avail;
        uint32_t w = 0;
        for (uint64_t i = rp.skip; i < avail && w < rp.top; ++i, ++w) {
            if (hits && w < cap) {
                hits[w].id = (uint32_t)(row[i] & 0xFFFFFFFFull);
                hits[w].score = vbit::key_score((uint32_t)(row[i] >> 32));
            }
        }
        return hits ? std::min(w, cap) : w;
    }

    // Step seam: every hit of request 0 by ascending anchor id (`all`), or its top-k in rank order.
    template <class Hit>
    void download_hits(uint32_t q, bool all, std::vector<Hit>& hits) {
        hits.clear();
        if (!all) {
            fetch();
            const uint64_t* row = &h_keys[(size_t)q * stride];
            for (uint32_t i = 0; i < stride && i < plan.programs[q].k && row[i]; ++i) hits.push_back(Hit{(uint32_t)(row[i] & 0xFFFFFFFFull), vbit::key_score((uint32_t)(row[i] >> 32))});
            return;
        }
        unsigned long long cnt = 0;
        VDEV_CUDA(cudaMemcpy(&cnt, d_counters.p + 4, 8, cudaMemcpyDeviceToHost));
        if (cnt > emit_capacity) throw std::runtime_error("emit buffer overflow");
        std::vector<unsigned long long> buf(cnt);
        if (cnt) VDEV_CUDA(cudaMemcpy(buf.data(), d_emit.p, cnt * 8, cudaMemcpyDeviceToHost));
        std::sort(buf.begin(), buf.end(), [](unsigned long long a, unsigned long long b) { return (uint32_t)a < (uint32_t)b; });
        for (auto v : buf) hits.push_back(Hit{(uint32_t)(v & 0xFFFFFFFFull), vbit::key_score((uint32_t)(v >> 32))});
    }

    // Step seam, device-resident: every hit of request 0 stays on the device as an anchor-sorted list (only the count is read back).
    void take_emitted(DeviceHitList& out) {
        unsigned long long cnt = 0;
        VDEV_CUDA(cudaMemcpy(&cnt, d_counters.p + 4, 8, cudaMemcpyDeviceToHost));
        if (cnt > emit_capacity) throw std::runtime_error("emit buffer overflow");
        if (cnt >= (1ull << 31)) throw vplan::Unsupported("hit lists of 2^31 entries or more");
        out.ix = ix;
        out.n = (uint32_t)cnt;
        out.nonneg = plan.programs[0].nonneg != 0 && plan.programs[0].n_boosts == 0;  // (a boost may turn a score negative)
        out.entries.alloc((size_t)cnt + 1);
        if (!cnt) return;
        DevBuf<unsigned long long> alt;
        DevBuf<unsigned char> temp;
        alt.alloc((size_t)cnt);
        const size_t temp_bytes = emitted_sort_temp_bytes((uint32_t)cnt);
        temp.alloc(temp_bytes + 1);
        VDEV_CUDA(launch_emitted_to_list(stream, d_emit.p, alt.p, out.entries.p, (uint32_t)cnt, temp.p, temp_bytes));
        VDEV_CUDA(cudaStreamSynchronize(stream));
    }
};

}  // namespace vdev
