// Step seam (one entry per PlanStep kind, src/plan_creator/plan_steps.rs:18-74) and
// facet materialisation, layered on the batch engine.
#pragma once
#include <string>
#include <vector>

#include "../../../include/veloci_b200.h"
#include "engine.hpp"
#include "explain.hpp"
#include "explain_plan.hpp"
#include "field_highlight.hpp"
#include "highlight.hpp"
#include "read_document.hpp"

namespace vsteps {

struct FacetGroup {
    uint32_t id = 0, count = 0;
    std::string text;
};
struct FacetGroups {
    std::string field;
    std::vector<FacetGroup> groups;
};

// SearchResult.facets: the device's (value id, count) groups plus their text (facet.rs:60-72)
inline void materialize_facets(vdev::Batch& b, std::vector<std::vector<FacetGroups>>& out) {
    if (!out.empty()) return;
    b.fetch();
    out.resize(b.n);
    for (uint32_t q = 0; q < b.n; ++q) {
        const vplan::RequestPlan& rp = b.plan.requests[q];
        if (rp.status != 0) continue;
        for (size_t f = 0; f < rp.facets.size(); ++f) {
            const uint32_t fi = rp.facet_begin + (uint32_t)f;
            FacetGroups fg;
            fg.field = rp.facets[f].field;
            for (uint32_t g = 0; g < b.h_facet_n[fi]; ++g) {
                FacetGroup grp;
                grp.id = b.h_facet_ids[(size_t)fi * b.facet_stride + g];
                grp.count = b.h_facet_counts[(size_t)fi * b.facet_stride + g];
                grp.text = b.ix->host->get_text_for_id(b.plan.facet_text_path[fi], grp.id);
                fg.groups.push_back(std::move(grp));
            }
            out[q].push_back(std::move(fg));
        }
    }
}

// SearchResult::why_found_terms of request q (src/search.rs:186): per dictionary path ("<field>.textindex") the texts of the
// terms its search parts matched (term_text_in_field, search_field.rs:386-389; merged over the parts, set_op.rs:49-64).
inline vhost::TermSets why_found_terms(vdev::Batch& b, uint32_t q) {
    vhost::TermSets out;
    const vdev::QueryProgram& qp = b.plan.programs[q];
    for (uint32_t l = 0; l < qp.n_leaves; ++l) {
        const uint32_t part = b.plan.leaf_part[qp.leaf_begin + l];
        if (b.plan.parts[part].flags & vdev::kPartList) continue;  // phrase / locality / boost lists are not search parts
        const std::string& path = b.plan.dict_names[b.plan.part_dict[part]];
        const vhost::TermDict& dict = b.ix->host->dict.at(path);
        std::vector<uint32_t> terms;
        std::vector<float> scores;
        b.download_matches(part, terms, scores);
        for (uint32_t id : terms) {
            size_t slot = 0;
            if (dict.find_id(id, slot)) out[path].insert(dict.term(slot));
        }
    }
    return out;
}

// term_id_hits_in_field of request q (search_field.rs:375-380, merged over the parts: set_op.rs:29-47): per dictionary path
// the ids of the terms its search parts matched -- what get_why_found highlights by when the request has `select`.
inline std::map<std::string, std::set<uint32_t>> why_found_term_ids(vdev::Batch& b, uint32_t q) {
    std::map<std::string, std::set<uint32_t>> out;
    const vdev::QueryProgram& qp = b.plan.programs[q];
    for (uint32_t l = 0; l < qp.n_leaves; ++l) {
        const uint32_t part = b.plan.leaf_part[qp.leaf_begin + l];
        if (b.plan.parts[part].flags & vdev::kPartList) continue;
        std::vector<uint32_t> terms;
        std::vector<float> scores;
        b.download_matches(part, terms, scores);
        out[b.plan.dict_names[b.plan.part_dict[part]]].insert(terms.begin(), terms.end());
    }
    return out;
}

// search::to_search_result (src/search.rs:65-110): the documents of request q's hits -- from the document store, or rebuilt
// from the indices when the request has `select` -- each with its hit and, when the request asked for why_found, the
// highlighted texts of the matched terms.
// {"num_hits": n, "execution_time_ns": t, "data": [{"doc": {..}, "hit": {"id": .., "score": ..}, "why_found": {"field": ["<b>..</b>"]}, "explain": [..]}]}
inline std::string result_docs(vdev::Batch& b, uint32_t q) {
    const vplan::RequestPlan& rp = b.plan.requests[q];
    uint64_t num_hits = 0;
    const uint32_t cap = (uint32_t)std::min<uint64_t>(rp.top, vdev::kMaxKLarge);
    std::vector<vdev::vgpu_hit_pod> hits((size_t)cap + 1);
    const uint32_t n = b.result(q, &num_hits, hits.data(), cap);
    vhost::TermSets terms;
    std::map<std::string, std::set<uint32_t>> term_ids;
    if (rp.why_found && n) {
        if (rp.select) term_ids = why_found_term_ids(b, q);  // with `select` the texts are highlighted by token ids (why_found.rs:11-49)
        else terms = why_found_terms(b, q);
    }
    // DocWithHit::explain (search.rs:86,96): the hit's explanations, null when the request did not ask or they cannot be rebuilt
    std::unique_ptr<vexplain::Explainer> explainer;
    if (rp.explain && n) {
        try {
            explainer.reset(new vexplain::Explainer(b, q));
        } catch (const vplan::Unsupported&) {
        }
    }
    // SearchResultWithDoc::execution_time_ns (search.rs:226, search_result_with_doc.rs): the requests of a batch are answered
    // together, so every one of them reports the device time of the batch's last execute
    double batch_ms = 0.0;
    for (int p = 0; p < vdev::kPhases; ++p) batch_ms += b.phase_ms[p];
    std::string out = "{\"num_hits\":" + std::to_string(num_hits) + ",\"execution_time_ns\":" + std::to_string((uint64_t)(batch_ms * 1.0e6)) + ",\"data\":[";
    for (uint32_t i = 0; i < n; ++i) {
        const std::string doc = rp.select ? vjson::to_string(vhost::read_data(*b.ix->host, hits[i].id, *rp.select)) : b.ix->host->get_doc(hits[i].id);
        if (i) out += ',';
        out += "{\"doc\":" + doc + ",\"hit\":{\"id\":" + std::to_string(hits[i].id) + ",\"score\":";
        char buf[48];
        snprintf(buf, sizeof buf, "%.9g", (double)hits[i].score);
        out += buf;
        out += "},\"why_found\":";
        if (rp.select) vhost::write_highlights(out, vhost::why_found_by_ids(*b.ix->host, hits[i].id, term_ids));
        else vhost::write_highlights(out, vhost::highlight_document(b.ix->host->metadata, doc, terms));
        if (rp.explain_asked) {
            out += ",\"explain\":";
            if (explainer && i < explainer->walk().n_anchors() && explainer->walk().anchor(i) == hits[i].id) vexplain::Walk::write_items(out, explainer->walk().explain_anchor(i));
            else out += "null";
        }
        out += '}';
    }
    return out + "]}";
}

inline vhost::SearchPart parse_part(const char* json) {
    vjson::Value v;
    try {
        v = vjson::parse(json, strlen(json));
    } catch (const vjson::ParseError& e) {
        throw vhost::RequestError(e.what());
    }
    return vhost::parse_search_part(v);
}

template <class T>
inline T* dup_array(const std::vector<T>& v) {
    T* p = (T*)malloc(std::max<size_t>(1, v.size()) * sizeof(T));
    if (!v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}

// get_term_ids_in_field (search_field.rs:277-398): fuzzy_match + scoring on the device,
// (term id, score) in ascending term-id order like the FST stream.
inline void field_search(vdev::DeviceIndex& ix, const char* part_json, bool get_scores, bool get_ids, vgpu_hitlist& out) {
    memset(&out, 0, sizeof out);
    vhost::SearchPart part = parse_part(part_json);
    vdev::Batch b;
    b.prepare_parts(&ix, {part});
    b.run_match();
    std::vector<uint32_t> terms;
    std::vector<float> scores;
    b.download_matches(0, terms, scores);
    std::vector<size_t> order(terms.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t x, size_t y) { return terms[x] < terms[y]; });
    std::vector<vgpu_hit> hits;
    std::vector<uint32_t> ids;
    for (size_t i : order) {
        if (get_scores) hits.push_back(vgpu_hit{terms[i], scores[i]});
        if (get_ids) ids.push_back(terms[i]);
    }
    out.hits = dup_array(hits), out.n_hits = (uint32_t)hits.size();
    out.ids = dup_array(ids), out.n_ids = (uint32_t)ids.size();
}

// suggest_multi / suggest (search_field.rs:147-228): every part's fuzzy match and scoring run on the device in one batch
// (get_term_ids_in_field, get_scores + return_term); the per-part top/skip bound (vdev::bound_part_hits), the merge of
// equal texts and the final order are a few hundred elements of host work.
struct Suggestion {
    std::string text;
    float score;
    uint32_t id;
};

inline std::vector<Suggestion> suggest(vdev::DeviceIndex& ix, const vhost::Request& req) {
    if (!req.suggest) throw vplan::InvalidRequest("only suggest allowed in suggest function");
    // the device matches and scores; bound, boost and token_value are applied to its output above
    std::vector<vhost::SearchPart> device_parts;
    for (const vhost::SearchPart& part : *req.suggest) {
        vhost::SearchPart p = part;
        p.top.reset(), p.skip.reset(), p.boost.reset(), p.token_value.reset();
        device_parts.push_back(std::move(p));
    }
    std::vector<Suggestion> all;
    if (!device_parts.empty()) {
        vdev::Batch b;
        std::vector<uint32_t> part_ids;
        b.prepare_parts(&ix, device_parts, &part_ids);
        b.run_match();
        for (size_t i = 0; i < device_parts.size(); ++i) {
            std::vector<vdev::TermHit> hits;
            b.download_part_hits(part_ids[i], hits);
            vdev::bound_part_hits((*req.suggest)[i], hits);
            vdev::apply_token_value(*ix.host, (*req.suggest)[i], hits);  // search_field.rs:391-395
            std::string path = device_parts[i].path;
            if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
            const vhost::TermDict& dict = ix.host->dict.at(path);
            for (const vdev::TermHit& h : hits) {
                size_t slot = 0;
                if (!dict.find_id(h.id, slot)) continue;
                all.push_back(Suggestion{vfmt::to_lowercase(dict.term(slot)), h.score, h.id});  // return_term_lowercase
            }
        }
    }
    // get_text_score_id_from_result: equal texts merge into the first of them with the largest score, then by score
    std::stable_sort(all.begin(), all.end(), [](const Suggestion& a, const Suggestion& b) { return b.text < a.text; });
    std::vector<Suggestion> merged;
    for (Suggestion& sg : all) {
        if (!merged.empty() && merged.back().text == sg.text) {
            if (sg.score > merged.back().score) merged.back().score = sg.score;
        } else {
            merged.push_back(std::move(sg));
        }
    }
    std::stable_sort(merged.begin(), merged.end(), [](const Suggestion& a, const Suggestion& b) { return a.score > b.score; });
    if (req.skip) merged.erase(merged.begin(), merged.begin() + (long)std::min<uint64_t>(*req.skip, merged.size()));  // apply_top_skip, search.rs:230-239
    if (req.top && merged.size() > *req.top) merged.resize((size_t)*req.top);
    return merged;
}

// search_field::highlight (search_field.rs:232-245): the part's terms are matched and scored on the device; the bound, the
// boost, the token values and resolve_token_hits_to_text_id with snippets follow on the host (host/field_highlight.hpp).
inline std::vector<Suggestion> highlight(vdev::DeviceIndex& ix, const char* part_json) {
    vjson::Value v;
    try {
        v = vjson::parse(part_json, strlen(part_json));
    } catch (const vjson::ParseError& e) {
        throw vhost::RequestError(e.what());
    }
    const vhost::HighlightRequest req = vhost::parse_highlight_request(v);
    vhost::SearchPart bare = req.part;
    bare.top.reset(), bare.skip.reset(), bare.boost.reset(), bare.token_value.reset();
    vdev::Batch b;
    std::vector<uint32_t> part_ids;
    b.prepare_parts(&ix, {bare}, &part_ids);
    b.run_match();
    std::vector<vdev::TermHit> hits;
    b.download_part_hits(part_ids[0], hits);
    vdev::bound_part_hits(req.part, hits);
    vdev::apply_token_value(*ix.host, req.part, hits);
    std::vector<Suggestion> out;
    for (vhost::FieldHighlight& h : vhost::highlight_field(*ix.host, req, hits)) out.push_back(Suggestion{std::move(h.text), h.score, h.id});
    return out;
}

// The remaining step entry points run a one-request batch whose leaves are explicit hit lists.
inline void run_lists(vdev::DeviceIndex& ix, const std::vector<vdev::ExplicitList>& lists, const std::vector<uint32_t>& code, const std::vector<vdev::BoostStep>& boosts, uint32_t k,
                      bool all_hits, vgpu_hitlist& out, const std::vector<uint32_t>& post = {}) {
    memset(&out, 0, sizeof out);
    vdev::Batch b;
    b.prepare_lists(&ix, lists, code, boosts, k, all_hits, nullptr, post);
    b.execute();
    std::vector<vgpu_hit> hits;
    b.download_hits(0, all_hits, hits);
    out.hits = dup_array(hits), out.n_hits = (uint32_t)hits.size();
    out.ids = dup_array(std::vector<uint32_t>()), out.n_ids = 0;
}

inline void resolve_to_anchor(vdev::DeviceIndex& ix, const char* part_json, const vgpu_hitlist& in, vgpu_hitlist& out) {
    memset(&out, 0, sizeof out);
    vhost::SearchPart part = parse_part(part_json);
    vdev::Batch b;
    std::vector<uint32_t> terms(in.n_hits);
    std::vector<float> scores(in.n_hits);
    for (uint32_t i = 0; i < in.n_hits; ++i) terms[i] = in.hits[i].id, scores[i] = in.hits[i].score;
    b.prepare_term_hits(&ix, part, terms, scores);
    b.execute();
    std::vector<vgpu_hit> hits;
    b.download_hits(0, true, hits);
    out.hits = dup_array(hits), out.n_hits = (uint32_t)hits.size();
    // hits_ids -> anchors through text_id_to_anchor (search_field.rs:468-496); small, host side
    std::vector<uint32_t> ids;
    if (in.n_ids) {
        std::string path = part.path;
        if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
        if (ix.host->is_anchor_identity_column(path)) ids.assign(in.ids, in.ids + in.n_ids);
        else {
            const vhost::KeyValueStore& t2a = ix.host->get_valueid_to_parent(path + ".text_id_to_anchor");
            for (uint32_t i = 0; i < in.n_ids; ++i) t2a.append_values(in.ids[i], ids);
        }
    }
    out.ids = dup_array(ids), out.n_ids = (uint32_t)ids.size();
}

inline std::vector<vdev::ExplicitList> to_lists(const vgpu_hitlist* inputs, uint32_t n) {
    std::vector<vdev::ExplicitList> lists(n);
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = 0; j < inputs[i].n_hits; ++j) {
            lists[i].anchors.push_back(inputs[i].hits[j].id);
            lists[i].scores.push_back(inputs[i].hits[j].score);
        }
    return lists;
}

// union_hits_score (set_op.rs:87-220) / intersect_hits_score (:368-446)
inline void set_op(vdev::DeviceIndex& ix, const vgpu_hitlist* inputs, const char* const* terms, uint32_t n, bool is_union, vgpu_hitlist& out) {
    memset(&out, 0, sizeof out);
    if (n == 0) {
        out.hits = dup_array(std::vector<vgpu_hit>()), out.ids = dup_array(std::vector<uint32_t>());
        return;
    }
    if (n > vdev::kMaxLeaves) throw vplan::Unsupported("more than " + std::to_string(vdev::kMaxLeaves) + " inputs");
    std::vector<vdev::ExplicitList> lists = to_lists(inputs, n);
    std::vector<uint32_t> code;
    for (uint32_t i = 0; i < n; ++i) code.push_back(vdev::kOpLeaf), code.push_back(i);
    if (n > 1) {
        if (is_union) {
            std::vector<std::string> ts;
            for (uint32_t i = 0; i < n; ++i) ts.push_back(terms[i] ? terms[i] : "");
            std::vector<std::string> sorted = ts;
            std::sort(sorted.begin(), sorted.end());
            sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
            code.push_back(vdev::kOpUnion), code.push_back(n), code.push_back((uint32_t)sorted.size());
            for (uint32_t i = 0; i < n; ++i) code.push_back((uint32_t)(std::find(sorted.begin(), sorted.end(), ts[i]) - sorted.begin()));
        } else {
            // first shortest input is summed last, the last input takes its place (set_op.rs:388-417)
            uint32_t shortest = 0;
            for (uint32_t i = 1; i < n; ++i)
                if (inputs[i].n_hits < inputs[shortest].n_hits) shortest = i;
            code.push_back(vdev::kOpIntersect), code.push_back(n);
            for (uint32_t i = 0; i + 1 < n; ++i) code.push_back(i == shortest ? n - 1 : i);
            code.push_back(shortest);
            for (uint32_t i = 0; i < n; ++i) code.push_back(vdev::kNoValue - 1);  // order is final: not a leaf reference
        }
    }
    run_lists(ix, lists, code, {}, 0, true, out);
}

inline std::vector<vdev::ExplicitList> ids_to_lists(const vgpu_hitlist* inputs, uint32_t n) {
    std::vector<vdev::ExplicitList> lists(n);
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = 0; j < inputs[i].n_ids; ++j) {
            lists[i].anchors.push_back(inputs[i].ids[j]);
            lists[i].scores.push_back(1.0f);
        }
    return lists;
}

inline void ids_out(const std::vector<uint32_t>& ids, vgpu_hitlist& out) {
    memset(&out, 0, sizeof out);
    out.hits = dup_array(std::vector<vgpu_hit>()), out.n_hits = 0;
    out.ids = dup_array(ids), out.n_ids = (uint32_t)ids.size();
}

// union_hits_ids (set_op.rs:222-258): one input passes through as given; otherwise the distinct ids, ascending.
// intersect_hits_ids (set_op.rs:468-510): one input passes through; otherwise the ids of the shortest input (sorted, duplicates
// kept) that every other non-empty input contains (an empty other input is skipped by the reference's iterator filter).
inline void set_op_ids(vdev::DeviceIndex& ix, const vgpu_hitlist* inputs, uint32_t n, bool is_union, vgpu_hitlist& out) {
    if (n == 0) return ids_out({}, out);
    if (n == 1) return ids_out(std::vector<uint32_t>(inputs[0].ids, inputs[0].ids + inputs[0].n_ids), out);
    if (n > vdev::kMaxLeaves) throw vplan::Unsupported("more than " + std::to_string(vdev::kMaxLeaves) + " inputs");
    std::vector<uint32_t> pick(n);
    uint32_t shortest = 0;
    for (uint32_t i = 0; i < n; ++i) {
        pick[i] = i;
        if (inputs[i].n_ids < inputs[shortest].n_ids) shortest = i;
    }
    if (!is_union) {
        if (inputs[shortest].n_ids == 0) return ids_out({}, out);
        pick.clear();
        for (uint32_t i = 0; i < n; ++i)
            if (i == shortest || inputs[i].n_ids != 0) pick.push_back(i);
    }
    std::vector<vgpu_hitlist> chosen;
    for (uint32_t i : pick) chosen.push_back(inputs[i]);
    const uint32_t m = (uint32_t)chosen.size();
    std::vector<vdev::ExplicitList> lists = ids_to_lists(chosen.data(), m);
    std::vector<uint32_t> code;
    for (uint32_t i = 0; i < m; ++i) code.push_back(vdev::kOpLeaf), code.push_back(i);
    if (m > 1) {
        if (is_union) {
            code.push_back(vdev::kOpUnion), code.push_back(m), code.push_back(m);
            for (uint32_t i = 0; i < m; ++i) code.push_back(i);
        } else {
            code.push_back(vdev::kOpIntersect), code.push_back(m);
            for (uint32_t i = 0; i < m; ++i) code.push_back(i);
            for (uint32_t i = 0; i < m; ++i) code.push_back(vdev::kNoValue - 1);
        }
    }
    vgpu_hitlist all;
    run_lists(ix, lists, code, {}, 0, true, all);  // the device's hits come back by ascending anchor id
    std::vector<uint32_t> ids;
    if (is_union) {
        for (uint32_t i = 0; i < all.n_hits; ++i) ids.push_back(all.hits[i].id);
    } else {
        std::vector<uint32_t> base(inputs[shortest].ids, inputs[shortest].ids + inputs[shortest].n_ids);
        std::sort(base.begin(), base.end());
        for (uint32_t id : base) {
            const vgpu_hit* b0 = all.hits;
            const vgpu_hit* e = all.hits + all.n_hits;
            const vgpu_hit* it = std::lower_bound(b0, e, id, [](const vgpu_hit& h, uint32_t v) { return h.id < v; });
            if (it != e && it->id == id) ids.push_back(id);
        }
    }
    vgpu_hitlist_free(&all);
    ids_out(ids, out);
}

// intersect_score_hits_with_ids (set_op.rs:311-326, the IntersectScoresWithIds step plan_steps.rs:330-345): the scored hits
// whose id is in `ids`, by ascending id; with no ids at all the reference keeps every hit.
inline void intersect_scores_with_ids(vdev::DeviceIndex& ix, const vgpu_hitlist& scores, const vgpu_hitlist& ids, vgpu_hitlist& out) {
    std::vector<vdev::ExplicitList> lists = to_lists(&scores, 1);
    if (ids.n_ids == 0) return run_lists(ix, lists, {vdev::kOpLeaf, 0u}, {}, 0, true, out);
    lists.push_back(ids_to_lists(&ids, 1)[0]);
    run_lists(ix, lists, {vdev::kOpLeaf, 0u, vdev::kOpLeaf, 1u, vdev::kOpFilter}, {}, 0, true, out);
}

// resolve_token_to_anchor with a FilterResult::Set (search_field.rs:423, 540-548): anchors outside the filter are skipped.
inline void resolve_to_anchor_filtered(vdev::DeviceIndex& ix, const char* part_json, const vgpu_hitlist& in, const uint32_t* filter, uint32_t n_filter, vgpu_hitlist& out) {
    vgpu_hitlist all;
    resolve_to_anchor(ix, part_json, in, all);
    vgpu_hitlist f;
    memset(&f, 0, sizeof f);
    f.ids = const_cast<uint32_t*>(filter), f.n_ids = n_filter;
    memset(&out, 0, sizeof out);
    if (n_filter == 0) {  // an empty set contains nothing
        out.hits = dup_array(std::vector<vgpu_hit>()), out.ids = all.ids, out.n_ids = all.n_ids;
        all.ids = nullptr;
    } else {
        intersect_scores_with_ids(ix, all, f, out);
        free(out.ids);
        out.ids = all.ids, out.n_ids = all.n_ids, all.ids = nullptr;
    }
    vgpu_hitlist_free(&all);
}

// get_facet (facet.rs:31-73) over the given hit ids: (text, count, value id) groups, count desc then value id asc.
inline std::vector<FacetGroup> facet(vdev::DeviceIndex& ix, const char* facet_json, const uint32_t* ids, uint32_t n_ids) {
    vjson::Value v;
    try {
        v = vjson::parse(facet_json, strlen(facet_json));
    } catch (const vjson::ParseError& e) {
        throw vhost::RequestError(e.what());
    }
    if (!v.is_object()) throw vhost::RequestError("facet must be an object");
    vhost::FacetRequest fr;
    const vjson::Value* fld = v.get("field");
    if (!fld || !fld->is_string()) throw vhost::RequestError("missing field `field`");
    fr.field = fld->str;
    if (const vjson::Value* t = v.get("top")) {
        if (t->is_null()) fr.top.reset();
        else if (t->is_number() && t->num >= 0) fr.top = (uint64_t)t->num;
        else throw vhost::RequestError("top must be an unsigned integer");
    }
    vdev::Batch b;
    vdev::ExplicitList l;
    l.anchors.assign(ids, ids + n_ids);
    l.scores.assign(n_ids, 1.0f);
    b.prepare_lists(&ix, {l}, {vdev::kOpLeaf, 0u}, {}, 0, false, &fr);
    b.execute();
    std::vector<std::vector<FacetGroups>> groups;
    materialize_facets(b, groups);
    return groups.at(0).at(0).groups;
}

// PlanStepPhrasePairToAnchorId (plan_steps.rs:279-293) = get_anchor_for_phrases_in_field (search_field.rs:263-275): the anchors
// of every (term id of the first part, term id of the second part) pair of the field's phrase-pair store, concatenated and
// sorted (duplicates stay).  The lookups and the gather run on the device; the final sort of the (small) list is host work.
inline void phrase_pairs_to_anchor(vdev::DeviceIndex& ix, const char* path_c, const uint32_t* ids1, uint32_t n1, const uint32_t* ids2, uint32_t n2, vgpu_hitlist& out) {
    std::string path = path_c;
    if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
    if (!vfmt::ends_with(path, ".phrase_pair_to_anchor")) path += ".phrase_pair_to_anchor";
    auto it = ix.phrases.find(path);
    if (it == ix.phrases.end()) ix.host->path_not_found(path);
    std::vector<uint32_t> anchors;
    const uint64_t pairs = (uint64_t)n1 * n2;
    if (pairs > (1ull << 26)) throw vplan::Unsupported("more than 2^26 term pairs in one phrase step");
    if (pairs) {
        VDEV_CUDA(cudaSetDevice(ix.device));
        vdev::DevBuf<uint32_t> d1, d2, d_count, d_off, d_out;
        d1.upload(std::vector<uint32_t>(ids1, ids1 + n1)), d2.upload(std::vector<uint32_t>(ids2, ids2 + n2));
        d_count.alloc((size_t)pairs);
        const vdev::PhraseView store = it->second.view();
        vdev::launch_phrase_lookup(cudaStreamLegacy, store, d1.p, n1, d2.p, n2, d_count.p, nullptr, nullptr);
        std::vector<uint32_t> count((size_t)pairs), off((size_t)pairs);
        VDEV_CUDA(cudaMemcpy(count.data(), d_count.p, (size_t)pairs * 4, cudaMemcpyDeviceToHost));
        uint64_t total = 0;
        for (size_t x = 0; x < count.size(); ++x) off[x] = (uint32_t)total, total += count[x];
        if (total > 0xFFFFFFF0ull) throw vplan::Unsupported("more than 2^32 anchors in one phrase step");
        if (total) {
            d_off.upload(off);
            d_out.alloc((size_t)total);
            vdev::launch_phrase_lookup(cudaStreamLegacy, store, d1.p, n1, d2.p, n2, nullptr, d_off.p, d_out.p);
            anchors.resize((size_t)total);
            VDEV_CUDA(cudaMemcpy(anchors.data(), d_out.p, (size_t)total * 4, cudaMemcpyDeviceToHost));
        }
        VDEV_CUDA(cudaGetLastError());
        std::sort(anchors.begin(), anchors.end());
    }
    ids_out(anchors, out);
}

// BoostAnchorFromPhraseResults (plan_steps.rs:260-277): the phrase results of one phrase (same `group`) are merged into one id
// set (kmerge + dedup, sort_and_group_boosts_by_phrase_terms :230-257), every set boosts the hits it contains by 5.0
// (boost_hits_ids_vec_multi, boost.rs:149-195: a hit in two sets is multiplied twice); the hits come back by ascending id.
inline void boost_anchor_from_phrase_results(vdev::DeviceIndex& ix, const vgpu_hitlist& hits, const vgpu_hitlist* phrase_results, const uint32_t* group, uint32_t n, vgpu_hitlist& out) {
    std::vector<uint32_t> groups(group, group + n);
    std::sort(groups.begin(), groups.end());
    groups.erase(std::unique(groups.begin(), groups.end()), groups.end());
    if (groups.size() + 1 > vdev::kMaxLeaves) throw vplan::Unsupported("more than " + std::to_string(vdev::kMaxLeaves - 1) + " phrases in one boost step");
    std::vector<vdev::ExplicitList> lists = to_lists(&hits, 1);
    std::vector<uint32_t> post;
    const float boost = 5.0f;  // plan_steps.rs:270
    uint32_t boost_bits;
    memcpy(&boost_bits, &boost, 4);
    for (uint32_t g : groups) {
        std::vector<uint32_t> ids;
        for (uint32_t i = 0; i < n; ++i)
            if (group[i] == g) ids.insert(ids.end(), phrase_results[i].ids, phrase_results[i].ids + phrase_results[i].n_ids);
        std::sort(ids.begin(), ids.end());
        ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
        vdev::ExplicitList l;
        l.anchors = std::move(ids);
        l.scores.assign(l.anchors.size(), 1.0f);
        post.push_back(vdev::kPostMulIfPresent), post.push_back((uint32_t)lists.size()), post.push_back(boost_bits);
        lists.push_back(std::move(l));
    }
    run_lists(ix, lists, {vdev::kOpLeaf, 0u}, {}, 0, true, out, post);
}

// The values of every id in a CSR store, by two launches of csr_expand_kernel (count, then fill); out[off[i] .. off[i + 1])
// belong to ids[i].  Step-seam helper: small lists, legacy stream, synchronous copies.
inline std::vector<uint32_t> expand_on_device(vdev::DeviceIndex& ix, const vdev::CsrDev& csr, const std::vector<uint32_t>& ids, bool self_if_empty, std::vector<uint32_t>* off_out = nullptr) {
    std::vector<uint32_t> result, off(ids.size() + 1, 0);
    if (!ids.empty()) {
        if (ids.size() > 0x7FFFFFFFull) throw vplan::Unsupported("too many ids in one step");
        VDEV_CUDA(cudaSetDevice(ix.device));
        const uint32_t n = (uint32_t)ids.size();
        vdev::DevBuf<uint32_t> d_ids, d_count, d_off, d_out;
        d_ids.upload(ids);
        d_count.alloc(n);
        vdev::launch_csr_expand(cudaStreamLegacy, csr.view(), d_ids.p, n, self_if_empty ? 1u : 0u, d_count.p, nullptr, nullptr);
        std::vector<uint32_t> count(n);
        VDEV_CUDA(cudaMemcpy(count.data(), d_count.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
        uint64_t total = 0;
        for (uint32_t i = 0; i < n; ++i) off[i] = (uint32_t)total, total += count[i];
        if (total > 0xFFFFFFF0ull) throw vplan::Unsupported("more than 2^32 values in one step");
        off[n] = (uint32_t)total;
        if (total) {
            d_off.upload(off);
            d_out.alloc((size_t)total);
            vdev::launch_csr_expand(cudaStreamLegacy, csr.view(), d_ids.p, n, self_if_empty ? 1u : 0u, nullptr, d_off.p, d_out.p);
            result.resize((size_t)total);
            VDEV_CUDA(cudaMemcpy(result.data(), d_out.p, (size_t)total * 4, cudaMemcpyDeviceToHost));
        }
        VDEV_CUDA(cudaGetLastError());
    }
    if (off_out) *off_out = std::move(off);
    return result;
}

// boost_text_locality (boost.rs:34-87) of one field, reduced like boost_text_locality_all (boost.rs:11-32): `term_hits[t].ids`
// are the token ids query term t matched in the field.  Every token's text ids (tokens_to_text_id), all terms together; a
// text reached c > 1 times boosts its anchors by 2 c c; an anchor reached through several texts keeps the SMALLEST boost
// (the reference's max_by has its comparator reversed).  out->hits: (anchor, boost) by ascending anchor.  Nothing for a
// single term.
inline void text_locality(vdev::DeviceIndex& ix, const char* path_c, const vgpu_hitlist* term_hits, uint32_t n_terms, vgpu_hitlist& out) {
    std::string path = path_c;
    if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
    std::vector<vgpu_hit> boosts;
    if (n_terms > 1) {
        auto store = [&](const std::string& p) -> const vdev::CsrDev& {
            auto it = ix.stores.find(p);
            if (it == ix.stores.end()) ix.host->path_not_found(p);
            return it->second;
        };
        std::vector<uint32_t> tokens;
        for (uint32_t t = 0; t < n_terms; ++t) tokens.insert(tokens.end(), term_hits[t].ids, term_hits[t].ids + term_hits[t].n_ids);
        std::vector<uint32_t> texts = expand_on_device(ix, store(path + ".tokens_to_text_id"), tokens, false);
        std::sort(texts.begin(), texts.end());
        std::vector<uint32_t> boosted_texts;
        std::vector<float> boost_of;
        for (size_t i = 0; i < texts.size();) {
            size_t j = i;
            while (j < texts.size() && texts[j] == texts[i]) ++j;
            if (j - i > 1) boosted_texts.push_back(texts[i]), boost_of.push_back(2.0f * (float)(j - i) * (float)(j - i));
            i = j;
        }
        if (ix.host->is_anchor_identity_column(path)) {
            for (size_t i = 0; i < boosted_texts.size(); ++i) boosts.push_back(vgpu_hit{boosted_texts[i], boost_of[i]});
        } else {
            std::vector<uint32_t> off;
            const std::vector<uint32_t> anchors = expand_on_device(ix, store(path + ".text_id_to_anchor"), boosted_texts, false, &off);
            for (size_t i = 0; i < boosted_texts.size(); ++i)
                for (uint32_t k = off[i]; k < off[i + 1]; ++k) boosts.push_back(vgpu_hit{anchors[k], boost_of[i]});
        }
        std::sort(boosts.begin(), boosts.end(), [](const vgpu_hit& a, const vgpu_hit& b) { return a.id != b.id ? a.id < b.id : a.score < b.score; });
        boosts.erase(std::unique(boosts.begin(), boosts.end(), [](const vgpu_hit& a, const vgpu_hit& b) { return a.id == b.id; }), boosts.end());  // the smallest per anchor
    }
    memset(&out, 0, sizeof out);
    out.hits = dup_array(boosts), out.n_hits = (uint32_t)boosts.size();
    out.ids = dup_array(std::vector<uint32_t>()), out.n_ids = 0;
}

inline vhost::BoostPart parse_boost(const char* boost_json) {
    vjson::Value v;
    try {
        v = vjson::parse(boost_json, strlen(boost_json));
    } catch (const vjson::ParseError& e) {
        throw vhost::RequestError(e.what());
    }
    return vhost::parse_boost_part(v);
}

// BoostToAnchor (plan_steps.rs:174-196): the part's term hits -> text ids (resolve_token_hits_to_text_id_ids_only,
// search_field.rs:640-687: tokenized fields only; untokenized ones pass the hits_ids through) -> the value ids those texts
// belong to (join_to_parent_ids, search.rs:281-315) -> the ones with a boost value, as (anchor, value) in value-id order
// (get_boost_ids_and_resolve_to_anchor, boost.rs:432-468).  Every join and the value lookup is a kernel over the id list;
// the sort + dedup between them is host work on the (small) lists.  out->hits = boost_ids.
inline void boost_to_anchor(vdev::DeviceIndex& ix, const char* part_json, const vgpu_hitlist& in, const char* boost_json, vgpu_hitlist& out) {
    const vhost::SearchPart part = parse_part(part_json);
    const vhost::BoostPart boost = parse_boost(boost_json);
    std::string path = part.path;
    if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
    auto store = [&](const std::string& p) -> const vdev::CsrDev& {
        auto it = ix.stores.find(p);
        if (it == ix.stores.end()) ix.host->path_not_found(p);
        return it->second;
    };
    VDEV_CUDA(cudaSetDevice(ix.device));
    auto expand = [&](const vdev::CsrDev& csr, const std::vector<uint32_t>& ids, bool self_if_empty) {
        std::vector<uint32_t> result = expand_on_device(ix, csr, ids, self_if_empty);
        std::sort(result.begin(), result.end());
        result.erase(std::unique(result.begin(), result.end()), result.end());
        return result;
    };
    std::vector<uint32_t> text_ids;
    if (ix.host->is_tokenized(path)) {
        std::vector<uint32_t> tokens;
        for (uint32_t i = 0; i < in.n_hits; ++i) tokens.push_back(in.hits[i].id);
        text_ids = expand(store(path + ".tokens_to_text_id"), tokens, true);
    } else {
        text_ids.assign(in.ids, in.ids + in.n_ids);
    }
    std::vector<uint32_t> value_ids = expand(store(path + ".value_id_to_parent"), text_ids, false);
    auto col = ix.boosts.find(boost.path + ".boost_valid_to_value");
    if (col == ix.boosts.end()) ix.host->path_not_found(boost.path + ".boost_valid_to_value");
    const vdev::CsrDev& v2a = store(boost.path + ".value_id_to_anchor");
    std::vector<vgpu_hit> pairs;
    if (!value_ids.empty()) {
        const uint32_t n = (uint32_t)value_ids.size();
        vdev::DevBuf<uint32_t> d_ids, d_anchor, d_bits;
        d_ids.upload(value_ids);
        d_anchor.alloc(n), d_bits.alloc(n);
        vdev::launch_boost_values(cudaStreamLegacy, col->second.bits.p, (uint32_t)col->second.n, v2a.view(), d_ids.p, n, d_anchor.p, d_bits.p);
        std::vector<uint32_t> anchor(n), bits(n);
        VDEV_CUDA(cudaMemcpy(anchor.data(), d_anchor.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
        VDEV_CUDA(cudaMemcpy(bits.data(), d_bits.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < n; ++i) {
            if (anchor[i] == vdev::kNoValue) continue;
            float v;
            memcpy(&v, &bits[i], 4);
            pairs.push_back(vgpu_hit{anchor[i], v});
        }
    }
    VDEV_CUDA(cudaGetLastError());
    memset(&out, 0, sizeof out);
    out.hits = dup_array(pairs), out.n_hits = (uint32_t)pairs.size();
    out.ids = dup_array(std::vector<uint32_t>()), out.n_ids = 0;
}

// ApplyAnchorBoost (plan_steps.rs:198-217) = apply_boost_values_anchor (boost.rs:255-281): the hits of a part (ascending
// anchors) and the (anchor, value) list BoostToAnchor made for the same part (ascending anchors, an anchor may have several
// values).  The reference's merge walk makes the number of values a hit takes depend on the run of boosted hits before it;
// the tile kernel's kOpLeafBoost reproduces that (apply_leaf_boost, tiles.cu).  The values are given, not looked up: they
// travel as a temporary column indexed by their position in the list.  out->hits: the hits by ascending anchor.
inline void apply_anchor_boost(vdev::DeviceIndex& ix, const char* boost_json, const vgpu_hitlist& hits, const vgpu_hitlist& boost_ids, vgpu_hitlist& out) {
    const vhost::BoostPart boost = parse_boost(boost_json);
    std::vector<vdev::ExplicitList> lists = to_lists(&hits, 1);
    if (boost_ids.n_hits == 0) return run_lists(ix, lists, {vdev::kOpLeaf, 0u}, {}, 0, true, out);
    if (boost_ids.n_hits >= 0x7FFFFFF0u) throw vplan::Unsupported("too many boost values in one ApplyAnchorBoost step");
    std::vector<uint32_t> sorted_hits = lists[0].anchors;
    std::sort(sorted_hits.begin(), sorted_hits.end());
    vdev::ExplicitList values;
    values.part_flags = vdev::kPartList | vdev::kPartListBoost;
    std::vector<uint32_t> column(boost_ids.n_hits);
    for (uint32_t i = 0; i < boost_ids.n_hits; ++i) {
        if (i && boost_ids.hits[i].id < boost_ids.hits[i - 1].id) throw vplan::Unsupported("ApplyAnchorBoost: the boost values must come by ascending anchor (as BoostToAnchor returns them)");
        if (!std::binary_search(sorted_hits.begin(), sorted_hits.end(), boost_ids.hits[i].id))
            throw vplan::Unsupported("ApplyAnchorBoost: a boost value for an anchor that is not among the hits (BoostToAnchor of the same part never returns one)");
        values.anchors.push_back(boost_ids.hits[i].id);
        values.scores.push_back(1.0f);
        values.raw_keys.push_back(0x7FFFFFFFu - i);
        memcpy(&column[i], &boost_ids.hits[i].score, 4);
    }
    lists.push_back(std::move(values));
    VDEV_CUDA(cudaSetDevice(ix.device));
    vdev::DevBuf<uint32_t> d_column;
    d_column.upload(column);
    vdev::BoostStep step;
    memset(&step, 0, sizeof step);
    step.column = d_column.p, step.n = (uint32_t)column.size();
    step.fun = (uint32_t)boost.boost_fun;
    step.param = boost.param.value_or(0.0f);
    step.list_only = 1;
    if (boost.expression) vplan::parse_expression(*boost.expression, step);
    run_lists(ix, lists, {vdev::kOpLeafBoost, 0u, 1u, 0u}, {step}, 0, true, out);
}

inline void add_boost(vdev::DeviceIndex& ix, const char* boost_json, vgpu_hitlist& inout) {
    vjson::Value v;
    try {
        v = vjson::parse(boost_json, strlen(boost_json));
    } catch (const vjson::ParseError& e) {
        throw vhost::RequestError(e.what());
    }
    vhost::BoostPart bp = vhost::parse_boost_part(v);
    vplan::BatchPlan plan;
    plan.ix = &ix;
    vdev::BoostStep step = plan.make_boost(bp);
    std::vector<vdev::ExplicitList> lists = to_lists(&inout, 1);
    vgpu_hitlist out;
    run_lists(ix, lists, {vdev::kOpLeaf, 0u}, {step}, 0, true, out);
    // add_boost keeps the order of its input; the engine returns hits by anchor id
    std::vector<std::pair<uint32_t, float>> by_id;
    for (uint32_t i = 0; i < out.n_hits; ++i) by_id.emplace_back(out.hits[i].id, out.hits[i].score);
    for (uint32_t i = 0; i < inout.n_hits; ++i) {
        auto it = std::lower_bound(by_id.begin(), by_id.end(), std::make_pair(inout.hits[i].id, -INFINITY));
        if (it != by_id.end() && it->first == inout.hits[i].id) inout.hits[i].score = it->second;
    }
    vgpu_hitlist_free(&out);
}

// top_n_sort + apply_top_skip (sort.rs:5-22, search.rs:230-239)
inline void top_n(vdev::DeviceIndex& ix, const vgpu_hitlist& in, uint32_t top, uint32_t skip, vgpu_hitlist& out) {
    if ((uint64_t)top + skip > (ix.n_shards > 1 ? vdev::kMaxK : vdev::kMaxKLarge)) throw vplan::Unsupported("top + skip above " + std::to_string(ix.n_shards > 1 ? vdev::kMaxK : vdev::kMaxKLarge));
    std::vector<vdev::ExplicitList> lists = to_lists(&in, 1);
    vgpu_hitlist all;
    run_lists(ix, lists, {vdev::kOpLeaf, 0u}, {}, top + skip, false, all);
    std::vector<vgpu_hit> hits;
    for (uint32_t i = skip; i < all.n_hits && hits.size() < top; ++i) hits.push_back(all.hits[i]);
    vgpu_hitlist_free(&all);
    memset(&out, 0, sizeof out);
    out.hits = dup_array(hits), out.n_hits = (uint32_t)hits.size();
    out.ids = dup_array(std::vector<uint32_t>()), out.n_ids = 0;
}

// ---- device-resident step seam ("device-resident handles between steps") -------------------------------
// The same steps over hit lists that stay on the device between them (vdev::DeviceHitList: anchor-sorted (anchor, score
// key) entries, consumed by the tile path like the postings of a rarely matched term).  Only sizes cross the bus between
// two steps; the chain ResolveTokenIdToAnchor -> Union / Intersect -> BoostPlanStepFromBoostRequest -> top_n runs from
// the matched terms to the k best hits without its intermediate lists leaving HBM.

inline vdev::ExplicitList dev_leaf(vdev::DeviceIndex& ix, const vdev::DeviceHitList& h) {
    if (h.ix != &ix) throw vplan::InvalidRequest("hit list handle of another index");
    vdev::ExplicitList l;
    l.dev = h.data(), l.dev_n = h.n, l.dev_nonneg = h.nonneg;
    return l;
}

inline void run_lists_dev(vdev::DeviceIndex& ix, const std::vector<vdev::ExplicitList>& lists, const std::vector<uint32_t>& code, const std::vector<vdev::BoostStep>& boosts, vdev::DeviceHitList& out) {
    vdev::Batch b;
    b.prepare_lists(&ix, lists, code, boosts, 0, true);
    b.execute();
    b.take_emitted(out);
}

// host list -> handle: by anchor id, a repeated anchor keeps its largest score (resolve_token_to_anchor's dedup), anchors
// outside the handle's shard are dropped
inline void dev_upload(vdev::DeviceIndex& ix, const vgpu_hitlist& in, vdev::DeviceHitList& out) {
    std::vector<std::pair<uint32_t, float>> hits;
    for (uint32_t i = 0; i < in.n_hits; ++i)
        if (in.hits[i].id >= ix.anchor_lo && in.hits[i].id < ix.anchor_hi) hits.emplace_back(in.hits[i].id, in.hits[i].score);
    std::stable_sort(hits.begin(), hits.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    std::vector<unsigned long long> entries;
    out.nonneg = true;
    for (size_t i = 0; i < hits.size(); ++i) {
        float score = hits[i].second;
        while (i + 1 < hits.size() && hits[i + 1].first == hits[i].first) score = std::max(score, hits[++i].second);
        if (!(score >= 0.0f)) out.nonneg = false;
        uint32_t key = vbit::score_key(score);
        if (key == 0) key = 1;
        entries.push_back(((unsigned long long)key << 32) | hits[i].first);  // SparseEntry {anchor, key}, little endian
    }
    VDEV_CUDA(cudaSetDevice(ix.device));
    out.ix = &ix;
    out.n = (uint32_t)entries.size();
    out.entries.upload(entries);
}

inline void dev_download(const vdev::DeviceHitList& h, vgpu_hitlist& out) {
    memset(&out, 0, sizeof out);
    std::vector<unsigned long long> entries(h.n);
    if (h.n) {
        VDEV_CUDA(cudaSetDevice(h.ix->device));
        VDEV_CUDA(cudaMemcpy(entries.data(), h.entries.p, (size_t)h.n * 8, cudaMemcpyDeviceToHost));
    }
    std::vector<vgpu_hit> hits;
    for (unsigned long long v : entries) hits.push_back(vgpu_hit{(uint32_t)(v & 0xFFFFFFFFull), vbit::key_score((uint32_t)(v >> 32))});
    out.hits = dup_array(hits), out.n_hits = (uint32_t)hits.size();
    out.ids = dup_array(std::vector<uint32_t>()), out.n_ids = 0;
}

// ResolveTokenIdToAnchor (search_field.rs:400-464) from the few (term id, score) hits of a part to a list on the device
inline void dev_resolve_to_anchor(vdev::DeviceIndex& ix, const char* part_json, const vgpu_hitlist& in, vdev::DeviceHitList& out) {
    vhost::SearchPart part = parse_part(part_json);
    std::vector<uint32_t> terms(in.n_hits);
    std::vector<float> scores(in.n_hits);
    for (uint32_t i = 0; i < in.n_hits; ++i) terms[i] = in.hits[i].id, scores[i] = in.hits[i].score;
    vdev::Batch b;
    b.prepare_term_hits(&ix, part, terms, scores);
    b.execute();
    b.take_emitted(out);
}

// Union / Intersect (set_op.rs:87-220, :368-446) over handles; the inputs' lengths decide the intersect's order as in the reference
inline void dev_set_op(vdev::DeviceIndex& ix, const vdev::DeviceHitList* const* inputs, const char* const* terms, uint32_t n, bool is_union, vdev::DeviceHitList& out) {
    if (n == 0) {
        VDEV_CUDA(cudaSetDevice(ix.device));
        out.ix = &ix, out.n = 0, out.nonneg = true;
        out.entries.alloc(1);
        return;
    }
    if (n > vdev::kMaxLeaves) throw vplan::Unsupported("more than " + std::to_string(vdev::kMaxLeaves) + " inputs");
    std::vector<vdev::ExplicitList> lists;
    for (uint32_t i = 0; i < n; ++i) lists.push_back(dev_leaf(ix, *inputs[i]));
    std::vector<uint32_t> code;
    for (uint32_t i = 0; i < n; ++i) code.push_back(vdev::kOpLeaf), code.push_back(i);
    if (n > 1) {
        if (is_union) {
            std::vector<std::string> ts;
            for (uint32_t i = 0; i < n; ++i) ts.push_back(terms && terms[i] ? terms[i] : "");
            std::vector<std::string> sorted = ts;
            std::sort(sorted.begin(), sorted.end());
            sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
            code.push_back(vdev::kOpUnion), code.push_back(n), code.push_back((uint32_t)sorted.size());
            for (uint32_t i = 0; i < n; ++i) code.push_back((uint32_t)(std::find(sorted.begin(), sorted.end(), ts[i]) - sorted.begin()));
        } else {
            uint32_t shortest = 0;
            for (uint32_t i = 1; i < n; ++i)
                if (inputs[i]->n < inputs[shortest]->n) shortest = i;
            code.push_back(vdev::kOpIntersect), code.push_back(n);
            for (uint32_t i = 0; i + 1 < n; ++i) code.push_back(i == shortest ? n - 1 : i);
            code.push_back(shortest);
            for (uint32_t i = 0; i < n; ++i) code.push_back(vdev::kNoValue - 1);  // order is final: not a leaf reference
        }
    }
    run_lists_dev(ix, lists, code, {}, out);
}

// BoostPlanStepFromBoostRequest (add_boost, boost.rs:470-504)
inline void dev_add_boost(vdev::DeviceIndex& ix, const char* boost_json, const vdev::DeviceHitList& in, vdev::DeviceHitList& out) {
    vhost::BoostPart bp = parse_boost(boost_json);
    vplan::BatchPlan plan;
    plan.ix = &ix;
    vdev::BoostStep step = plan.make_boost(bp);
    run_lists_dev(ix, {dev_leaf(ix, in)}, {vdev::kOpLeaf, 0u}, {step}, out);
}

// top_n_sort + apply_top_skip (sort.rs:5-22, search.rs:230-239): the k best hits of a handle, to the host
inline void dev_top_n(vdev::DeviceIndex& ix, const vdev::DeviceHitList& in, uint32_t top, uint32_t skip, vgpu_hitlist& out) {
    if ((uint64_t)top + skip > (ix.n_shards > 1 ? vdev::kMaxK : vdev::kMaxKLarge)) throw vplan::Unsupported("top + skip above " + std::to_string(ix.n_shards > 1 ? vdev::kMaxK : vdev::kMaxKLarge));
    vdev::Batch b;
    b.prepare_lists(&ix, {dev_leaf(ix, in)}, {vdev::kOpLeaf, 0u}, {}, top + skip, false);
    b.execute();
    std::vector<vgpu_hit> all, hits;
    b.download_hits(0, false, all);
    for (uint32_t i = skip; i < all.size() && hits.size() < top; ++i) hits.push_back(all[i]);
    memset(&out, 0, sizeof out);
    out.hits = dup_array(hits), out.n_hits = (uint32_t)hits.size();
    out.ids = dup_array(std::vector<uint32_t>()), out.n_ids = 0;
}

}  // namespace vsteps
