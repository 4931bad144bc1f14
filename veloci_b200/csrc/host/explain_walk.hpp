// `explain` (src/search/result/explain.rs:1-21): for the hits a request returns, how their scores came about, as the
// reference's plan steps record it while they run -- FieldSearch (search_field.rs:334-344), ResolveTokenIdToAnchor
// (:429-441), Union (set_op.rs:120-137,187-208), Intersect (:384,421-432), the anchor-level request boosts and the
// token_value boost of a part (boost.rs:283-377,470-504).  The reference carries a map over *all* hits through the plan and
// reads the returned ones at the end (search.rs:174, :86); here the plan runs on the device without such a map, and the
// explanations of the k returned anchors are rebuilt afterwards: given the parts' matched terms (a device match) and the
// posting weight of every (matched term, returned anchor) pair (posting_lookup_kernel), the tree is walked per anchor in the
// reference's order of operations (f32, same order of additions), a few dozen values per request.
//
// This header is the walk: host arithmetic only, fed by host/explain.hpp on the device and by the index helper library in
// the CPU tests (vidx_explain_walk, against the oracle's explain).
//
// Outside this reconstruction (Unsupported for the explanation; the search itself is not affected): 1:n boosts and phrase
// boosts (what the reference records for them depends on the merge walk over the whole hit list, boost.rs:197-281).
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "part_hits.hpp"
#include "persistence.hpp"
#include "request.hpp"

namespace vexplain {

struct Item {  // one Explain value
    enum Kind { Boost, TermToAnchor, LevenshteinScore, OrSumOverDistinctTerms } kind;
    float a = 0.f, b = 0.f, c = 0.f;  // Boost / OrSum: a.  TermToAnchor: term_score, anchor_score, final_score.  LevenshteinScore: score
    uint32_t term_id = 0;
    std::string text;
    explicit Item(Kind k, float value = 0.f) : kind(k), a(value) {}
};
typedef std::vector<Item> Items;

inline bool wants(const vhost::SearchPart& p) { return vhost::part_explains(p); }

inline void propagate(vhost::SearchRequest& r) {  // execution_plan.rs:46-90
    if (r.kind == vhost::SearchRequest::Search) r.part.options.present = true, r.part.options.explain = true;
    for (auto& q : r.queries) propagate(q);
}

// One search part of the tree: its final term hits in hits_scores order, what FieldSearch recorded per term, the posting
// weight of every (hit, anchor) pair.
struct Leaf {
    std::vector<vdev::TermHit> hits;
    std::map<uint32_t, Items> term_items;
    std::vector<float> weight;  // [hits.size() * n_anchors], -1: no posting
    uint64_t est = 0;           // postings of the matched terms: the length estimate the engine orders `and` inputs by (fuzzy.cu: finalize_programs_kernel)
};

struct Node {  // a sub-result of the tree, seen from one anchor
    bool present = false;
    float score = 0.f;
    Items items;
    const vhost::SearchPart* request = nullptr;  // SearchFieldResult::request: the part whose settings the sub-result carries on
    bool is_leaf = false;
    uint64_t est = 0;
};

class Walk {
  public:
    Walk(const vhost::Persistence& host, const vhost::Request& request, std::vector<uint32_t> anchors) : host_(host), request_(request), anchors_(std::move(anchors)) {
        if (!request_.search_req) throw vplan::InvalidRequest("search_req is None, but is required in search");
        if (request_.phrase_boosts && !request_.phrase_boosts->empty()) throw vplan::Unsupported("explain of phrase boosts is outside the accelerated path");
        if (request_.explain) {
            propagate(*request_.search_req);
            if (request_.filter) propagate(*request_.filter);
        }
        if (request_.boost)
            for (const vhost::BoostPart& bp : *request_.boost)
                if (bp.path.find("[]") != std::string::npos) throw vplan::Unsupported("explain of 1:n boosts is outside the accelerated path");
        check_no_tree_boosts(*request_.search_req);
        collect(*request_.search_req);
        leaves_.resize(parts_.size());
    }

    // the tree's search parts, in tree order; part(i) with top / skip / boost / token_value / options taken off is what the
    // matcher is asked for
    size_t n_parts() const { return parts_.size(); }
    const vhost::SearchPart& part(size_t i) const { return *parts_[i]; }
    vhost::SearchPart bare_part(size_t i) const {
        vhost::SearchPart d = *parts_[i];
        d.top.reset(), d.skip.reset(), d.boost.reset(), d.token_value.reset();
        d.options = vhost::SearchOptions();
        return d;
    }

    // get_term_ids_in_field's tail for part i over the bare part's (term id, score) hits in ascending term id order: what
    // FieldSearch records (search_field.rs:334-344: the score before the part's boost), the per-part bound, the part boost,
    // the token_value boost.  Returns the part's final hits, in hits_scores order: the caller supplies their weights next.
    const std::vector<vdev::TermHit>& set_hits(size_t i, std::vector<vdev::TermHit> raw) {
        const vhost::SearchPart& part = *parts_[i];
        Leaf& leaf = leaves_[i];
        leaf.hits = std::move(raw);
        std::string path = part.path;
        if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
        auto dict = host_.dict.find(path);
        if (dict == host_.dict.end()) throw vhost::FstNotFound(path);
        if (wants(part))
            for (const vdev::TermHit& h : leaf.hits) {
                Item e(Item::LevenshteinScore, h.score);
                e.term_id = h.id;
                size_t slot = 0;
                if (dict->second.find_id(h.id, slot)) e.text = dict->second.term(slot);
                leaf.term_items[h.id] = {e};
            }
        vdev::bound_part_hits(part, leaf.hits);
        std::map<uint32_t, std::vector<float>> boost_log;  // add_boost on the term hits records under the term id (boost.rs:484)
        vdev::apply_token_value(host_, part, leaf.hits, wants(part) ? &boost_log : nullptr);
        for (auto& kv : boost_log)
            for (float f : kv.second) leaf.term_items[kv.first].push_back(Item(Item::Boost, f));
        return leaf.hits;
    }
    // weight[t * n_anchors + a]: posting weight of the part's t-th final hit on the a-th anchor, -1 without a posting;
    // `postings`: the lengths of the hits' posting lists added up
    void set_weights(size_t i, std::vector<float> weight, uint64_t postings) {
        if (weight.size() != leaves_[i].hits.size() * anchors_.size()) throw std::runtime_error("explain: weight table of the wrong size");
        leaves_[i].weight = std::move(weight), leaves_[i].est = postings;
    }

    // {"<anchor>": [Explain, ...]} of the returned hits, serde's externally tagged form of the enum
    std::string to_json() {
        std::string out = "{";
        for (size_t a = 0; a < anchors_.size(); ++a) {
            const Items items = explain_anchor(a);
            out += (a ? ",\"" : "\"") + std::to_string(anchors_[a]) + "\":";
            write_items(out, items);
        }
        return out + "}";
    }

    // The anchor's explanations; empty when nothing on its way through the tree recorded any.
    Items explain_anchor(size_t a) {
        Node root = eval(*request_.search_req, a);
        // (the filter drops hits and leaves the explanations alone: intersect_score_hits_with_ids, set_op.rs:311-326)
        if (request_.boost && root.present)
            for (const vhost::BoostPart& bp : *request_.boost) add_boost(bp, root, anchors_[a]);
        last_score_ = root.score;
        return root.items;
    }
    float last_score() const { return last_score_; }  // the score the walk arrived at for the last anchor explained (equals the hit's unless boost_term / text locality follow)
    size_t n_anchors() const { return anchors_.size(); }
    uint32_t anchor(size_t a) const { return anchors_[a]; }
    const std::vector<uint32_t>& anchors() const { return anchors_; }

    static void write_items(std::string& out, const Items& items) {
        out += '[';
        char buf[192];
        for (size_t i = 0; i < items.size(); ++i) {
            const Item& e = items[i];
            if (i) out += ',';
            switch (e.kind) {
                case Item::Boost: snprintf(buf, sizeof buf, "{\"Boost\":%.9g}", (double)e.a), out += buf; break;
                case Item::OrSumOverDistinctTerms: snprintf(buf, sizeof buf, "{\"OrSumOverDistinctTerms\":%.9g}", (double)e.a), out += buf; break;
                case Item::TermToAnchor:
                    snprintf(buf, sizeof buf, "{\"TermToAnchor\":{\"term_score\":%.9g,\"anchor_score\":%.9g,\"final_score\":%.9g,\"term_id\":%u}}", (double)e.a, (double)e.b, (double)e.c, e.term_id);
                    out += buf;
                    break;
                case Item::LevenshteinScore:
                    snprintf(buf, sizeof buf, "{\"LevenshteinScore\":{\"score\":%.9g,\"text_or_token_id\":", (double)e.a);
                    out += buf;
                    vjson::write_string(out, e.text);
                    out += ",\"term_id\":" + std::to_string(e.term_id) + "}}";
                    break;
            }
        }
        out += ']';
    }

  private:
    const vhost::Persistence& host_;
    vhost::Request request_;
    std::vector<uint32_t> anchors_;
    std::vector<const vhost::SearchPart*> parts_;
    std::vector<Leaf> leaves_;
    float last_score_ = 0.f;

    static void check_no_tree_boosts(const vhost::SearchRequest& r) {
        // boosts in a sub-query's options only ever act as 1:n boosts (execution_plan.rs:263-270,422-436)
        if (r.get_boost() && !r.get_boost()->empty()) throw vplan::Unsupported("explain of 1:n boosts is outside the accelerated path");
        for (auto& q : r.queries) check_no_tree_boosts(q);
    }

    void collect(const vhost::SearchRequest& r) {
        if (r.kind == vhost::SearchRequest::Search) parts_.push_back(&r.part);
        for (auto& q : r.queries) collect(q);
    }

    Node eval(const vhost::SearchRequest& r, size_t a) {
        if (r.kind == vhost::SearchRequest::Search) return eval_leaf(r.part, a);
        std::vector<Node> in;
        for (auto& q : r.queries) in.push_back(eval(q, a));
        if (in.empty()) return Node();
        if (in.size() == 1) return std::move(in[0]);  // set_op.rs:90-96, :372-375
        return r.kind == vhost::SearchRequest::Or ? eval_or(in) : eval_and(in);
    }

    // resolve_token_to_anchor (search_field.rs:418-464): every matched term with a posting on the anchor, in hit order; the
    // anchor keeps the largest of their scores
    Node eval_leaf(const vhost::SearchPart& part, size_t a) {
        const Leaf& leaf = leaves_[(size_t)(std::find(parts_.begin(), parts_.end(), &part) - parts_.begin())];
        Node n;
        n.request = &part, n.is_leaf = true, n.est = leaf.est;
        const size_t na = anchors_.size();
        for (size_t t = 0; t < leaf.hits.size(); ++t) {
            const float w = leaf.weight[t * na + a];
            if (w < 0.0f) continue;
            const vdev::TermHit& h = leaf.hits[t];
            const float final_score = h.score * w;
            if (wants(part)) {
                Item e(Item::TermToAnchor);
                e.term_id = h.id, e.a = h.score, e.b = w, e.c = final_score;
                n.items.push_back(e);
                auto it = leaf.term_items.find(h.id);
                if (it != leaf.term_items.end()) n.items.insert(n.items.end(), it->second.begin(), it->second.end());
            }
            if (!n.present || final_score > n.score) n.score = final_score;
            n.present = true;
        }
        return n;
    }

    // union_hits_score (set_op.rs:87-220)
    Node eval_or(std::vector<Node>& in) {
        Node out;
        out.request = in[0].request;
        std::vector<std::string> terms;
        for (const Node& c : in) terms.push_back(c.request && !c.request->terms.empty() ? c.request->terms[0] : std::string());
        std::vector<std::string> slots = terms;
        std::sort(slots.begin(), slots.end());
        slots.erase(std::unique(slots.begin(), slots.end()), slots.end());
        std::vector<float> max_per_term(slots.size(), 0.0f);
        for (size_t i = 0; i < in.size(); ++i)
            if (in[i].present) {
                float& m = max_per_term[(size_t)(std::lower_bound(slots.begin(), slots.end(), terms[i]) - slots.begin())];
                m = fmaxf(m, in[i].score);
                out.present = true;
            }
        if (!out.present) return out;
        float n = 0.f, sum = 0.0f;
        for (float m : max_per_term) n += m >= 0.00001f ? 1.f : 0.f;
        for (float m : max_per_term) sum += m;
        out.score = sum * n * n;
        if (in[0].request && wants(*in[0].request)) {  // :120: the first input decides
            for (const Node& c : in)  // :133-137: a later input's explanations replace an earlier one's
                if (c.present && !c.items.empty()) out.items = c.items;
            out.items.push_back(Item(Item::OrSumOverDistinctTerms, sum));
            for (const Node& c : in)
                if (c.present) out.items.insert(out.items.end(), c.items.begin(), c.items.end());
        }
        return out;
    }

    // intersect_hits_score (set_op.rs:368-446): the first shortest input is taken out (the last one takes its place), its score
    // is added last and its explanations are not carried on.  Lengths as the engine orders them: the parts' posting counts,
    // request order when an input is a sub-tree.
    Node eval_and(std::vector<Node>& in) {
        Node out;
        const bool should_explain = in[0].request && wants(*in[0].request);  // :384, before the removal
        size_t shortest = 0;
        bool known = true;
        for (const Node& c : in) known = known && c.is_leaf;
        if (!known) {
            shortest = in.size() - 1;
        } else {
            for (size_t i = 1; i < in.size(); ++i)
                if (in[i].est < in[shortest].est) shortest = i;
        }
        for (const Node& c : in)
            if (!c.present) return out;
        Node last = std::move(in[shortest]);
        if (shortest != in.size() - 1) in[shortest] = std::move(in.back());
        in.pop_back();
        out.present = true;
        float score = 0.0f;
        for (const Node& c : in) score += c.score;
        score += last.score;
        out.score = score;
        out.request = in[0].request;
        if (should_explain)
            for (const Node& c : in) out.items.insert(out.items.end(), c.items.begin(), c.items.end());
        return out;
    }

    // add_boost on the final hits (boost.rs:470-504, apply_boost :283-377)
    void add_boost(const vhost::BoostPart& bp, Node& hit, uint32_t anchor) {
        const vhost::KeyValueStore& store = host_.get_boost(bp.path + ".boost_valid_to_value");
        if (bp.skip_when_score)
            for (float x : *bp.skip_when_score)
                if (fabsf(x - hit.score) < 0.00001f) return;
        uint32_t bits = 0;
        if (!store.get_value(anchor, bits)) return;
        float v;
        memcpy(&v, &bits, 4);
        std::vector<float> log;
        vdev::apply_boost_value(bp, v, hit.score, &log);
        if (hit.request && wants(*hit.request))
            for (float f : log) hit.items.push_back(Item(Item::Boost, f));
    }
};

}  // namespace vexplain
