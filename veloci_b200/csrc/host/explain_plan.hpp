// search::explain_plan (src/search.rs:132-141): the execution plan of a request as a Graphviz dot graph, as
// plan_creator (src/plan_creator/execution_plan.rs:132-200) lays it out and render_plan_to (plan.rs:81-125) prints it --
// one node per plan step labelled with the step's Display text (plan_steps.rs:76-135), one edge per dependency.  The
// reference's plan is what a host that drives the step symbols itself would run; vgpu_batch_execute fuses the same steps
// into one pass per tile, so this graph documents the request, not a schedule.  Host code, no index needed.
// Pinned by the reference only through tests.rs:1210-1230 (the text contains the term, the field and "boost"); node
// numbering and edge order follow the order of add_step / add_dependency calls in plan_creator.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "part_hits.hpp"
#include "request.hpp"

namespace vhost {

class PlanGraph {
  public:
    explicit PlanGraph(Request request) {
        if (!request.search_req) throw vplan::InvalidRequest("search_req is None, but is required in search");
        if (request.explain) {  // settings are propagated before the parts are collected: they are part of a part's identity
            propagate_explain(*request.search_req);
            if (request.filter) propagate_explain(*request.filter);
            if (request.phrase_boosts)
                for (PhraseBoost& pb : *request.phrase_boosts) set_explain(pb.search1), set_explain(pb.search2);
        }
        // collect_all_field_request_into_cache (:91-130): phrase boost parts, the tree's parts, then the filter's
        if (request.phrase_boosts)
            for (const PhraseBoost& pb : *request.phrase_boosts) field_search(pb.search1), field_search(pb.search2);
        collect(*request.search_req);
        if (request.filter) collect(*request.filter);
        long filter_final = -1;
        if (request.filter) filter_final = create(true, *request.filter, {}, -1, -1);
        const std::vector<BoostPart> boosts = request.boost.value_or(std::vector<BoostPart>());
        long final_step = create(false, *request.search_req, boosts, -1, filter_final);
        if (filter_final >= 0) {
            const long id = add_step("IntersectScoresWithIds");
            add_dependency(id, filter_final), add_dependency(id, final_step);
            final_step = id;
        }
        for (const BoostPart& b : boosts)  // :175-189: boosts on the anchor level follow the tree
            if (b.path.find("[]") == std::string::npos) {
                const long id = add_step("BoostPlanStepFromBoostRequest");
                add_dependency(id, final_step);
                final_step = id;
            }
        if (request.phrase_boosts) {  // add_phrase_boost_plan_steps :202-262
            for (const PhraseBoost& pb : *request.phrase_boosts) {
                const long id = add_step("PlanStepPhrasePairToAnchorId");
                add_dependency(id, field_search(pb.search1)), add_dependency(id, field_search(pb.search2));
            }
            const long id = add_step("BoostAnchorFromPhraseResults");
            add_dependency(id, final_step);
        }
    }

    // dot::render: ids N<i>, labels escaped like Rust's escape_default
    std::string to_dot() const {
        std::string out = "digraph example2 {\n";
        for (size_t i = 0; i < nodes_.size(); ++i) out += "    N" + std::to_string(i) + "[label=\"" + escape(nodes_[i] + "\n") + "\"];\n";
        for (auto& e : deps_) out += "    N" + std::to_string(e.second) + " -> N" + std::to_string(e.first) + "[label=\"\"];\n";
        return out + "}\n";
    }
    size_t n_steps() const { return nodes_.size(); }

  private:
    std::vector<std::string> nodes_;
    std::vector<std::pair<long, long>> deps_;  // (step, depends on)
    std::map<std::string, long> cache_;         // FieldRequestCache: part -> its PlanStepFieldSearchToTokenIds

    static void set_explain(SearchPart& p) { p.options.present = true, p.options.explain = true; }
    static void propagate_explain(SearchRequest& r) {
        if (r.kind == SearchRequest::Search) set_explain(r.part);
        for (SearchRequest& q : r.queries) propagate_explain(q);
    }
    long add_step(const std::string& label) {
        nodes_.push_back(label);
        return (long)nodes_.size() - 1;
    }
    void add_dependency(long step, long depends_on) { deps_.emplace_back(step, depends_on); }
    long field_search(const SearchPart& part) {
        auto it = cache_.find(part.key());
        if (it != cache_.end()) return it->second;
        const long id = add_step("search " + part.path + " " + (part.terms.empty() ? std::string() : part.terms[0]));
        cache_.emplace(part.key(), id);
        return id;
    }
    void collect(const SearchRequest& r) {
        if (r.kind == SearchRequest::Search) field_search(r.part);
        for (const SearchRequest& q : r.queries) collect(q);
    }

    // plan_creator_2 (:272-387) and plan_creator_search_part (:389-534)
    long create(bool is_filter, const SearchRequest& r, std::vector<BoostPart> boosts, long parent, long depends_on) {
        if (r.kind != SearchRequest::Search) {
            const long step = add_step(r.kind == SearchRequest::Or ? "Union" : "Intersect");
            for (const SearchRequest& q : r.queries) {
                std::vector<BoostPart> b = boosts;  // merge_vec :263-270
                if (q.get_boost()) b.insert(b.end(), q.get_boost()->begin(), q.get_boost()->end());
                create(is_filter, q, b, step, depends_on);
            }
            if (parent >= 0) add_dependency(parent, step);
            if (depends_on >= 0) add_dependency(step, depends_on);
            return step;
        }
        const SearchPart& part = r.part;
        const long fs = field_search(part);
        const size_t pos = part.path.rfind("[]");
        if (pos != std::string::npos) {
            const std::string level = part.path.substr(0, pos);
            const BoostPart* on_level = nullptr;
            for (const BoostPart& b : boosts) {
                const size_t bp = b.path.rfind("[]");
                if (bp != std::string::npos && b.path.substr(0, bp) == level && !on_level) on_level = &b;
            }
            if (on_level) {  // a boost on the part's 1:n level: BoostToAnchor + ApplyAnchorBoost (:438-509)
                const long resolve = add_step("token to anchor");
                add_dependency(resolve, fs);
                if (depends_on >= 0) add_dependency(resolve, depends_on);
                const long to_anchor = add_step("BoostToAnchor " + on_level->path);
                add_dependency(to_anchor, fs);
                const long apply = add_step("ApplyAnchorBoost");
                add_dependency(apply, to_anchor), add_dependency(apply, resolve);
                if (parent >= 0) add_dependency(parent, apply);
                if (depends_on >= 0) add_dependency(apply, depends_on);
                return apply;
            }
        }
        const long resolve = add_step("token to anchor");
        add_dependency(resolve, fs);
        if (parent >= 0) add_dependency(parent, resolve);
        if (depends_on >= 0) add_dependency(resolve, depends_on);
        return resolve;
    }

    static std::string escape(const std::string& s) {  // char::escape_default over the string's scalars
        std::vector<uint32_t> cps;
        vfmt::utf8_decode(s, cps);
        std::string out;
        char buf[16];
        for (uint32_t c : cps) {
            if (c == '\t') out += "\\t";
            else if (c == '\r') out += "\\r";
            else if (c == '\n') out += "\\n";
            else if (c == '\'' || c == '"' || c == '\\') out += '\\', out += (char)c;
            else if (c >= 0x20 && c <= 0x7E) out += (char)c;
            else snprintf(buf, sizeof buf, "\\u{%x}", c), out += buf;
        }
        return out;
    }
};

inline std::string explain_plan(const Request& request) { return PlanGraph(request).to_dot(); }

}  // namespace vhost
