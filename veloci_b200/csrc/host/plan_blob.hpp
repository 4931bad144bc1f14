// Position-independent form of a BatchPlan.
//
// The reference plans a request where it executes it (plan_creator, src/plan_creator/execution_plan.rs:132-200).  With
// one process per GPU on a box, every process would parse and plan the same batch: N ranks x 16 planner threads fight
// for the host cores and the end-to-end rate falls with N.  Instead one process plans, exports the plan as a byte blob,
// and the other processes import it against their own handle of the same index directory.
//
// The plan tables carry device pointers of index structures (boost columns and their level headers, id -> ids stores,
// phrase stores).  They are all base pointers of allocations that every handle of the same directory owns in the same
// order (DeviceIndex::reloc_ptrs, filled in name order at open), so the blob stores the allocation's ordinal + 1 in the
// pointer's place and the importer puts its own pointer back.  The blob starts with a fingerprint of the directory's
// structure names and sizes: importing against another index fails instead of reading foreign memory.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "planner.hpp"

namespace vplan {

struct BlobError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class BlobWriter {
   public:
    std::vector<uint8_t> out;
    void raw(const void* p, size_t n) {
        const uint8_t* b = static_cast<const uint8_t*>(p);
        out.insert(out.end(), b, b + n);
    }
    void u64(uint64_t v) { raw(&v, 8); }
    void u32(uint32_t v) { raw(&v, 4); }
    void str(const std::string& s) {
        u32((uint32_t)s.size());
        raw(s.data(), s.size());
    }
    template <class T>
    void pod_vec(const std::vector<T>& v) {
        u64(v.size());
        if (!v.empty()) raw(v.data(), v.size() * sizeof(T));
        while (out.size() & 7) out.push_back(0);
    }
    void str_vec(const std::vector<std::string>& v) {
        u64(v.size());
        for (auto& s : v) str(s);
    }
    template <class T>
    void opt(const std::optional<T>& o) {
        u32(o ? 1u : 0u);
        T v = o ? *o : T();
        raw(&v, sizeof(T));
    }
};

class BlobReader {
   public:
    BlobReader(const void* p, size_t n) : begin_(static_cast<const uint8_t*>(p)), p_(begin_), end_(begin_ + n) {}
    void raw(void* dst, size_t n) {
        if ((size_t)(end_ - p_) < n) throw BlobError("plan blob is truncated");
        memcpy(dst, p_, n);
        p_ += n;
    }
    uint64_t u64() {
        uint64_t v;
        raw(&v, 8);
        return v;
    }
    uint32_t u32() {
        uint32_t v;
        raw(&v, 4);
        return v;
    }
    std::string str() {
        const uint32_t n = u32();
        if ((size_t)(end_ - p_) < n) throw BlobError("plan blob is truncated");
        std::string s(reinterpret_cast<const char*>(p_), n);
        p_ += n;
        return s;
    }
    template <class T>
    void pod_vec(std::vector<T>& v) {
        const uint64_t n = u64();
        if (n > (uint64_t)(end_ - p_) / sizeof(T)) throw BlobError("plan blob is truncated");
        v.resize((size_t)n);
        if (n) raw(v.data(), (size_t)n * sizeof(T));
        while (((size_t)(p_ - begin_) & 7) && p_ < end_) ++p_;  // the writer pads every table to 8 bytes
    }
    void str_vec(std::vector<std::string>& v) {
        const uint64_t n = u64();
        v.clear();
        for (uint64_t i = 0; i < n; ++i) v.push_back(str());
    }
    template <class T>
    void opt(std::optional<T>& o) {
        const uint32_t has = u32();
        T v;
        raw(&v, sizeof(T));
        if (has) o = v;
        else o.reset();
    }

   private:
    const uint8_t* begin_;
    const uint8_t* p_;
    const uint8_t* end_;
};

// Every device pointer of a plan table, visited in place.
template <class F>
inline void visit_ptrs(vdev::CsrView& v, F&& f) {
    f(reinterpret_cast<const void*&>(v.off)), f(reinterpret_cast<const void*&>(v.val));
}
template <class F>
inline void visit_ptrs(vdev::PhraseView& v, F&& f) {
    f(reinterpret_cast<const void*&>(v.keys)), f(reinterpret_cast<const void*&>(v.off)), f(reinterpret_cast<const void*&>(v.anchors));
}
template <class F>
inline void visit_ptrs(BoostStep& b, F&& f) {
    f(reinterpret_cast<const void*&>(b.column)), f(reinterpret_cast<const void*&>(b.levels));
}
template <class F>
inline void visit_ptrs(QueryProgram& q, F&& f) {
    f(reinterpret_cast<const void*&>(q.fb_col)), f(reinterpret_cast<const void*&>(q.fb_lev));
}
template <class F>
inline void visit_ptrs(vdev::PhraseMember& m, F&& f) {
    visit_ptrs(m.store, f);
}
template <class F>
inline void visit_ptrs(vdev::IdsMember& m, F&& f) {
    visit_ptrs(m.text_id_to_anchor, f);
}
template <class F>
inline void visit_ptrs(vdev::BoostListMember& m, F&& f) {
    visit_ptrs(m.tokens_to_text_id, f), visit_ptrs(m.value_id_to_parent, f), visit_ptrs(m.value_id_to_anchor, f);
    f(reinterpret_cast<const void*&>(m.column));
}
template <class F>
inline void visit_ptrs(vdev::TlInstance& t, F&& f) {
    visit_ptrs(t.tokens_to_text_id, f), visit_ptrs(t.text_id_to_anchor, f);
}
template <class F>
inline void visit_ptrs(vdev::FacetStep& s, F&& f) {
    for (uint32_t i = 0; i < vdev::kMaxFacetSteps; ++i) visit_ptrs(s.step[i], f);
}

template <class T, class F>
inline void visit_all(std::vector<T>& v, F&& f) {
    for (auto& x : v) visit_ptrs(x, f);
}

static const uint64_t kBlobMagic = 0x33424c504f4c4556ull;  // "VELOPLB3"

inline void write_search_part(BlobWriter& w, const vhost::SearchPart& p) {
    if (p.options.present) throw BlobError("search part with options cannot be exported");
    w.str(p.path);
    w.u32(p.is_regex ? 1u : 0u);
    w.str_vec(p.terms);
    w.opt(p.levenshtein_distance);
    w.u32(p.starts_with ? 1u : 0u);
    w.opt(p.boost);
    w.opt(p.ignore_case);
    w.opt(p.top);
    w.opt(p.skip);
    w.u32(p.token_value ? 1u : 0u);
    if (p.token_value) {
        const vhost::BoostPart& b = *p.token_value;
        w.str(b.path);
        w.u32((uint32_t)b.boost_fun);
        w.opt(b.param);
        w.u32(b.skip_when_score ? (uint32_t)b.skip_when_score->size() + 1u : 0u);
        if (b.skip_when_score)
            for (float f : *b.skip_when_score) w.raw(&f, 4);
        w.u32(b.expression ? 1u : 0u);
        if (b.expression) w.str(*b.expression);
    }
}
inline void read_search_part(BlobReader& r, vhost::SearchPart& p) {
    p.path = r.str();
    p.is_regex = r.u32() != 0;
    r.str_vec(p.terms);
    r.opt(p.levenshtein_distance);
    p.starts_with = r.u32() != 0;
    r.opt(p.boost);
    r.opt(p.ignore_case);
    r.opt(p.top);
    r.opt(p.skip);
    p.token_value.reset();
    if (r.u32() != 0) {
        vhost::BoostPart b;
        b.path = r.str();
        const uint32_t fun = r.u32();
        if (fun > (uint32_t)vhost::BoostFun::Replace) throw BlobError("plan blob is inconsistent (boost function)");
        b.boost_fun = (vhost::BoostFun)fun;
        r.opt(b.param);
        const uint32_t n_skip = r.u32();
        if (n_skip) {
            b.skip_when_score.emplace();
            for (uint32_t i = 1; i < n_skip; ++i) {
                float f;
                r.raw(&f, 4);
                b.skip_when_score->push_back(f);
            }
        }
        if (r.u32() != 0) b.expression = r.str();
        p.token_value = std::move(b);
    }
}

// The plan as bytes without process-local addresses.  `plan` is not modified.
inline std::vector<uint8_t> export_plan(const BatchPlan& plan_in) {
    const vdev::DeviceIndex& ix = *plan_in.ix;
    BlobWriter w;
    w.u64(kBlobMagic);
    w.u64(ix.reloc_fingerprint);
    auto to_ordinal = [&](const void*& p) {
        if (!p) return;
        auto it = ix.reloc_index.find(p);
        if (it == ix.reloc_index.end()) throw BlobError("plan table points outside the index structures");
        p = reinterpret_cast<const void*>((uintptr_t)it->second + 1);
    };
    auto pods = [&](auto v) {  // by value: relocated copy
        visit_all(v, to_ordinal);
        w.pod_vec(v);
    };
    w.u64(plan_in.requests.size());
    for (auto& rq : plan_in.requests) {
        w.u32((uint32_t)rq.status);
        w.str(rq.message);
        w.u64(rq.top), w.u64(rq.skip);
        w.u32((rq.has_facets ? 1u : 0u) | (rq.why_found ? 2u : 0u) | (rq.explain_asked ? 4u : 0u)), w.u32(rq.facet_begin);
        w.u64(rq.facets.size());
        for (auto& f : rq.facets) w.str(f.field), w.opt(f.top);
        w.u32(rq.select ? 1u : 0u);
        if (rq.select) w.str_vec(*rq.select);
    }
    w.pod_vec(plan_in.parts);
    w.pod_vec(plan_in.part_dict);
    w.str_vec(plan_in.dict_names);
    w.str_vec(plan_in.postings_names);
    pods(plan_in.programs);
    w.pod_vec(plan_in.leaf_part);
    w.pod_vec(plan_in.prog);
    pods(plan_in.boosts);
    pods(plan_in.phrase_members);
    pods(plan_in.ids_members);
    pods(plan_in.boost_members);
    pods(plan_in.tl_instances);
    w.pod_vec(plan_in.tl_term_parts);
    {
        std::vector<vdev::FacetStep> f = plan_in.facets;
        for (auto& s : f) s.hist = nullptr;  // set by the importing engine
        visit_all(f, to_ordinal);
        w.pod_vec(f);
    }
    w.pod_vec(plan_in.facet_top);
    w.str_vec(plan_in.facet_text_path);
    w.u32(plan_in.max_leaves), w.u32(plan_in.max_k);
    w.u64(plan_in.bounded.size());
    for (auto& b : plan_in.bounded) w.u32(b.part), write_search_part(w, b.request);
    w.u64(plan_in.regex_parts.size());
    for (auto& g : plan_in.regex_parts) w.u32(g.part), w.str(g.pattern), w.u32((g.case_insensitive ? 1u : 0u) | (g.starts_with ? 2u : 0u));
    return std::move(w.out);
}

inline void import_plan(const vdev::DeviceIndex* ix, const void* blob, size_t len, BatchPlan& plan) {
    BlobReader r(blob, len);
    if (r.u64() != kBlobMagic) throw BlobError("not a veloci_b200 plan blob");
    if (r.u64() != ix->reloc_fingerprint) throw BlobError("the plan was made for another index (structure fingerprint differs)");
    auto to_pointer = [&](const void*& p) {
        if (!p) return;
        const uintptr_t ord = reinterpret_cast<uintptr_t>(p) - 1;
        if (ord >= ix->reloc_ptrs.size()) throw BlobError("plan blob references an unknown index structure");
        p = ix->reloc_ptrs[ord];
    };
    auto pods = [&](auto& v) {
        r.pod_vec(v);
        visit_all(v, to_pointer);
    };
    plan.ix = ix;
    const uint64_t n_req = r.u64();
    plan.requests.clear();
    plan.requests.reserve((size_t)n_req);
    for (uint64_t i = 0; i < n_req; ++i) {
        RequestPlan rq;
        rq.status = (int32_t)r.u32();
        rq.message = r.str();
        rq.top = r.u64(), rq.skip = r.u64();
        const uint32_t request_flags = r.u32();
        rq.has_facets = request_flags & 1u, rq.why_found = request_flags & 2u, rq.explain_asked = request_flags & 4u, rq.facet_begin = r.u32();
        const uint64_t nf = r.u64();
        for (uint64_t j = 0; j < nf; ++j) {
            vhost::FacetRequest f;
            f.field = r.str();
            r.opt(f.top);
            rq.facets.push_back(std::move(f));
        }
        if (r.u32()) {
            std::vector<std::string> fields;
            r.str_vec(fields);
            rq.select = std::move(fields);
        }
        plan.requests.push_back(std::move(rq));
    }
    r.pod_vec(plan.parts);
    r.pod_vec(plan.part_dict);
    r.str_vec(plan.dict_names);
    r.str_vec(plan.postings_names);
    for (auto& name : plan.dict_names)
        if (!ix->dicts.count(name)) throw BlobError("plan blob names a dictionary this index does not have: " + name);
    for (auto& name : plan.postings_names)
        if (!ix->postings.count(name)) throw BlobError("plan blob names a postings store this index does not have: " + name);
    pods(plan.programs);
    r.pod_vec(plan.leaf_part);
    r.pod_vec(plan.prog);
    pods(plan.boosts);
    pods(plan.phrase_members);
    pods(plan.ids_members);
    pods(plan.boost_members);
    pods(plan.tl_instances);
    r.pod_vec(plan.tl_term_parts);
    pods(plan.facets);
    r.pod_vec(plan.facet_top);
    r.str_vec(plan.facet_text_path);
    plan.max_leaves = r.u32(), plan.max_k = r.u32();
    const uint64_t nb = r.u64();
    plan.bounded.clear();
    for (uint64_t i = 0; i < nb; ++i) {
        BatchPlan::BoundedPart b;
        b.part = r.u32();
        read_search_part(r, b.request);
        plan.bounded.push_back(std::move(b));
    }
    const uint64_t ng = r.u64();
    plan.regex_parts.clear();
    for (uint64_t i = 0; i < ng; ++i) {
        BatchPlan::RegexPart g;
        g.part = r.u32();
        g.pattern = r.str();
        const uint32_t f = r.u32();
        g.case_insensitive = f & 1u, g.starts_with = f & 2u;
        if (g.part >= plan.parts.size()) throw BlobError("plan blob is inconsistent (regex part)");
        plan.regex_parts.push_back(std::move(g));
    }
    if (plan.programs.size() != plan.requests.size() || plan.part_dict.size() != plan.parts.size()) throw BlobError("plan blob is inconsistent");
}

}  // namespace vplan
