// Request -> batch programs.  The host half of plan_creator
// (src/plan_creator/execution_plan.rs:132-534): the request tree becomes a postfix
// program over "leaves" (one per `search` node), identical RequestSearchParts are
// searched once per *batch* (FieldRequestCache, :13,91-130, widened from one request
// to the whole batch), anchor-level boosts are chained after the root (:175-189).
#pragma once
#include <cmath>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../cuda/device_types.cuh"
#include "device_index.hpp"
#include "part_hits.hpp"
#include "regex_dfa.hpp"
#include "request.hpp"

namespace vplan {

using vdev::BoostStep;
using vdev::PartQuery;
using vdev::QueryProgram;

struct RequestPlan {
    int32_t status = 0;
    std::string message;
    uint64_t top = 10, skip = 0;
    std::vector<vhost::FacetRequest> facets;
    bool has_facets = false;
    uint32_t facet_begin = 0;  // first of the request's entries in BatchPlan::facets
    std::shared_ptr<const vhost::Request> explain;  // the request itself when it asks for explain (host/explain.hpp rebuilds the explanations of its hits)
    bool explain_asked = false;  // ... it did (an imported plan carries only this)
    bool why_found = false;    // Request::why_found: the matched term texts are kept for highlighting (execution_plan.rs:416)
    std::optional<std::vector<std::string>> select;  // Request::select: the hits' documents are rebuilt from the indices (search.rs:82-88)
};

struct BatchPlan {
    const vdev::DeviceIndex* ix = nullptr;
    std::vector<RequestPlan> requests;
    // distinct search parts of the batch
    std::vector<PartQuery> parts;
    std::vector<uint32_t> part_dict;  // part -> dictionary table index
    // Parts are unified by what the device sees of them: two RequestSearchParts with the same PartQuery on the same
    // dictionary produce the same hits (the reference's FieldRequestCache keys on the request text instead,
    // execution_plan.rs:13,108-130; sharing more is not observable).  Open addressing over the part's hash.
    std::vector<uint64_t> part_hash;   // part -> hash of its PartQuery (without the table indexes), 0 for list parts
    std::vector<uint32_t> part_slot;   // hash table: part id + 1, 0 = empty
    size_t part_slots_used = 0;

    static uint64_t hash_part(const PartQuery& q) {
        uint64_t h = 0x9E3779B97F4A7C15ull;
        auto mix = [&](uint64_t v) {
            h ^= v;
            h *= 0xFF51AFD7ED558CCDull;
            h ^= h >> 32;
        };
        const uint32_t m = std::min<uint32_t>(std::max(q.m, q.m_score), 64);
        for (uint32_t j = 0; j < m; ++j) mix(((uint64_t)q.match_sym[j] << 16) | q.score_sym[j]);
        mix(q.m_score);
        uint32_t boost_bits;
        memcpy(&boost_bits, &q.boost, 4);
        mix(((uint64_t)q.m << 32) | q.d_score);
        mix(((uint64_t)q.flags << 32) | boost_bits);
        mix(q.lower_bytes);
        return h | 1;  // never 0
    }
    void grow_part_slots(size_t want_parts) {
        size_t cap = 64;
        while (cap < want_parts * 2 + 2) cap <<= 1;
        if (cap <= part_slot.size()) return;
        part_slot.assign(cap, 0);
        part_slots_used = 0;
        for (size_t i = 0; i < parts.size(); ++i)
            if (part_hash[i]) {
                size_t at = part_hash[i] & (cap - 1);
                while (part_slot[at]) at = (at + 1) & (cap - 1);
                part_slot[at] = (uint32_t)i + 1, ++part_slots_used;
            }
    }
    // The part equal to (q, dictionary), added if new.  `q` must have been zero-filled before its fields were set.
    uint32_t find_or_add_part(const PartQuery& q, uint32_t dict, uint64_t h) {
        if ((part_slots_used + 1) * 2 > part_slot.size()) grow_part_slots(std::max<size_t>(parts.size() + 1, part_slots_used * 2));
        const size_t mask = part_slot.size() - 1;
        size_t at = h & mask;
        for (; part_slot[at]; at = (at + 1) & mask) {
            const uint32_t id = part_slot[at] - 1;
            if (part_hash[id] == h && part_dict[id] == dict && memcmp(&parts[id], &q, sizeof q) == 0) return id;
        }
        const uint32_t id = (uint32_t)parts.size();
        parts.push_back(q);
        part_dict.push_back(dict);
        part_hash.push_back(h);
        part_slot[at] = id + 1, ++part_slots_used;
        return id;
    }
    // tables of device views referenced by index
    std::vector<std::string> dict_names, postings_names;
    std::unordered_map<std::string, uint32_t> dict_index, postings_index;
    // programs
    std::vector<QueryProgram> programs;
    std::vector<uint32_t> leaf_part;
    std::vector<uint32_t> prog;
    std::vector<BoostStep> boosts;
    std::vector<vdev::PhraseMember> phrase_members;
    std::vector<vdev::IdsMember> ids_members;
    std::vector<vdev::BoostListMember> boost_members;
    std::vector<vdev::TlInstance> tl_instances;
    std::vector<uint32_t> tl_term_parts;
    struct TlTerm {  // a search part of the request tree, for term_id_hits_in_field (search_field.rs:379-383)
        std::string path, term;
        uint32_t part;
    };
    std::vector<TlTerm> tl_seen;  // of the request being planned
    std::vector<vdev::FacetStep> facets;   // hist pointers are set by the engine
    std::vector<uint32_t> facet_top;       // groups wanted per facet
    std::vector<std::string> facet_text_path;  // dictionary the value ids of the facet belong to
    uint32_t max_leaves = 1, max_k = 1;

    uint32_t dict_id(const std::string& path) {
        auto it = dict_index.find(path);
        if (it != dict_index.end()) return it->second;
        if (!ix->dicts.count(path)) throw vhost::FstNotFound(path);
        uint32_t id = (uint32_t)dict_names.size();
        dict_names.push_back(path);
        dict_index.emplace(path, id);
        return id;
    }
    uint32_t postings_id(const std::string& path) {
        auto it = postings_index.find(path);
        if (it != postings_index.end()) return it->second;
        if (!ix->postings.count(path)) ix->host->path_not_found(path);
        uint32_t id = (uint32_t)postings_names.size();
        postings_names.push_back(path);
        postings_index.emplace(path, id);
        return id;
    }

    // What a field's search parts look up, resolved once per plan and field.
    struct FieldTables {
        std::string field;  // as written in the request
        uint32_t dict_id, postings_id;
        const vdev::DictDev* dict;
        uint16_t ascii_code[128];  // alphabet code of the ASCII scalars
    };
    std::vector<FieldTables> fields_seen;
    const FieldTables& field_tables(const std::string& field) {
        for (auto& f : fields_seen)
            if (f.field == field) return f;
        std::string path = field;
        if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
        FieldTables f;
        f.field = field;
        f.dict_id = dict_id(path);  // FstNotFound
        f.dict = &ix->dicts.at(path);
        f.postings_id = postings_id(path + ".to_anchor_id_score");
        for (uint32_t c = 0; c < 128; ++c) f.ascii_code[c] = f.dict->code_of(c);
        fields_seen.push_back(std::move(f));
        return fields_seen.back();
    }

    // get_term_ids_in_field's request normalisation (search_field.rs:277-300)
    uint32_t add_part(const vhost::SearchPart& part_in) {
        const vhost::SearchPart& req = part_in;
        if (req.terms.empty()) throw InvalidRequest("search part without terms");
        if (req.token_value) {  // search_field.rs:391-395: the store must exist (persistence.rs:454-458), the expression parse
            ix->host->get_boost(req.token_value->path + ".textindex.token_values.boost_valid_to_value");
            BoostStep probe;
            memset(&probe, 0, sizeof probe);
            if (req.token_value->expression) parse_expression(*req.token_value->expression, probe);
        }
        const FieldTables& ft = field_tables(req.path);
        const uint32_t did = ft.dict_id, pid = ft.postings_id;
        const std::string& term = req.terms[0];
        bool ascii = true;
        for (unsigned char c : term) ascii &= c < 0x80;
        std::vector<uint32_t>&raw = scratch_raw, &low = scratch_low;
        raw.clear(), low.clear();
        size_t lower_bytes = term.size();
        if (ascii) {
            for (unsigned char c : term) raw.push_back(c), low.push_back((c >= 'A' && c <= 'Z') ? c + 32u : c);
        } else {
            const std::string lower_term = vfmt::to_lowercase(term);
            vfmt::utf8_decode(term, raw);
            vfmt::utf8_decode(lower_term, low);
            lower_bytes = lower_term.size();
        }
        if (raw.size() > 64 || low.size() > 64) throw Unsupported("search terms longer than 64 characters are outside the accelerated path");
        if (req.is_regex) {
            // search_field.rs:72-78: the pattern must compile (the reference unwraps the build); the tables for the
            // dictionary's alphabet are made when the plan is uploaded (Batch::prepare_tables), so the plan stays exportable
            try {
                vregex::compile(term, req.ignore_case.value_or(true));
            } catch (const vregex::RegexError& e) {
                throw InvalidRequest(std::string("regex \"") + term + "\": " + e.what());
            } catch (const vregex::RegexUnsupported& e) {
                throw Unsupported(std::string("regex \"") + term + "\": " + e.what() + " is outside the accelerated path");
            }
        }
        PartQuery q;
        memset(&q, 0, sizeof q);
        q.m = (uint32_t)raw.size();
        q.m_score = (uint32_t)low.size();  // (Rust's to_lowercase: U+0130 becomes two scalars)
        uint32_t d = 0;
        if (req.levenshtein_distance) d = std::min<uint32_t>(*req.levenshtein_distance, (uint32_t)low.size() - 1u);  // wraps for "" like release Rust
        q.d_score = d;
        q.d_match = std::min<uint32_t>(d, 4);
        const bool transposition = req.ignore_case.value_or(false);   // search_field.rs:87 (sic)
        const bool case_insensitive = req.ignore_case.value_or(true);  // :88
        q.flags = (req.starts_with ? vdev::kPartPrefix : 0u) | (transposition ? vdev::kPartTransposition : 0u) | (case_insensitive ? 0u : vdev::kPartRawCase) |
                  ((req.starts_with || d != 0) ? vdev::kPartCheckPrefix : 0u) | (req.boost ? vdev::kPartHasBoost : 0u);
        q.boost = req.boost.value_or(1.0f);
        q.lower_bytes = (uint32_t)lower_bytes;
        q.postings = pid;
        auto code_of = [&](uint32_t scalar) { return scalar < 128 ? ft.ascii_code[scalar] : ft.dict->code_of(scalar); };
        for (size_t j = 0; j < raw.size(); ++j) q.match_sym[j] = code_of(case_insensitive ? vfmt::lower_scalar(raw[j]) : raw[j]);
        for (size_t j = 0; j < low.size(); ++j) q.score_sym[j] = code_of(low[j]);
        if (req.is_regex && !req.top) {  // never shared: its matches come from its own DFA, which PartQuery does not describe
            q.flags |= vdev::kPartRegex;
            const uint32_t id = (uint32_t)parts.size();
            parts.push_back(q);
            part_dict.push_back(did);
            part_hash.push_back(0);
            regex_parts.push_back(RegexPart{id, term, req.ignore_case.value_or(true), req.starts_with});
            return id;
        }
        if (req.top || req.token_value) {
            // The per-part bound (search_field.rs:292-294,:322-331,:366-369) depends on the order the hits arrive in, and the
            // token_value boost (:391-395) comes after it: the engine matches such a part on its own first, bounds and boosts
            // its hits on the host and gives them to the batch (kPartInjected).
            q.flags |= vdev::kPartInjected | (req.token_value ? vdev::kPartAnySign : 0u);
            const uint32_t id = (uint32_t)parts.size();
            parts.push_back(q);
            part_dict.push_back(did);
            part_hash.push_back(0);  // never shared
            bounded.push_back(BoundedPart{id, req});
            return id;
        }
        return find_or_add_part(q, did, hash_part(q));
    }
    struct BoundedPart {
        uint32_t part;
        vhost::SearchPart request;
    };
    std::vector<BoundedPart> bounded;
    struct RegexPart {  // a kPartRegex part and what its DFA is built from
        uint32_t part;
        std::string pattern;
        bool case_insensitive, starts_with;
    };
    std::vector<RegexPart> regex_parts;

    // A part whose hits are produced by a list kernel (phrase pairs, text locality, 1:n boosts) instead of a field search.
    uint32_t add_list_part() {
        PartQuery q;
        memset(&q, 0, sizeof q);
        q.flags = vdev::kPartList;
        q.postings = vdev::kNoValue;
        const uint32_t id = (uint32_t)parts.size();
        parts.push_back(q);
        part_dict.push_back(0);
        part_hash.push_back(0);  // never shared
        return id;
    }

    // Room for the merged tables of `chunks`, so that merging does not reallocate.
    void reserve_for(const std::vector<BatchPlan>& chunks) {
        size_t n_parts = 0, n_leaf = 0, n_prog = 0, n_boost = 0, n_req = 0;
        for (auto& c : chunks) n_parts += c.parts.size(), n_leaf += c.leaf_part.size(), n_prog += c.prog.size(), n_boost += c.boosts.size(), n_req += c.requests.size();
        parts.reserve(n_parts), part_dict.reserve(n_parts), part_hash.reserve(n_parts);
        grow_part_slots(n_parts);
        leaf_part.reserve(n_leaf), prog.reserve(n_prog), boosts.reserve(n_boost), programs.reserve(n_req), requests.reserve(n_req);
    }

    // Appends the plan of another chunk of the same batch (built on another thread): its parts are unified with the
    // parts already known, its tables are appended with their indexes shifted.
    void merge(BatchPlan&& o) {
        std::vector<uint32_t> dict_map(o.dict_names.size()), post_map(o.postings_names.size()), part_map(o.parts.size());
        for (size_t i = 0; i < o.dict_names.size(); ++i) dict_map[i] = dict_id(o.dict_names[i]);
        for (size_t i = 0; i < o.postings_names.size(); ++i) post_map[i] = postings_id(o.postings_names[i]);
        for (size_t i = 0; i < o.parts.size(); ++i) {
            if (o.parts[i].flags & vdev::kPartList) {  // list parts are never shared
                part_map[i] = add_list_part();
                parts[part_map[i]].flags = o.parts[i].flags;
                continue;
            }
            PartQuery q = o.parts[i];
            if (q.postings != vdev::kNoValue) q.postings = post_map[q.postings];
            if (o.part_hash[i] == 0) {  // bounded parts neither
                part_map[i] = (uint32_t)parts.size();
                parts.push_back(q);
                part_dict.push_back(dict_map[o.part_dict[i]]);
                part_hash.push_back(0);
                continue;
            }
            part_map[i] = find_or_add_part(q, dict_map[o.part_dict[i]], o.part_hash[i]);
        }
        for (BoundedPart& bp : o.bounded) bounded.push_back(BoundedPart{part_map[bp.part], std::move(bp.request)});
        for (RegexPart& rp : o.regex_parts) regex_parts.push_back(RegexPart{part_map[rp.part], std::move(rp.pattern), rp.case_insensitive, rp.starts_with});
        const uint32_t leaf_base = (uint32_t)leaf_part.size(), prog_base = (uint32_t)prog.size(), boost_base = (uint32_t)boosts.size(), facet_base = (uint32_t)facets.size();
        for (vdev::PhraseMember m : o.phrase_members) {
            m.part1 = part_map[m.part1], m.part2 = part_map[m.part2], m.list_part = part_map[m.list_part];
            phrase_members.push_back(m);
        }
        const uint32_t tl_term_base = (uint32_t)tl_term_parts.size(), request_base = (uint32_t)requests.size();
        for (uint32_t p : o.tl_term_parts) tl_term_parts.push_back(part_map[p]);
        for (vdev::TlInstance t : o.tl_instances) {
            t.list_part = part_map[t.list_part], t.term_begin += tl_term_base, t.request += request_base;
            tl_instances.push_back(t);
        }
        for (vdev::IdsMember m : o.ids_members) {
            m.part = part_map[m.part], m.list_part = part_map[m.list_part];
            ids_members.push_back(m);
        }
        for (vdev::BoostListMember m : o.boost_members) {
            m.part = part_map[m.part], m.list_part = part_map[m.list_part];
            boost_members.push_back(m);
        }
        facets.insert(facets.end(), o.facets.begin(), o.facets.end());
        facet_top.insert(facet_top.end(), o.facet_top.begin(), o.facet_top.end());
        facet_text_path.insert(facet_text_path.end(), o.facet_text_path.begin(), o.facet_text_path.end());
        for (uint32_t p : o.leaf_part) leaf_part.push_back(part_map[p]);
        prog.insert(prog.end(), o.prog.begin(), o.prog.end());
        boosts.insert(boosts.end(), o.boosts.begin(), o.boosts.end());
        for (size_t i = 0; i < o.programs.size(); ++i) {
            QueryProgram qp = o.programs[i];
            if (qp.active) qp.leaf_begin += leaf_base, qp.prog_begin += prog_base, qp.boost_begin += boost_base, qp.post_begin += prog_base, qp.facet_begin += facet_base;
            programs.push_back(qp);
            requests.push_back(std::move(o.requests[i]));
            requests.back().facet_begin += facet_base;
        }
        max_leaves = std::max(max_leaves, o.max_leaves);
        max_k = std::max(max_k, o.max_k);
    }

    // A part whose hits_ids are used: they are collected before the per-part bound (search_field.rs:305-307).
    uint32_t add_part_for_ids(const vhost::SearchPart& part) {
        if (!part.top) return add_part(part);
        vhost::SearchPart unbounded = part;
        unbounded.top.reset(), unbounded.skip.reset();
        return add_part(unbounded);
    }

    // A part searched for ids only (filter trees, boost_term): its matched term ids, taken as text ids, resolve to
    // anchors through text_id_to_anchor (search_field.rs:468-498).  Returns the list part that receives them.
    uint32_t add_ids_part(const vhost::SearchPart& part) {
        std::string path = part.path;
        if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
        vdev::IdsMember m;
        memset(&m, 0, sizeof m);
        m.part = add_part_for_ids(part);
        m.identity = ix->host->is_anchor_identity_column(path) ? 1u : 0u;
        if (!m.identity) {
            auto it = ix->stores.find(path + ".text_id_to_anchor");
            if (it == ix->stores.end()) ix->host->path_not_found(path + ".text_id_to_anchor");
            m.text_id_to_anchor = it->second.view();
        }
        m.list_part = add_list_part();
        ids_members.push_back(m);
        return m.list_part;
    }

    struct Node {  // emitted subtree
        std::string term;  // request.terms[0] the result carries (set_op.rs:122-124)
        int leaf = -1;     // leaf index when the subtree is a single search part
        bool empty = false;
    };

    // state of the request being planned
    std::vector<uint32_t> scratch_raw, scratch_low;    // decoded term of the part being added
    std::vector<uint32_t> scratch_leaves, scratch_code, scratch_post;
    std::vector<BoostStep> scratch_steps;
    bool cur_want_terms = false;                       // text locality asked for: remember the tree's (path, term, part)
    std::vector<BoostStep>* cur_steps = nullptr;       // its boost steps (1:n boosts are appended here)
    std::vector<std::string> cur_ids_keys;             // parts that are also searched for ids (filter, phrase boosts)
    bool cur_leaf_boost = false;
    uint32_t cur_n_leaf_boosts = 0;
    const vhost::SearchRequest* cur_root = nullptr;    // the request's root node
    uint32_t cur_must_mask = 0;                        // leaves that are direct search-part children of a root `and`

    Node emit(const vhost::SearchRequest& r, std::vector<uint32_t>& leaves, std::vector<uint32_t>& code, bool ids_only = false,
              const std::vector<vhost::BoostPart>& boosts = std::vector<vhost::BoostPart>()) {
        if (r.kind == vhost::SearchRequest::Search) {
            uint32_t part = ids_only ? add_ids_part(r.part) : add_part(r.part);
            Node n;
            n.term = r.part.terms[0];
            n.leaf = (int)leaves.size();
            leaves.push_back(part);
            if (!ids_only && (cur_want_terms || !boosts.empty())) {
                std::string path = r.part.path;
                if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
                if (cur_want_terms) tl_seen.push_back(TlTerm{path, r.part.terms[0], part});
                // a boost on the same 1:n level as the part (execution_plan.rs:422-436): BoostToAnchor + ApplyAnchorBoost
                const size_t pos = boosts.empty() ? std::string::npos : r.part.path.rfind("[]");
                if (pos != std::string::npos) {
                    const std::string end_obj = r.part.path.substr(0, pos);
                    const vhost::BoostPart* found = nullptr;
                    for (auto& b : boosts) {
                        const size_t bp = b.path.rfind("[]");
                        if (bp == std::string::npos || b.path.substr(0, bp) != end_obj) continue;
                        if (found) throw InvalidRequest("more than one boost on the same 1:n level");
                        found = &b;
                    }
                    if (found) {
                        // apply_boost_values_anchor (boost.rs:255-281) depends on the run of boosted hits before an anchor; a
                        // run that starts in the previous anchor-range shard cannot be seen from this one
                        if (ix->n_shards > 1) throw Unsupported("1:n boosts (a boost on the search part's own [] level) are outside the accelerated path on a sharded index");
                        vdev::BoostListMember m;
                        memset(&m, 0, sizeof m);
                        m.part = part;
                        m.tokenized = ix->host->is_tokenized(path) ? 1u : 0u;
                        m.use_ids = std::find(cur_ids_keys.begin(), cur_ids_keys.end(), r.part.key()) != cur_ids_keys.end() ? 1u : 0u;
                        auto need = [&](const std::string& p) -> const vdev::CsrDev& {
                            auto it = ix->stores.find(p);
                            if (it == ix->stores.end()) ix->host->path_not_found(p);
                            return it->second;
                        };
                        if (m.tokenized) m.tokens_to_text_id = need(path + ".tokens_to_text_id").view();
                        m.value_id_to_parent = need(r.part.path + ".textindex.value_id_to_parent").view();
                        auto col = ix->boosts.find(found->path + ".boost_valid_to_value");
                        if (col == ix->boosts.end()) ix->host->path_not_found(found->path + ".boost_valid_to_value");
                        m.column = col->second.bits.p, m.column_n = (uint32_t)col->second.n;
                        m.value_id_to_anchor = need(found->path + ".value_id_to_anchor").view();
                        m.list_part = add_list_part();
                        parts[m.list_part].flags |= vdev::kPartListBoost;
                        boost_members.push_back(m);
                        if (++cur_n_leaf_boosts > vdev::kMaxLeafBoosts) throw Unsupported("more than 4 search parts with a 1:n boost in one request");
                        BoostStep step;
                        memset(&step, 0, sizeof step);
                        step.column = col->second.bits.p, step.n = (uint32_t)col->second.n;
                        step.fun = (uint32_t)found->boost_fun;
                        step.param = found->param.value_or(0.0f);
                        step.list_only = 1;
                        if (found->expression) parse_expression(*found->expression, step);
                        const uint32_t step_index = (uint32_t)cur_steps->size();
                        cur_steps->push_back(step);
                        const uint32_t list_leaf = (uint32_t)leaves.size();
                        leaves.push_back(m.list_part);
                        code.push_back(vdev::kOpLeafBoost), code.push_back((uint32_t)n.leaf), code.push_back(list_leaf), code.push_back(step_index);
                        cur_leaf_boost = true;
                        return n;
                    }
                }
            }
            code.push_back(vdev::kOpLeaf);
            code.push_back((uint32_t)n.leaf);
            return n;
        }
        if (r.queries.empty()) throw Unsupported("empty or/and list");
        std::vector<Node> kids;
        kids.reserve(r.queries.size());
        for (auto& q : r.queries) {
            if (q.get_boost() && !q.get_boost()->empty()) {  // merge_vec (execution_plan.rs:263-270)
                std::vector<vhost::BoostPart> b = boosts;
                b.insert(b.end(), q.get_boost()->begin(), q.get_boost()->end());
                kids.push_back(emit(q, leaves, code, ids_only, b));
            } else {
                kids.push_back(emit(q, leaves, code, ids_only, boosts));
            }
        }
        if (kids.size() == 1) return kids[0];  // passthrough (set_op.rs:93-96, :371-374)
        if (kids.size() > vdev::kMaxLeaves) throw Unsupported("more than " + std::to_string(vdev::kMaxLeaves) + " sub-queries in one or/and");
        Node out;
        out.term = kids[0].term;
        if (r.kind == vhost::SearchRequest::Or) {
            // term slots of the union (set_op.rs:122-137): the distinct terms in sorted order
            const std::string* terms[vdev::kMaxLeaves];
            size_t n_terms = 0;
            for (auto& k : kids) terms[n_terms++] = &k.term;
            auto less = [](const std::string* a, const std::string* b) { return *a < *b; };
            auto same = [](const std::string* a, const std::string* b) { return *a == *b; };
            std::sort(terms, terms + n_terms, less);
            n_terms = (size_t)(std::unique(terms, terms + n_terms, same) - terms);
            code.push_back(vdev::kOpUnion);
            code.push_back((uint32_t)kids.size());
            code.push_back((uint32_t)n_terms);
            for (auto& k : kids) code.push_back((uint32_t)(std::lower_bound(terms, terms + n_terms, &k.term, less) - terms));
        } else {
            if (&r == cur_root && !ids_only)
                for (auto& k : kids)
                    if (k.leaf >= 0 && k.leaf < 32) cur_must_mask |= 1u << k.leaf;
            code.push_back(vdev::kOpIntersect);
            code.push_back((uint32_t)kids.size());
            for (size_t i = 0; i < kids.size(); ++i) code.push_back((uint32_t)i);  // sum order, patched on the device
            for (auto& k : kids) code.push_back(k.leaf >= 0 ? (uint32_t)k.leaf : vdev::kNoValue);
        }
        return out;
    }

    std::vector<std::pair<std::string, const vdev::ColumnDev*>> columns_seen;  // boost path -> its column, per plan
    BoostStep make_boost(const vhost::BoostPart& b) {
        BoostStep s;
        memset(&s, 0, sizeof s);
        const vdev::ColumnDev* found = nullptr;
        for (auto& seen : columns_seen)
            if (seen.first == b.path) found = seen.second;
        if (!found) {
            const std::string path = b.path + ".boost_valid_to_value";
            auto it = ix->boosts.find(path);
            if (it == ix->boosts.end()) ix->host->path_not_found(path);
            found = &it->second;
            columns_seen.emplace_back(b.path, found);
        }
        const vdev::ColumnDev& col = *found;
        s.column = col.bits.p;
        s.levels = col.level_hdr.p;
        s.n = (uint32_t)col.n;
        s.fun = (uint32_t)b.boost_fun;
        s.param = b.param.value_or(0.0f);
        if (b.skip_when_score) {
            if (b.skip_when_score->size() > vdev::kMaxSkipWhenScore) throw Unsupported("more than " + std::to_string(vdev::kMaxSkipWhenScore) + " skip_when_score values");
            s.n_skip = (uint32_t)b.skip_when_score->size();
            for (size_t i = 0; i < b.skip_when_score->size(); ++i) s.skip[i] = (*b.skip_when_score)[i];
        }
        if (b.expression) parse_expression(*b.expression, s);
        // upper bound of the multiplier (used by the tile kernel to skip hopeless anchors)
        s.can_prune = 0, s.max_mult = 0.0f;
        if (!b.expression && s.n_skip == 0 && col.non_negative && s.param >= 0.0f && std::isfinite(s.param) && std::isfinite(col.vmax)) {
            float m = 0.0f;
            bool ok = true;
            switch (b.boost_fun) {
                case vhost::BoostFun::Log10: m = log10f(col.vmax + s.param); break;
                case vhost::BoostFun::Log2: m = log2f(col.vmax + s.param); break;
                case vhost::BoostFun::Multiply: m = col.vmax + s.param; break;
                default: ok = false; break;
            }
            if (ok && std::isfinite(m)) {
                m = m + fabsf(m) * 1e-6f + 1e-30f;  // device log10f/log2f may differ from the host's by a few ulp
                s.can_prune = 1;
                s.max_mult = std::max(m, 1.0f);  // anchors without a boost value keep their score
            }
        }
        return s;
    }

    // get_facet's index walk (facet.rs:31-83): one id -> ids join, or the chain of parent_to_value_id joins
    void add_facet(const vhost::FacetRequest& fr) {
        const std::vector<std::string> steps = vfmt::get_steps_to_anchor(fr.field);
        std::vector<std::string> paths;
        if (steps.size() == 1 || ix->host->has_index(steps.back() + ".anchor_to_text_id"))
            paths.push_back(steps.size() == 1 ? steps.front() + ".parent_to_value_id" : steps.back() + ".anchor_to_text_id");
        else
            for (auto& st : steps) paths.push_back(st + ".parent_to_value_id");
        if (paths.size() > vdev::kMaxFacetSteps) throw Unsupported("facet fields nested deeper than three levels are outside the accelerated path");
        vdev::FacetStep fs;
        memset(&fs, 0, sizeof fs);
        fs.n_steps = (uint32_t)paths.size();
        for (size_t i = 0; i < paths.size(); ++i) {
            auto it = ix->stores.find(paths[i]);
            if (it == ix->stores.end()) ix->host->path_not_found(paths[i]);
            fs.step[i] = it->second.view();
            fs.hist_size = it->second.n_values;
        }
        const uint64_t top = fr.top ? *fr.top : (uint64_t)fs.hist_size;
        // the groups are picked one arg-max round each (facet_topk_kernel): fine for the usual handful, seconds beyond this
        if (std::min<uint64_t>(top, fs.hist_size) > 65536) throw Unsupported("facets with more than 65536 groups are outside the accelerated path");
        facets.push_back(fs);
        facet_top.push_back((uint32_t)std::min<uint64_t>(top, fs.hist_size));
        facet_text_path.push_back(steps.back());
    }

    // a single boost step without skip list or expression travels inside the QueryProgram
    static void set_fast_boost(QueryProgram& qp, const std::vector<BoostStep>& steps) {
        qp.fb_flags = 0;
        if (steps.size() != 1 || steps[0].n_skip != 0 || steps[0].expr_op != vdev::kExprNone || steps[0].list_only) return;
        const BoostStep& s = steps[0];
        qp.fb_flags = 1u | ((s.can_prune && s.max_mult > 0.0f) ? 2u : 0u);
        qp.fb_col = s.column, qp.fb_lev = s.levels, qp.fb_n = s.n, qp.fb_fun = s.fun, qp.fb_param = s.param, qp.fb_max_mult = s.max_mult;
    }

    static void collect_keys(const vhost::SearchRequest& r, std::vector<std::string>& out) {
        if (r.kind == vhost::SearchRequest::Search) out.push_back(r.part.key());
        for (auto& q : r.queries) collect_keys(q, out);
    }

    void plan_request(const vhost::Request& request, RequestPlan& rp, QueryProgram& qp) {
        rp.top = request.top.value_or(10);  // search.rs:146
        rp.skip = request.skip.value_or(0);
        rp.why_found = request.why_found;
        rp.select = request.select;
        if (vhost::request_explains(request)) rp.explain = std::make_shared<const vhost::Request>(request), rp.explain_asked = true;
        if (!request.search_req) throw InvalidRequest("search_req is None, but is required in search");
        {  // the top-k heap: merged in shared memory up to 256 keys, in global memory up to 4096 (one device only: shards gather every heap)
            const uint64_t limit = ix->n_shards > 1 ? vdev::kMaxK : vdev::kMaxKLarge;
            if (rp.top > limit || rp.skip > limit || rp.top + rp.skip > limit)
                throw Unsupported("top + skip above " + std::to_string(limit) + (ix->n_shards > 1 ? " on a sharded index" : "") + " is outside the accelerated path");
        }

        std::vector<uint32_t>&leaves = scratch_leaves, &code = scratch_code, &post = scratch_post;
        leaves.clear(), code.clear(), post.clear();
        const vhost::SearchRequest& root = *request.search_req;
        tl_seen.clear();
        cur_want_terms = request.text_locality;
        std::vector<BoostStep>& steps = scratch_steps;
        steps.clear();
        cur_steps = &steps, cur_leaf_boost = false, cur_n_leaf_boosts = 0;
        cur_root = &root, cur_must_mask = 0;
        cur_ids_keys.clear();
        if (request.phrase_boosts)
            for (auto& pb : *request.phrase_boosts) cur_ids_keys.push_back(pb.search1.key()), cur_ids_keys.push_back(pb.search2.key());
        if (request.filter) collect_keys(*request.filter, cur_ids_keys);
        {
            // the request's boosts go down the tree; sub-queries add their own options (execution_plan.rs:263-270), the root's own do not apply
            static const std::vector<vhost::BoostPart> none;
            emit(root, leaves, code, false, request.boost ? *request.boost : none);
        }
        cur_want_terms = false;  // the filter tree and the extra parts below are not part of term_id_hits_in_field
        const std::vector<TlTerm> tree_terms = std::move(tl_seen);
        tl_seen.clear();
        bool extras = cur_leaf_boost;
        if (request.filter) {  // the filter tree is evaluated for presence only; hits outside it are dropped (set_op.rs:311-326)
            emit(*request.filter, leaves, code, true);
            code.push_back(vdev::kOpFilter);
            extras = true;
        }
        if (request.phrase_boosts && !request.phrase_boosts->empty()) {
            // add_phrase_boost_plan_steps (execution_plan.rs:202-262) + sort_and_group_boosts_by_phrase_terms (plan_steps.rs:235-258):
            // the entries with the same (term1, term2) texts form one group; a hit in the group's anchors is multiplied by 5.0
            struct Entry {
                std::string t1, t2;
                uint32_t p1, p2;
                vdev::PhraseView store;
            };
            std::vector<Entry> entries;
            for (auto& pb : *request.phrase_boosts) {
                if (pb.search1.path != pb.search2.path) throw InvalidRequest("phrase boost parts must be on the same path");
                if (pb.search1.terms.empty() || pb.search2.terms.empty()) throw InvalidRequest("search part without terms");
                std::string path = pb.search1.path;
                if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
                if (!vfmt::ends_with(path, ".phrase_pair_to_anchor")) path += ".phrase_pair_to_anchor";
                auto it = ix->phrases.find(path);
                if (it == ix->phrases.end()) ix->host->path_not_found(path);
                entries.push_back(Entry{pb.search1.terms[0], pb.search2.terms[0], add_part_for_ids(pb.search1), add_part_for_ids(pb.search2), it->second.view()});
            }
            std::stable_sort(entries.begin(), entries.end(), [](const Entry& a, const Entry& b) { return std::tie(a.t1, a.t2) < std::tie(b.t1, b.t2); });
            for (size_t i = 0; i < entries.size();) {
                const uint32_t list_part = add_list_part();
                const uint32_t leaf = (uint32_t)leaves.size();
                leaves.push_back(list_part);
                size_t j = i;
                for (; j < entries.size() && entries[j].t1 == entries[i].t1 && entries[j].t2 == entries[i].t2; ++j)
                    phrase_members.push_back(vdev::PhraseMember{entries[j].p1, entries[j].p2, list_part, 0u, entries[j].store});
                const float five = 5.0f;
                uint32_t bits;
                memcpy(&bits, &five, 4);
                post.push_back(vdev::kPostMulIfPresent), post.push_back(leaf), post.push_back(bits);
                i = j;
            }
            extras = true;
        }
        if (request.boost_term)  // search.rs:176, boost.rs:89-195: hits that the part also finds are multiplied by its boost (default 2.0)
            for (auto& part : *request.boost_term) {
                const uint32_t list_part = add_ids_part(part);
                const uint32_t leaf = (uint32_t)leaves.size();
                leaves.push_back(list_part);
                const float v = part.boost.value_or(2.0f);
                uint32_t bits;
                memcpy(&bits, &v, 4);
                post.push_back(vdev::kPostMulIfPresent), post.push_back(leaf), post.push_back(bits);
                extras = true;
            }
        if (request.text_locality) {
            // boost_text_locality_all (boost.rs:11-32): per field with at least two query terms, one list; the smallest boost wins
            std::map<std::string, std::map<std::string, uint32_t>> by_path;  // later parts replace earlier ones (set_op.rs:29-47)
            for (auto& t : tree_terms) by_path[t.path][t.term] = t.part;
            uint32_t list_part = vdev::kNoValue;
            for (auto& kv : by_path) {
                if (kv.second.size() <= 1) continue;
                vdev::TlInstance t;
                memset(&t, 0, sizeof t);
                auto t2t = ix->stores.find(kv.first + ".tokens_to_text_id");
                if (t2t == ix->stores.end()) ix->host->path_not_found(kv.first + ".tokens_to_text_id");
                t.tokens_to_text_id = t2t->second.view();
                t.identity = ix->host->is_anchor_identity_column(kv.first) ? 1u : 0u;
                if (!t.identity) {
                    auto t2a = ix->stores.find(kv.first + ".text_id_to_anchor");
                    if (t2a == ix->stores.end()) ix->host->path_not_found(kv.first + ".text_id_to_anchor");
                    t.text_id_to_anchor = t2a->second.view();
                }
                if (list_part == vdev::kNoValue) list_part = add_list_part();
                t.list_part = list_part;
                t.term_begin = (uint32_t)tl_term_parts.size(), t.n_terms = (uint32_t)kv.second.size();
                t.request = (uint32_t)requests.size();
                for (auto& tp : kv.second) tl_term_parts.push_back(tp.second);
                tl_instances.push_back(t);
            }
            if (list_part != vdev::kNoValue) {
                const uint32_t leaf = (uint32_t)leaves.size();
                leaves.push_back(list_part);
                post.push_back(vdev::kPostMulValue), post.push_back(leaf);
                extras = true;
            }
        }
        if (leaves.size() > vdev::kMaxLeaves) throw Unsupported("more than " + std::to_string(vdev::kMaxLeaves) + " search parts in one request");
        if (request.facets && !request.facets->empty()) {
            rp.facets = *request.facets;
            rp.has_facets = true;
            rp.facet_begin = (uint32_t)facets.size();
            for (auto& fr : rp.facets) add_facet(fr);
            extras = true;
        }

        if (request.boost)
            for (auto& b : *request.boost)  // anchor-level boosts run after the root (execution_plan.rs:175-189); 1:n boosts were matched to their parts above
                if (b.path.find("[]") == std::string::npos) steps.push_back(make_boost(b));

        // flat `or` of search parts with pairwise distinct terms: leaves in slot (= sorted term) order, no program
        bool flat = false;
        if (extras) {
            // filters, post ops and facets run on the program path
        } else if (leaves.size() == 1 && code.size() == 2) {
            flat = true;
        } else if (root.kind == vhost::SearchRequest::Or && code.size() == 2 * leaves.size() + 3 + leaves.size() && code[2 * leaves.size()] == vdev::kOpUnion) {
            // Parts that share a term slot must be the very same search part (then their hit lists are identical and
            // max(x, x) = x, so one leaf stands for the slot).
            const uint32_t n_slots = code[2 * leaves.size() + 2];
            std::vector<uint32_t> by_slot(n_slots, vdev::kNoValue);
            bool same = true;
            for (size_t c = 0; c < leaves.size() && same; ++c) {
                uint32_t& slot = by_slot[code[2 * leaves.size() + 3 + c]];
                if (slot == vdev::kNoValue) slot = leaves[c];
                else if (slot != leaves[c]) same = false;
            }
            if (same) {
                if (by_slot.size() == 1 && leaves.size() > 1) qp.union1 = 1;  // [A, A, ..]: still a union (n * n with n in {0, 1}), not the one-input passthrough
                leaves = by_slot;
                flat = true;
            }
        }
        qp.leaf_begin = (uint32_t)leaf_part.size();
        qp.n_leaves = (uint32_t)leaves.size();
        leaf_part.insert(leaf_part.end(), leaves.begin(), leaves.end());
        qp.prog_begin = (uint32_t)prog.size();
        qp.prog_len = flat ? 0u : (uint32_t)code.size();
        if (!flat) prog.insert(prog.end(), code.begin(), code.end());
        qp.post_begin = (uint32_t)prog.size();
        qp.post_len = (uint32_t)post.size();
        prog.insert(prog.end(), post.begin(), post.end());
        qp.n_leaf_boosts = cur_n_leaf_boosts;
        qp.must_mask = cur_must_mask;
        qp.facet_begin = rp.facet_begin;
        qp.n_facets = (uint32_t)rp.facets.size();
        qp.boost_begin = (uint32_t)boosts.size();
        qp.n_boosts = (uint32_t)steps.size();
        boosts.insert(boosts.end(), steps.begin(), steps.end());
        set_fast_boost(qp, steps);
        qp.k = (uint32_t)(rp.top + rp.skip);
        qp.active = 1;
        qp.nonneg = 1;
        for (uint32_t part : leaves)
            if (!(parts[part].boost >= 0.0f) || (parts[part].flags & vdev::kPartAnySign)) qp.nonneg = 0;
        max_leaves = std::max<uint32_t>(max_leaves, qp.n_leaves);
        max_k = std::max<uint32_t>(max_k, std::max<uint32_t>(qp.k, 1));
    }

    // Request JSON -> Request: independent of the batch, so a batch parses its requests on several threads.
    struct Parsed {
        vhost::Request request;
        int32_t status = 0;
        std::string message;
    };
    static void parse_into(const char* json, Parsed& out) {
        try {
            out.request = vhost::read_request_json(json, strlen(json));
        } catch (const vhost::RequestError& e) {
            out.status = 5, out.message = e.what();
        } catch (const std::exception& e) {
            out.status = 9, out.message = e.what();
        }
    }

    void add_request(const char* json) {
        Parsed p;
        parse_into(json, p);
        add_parsed(p);
    }

    void add_parsed(const Parsed& parsed) {
        RequestPlan rp;
        QueryProgram qp;
        memset(&qp, 0, sizeof qp);
        // a failing request must not leave half-registered leaves behind
        const size_t leaf_mark = leaf_part.size(), prog_mark = prog.size(), boost_mark = boosts.size(), facet_mark = facets.size(), phrase_mark = phrase_members.size(), ids_mark = ids_members.size(), tl_mark = tl_instances.size(),
                     tl_term_mark = tl_term_parts.size(), bm_mark = boost_members.size(), bounded_mark = bounded.size(), regex_mark = regex_parts.size();
        try {
            if (parsed.status != 0) {
                rp.status = parsed.status, rp.message = parsed.message;
            } else {
                plan_request(parsed.request, rp, qp);
            }
        } catch (const InvalidRequest& e) {
            rp.status = 1, rp.message = e.what();
        } catch (const vhost::FstNotFound& e) {
            rp.status = 2, rp.message = e.what();
        } catch (const vhost::PathNotFound& e) {
            rp.status = 3, rp.message = e.what();
        } catch (const vhost::RequestError& e) {
            rp.status = 5, rp.message = e.what();
        } catch (const Unsupported& e) {
            rp.status = 8, rp.message = e.what();
        } catch (const std::exception& e) {
            rp.status = 9, rp.message = e.what();
        }
        if (rp.status != 0) {
            leaf_part.resize(leaf_mark), prog.resize(prog_mark), boosts.resize(boost_mark);
            facets.resize(facet_mark), facet_top.resize(facet_mark), facet_text_path.resize(facet_mark);
            phrase_members.resize(phrase_mark), ids_members.resize(ids_mark), tl_instances.resize(tl_mark), tl_term_parts.resize(tl_term_mark), boost_members.resize(bm_mark);
            // per-part-top parts of the failed request: nothing references them any more, so they must not be matched and
            // bounded on every execute (their PartQuery stays behind as an injected part without hits)
            bounded.resize(bounded_mark);
            regex_parts.resize(regex_mark);
            rp.facets.clear(), rp.has_facets = false;
            memset(&qp, 0, sizeof qp);
        }
        requests.push_back(std::move(rp));
        programs.push_back(qp);
    }
};

}  // namespace vplan
