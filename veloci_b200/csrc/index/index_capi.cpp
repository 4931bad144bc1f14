// C entry points of the index-building helper library (libveloci_b200_index.so).
// Test/bench infrastructure: produces index directories in the reference's
// on-disk layout; not on the accelerated query path.
#include <cstring>
#include <sstream>

#include "../host/explain_plan.hpp"
#include "../host/explain_walk.hpp"
#include "../host/field_highlight.hpp"
#include "../host/highlight.hpp"
#include "../host/part_hits.hpp"
#include "../host/query_generator.hpp"
#include "../host/read_document.hpp"
#include "../host/regex_dfa.hpp"
#include "../host/request.hpp"
#include "indexer.hpp"
#include "synth.hpp"

extern "C" {

static void set_err(char* err, size_t n, const char* msg) {
    if (err && n) snprintf(err, n, "%s", msg);
}

// Persistence::create_mmap + create_indices_from_str (src/create.rs:929-965):
// `jsonl` is one JSON document per line (or a stream of JSON values), `config`
// the field configuration as JSON.
int vidx_create_from_jsonl(const char* dir, const char* jsonl, const char* config, char* err, size_t errlen) {
    try {
        vindex::Indexer ix(dir, config ? config : "{}");
        vjson::Parser parser(jsonl, strlen(jsonl));
        vjson::Value v;
        while (parser.parse_next(v)) {
            if (v.is_array()) {
                for (auto& e : v.arr) ix.add_document(e);
            } else {
                ix.add_document(v);
            }
        }
        ix.finish();
        return 0;
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return 1;
    }
}

// add_token_values_to_tokens (src/create/token_values_to_tokens.rs:26-82): `data_json` = [{"text": ..., "value": f32 | null}],
// `config_json` = {"path": <field>}.  Every text is looked up exactly (levenshtein 0, case-sensitive) in the field's
// dictionary; the found term ids get `value.to_bits()` in the 1:1 packed store
// `<path>.textindex.token_values.boost_valid_to_value` (Boost, SingleValue, U32), appended to the column's indices.
int vidx_add_token_values(const char* dir, const char* data_json, const char* config_json, char* err, size_t errlen) {
    try {
        std::unique_ptr<vhost::Persistence> p = vhost::Persistence::load(dir);
        const vjson::Value data = vjson::parse(data_json, strlen(data_json));
        const vjson::Value config = vjson::parse(config_json, strlen(config_json));
        const vjson::Value* path_v = config.get("path");
        if (!data.is_array() || !path_v || !path_v->is_string()) throw std::runtime_error("token values: expected an array of {text, value} and a config {path}");
        const std::string field = path_v->str;
        auto fst = p->fst.find(field + ".textindex");
        if (fst == p->fst.end()) throw std::runtime_error("field does not exist " + field + ".textindex (fst not found)");
        vfmt::PackedWriter w;
        std::vector<std::pair<uint32_t, uint32_t>> entries;
        for (const vjson::Value& el : data.arr) {
            const vjson::Value* text = el.get("text");
            const vjson::Value* value = el.get("value");
            if (!text || !text->is_string()) throw std::runtime_error("token values: entry without text");
            if (!value || value->is_null()) continue;
            if (!value->is_number()) throw std::runtime_error("token values: value must be a number");
            uint64_t id = 0;
            if (!fst->second.get(text->str, id)) continue;
            const float f = (float)value->num;
            uint32_t bits;
            memcpy(&bits, &f, 4);
            entries.emplace_back((uint32_t)id, bits);
        }
        std::stable_sort(entries.begin(), entries.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        for (auto& e : entries) w.add(e.first, e.second);
        const std::string path = field + ".textindex.token_values.boost_valid_to_value";
        const std::vector<uint8_t> bytes = w.encode();
        vhost::write_file(std::string(dir) + "/" + path, bytes.data(), bytes.size());
        vhost::IndexMetadata im;
        im.path = path;
        im.category = vhost::IndexCategory::Boost;
        im.cardinality = vhost::IndexCardinality::SingleValue;
        im.is_empty = w.cache.empty();
        im.meta = w.meta;
        vhost::Metadata meta = p->metadata;
        auto col = meta.columns.find(field);
        if (col == meta.columns.end()) {
            vhost::FieldInfo fi;
            fi.has_fst = false;
            col = meta.columns.emplace(field, fi).first;
        }
        col->second.indices.push_back(im);
        const std::string mj = vjson::to_string(vhost::metadata_to_json(meta), 2);
        vhost::write_file(std::string(dir) + "/metaData.json", mj.data(), mj.size());
        return 0;
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return 1;
    }
}

// The product's host arithmetic on a part's term hits (host/part_hits.hpp: bound_part_hits, apply_token_value), for the CPU
// tests that hold it against the oracle's get_term_ids_in_field: `hits_json` = [[term id, score], ...] as the device match
// delivers them for the part without top / skip / boost / token_value; writes the part's final hits in the same form.
int vidx_bound_part_hits(const char* dir, const char* part_json, const char* hits_json, char* out, size_t outlen) {
    try {
        std::unique_ptr<vhost::Persistence> p = vhost::Persistence::load(dir);
        const vhost::SearchPart part = vhost::parse_search_part(vjson::parse(part_json, strlen(part_json)));
        const vjson::Value hv = vjson::parse(hits_json, strlen(hits_json));
        std::vector<vdev::TermHit> hits;
        for (const vjson::Value& h : hv.arr) hits.push_back(vdev::TermHit{(uint32_t)h.arr.at(0).num, (float)h.arr.at(1).num});
        std::sort(hits.begin(), hits.end(), [](const vdev::TermHit& a, const vdev::TermHit& b) { return a.id < b.id; });
        vdev::bound_part_hits(part, hits);
        vdev::apply_token_value(*p, part, hits);
        std::string text = "[";
        for (size_t i = 0; i < hits.size(); ++i) {
            char buf[64];
            snprintf(buf, sizeof buf, "%s[%u,%.9g]", i ? "," : "", hits[i].id, (double)hits[i].score);
            text += buf;
        }
        text += "]";
        if (text.size() + 1 > outlen) throw std::runtime_error("output buffer too small");
        set_err(out, outlen, text.c_str());
        return 0;
    } catch (const std::exception& e) {
        set_err(out, outlen, e.what());
        return 1;
    }
}

// The product's explain walk (host/explain_walk.hpp) without a device, for the CPU tests that hold it against the oracle's
// explain: `anchors_json` = the ids of the request's returned hits, `leaves_json` = per search part of the tree (tree order)
// the [[term id, score], ...] hits of the bare part, which the device match delivers in the product; the posting weights
// come from the index files here (the product looks them up on the device).  Writes {"<anchor>": [Explain, ...]}.
int vidx_explain_walk(const char* dir, const char* request_json, const char* anchors_json, const char* leaves_json, char* out, size_t outlen) {
    try {
        std::unique_ptr<vhost::Persistence> p = vhost::Persistence::load(dir);
        const vhost::Request request = vhost::read_request_json(request_json, strlen(request_json));
        const vjson::Value av = vjson::parse(anchors_json, strlen(anchors_json));
        const vjson::Value lv = vjson::parse(leaves_json, strlen(leaves_json));
        std::vector<uint32_t> anchors;
        for (const vjson::Value& a : av.arr) anchors.push_back((uint32_t)a.num);
        vexplain::Walk walk(*p, request, anchors);
        if (lv.arr.size() != walk.n_parts()) throw std::runtime_error("explain walk: one hit list per search part expected");
        for (size_t i = 0; i < walk.n_parts(); ++i) {
            std::vector<vdev::TermHit> raw;
            for (const vjson::Value& h : lv.arr[i].arr) raw.push_back(vdev::TermHit{(uint32_t)h.arr.at(0).num, (float)h.arr.at(1).num});
            std::sort(raw.begin(), raw.end(), [](const vdev::TermHit& a, const vdev::TermHit& b) { return a.id < b.id; });
            const std::vector<vdev::TermHit>& hits = walk.set_hits(i, std::move(raw));
            std::string path = walk.part(i).path;
            if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
            const vfmt::AnchorScoreView& store = p->get_token_to_anchor(path);
            std::vector<float> weight(hits.size() * anchors.size(), -1.0f);
            uint64_t postings = 0;
            for (size_t t = 0; t < hits.size(); ++t)
                store.for_each(hits[t].id, [&](uint32_t anchor, uint32_t raw_score) {
                    ++postings;
                    for (size_t a = 0; a < anchors.size(); ++a)
                        if (anchors[a] == anchor) weight[t * anchors.size() + a] = (float)(_Float16)(float)raw_score / 100.0f;  // AnchorScore keeps an f16 (search_field.rs:426)
                });
            walk.set_weights(i, std::move(weight), postings);
        }
        const std::string text = walk.to_json();
        if (text.size() + 1 > outlen) throw std::runtime_error("output buffer too small");
        set_err(out, outlen, text.c_str());
        return 0;
    } catch (const vplan::Unsupported& e) {
        set_err(out, outlen, e.what());
        return 8;
    } catch (const std::exception& e) {
        set_err(out, outlen, e.what());
        return 1;
    }
}

// The host half of search_field::highlight (host/field_highlight.hpp) without a device: `hits_json` = the part's
// [[term id, score], ...] as the device match delivers them for the bare part (terms normalised); writes
// [[highlighted text, score, text id], ...].  what = 1: only normalize_text of `part_json` (a JSON string).
int vidx_field_highlight(const char* dir, const char* part_json, const char* hits_json, int what, char* out, size_t outlen) {
    try {
        std::string text;
        if (what == 1) {
            const vjson::Value v = vjson::parse(part_json, strlen(part_json));
            vjson::write_string(text, vhost::normalize_text(v.str));
        } else {
            std::unique_ptr<vhost::Persistence> p = vhost::Persistence::load(dir);
            const vhost::HighlightRequest req = vhost::parse_highlight_request(vjson::parse(part_json, strlen(part_json)));
            const vjson::Value hv = vjson::parse(hits_json, strlen(hits_json));
            std::vector<vdev::TermHit> hits;
            for (const vjson::Value& h : hv.arr) hits.push_back(vdev::TermHit{(uint32_t)h.arr.at(0).num, (float)h.arr.at(1).num});
            std::sort(hits.begin(), hits.end(), [](const vdev::TermHit& a, const vdev::TermHit& b) { return a.id < b.id; });
            vdev::bound_part_hits(req.part, hits);
            vdev::apply_token_value(*p, req.part, hits);
            text = "[";
            bool first = true;
            for (const vhost::FieldHighlight& h : vhost::highlight_field(*p, req, hits)) {
                text += first ? "[" : ",[";
                first = false;
                vjson::write_string(text, h.text);
                char buf[64];
                snprintf(buf, sizeof buf, ",%.9g,%u]", (double)h.score, h.id);
                text += buf;
            }
            text += "]";
        }
        if (text.size() + 1 > outlen) throw std::runtime_error("output buffer too small");
        set_err(out, outlen, text.c_str());
        return 0;
    } catch (const vplan::InvalidRequest& e) {
        set_err(out, outlen, e.what());
        return 2;
    } catch (const std::exception& e) {
        set_err(out, outlen, e.what());
        return 1;
    }
}

// search::explain_plan of the product's host code (host/explain_plan.hpp), for the CPU tests.
int vidx_explain_plan(const char* request_json, char* out, size_t outlen) {
    try {
        const std::string text = vhost::explain_plan(vhost::read_request_json(request_json, strlen(request_json)));
        if (text.size() + 1 > outlen) throw std::runtime_error("output buffer too small");
        set_err(out, outlen, text.c_str());
        return 0;
    } catch (const std::exception& e) {
        set_err(out, outlen, e.what());
        return 1;
    }
}

// Synthetic Zipfian corpus written straight into the on-disk layout (see synth.hpp).
int vidx_create_synthetic(const char* dir, const char* params_json, char* err, size_t errlen) {
    try {
        vindex::SynthParams p = vindex::SynthParams::from_json(params_json ? params_json : "{}");
        vindex::write_synthetic_index(dir, p);
        return 0;
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return 1;
    }
}

// Writes `n` request JSON lines for the synthetic corpus into `out_path`.
int vidx_write_synthetic_requests(const char* out_path, const char* params_json, char* err, size_t errlen) {
    try {
        vindex::SynthParams p = vindex::SynthParams::from_json(params_json ? params_json : "{}");
        vindex::write_synthetic_requests(out_path, p);
        return 0;
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return 1;
    }
}

// The request surface, for the tests that hold its two parsers against each other: `reader` = 0 parses through the JSON
// DOM (parse_request, what the oracle uses), 1 through the one-pass RequestReader (what the planner uses).  Writes the
// canonical text of the parsed request (or the error message) into `out`; returns 0, or 5 for a RequestError.
int vidx_describe_request(const char* json, int reader, char* out, size_t outlen) {
    try {
        const vhost::Request r = reader ? vhost::read_request_json(json, strlen(json)) : vhost::parse_request_json(json, strlen(json));
        set_err(out, outlen, vhost::describe(r).c_str());
        return 0;
    } catch (const vhost::RequestError& e) {
        set_err(out, outlen, e.what());
        return 5;
    } catch (const std::exception& e) {
        set_err(out, outlen, e.what());
        return 9;
    }
}

// The query language and the request generator (host/query_parser.hpp, host/query_generator.hpp), for the CPU tests:
// `what` 0 = Debug text of the parsed tree, 1 = its phrase pairs as a JSON list, 2 = its terms as a JSON list, 3 = the tokens.
// `options`: bit 0 no_attributes, bit 1 no_parentheses, bit 2 no_levensthein.  Returns 0, or 1 with the ParseError text.
int vidx_query_parse(const char* text, int options, int what, char* out, size_t outlen) {
    try {
        vquery::ParserOptions o;
        o.no_attributes = options & 1, o.no_parentheses = options & 2, o.no_levensthein = options & 4;
        std::string s;
        if (what == 3) {  // the tokens alone: [[text, type], ...]
            const std::string query(text);
            s = "[";
            for (auto& t : vquery::Lexer(query, o).tokens()) {
                if (s.size() > 1) s += ',';
                s += '[';
                vjson::write_string(s, query.substr(t.begin, t.end - t.begin));
                s += ",\"" + std::string(vquery::token_type_name(t.type)) + "\"]";
            }
            s += ']';
            set_err(out, outlen, s.c_str());
            return 0;
        }
        const vquery::Ast ast = vquery::parse(text, o);
        if (what == 0) {
            s = ast.debug();
        } else if (what == 1) {
            s = "[";
            for (auto& p : ast.phrase_pairs()) {
                if (s.size() > 1) s += ',';
                s += '[';
                vjson::write_string(s, p.first);
                s += ',';
                vjson::write_string(s, p.second);
                s += ']';
            }
            s += ']';
        } else {
            s = "[";
            ast.walk_terms(ast.root, [&](const std::string& t) {
                if (s.size() > 1) s += ',';
                vjson::write_string(s, t);
            });
            s += ']';
        }
        set_err(out, outlen, s.c_str());
        return 0;
    } catch (const std::exception& e) {
        set_err(out, outlen, e.what());
        return 1;
    }
}

// UserAST::filter_ast with "drop the leaves whose lower-cased phrase is in `stopwords_json`" (the predicate of
// query_parser_to_veloci_request.rs:117-133): Debug text of what is left, "None" when nothing is.
int vidx_query_filter_stopwords(const char* text, const char* stopwords_json, char* out, size_t outlen) {
    try {
        vquery::Ast ast = vquery::parse(text);
        const vjson::Value words = vjson::parse(stopwords_json, strlen(stopwords_json));
        const int32_t kept = ast.filter(ast.root, [&](const vquery::Ast& a, int32_t i, const std::string*) {
            if (a.nodes[i].kind != vquery::Node::Leaf) return false;
            const std::string low = vfmt::to_lowercase(a.nodes[i].text);
            for (auto& w : words.arr)
                if (w.str == low) return true;
            return false;
        });
        std::string s = "None";
        if (kept >= 0) s.clear(), ast.debug(kept, s);
        set_err(out, outlen, s.c_str());
        return 0;
    } catch (const std::exception& e) {
        set_err(out, outlen, e.what());
        return 1;
    }
}

// search_query (what = 0) / suggest_query (what = 1) over the index in `dir` (host only: metaData.json and the index
// files, no device).  Returns 0 with the request JSON, 1 with the message of a parse error, 2 of a generator error
// (FieldNotFound / AllFieldsFiltered), 5 for parameters that do not deserialize.
int vidx_generate_request(const char* dir, const char* params_json, int what, char* out, size_t outlen) {
    try {
        const auto p = vhost::Persistence::load(dir);
        const vquery::FieldCatalog cat = vquery::FieldCatalog::of(*p);
        const std::string s = what ? vquery::suggest_query_json(cat, params_json, strlen(params_json)) : vquery::search_query_json(cat, params_json, strlen(params_json));
        set_err(out, outlen, s.c_str());
        return 0;
    } catch (const vquery::ParseError& e) {
        set_err(out, outlen, e.what());
        return 1;
    } catch (const vquery::GeneratorError& e) {
        set_err(out, outlen, e.what());
        return 2;
    } catch (const vquery::ParamsError& e) {
        set_err(out, outlen, e.what());
        return 5;
    } catch (const std::exception& e) {
        set_err(out, outlen, e.what());
        return 9;
    }
}

// The product's regex DFA (host/regex_dfa.hpp) on the host: which of `terms_json` (a JSON list of strings) the DFA of
// `pattern` accepts, as a string of '0' / '1'.  Returns 0; 1 = RegexError (the reference could not build it either);
// 8 = syntax outside the implemented subset.  `stats` (optional, 2 words) receives the DFA's states and classes.
int vidx_regex_match(const char* pattern, int case_insensitive, int starts_with, const char* terms_json, char* out, size_t outlen, uint32_t* stats) {
    try {
        const vregex::Dfa dfa = vregex::compile(pattern, case_insensitive != 0);
        if (stats) stats[0] = dfa.n_states, stats[1] = dfa.n_classes;
        const vjson::Value terms = vjson::parse(terms_json, strlen(terms_json));
        std::string bits;
        std::vector<uint32_t> scalars;
        for (auto& t : terms.arr) {
            scalars.clear();
            vfmt::utf8_decode(t.str, scalars);
            bits += dfa.matches(scalars, starts_with != 0) ? '1' : '0';
        }
        set_err(out, outlen, bits.c_str());
        return 0;
    } catch (const vregex::RegexError& e) {
        set_err(out, outlen, e.what());
        return 1;
    } catch (const vregex::RegexUnsupported& e) {
        set_err(out, outlen, e.what());
        return 8;
    } catch (const std::exception& e) {
        set_err(out, outlen, e.what());
        return 9;
    }
}

// DocLoader::get_doc over the `data` file of `dir` (host/doc_store.hpp, the product's reader): the document text.
int vidx_get_doc(const char* dir, uint32_t doc_id, char* out, size_t outlen) {
    try {
        const std::vector<uint8_t> bytes = vhost::read_file(std::string(dir) + "/data");
        const vhost::DocLoader loader(bytes.data(), bytes.size());
        const std::string doc = loader.get_doc(doc_id);
        if (doc.size() + 1 > outlen) throw std::runtime_error("document larger than the output buffer");
        set_err(out, outlen, doc.c_str());
        return 0;
    } catch (const std::exception& e) {
        set_err(out, outlen, e.what());
        return 1;
    }
}

// DocStoreWriter over the given documents (a JSON list of strings) into `path`; lz4_block_decompress over a hex block.
int vidx_write_doc_store(const char* path, const char* docs_json, char* err, size_t errlen) {
    try {
        const vjson::Value docs = vjson::parse(docs_json, strlen(docs_json));
        vhost::DocStoreWriter w;
        std::vector<uint8_t> bytes;
        for (auto& d : docs.arr) w.add_doc(d.str, bytes);
        w.finish(bytes);
        vhost::write_file(path, bytes.data(), bytes.size());
        return 0;
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return 1;
    }
}
int vidx_lz4_decompress(const uint8_t* block, size_t n, uint8_t* out, size_t out_len, char* err, size_t errlen) {
    try {
        vfmt::lz4_block_decompress(block, n, out, out_len);
        return 0;
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return 1;
    }
}
int vidx_lz4_compress(const uint8_t* src, size_t n, uint8_t* out, size_t cap, size_t* out_len) {
    std::vector<uint8_t> bytes;
    vfmt::lz4_compress_prepend_size(src, n, bytes);
    if (bytes.size() > cap) return 1;
    memcpy(out, bytes.data(), bytes.size());
    *out_len = bytes.size();
    return 0;
}

// highlight_on_original_document (host/highlight.hpp) with the field options of `dir`'s metaData.json: `terms_json` is
// {"<field>.textindex": ["term", ...]}; writes {"<field>": ["<b>..</b> ..", ...]}.
int vidx_highlight_doc(const char* dir, const char* doc_json, const char* terms_json, char* out, size_t outlen) {
    try {
        const std::vector<uint8_t> mj = vhost::read_file(std::string(dir) + "/metaData.json");
        const vhost::Metadata meta = vhost::metadata_from_json(vjson::parse((const char*)mj.data(), mj.size()));
        const vjson::Value terms = vjson::parse(terms_json, strlen(terms_json));
        vhost::TermSets sets;
        for (auto& kv : terms.obj)
            for (auto& t : kv.second.arr) sets[kv.first].insert(t.str);
        std::string s;
        vhost::write_highlights(s, vhost::highlight_document(meta, doc_json, sets));
        set_err(out, outlen, s.c_str());
        return 0;
    } catch (const std::exception& e) {
        set_err(out, outlen, e.what());
        return 1;
    }
}

// vfmt::to_lowercase (format/unicode.hpp: the product's restatement of Rust's str::to_lowercase)
int vidx_to_lowercase(const char* text, char* out, size_t outlen) {
    const std::string low = vfmt::to_lowercase(text);
    if (low.size() + 1 > outlen) return 1;
    memcpy(out, low.c_str(), low.size() + 1);
    return 0;
}

// vfmt::FstReader (format/fst.hpp, the product's dictionary reader) over the file `path`: every "hex(key)<TAB>value" line in key
// order, then get() and ord_to_term() of every key as a self check ("ok" / the first disagreement) in the last line.
int vidx_fst_dump(const char* path, char* out, size_t outlen) {
    try {
        const std::vector<uint8_t> bytes = vhost::read_file(path);
        const vfmt::FstReader r(bytes.data(), bytes.size());
        std::string s;
        std::vector<std::pair<std::string, uint64_t>> items;
        r.for_each([&](const std::string& k, uint64_t v) { items.emplace_back(k, v); });
        std::string check = "ok";
        for (auto& kv : items) {
            static const char* hex = "0123456789abcdef";
            for (unsigned char c : kv.first) s += hex[c >> 4], s += hex[c & 15];
            s += "\t" + std::to_string(kv.second) + "\n";
            uint64_t v = 0;
            if (!r.get(kv.first, v) || v != kv.second) check = "get(" + kv.first + ") disagrees";
        }
        if (items.size() != r.len()) check = "key count differs from the footer";
        s += check;
        if (s.size() + 1 > outlen) throw std::runtime_error("output buffer too small");
        set_err(out, outlen, s.c_str());
        return 0;
    } catch (const std::exception& e) {
        set_err(out, outlen, e.what());
        return 1;
    }
}

// read_data (host/read_document.hpp) over the index in `dir`: `fields_json` = the request's `select`; with `term_ids_json`
// ({"<field>.textindex": [term ids]}) instead the token-id based why_found of anchor `id` (why_found_by_ids).
int vidx_read_doc(const char* dir, uint32_t id, const char* fields_json, const char* term_ids_json, char* out, size_t outlen) {
    try {
        const auto p = vhost::Persistence::load(dir);
        std::string s;
        if (term_ids_json) {
            const vjson::Value t = vjson::parse(term_ids_json, strlen(term_ids_json));
            std::map<std::string, std::set<uint32_t>> ids;
            for (auto& kv : t.obj)
                for (auto& x : kv.second.arr) ids[kv.first].insert((uint32_t)x.num);
            vhost::write_highlights(s, vhost::why_found_by_ids(*p, id, ids));
        } else {
            const vjson::Value f = vjson::parse(fields_json, strlen(fields_json));
            std::vector<std::string> fields;
            for (auto& e : f.arr) fields.push_back(e.str);
            s = vjson::to_string(vhost::read_data(*p, id, fields));
        }
        if (s.size() + 1 > outlen) throw std::runtime_error("output buffer too small");
        set_err(out, outlen, s.c_str());
        return 0;
    } catch (const std::exception& e) {
        set_err(out, outlen, e.what());
        return 1;
    }
}

}  // extern "C"
