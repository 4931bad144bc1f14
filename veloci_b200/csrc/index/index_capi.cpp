// C entry points of the index-building helper library (libveloci_b200_index.so).
// Test/bench infrastructure: produces index directories in the reference's
// on-disk layout; not on the accelerated query path.
#include <cstring>
#include <sstream>

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

}  // extern "C"
