"""ctypes bindings used by the tests: the index builder and the CPU oracle.

The oracle (oracle/veloci_oracle.cpp) is test infrastructure: it is only ever
loaded from tests/, __graft_entry__.smoke() and bench.py's CPU baseline.
"""
import ctypes
import json
import os
import tempfile

from veloci_b200 import build

_ERRLEN = 4096


class OracleError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(message)
        self.status = status
        self.message = message


def _index_lib():
    lib = ctypes.CDLL(build.build_index_lib())
    lib.vidx_create_from_jsonl.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    lib.vidx_add_token_values.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    lib.vidx_bound_part_hits.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    lib.vidx_explain_walk.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    lib.vidx_field_highlight.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
    lib.vidx_explain_plan.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    lib.vidx_create_synthetic.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    lib.vidx_write_synthetic_requests.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    return lib


def create_index(directory, docs, config=None):
    """docs: list of JSON-able documents (or a JSON-lines string); config: dict of field configs."""
    lib = _index_lib()
    text = docs if isinstance(docs, str) else "\n".join(json.dumps(d, ensure_ascii=False) for d in docs)
    err = ctypes.create_string_buffer(_ERRLEN)
    rc = lib.vidx_create_from_jsonl(directory.encode(), text.encode("utf-8"), json.dumps(config or {}).encode(), err, _ERRLEN)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return directory


def add_token_values(directory, data, config):
    """create/token_values_to_tokens.rs:26-82 on an index directory: data = [{"text", "value"}], config = {"path": field}."""
    lib = _index_lib()
    err = ctypes.create_string_buffer(_ERRLEN)
    rc = lib.vidx_add_token_values(directory.encode(), json.dumps(data, ensure_ascii=False).encode("utf-8"), json.dumps(config).encode(), err, _ERRLEN)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return directory


def bound_part_hits(directory, part, hits):
    """The product's host arithmetic on a part's term hits (csrc/host/part_hits.hpp; no device): per-part top / skip bound,
    part boost, token_value boost.  `hits` = [(term id, score)] of the part without those; returns its final hits."""
    lib = _index_lib()
    out = ctypes.create_string_buffer(1 << 22)
    rc = lib.vidx_bound_part_hits(directory.encode(), json.dumps(part, ensure_ascii=False).encode("utf-8"), json.dumps([[int(i), float(s)] for i, s in hits]).encode(), out, len(out))
    if rc != 0:
        raise RuntimeError(out.value.decode())
    return [(int(i), float(s)) for i, s in json.loads(out.value.decode())]


def explain_walk(directory, request, anchors, leaves):
    """The product's explain walk (csrc/host/explain_walk.hpp; no device) for the returned `anchors` of `request`, given per
    search part of the tree (tree order) the (term id, score) hits of the bare part.  -> {anchor: [Explain, ...]}"""
    lib = _index_lib()
    out = ctypes.create_string_buffer(1 << 22)
    rc = lib.vidx_explain_walk(directory.encode(), json.dumps(request, ensure_ascii=False).encode("utf-8"), json.dumps([int(a) for a in anchors]).encode(),
                               json.dumps([[[int(i), float(s)] for i, s in hits] for hits in leaves]).encode(), out, len(out))
    if rc != 0:
        raise OracleError(rc, out.value.decode())
    return {int(k): v for k, v in json.loads(out.value.decode()).items()}


def tree_parts(search_request):
    """The search parts of a request tree in tree order."""
    if "search" in search_request:
        return [search_request["search"]]
    node = search_request.get("or") or search_request.get("and")
    return [p for q in node["queries"] for p in tree_parts(q)]


def bare_part(part):
    return {k: v for k, v in part.items() if k not in ("top", "skip", "boost", "token_value", "options")}


def same_explain(got, want, rel=1e-5):
    """Two explain values ({anchor: [Explain]} or parts of them): same structure, texts and ids, floats within `rel`."""
    if isinstance(want, dict):
        return isinstance(got, dict) and set(map(str, got)) == set(map(str, want)) and all(same_explain(got[k] if k in got else got[str(k)], v, rel) for k, v in want.items())
    if isinstance(want, list):
        return isinstance(got, list) and len(got) == len(want) and all(same_explain(g, w, rel) for g, w in zip(got, want))
    if isinstance(want, float) or isinstance(got, float):
        return abs(float(got) - float(want)) <= rel * max(abs(float(want)), 1e-30)
    return got == want


def field_highlight(directory, part, hits):
    """The host half of the product's search_field::highlight (csrc/host/field_highlight.hpp; no device) over the bare
    part's (term id, score) hits -> [(highlighted text, score, text id)]."""
    lib = _index_lib()
    out = ctypes.create_string_buffer(1 << 22)
    rc = lib.vidx_field_highlight(directory.encode(), json.dumps(part, ensure_ascii=False).encode("utf-8"), json.dumps([[int(i), float(s)] for i, s in hits]).encode(), 0, out, len(out))
    if rc != 0:
        raise OracleError(rc, out.value.decode())
    return [(t, float(s), int(i)) for t, s, i in json.loads(out.value.decode("utf-8"))]


def normalize_text(text):
    """util::normalize_text of the product's host code."""
    lib = _index_lib()
    out = ctypes.create_string_buffer(1 << 16)
    rc = lib.vidx_field_highlight(b"", json.dumps(text, ensure_ascii=False).encode("utf-8"), b"[]", 1, out, len(out))
    if rc != 0:
        raise RuntimeError(out.value.decode())
    return json.loads(out.value.decode("utf-8"))


def explain_plan(request):
    """search::explain_plan of the product's host code (csrc/host/explain_plan.hpp): a Graphviz dot text."""
    lib = _index_lib()
    out = ctypes.create_string_buffer(1 << 20)
    rc = lib.vidx_explain_plan(json.dumps(request, ensure_ascii=False).encode("utf-8"), out, len(out))
    if rc != 0:
        raise RuntimeError(out.value.decode())
    return out.value.decode("utf-8")


def create_synthetic_index(directory, **params):
    lib = _index_lib()
    err = ctypes.create_string_buffer(_ERRLEN)
    rc = lib.vidx_create_synthetic(directory.encode(), json.dumps(params).encode(), err, _ERRLEN)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return directory


def synthetic_requests(**params):
    lib = _index_lib()
    err = ctypes.create_string_buffer(_ERRLEN)
    with tempfile.NamedTemporaryFile(suffix=".jsonl", delete=False) as f:
        path = f.name
    try:
        rc = lib.vidx_write_synthetic_requests(path.encode(), json.dumps(params).encode(), err, _ERRLEN)
        if rc != 0:
            raise RuntimeError(err.value.decode())
        with open(path) as f:
            return [line for line in f.read().split("\n") if line]
    finally:
        os.unlink(path)


def query_parse(text, what=0, no_attributes=False, no_parentheses=False, no_levensthein=False):
    """The product's query parser (csrc/host/query_parser.hpp) through the host-only helper library.
    what: 0 Debug text of the tree, 1 phrase pairs, 2 terms, 3 tokens.  Returns (ok, text-or-JSON)."""
    lib = _index_lib()
    lib.vidx_query_parse.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
    out = ctypes.create_string_buffer(1 << 16)
    rc = lib.vidx_query_parse(text.encode("utf-8"), int(no_attributes) | int(no_parentheses) << 1 | int(no_levensthein) << 2, what, out, len(out))
    s = out.value.decode("utf-8")
    return rc == 0, (json.loads(s) if rc == 0 and what else s)


def query_filter_stopwords(text, stopwords):
    lib = _index_lib()
    lib.vidx_query_filter_stopwords.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    out = ctypes.create_string_buffer(1 << 16)
    rc = lib.vidx_query_filter_stopwords(text.encode("utf-8"), json.dumps(sorted(stopwords), ensure_ascii=False).encode("utf-8"), out, len(out))
    assert rc == 0, out.value
    return out.value.decode("utf-8")


def generate_request(directory, params, suggest=False):
    """search_query / suggest_query of the product's host code over the index in `directory` (no device).
    Returns (status, request dict or message, raw text)."""
    lib = _index_lib()
    lib.vidx_generate_request.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
    out = ctypes.create_string_buffer(1 << 20)
    rc = lib.vidx_generate_request(directory.encode(), json.dumps(params, ensure_ascii=False).encode("utf-8"), int(suggest), out, len(out))
    s = out.value.decode("utf-8")
    return rc, (json.loads(s) if rc == 0 else s), s


def regex_match(pattern, terms, case_insensitive=True, starts_with=False):
    """The product's regex DFA (csrc/host/regex_dfa.hpp) on the host: (status, [bool per term] or message, (states, classes))."""
    lib = _index_lib()
    lib.vidx_regex_match.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32)]
    out = ctypes.create_string_buffer(max(4096, len(terms) + 16))
    stats = (ctypes.c_uint32 * 2)()
    rc = lib.vidx_regex_match(pattern.encode("utf-8"), int(case_insensitive), int(starts_with), json.dumps(terms, ensure_ascii=False).encode("utf-8"), out, len(out), stats)
    s = out.value.decode("utf-8")
    return rc, ([c == "1" for c in s] if rc == 0 else s), (stats[0], stats[1])


class Oracle:
    def __init__(self, directory=None):
        self.lib = ctypes.CDLL(build.build_oracle())
        L = self.lib
        L.vo_open.restype = ctypes.c_void_p
        L.vo_open.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        L.vo_close.argtypes = [ctypes.c_void_p]
        L.vo_free.argtypes = [ctypes.c_void_p]
        L.vo_search.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p, ctypes.c_size_t]
        L.vo_call.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p, ctypes.c_size_t]
        L.vo_search_batch.restype = ctypes.c_double
        L.vo_search_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_char_p), ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        self.h = None
        if directory is not None:
            err = ctypes.create_string_buffer(_ERRLEN)
            self.h = L.vo_open(directory.encode(), err, _ERRLEN)
            if not self.h:
                raise OracleError(4, err.value.decode())

    def close(self):
        if self.h:
            self.lib.vo_close(self.h)
            self.h = None

    def _take(self, rc, out, err):
        if rc != 0:
            raise OracleError(rc, err.value.decode())
        s = ctypes.string_at(out.value).decode("utf-8")
        self.lib.vo_free(out)
        return s

    def search(self, request):
        """request: dict or JSON string -> {"num_hits", "data": [[id, score, score_bits]...], "facets"}"""
        text = request if isinstance(request, str) else json.dumps(request, ensure_ascii=False)
        out = ctypes.c_void_p()
        err = ctypes.create_string_buffer(_ERRLEN)
        rc = self.lib.vo_search(self.h, text.encode("utf-8"), ctypes.byref(out), err, _ERRLEN)
        return json.loads(self._take(rc, out, err))

    def call(self, fn, **args):
        out = ctypes.c_void_p()
        err = ctypes.create_string_buffer(_ERRLEN)
        rc = self.lib.vo_call(self.h, fn.encode(), json.dumps(args, ensure_ascii=False).encode("utf-8"), ctypes.byref(out), err, _ERRLEN)
        return json.loads(self._take(rc, out, err))

    def search_batch(self, requests, threads=1, k=10):
        import numpy as np

        n = len(requests)
        arr = (ctypes.c_char_p * n)(*[r.encode("utf-8") if isinstance(r, str) else json.dumps(r).encode("utf-8") for r in requests])
        ids = np.zeros((n, k), dtype=np.uint32)
        scores = np.zeros((n, k), dtype=np.float32)
        num_hits = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        secs = self.lib.vo_search_batch(self.h, arr, n, threads, k, ids.ctypes.data, scores.ctypes.data, num_hits.ctypes.data, status.ctypes.data)
        return {"seconds": secs, "ids": ids, "scores": scores, "num_hits": num_hits, "status": status}


# ---- result parity (bench.py, __graft_entry__.smoke and the GPU tests share one rule)
REL_TOL = 1e-5  # f32 scores: relative tolerance north_star states


def _close(a, b):
    return abs(float(a) - float(b)) <= REL_TOL * max(abs(float(a)), abs(float(b)), 1e-30)


def same_topk(got, ref):
    """got / ref: lists of (id, score) in rank order.  Equal when the scores agree rank by rank within REL_TOL and every id
    that differs is a tie: an id of `got` the reference also returned must tie with the reference's hit at that rank, an id
    the reference did not return must tie with the reference's last (boundary) hit.  -> (ok, reason)"""
    if len(got) != len(ref):
        return False, f"{len(got)} hits vs {len(ref)}"
    ref_scores = {i: s for i, s in ref}
    for pos, ((gi, gs), (ci, cs)) in enumerate(zip(got, ref)):
        if not _close(gs, cs):
            return False, f"score at rank {pos}: {gs!r} vs {cs!r}"
        if gi != ci:
            if gi in ref_scores:
                if not _close(ref_scores[gi], cs):
                    return False, f"id {gi} at rank {pos} is not a tie with {ci}"
            elif not _close(gs, ref[-1][1]):
                return False, f"id {gi} at rank {pos} is neither in the reference's top-k nor a boundary tie"
    return True, ""


def batch_parity(got, ref, rows=None):
    """got / ref: results_flat-style dicts (ids [n][k] padded with 0xFFFFFFFF, scores, num_hits, status) over the same
    requests.  -> {"checked", "equal", "num_hits_equal", "first_mismatch"}: hit counts exact, top-k per same_topk."""
    n = len(ref["num_hits"]) if rows is None else len(rows)
    equal = hits_equal = 0
    first = None
    for j in range(n):
        q = j if rows is None else rows[j]
        ok = int(got["status"][q]) == int(ref["status"][j])
        why = "status" if not ok else ""
        if ok and int(got["num_hits"][q]) == int(ref["num_hits"][j]):
            hits_equal += 1
        elif ok:
            ok, why = False, f"num_hits {int(got['num_hits'][q])} vs {int(ref['num_hits'][j])}"
        if ok:
            g = [(int(i), float(s)) for i, s in zip(got["ids"][q], got["scores"][q]) if int(i) != 0xFFFFFFFF]
            c = [(int(i), float(s)) for i, s in zip(ref["ids"][j], ref["scores"][j]) if int(i) != 0xFFFFFFFF]
            ok, why = same_topk(g, c)
        if ok:
            equal += 1
        elif first is None:
            first = {"request": int(q), "why": why}
    return {"checked": n, "equal": equal, "num_hits_equal": hits_equal, "first_mismatch": first}
