"""The request surface has two parsers: `parse_request` over the JSON DOM (the oracle's way in) and the one-pass
`RequestReader` (the planner's way in).  They must agree on every request: same parsed fields, same error class and
message.  Held against each other on hand-written requests of every field of `search::Request`
(src/search/request/mod.rs:14-87), on type errors, on malformed JSON, and on randomly mutated requests."""
import ctypes
import json
import random

import helpers

P = lambda t, **kw: {"search": {"terms": [t], "path": "body", **kw}}
BOOST = {"path": "commonness", "boost_fun": "Log10", "param": 1}

REQUESTS = [
    {"search_req": P("abc")},
    {"search_req": P("Abc", levenshtein_distance=1, starts_with=True, ignore_case=False, boost=2.5, top=3, skip=1)},
    {"search_req": P("abc", is_regex=True, token_value={"path": "tv", "boost_fun": "Multiply", "param": 2, "skip_when_score": [1, 2.5], "expression": "$SCORE * 2"})},
    {"search_req": P("abc", options={"explain": True, "top": 4, "skip": 2, "boost": [BOOST]})},
    {"search_req": {"or": {"queries": [P("a"), P("b", levenshtein_distance=1), {"and": {"queries": [P("c"), P("d")], "options": {"boost": [BOOST]}}}]}}, "boost": [BOOST], "top": 10},
    {"search_req": {"and": {"queries": []}}},
    {"search_req": P("a"), "top": None, "skip": 5},
    {"search_req": P("a"), "top": 0},
    {"search_req": P("a"), "facets": [{"field": "tags[]"}, {"field": "x", "top": None}, {"field": "y", "top": 3}]},
    {"search_req": P("a"), "phrase_boosts": [{"search1": P("a")["search"], "search2": P("b")["search"]}]},
    {"search_req": P("a"), "boost_term": [P("b", boost=3.0)["search"]], "select": ["a", "b[].c"], "why_found": True, "text_locality": True, "explain": True},
    {"search_req": P("a"), "filter": {"or": {"queries": [P("x"), P("y")]}}},
    {"search_req": P("a"), "unknown_key": {"nested": [1, 2, {"x": None}]}, "boost": None, "facets": None, "filter": None},
    {"search_req": P("äöü 日本 \"quoted\" \\ back\n\ttab \U0001F600")},
    {"search_req": {"search": {"terms": ["a", "b"], "path": "body", "levenshtein_distance": None, "boost": None, "top": None}}},
    {},
    {"search_req": None},
]

BROKEN = [
    '{"search_req": {"search": {"terms": ["a"]}}}',                       # missing path
    '{"search_req": {"search": {"path": "body"}}}',                       # missing terms
    '{"search_req": {"search": {"terms": "a", "path": "body"}}}',         # terms not an array
    '{"search_req": {"search": {"terms": [1], "path": "body"}}}',
    '{"search_req": {"search": {"terms": ["a"], "path": 5}}}',
    '{"search_req": {"search": {"terms": ["a"], "path": "body", "levenshtein_distance": -1}}}',
    '{"search_req": {"search": {"terms": ["a"], "path": "body", "levenshtein_distance": 1.5}}}',
    '{"search_req": {"search": {"terms": ["a"], "path": "body", "starts_with": 1}}}',
    '{"search_req": {"search": {"terms": ["a"], "path": "body", "boost": "x"}}}',
    '{"search_req": {"xor": {"queries": []}}}',
    '{"search_req": {"or": {"queries": []}, "and": {"queries": []}}}',
    '{"search_req": {}}',
    '{"search_req": []}',
    '{"search_req": {"or": {}}}',
    '{"search_req": {"or": {"queries": 5}}}',
    '{"search_req": {"or": []}}',
    '{"search_req": {"search": {"terms": ["a"], "path": "body"}}, "top": "ten"}',
    '{"search_req": {"search": {"terms": ["a"], "path": "body"}}, "boost": [{"boost_fun": "Log10"}]}',
    '{"search_req": {"search": {"terms": ["a"], "path": "body"}}, "boost": [{"path": "c", "boost_fun": "Log3"}]}',
    '{"search_req": {"search": {"terms": ["a"], "path": "body"}}, "boost": {"path": "c"}}',
    '{"search_req": {"search": {"terms": ["a"], "path": "body"}}, "facets": [{"top": 3}]}',
    '{"search_req": {"search": {"terms": ["a"], "path": "body"}}, "facets": [5]}',
    '{"search_req": {"search": {"terms": ["a"], "path": "body"}}, "phrase_boosts": [{"search1": {"terms": ["a"], "path": "b"}}]}',
    '{"search_req": {"search": {"terms": ["a"], "path": "body"}}, "why_found": "yes"}',
    '[1, 2]',
    '"text"',
    '',
    '{',
    '{"search_req": {"search": {"terms": ["a"], "path": "body"}}',
    '{"search_req": {"search": {"terms": ["a"], "path": "body"}}} trailing',
    '{"search_req": {"search": {"terms": ["a\\x"], "path": "body"}}}',
    '{"search_req": {"search": {"terms": ["a"], "path": "body",}}}',
    '{"top": 1e3, "search_req": {"search": {"terms": ["a"], "path": "body"}}}',
    '{"top": 18446744073709551615, "skip": 00012, "search_req": {"search": {"terms": ["a"], "path": "body"}}}',
]


def _lib():
    lib = helpers._index_lib()
    lib.vidx_describe_request.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
    return lib


def _both(lib, text):
    out = []
    for reader in (0, 1):
        buf = ctypes.create_string_buffer(1 << 16)
        rc = lib.vidx_describe_request(text.encode("utf-8"), reader, buf, len(buf))
        out.append((rc, buf.value.decode("utf-8", "replace")))
    return out


def test_reader_agrees_with_dom_parser_on_requests():
    lib = _lib()
    for r in REQUESTS:
        for text in (json.dumps(r), json.dumps(r, ensure_ascii=False, indent=2), json.dumps(r, separators=(",", ":"))):
            dom, reader = _both(lib, text)
            assert dom[0] == 0, (text, dom)
            assert dom == reader, text


def test_reader_agrees_with_dom_parser_on_errors():
    lib = _lib()
    for text in BROKEN:
        dom, reader = _both(lib, text)
        assert dom[0] == reader[0], (text, dom, reader)
        if dom[0] != 0 and not dom[1].startswith("json:"):
            assert dom[1] == reader[1], text  # type and shape errors carry the same message
    # the listed cases are failures, except the numeric spellings at the end
    assert sum(1 for t in BROKEN if _both(lib, t)[0][0] != 0) >= len(BROKEN) - 2


def test_reader_agrees_with_dom_parser_on_mutations():
    """Random edits of valid request texts: both parsers accept or both refuse, and when they accept they see the same
    request."""
    lib = _lib()
    rng = random.Random(5)
    texts = [json.dumps(r) for r in REQUESTS if r.get("search_req")]
    alphabet = '{}[]",:0123456789.-eE nulltruefalse\\'
    for _ in range(3000):
        t = list(rng.choice(texts))
        for _ in range(rng.randint(1, 3)):
            i = rng.randrange(len(t))
            op = rng.random()
            if op < 0.4:
                del t[i]
            elif op < 0.8:
                t[i] = rng.choice(alphabet)
            else:
                t.insert(i, rng.choice(alphabet))
        text = "".join(t)
        if "\x00" in text:
            continue
        dom, reader = _both(lib, text)
        assert (dom[0] == 0) == (reader[0] == 0), (text, dom, reader)
        if dom[0] == 0:
            assert dom[1] == reader[1], text
