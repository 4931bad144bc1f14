"""Two oracles against each other on the hot path: oracle/veloci_oracle.cpp (C++, reads index directories through the
product's reader and parses requests with the product's DOM parser) and oracle/search_py.py (plain Python + numpy float32
over the oracle's own decoder of the index files, its own reading of the request JSON, its own matching and scoring).
They share no code; agreement on hit counts, ids and scores means the shared headers under the C++ oracle are not hiding a
common mistake on this path."""
import json
import os
import sys
import tempfile

import pytest

import helpers
import ref_fixtures as fx

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import search_py  # noqa: E402  (test infrastructure)

S = lambda term, path, **kw: {"search": {"terms": [term], "path": path, **kw}}
GER, ENG = "meanings.ger[]", "meanings.eng[]"


def agree(cpp, py, request):
    a, b = cpp.search(request), py.search(request)
    assert a["num_hits"] == b["num_hits"], (request, a["num_hits"], b["num_hits"])
    assert helpers.same_topk([(h[0], h[1]) for h in a["data"]], b["data"]), (request, a["data"], b["data"])
    if request.get("facets"):  # (every group asked for: the order among equal counts is open in the reference)
        for field, groups in a["facets"].items():
            assert sorted((t, n, i) for t, n, i in groups) == sorted(b["facets"][field]), (request, field)
    return a["num_hits"]


def test_reference_corpus(native_libs):
    d = tempfile.mkdtemp(prefix="vb200_py_")
    helpers.create_index(d, fx.TEST_ALL_DOCS, fx.TEST_ALL_CONFIG)
    cpp, py = helpers.Oracle(d), search_py.PySearch(d)
    c1 = {"path": "commonness", "boost_fun": "Log10", "param": 1}
    requests = [
        {"search_req": S("urge", ENG)},
        {"search_req": S("Urge", ENG, ignore_case=False)},
        {"search_req": S("majestätischer", GER, levenshtein_distance=1)},
        {"search_req": S("majestät", GER, levenshtein_distance=2), "top": 3, "skip": 1},
        {"search_req": S("will", GER, starts_with=True), "top": None},
        {"search_req": S("wil", GER, starts_with=True, levenshtein_distance=1, boost=2.5)},
        {"search_req": S("ewsome", "field1[].text", levenshtein_distance=1, ignore_case=True)},   # transposition only with ignore_case (search_field.rs:87)
        {"search_req": S("awesoem", "field1[].text", levenshtein_distance=1, ignore_case=True)},
        {"search_req": S("awesoem", "field1[].text", levenshtein_distance=1)},
        {"search_req": {"or": {"queries": [S("majestät", GER), S("urge", ENG)]}}},
        {"search_req": {"or": {"queries": [S("majestät", GER, levenshtein_distance=1), S("majestät", ENG, levenshtein_distance=1), S("urge", ENG), S("weich", GER, levenshtein_distance=1)]}}, "boost": [c1]},
        {"search_req": {"and": {"queries": [S("majestät", GER, levenshtein_distance=2), S("majestätisches", GER, levenshtein_distance=3)]}}},
        {"search_req": {"or": {"queries": [{"and": {"queries": [S("will", GER, starts_with=True), S("will", ENG, starts_with=True)]}}, S("urge", ENG)]}}, "boost": [c1]},
        {"search_req": S("weich", GER, levenshtein_distance=1), "boost": [{"path": "commonness", "boost_fun": "Log2", "param": 2}]},
        {"search_req": S("weich", GER, levenshtein_distance=1), "boost": [{"path": "commonness", "boost_fun": "Multiply", "expression": "$SCORE * 2", "skip_when_score": [7.5]}, {"path": "commonness", "boost_fun": "Add", "param": 3}]},
        {"search_req": S("weich", GER, levenshtein_distance=1), "boost": [{"path": "commonness", "boost_fun": "Replace"}, {"path": "commonness", "expression": "10 / $SCORE"}]},
        {"search_req": S("nothing matches this", GER)},
        # boost_term (tests.rs:1232-1256), per-part top / skip and part boosts (tests.rs:1276-1288)
        {"search_req": S("majestät", GER, levenshtein_distance=2), "boost_term": [{"terms": ["9555"], "path": "ent_seq", "boost": 5.0}, {"terms": ["majestät"], "path": GER, "levenshtein_distance": 1}]},
        {"search_req": S("will", GER, starts_with=True, top=2, skip=1, boost=0.5)},
        {"search_req": {"or": {"queries": [S("majestät", GER, levenshtein_distance=2, top=1), S("will", GER, starts_with=True, top=3)]}}},
        # filters (tests.rs:753-824) and facets (tests_facet.rs:60-117 run on their own corpus below)
        {"search_req": S("majestät", GER, levenshtein_distance=1), "filter": S("20", "commonness")},
        {"search_req": {"or": {"queries": [S("majestät", GER, levenshtein_distance=2), S("urge", ENG), S("will", GER, starts_with=True)]}},
         "filter": {"or": {"queries": [S("20", "commonness"), S("1587690", "ent_seq")]}}, "boost": [c1]},
        {"search_req": S("will", GER, starts_with=True), "filter": {"and": {"queries": [S("will", ENG, starts_with=True), S("will", GER, starts_with=True)]}}},
        {"search_req": {"or": {"queries": [S("majestät", GER, levenshtein_distance=2), S("will", GER, starts_with=True)]}}, "facets": [{"field": "tags[]", "top": 100}, {"field": "commonness", "top": 100}]},
    ]
    total = sum(agree(cpp, py, r) for r in requests)
    assert total > 25
    with pytest.raises(KeyError):
        py.search({"search_req": S("a", "notexisting")})


def test_synthetic_config2_shape(native_libs):
    """A seeded Zipfian corpus in the shape of BASELINE config 2 (3-term `or`, levenshtein 1, Log10 boost) and its `and` /
    single-term relatives, small enough for the Python oracle (500-term dictionary)."""
    params = dict(num_docs=4000, vocab=500, seed=7)
    d = tempfile.mkdtemp(prefix="vb200_py_")
    helpers.create_synthetic_index(d, **params)
    cpp, py = helpers.Oracle(d), search_py.PySearch(d)
    reqs = helpers.synthetic_requests(num_queries=40, query_kind="or3", levenshtein=1, query_seed=3, **params)
    reqs += helpers.synthetic_requests(num_queries=15, query_kind="single", levenshtein=2, query_seed=4, **params)
    params["tags"] = 30  # the `and` requests count facets on tags[]
    helpers.create_synthetic_index(d, **params)
    cpp, py = helpers.Oracle(d), search_py.PySearch(d)
    for r in helpers.synthetic_requests(num_queries=15, query_kind="and", levenshtein=1, query_seed=5, **params):
        r = json.loads(r)
        r["facets"][0]["top"] = 1000
        reqs.append(json.dumps(r))
    total = sum(agree(cpp, py, json.loads(r)) for r in reqs)
    assert total > 20000


def test_phrase_and_text_locality_corpora(native_libs):
    """The reference's phrase tests (test_phrase.rs) and its text locality test (tests.rs:1296-1313) through both oracles."""
    d = tempfile.mkdtemp(prefix="vb200_py_")
    helpers.create_index(d, fx.TEST_PHRASE_DOCS, fx.TEST_PHRASE_CONFIG)
    cpp, py = helpers.Oracle(d), search_py.PySearch(d)
    T = lambda *terms: [S(t, "tags[]") for t in terms]
    pb = lambda path, a, b: {"path": path, "search1": {"terms": [a], "path": path}, "search2": {"terms": [b], "path": path}}
    requests = [
        {"search_req": S("erbin", "title"), "phrase_boosts": [pb("title", "die", "erbin")]},
        {"search_req": {"or": {"queries": [S(t, p) for p in ("title", "tags[]") for t in ("die", "erbin")]}}, "phrase_boosts": [pb("title", "die", "erbin"), pb("tags[]", "die", "erbin")]},
        {"search_req": {"or": {"queries": T("greg", "tagebuch", "05")}}, "phrase_boosts": [pb("tags[]", "greg", "tagebuch"), pb("tags[]", "tagebuch", "05")]},
        {"search_req": {"and": {"queries": T("greg", "tagebuch", "05")}}, "phrase_boosts": [pb("tags[]", "greg", "tagebuch")]},
        {"search_req": {"or": {"queries": T("greg", "tagebuch", "05") + [S(t, "title") for t in ("greg", "tagebuch", "05")]}}, "text_locality": True,
         "phrase_boosts": [pb("tags[]", "greg", "tagebuch"), pb("title", "greg", "tagebuch"), pb("tags[]", "tagebuch", "05"), pb("title", "tagebuch", "05")]},
        {"search_req": {"or": {"queries": T("greg", "tagebuch", "05")}}, "text_locality": True},
    ]
    assert sum(agree(cpp, py, r) for r in requests) > 10
    d = tempfile.mkdtemp(prefix="vb200_py_")
    helpers.create_index(d, fx.TEST_ALL_DOCS, fx.TEST_ALL_CONFIG)
    cpp, py = helpers.Oracle(d), search_py.PySearch(d)
    req = {"search_req": {"or": {"queries": [S("text", GER), S("localität", GER), S("text", ENG)]}}, "text_locality": True}
    assert agree(cpp, py, req) >= 2


def test_synthetic_config3_shape(native_libs):
    """BASELINE config 3 in small: `and` of 2-3 parts (levenshtein 0 / 1), phrase boosts for adjacent parts, text locality,
    facets on tags[] -- every piece of that request through both oracles."""
    params = dict(num_docs=3000, vocab=400, seed=9, tags=25, text_locality=True, phrase=True)
    d = tempfile.mkdtemp(prefix="vb200_py_")
    helpers.create_synthetic_index(d, **params)
    cpp, py = helpers.Oracle(d), search_py.PySearch(d)
    total = boosted = 0
    for r in helpers.synthetic_requests(num_queries=40, query_kind="and", levenshtein=1, query_seed=6, **params):
        r = json.loads(r)
        assert r.get("phrase_boosts") and r.get("text_locality") and r.get("facets")
        r["facets"][0]["top"] = 1000
        total += agree(cpp, py, r)
        plain = {k: v for k, v in r.items() if k not in ("phrase_boosts", "text_locality")}
        boosted += [h[1] for h in cpp.search(r)["data"]] != [h[1] for h in cpp.search(plain)["data"]]
    assert total > 300 and boosted >= 5, (total, boosted)


def test_per_part_top_and_token_values(native_libs):
    """The per-part bound (keep top + skip + 200, cut, drop below the worst kept: search_field.rs:292-294,322-331,366-369) on
    parts that match hundreds of terms, and token values (:391-395), through both oracles as term hit lists and as whole requests."""
    from test_part_hits import make_valued_index, token_value_parts
    d, cpp, words, _ = make_valued_index()
    py = search_py.PySearch(d)
    parts = list(token_value_parts(words[:8]))
    # one-letter prefixes match hundreds of terms: the bound's cut at top + skip + 200 hits is reached
    parts += [{"terms": [w[:1]], "path": "body", "starts_with": True, "levenshtein_distance": i % 2, "top": 2 + 3 * i, "skip": i % 3,
               **({"boost": -1.5} if i == 2 else {}), **({"token_value": {"path": "body", "boost_fun": "Multiply"}} if i % 2 else {})} for i, w in enumerate(words[8:13])]
    n_long = 0
    for part in parts:
        a = sorted(cpp.call("field_search", part=part)["hits_scores"])
        b = sorted((t, float(s)) for t, s in py.field_search(part)[1])
        assert [t for t, _ in a] == [t for t, _ in b], part
        assert all(abs(x[1] - y[1]) <= 1e-5 * max(abs(x[1]), 1e-30) for x, y in zip(a, b)), part
        agree(cpp, py, {"search_req": {"search": part}, "top": 20})
        n_long += len(cpp.call("field_search", part={k: v for k, v in part.items() if k not in ("top", "skip", "token_value")})["hits_scores"]) > 210
    assert n_long >= 3, "no part reached the bound"


def test_one_to_n_boosts(native_libs):
    """Boosts on a 1:n level (tests.rs:839-931): BoostToAnchor's joins and ApplyAnchorBoost's merge walk, with its
    position dependence when a document has several boosted values, through both oracles."""
    import numpy as np
    d = tempfile.mkdtemp(prefix="vb200_py_")
    helpers.create_index(d, fx.TEST_ALL_DOCS, fx.TEST_ALL_CONFIG)
    cpp, py = helpers.Oracle(d), search_py.PySearch(d)
    c1 = {"path": "commonness", "boost_fun": "Log10", "param": 1}
    requests = [
        {"search_req": S("意慾", "kanji[].text"), "boost": [{"path": "kanji[].commonness", "boost_fun": "Log10", "param": 1}]},
        {"search_req": S("awesome", "field1[].text"), "boost": [c1, {"path": "field1[].rank", "expression": "10 / $SCORE", "skip_when_score": [0]}]},
        {"search_req": {"or": {"queries": [
            {"search": {"terms": ["awesome"], "path": "field1[].text", "options": {"boost": [{"path": "field1[].rank", "boost_fun": "Log10", "param": 1}]}}},
            {"search": {"terms": ["urge"], "path": ENG, "options": {"boost": [c1]}}}]}}},
        {"search_req": {"or": {"queries": [S("awesome", "field1[].text"), S("意慾", "kanji[].text", levenshtein_distance=1)]}},
         "boost": [{"path": "field1[].rank", "boost_fun": "Multiply"}, {"path": "kanji[].commonness", "boost_fun": "Add", "param": 2}, {"path": "commonness", "boost_fun": "Log2", "param": 2}]},
    ]
    assert sum(agree(cpp, py, r) for r in requests) >= 4
    rng = np.random.default_rng(11)
    syll = ["ka", "ki", "ku", "mi", "mo", "ra", "ri", "ru", "sa", "to"]
    words = ["".join(rng.choice(syll, size=int(rng.integers(2, 4)))) for _ in range(120)]
    docs = []
    for i in range(1500):
        n = int(rng.integers(0, 4))
        doc = {"ent_seq": str(i)}
        if n:
            doc["kana"] = [{"text": str(rng.choice(words)), **({"commonness": int(rng.integers(1, 900))} if rng.random() < 0.75 else {})} for _ in range(n)]
        docs.append(doc)
    d = tempfile.mkdtemp(prefix="vb200_py_")
    helpers.create_index(d, docs, {"kana[].text": {"fulltext": {"tokenize": True}}, "kana[].commonness": dict(fx.BOOST)})
    cpp, py = helpers.Oracle(d), search_py.PySearch(d)
    total = 0
    for t in [str(w) for w in rng.choice(words, size=8)] + ["mi", "ka"]:
        for fun in ("Log10", "Multiply", "Add"):
            total += agree(cpp, py, {"search_req": S(t, "kana[].text", levenshtein_distance=1, starts_with=True), "boost": [{"path": "kana[].commonness", "boost_fun": fun, "param": 1}], "top": 50})
    assert total > 1000


def test_random_request_trees(native_libs):
    """300 seeded random requests on the reference corpus -- nested or / and trees of fuzzy, prefix and exact parts over
    several fields, optional filter tree, anchor and 1:n boosts, boost_term, text locality, facets, top / skip -- through both oracles."""
    import random
    rng = random.Random(20261019)
    d = tempfile.mkdtemp(prefix="vb200_py_")
    helpers.create_index(d, fx.TEST_ALL_DOCS, fx.TEST_ALL_CONFIG)
    cpp, py = helpers.Oracle(d), search_py.PySearch(d)
    terms = {GER: ["majestät", "majestätischer", "will", "wille", "weich", "welch", "text", "localität", "begeisterung", "der", "test", "treffer"],
             ENG: ["urge", "will", "majesty", "test1", "text", "awesome"], "field1[].text": ["awesome", "nice"], "kanji[].text": ["意慾", "偉容"],
             "ent_seq": ["1587690", "1587700", "9555"], "tags[]": ["nice", "cool", "awesome"]}

    def part():
        path = rng.choice(list(terms))
        t = rng.choice(terms[path])
        p = {"terms": [t[:rng.randrange(2, len(t) + 1)] if rng.random() < 0.3 else t], "path": path}
        if rng.random() < 0.5:
            p["levenshtein_distance"] = rng.randrange(0, 3)
        if rng.random() < 0.3:
            p["starts_with"] = True
        if rng.random() < 0.15:
            p["ignore_case"] = rng.random() < 0.5
        if rng.random() < 0.15:
            p["boost"] = rng.choice([0.5, 2.0, 3.5])
        if rng.random() < 0.1:
            p["top"], p["skip"] = rng.randrange(1, 4), rng.randrange(0, 2)
        return {"search": p}

    def tree(depth):
        if depth == 0 or rng.random() < 0.35:
            return part()
        return {rng.choice(["or", "and"]): {"queries": [tree(depth - 1) for _ in range(rng.randrange(1, 4))]}}

    total = 0
    for _ in range(300):
        r = {"search_req": tree(2)}
        if rng.random() < 0.2:
            r["filter"] = tree(1)
        boosts = []
        if rng.random() < 0.4:
            boosts.append({"path": "commonness", "boost_fun": rng.choice(["Log10", "Log2", "Multiply", "Add", "Replace"]), "param": rng.choice([1, 2, 0.5])})
        if rng.random() < 0.15:
            boosts.append({"path": rng.choice(["kanji[].commonness", "field1[].rank"]), "boost_fun": rng.choice(["Log10", "Multiply", "Add"]), "param": 1})
        if rng.random() < 0.1:
            boosts.append({"path": "commonness", "expression": rng.choice(["$SCORE * 2", "10 / $SCORE", "$SCORE + 1.5"]), "skip_when_score": [10.0]})
        if boosts:
            r["boost"] = boosts
        if rng.random() < 0.15:
            r["boost_term"] = [part()["search"]]
        if rng.random() < 0.2:
            r["text_locality"] = True
        if rng.random() < 0.2:
            r["facets"] = [{"field": rng.choice(["tags[]", "commonness"]), "top": 100}]
        if rng.random() < 0.3:
            r["top"], r["skip"] = rng.randrange(1, 6), rng.randrange(0, 3)
        total += agree(cpp, py, r)
    assert total > 300


def test_suggest(native_libs):  # tests.rs:1087-1158 and the shapes of tests/test_gpu_features.py::test_suggest_*
    d = tempfile.mkdtemp(prefix="vb200_py_")
    helpers.create_index(d, fx.TEST_ALL_DOCS, fx.TEST_ALL_CONFIG)
    helpers.add_token_values(d, *fx.TEST_ALL_TOKEN_VALUES)
    cpp, py = helpers.Oracle(d), search_py.PySearch(d)

    def same(a, b, ctx):
        assert [(t, i) for t, _, i in a] == [(t, i) for t, _, i in b], (ctx, a, b)
        assert all(abs(x[1] - y[1]) <= 1e-5 * max(abs(x[1]), 1e-30) for x, y in zip(a, b)), ctx

    part = {"terms": ["majes"], "path": GER, "levenshtein_distance": 0, "starts_with": True, "top": 10, "skip": 0}
    got = py.suggest(part)
    assert sorted(t for t, _, _ in got) == sorted(["majestät", "majestät (f)", "majestätisches", "majestätischer", "majestätischer anblick (m)", "majestätisches aussehen (n)"])
    same(cpp.call("suggest", part=part), got, part)
    tv = {"terms": ["begeist"], "path": GER, "levenshtein_distance": 0, "starts_with": True, "token_value": {"path": GER, "boost_fun": "Log10", "param": 1}, "top": 10, "skip": 0}
    assert [t for t, _, _ in py.suggest(tv)] == ["begeisterung", "begeistern", "begeisterung (f)"]  # tests.rs:1134-1158
    for req in ({"suggest": [{"terms": ["will"], "path": GER, "levenshtein_distance": 0, "starts_with": True}, {"terms": ["will"], "path": ENG, "levenshtein_distance": 0, "starts_with": True}], "top": 10, "skip": 0},
                {"suggest": [{"terms": ["majes"], "path": GER, "starts_with": True}], "top": 2, "skip": 1},
                {"suggest": [{"terms": ["Majestat"], "path": GER, "levenshtein_distance": 2}], "top": None},
                {"suggest": [{"terms": ["will"], "path": GER, "starts_with": True, "boost": 3.0, "top": 2}, {"terms": ["wille"], "path": GER, "levenshtein_distance": 1}]},
                {"suggest": [], "top": 10}):
        same(cpp.call("suggest_multi", request=req), py.suggest_multi(req), req)
