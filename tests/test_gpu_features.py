"""GPU parity of the plan steps around the core path: filters,
boost_term, phrase boosts, text locality, facets.  Every request runs through the C ABI on the
device and through the CPU oracle on the same index directory; hit counts, hit ids, facet
groups are compared exactly, scores within 1e-5 relative."""
import json
import tempfile

import numpy as np
import pytest

import helpers
import ref_fixtures as fx
from test_gpu_parity import assert_same_topk, close, compare_batch  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import veloci_b200

    assert veloci_b200.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return veloci_b200


def _pair(gpu, docs=None, config=None, synth=None):
    d = tempfile.mkdtemp(prefix="vb200_gpu_feat_")
    if synth is not None:
        helpers.create_synthetic_index(d, **synth)
    else:
        helpers.create_index(d, docs, config)
    return gpu.Index(d), helpers.Oracle(d)


def compare(index, oracle, requests):
    reqs = [r if isinstance(r, str) else json.dumps(r, ensure_ascii=False) for r in requests]
    b = index.prepare(reqs).execute()
    for q, r in enumerate(reqs):
        assert b.status(q) == 0, (r, b.message(q))
        g = b.result(q)
        c = oracle.search(r)
        assert g["num_hits"] == c["num_hits"], (r, g["num_hits"], c["num_hits"])
        assert_same_topk(g["data"], [(h[0], np.float32(h[1])) for h in c["data"]], ctx=r)
        if "facets" in c:
            assert "facets" in g, r
            assert set(g["facets"]) == set(c["facets"]), r
            for field, groups in c["facets"].items():
                assert [(t, n, i) for t, n, i in g["facets"][field]] == [(t, n, i) for t, n, i in groups], (r, field)
    return b


S = lambda term, path, **kw: {"search": {"terms": [term], "path": path, **kw}}
OR_MAJ_URGE = {"or": {"queries": [S("majestät", "meanings.ger[]"), S("urge", "meanings.eng[]")]}}


@pytest.fixture(scope="module")
def test_all(gpu, native_libs):
    return _pair(gpu, fx.TEST_ALL_DOCS, fx.TEST_ALL_CONFIG)


def test_filters(test_all):  # tests/all/tests.rs:753-824
    index, oracle = test_all
    compare(index, oracle, [
        {"search_req": S("urge", "meanings.eng[]"), "filter": S("1587690", "ent_seq")},
        {"search_req": OR_MAJ_URGE, "filter": S("1587690", "ent_seq")},
        {"search_req": OR_MAJ_URGE, "filter": S("urge", "meanings.eng[]")},
        {"search_req": OR_MAJ_URGE, "filter": {"or": {"queries": [S("1587690", "ent_seq"), S("urge", "meanings.eng[]")]}}},
        {"search_req": OR_MAJ_URGE, "filter": {"and": {"queries": [S("majestät", "meanings.ger[]"), S("urge", "meanings.eng[]")]}}},
        {"search_req": S("will", "meanings.eng[]"), "filter": S("nothing-matches", "ent_seq"), "boost": [{"path": "commonness", "boost_fun": "Log10", "param": 1}]},
        # filter parts are searched for ids: a term that only occurs as a token of longer texts selects nothing (search_field.rs:468-498)
        {"search_req": OR_MAJ_URGE, "filter": S("majestätischer", "meanings.ger[]")},
        {"search_req": OR_MAJ_URGE, "filter": S("majestät", "meanings.ger[]", levenshtein_distance=1)},
        {"search_req": S("will", "meanings.eng[]"), "filter": S("will", "meanings.eng[]")},
    ])


def test_boost_term(test_all):  # tests/all/tests.rs:1232-1256
    index, oracle = test_all
    compare(index, oracle, [
        {"search_req": S("will", "meanings.eng[]"), "boost_term": [{"terms": ["9555"], "path": "ent_seq", "boost": 5.0}]},
        {"search_req": OR_MAJ_URGE, "boost_term": [{"terms": ["urge"], "path": "meanings.eng[]"}, {"terms": ["1587690"], "path": "ent_seq", "boost": 3.5}]},
        {"search_req": OR_MAJ_URGE, "boost_term": [{"terms": ["will"], "path": "meanings.eng[]", "levenshtein_distance": 1}],
         "boost": [{"path": "commonness", "boost_fun": "Log2", "param": 2}], "filter": S("urge", "meanings.eng[]")},
    ])


def test_facets_reference_fixture(gpu, native_libs):  # tests/all/tests_facet.rs:60-117
    index, oracle = _pair(gpu, fx.TEST_FACET_DOCS, fx.TEST_FACET_CONFIG)
    b = compare(index, oracle, [
        {"search_req": S("will", "meanings.eng[]"), "facets": [{"field": "tags[]"}, {"field": "commonness"}]},
        {"search_req": S("test", "meanings.ger[]"), "facets": [{"field": "meanings.eng[]"}]},
        {"search_req": S("will", "meanings.eng[]"), "facets": [{"field": "tags[]", "top": 1}]},
    ])
    assert [f[:2] for f in b.result(0)["facets"]["tags[]"]] == [("nice", 2), ("cool", 1)]
    assert [f[:2] for f in b.result(0)["facets"]["commonness"]] == [("20", 2)]
    assert [f[:2] for f in b.result(1)["facets"]["meanings.eng[]"]] == [("test1", 1)]


def test_facets_large(gpu, native_libs):  # tests/all/tests_large.rs:40-112
    docs = [{"category": "superb", "tags": ["nice", "cool"]}] * 300 + [{"category": "awesome", "tags": ["is", "cool"]}] * 300
    index, oracle = _pair(gpu, docs, {"*GLOBAL*": {"features": ["All"]}, "tags[]": {"facet": True}})
    b = compare(index, oracle, [
        {"search_req": S("superb", "category"), "facets": [{"field": "tags[]"}]},
        {"search_req": {"or": {"queries": [S("superb", "category"), S("awesome", "category")]}}, "facets": [{"field": "tags[]"}]},
        {"search_req": S("superb", "category"), "filter": S("awesome", "category"), "facets": [{"field": "tags[]"}]},
    ])
    assert sorted(f[:2] for f in b.result(0)["facets"]["tags[]"]) == sorted([("nice", 300), ("cool", 300)])
    assert b.result(1)["num_hits"] == 600


def _pb(path, t1="die", t2="erbin"):
    return {"path": path, "search1": {"terms": [t1], "path": path}, "search2": {"terms": [t2], "path": path}}


def test_phrase_boosts(gpu, native_libs):  # tests/all/test_phrase.rs:39-99, test_scores.rs:106-126
    index, oracle = _pair(gpu, fx.TEST_PHRASE_DOCS, fx.TEST_PHRASE_CONFIG)
    b = compare(index, oracle, [
        {"search_req": S("erbin", "title"), "phrase_boosts": [_pb("title")]},
        {"search_req": {"or": {"queries": [S(t, p) for p in ("title", "tags[]") for t in ("die", "erbin")]}}, "phrase_boosts": [_pb("title"), _pb("tags[]")]},
        {"search_req": {"and": {"queries": [S("die", "title"), S("erbin", "title")]}}, "phrase_boosts": [_pb("title")]},
        {"search_req": {"or": {"queries": [S(t, "tags[]") for t in ("greg", "tagebuch", "05")]}},
         "phrase_boosts": [_pb("tags[]", "greg", "tagebuch"), _pb("tags[]", "tagebuch", "05")]},
        {"search_req": {"or": {"queries": [S("greg", "tags[]", levenshtein_distance=1), S("tagebuch", "tags[]")]}},
         "phrase_boosts": [{"path": "tags[]", "search1": {"terms": ["greg"], "path": "tags[]", "levenshtein_distance": 1}, "search2": {"terms": ["tagebuch"], "path": "tags[]"}}],
         "boost_term": [{"terms": ["05"], "path": "tags[]"}]},
    ])
    first = b.result(0)["data"][0][0]
    assert fx.TEST_PHRASE_DOCS[first]["title"] == "die erbin"
    index2, oracle2 = _pair(gpu, fx.TEST_SCORE_DOCS, fx.TEST_SCORE_CONFIG)
    compare(index2, oracle2, [
        {"search_req": {"or": {"queries": [S(t, "title") for t in ("greg", "tagebuch", "05")]}}, "phrase_boosts": [_pb("title", "greg", "tagebuch")]},
    ])


def test_text_locality(test_all, gpu, native_libs):  # tests/all/tests.rs:1296-1313
    index, oracle = test_all
    b = compare(index, oracle, [
        {"search_req": {"or": {"queries": [S("text", "meanings.ger[]"), S("localität", "meanings.ger[]")]}}, "text_locality": True},
        {"search_req": {"or": {"queries": [S("text", "meanings.ger[]"), S("localität", "meanings.ger[]"), S("urge", "meanings.eng[]")]}}, "text_locality": True,
         "boost": [{"path": "commonness", "boost_fun": "Log10", "param": 1}]},
        {"search_req": {"and": {"queries": [S("alle", "meanings.ger[]"), S("meine", "meanings.ger[]"), S("words", "meanings.ger[]")]}}, "text_locality": True},
        {"search_req": S("text", "meanings.ger[]"), "text_locality": True},
    ])
    first = b.result(0)["data"][0][0]
    assert fx.TEST_ALL_DOCS[first]["meanings"]["ger"][0] == "text localität"
    index2, oracle2 = _pair(gpu, fx.TEST_PHRASE_DOCS, fx.TEST_PHRASE_CONFIG)
    compare(index2, oracle2, [
        {"search_req": {"or": {"queries": [S(t, p) for p in ("title", "tags[]") for t in ("greg", "tagebuch", "05")]}}, "text_locality": True,
         "phrase_boosts": [_pb("tags[]", "greg", "tagebuch")]},
    ])


def test_config3_shape(gpu, native_libs):  # BASELINE config 3: AND + phrase + text locality + facets on a synthetic corpus
    params = dict(num_docs=30000, vocab=2500, seed=13, tags=40, text_locality=True, phrase=True)
    index, oracle = _pair(gpu, synth=params)
    reqs = helpers.synthetic_requests(num_queries=150, query_kind="and", levenshtein=1, query_seed=17, **params)
    assert "phrase_boosts" in reqs[0] and "text_locality" in reqs[0] and "facets" in reqs[0]
    b = compare(index, oracle, reqs)
    assert sum(b.result(q)["num_hits"] for q in range(len(reqs))) > 0


def test_one_to_n_boosts(test_all):  # tests/all/tests.rs:839-931 (BoostToAnchor + ApplyAnchorBoost)
    index, oracle = test_all
    compare(index, oracle, [
        {"search_req": S("意慾", "kanji[].text"), "boost": [{"path": "kanji[].commonness", "boost_fun": "Log10", "param": 1}]},
        {"search_req": S("awesome", "field1[].text"),
         "boost": [{"path": "commonness", "boost_fun": "Log10", "param": 1}, {"path": "field1[].rank", "expression": "10 / $SCORE", "skip_when_score": [0]}]},
        {"search_req": {"or": {"queries": [
            {"search": {"terms": ["awesome"], "path": "field1[].text", "options": {"boost": [{"path": "field1[].rank", "boost_fun": "Log10", "param": 1}]}}},
            {"search": {"terms": ["urge"], "path": "meanings.eng[]", "options": {"boost": [{"path": "commonness", "boost_fun": "Log10", "param": 1}]}}},
        ]}}},
        {"search_req": {"or": {"queries": [S("awesome", "field1[].text"), S("意慾", "kanji[].text", levenshtein_distance=1)]}},
         "boost": [{"path": "field1[].rank", "boost_fun": "Multiply"}, {"path": "kanji[].commonness", "boost_fun": "Add", "param": 2}, {"path": "commonness", "boost_fun": "Log2", "param": 2}]},
    ])


def _jmdict_request(term, lev):  # benches/bench_jmdict.rs:115-235, verbatim shape
    def part(path, boosts, starts_with):
        p = {"terms": [term], "path": path, "levenshtein_distance": lev, "options": {"boost": boosts}}
        if starts_with:
            p["starts_with"] = True
        return {"search": p}
    c1 = {"path": "commonness", "boost_fun": "Log10", "param": 1}
    return {"search_req": {"or": {"queries": [
        part("kanji[].text", [c1, {"path": "kanji[].commonness", "boost_fun": "Log10", "param": 1}], True),
        part("kana[].text", [c1, {"path": "kana[].commonness", "boost_fun": "Log10", "param": 1}], True),
        part("kana[].text", [c1, {"path": "kana[].commonness", "boost_fun": "Log10", "param": 1}], True),
        part("meanings.ger[].text", [{"path": "commonness", "boost_fun": "Log10", "param": 0}, {"path": "meanings.ger[].rank", "expression": "10 / $SCORE"}], False),
        part("meanings.eng[]", [c1], False),
    ], "options": {"top": 10, "skip": 0}}}}


def test_config1_jmdict_shape(gpu, native_libs):  # BASELINE config 1: jmdict-shaped corpus, the 5-way OR of bench_jmdict
    rng = np.random.default_rng(7)
    syll = ["ka", "ki", "ku", "ke", "ko", "sa", "shi", "su", "ta", "chi", "to", "na", "ni", "no", "ma", "mi", "mo", "ra", "ri", "ru"]
    words = ["".join(rng.choice(syll, size=int(rng.integers(2, 5)))) for _ in range(600)]
    eng = ["will", "urge", "house", "water", "test", "light", "dark", "river", "stone", "tree", "majesty", "appearance", "long", "torso"]
    docs = []
    for i in range(2500):
        d = {"ent_seq": str(100000 + i)}
        if rng.random() < 0.7:
            d["commonness"] = int(rng.integers(1, 3000))
        # at most one boosted value per 1:n field and document: with several, the reference's merge join
        # (boost.rs:255-281) applies the first or all of them depending on the position in the hit list, see DESIGN.md section 8
        def values(texts, key, lo, hi):
            boosted = int(rng.integers(0, len(texts) + 1))
            return [{"text": t, **({key: int(rng.integers(lo, hi))} if i == boosted else {})} for i, t in enumerate(texts)]
        if rng.random() < 0.8:
            d["kanji"] = values([str(rng.choice(words)) for _ in range(int(rng.integers(1, 3)))], "commonness", 1, 500)
        d["kana"] = values([str(rng.choice(words)) for _ in range(int(rng.integers(1, 3)))], "commonness", 1, 500)
        d["meanings"] = {
            "ger": values([" ".join(rng.choice(words, size=int(rng.integers(1, 4)))) for _ in range(int(rng.integers(0, 3)))], "rank", 1, 10),
            "eng": [" ".join(rng.choice(eng, size=int(rng.integers(1, 4)))) for _ in range(int(rng.integers(1, 3)))],
        }
        docs.append(d)
    config = {
        "kanji[].text": {"fulltext": {"tokenize": True}}, "kana[].text": {"fulltext": {"tokenize": True}},
        "meanings.ger[].text": {"fulltext": {"tokenize": True}}, "meanings.eng[]": {"fulltext": {"tokenize": True}},
        "commonness": dict(fx.BOOST), "kanji[].commonness": dict(fx.BOOST), "kana[].commonness": dict(fx.BOOST), "meanings.ger[].rank": dict(fx.BOOST),
    }
    index, oracle = _pair(gpu, docs, config)
    terms = [str(w) for w in rng.choice(words, size=25)] + [str(w)[:3] for w in rng.choice(words, size=10)] + ["will", "urge", "wate", "majesty"]
    reqs = [_jmdict_request(t, lev) for t in terms for lev in (0, 1)]
    b = compare(index, oracle, reqs)
    assert sum(b.result(q)["num_hits"] for q in range(len(reqs))) > 0


def test_one_to_n_boost_with_several_values_per_document(gpu, native_libs):
    """Several boosted values of one document in one 1:n field: the reference applies the first or all of them depending on
    the run of boosted hits before the anchor (apply_boost_values_anchor, boost.rs:255-281)."""
    rng = np.random.default_rng(11)
    syll = ["ka", "ki", "ku", "mi", "mo", "ra", "ri", "ru", "sa", "to"]
    words = ["".join(rng.choice(syll, size=int(rng.integers(2, 4)))) for _ in range(120)]
    docs = []
    for i in range(1500):
        n = int(rng.integers(0, 4))
        d = {"ent_seq": str(i)}
        if n:
            d["kana"] = [{"text": str(rng.choice(words)), **({"commonness": int(rng.integers(1, 900))} if rng.random() < 0.75 else {})} for _ in range(n)]
        docs.append(d)
    index, oracle = _pair(gpu, docs, {"kana[].text": {"fulltext": {"tokenize": True}}, "kana[].commonness": dict(fx.BOOST)})
    reqs = []
    for t in [str(w) for w in rng.choice(words, size=12)] + ["mi", "ka", "ri"]:
        for fun in ("Log10", "Multiply", "Add"):
            reqs.append({"search_req": {"search": {"terms": [t], "path": "kana[].text", "levenshtein_distance": 1, "starts_with": True}},
                         "boost": [{"path": "kana[].commonness", "boost_fun": fun, "param": 1}], "top": 50})
    compare(index, oracle, reqs)


# ---- suggest (search_field.rs:147-228)
def _same_suggestions(got, want, ctx):
    assert [(t, i) for t, _, i in got] == [(t, i) for t, _, i in want], ctx
    for (_, gs, _), (_, cs, _) in zip(got, want):
        assert abs(float(gs) - float(cs)) <= 1e-5 * max(abs(float(cs)), 1e-30), ctx


def test_suggest_on_the_reference_corpus(gpu, native_libs):
    """The reference's suggest tests (tests.rs:1087-1132) on its own corpus: the golden lists as sets (its order among equal
    scores is its unstable sort's), and the CUDA path against the oracle element by element (both break ties the same way)."""
    import ref_fixtures as fx
    d = tempfile.mkdtemp(prefix="vb200_gpu_")
    helpers.create_index(d, fx.TEST_ALL_DOCS, fx.TEST_ALL_CONFIG)
    index, oracle = gpu.Index(d), helpers.Oracle(d)
    part = {"terms": ["majes"], "path": "meanings.ger[]", "levenshtein_distance": 0, "starts_with": True, "top": 10, "skip": 0}
    got = index.suggest(part)
    assert sorted(t for t, _, _ in got) == sorted(["majestät", "majestät (f)", "majestätisches", "majestätischer", "majestätischer anblick (m)", "majestätisches aussehen (n)"])
    _same_suggestions(got, oracle.call("suggest", part=part), part)
    req = {"suggest": [{"terms": ["will"], "path": "meanings.ger[]", "levenshtein_distance": 0, "starts_with": True},
                       {"terms": ["will"], "path": "meanings.eng[]", "levenshtein_distance": 0, "starts_with": True}], "top": 10, "skip": 0}
    got = index.suggest_multi(req)
    assert sorted(t for t, _, _ in got) == sorted(["will", "wille", "wille (m)", "will testo"])
    _same_suggestions(got, oracle.call("suggest_multi", request=req), req)
    for req in ({"suggest": [{"terms": ["majes"], "path": "meanings.ger[]", "starts_with": True}], "top": 2, "skip": 1},
                {"suggest": [{"terms": ["Majestat"], "path": "meanings.ger[]", "levenshtein_distance": 2}], "top": None},
                {"suggest": [], "top": 10},
                {"suggest": [{"terms": ["zzzzzz"], "path": "meanings.ger[]"}]}):
        _same_suggestions(index.suggest_multi(req), oracle.call("suggest_multi", request=req), req)
    with pytest.raises(Exception):
        index.suggest_multi({"search_req": {"search": {"terms": ["a"], "path": "meanings.ger[]"}}})  # "only suggest allowed in suggest function"
    with pytest.raises(Exception):
        index.suggest({"terms": ["a"], "path": "nope"})


def test_token_value_boosts(gpu, native_libs):
    """`token_value` of a search part (search_field.rs:391-395): add_boost over the part's term hits with the
    1:1 store create/token_values_to_tokens.rs writes.  The reference's own case (tests.rs:1134-1158) on its corpus, then a
    20k-term dictionary where a third of the terms carry values: every boost function, an expression, skip_when_score,
    together with the per-part top bound and the part boost, in suggest, in the FieldSearch step and inside whole
    requests (alone, in or / and trees, with a request boost), from an exported plan too."""
    d = tempfile.mkdtemp(prefix="vb200_gpu_")
    helpers.create_index(d, fx.TEST_ALL_DOCS, fx.TEST_ALL_CONFIG)
    helpers.add_token_values(d, *fx.TEST_ALL_TOKEN_VALUES)
    index, oracle = gpu.Index(d), helpers.Oracle(d)
    tv = {"path": "meanings.ger[]", "boost_fun": "Log10", "param": 1}
    part = {"terms": ["begeist"], "path": "meanings.ger[]", "levenshtein_distance": 0, "starts_with": True, "token_value": tv, "top": 10, "skip": 0}
    got = index.suggest(part)
    assert [t for t, _, _ in got] == ["begeisterung", "begeistern", "begeisterung (f)"]
    _same_suggestions(got, oracle.call("suggest", part=part), part)
    compare(index, oracle, [{"search_req": {"search": {k: v for k, v in part.items() if k not in ("top", "skip")}}},
                            {"search_req": {"search": part}},
                            {"search_req": {"or": {"queries": [{"search": part}, S("urge", "meanings.eng[]")]}}}])
    with pytest.raises(Exception):  # no token values on this field: "Did not found path in indices"
        index.suggest({"terms": ["will"], "path": "meanings.eng[]", "token_value": {"path": "meanings.eng[]", "boost_fun": "Log10"}})
    b = index.prepare([json.dumps({"search_req": {"search": {"terms": ["will"], "path": "meanings.eng[]", "token_value": {"path": "meanings.eng[]"}}}}),
                       json.dumps({"search_req": S("urge", "meanings.eng[]")})]).execute()
    assert b.status(0) == 3 and b.status(1) == 0, (b.status(0), b.message(0))

    from test_part_hits import make_valued_index, token_value_parts
    d, oracle, words, _ = make_valued_index()
    index = gpu.Index(d)
    P = lambda t, **kw: {"search": {"terms": [t], "path": "body", **kw}}
    reqs = []
    for i, part in enumerate(token_value_parts(words)):
        w = words[i]
        _same_suggestions(index.suggest(part), oracle.call("suggest", part=part), part)
        hits, _ = index.field_search(part)
        ref = sorted(oracle.call("field_search", part=part)["hits_scores"])
        assert [h[0] for h in sorted(hits)] == [h[0] for h in ref], part
        for (_, gs), (_, cs) in zip(sorted(hits), ref):
            assert abs(float(gs) - float(cs)) <= 1e-5 * max(abs(float(cs)), 1e-30), part
        shape = i % 4
        other = P(words[(i + 5) % len(words)], levenshtein_distance=1)
        if shape == 0:
            reqs.append({"search_req": {"search": part}})
        elif shape == 1:
            reqs.append({"search_req": {"or": {"queries": [{"search": part}, other]}}, "boost": [{"path": "commonness", "boost_fun": "Log10", "param": 1}]})
        elif shape == 2:
            reqs.append({"search_req": {"and": {"queries": [{"search": part}, P(w[:1], starts_with=True)]}}})
        else:
            reqs.append({"search_req": {"or": {"queries": [{"search": part}, {"search": {**part, "token_value": {"path": "body", "boost_fun": "Multiply"}}}]}}, "top": 5})
    batch = compare(index, oracle, reqs)
    texts = [json.dumps(r) for r in reqs]
    again = index.prepare(len(texts), plan=batch.export_plan()).execute()
    for q in range(len(texts)):
        assert again.status(q) == 0 and again.result(q) == batch.result(q), texts[q]


def test_suggest_bounds_and_boosts_on_a_large_dictionary(gpu, native_libs):
    """Prefix and fuzzy suggestions over a 20k-term dictionary: hundreds of matches per part, so the per-part top bound
    (keep top + skip + 200, cut, drop what scores below the worst kept) is exercised, with part boosts (also negative),
    several parts on one field merging equal texts, and skip."""
    synth = dict(num_docs=30000, vocab=20000, seed=11)
    d = tempfile.mkdtemp(prefix="vb200_gpu_")
    helpers.create_synthetic_index(d, **synth)
    index, oracle = gpu.Index(d), helpers.Oracle(d)
    words = [json.loads(r)["search_req"]["search"]["terms"][0] for r in helpers.synthetic_requests(num_queries=40, query_kind="single", query_seed=13, **synth)]
    n_long = 0
    for i, w in enumerate(words):
        parts = [{"terms": [w[:1 + i % 3]], "path": "body", "starts_with": True, "levenshtein_distance": i % 2}]
        if i % 2 == 0:
            parts[0]["top"], parts[0]["skip"] = 1 + i % 7, i % 3
        if i % 3 == 0:
            parts.append({"terms": [w], "path": "body", "levenshtein_distance": 2, "boost": -1.5 if i % 6 == 0 else 3.0})
        if i % 5 == 0:
            parts.append({"terms": [w[:2]], "path": "body", "starts_with": True, "top": 300})
        req = {"suggest": parts, "top": 25 if i % 4 else None, "skip": i % 4}
        got = index.suggest_multi(req)
        want = oracle.call("suggest_multi", request=req)
        n_long += len(index.field_search({k: v for k, v in parts[0].items() if k not in ("top", "skip")})[0]) > 210
        _same_suggestions(got, want, req)
        one = dict(parts[0])
        _same_suggestions(index.suggest(one), oracle.call("suggest", part=one), one)
    assert n_long >= 5, "no part matched enough terms to reach the bound"


def test_per_part_top_and_skip_in_search_requests(gpu, native_libs):
    """a RequestSearchPart with `top` keeps only its best top + skip term matches (with the reference's
    200-hit slack and its drop-below-the-worst rule, search_field.rs:292-294,322-331,366-369) before they are resolved to
    anchors.  Parts that match hundreds of terms, alone, inside or/and trees with unbounded parts, with part boosts and
    with request boosts."""
    synth = dict(num_docs=30000, vocab=20000, seed=11)
    d = tempfile.mkdtemp(prefix="vb200_gpu_")
    helpers.create_synthetic_index(d, **synth)
    index, oracle = gpu.Index(d), helpers.Oracle(d)
    words = [json.loads(r)["search_req"]["search"]["terms"][0] for r in helpers.synthetic_requests(num_queries=60, query_kind="single", query_seed=17, **synth)]
    P = lambda t, **kw: {"search": {"terms": [t], "path": "body", **kw}}
    boost = [{"path": "commonness", "boost_fun": "Log10", "param": 1}]
    reqs = []
    for i, w in enumerate(words):
        bounded = P(w[:1 + i % 3], starts_with=True, top=1 + i % 9, skip=i % 3)
        fuzzy = P(w, levenshtein_distance=2, top=2 + i % 4)
        if i % 4 == 0:
            bounded["search"]["boost"] = 2.0 if i % 8 else -1.0
        shape = i % 6
        if shape == 0:
            r = {"search_req": bounded}
        elif shape == 1:
            r = {"search_req": fuzzy, "boost": boost}
        elif shape == 2:
            r = {"search_req": {"or": {"queries": [bounded, P(words[(i + 1) % len(words)], levenshtein_distance=1), fuzzy]}}, "boost": boost}
        elif shape == 3:
            r = {"search_req": {"and": {"queries": [P(w[:2], starts_with=True, top=400), P(words[(i + 7) % len(words)][:1], starts_with=True)]}}}
        elif shape == 4:
            r = {"search_req": {"or": {"queries": [fuzzy, P(w[:3], starts_with=True, levenshtein_distance=1, top=7)]}}, "top": 3}
        else:
            r = {"search_req": {"or": {"queries": [bounded, bounded, P(w[:2], starts_with=True, top=5)]}}, "top": 20, "skip": 3}
        reqs.append(json.dumps(r))
    out, ref = compare_batch(index, oracle, reqs, k=10)
    assert (out["num_hits"] > 0).sum() > len(reqs) // 2
    # the bound is visible: the same part without `top` finds more
    unbounded = index.search_batch([json.dumps({"search_req": P(words[0][:1], starts_with=True)})], k=10)
    assert int(unbounded["num_hits"][0]) > int(out["num_hits"][0])
    # the step seam takes bounded parts too
    for i, w in enumerate(words[:12]):
        part = {"terms": [w[:1 + i % 2]], "path": "body", "starts_with": True, "top": 3 + i, "skip": i % 2}
        if i % 3 == 0:
            part["boost"] = 1.5
        hits, _ = index.field_search(part)
        want = oracle.call("field_search", part=part)["hits_scores"]
        assert sorted((h[0], round(float(h[1]), 4)) for h in hits) == sorted((h[0], round(float(h[1]), 4)) for h in want), part


def test_per_part_top_does_not_bound_ids(test_all):
    """Filter, boost_term and phrase parts are searched for ids, which are collected before the per-part bound
    (search_field.rs:305-307): `top` on them changes nothing."""
    index, oracle = test_all
    compare(index, oracle, [
        {"search_req": OR_MAJ_URGE, "filter": S("1587690", "ent_seq", top=1)},
        {"search_req": OR_MAJ_URGE, "filter": {"or": {"queries": [S("15", "ent_seq", starts_with=True, top=1), S("urge", "meanings.eng[]", top=0)]}}},
        {"search_req": S("will", "meanings.eng[]"), "boost_term": [{"terms": ["9"], "path": "ent_seq", "starts_with": True, "boost": 5.0, "top": 1}]},
        {"search_req": S("will", "meanings.eng[]", top=1), "boost_term": [{"terms": ["9555"], "path": "ent_seq", "boost": 5.0, "top": 0}]},
        {"search_req": S("maje", "meanings.ger[]", starts_with=True, top=1)},
        {"search_req": {"or": {"queries": [S("maje", "meanings.ger[]", starts_with=True, top=2, skip=1), S("urge", "meanings.eng[]")]}}},
    ])
