"""`explain` (src/search/result/explain.rs): the oracle's explain maps pinned on the reference's two explain tests, and the
product's explain walk (csrc/host/explain_walk.hpp) against the oracle on the CPU.  In the product the parts' term hits come
from the device match and the posting weights from posting_lookup_kernel (tests/test_gpu_round2.py::test_explain); here the
oracle supplies the former and the index files the latter, so the walk itself -- order of the explanations through
FieldSearch, Resolve, Union, Intersect, request boosts and token values -- is checked without a device."""
import json
import tempfile

import pytest

import helpers
import ref_fixtures as fx
from test_part_hits import SYNTH, make_valued_index


@pytest.fixture(scope="module")
def test_all(native_libs):
    d = tempfile.mkdtemp(prefix="vb200_explain_")
    helpers.create_index(d, fx.TEST_ALL_DOCS, fx.TEST_ALL_CONFIG)
    helpers.add_token_values(d, *fx.TEST_ALL_TOKEN_VALUES)
    return d, helpers.Oracle(d)


S = lambda term, path, **kw: {"search": {"terms": [term], "path": path, **kw}}


def test_reference_explain_tests(test_all):
    _, o = test_all
    # tests.rs:347-364 simple_search_explained
    res = o.search({"search_req": S("urge", "meanings.eng[]", options={"explain": True})})
    assert len(res["data"]) == 1 and fx.TEST_ALL_DOCS[res["data"][0][0]]["ent_seq"] == "1587690"
    ex = res["explain"][str(res["data"][0][0])]
    assert len(ex) == 2
    assert list(ex[0]) == ["TermToAnchor"] and ex[1] == {"LevenshteinScore": {"score": 10, "text_or_token_id": "urge", "term_id": ex[0]["TermToAnchor"]["term_id"]}}
    assert ex[0]["TermToAnchor"]["term_score"] == 10 and abs(ex[0]["TermToAnchor"]["final_score"] - 10 * ex[0]["TermToAnchor"]["anchor_score"]) < 1e-4
    # tests.rs:366-391 or_query_explained
    res = o.search({"search_req": {"or": {"queries": [S("majestät", "meanings.ger[]"), S("urge", "meanings.eng[]")]}}, "explain": True})
    assert len(res["data"]) == 2 and fx.TEST_ALL_DOCS[res["data"][0][0]]["ent_seq"] == "1587690"
    ex = res["explain"][str(res["data"][0][0])]
    assert len(ex) == 5 and [list(e)[0] for e in ex] == ["TermToAnchor", "LevenshteinScore", "OrSumOverDistinctTerms", "TermToAnchor", "LevenshteinScore"]
    # no explain asked: no explanations
    assert "explain" not in o.search({"search_req": S("urge", "meanings.eng[]")})


def walk_against_oracle(d, o, request):
    res = o.search(request)
    anchors = [h[0] for h in res["data"]]
    leaves = [o.call("field_search", part=helpers.bare_part(p))["hits_scores"] for p in helpers.tree_parts(request["search_req"])]
    got = helpers.explain_walk(d, request, anchors, leaves)
    want = {int(k): v for k, v in res.get("explain", {}).items()}
    for a in anchors:  # an anchor nothing was recorded for has no entry in the reference's map
        want.setdefault(a, [])
    assert helpers.same_explain(got, want), (json.dumps(request, ensure_ascii=False), got, want)
    return want


def reference_corpus_requests():
    ger, eng = "meanings.ger[]", "meanings.eng[]"
    boost = [{"path": "commonness", "boost_fun": "Log10", "param": 1}]
    return [
        {"search_req": S("urge", eng, options={"explain": True})},
        {"search_req": {"or": {"queries": [S("majestät", ger), S("urge", eng)]}}, "explain": True},
        {"search_req": S("majestätischer", ger, levenshtein_distance=1), "explain": True},  # tests.rs:455-468: two tokens hit the same anchor
        {"search_req": {"or": {"queries": [S("majestät", ger, levenshtein_distance=1), S("majestät", eng, levenshtein_distance=1), S("urge", eng)]}}, "explain": True, "boost": boost},
        {"search_req": {"or": {"queries": [S("will", ger, starts_with=True), S("will", eng, starts_with=True, boost=2.5)]}}, "explain": True, "top": 20},  # the same term in two fields: one slot
        {"search_req": {"and": {"queries": [S("majestät", ger, levenshtein_distance=2), S("majestätisches", ger, levenshtein_distance=3)]}}, "explain": True},
        {"search_req": {"or": {"queries": [S("urge", eng), S("majestät", ger, options={"explain": True})]}}},  # only one part asks: the first input decides (set_op.rs:120)
        {"search_req": {"or": {"queries": [S("urge", eng, options={"explain": True}), S("majestät", ger)]}}},
        {"search_req": S("weich", ger, levenshtein_distance=1), "explain": True, "boost": [{"path": "commonness", "boost_fun": "Multiply", "expression": "$SCORE * 2", "skip_when_score": [7.5]},
                                                                                           {"path": "commonness", "boost_fun": "Add", "param": 3}]},
        {"search_req": S("begeist", ger, starts_with=True, token_value={"path": ger, "boost_fun": "Log10", "param": 1}), "explain": True},
        {"search_req": S("begeist", ger, starts_with=True, top=2, boost=0.5, token_value={"path": ger, "boost_fun": "Multiply"}), "explain": True},
        {"search_req": S("majestät", ger, levenshtein_distance=1), "filter": S("20", "commonness"), "explain": True},
        {"search_req": S("nothing matches this", ger), "explain": True},
    ]


UNSUPPORTED_REQUESTS = [  # outside the reconstruction: 1:n boosts, phrase boosts
    {"search_req": S("意慾", "kanji[].text"), "explain": True, "boost": [{"path": "kanji[].commonness", "boost_fun": "Log10", "param": 1}]},  # tests.rs:839-860
    {"search_req": S("urge", "meanings.eng[]"), "explain": True, "phrase_boosts": [{"search1": {"terms": ["a"], "path": "meanings.eng[]"}, "search2": {"terms": ["b"], "path": "meanings.eng[]"}}]},
]


def test_walk_on_the_reference_corpus(test_all):
    d, o = test_all
    requests = reference_corpus_requests()
    n_items = 0
    for r in requests:
        want = walk_against_oracle(d, o, r)
        n_items += sum(len(v) for v in want.values())
    assert n_items > 60
    for r in UNSUPPORTED_REQUESTS:
        with pytest.raises(helpers.OracleError) as e:
            helpers.explain_walk(d, r, [1], [[]])
        assert e.value.status == 8  # VGPU_ERR_UNSUPPORTED: outside the reconstruction


def large_dictionary_requests(words):
    P = lambda t, **kw: {"search": {"terms": [t], "path": "body", **kw}}
    boost = [{"path": "commonness", "boost_fun": "Log10", "param": 1}]
    for i, w in enumerate(words[:24]):
        wide, narrow, other = P(w[:2], starts_with=True), P(w, levenshtein_distance=1), P(words[(i + 3) % len(words)], levenshtein_distance=1)
        shape = i % 6
        if shape == 0:
            r = {"search_req": {"or": {"queries": [narrow, other, P(words[(i + 9) % len(words)])]}}, "boost": boost}
        elif shape == 1:
            r = {"search_req": {"and": {"queries": [wide, narrow]}}}
        elif shape == 2:
            r = {"search_req": {"and": {"queries": [narrow, wide]}}, "boost": boost}
        elif shape == 3:
            r = {"search_req": {"or": {"queries": [{"and": {"queries": [wide, narrow]}}, other]}}}
        elif shape == 4:
            wide["search"]["token_value"] = {"path": "body", "boost_fun": ["Log10", "Multiply", "Replace"][i % 3], "param": 1.5}
            wide["search"]["top"] = 5 + i
            r = {"search_req": {"or": {"queries": [wide, other]}}, "top": 15}
        else:
            r = {"search_req": wide, "boost": boost, "top": 12, "skip": 3}
        r["explain"] = True
        yield r


def test_walk_on_a_large_dictionary(native_libs):
    """Parts that match hundreds of terms, several of them on one anchor; `and` inputs of clearly different lengths (the
    engine orders them by posting counts, the reference by distinct anchors); token values; per-part top; request boosts."""
    d, o, words, _ = make_valued_index()
    n_items = n_multi = 0
    for r in large_dictionary_requests(words):
        want = walk_against_oracle(d, o, r)
        n_items += sum(len(v) for v in want.values())
        n_multi += sum(1 for v in want.values() if sum(1 for e in v if "TermToAnchor" in e) > 2)
    assert n_items > 500 and n_multi > 5


def test_reference_explain_tests_through_the_query_generator(native_libs):
    """test_query_generator.rs:139-152 (5 entries) and :154-168 (3 hits, 7 entries for the first): requests made by the
    product's query generator over every field of the reference's query-generator corpus; the oracle's explain lengths as
    the reference asserts them, and the product's walk against the oracle on these many-part `or`s."""
    d = tempfile.mkdtemp(prefix="vb200_explain_qg_")
    helpers.create_index(d, fx.TEST_QG_DOCS, fx.TEST_QG_CONFIG)
    o = helpers.Oracle(d)
    for params, n_hits, n_explain in (({"search_term": "urge", "explain": True}, 1, 5), ({"search_term": "urge OR いよく", "explain": True}, 3, 7)):
        rc, req, raw = helpers.generate_request(d, params)
        assert rc == 0 and req["explain"] is True, raw
        res = o.search(req)
        assert len(res["data"]) == n_hits and fx.TEST_QG_DOCS[res["data"][0][0]]["ent_seq"] == "1587690"
        assert len(res["explain"][str(res["data"][0][0])]) == n_explain
        walk_against_oracle(d, o, req)
