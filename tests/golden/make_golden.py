#!/usr/bin/env python
"""Writes the golden result files of this directory with the CPU oracle (oracle/veloci_oracle.cpp, pinned against the
reference's own vectors by tests/test_oracle_kats.py and tests/test_oracle_reference_fixtures.py).

    python tests/golden/make_golden.py

* reference_corpus.json : the corpus of the reference's tests/all/tests.rs (tests/ref_fixtures.py) with the requests of
  tests/test_gpu_parity.py::_requests_test_all and the feature requests (filters, boost_term, facets, suggest).
* synthetic_small.json  : a seeded synthetic index in the shape of BASELINE config 2 (60k docs, 5k terms) with 3-term OR,
  AND and single-term fuzzy requests.

Each file holds {"index": how to rebuild the index, "cases": [{"request", "num_hits", "data": [[anchor, score], ...]}]}
(suggest cases hold {"suggest", "items": [[text, score, term id], ...]}).  tests/test_golden.py checks the oracle (CPU) and
the CUDA path (GPU) against them; nothing here is read from /root/reference at test time.
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import helpers  # noqa: E402
import ref_fixtures as fx  # noqa: E402

SMALL = dict(num_docs=60000, vocab=5000, seed=7)


def S(term, path, **kw):
    return {"search": {"terms": [term], "path": path, **kw}}


def reference_requests():
    or_maj_urge = {"or": {"queries": [S("majestät", "meanings.ger[]"), S("urge", "meanings.eng[]")]}}
    boost = [{"path": "commonness", "boost_fun": "Log10", "param": 1}]
    return [
        {"search_req": S("urge", "meanings.eng[]")},
        {"search_req": S("majestätischer", "meanings.ger[]", levenshtein_distance=1)},
        {"search_req": S("Majestätischer", "meanings.ger[]", ignore_case=False)},
        {"search_req": or_maj_urge},
        {"search_req": or_maj_urge, "top": 1},
        {"search_req": {"and": {"queries": [S("alle", "meanings.ger[]"), S("meine", "meanings.ger[]"), S("words", "meanings.ger[]")]}}},
        {"search_req": S("awesome", "field1[].text"), "boost": boost},
        {"search_req": S("意慾", "kanji[].text"), "boost": boost},
        {"search_req": S("COllectif", "title"), "boost": [{"path": "commonness", "boost_fun": "Log2", "param": 2}]},
        {"search_req": S("weich", "meanings.ger[]", levenshtein_distance=1)},
        {"search_req": S("ein", "meanings.ger[]", starts_with=True)},
        {"search_req": S("will", "meanings.eng[]", boost=3.0)},
        {"search_req": or_maj_urge, "filter": S("1587690", "ent_seq")},
        {"search_req": or_maj_urge, "filter": {"or": {"queries": [S("1587690", "ent_seq"), S("urge", "meanings.eng[]")]}}},
        {"search_req": S("will", "meanings.eng[]"), "boost_term": [{"terms": ["9555"], "path": "ent_seq", "boost": 5.0}]},
        {"search_req": S("maje", "meanings.ger[]", starts_with=True, top=2, skip=1)},
    ]


def suggest_requests():
    return [
        {"suggest": [{"terms": ["majes"], "path": "meanings.ger[]", "levenshtein_distance": 0, "starts_with": True}], "top": 10, "skip": 0},
        {"suggest": [{"terms": ["will"], "path": "meanings.ger[]", "levenshtein_distance": 0, "starts_with": True},
                     {"terms": ["will"], "path": "meanings.eng[]", "levenshtein_distance": 0, "starts_with": True}], "top": 10, "skip": 0},
    ]


def synthetic_requests():
    reqs = helpers.synthetic_requests(num_queries=40, query_kind="or3", levenshtein=1, query_seed=101, **SMALL)
    reqs += helpers.synthetic_requests(num_queries=20, query_kind="and", levenshtein=1, query_seed=102, **SMALL)
    reqs += helpers.synthetic_requests(num_queries=20, query_kind="single", levenshtein=2, query_seed=103, **SMALL)
    return reqs


def build_reference_index():
    d = tempfile.mkdtemp(prefix="vb200_golden_")
    helpers.create_index(d, fx.TEST_ALL_DOCS, fx.TEST_ALL_CONFIG)
    return d


def build_synthetic_index():
    d = tempfile.mkdtemp(prefix="vb200_golden_")
    helpers.create_synthetic_index(d, **SMALL)
    return d


def oracle_cases(oracle, requests):
    cases = []
    for r in requests:
        text = r if isinstance(r, str) else json.dumps(r, ensure_ascii=False)
        res = oracle.search(text)
        cases.append({"request": json.loads(text), "num_hits": res["num_hits"], "data": [[h[0], float(h[1])] for h in res["data"]]})
    return cases


def main():
    d = build_reference_index()
    o = helpers.Oracle(d)
    cases = oracle_cases(o, reference_requests())
    for r in suggest_requests():
        cases.append({"suggest": r, "items": [[t, float(s), i] for t, s, i in o.call("suggest_multi", request=r)]})
    json.dump({"index": "tests/ref_fixtures.py TEST_ALL_DOCS / TEST_ALL_CONFIG", "cases": cases}, open(os.path.join(HERE, "reference_corpus.json"), "w"), ensure_ascii=False, indent=1)
    d = build_synthetic_index()
    o = helpers.Oracle(d)
    json.dump({"index": {"synthetic": SMALL}, "cases": oracle_cases(o, synthetic_requests())}, open(os.path.join(HERE, "synthetic_small.json"), "w"), ensure_ascii=False, indent=1)


if __name__ == "__main__":
    main()
