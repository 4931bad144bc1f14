"""Request generation: the query language and `search_query` / `suggest_query`.

Three implementations meet here: the reference's own unit vectors (restated below with the file:line they come from),
the plain-Python oracle (oracle/query_generator.py) and the product's host code (csrc/host/query_parser.hpp,
query_generator.hpp, reached through the host-only helper library: no device needed).  The vectors pin the oracle AND
the product; random queries hold the product against the oracle; the generated requests are then run through the CPU
search oracle on the reference's query-generator corpus (tests/all/test_query_generator.rs) for the hits the reference
asserts.  The same requests run on the GPU in tests/test_gpu_round2.py.
"""
import json
import os
import random
import sys
import tempfile

import pytest

import helpers
import ref_fixtures as fx

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import query_generator as qg  # noqa: E402  (test infrastructure)

LIT, ATTR, PO, PC, TILDE, OR, AND = "Literal", "AttributeLiteral", "ParenthesesOpen", "ParenthesesClose", "Tilde", "Or", "And"


def both_tokens(text):
    raw = text.encode("utf-8")
    ours = [(raw[a:b].decode("utf-8"), t) for a, b, t in qg.lex(text)]
    ok, got = helpers.query_parse(text, what=3)
    assert ok and [tuple(x) for x in got] == ours, (text, got, ours)
    return ours


# ---- query_parser/src/lexer.rs:248-325
LEXER_TEXTS = [
    ("    ", []),
    ("schlau (", ["schlau", "("]),
    (" schön und schlau", ["schön", "und", "schlau"]),
    ("schlau", ["schlau"]),
    ("schlau(", ["schlau", "("]),
    ("coolAND AND (", ["coolAND", "AND", "("]),
    ("ANDand AND    ", ["ANDand", "AND"]),
    ("(cool)", ["(", "cool", ")"]),
    ("(cool OR nice)AND", ["(", "cool", "OR", "nice", ")", "AND"]),
    ('"my quote"', ["my quote"]),
    ('asdf"', ['asdf"']),
    ('"asdf"', ["asdf"]),
    ('tes"tco"ol', ['tes"tco"ol']),
    ("cool:nice", ["cool", "nice"]),
    ('"cool":nice', ["cool", "nice"]),
]
LEXER_TYPES = [
    ("coolAND AND (", [LIT, AND, PO]),
    ("or OR", [LIT, LIT]),
    ("OR OR", [LIT, LIT]),
    ("OR OR OR", [LIT, OR, LIT]),
    ("AND AND", [LIT, LIT]),
    ("AND AND AND", [LIT, AND, LIT]),
    ("ANDand AND    ", [LIT, AND]),
    ("(cool)", [PO, LIT, PC]),
    ("(cool OR nice)AND", [PO, LIT, OR, LIT, PC, LIT]),
    ("or~", [LIT, TILDE]),
    ("~~", [TILDE, TILDE]),
    ("~  ~", [TILDE, TILDE]),
    ("~a~", [TILDE, LIT, TILDE]),
    ("cool:nice", [ATTR, LIT]),
    ('"cool":nice', [ATTR, LIT]),
]


@pytest.mark.parametrize("text,expected", LEXER_TEXTS)
def test_lexer_token_texts(native_libs, text, expected):
    assert [t for t, _ in both_tokens(text)] == expected


@pytest.mark.parametrize("text,expected", LEXER_TYPES)
def test_lexer_token_types(native_libs, text, expected):
    assert [k for _, k in both_tokens(text)] == expected


# ---- query_parser/src/parser.rs:203-480: (query, options, Debug text of the expected tree)
SAME = {}
PARSER = [
    ("hallo", SAME, '"hallo"'),
    ('"cool")', SAME, '"cool"'),
    ('"cooles teil")', SAME, '"cooles teil"'),
    ("(cool)", SAME, '"cool"'),
    ("((((((cool))))))", SAME, '"cool"'),
    ("((((((cool)))))) AND ((((((cool))))))", SAME, '("cool" AND "cool")'),
    ("(super AND cool) OR fancy", SAME, '(("super" AND "cool") OR "fancy")'),
    ("(super AND cool) OR (fancy)", SAME, '(("super" AND "cool") OR "fancy")'),
    ("((super AND cool)) OR (fancy)", SAME, '(("super" AND "cool") OR "fancy")'),
    ("(cool)", {"no_parentheses": True}, '"(cool)"'),
    ("((((((cool)))))) AND ((((((cool))))))", {"no_parentheses": True}, '("((((((cool))))))" AND "((((((cool))))))")'),
    ("super AND cool OR fancy", SAME, '("super" AND ("cool" OR "fancy"))'),
    ("super OR cool AND fancy", SAME, '("super" OR ("cool" AND "fancy"))'),
    ("super cool OR fancy", SAME, '("super" OR ("cool" OR "fancy"))'),
    ("super cool", SAME, '("super" OR "cool")'),
    ("super OR cool", SAME, '("super" OR "cool")'),
    ("fancy~1", SAME, '"fancy"~1'),
    ("super cool OR fancy~1", SAME, '("super" OR ("cool" OR "fancy"~1))'),
    ("fancy~1", {"no_levensthein": True}, '"fancy~1"'),
    ("field:fancy~1", SAME, 'field:"fancy"~1'),
    ('"field":fancy unlimited', SAME, '(field:"fancy" OR "unlimited")'),
    ('"field""cool"', SAME, '("field" OR "cool")'),
    ("field:fancy", SAME, 'field:"fancy"'),
    ("field:fancy", {"no_attributes": True}, '"field:fancy"'),
    ("freestyle myattr:(super cool)", SAME, '("freestyle" OR myattr:("super" OR "cool"))'),
    ("field:(fancy unlimited)", SAME, 'field:("fancy" OR "unlimited")'),
    ("a AND  b", SAME, '("a" AND "b")'),
    ("die drei ???", SAME, '("die" OR ("drei" OR "???"))'),
    ("a+", SAME, '"a+"'),
    ("a AND b AND c", SAME, '("a" AND ("b" AND "c"))'),
    ("a OR b OR c", SAME, '("a" OR ("b" OR "c"))'),
    ("a AND b", SAME, '("a" AND "b")'),
    ("a:b", SAME, 'a:"b"'),
    ("a:b OR c", SAME, '(a:"b" OR "c")'),
    ("a", SAME, '"a"'),
    ("食べる AND b", SAME, '("食べる" AND "b")'),
    ("a OR b AND c", SAME, '("a" OR ("b" AND "c"))'),
    ("a b", SAME, '("a" OR "b")'),
    ('"a b"', SAME, '"a b"'),
    ("feld:10 b", SAME, '(feld:"10" OR "b")'),
]
PARSER_ERRORS = [
    ("field:what:ok", None),  # parser.rs:208-210: is_err
    ("fancy~", 'UnexpectedTokenType("fancy~﹏﹏", "Expecting a levenshtein number after a \'~\' ")'),  # parser.rs:292-298
    ("fancy:", 'UnexpectedTokenType("fancy:﹏﹏", "only token or ( allowed after attribute (\'attr:\') ")'),  # parser.rs:418-425
]


@pytest.mark.parametrize("text,opts,expected", PARSER)
def test_parser_vectors(native_libs, text, opts, expected):
    assert qg.debug(qg.parse(text, qg.Options(**opts))) == expected
    assert helpers.query_parse(text, **opts) == (True, expected)


@pytest.mark.parametrize("text,expected", PARSER_ERRORS)
def test_parser_error_vectors(native_libs, text, expected):
    with pytest.raises(qg.ParseError) as e:
        qg.parse(text)
    ok, msg = helpers.query_parse(text)
    assert not ok
    if expected is not None:
        assert str(e.value) == expected and msg == expected


# ---- query_parser/src/ast.rs:283-312 (phrase pairs, terms), :188-216 (filter_ast)
def test_phrase_pairs_and_terms(native_libs):
    cases = [
        ("super cool fancy", [["super", "cool"], ["cool", "fancy"]]),
        ("super cool fancy great", [["super", "cool"], ["cool", "fancy"], ["fancy", "great"]]),
        ("super cool nice great", [["super", "cool"], ["cool", "nice"], ["nice", "great"]]),
        ("myattr:(super cool)", [["super", "cool"]]),
        ("myattr:(super cool) different scope", [["super", "cool"], ["cool", "different"], ["different", "scope"]]),
    ]
    for text, pairs in cases:
        assert [list(p) for p in qg.phrase_pairs(qg.parse(text))] == pairs
        assert helpers.query_parse(text, what=1) == (True, pairs)
    assert qg.walk_terms(qg.parse("myattr:(super cool) AND fancy")) == ["super", "cool", "fancy"]
    assert helpers.query_parse("myattr:(super cool) AND fancy", what=2) == (True, ["super", "cool", "fancy"])


def test_filter_ast(native_libs):
    def drop(words):
        return lambda ast, attr: ast[0] == "leaf" and ast[1].lower() in words

    # ast.rs:188-199, :201-216
    assert qg.debug(qg.filter_ast(qg.parse("super cool fancy"), drop({"cool"}))) == '("super" OR "fancy")'
    assert helpers.query_filter_stopwords("super cool fancy", {"cool"}) == '("super" OR "fancy")'
    assert qg.filter_ast(qg.parse("myattr:(super cool)"), lambda a, b: True) is None
    assert helpers.query_filter_stopwords("myattr:(super cool)", {"super", "cool"}) == "None"
    assert qg.debug(qg.filter_ast(qg.parse("myattr:(super cool)"), drop({"cool"}))) == 'myattr:"super"'
    assert helpers.query_filter_stopwords("myattr:(super cool)", {"cool"}) == 'myattr:"super"'
    # query_parser_to_veloci_request.rs:181-199 (stop words from a user list; the bundled language lists are data files of the reference)
    assert qg.debug(qg.filter_ast(qg.parse("die erbin"), drop({"die"}))) == '"erbin"'
    assert helpers.query_filter_stopwords("Die erbin", {"die"}) == '"erbin"'


def test_random_queries_product_equals_oracle(native_libs):
    rng = random.Random(7)
    pieces = ["a", "b", "cool", "AND", "OR", "AND ", "OR ", " ", "  ", "(", ")", "~", "1", "2", "300", ":", '"', "*", "ü", "食べる", "\t", "　", "attr", "+", "x:y"]
    n_ok = n_err = 0
    for _ in range(4000):
        text = "".join(rng.choice(pieces) for _ in range(rng.randint(1, 9)))
        opts = {k: rng.random() < 0.2 for k in ("no_attributes", "no_parentheses", "no_levensthein")}
        try:
            ast = qg.parse(text, qg.Options(**opts))
        except qg.ParseError as e:
            ok, msg = helpers.query_parse(text, **opts)
            assert not ok, (text, opts, msg)
            if not str(e).startswith("panic"):
                assert msg == str(e), (text, opts)
            n_err += 1
            continue
        assert helpers.query_parse(text, **opts) == (True, qg.debug(ast)), (text, opts)
        assert helpers.query_parse(text, what=1, **opts) == (True, [list(p) for p in qg.phrase_pairs(ast)]), (text, opts)
        assert helpers.query_parse(text, what=2, **opts) == (True, qg.walk_terms(ast)), (text, opts)
        n_ok += 1
    assert n_ok > 1000 and n_err > 300


# ---- search_query on the reference's corpus
@pytest.fixture(scope="module")
def qg_index(native_libs):
    d = tempfile.mkdtemp(prefix="vb200_qg_")
    helpers.create_index(d, fx.TEST_QG_DOCS, fx.TEST_QG_CONFIG)
    meta = json.load(open(os.path.join(d, "metaData.json")))
    all_fields = list(meta["columns"])
    search_fields = [f for f, c in meta["columns"].items() if any(i["path"] == f + ".textindex.to_anchor_id_score" for i in c["indices"])]
    return d, qg.Catalog(all_fields, search_fields), helpers.Oracle(d)


def as_f32(v):
    """every float of a JSON value rounded to f32 (the product writes the shortest text of the f32, Python its f64 digits)"""
    if isinstance(v, float):
        return qg.f32(v)
    if isinstance(v, list):
        return [as_f32(x) for x in v]
    if isinstance(v, dict):
        return {k: as_f32(x) for k, x in v.items()}
    return v


def generated(qg_index, params):
    d, cat, _ = qg_index
    rc, req, raw = helpers.generate_request(d, params)
    assert rc == 0, req
    assert as_f32(req) == as_f32(json.loads(json.dumps(qg.search_query(cat, params)))), (params, raw)
    return req


def hits(qg_index, params):
    req = generated(qg_index, params)
    res = qg_index[2].search(req)
    return [fx.TEST_QG_DOCS[h[0]] for h in res["data"]]


def test_catalog_is_what_the_index_says(qg_index):
    _, cat, _ = qg_index
    assert "meanings.eng[]" in cat.search_fields and "tags[]" in cat.search_fields and "commonness" in cat.all_fields


def test_field_expand_and_simplify(qg_index):  # query_parser_to_veloci_request.rs:201-223
    fields = ["Title", "Author[].name"]
    assert qg.debug(qg.expand_fields(("leaf", "Fred", None), fields)) == '(Author[].name:"Fred" OR Title:"Fred")'
    assert qg.debug(qg.expand_fields(("attr", "Title", ("leaf", "Fred", None)), fields)) == 'Title:"Fred"'
    req = generated(qg_index, {"search_term": "urge will", "fields": ["meanings.eng[]", "meanings.ger[]"]})
    paths = [(q["search"]["path"], q["search"]["terms"][0]) for q in req["search_req"]["or"]["queries"]]
    assert paths == [("meanings.ger[]", "will"), ("meanings.eng[]", "will"), ("meanings.ger[]", "urge"), ("meanings.eng[]", "urge")]


def test_simple_search_querygenerator(qg_index):  # test_query_generator.rs:170-179
    h = hits(qg_index, {"search_term": "urge"})
    assert len(h) == 1 and h[0]["ent_seq"] == "1587690" and h[0]["commonness"] == 20 and h[0]["tags"] == ["nice"]


def test_attributed_search(qg_index):  # test_query_generator.rs:182-189, :192-204
    h = hits(qg_index, {"search_term": "ent_seq:99999"})
    assert len(h) == 1 and h[0]["ent_seq"] == "99999"
    h = hits(qg_index, {"search_term": "ent_seq:99999", "parser_options": {"no_attributes": True}})
    assert len(h) == 1 and h[0]["ent_seq"] == "1337"


def test_or_and_connect(qg_index):  # test_query_generator.rs:207-228, :230-254, :296-303
    h = hits(qg_index, {"search_term": "urge OR いよく"})
    assert len(h) == 3 and h[0]["ent_seq"] == "1587690"
    for extra in ({}, {"stopword_lists": []}, {"stopword_lists": ["en"]}):
        h = hits(qg_index, {"search_term": "urge AND いよく", **extra})
        assert len(h) == 1 and h[0]["ent_seq"] == "1587690" and h[0]["tags"] == ["nice"]
    assert hits(qg_index, {"search_term": "urge AND いよく AND awesome"}) == []


def test_complex_from_json(qg_index):  # test_query_generator.rs:270-294
    p = {"search_term": "will", "top": 10, "facets": ["commonness", "kanji[].commonness"], "levenshtein": 0, "boost_fields": {"meanings.eng[]": 1.5}}
    h = hits(qg_index, p)
    assert len(h) == 2 and h[0]["meanings"]["eng"][0] == "will"
    p["boost_terms"] = {"meanings.ger[]:majestätisches Aussehen (n)": 20.0}
    h = hits(qg_index, p)
    assert len(h) == 2 and h[0]["meanings"]["eng"][0] == "will testo"
    req = generated(qg_index, p)
    assert req["facets"] == [{"field": "commonness", "top": 5}, {"field": "kanji[].commonness", "top": 5}]
    assert req["boost_term"] == [{"path": "meanings.ger[]", "terms": ["majestätisches Aussehen (n)"], "boost": 20.0}]


def test_wildcards(qg_index):  # test_query_generator.rs:306-326, the requests of :328-356
    assert len(hits(qg_index, {"search_term": "awes*"})) == 1
    assert len(hits(qg_index, {"search_term": "いよ*"})) == 3
    assert len(hits(qg_index, {"search_term": "awesam*"})) == 1
    req = generated(qg_index, {"search_term": "*wesom*", "fields": ["tags[]"]})
    assert req["search_req"] == {"search": {"path": "tags[]", "terms": [".*wesom.*"], "is_regex": True}}
    req = generated(qg_index, {"search_term": "*we*some", "fields": ["tags[]"]})
    assert req["search_req"]["search"]["terms"] == [".*we.*some"] and "levenshtein_distance" not in req["search_req"]["search"]
    req = generated(qg_index, {"search_term": "tags[]:a.b*c(d)*", "parser_options": {"no_parentheses": True}})
    assert req["search_req"]["search"]["terms"] == ["a\\.b.*c\\(d\\).*"]


def test_errors(qg_index):  # test_query_generator.rs:358-378
    d, cat, _ = qg_index
    rc, msg, _ = helpers.generate_request(d, {"search_term": "awes*", "fields": ["notexistingfield"]})
    assert rc == 2 and "All fields filtered" in msg
    with pytest.raises(qg.GeneratorError) as e:
        qg.search_query(cat, {"search_term": "awes*", "fields": ["notexistingfield"]})
    assert str(e.value) == msg
    rc, msg, _ = helpers.generate_request(d, {"search_term": "notexistingfield:awes*"})
    assert rc == 2 and "Field notexistingfield not found in" in msg
    with pytest.raises(qg.GeneratorError) as e:
        qg.search_query(cat, {"search_term": "notexistingfield:awes*"})
    assert str(e.value) == msg
    assert helpers.generate_request(d, {"search_term": "a AND "})[0] == 1
    assert helpers.generate_request(d, {"search_term": ""})[0] == 1
    assert helpers.generate_request(d, {"search_term": 5})[0] == 5
    assert helpers.generate_request(d, {"search_term": "a", "top": -1})[0] == 5


def test_every_parameter_reaches_the_request(qg_index):
    p = {
        "search_term": 'will urge~2 "long torso" tags[]:(nice cool) X',
        "top": 7, "skip": 2, "ignore_case": False, "levenshtein_auto_limit": 2, "facetlimit": 3, "why_found": True, "text_locality": True,
        "boost_queries": [{"path": "commonness", "boost_fun": "Log10", "param": 1}, {"path": "kanji[].commonness", "expression": "$SCORE * 2", "skip_when_score": [0, 0.25]}],
        "facets": ["tags[]"], "fields": ["meanings.eng[]", "meanings.ger[]", "tags[]"], "boost_fields": {"meanings.eng[]": 2.5, "tags[]": 0.1},
        "boost_terms": {"cool": 3, "tags[]:nice": 1e-7}, "phrase_pairs": True, "explain": True, "filter": "commonness:20 OR ent_seq:(25 26)",
        "operator": "and", "select": "ent_seq", "stopwords": ["x"],
    }
    req = generated(qg_index, p)
    assert req["top"] == 7 and req["skip"] == 2 and req["why_found"] and req["text_locality"] and req["explain"] and req["select"] is None
    assert req["boost"][1] == {"path": "kanji[].commonness", "boost_fun": None, "param": None, "skip_when_score": [0.0, 0.25], "expression": "$SCORE * 2"}
    assert len(req["phrase_boosts"]) == 5 * 3 and req["phrase_boosts"][0]["search1"]["boost"] == 2.5
    assert req["filter"]["or"]["queries"][0] == {"search": {"path": "commonness", "terms": ["20"], "levenshtein_distance": 0}}
    leaves = req["search_req"]["or"]["queries"]
    assert {"search": {"path": "tags[]", "terms": ["nice"], "levenshtein_distance": 1, "boost": 0.1, "ignore_case": False}} in leaves
    assert {"search": {"path": "meanings.ger[]", "terms": ["urge"], "levenshtein_distance": 2, "ignore_case": False}} in leaves
    # "attr:( ... ) more" keeps `more` under the attribute (parser.rs:146-152 returns without looking past the group)
    assert [q["search"]["path"] for q in leaves if q["search"]["terms"] == ["X"]] == ["tags[]"]
    # the text itself: key order and number formats of the serde derive
    _, _, raw = helpers.generate_request(qg_index[0], {"search_term": "urge", "fields": ["meanings.eng[]"], "boost_fields": {"meanings.eng[]": 2}, "boost_terms": {"x": 1e-7, "y": 1.5e10, "z": 0.001}})
    assert raw.startswith('{"search_req":{"search":{"path":"meanings.eng[]","terms":["urge"],"levenshtein_distance":1,"boost":2.0}},"boost_term":[{"path":"commonness","terms":["x"],"boost":1e-7}')
    assert '"boost":15000000000.0' in raw and '"boost":0.001' in raw and raw.endswith('"select":null}')
    # the oracle searches what was generated (every part is one the path supports)
    res = qg_index[2].search(generated(qg_index, {**p, "explain": False, "why_found": False}))
    assert res["num_hits"] >= 1


def test_suggest_query(qg_index):  # src/query_generator.rs:297-322
    d, cat, oracle = qg_index
    for kw in ({}, {"top": 3, "skip": 1}, {"levenshtein": 0, "fields": ["meanings.ger[]"]}, {"levenshtein_auto_limit": 0}):
        for text in ("Begeisteru", "wil", "majes"):
            rc, req, _ = helpers.generate_request(d, {"request": text, **kw}, suggest=True)
            assert rc == 0 and req == qg.suggest_query(cat, text, **kw)
    rc, req, _ = helpers.generate_request(d, {"request": "Begeisteru", "top": 10, "levenshtein": 1, "fields": ["meanings.ger[]"]}, suggest=True)
    assert req["suggest"] == [{"path": "meanings.ger[]", "terms": ["Begeisteru"], "levenshtein_distance": 1, "starts_with": True, "top": 10}]
    out = oracle.call("suggest_multi", request=req)
    assert "begeisterung" in [s[0] for s in out]
