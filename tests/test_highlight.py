"""why_found on the stored document (second half; src/highlight_field.rs:98-186, src/search.rs:65-101).

The reference's why_found tests (tests/all/test_why_found.rs) restated: the request's parts are matched by the CPU search
oracle, the matched term texts highlight the stored documents of the hits.  Two implementations are held to the reference's
expected strings and to each other: the plain-Python oracle (oracle/highlight.py) and the product's host code
(csrc/host/highlight.hpp through the host-only helper library).  The GPU path (vgpu_batch_result_docs) is checked in
tests/test_gpu_round2.py."""
import ctypes
import json
import os
import random
import sys
import tempfile

import pytest

import helpers
import ref_fixtures as fx

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import highlight as ohl  # noqa: E402  (test infrastructure)


@pytest.fixture(scope="module")
def corpus(native_libs):
    d = tempfile.mkdtemp(prefix="vb200_wf_")
    helpers.create_index(d, fx.TEST_WHYFOUND_DOCS, fx.TEST_WHYFOUND_CONFIG)
    columns = json.load(open(os.path.join(d, "metaData.json")))["columns"]
    return d, columns, helpers.Oracle(d)


def product_highlight(directory, doc, terms):
    lib = helpers._index_lib()
    lib.vidx_highlight_doc.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    out = ctypes.create_string_buffer(1 << 18)
    rc = lib.vidx_highlight_doc(directory.encode(), json.dumps(doc, ensure_ascii=False).encode("utf-8"), json.dumps({k: sorted(v) for k, v in terms.items()}, ensure_ascii=False).encode("utf-8"), out, len(out))
    assert rc == 0, out.value
    return json.loads(out.value.decode("utf-8"))


def leaves(req):
    if "search" in req:
        return [req["search"]]
    return [p for q in (req.get("or") or req.get("and"))["queries"] for p in leaves(q)]


def why_found(corpus, request):
    """hits of the request with their why_found maps: [(doc, why_found)] -- what search_testo_to_doc! gives the tests"""
    d, columns, oracle = corpus
    terms = {}
    for part in leaves(request["search_req"]):  # term_text_in_field of every part (search_field.rs:386-389), merged (set_op.rs:49-64)
        res = oracle.call("field_search", part=part)
        terms.setdefault(part["path"] + ".textindex", set()).update(res["terms"])
    out = []
    for hit in oracle.search(request)["data"]:
        doc = fx.TEST_WHYFOUND_DOCS[hit[0]]
        ref = ohl.highlight_on_original_document(columns, doc, terms)
        assert product_highlight(d, doc, terms) == ref, (request, doc)
        out.append((doc, ref))
    return out


def S(term, path, **kw):
    return {"search_req": {"search": {"terms": [term], "path": path, **kw}}, "why_found": True}


def test_should_tokenize_url(corpus):  # test_why_found.rs:73-92
    hits = why_found(corpus, S("veloci", "url"))
    assert len(hits) == 1 and hits[0][1]["url"] == ["https://github.com/PSeitz/<b>veloci</b>"]


def test_custom_tokenized(corpus):  # test_why_found.rs:94-142
    assert why_found(corpus, S("test", "custom_tokenized"))[0][1]["custom_tokenized"] == ["<b>test</b>§_ cool _"]
    assert why_found(corpus, S("§", "custom_tokenized"))[0][1]["custom_tokenized"] == ["test<b>§</b>_ cool _"]
    assert why_found(corpus, S("_ cool _", "custom_tokenized"))[0][1]["custom_tokenized"] == ["test§<b>_ cool _</b>"]
    assert why_found(corpus, S("<<", "custom_tokenized"))[0][1]["custom_tokenized"] == ["<b><<</b>cool>>"]
    assert why_found(corpus, S("cool", "custom_tokenized")) == []  # :223-235: a space is not a separator there


def test_complete_text_hits(corpus):  # test_why_found.rs:149-221 (the `select` variants are outside this path)
    assert why_found(corpus, S("<<cool>>", "custom_tokenized"))[0][1]["custom_tokenized"] == ["<b><<cool>></b>"]
    assert why_found(corpus, S("ID1000", "not_tokenized"))[0][1]["not_tokenized"] == ["<b>ID1000</b>"]
    assert why_found(corpus, S("ID1000", "not_tokenized_1_n[]"))[0][1]["not_tokenized_1_n[]"] == ["<b>ID1000</b>"]


def test_tokens_and_text_ids(corpus):  # test_why_found.rs:237-266
    hits = why_found(corpus, S("schön", "richtig", levenshtein_distance=1))
    assert hits[0][1]["richtig"] == ["<b>schön</b> super"] and hits[1][1]["richtig"] == ["<b>shön</b>"]
    hits = why_found(corpus, S("treffers", "viele[]", levenshtein_distance=1))
    assert hits[0][1]["viele[]"] == ["<b>treffers</b>", "super <b>treffers</b>"]


def test_window_and_ellipsis(corpus):  # test_why_found.rs:268-300 (second request), :302-316
    hits = why_found(corpus, S("umsortiert", "viele[]", levenshtein_distance=0))
    assert hits[0][0]["richtig"] == "shön"
    assert hits[0][1]["viele[]"] == [" ... zu checken, dass da nicht <b>umsortiert</b> wird"]
    assert why_found(corpus, S("Taschenbuch", "buch", levenshtein_distance=1))[0][1]["buch"] == ["<b>Taschenbuch</b> (kartoniert)"]


def test_multi_terms(corpus):  # test_why_found.rs:318-350
    req = {"search_req": {"or": {"queries": [{"search": {"terms": ["Taschenbuch"], "path": "buch", "levenshtein_distance": 1}},
                                             {"search": {"terms": ["kartoniert"], "path": "buch", "levenshtein_distance": 1}}]}}, "why_found": True}
    assert why_found(corpus, req)[0][1]["buch"] == ["<b>Taschenbuch</b> (<b>kartoniert</b>)"]


def test_random_texts_product_equals_oracle(corpus):
    d, columns, _ = corpus
    rng = random.Random(21)
    words = ["alpha", "beta", "gamma", "delta", "schön", "食べる", "x", "und", "so"]
    seps = [" ", ", ", " - ", ".", "  ", ":", "…"]
    n_some = 0
    for _ in range(600):
        text = "".join(rng.choice(words) + rng.choice(seps) for _ in range(rng.randint(1, 60)))
        if rng.random() < 0.3:
            text = rng.choice(seps) + text
        if rng.random() < 0.3:
            text = text.rstrip(" ,.-:…")
        terms = set(rng.sample(words, rng.randint(1, 3)))
        if rng.random() < 0.2:
            terms.add(rng.choice(seps))
        if rng.random() < 0.1:
            terms = {text}
        doc = {"richtig": text, "viele": [text, "nichts"], "not_tokenized": text}
        sets = {"richtig.textindex": terms, "viele[].textindex": terms, "not_tokenized.textindex": terms}
        ref = ohl.highlight_on_original_document(columns, doc, sets)
        assert product_highlight(d, doc, sets) == ref, (text, terms)
        n_some += bool(ref)
    assert n_some > 400


def test_regex_why_found(corpus):  # test_why_found.rs:345-360: the pattern spans the tokens, only the whole text matches: all of it is marked
    hits = why_found(corpus, S(".*github.com.*", "url", is_regex=True))
    assert len(hits) == 1 and hits[0][1]["url"] == ["<b>https://github.com/PSeitz/veloci</b>"]


def test_regex_why_found_token(corpus):  # test_why_found.rs:362-377: a token and the whole text match: the token is what gets marked
    hits = why_found(corpus, S(".*PSeitz.*", "url", is_regex=True))
    assert len(hits) == 1 and hits[0][1]["url"] == ["https://github.com/<b>PSeitz</b>/veloci"]


def test_highlight_text_unit_vectors(corpus):  # highlight_field.rs:279-300 (the default tokenizer)
    d, _, _ = corpus
    assert ohl.highlight_text("mein treffer", {"treffer"}, ohl.DEFAULT_SEPERATORS) == "mein <b>treffer</b>"
    assert ohl.highlight_text("mein treffer treffers", {"treffers", "treffer"}, ohl.DEFAULT_SEPERATORS) == "mein <b>treffer</b> <b>treffers</b>"
    # the product through a document of one tokenized field
    for text, terms, want in (("mein treffer", {"treffer"}, "mein <b>treffer</b>"), ("mein treffer treffers", {"treffers", "treffer"}, "mein <b>treffer</b> <b>treffers</b>")):
        assert product_highlight(d, {"richtig": text}, {"richtig.textindex": terms}) == {"richtig": [want]}
