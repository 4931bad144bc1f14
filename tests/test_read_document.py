"""`select` (documents rebuilt from the indices, src/search/read_document.rs) and the token-id based why_found that goes
with it (src/search/why_found.rs, highlight_document) -- SURVEY §8 f.4.  The product's host code
(csrc/host/read_document.hpp) against a plain-Python oracle that reads the index files with its own decoder
(oracle/read_document.py over oracle/index_files.py), on the reference's own assertions (tests/all/tests.rs:439-453,
:1050-1066, tests_large.rs:52-70, test_why_found.rs:177-221, :267-284) and on every document of the fixture corpora."""
import ctypes
import json
import os
import sys
import tempfile

import pytest

import helpers
import ref_fixtures as fx

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import read_document as ord_  # noqa: E402  (test infrastructure)


def product(directory, doc_id, fields=None, term_ids=None):
    lib = helpers._index_lib()
    lib.vidx_read_doc.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    out = ctypes.create_string_buffer(1 << 20)
    rc = lib.vidx_read_doc(directory.encode(), doc_id, json.dumps(fields or []).encode(), None if term_ids is None else json.dumps(term_ids).encode(), out, len(out))
    assert rc == 0, out.value
    return json.loads(out.value.decode("utf-8"))


def make(docs, config):
    d = tempfile.mkdtemp(prefix="vb200_sel_")
    helpers.create_index(d, docs, config)
    return d, ord_.Reader(d), helpers.Oracle(d)


@pytest.fixture(scope="module")
def test_all(native_libs):
    return make(fx.TEST_ALL_DOCS, fx.TEST_ALL_CONFIG)


def test_select_fields(test_all):  # tests.rs:439-453
    d, reader, oracle = test_all
    hit = oracle.search({"search_req": {"search": {"terms": ["urge"], "path": "meanings.eng[]"}}})["data"][0][0]
    for doc in (product(d, hit, ["ent_seq", "tags[]"]), reader.read_data(hit, ["ent_seq", "tags[]"])):
        assert doc["ent_seq"] == "1587690" and "commonness" not in doc and doc["tags"] == ["nice"]


def test_select_on_long_text(test_all):  # tests.rs:1050-1066 (a text longer than do_not_store_text_longer_than: rebuilt from its tokens)
    d, reader, oracle = test_all
    hit = oracle.search({"search_req": {"search": {"terms": ["story"], "path": "mylongtext"}}})["data"][0][0]
    for doc in (product(d, hit, ["mylongtext"]), reader.read_data(hit, ["mylongtext"])):
        assert doc["mylongtext"] == fx.LONG


def test_every_document_every_field(test_all):
    """every field of every document of the fixture corpus, alone and all together: the product equals the oracle, and both
    equal the source document's values as text"""
    d, reader, _ = test_all
    fields = sorted(reader.ix.meta["columns"])
    assert "kanji[].text" in fields and "address[].line[]" in fields
    n = 0
    for doc_id, src in enumerate(fx.TEST_ALL_DOCS):
        whole = product(d, doc_id, fields)
        assert whole == reader.read_data(doc_id, fields), doc_id
        for f in fields:
            one = product(d, doc_id, [f])
            assert one == reader.read_data(doc_id, [f]), (doc_id, f)
            n += bool(one)
        if "ent_seq" in src:
            assert whole["ent_seq"] == src["ent_seq"]
        if "tags" in src:
            assert whole["tags"] == src["tags"]
        if "kanji" in src:
            assert [k["text"] for k in whole["kanji"]] == [k["text"] for k in src["kanji"]]
            assert [k.get("commonness") for k in whole["kanji"]] == [None if "commonness" not in k else str(k["commonness"]) for k in src["kanji"]]
        if "meanings" in src:
            assert whole["meanings"].get("eng", []) == src["meanings"].get("eng", [])
    assert n > 60
    # a selected field that is the prefix of another wins; unknown fields are ignored (search.rs:272-279)
    assert product(d, 1, ["nofield", "ent_seq"]) == product(d, 1, ["ent_seq"])
    assert product(d, 1, ["kanji[]", "kanji[].text"]) == reader.read_data(1, ["kanji[]", "kanji[].text"])


def test_select_on_large_text(native_libs):  # tests_large.rs:52-70
    docs = [{"category": "superb", "tags": ["nice", "cool"], "text": "hallo " * 400}] * 3
    d, reader, oracle = make(docs, {"*GLOBAL*": {"features": ["All"]}, "tags[]": {"facet": True}})
    hit = oracle.search({"search_req": {"search": {"terms": ["superb"], "path": "category"}}})["data"][0][0]
    for doc in (product(d, hit, ["text"]), reader.read_data(hit, ["text"])):
        assert doc == {"text": "hallo " * 400}


def test_why_found_with_select(native_libs):  # test_why_found.rs:177-191, :206-221, :267-284
    d, reader, oracle = make(fx.TEST_WHYFOUND_DOCS, fx.TEST_WHYFOUND_CONFIG)

    def why(part):
        res = oracle.call("field_search", part=part)
        term_ids = {part["path"] + ".textindex": [h[0] for h in res["hits_scores"]]}
        anchor = oracle.search({"search_req": {"search": part}})["data"][0][0]
        ref = reader.get_why_found(anchor, term_ids)
        assert product(d, anchor, term_ids=term_ids) == ref, part
        return anchor, ref

    assert why({"terms": ["ID1000"], "path": "not_tokenized"})[1]["not_tokenized"] == ["<b>ID1000</b>"]
    assert why({"terms": ["ID1000"], "path": "not_tokenized_1_n[]"})[1]["not_tokenized_1_n[]"] == ["<b>ID1000</b>"]
    anchor, ref = why({"terms": ["umsortiert"], "path": "viele[]", "levenshtein_distance": 0})
    assert ref["viele[]"] == [" ... zu checken, dass da nicht <b>umsortiert</b> wird"]
    assert product(d, anchor, ["richtig"]) == reader.read_data(anchor, ["richtig"]) == {"richtig": "shön"}
    assert why({"terms": ["treffers"], "path": "viele[]", "levenshtein_distance": 1})[1]["viele[]"] == ["<b>treffers</b>", "super <b>treffers</b>"]
    assert why({"terms": ["schön"], "path": "richtig", "levenshtein_distance": 1})[1]["richtig"] == ["<b>schön</b> super"]
