"""search_field::highlight (src/search/search_field.rs:232-245; the reference's tests.rs:1009-1045): the Python oracle
(oracle/field_highlight.py over the oracle's own index decoder, term hits from the C++ oracle) pinned on the reference's two
tests, and the host half of the product (csrc/host/field_highlight.hpp through the host-only helper library) against it.
On the GPU the term hits come from the device match: tests/test_gpu_round2.py::test_field_highlight_on_the_gpu."""
import os
import random
import sys
import tempfile

import pytest

import helpers
import ref_fixtures as fx

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import field_highlight as ofh  # noqa: E402  (test infrastructure)
import read_document as ord_  # noqa: E402

STORY = "Prolog:\nthis is a <b>story</b> of a guy who went ... "

PARTS = [
    {"terms": ["story"], "path": "mylongtext", "levenshtein_distance": 0, "starts_with": True, "snippet": True, "top": 10, "skip": 0},   # tests.rs:1009-1027
    {"terms": ["story"], "path": "tags[]", "levenshtein_distance": 0, "starts_with": True, "snippet": True, "top": 10, "skip": 0},       # tests.rs:1029-1045
    {"terms": ["Story."], "path": "mylongtext", "snippet": True},                                     # normalised to "story" (search_field.rs:234)
    {"terms": ["guy"], "path": "mylongtext", "snippet": True, "snippet_info": {"num_words_around_snippet": 2, "snippet_start_tag": "[", "snippet_end_tag": "]", "snippet_connector": "~"}},
    {"terms": ["the"], "path": "mylongtext", "snippet": True, "levenshtein_distance": 1},            # several tokens of one text hit
    {"terms": ["majestät"], "path": "meanings.ger[]", "snippet": True, "levenshtein_distance": 1},   # texts of several documents, ordered by score
    {"terms": ["will"], "path": "meanings.ger[]", "snippet": True, "starts_with": True, "top": 2, "skip": 1},
    {"terms": ["majestät"], "path": "meanings.ger[]", "snippet": True, "boost": 2.0, "top": 3},
    {"terms": ["nothing here"], "path": "meanings.ger[]", "snippet": True},
]


def make_index():
    d = tempfile.mkdtemp(prefix="vb200_fh_")
    helpers.create_index(d, fx.TEST_ALL_DOCS, fx.TEST_ALL_CONFIG)
    return d


def oracle_highlight(d, oracle, part):
    reader = ord_.Reader(d)
    return ofh.highlight(reader, lambda p: oracle.call("field_search", part=p)["hits_scores"], part)


def bare(part):
    return {k: v for k, v in part.items() if k not in ("top", "skip", "boost", "token_value", "snippet", "snippet_info")}


def same(got, want):
    return [(t, i) for t, _, i in got] == [(t, i) for t, _, i in want] and all(abs(g[1] - w[1]) <= 1e-6 * max(abs(w[1]), 1e-30) for g, w in zip(got, want))


@pytest.fixture(scope="module")
def index(native_libs):
    d = make_index()
    return d, helpers.Oracle(d)


def test_reference_highlight_tests(index):
    d, o = index
    assert [t for t, _, _ in oracle_highlight(d, o, PARTS[0])] == [STORY]
    assert [t for t, _, _ in oracle_highlight(d, o, PARTS[1])] == [STORY]


def test_normalize_text():  # util.rs:11-30
    rng = random.Random(5)
    alphabet = list("ab (f)(m)(n)(1)(x)){}'\"“ \t\n  ,.…;・’-ÄÖİΣσ　 x9") + ["(", ")", "  "]
    texts = ["Majestät (f)", "  der (1) (große)   Wurf…  ", "it's {a} \"test\"", "a\t\tb", "a\tb", "x-y,z.", "ＡＢ　　Ｃ", "ΑΣ ΟΔΟΣ", ""]
    texts += ["".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 14))) for _ in range(400)]
    for t in texts:
        assert helpers.normalize_text(t) == ofh.normalize_text(t), repr(t)


def test_host_half_against_the_oracle(index):
    d, o = index
    n = 0
    for part in PARTS:
        normalized = dict(part, terms=[ofh.normalize_text(t) for t in part["terms"]])
        raw = o.call("field_search", part=bare(normalized))["hits_scores"]
        got = helpers.field_highlight(d, part, raw)
        want = oracle_highlight(d, o, part)
        assert same(got, want), (part, got, want)
        n += len(want)
    assert n >= 6
    for part in ({"terms": ["story"], "path": "mylongtext"},                       # no snippet: nothing to return for a hit
                 {"terms": ["1587690"], "path": "nofulltext", "snippet": True}):    # not tokenized
        with pytest.raises(helpers.OracleError) as e:
            helpers.field_highlight(d, part, [(1, 1.0)])
        assert e.value.status == 2
