"""Regex search parts (`is_regex`, src/search/search_field.rs:72-83).

The reference's matcher is the dense DFA of regex-automata 0.1.9 (unanchored, leftmost-first) run over the term dictionary:
a term matches when the DFA sits in a match state after the term's last byte.  Two independent restatements are held
against each other here: the oracle's thread-list simulation (oracle/regex_sim.hpp) and the product's scalar-class DFA
(csrc/host/regex_dfa.hpp, the tables regex_match_kernel runs on the GPU), on the reference's own vectors, on hand cases of
the leftmost-first cut, and on random patterns x random terms.  The GPU kernel itself is checked in tests/test_gpu_round2.py.
"""
import json
import random
import tempfile

import pytest

import helpers
import ref_fixtures as fx


@pytest.fixture(scope="module")
def oracle(native_libs):
    return helpers.Oracle()


def both(oracle, pattern, terms, ci=True, starts_with=False):
    rc, got, stats = helpers.regex_match(pattern, terms, ci, starts_with)
    ref = oracle.call("regex_accepts", pattern=pattern, case_insensitive=ci, starts_with=starts_with, terms=terms)
    if ref in ("bad", "outside"):
        assert rc == (1 if ref == "bad" else 8), (pattern, rc, got, ref)
        return None
    assert rc == 0, (pattern, got)
    assert got == [c == "1" for c in ref], (pattern, ci, starts_with, [t for t, g, r in zip(terms, got, ref) if g != (r == "1")])
    return got


def test_reference_vectors(oracle):
    # search_field.rs:101-120: ".*wesom.*" finds "awesome"
    assert both(oracle, ".*wesom.*", ["awesome"]) == [True]
    # search_field.rs:121-141: ".*wesom" finds it only as a prefix match (starts_with)
    assert both(oracle, ".*wesom", ["awesome"], starts_with=True) == [True]
    assert both(oracle, ".*wesom", ["awesome"]) == [False]
    # tests/all/test_query_generator.rs:328-356: "*wesom*" -> ".*wesom.*", "*we*some" -> ".*we.*some", "*wesam*" has no hit
    tags = ["nice", "cool", "awesome", "Eis", "coolo", "ent_seq:99999"]
    assert both(oracle, ".*wesom.*", tags) == [False, False, True, False, False, False]
    assert both(oracle, ".*we.*some", tags) == [False, False, True, False, False, False]
    assert both(oracle, ".*wesam.*", tags) == [False] * 6


def test_unanchored_and_leftmost_first(oracle):
    # unanchored: the pattern may start anywhere, but the match state must hold at the END of the term
    assert both(oracle, "some", ["awesome", "some", "somewhat", "awesom"]) == [True, True, False, False]
    assert both(oracle, "som", ["awesome"], starts_with=True) == [True]
    # leftmost-first: once `a` has matched, the lower-priority `ab` branch (and the restart thread) are gone
    assert both(oracle, "a|ab", ["a", "ab", "b", "xa"]) == [True, False, False, True]
    assert both(oracle, "ab|a", ["a", "ab", "xab"]) == [True, True, True]
    # lazy vs greedy: "a+?" is satisfied by the first a, a second one finds no thread left
    assert both(oracle, "a+?", ["a", "aa", "ba"]) == [True, False, True]
    assert both(oracle, "a+", ["a", "aa", "ba", "ab"]) == [True, True, True, False]
    assert both(oracle, "a??b", ["b", "ab", "aab"]) == [True, True, True]
    # the empty pattern matches before the first scalar; every later state is dead
    assert both(oracle, "", ["", "a"]) == [True, False]
    assert both(oracle, "", ["", "a"], starts_with=True) == [True, True]


def test_case_folding_and_classes(oracle):
    assert both(oracle, "STRASSE", ["strasse", "Strasse", "straße"]) == [True, True, False]
    assert both(oracle, "k", ["k", "K", "K"]) == [True, True, True]            # Kelvin sign folds to k
    assert both(oracle, "k", ["k", "K", "K"], ci=False) == [True, False, False]
    assert both(oracle, "σ", ["σ", "Σ", "ς"]) == [True, True, True]
    assert both(oracle, "[a-c]x", ["ax", "Bx", "dx"]) == [True, True, False]
    assert both(oracle, "[^a-c]x", ["ax", "Bx", "dx", "Dx"]) == [False, False, True, True]
    assert both(oracle, "(?-i)[a-c]x", ["ax", "Bx"]) == [True, False]
    assert both(oracle, r"\d+", ["2024", "x7", "٣", "7x", ""]) == [True, True, True, False, False]
    assert both(oracle, r"\D", ["a", "1"]) == [True, False]
    assert both(oracle, r"a\s\S", ["a b", "a  ", "a　b"]) == [True, False, True]
    assert both(oracle, r"[\d.]+x", ["1.5x", "x"]) == [True, False]
    assert both(oracle, r"a\.b\*", ["a.b*", "axb*"]) == [True, False]
    assert both(oracle, ".", ["a", "\n", "食"]) == [True, False, True]
    assert both(oracle, "(?s).", ["\n"]) == [True]
    assert both(oracle, r"\x41\u{98DF}", ["a食", "A食"]) == [True, True]


def test_counted_repetition_and_groups(oracle):
    assert both(oracle, "a{2}", ["a", "aa", "aaa", "baa"]) == [False, True, False, True]  # "aaa": the match after "aa" cut the restart thread
    assert both(oracle, "x(ab){1,2}", ["xab", "xabab", "xababab"]) == [True, True, False]
    assert both(oracle, "x(ab){2,}y", ["xaby", "xababy", "xabababy"]) == [False, True, True]
    assert both(oracle, "x(?:a|b)*?y", ["xy", "xabay", "xcy"]) == [True, True, False]
    assert both(oracle, "(?P<w>will)|urge", ["will", "urge", "wil"]) == [True, True, False]
    assert both(oracle, "(a|b)*c{0,1}", ["abc", "ab", "abcc"]) == [True, True, False]
    assert both(oracle, "a{0}", ["", "a"]) == [True, False]


def test_patterns_the_reference_cannot_build(oracle):
    for pattern in ("^abc", "abc$", r"\bword", "(abc", "abc)", "a**b)", "[a-", "a{2,1}", r"\q", "*a", "a{x}", "[z-a]"):
        assert both(oracle, pattern, ["abc"]) is None
        assert helpers.regex_match(pattern, ["abc"])[0] == 1, pattern
    for pattern in (r"\w+", r"\p{L}", "[[:alpha:]]", "[a&&b]", "(?m)a", "a{2000}"):
        assert helpers.regex_match(pattern, ["abc"])[0] == 8, pattern
        assert both(oracle, pattern, ["abc"]) is None


def test_random_patterns_product_equals_oracle(oracle):
    rng = random.Random(11)
    atoms = ["a", "b", "c", ".", "[ab]", "[^a]", "ß", "K", r"\d", "食", "(?:ab|a)", "(a|bc)", "(b|)", "x"]
    quants = ["", "", "", "*", "+", "?", "*?", "+?", "??", "{2}", "{1,2}", "{0,2}?", "{2,}"]
    alphabet = "abcxK1ß食\n"
    terms = sorted({"".join(rng.choice(alphabet) for _ in range(rng.randint(0, 7))) for _ in range(300)} | {"", "a", "ab", "abc", "aab", "kk"})
    n = 0
    for _ in range(700):
        pattern = "".join(rng.choice(atoms) + rng.choice(quants) for _ in range(rng.randint(1, 4)))
        if rng.random() < 0.3:
            pattern = pattern + "|" + "".join(rng.choice(atoms) + rng.choice(quants) for _ in range(rng.randint(1, 3)))
        for ci in (True, False):
            for sw in (False, True):
                got = both(oracle, pattern, terms, ci, sw)
                assert got is not None, pattern
                n += sum(got)
    assert n > 10000  # (the comparison saw plenty of accepted terms, not only rejections)


def test_dfa_stays_small_for_generated_wildcards(oracle):
    # what query_generator emits for "*foo*bar*" style terms: ".*" between escaped pieces
    rc, got, (states, classes) = helpers.regex_match(r".*foo.*bar\.baz.*", ["xfooybar.bazz", "foobar"], True, False)
    assert rc == 0 and got == [True, False] and states < 40 and classes < 12


# ---- end to end on the CPU oracle: the reference's wildcard query-generator tests (tests/all/test_query_generator.rs:328-356)
def test_wildcard_requests_through_the_oracle(native_libs):
    d = tempfile.mkdtemp(prefix="vb200_rx_")
    helpers.create_index(d, fx.TEST_QG_DOCS, fx.TEST_QG_CONFIG)
    o = helpers.Oracle(d)
    for term, n_hits in (("*wesom*", 1), ("*we*some", 1)):
        rc, req, _ = helpers.generate_request(d, {"search_term": term, "fields": ["tags[]"]})
        assert rc == 0 and req["search_req"]["search"]["is_regex"] is True
        assert len(o.search(req)["data"]) == n_hits, term
    rc, req, _ = helpers.generate_request(d, {"search_term": "tags[]:*wesam*"})
    assert len(o.search(req)["data"]) == 0
    res = o.call("field_search", part={"path": "tags[]", "terms": [".*oo.*"], "is_regex": True})
    assert len(res["hits_scores"]) == 2  # cool, coolo
