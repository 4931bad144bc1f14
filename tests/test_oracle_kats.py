"""Known-answer vectors of the reference's inline unit tests, run against the
oracle's restatements and the shared on-disk codecs (file:line cited per test)."""
import math
import struct

import pytest

import helpers


@pytest.fixture(scope="module")
def o(native_libs):
    return helpers.Oracle()


def f32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


def test_union_hits_ids(o):  # set_op.rs:260-276
    r = o.call("union_hits_ids", inputs=[{"hits_ids": [10, 0, 5]}, {"hits_ids": [0, 3, 10, 20]}])
    assert r == [0, 3, 5, 10, 20]


def test_intersect_score_hits_with_ids(o):  # set_op.rs:328-345
    r = o.call("intersect_score_hits_with_ids", inputs=[{"hits_scores": [[10, 20.0], [0, 20.0], [5, 20.0]]}, {"hits_ids": [0, 10]}])
    assert r == [[0, 20.0], [10, 20.0]]


def test_intersect_hits_ids(o):  # set_op.rs:511-530
    assert o.call("intersect_hits_ids", inputs=[{"hits_ids": [10, 0, 5]}, {"hits_ids": [0, 3, 10, 20]}]) == [0, 10]


def test_intersect_hits_scores(o):  # set_op.rs:532-551
    r = o.call("intersect_hits_score", inputs=[{"hits_scores": [[10, 20.0], [0, 20.0], [5, 20.0]]}, {"hits_scores": [[0, 20.0], [3, 20.0], [10, 30.0], [20, 30.0]]}])
    assert r == [[0, 40.0], [10, 50.0]]


def test_intersect_hits_scores_regression(o):  # set_op.rs:553-580
    h1 = [[704, 13.7], [19921, 39.4], [20000, 13.7], [44650, 39.4]]
    h2 = [[18779, 28.199999], [20000, 14.400001], [32606, 39.4], [130721, 13.3], [168854, 2.0666666]]
    r = o.call("intersect_hits_score", inputs=[{"hits_scores": h1}, {"hits_scores": h2}])
    assert len(r) == 1 and r[0][0] == 20000


def test_union_hits_score_max_per_term_times_n_squared(o):  # set_op.rs:166-186 (+ the commented vector :278-309, "max_score" line)
    h1 = [[10, 20.0], [0, 10.0], [5, 20.0]]
    h2 = [[0, 20.0], [3, 20.0], [10, 30.0], [20, 30.0]]
    r = o.call("union_hits_score", inputs=[{"term": "a", "hits_scores": h1}, {"term": "b", "hits_scores": h2}])
    assert r == [[0, 120.0], [3, 20.0], [5, 20.0], [10, 200.0], [20, 30.0]]
    # same term text in both inputs: one slot, max, n = 1
    r = o.call("union_hits_score", inputs=[{"term": "a", "hits_scores": h1}, {"term": "a", "hits_scores": h2}])
    assert r == [[0, 20.0], [3, 20.0], [5, 20.0], [10, 30.0], [20, 30.0]]
    # a single input passes through untouched and unsorted (:90-96)
    assert o.call("union_hits_score", inputs=[{"term": "a", "hits_scores": h1}]) == h1


def test_apply_boost_values_anchor(o):  # boost.rs:239-253
    r = o.call("apply_boost_values_anchor", hits_scores=[[1, 10.0], [3, 20.0], [5, 20.0]], boost={"path": "x", "boost_fun": "Multiply"},
               boost_ids=[[1, 2.0], [2, 20.0], [5, 3.0], [6, 3.0]])
    assert r == [[1, 20.0], [3, 20.0], [5, 60.0]]


def test_boost_hits_ids_vec_multi(o):  # boost.rs:404-430
    r = o.call("boost_hits_ids_vec_multi", hits_scores=[[10, 20.0], [0, 20.0], [5, 20.0], [60, 20.0]],
               boosts=[{"hits_ids": [0, 3, 10, 10, 70]}, {"hits_ids": [10, 60]}])
    assert r == [[0, 40.0], [5, 20.0], [10, 160.0], [60, 40.0]]


def test_distance(o):  # search_field.rs:734-744
    for a, b, d in [("a", "a", 0), ("a", "b", 1), ("", "a", 1), ("a", "", 1), ("aa", "a", 1), ("a", "aa", 1), ("a", "bbb", 3), ("bbb", "a", 3)]:
        assert o.call("distance", a=a, b=b) == d


def test_distance_dfa_transposition_and_fallback(o):  # search_field.rs:298-300,691-702
    assert o.call("distance_dfa", hit="saucissonsec", term="saucisson sec", d=2) == 1  # the commented dfa test :746-759
    assert o.call("distance_dfa", hit="ab", term="ba", d=1) == 1  # transposition costs one inside the DFA
    assert o.call("distance_dfa", hit="ab", term="ba", d=0) == 2  # beyond d: plain Levenshtein fallback
    assert o.call("distance_dfa", hit="awesome sauce", term="awe", d=1) == 10


def test_default_score_kats(o):  # search_field.rs:27-33, values derived in SURVEY.md section 8(a5)
    assert o.call("default_score", distance=0, prefix=False) == 10.0
    assert o.call("default_score", distance=0, prefix=True) == 10.0
    assert f32(o.call("default_score", distance=1, prefix=False)) == f32(2.0 / f32(1.2))
    assert f32(o.call("default_score", distance=1, prefix=True)) == f32(2.0 / f32(1.2))
    assert abs(o.call("default_score", distance=2, prefix=False) - 0.9090909) < 1e-6
    assert abs(o.call("default_score", distance=2, prefix=True) - 1.1204717) < 1e-6
    assert o.call("default_score", distance=3, prefix=False) == 0.625
    assert abs(o.call("default_score", distance=3, prefix=True) - 0.9090909) < 1e-6


def test_expression(o):  # expression.rs:108-123
    assert o.call("expression", expr="$SCORE + 2.0", value=10.0) == 12.0
    assert o.call("expression", expr="10.0 / $SCORE", value=10.0) == 1.0
    assert o.call("expression", expr="$SCORE * $SCORE", value=10.0) == 100.0


def test_token_score_kats(o):  # calculate_score.rs:34-49, SURVEY.md section 8(c)
    kats = [((0, 1, 1, True), 395), ((0, 1, 1, False), 148), ((0, 1, 3, False), 145), ((2, 1, 3, False), 142), ((0, 300, 1, True), 382), ((5, 100000, 8, False), 83)]
    for (pos, nocc, ntok, exact), want in kats:
        assert o.call("token_score", pos=pos, nocc=nocc, ntok=ntok, exact=exact) == want


def test_top_n_sort_order(o):  # sort.rs:5-22 + search.rs:123-130: score desc, then id desc
    hits = [[i, float(i % 7)] for i in range(1000)]
    r = o.call("top_n_sort", hits=hits, top=5)
    want = sorted(hits, key=lambda h: (-h[1], -h[0]))
    assert r[:5] == want[:5]
    assert len(r) >= 5


def test_f16_roundtrip(o):  # persistence_score/mod.rs:7-17, half::f16 semantics
    assert o.call("f16_roundtrip", value=395.0) == 395.0
    assert o.call("f16_roundtrip", value=2049.0) == 2048.0  # ties to even
    assert o.call("f16_roundtrip", value=2051.0) == 2052.0
    assert o.call("f16_roundtrip", value=65504.0) == 65504.0  # largest finite half
    assert o.call("f16_roundtrip", value=1.0) == 1.0


def test_steps_to_anchor(o):  # util.rs:173-188
    assert o.call("steps_to_anchor", path="meanings.ger[]") == ["meanings.ger[]", "meanings.ger[].textindex"]
    assert o.call("steps_to_anchor", path="address[].line[]") == ["address[]", "address[].line[]", "address[].line[].textindex"]
    assert o.call("steps_to_anchor", path="title") == ["title.textindex"]


def test_tokenizer(o):  # tokenizer/mod.rs:39-77
    assert o.call("tokenize", text="das \n ist ein txt, test") == ["das", " \n ", "ist", " ", "ein", " ", "txt", ", ", "test"]
    assert o.call("tokenize", text=" Taschenbuch (kartoniert)") == [" ", "Taschenbuch", " (", "kartoniert", ")"]
    assert o.call("tokenize", text="T oll") == ["T", " ", "oll"]


# ------------------------------------------------------------------ codecs --
def test_indirect_codec(o):  # indirect/mod.rs:28-72
    adds = [[0, [5, 6]], [1, [9]], [2, [9]], [3, [9, 50000]], [5, [80]], [9, [0]], [10, [0]]]
    r = o.call("codec_indirect", adds=adds, queries=[0, 1, 2, 3, 4, 5, 6, 9, 10, 11])
    assert r["values"] == [[5, 6], [9], [9], [9, 50000], None, [80], None, [0], [0], None]
    r = o.call("codec_indirect", adds=adds, queries=[0, 1, 2, 3, 4, 5])  # count_values_for_ids
    assert r["counts"]["5"] == 1 and r["counts"]["9"] == 3


def test_packed_codec(o):  # single_array.rs:65-91
    r = o.call("codec_packed", stored=[123, 33, 545, 99], queries=[0, 1, 2, 3, 4, 5])
    assert r["values"] == [122, 32, 544, 98, None, None] and r["width"] == 2
    r = o.call("codec_packed", stored=[50001, 33], queries=[0, 1, 2])
    assert r["values"] == [50000, 32, None] and r["width"] == 3


def test_phrase_pair_codec(o):  # persistence_data_binary_search.rs:213-263
    adds = [[0, 0, [5, 6]], [0, 1, [9]], [2, 0, [9]], [2, 3, [9, 50000]], [5, 0, [80]], [5, 9, [0]], [5, 10, [0]]]
    qs = [[0, 0], [0, 1], [0, 2], [2, 0], [2, 3], [5, 0], [5, 9], [5, 10]]
    r = o.call("codec_phrase", adds=adds, queries=qs)
    assert r["size"] == 7
    assert r["offsets"][:2] == [1, 4]  # decode_pos(0) == ((0,0),1), decode_pos(1) == ((0,1),4): vint(len) + 2 vints
    assert r["values"] == [[5, 6], [9], None, [9], [9, 50000], [80], [0], [0]]


def test_anchor_score_codec(o):  # token_to_anchor_score_vint.rs:212-232
    assert o.call("codec_anchor_score", adds=[[1, [1, 1]]], queries=[0, 1, 2]) == [[], [[1, 1]], []]
    r = o.call("codec_anchor_score", adds=[[5, [1, 1, 2, 3]]], queries=[4, 5] + list(range(6, 18)))
    assert r[0] == [] and r[1] == [[1, 1], [2, 3]] and all(x == [] for x in r[2:])
    big = [[0, [0, 395, 7, 148, 100000, 148, 100001, 70000]]]
    assert o.call("codec_anchor_score", adds=big, queries=[0]) == [[[0, 395], [7, 148], [100000, 148], [100001, 70000]]]
    # the same store with 8-byte start positions (data_type U64, token_to_anchor_score_vint.rs:242-248)
    r = o.call("codec_anchor_score", adds=[[1, [1, 1]], [5, [1, 1, 2, 3]]], queries=[0, 1, 4, 5, 6], wide=True)
    assert r == [[], [[1, 1]], [], [[1, 1], [2, 3]], []]


def test_fst_roundtrip(o):  # search_field.rs:36-51,101-141 (fst::Map::from_iter + ord_to_term)
    keys = sorted(["awesome", "awe", "a", "b", "bar", "baz", "majestät", "majestätischer", "zebra", "偉容", ""], key=lambda s: s.encode())
    r = o.call("fst_roundtrip", keys=keys, ords=list(range(len(keys))) + [len(keys)])
    assert r["len"] == len(keys)
    assert r["items"] == [[k, i] for i, k in enumerate(keys)]
    assert r["ord_to_term"] == keys + [None]
    # ids with gaps (long tokens are left out of the FST, create_fulltext.rs:60-64)
    r = o.call("fst_roundtrip", keys=["a", "b", "c"], values=[0, 5, 9], ords=[0, 5, 9])
    assert r["items"] == [["a", 0], ["b", 5], ["c", 9]]
    assert r["ord_to_term"] == ["a", "b", "c"]
    many = sorted({"w%05d" % (i * 7919 % 30011) for i in range(3000)})
    r = o.call("fst_roundtrip", keys=many)
    assert [k for k, _ in r["items"]] == many and r["bytes"] < sum(len(k) for k in many)
