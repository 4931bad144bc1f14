"""The product's host arithmetic on a search part's term hits (csrc/host/part_hits.hpp: the per-part top / skip bound,
the part boost and the token_value boost of get_term_ids_in_field, search_field.rs:292-294,322-331,359-376,391-395) against
the oracle's get_term_ids_in_field, on the CPU: the device only delivers the unbounded (term id, score) list, which the oracle
supplies here."""
import json
import tempfile

import pytest

import helpers

SYNTH = dict(num_docs=30000, vocab=20000, seed=11)


def make_valued_index():
    """A 20k-term synthetic index where a third of the terms that the parts of token_value_parts() match carry token values
    (some below one: negative logarithms)."""
    d = tempfile.mkdtemp(prefix="vb200_tv_")
    helpers.create_synthetic_index(d, **SYNTH)
    words = [json.loads(r)["search_req"]["search"]["terms"][0] for r in helpers.synthetic_requests(num_queries=48, query_kind="single", query_seed=23, **SYNTH)]
    probe = helpers.Oracle(d)
    valued = []
    for i, w in enumerate(words):
        for j, (t, _, _) in enumerate(probe.call("suggest", part={"terms": [w[:2]], "path": "body", "starts_with": True, "top": 600})):
            if (i + j) % 3 == 0:
                valued.append({"text": t, "value": [0.125, 0.25, 3.0, 17.5, 1000.0][(i + j) % 5]})
    valued += [{"text": "not in the dictionary", "value": 7}, {"text": words[0], "value": None}]
    probe.close()
    helpers.add_token_values(d, valued, {"path": "body"})
    return d, helpers.Oracle(d), words, valued


@pytest.fixture(scope="module")
def valued_index(native_libs):
    return make_valued_index()


def token_value_parts(words):
    funs = ["Log10", "Log2", "Multiply", "Add", "Replace", None]
    for i, w in enumerate(words):
        tv = {"path": "body", "param": [1.0, 2.0, 0.5][i % 3]}
        if funs[i % 6]:
            tv["boost_fun"] = funs[i % 6]
        if i % 4 == 1:
            tv["expression"] = ["$SCORE * 2", "10 / $SCORE", "$SCORE - 3.5", "$SCORE + $SCORE"][(i // 4) % 4]
        if i % 5 == 2:
            tv["skip_when_score"] = [10.0, 2.0 / 1.2]
        part = {"terms": [w[:2]], "path": "body", "starts_with": True, "levenshtein_distance": i % 2, "token_value": tv}
        if i % 2 == 0:
            part["top"], part["skip"] = 3 + i % 11, i % 3
        if i % 7 == 0:
            part["boost"] = 2.5
        yield part


def test_token_values_store_is_written_like_the_reference(valued_index):  # token_values_to_tokens.rs:26-82
    d, oracle, words, valued = valued_index
    meta = json.load(open(d + "/metaData.json"))
    entry = [e for e in meta["columns"]["body"]["indices"] if e["path"] == "body.textindex.token_values.boost_valid_to_value"]
    assert len(entry) == 1 and entry[0]["index_category"] == "Boost" and entry[0]["index_cardinality"] == "SingleValue" and not entry[0]["is_empty"]
    import struct
    raw = open(d + "/body.textindex.token_values.boost_valid_to_value", "rb").read()
    assert len(raw) % 4 == 0  # values are f32 bit patterns: 4 bytes wide (single_array.rs:17-28)
    stored = {i: struct.unpack("<f", struct.pack("<I", v - 1))[0] for i, v in enumerate(struct.unpack("<%dI" % (len(raw) // 4), raw)) if v}
    want = {}
    for e in valued:
        if e["value"] is None:
            continue
        hits = oracle.call("field_search", part={"terms": [e["text"]], "path": "body", "levenshtein_distance": 0, "ignore_case": False})["hits_scores"]
        if hits:
            want[hits[0][0]] = float(e["value"])
    assert stored == want and len(want) > 100


def test_bound_and_token_value_against_the_oracle(valued_index):
    d, oracle, words, _ = valued_index
    changed = 0
    for part in token_value_parts(words):
        plain = {k: v for k, v in part.items() if k not in ("top", "skip", "boost", "token_value")}
        unbounded = oracle.call("field_search", part=plain)["hits_scores"]
        got = sorted(helpers.bound_part_hits(d, part, unbounded))
        want = sorted((i, s) for i, s in oracle.call("field_search", part=part)["hits_scores"])
        assert [i for i, _ in got] == [i for i, _ in want], part
        for (_, gs), (_, ws) in zip(got, want):
            assert abs(gs - ws) <= 1e-6 * max(abs(ws), 1e-30), part
        base = dict(oracle.call("field_search", part={k: v for k, v in part.items() if k != "token_value"})["hits_scores"])
        changed += sum(1 for i, s in want if i in base and base[i] != s)
    assert changed > 200, "the token values did not reach the parts' hits"


def test_errors(valued_index):
    d, oracle, words, _ = valued_index
    with pytest.raises(RuntimeError, match="Did not found path in indices"):  # persistence.rs:454-458
        helpers.bound_part_hits(d, {"terms": ["ab"], "path": "body", "token_value": {"path": "commonness"}}, [(1, 1.0)])
    with pytest.raises(RuntimeError):  # expression.rs: not `x op y`
        helpers.bound_part_hits(d, {"terms": ["ab"], "path": "body", "token_value": {"path": "body", "expression": "$SCORE *"}}, [(1, 1.0)])
