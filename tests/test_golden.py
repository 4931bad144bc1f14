"""Committed golden results (tests/golden/*.json, written by tests/golden/make_golden.py with the CPU oracle): the oracle
must keep reproducing them (CPU test), and the CUDA path, through the C ABI, must match them without the oracle in the
loop (GPU test).  Hit counts and ids exact; f32 scores within 1e-5 relative, order identical except among ties."""
import importlib.util
import json
import os

import numpy as np
import pytest

import helpers

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(golden)

REL = 1e-5


def close(a, b):
    return abs(float(a) - float(b)) <= REL * max(abs(float(a)), abs(float(b)), 1e-30)


def load(name):
    return json.load(open(os.path.join(HERE, "golden", name + ".json"), encoding="utf-8"))["cases"]


def check_hits(got_num_hits, got, case, ctx):
    want = case["data"]
    assert int(got_num_hits) == case["num_hits"], ctx
    assert len(got) == len(want), ctx
    want_scores = {i: s for i, s in want}
    for pos, ((gi, gs), (wi, ws)) in enumerate(zip(got, want)):
        assert close(gs, ws), (ctx, pos, gs, ws)
        if gi != wi:  # only a tie may reorder or swap the boundary element
            assert close(want_scores.get(gi, want[-1][1]), ws), (ctx, pos, gi, wi)


def check_suggestions(got, case, ctx):
    assert [(t, i) for t, _, i in got] == [(t, i) for t, _, i in case["items"]], ctx
    for (_, gs, _), (_, ws, _) in zip(got, case["items"]):
        assert close(gs, ws), ctx


@pytest.mark.parametrize("name,build", [("reference_corpus", golden.build_reference_index), ("synthetic_small", golden.build_synthetic_index)])
def test_oracle_reproduces_golden(native_libs, name, build):
    oracle = helpers.Oracle(build())
    for case in load(name):
        if "suggest" in case:
            check_suggestions(oracle.call("suggest_multi", request=case["suggest"]), case, case["suggest"])
            continue
        res = oracle.search(json.dumps(case["request"], ensure_ascii=False))
        check_hits(res["num_hits"], [(h[0], h[1]) for h in res["data"]], case, case["request"])


@pytest.mark.parametrize("name,build", [("reference_corpus", golden.build_reference_index), ("synthetic_small", golden.build_synthetic_index)])
def test_python_oracle_reproduces_golden(native_libs, name, build):
    """The goldens were written by the C++ oracle; the independent Python restatement (oracle/search_py.py: its own index
    decoder, request reading, matching and scoring) must arrive at the same hits: what the CUDA path is held to in the
    GPU test below does not rest on one implementation."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
    import search_py  # noqa: E402  (test infrastructure)
    py = search_py.PySearch(build())
    checked = 0
    for case in load(name):
        if "suggest" in case:
            check_suggestions(py.suggest_multi(case["suggest"]), case, case["suggest"])
            checked += 1
            continue
        try:
            res = py.search(case["request"])
        except search_py.Unsupported:  # regex parts
            continue
        check_hits(res["num_hits"], res["data"], case, case["request"])
        checked += 1
    assert checked >= 10, checked


@pytest.mark.gpu
@pytest.mark.parametrize("name,build", [("reference_corpus", golden.build_reference_index), ("synthetic_small", golden.build_synthetic_index)])
def test_cuda_path_matches_golden(native_libs, name, build):
    import veloci_b200

    assert veloci_b200.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    index = veloci_b200.Index(build())
    cases = load(name)
    searches = [c for c in cases if "request" in c]
    batch = index.prepare([json.dumps(c["request"], ensure_ascii=False) for c in searches]).execute()
    for q, case in enumerate(searches):
        assert batch.status(q) == 0, (case["request"], batch.message(q))
        res = batch.result(q)
        check_hits(res["num_hits"], [(int(i), np.float32(s)) for i, s in res["data"]], case, case["request"])
    batch.close()
    for case in cases:
        if "suggest" in case:
            check_suggestions(index.suggest_multi(case["suggest"]), case, case["suggest"])
