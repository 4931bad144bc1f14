"""search::explain_plan (src/search.rs:132-141): the request's plan as a Graphviz dot graph (csrc/host/explain_plan.hpp).
The reference pins its contents only by containment (tests.rs:1210-1230); beyond that the graph is checked for the
structure plan_creator builds: one field search per distinct part, a resolve per leaf, set operations before their
inputs, filter, boosts and phrase steps chained after the tree."""
import re

import helpers

S = lambda term, path, **kw: {"search": {"terms": [term], "path": path, **kw}}


def graph(request):
    text = helpers.explain_plan(request)
    assert text.startswith("digraph example2 {\n") and text.endswith("}\n")
    nodes = {int(i): label for i, label in re.findall(r'^    N(\d+)\[label="(.*)"\];$', text, re.M)}
    edges = [(int(a), int(b)) for a, b in re.findall(r'^    N(\d+) -> N(\d+)\[label=""\];$', text, re.M)]
    assert len(nodes) + len(edges) + 2 == len(text.splitlines())
    return text, nodes, edges


def test_reference_containment():  # tests.rs:1210-1230
    req = {"search_req": {"search": {"terms": ["weich"], "path": "meanings.ger[]", "levenshtein_distance": 1, "firstCharExactMatch": True}},
           "boost": [{"path": "commonness", "boost_fun": "Log2", "param": 2}]}
    text, nodes, edges = graph(req)
    low = text.lower()
    assert "weich" in low and "meanings.ger[]" in low and "boost" in low
    assert nodes == {0: "search meanings.ger[] weich\\n", 1: "token to anchor\\n", 2: "BoostPlanStepFromBoostRequest\\n"}
    assert edges == [(0, 1), (1, 2)]


def test_tree_filter_boosts_and_phrases():
    a, b, c = S("majestät", "meanings.ger[]"), S("urge", "meanings.eng[]"), S("意慾", "kanji[].text")
    req = {"search_req": {"or": {"queries": [{"and": {"queries": [a, b]}}, c, a]}}, "filter": S("20", "commonness"),
           "boost": [{"path": "kanji[].commonness", "boost_fun": "Log10"}, {"path": "commonness", "boost_fun": "Add"}],
           "phrase_boosts": [{"search1": a["search"], "search2": b["search"]}]}
    text, nodes, edges = graph(req)
    labels = [nodes[i] for i in sorted(nodes)]
    assert labels[:4] == ["search meanings.ger[] majest\\u{e4}t\\n", "search meanings.eng[] urge\\n", "search kanji[].text \\u{610f}\\u{617e}\\n", "search commonness 20\\n"]  # one per distinct part
    rest = labels[4:]
    assert rest == ["token to anchor\\n",                                       # the filter's leaf
                    "Union\\n", "Intersect\\n", "token to anchor\\n", "token to anchor\\n",    # set operations come before their inputs
                    "token to anchor\\n", "BoostToAnchor kanji[].commonness\\n", "ApplyAnchorBoost\\n",  # the part on the boost's 1:n level
                    "token to anchor\\n",                                       # `a` again: searched once, resolved twice
                    "IntersectScoresWithIds\\n", "BoostPlanStepFromBoostRequest\\n", "PlanStepPhrasePairToAnchorId\\n", "BoostAnchorFromPhraseResults\\n"]
    union, filt = 5, 4
    into_union = sorted(src for src, dst in edges if dst == union)
    assert into_union == [4, 6, 11, 12]  # its three inputs and the filter it waits for
    assert (0, 15) in edges and (1, 15) in edges  # the phrase pair step reads both field searches
    assert all(src < len(nodes) and dst < len(nodes) for src, dst in edges)
    assert sum(1 for src, dst in edges if src == filt) >= 4  # every resolve of the tree waits for the filter


def test_errors():
    import pytest
    with pytest.raises(RuntimeError):
        helpers.explain_plan({"top": 3})


def test_through_the_c_abi(native_libs):  # vgpu_explain_plan needs neither an index nor a device
    import veloci_b200.api as api
    req = {"search_req": {"or": {"queries": [S("majestät", "meanings.ger[]"), S("urge", "meanings.eng[]")]}}, "explain": True}
    assert api.explain_plan(req) == helpers.explain_plan(req)
