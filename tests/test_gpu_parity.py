"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same
index directories and the same request JSON.

Bar (BASELINE.json north_star): term matches, hit-id sets and counts bit-exact;
f32 scores within 1e-5 relative; top-k order identical except among ties within
that tolerance.
"""
import json
import tempfile

import numpy as np
import pytest

import helpers
import ref_fixtures as fx

pytestmark = pytest.mark.gpu

REL = 1e-5


def close(a, b):
    a, b = float(a), float(b)
    return abs(a - b) <= REL * max(abs(a), abs(b), 1e-30)


def assert_same_topk(gpu, cpu, ctx=""):
    """gpu/cpu: lists of (id, score) in rank order."""
    assert len(gpu) == len(cpu), f"{ctx}: {len(gpu)} hits vs {len(cpu)}"
    cpu_scores = {i: s for i, s in cpu}
    for pos, ((gi, gs), (ci, cs)) in enumerate(zip(gpu, cpu)):
        assert close(gs, cs), f"{ctx}: score at rank {pos}: gpu {gs!r} cpu {cs!r}"
        if gi != ci:
            # only a tie (within tolerance) may reorder or swap the boundary element
            if gi in cpu_scores:
                assert close(cpu_scores[gi], cs), f"{ctx}: id {gi} at rank {pos} is not a tie with {ci}"
            else:
                assert close(gs, cpu[-1][1]), f"{ctx}: id {gi} at rank {pos} missing from the oracle's top-k and not a boundary tie"


@pytest.fixture(scope="module")
def gpu():
    import veloci_b200

    assert veloci_b200.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return veloci_b200


def _pair(gpu, docs=None, config=None, synth=None):
    d = tempfile.mkdtemp(prefix="vb200_gpu_")
    if synth is not None:
        helpers.create_synthetic_index(d, **synth)
    else:
        helpers.create_index(d, docs, config)
    return gpu.Index(d), helpers.Oracle(d), d


def compare_batch(index, oracle, requests, k=10, expect_ok=True):
    out = index.search_batch(requests, k=k)
    ref = oracle.search_batch(requests, threads=4, k=k)
    for q, r in enumerate(requests):
        if expect_ok:
            assert out["status"][q] == 0, f"request {q} failed on the GPU path: {r}"
        if out["status"][q] != 0:
            continue
        assert ref["status"][q] == 0, f"oracle failed on request {q}"
        assert int(out["num_hits"][q]) == int(ref["num_hits"][q]), f"num_hits of request {q}: gpu {out['num_hits'][q]} cpu {ref['num_hits'][q]}: {r}"
        g = [(int(i), s) for i, s in zip(out["ids"][q], out["scores"][q]) if i != 0xFFFFFFFF]
        c = [(int(i), s) for i, s in zip(ref["ids"][q], ref["scores"][q]) if i != 0xFFFFFFFF]
        assert_same_topk(g, c, ctx=f"request {q} {r}")
    return out, ref


SMALL = dict(num_docs=60000, vocab=5000, seed=7)


@pytest.fixture(scope="module")
def small(gpu, native_libs):
    return _pair(gpu, synth=SMALL)


def test_or3_fuzzy_boost(small):  # BASELINE config 2 shape
    index, oracle, _ = small
    reqs = helpers.synthetic_requests(num_queries=400, query_kind="or3", levenshtein=1, **SMALL)
    out, _ = compare_batch(index, oracle, reqs)
    assert out["num_hits"].sum() > 0


def test_single_term_lev2(small):  # BASELINE config 4 shape
    index, oracle, _ = small
    reqs = helpers.synthetic_requests(num_queries=300, query_kind="single", levenshtein=2, query_seed=5, **SMALL)
    compare_batch(index, oracle, reqs)


def test_and_queries(small):  # BASELINE config 3 core (intersection)
    index, oracle, _ = small
    reqs = helpers.synthetic_requests(num_queries=300, query_kind="and", levenshtein=1, query_seed=9, **SMALL)
    compare_batch(index, oracle, reqs)


def test_exact_no_boost_top_skip(small):
    index, oracle, _ = small
    base = [json.loads(r) for r in helpers.synthetic_requests(num_queries=60, query_kind="or3", levenshtein=0, query_seed=11, **SMALL)]
    reqs = []
    for i, r in enumerate(base):
        r.pop("boost", None)
        r["top"] = [1, 3, 10, 25][i % 4]
        if i % 3 == 0:
            r["skip"] = [0, 2, 7][(i // 3) % 3]
        reqs.append(json.dumps(r))
    index_out = index.prepare(reqs).execute()
    for q, r in enumerate(reqs):
        g = index_out.result(q)
        c = oracle.search(r)
        assert g["num_hits"] == c["num_hits"]
        assert_same_topk(g["data"], [(h[0], np.float32(h[1])) for h in c["data"]], ctx=r)


def test_field_search_matches_term_sets(small):  # get_term_ids_in_field, exact term-id sets and scores
    index, oracle, _ = small
    words = [json.loads(r)["search_req"]["search"]["terms"][0] for r in helpers.synthetic_requests(num_queries=120, query_kind="single", query_seed=3, **SMALL)]
    for i, w in enumerate(words):
        part = {"terms": [w], "path": "body", "levenshtein_distance": i % 3}
        if i % 7 == 0:
            part["starts_with"] = True
            part["terms"] = [w[:3]]
        if i % 5 == 0:
            part["ignore_case"] = i % 10 == 0
        if i % 11 == 0:
            part["boost"] = 2.5
        hits, _ = index.field_search(part)
        ref = oracle.call("field_search", part=part)["hits_scores"]
        assert [h[0] for h in hits] == [h[0] for h in ref], part
        for (_, gs), (_, cs) in zip(hits, ref):
            assert close(gs, cs), (part, gs, cs)


def test_errors_per_request(small):
    index, _, _ = small
    reqs = [
        json.dumps({"search_req": {"search": {"terms": ["abc"], "path": "nope"}}}),
        json.dumps({"top": 3}),
        "{not json",
        json.dumps({"search_req": {"search": {"terms": ["abc"], "path": "body"}}, "boost": [{"path": "missing", "boost_fun": "Log10"}]}),
        json.dumps({"search_req": {"search": {"terms": ["abc"], "path": "body"}}}),
    ]
    b = index.prepare(reqs).execute()
    assert [b.status(i) for i in range(5)] == [2, 1, 5, 3, 0]
    assert "field does not exist nope.textindex (fst not found)" in b.message(0)  # tests/all/tests.rs:435


# ---- the reference's own integration fixtures ----------------------------------

def _requests_test_all():
    S = lambda terms, path, **kw: {"search": {"terms": [terms], "path": path, **kw}}
    return [
        {"search_req": S("urge", "meanings.eng[]")},
        {"search_req": S("majestätischer", "meanings.ger[]", levenshtein_distance=1)},
        {"search_req": S("Majestätischer", "meanings.ger[]", ignore_case=False)},
        {"search_req": {"or": {"queries": [S("majestät", "meanings.ger[]"), S("urge", "meanings.eng[]")]}}},
        {"search_req": {"or": {"queries": [S("majestät", "meanings.ger[]"), S("urge", "meanings.eng[]")]}}, "top": 1},
        {"search_req": {"and": {"queries": [S("alle", "meanings.ger[]"), S("meine", "meanings.ger[]"), S("words", "meanings.ger[]")]}}},
        {"search_req": {"and": {"queries": [S("majestät", "meanings.ger[]"), S("majestätischer", "meanings.ger[]")]}}},
        {"search_req": {"and": {"queries": [S("majestät", "meanings.ger[]"), S("urge", "meanings.eng[]")]}}},
        {"search_req": S("awesome", "field1[].text"), "boost": [{"path": "commonness", "boost_fun": "Log10", "param": 1}]},
        {"search_req": S("意慾", "kanji[].text"), "boost": [{"path": "commonness", "boost_fun": "Log10", "param": 1}]},
        {"search_req": {"or": {"queries": [S("awesome", "field1[].text"), S("awesome", "field1[].text")]}}},
        {"search_req": S("COllectif", "title"), "boost": [{"path": "commonness", "boost_fun": "Log2", "param": 2}]},
        {"search_req": S("boostemich", "meanings.ger[]"), "boost": [{"path": "commonness", "boost_fun": "Log2", "param": 2}]},
        {"search_req": S("weich", "meanings.ger[]", levenshtein_distance=1)},
        {"search_req": S("ein", "meanings.ger[]", starts_with=True)},
        {"search_req": {"or": {"queries": [{"and": {"queries": [S("alle", "meanings.ger[]"), S("meine", "meanings.ger[]")]}}, S("urge", "meanings.eng[]")]}}},
        {"search_req": S("will", "meanings.eng[]", boost=3.0)},
    ]


def test_reference_fixture_corpus(gpu, native_libs):  # tests/all/tests.rs corpus
    index, oracle, _ = _pair(gpu, fx.TEST_ALL_DOCS, fx.TEST_ALL_CONFIG)
    reqs = [json.dumps(r, ensure_ascii=False) for r in _requests_test_all()]
    b = index.prepare(reqs).execute()
    for q, r in enumerate(reqs):
        assert b.status(q) == 0, (r, b.message(q))
        g = b.result(q)
        c = oracle.search(r)
        assert g["num_hits"] == c["num_hits"], r
        assert_same_topk(g["data"], [(h[0], np.float32(h[1])) for h in c["data"]], ctx=r)


def test_set_op_step_seam(small):  # set_op.rs:532-551 vector + random lists
    index, oracle, _ = small
    got = index.intersect_hits_score([[(0, 20.0), (10, 20.0)], [(0, 20.0), (3, 20.0), (10, 30.0), (20, 30.0)]])
    assert got == [(0, np.float32(40.0)), (10, np.float32(50.0))]
    rng = np.random.default_rng(3)
    lists, terms = [], ["b", "a", "b", "c"]
    for _ in range(4):
        ids = np.unique(rng.integers(0, SMALL["num_docs"], 500))
        lists.append([(int(i), float(np.float32(rng.random() * 10))) for i in ids])
    inputs = [{"hits_scores": [[i, s] for i, s in l], "term": t} for l, t in zip(lists, terms)]
    ref = oracle.call("union_hits_score", inputs=inputs)
    got = index.union_hits_score(lists, terms)
    assert [g[0] for g in got] == [r[0] for r in ref]
    assert all(close(g[1], r[1]) for g, r in zip(got, ref))
    ref = oracle.call("intersect_hits_score", inputs=inputs[:3])
    got = index.intersect_hits_score(lists[:3])
    assert [g[0] for g in got] == [r[0] for r in ref]
    assert all(close(g[1], r[1]) for g, r in zip(got, ref))


def test_resolve_and_boost_and_topn_step_seam(small):
    index, oracle, _ = small
    part = {"terms": ["abcd"], "path": "body", "levenshtein_distance": 2}
    hits, _ = index.field_search(part)
    ref = oracle.call("resolve_token_to_anchor", part=part)["hits_scores"]
    got, _ = index.resolve_to_anchor(part, hits)
    assert [g[0] for g in got] == [r[0] for r in ref]
    assert all(close(g[1], r[1]) for g, r in zip(got, ref))
    boosted = index.add_boost({"path": "commonness", "boost_fun": "Log10", "param": 1}, got)
    assert len(boosted) == len(got) and any(b[1] != g[1] for b, g in zip(boosted, got)) or not got
    top = index.top_n(got, 7, 2)
    ref_top = oracle.call("top_n_sort", hits=[[i, float(s)] for i, s in got], top=9)[2:9]
    assert_same_topk(top, [(r[0], np.float32(r[1])) for r in ref_top])


def test_plane_path_equals_posting_path(gpu, native_libs, monkeypatch):
    """The head-term plane path and the general posting path are two evaluations of the same
    arithmetic: hit counts, ids and score bits must be identical."""
    d = tempfile.mkdtemp(prefix="vb200_gpu_planes_")
    params = dict(num_docs=150000, vocab=8000, seed=21)
    helpers.create_synthetic_index(d, **params)
    reqs = helpers.synthetic_requests(num_queries=500, query_kind="or3", levenshtein=1, query_seed=3, **params)
    reqs += helpers.synthetic_requests(num_queries=200, query_kind="single", levenshtein=1, query_seed=4, **params)
    with_planes = gpu.Index(d)
    b = with_planes.prepare(reqs)
    b.execute()
    got = b.results_flat(10)
    stats = b.path_stats()
    assert stats["plane_items"] > 0, "the plane path did not run"
    without = gpu.Index(d, planes=False)
    b2 = without.prepare(reqs)
    b2.execute()
    ref = b2.results_flat(10)
    assert b2.path_stats()["plane_items"] == 0
    assert (got["status"] == 0).all() and (ref["status"] == 0).all()
    assert (got["num_hits"] == ref["num_hits"]).all()
    assert (got["ids"] == ref["ids"]).all()
    assert (got["scores"].view(np.uint32) == ref["scores"].view(np.uint32)).all()


def test_deletion_index_equals_dictionary_scan(gpu, native_libs, monkeypatch):
    """Fuzzy matching through the deletion-neighbourhood index finds exactly the terms the dictionary scan finds."""
    d = tempfile.mkdtemp(prefix="vb200_gpu_delidx_")
    params = dict(num_docs=40000, vocab=30000, seed=5)
    helpers.create_synthetic_index(d, **params)
    reqs = []
    for lev, seed in ((0, 1), (1, 2), (2, 3)):
        reqs += helpers.synthetic_requests(num_queries=150, query_kind="single", levenshtein=lev, query_seed=seed, **params)
    probe = gpu.Index(d).search_batch(reqs, k=10)
    scan = gpu.Index(d, deletion_index=False).search_batch(reqs, k=10)
    assert (probe["status"] == 0).all() and (scan["status"] == 0).all()
    assert probe["num_hits"].sum() > 0
    assert (probe["num_hits"] == scan["num_hits"]).all()
    assert (probe["ids"] == scan["ids"]).all()
    assert (probe["scores"].view(np.uint32) == scan["scores"].view(np.uint32)).all()


class _DevArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def test_sharded_index_equals_unsharded(gpu, native_libs):
    """Two anchor-range shards (both on this device), local top-k rows gathered and merged by
    merge_heaps_kernel: identical to the unsharded index."""
    import torch

    d = tempfile.mkdtemp(prefix="vb200_gpu_shards_")
    params = dict(num_docs=90000, vocab=6000, seed=31, tags=30)
    helpers.create_synthetic_index(d, **params)
    reqs = helpers.synthetic_requests(num_queries=300, query_kind="or3", levenshtein=1, query_seed=8, **params)
    reqs += helpers.synthetic_requests(num_queries=100, query_kind="and", levenshtein=1, query_seed=9, **params)  # with facets on tags[]
    assert "facets" in reqs[-1]
    whole_batch = gpu.Index(d).prepare(reqs).execute()
    whole = whole_batch.results_flat(10)
    shards = [gpu.Index(d, shard_rank=r, n_shards=2) for r in range(2)]
    batches = [s.prepare(reqs) for s in shards]
    keys, hits = [], []
    # threshold exchange between the shards after their first tiles (MAX in unsigned order), as bench.py does over NCCL
    sign = -(1 << 63)
    taus = []
    for b in batches:
        b.execute_begin()
        ptr, cnt = b.thresholds()
        taus.append(torch.as_tensor(_DevArray(ptr, cnt), device="cuda"))
    shared = torch.maximum(taus[0] ^ sign, taus[1] ^ sign) ^ sign
    assert (shared != 0).any()
    for t in taus:
        t.copy_(shared)
    torch.cuda.synchronize()
    for b in batches:
        b.execute_finish()
        kp, hp, stride = b.local_topk()
        keys.append(torch.as_tensor(_DevArray(kp, len(reqs) * stride), device="cuda").clone())
        hits.append(torch.as_tensor(_DevArray(hp, len(reqs)), device="cuda").clone())
    g_keys, g_hits = torch.cat(keys), torch.cat(hits)
    # facet histograms: summed over the shards (all-reduce SUM), written back to every shard
    hists = []
    for b in batches:
        ptr, cnt = b.facet_histograms()
        assert cnt > 0  # (the array has one pad word when the histograms' total is even; it is dropped or harmlessly summed)
        hists.append(torch.as_tensor(_DevArray(ptr, cnt // 2), device="cuda").view(torch.int32))
    total = hists[0] + hists[1]
    for h in hists:
        h.copy_(total)
    torch.cuda.synchronize()
    for b in batches:  # every rank merges the same gathered buffers
        b.merge_gathered(g_keys.data_ptr(), g_hits.data_ptr(), 2)
        got = b.results_flat(10)
        assert (got["num_hits"] == whole["num_hits"]).all()
        assert (got["ids"] == whole["ids"]).all()
        assert (got["scores"].view(np.uint32) == whole["scores"].view(np.uint32)).all()
        for q in range(300, 400, 7):
            assert b.result(q)["facets"] == whole_batch.result(q)["facets"]


def test_edge_requests(small):
    """Boundary shapes of the request surface on the bench-shaped index: no hits, top 0, skip past the end, many parts,
    long and non-ASCII terms, a part repeated many times, distance larger than the term."""
    index, oracle, _ = small
    words = [json.loads(r)["search_req"]["search"]["terms"][0] for r in helpers.synthetic_requests(num_queries=14, query_kind="single", query_seed=21, **SMALL)]
    P = lambda w, **kw: {"search": {"terms": [w], "path": "body", **kw}}
    boost = [{"path": "commonness", "boost_fun": "Log10", "param": 1}]
    reqs = [
        {"search_req": P("zzzzzzzzzzzzqqqq")},                                             # no term matches
        {"search_req": P(words[0]), "top": 0},                                              # only the count
        {"search_req": P("zzzzzzzzzzzzqqqq", levenshtein_distance=1), "top": 5, "skip": 250},   # skip past every hit
        {"search_req": P(words[2]), "top": 64, "boost": boost},                             # largest k of the plane path
        {"search_req": P(words[2]), "top": 65, "boost": boost},                             # first k of the general path
        {"search_req": P(words[3]), "top": 200, "skip": 56},                                # k = 256
        {"search_req": {"or": {"queries": [P(w) for w in words[:12]]}}, "boost": boost},    # 12 parts
        {"search_req": {"or": {"queries": [P(words[4])] * 6}}},                             # one part six times
        {"search_req": {"and": {"queries": [P(words[5]), P(words[5])]}}},                   # intersection with itself
        {"search_req": P("ab", levenshtein_distance=5)},                                    # distance clamped to len - 1
        {"search_req": P("a" * 64, levenshtein_distance=2)},                                # longest supported term
        {"search_req": P("größe", levenshtein_distance=1)},                                 # scalars outside the dictionary alphabet
        {"search_req": P("日本語", levenshtein_distance=1)},
        {"search_req": {"or": {"queries": [P(words[6], boost=0.5), P(words[7], boost=3.0), P(words[8])]}}, "boost": boost},
        {"search_req": {"or": {"queries": [{"and": {"queries": [P(words[9]), P(words[10])]}}, {"or": {"queries": [P(words[11]), P(words[12])]}}]}}},
    ]
    texts = [json.dumps(r, ensure_ascii=False) for r in reqs]
    b = index.prepare(texts).execute()
    for q, r in enumerate(texts):
        assert b.status(q) == 0, (r, b.message(q))
        g = b.result(q, cap=256)
        c = oracle.search(r)
        assert g["num_hits"] == c["num_hits"], r
        assert_same_topk(g["data"], [(h[0], np.float32(h[1])) for h in c["data"]], ctx=r)
    # an empty batch and a batch whose every request fails are not errors of the call
    assert index.prepare([]).execute().n == 0
    bad = index.prepare(["{", json.dumps({"top": 1})]).execute()
    assert [bad.status(0), bad.status(1)] == [5, 1]


def test_search_stream_equals_batches(small):
    """Index.search_stream (planning of batch i+1 overlapped with the GPU work of batch i) returns, batch by batch, what
    search_batch returns for the same requests."""
    index, oracle, _ = small
    batches = [helpers.synthetic_requests(num_queries=150, query_kind=kind, levenshtein=1, query_seed=seed, **SMALL) for kind, seed in (("or3", 31), ("and", 32), ("single", 33), ("or3", 34))]
    batches.append([])  # an empty batch in the middle of the stream
    batches.append(batches[0])
    streamed = list(index.search_stream(iter(batches), k=10))
    assert len(streamed) == len(batches)
    for reqs, got in zip(batches, streamed):
        want = index.search_batch(reqs, k=10) if reqs else None
        assert len(got["num_hits"]) == len(reqs)
        if not reqs:
            continue
        assert np.array_equal(got["num_hits"], want["num_hits"])
        assert np.array_equal(got["ids"][:, :10], want["ids"])
        assert np.array_equal(got["scores"][:, :10], want["scores"])
    compare_batch(index, oracle, batches[0])


def test_parts_shared_by_what_the_device_sees(small):
    """The planner unifies search parts by their device form (symbols, distance, flags, boost), not by their JSON: parts
    that differ only in fields that do not change the hits (ignore_case spelled out, a distance above len - 1) share
    one part, parts that differ in a field that does change them (boost, starts_with, distance) do not."""
    index, oracle, _ = small
    w = [json.loads(r)["search_req"]["search"]["terms"][0] for r in helpers.synthetic_requests(num_queries=3, query_kind="single", query_seed=77, **SMALL)]
    P = lambda t, **kw: {"search": {"terms": [t], "path": "body", **kw}}
    reqs = [
        {"search_req": P(w[0], levenshtein_distance=1)},
        {"search_req": P(w[0], levenshtein_distance=1, ignore_case=True)},
        {"search_req": P(w[0], levenshtein_distance=1, boost=2.0)},
        {"search_req": P(w[0], levenshtein_distance=1, starts_with=True)},
        {"search_req": P(w[0], levenshtein_distance=2)},
        {"search_req": P(w[0])},
        {"search_req": P("ab", levenshtein_distance=1)},
        {"search_req": P("ab", levenshtein_distance=7)},
        {"search_req": {"or": {"queries": [P(w[0], levenshtein_distance=1), P(w[0], levenshtein_distance=1, boost=3.0), P(w[1])]}}},
        {"search_req": {"and": {"queries": [P(w[1], levenshtein_distance=1, ignore_case=True), P(w[1], levenshtein_distance=1), P(w[2], starts_with=True)]}}},
    ]
    compare_batch(index, oracle, [json.dumps(r) for r in reqs] * 30, k=10)


def test_batches_as_lines_with_line_feeds_and_empty_requests(small):
    """Batches of 64+ request strings travel as one line-feed separated buffer (vgpu_batch_prepare_lines); a request that
    contains a line feed (pretty-printed JSON) sends the batch down the one-string-per-request path instead, an empty
    request fails alone.  Results are the same either way."""
    index, oracle, _ = small
    reqs = helpers.synthetic_requests(num_queries=100, query_kind="or3", levenshtein=1, query_seed=41, **SMALL)
    plain = index.search_batch(reqs, k=10)
    b = index.prepare(reqs).execute()
    lines = b.results_flat(10)
    b.close()
    assert np.array_equal(lines["ids"], plain["ids"]) and np.array_equal(lines["num_hits"], plain["num_hits"])
    pretty = list(reqs)
    pretty[7] = json.dumps(json.loads(reqs[7]), indent=2)
    assert "\n" in pretty[7]
    pretty[50] = ""
    pretty[99] = "\n"
    b = index.prepare(pretty).execute()
    got = b.results_flat(10)
    assert [int(s) for s in got["status"][[7, 50, 99]]] == [0, 5, 5]
    ok = got["status"] == 0
    assert ok.sum() == 98
    assert np.array_equal(got["ids"][ok], plain["ids"][ok]) and np.array_equal(got["num_hits"][ok], plain["num_hits"][ok])
    b.close()
    with_empty = list(reqs)
    with_empty[0] = ""
    with_empty[63] = ""
    b = index.prepare(with_empty).execute()   # stays on the lines path: an empty line is an empty request
    got = b.results_flat(10)
    assert [int(s) for s in got["status"][[0, 63]]] == [5, 5] and (got["status"][1:63] == 0).all()
    assert np.array_equal(got["ids"][1:63], plain["ids"][1:63])
    b.close()
