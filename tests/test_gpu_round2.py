"""GPU tests of the round-2 boundary: many anchor-range shards, plan export / import, the shared-memory plan channel,
the in-library NCCL exchange (needs two GPUs), open flags, load-time validation.  All through the C ABI."""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHARDED = dict(num_docs=120000, vocab=7000, seed=33, tags=30)


@pytest.fixture(scope="module")
def gpu():
    import veloci_b200

    assert veloci_b200.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return veloci_b200


@pytest.fixture(scope="module")
def sharded_corpus(gpu, native_libs):
    d = tempfile.mkdtemp(prefix="vb200_r2_shards_")
    helpers.create_synthetic_index(d, **SHARDED)
    reqs = helpers.synthetic_requests(num_queries=400, query_kind="or3", levenshtein=1, query_seed=8, **SHARDED)
    reqs += helpers.synthetic_requests(num_queries=120, query_kind="and", levenshtein=1, query_seed=9, **SHARDED)  # with facets on tags[]
    reqs += helpers.synthetic_requests(num_queries=80, query_kind="single", levenshtein=2, query_seed=10, **SHARDED)
    assert "facets" in reqs[450]
    return d, reqs


class _DevArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def run_shards_on_one_device(gpu, d, reqs, n_shards, k=10):
    """The host-driven exchange (execute_begin / thresholds / execute_finish / local_topk / merge_gathered) over
    `n_shards` handles on this device: what vgpu_batch_execute does over NCCL when every shard has its own GPU."""
    import torch

    shards = [gpu.Index(d, shard_rank=r, n_shards=n_shards) for r in range(n_shards)]
    batches = [s.prepare(reqs) for s in shards]
    sign = -(1 << 63)
    taus = []
    for b in batches:
        b.execute_begin()
        ptr, cnt = b.thresholds()
        taus.append(torch.as_tensor(_DevArray(ptr, cnt), device="cuda"))
    shared = taus[0] ^ sign
    for t in taus[1:]:
        shared = torch.maximum(shared, t ^ sign)
    shared = shared ^ sign
    for t in taus:
        t.copy_(shared)
    torch.cuda.synchronize()
    keys, hits, hists = [], [], []
    for b in batches:
        b.execute_finish()
        kp, hp, stride = b.local_topk()
        keys.append(torch.as_tensor(_DevArray(kp, len(reqs) * stride), device="cuda").clone())
        hits.append(torch.as_tensor(_DevArray(hp, len(reqs)), device="cuda").clone())
        ptr, cnt = b.facet_histograms()
        if cnt:
            hists.append(torch.as_tensor(_DevArray(ptr, cnt // 2), device="cuda").view(torch.int32))
    g_keys, g_hits = torch.cat(keys), torch.cat(hits)
    if hists:
        total = hists[0].clone()
        for h in hists[1:]:
            total += h
        for h in hists:
            h.copy_(total)
    torch.cuda.synchronize()
    out = []
    for b in batches:  # every rank merges the same gathered buffers
        b.merge_gathered(g_keys.data_ptr(), g_hits.data_ptr(), n_shards)
        out.append(b)
    return shards, out


@pytest.mark.parametrize("n_shards", [3, 8])
def test_many_shards_equal_unsharded_and_oracle(gpu, sharded_corpus, n_shards):
    """Eight (and three: uneven cuts) anchor-range shards of one index, merged on the device: identical, bit for bit, to
    the unsharded index -- ids, score bits, hit counts, facet groups -- and equal to the CPU oracle."""
    d, reqs = sharded_corpus
    whole_batch = gpu.Index(d).prepare(reqs).execute()
    whole = whole_batch.results_flat(10)
    assert (whole["status"] == 0).all()
    shards, batches = run_shards_on_one_device(gpu, d, reqs, n_shards)
    ranges = [s.info() for s in shards]
    assert ranges[0]["anchor_lo"] == 0 and ranges[-1]["anchor_hi"] == SHARDED["num_docs"]
    assert all(ranges[i]["anchor_hi"] == ranges[i + 1]["anchor_lo"] for i in range(n_shards - 1))
    for b in (batches[0], batches[-1]):
        got = b.results_flat(10)
        assert (got["status"] == 0).all()
        assert (got["num_hits"] == whole["num_hits"]).all()
        assert (got["ids"] == whole["ids"]).all()
        assert (got["scores"].view(np.uint32) == whole["scores"].view(np.uint32)).all()
        for q in range(400, 520, 7):
            assert b.result(q)["facets"] == whole_batch.result(q)["facets"]
    rows = list(range(0, len(reqs), 5))
    ref = helpers.Oracle(d).search_batch([reqs[q] for q in rows], threads=4, k=10)
    par = helpers.batch_parity(batches[0].results_flat(10), ref, rows)
    assert par["equal"] == par["checked"], par


def test_one_to_n_boost_is_refused_on_a_sharded_index(gpu, native_libs):
    """apply_boost_values_anchor (boost.rs:255-281) depends on the run of boosted hits before an anchor, which may begin in
    the previous shard: the sharded handle answers VGPU_ERR_UNSUPPORTED for that request shape instead of a result that
    could differ; the unsharded handle answers it."""
    rng = np.random.default_rng(5)
    words = ["kami", "kumo", "kawa", "mori", "sora", "tori"]
    docs = [{"ent_seq": str(i), "kana": [{"text": str(rng.choice(words)), "commonness": int(rng.integers(1, 90))} for _ in range(int(rng.integers(1, 3)))]} for i in range(400)]
    d = tempfile.mkdtemp(prefix="vb200_r2_1n_")
    helpers.create_index(d, docs, {"kana[].text": {"fulltext": {"tokenize": True}}, "kana[].commonness": {"boost": {"boost_type": "f32"}}})
    req = {"search_req": {"search": {"terms": ["kami"], "path": "kana[].text"}}, "boost": [{"path": "kana[].commonness", "boost_fun": "Log10", "param": 1}]}
    plain = {"search_req": {"search": {"terms": ["kami"], "path": "kana[].text"}}}
    whole = gpu.Index(d).search_batch([json.dumps(req), json.dumps(plain)])
    assert list(whole["status"]) == [0, 0] and whole["num_hits"][0] > 0
    shard = gpu.Index(d, shard_rank=1, n_shards=2)
    b = shard.prepare([json.dumps(req), json.dumps(plain)]).execute()
    assert b.status(0) == 8 and "1:n" in b.message(0)
    assert b.status(1) == 0


def test_plan_export_import(gpu, sharded_corpus):
    """A plan exported by one handle and imported by other handles of the same directory (unsharded and a shard) gives the
    results of planning there; a plan of another index is refused."""
    d, reqs = sharded_corpus
    a = gpu.Index(d)
    ba = a.prepare(reqs)
    blob = ba.export_plan()
    assert len(blob) > 1000
    want = ba.execute().results_flat(10)
    b = gpu.Index(d)
    bb = b.prepare(len(reqs), plan=blob).execute()
    got = bb.results_flat(10)
    for key in ("status", "num_hits", "ids"):
        assert (got[key] == want[key]).all(), key
    assert (got["scores"].view(np.uint32) == want["scores"].view(np.uint32)).all()
    for q in range(400, 520, 11):
        assert bb.result(q)["facets"] == ba.result(q)["facets"]
    # a shard imports the unsharded handle's plan: same local rows as planning on the shard
    s1 = gpu.Index(d, shard_rank=1, n_shards=2)
    own = s1.prepare(reqs).execute().results_flat(10)
    imp = s1.prepare(len(reqs), plan=blob).execute().results_flat(10)
    for key in ("status", "num_hits", "ids"):
        assert (imp[key] == own[key]).all(), key
    # failed requests travel with their status and message
    mixed = [reqs[0], "{not json", json.dumps({"search_req": {"search": {"terms": ["x"], "path": "nosuchfield"}}}), reqs[1]]
    bm = a.prepare(mixed)
    bi = b.prepare(len(mixed), plan=bm.export_plan()).execute()
    bm.execute()
    assert [bi.status(q) for q in range(4)] == [bm.status(q) for q in range(4)] == [0, 5, 2, 0]
    assert bi.message(2) == bm.message(2)
    # another index: refused
    other = tempfile.mkdtemp(prefix="vb200_r2_other_")
    helpers.create_synthetic_index(other, num_docs=5000, vocab=500, seed=1)
    with pytest.raises(gpu.VelociGpuError) as e:
        gpu.Index(other).prepare(len(reqs), plan=blob)
    assert e.value.status == 1
    with pytest.raises(gpu.VelociGpuError):
        b.prepare(3, plan=blob[: len(blob) // 2])


_CHANNEL_WORKER = r"""
import json, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import numpy as np
import veloci_b200
d, name, n_batches, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
index = veloci_b200.Index(d)
ch = veloci_b200.PlanChannel(name, 1, 2, capacity=32 << 20)
res = []
for out_flat in index.search_stream(([] for _ in range(n_batches)), k=10, channel=ch):
    res.append(out_flat)
np.savez(out, **{{f"{{k}}{{i}}": v for i, r in enumerate(res) for k, v in r.items()}})
"""


def test_plan_channel_between_two_processes(gpu, sharded_corpus):
    """Local rank 0 (this process) plans and publishes five different batches through the shared-memory channel; a second
    process imports each plan (it never sees the request text) and must produce the same results.  Five batches through
    two slots exercise the slot hand-over in both directions."""
    d, reqs = sharded_corpus
    n_batches = 5
    batches = [reqs[i * 100:(i + 1) * 100 + 37] for i in range(n_batches)]
    name = f"/vb200_test_{os.getpid()}"
    out = tempfile.mktemp(suffix=".npz")
    child = subprocess.Popen([sys.executable, "-c", _CHANNEL_WORKER.format(root=ROOT), d, name, str(n_batches), out])
    try:
        index = gpu.Index(d)
        ch = gpu.PlanChannel(name, 0, 2, capacity=32 << 20)
        mine = list(index.search_stream(iter(batches), k=10, channel=ch))
        assert child.wait(timeout=300) == 0
    finally:
        if child.poll() is None:
            child.kill()
    theirs = np.load(out)
    for i, r in enumerate(mine):
        direct = index.search_batch(batches[i], k=10)
        for key in ("status", "num_hits", "ids"):
            assert (r[key] == direct[key]).all(), (i, key)
            assert (theirs[f"{key}{i}"] == direct[key]).all(), (i, key)
        assert (theirs[f"scores{i}"].view(np.uint32) == direct["scores"].view(np.uint32)).all()
    ch.close()


def test_load_rejects_postings_beyond_num_docs(gpu, native_libs):
    """Planes, level bitmaps and tile buckets are sized from metaData.json's num_docs: an index whose postings reach past it
    (stale metadata) must fail to load (VGPU_ERR_IO), not write out of bounds."""
    src = tempfile.mkdtemp(prefix="vb200_r2_meta_")
    helpers.create_synthetic_index(src, num_docs=3000, vocab=300, seed=2)
    meta_path = os.path.join(src, "metaData.json")
    meta = json.load(open(meta_path))
    assert meta["num_docs"] == 3000
    gpu.Index(src).close()
    meta["num_docs"] = 2000
    json.dump(meta, open(meta_path, "w"))
    with pytest.raises(gpu.VelociGpuError) as e:
        gpu.Index(src)
    assert e.value.status == 4 and "num_docs" in e.value.message
    shutil.rmtree(src, ignore_errors=True)


def test_kernel_profile_and_work_stats(gpu, sharded_corpus):
    d, reqs = sharded_corpus
    b = gpu.Index(d).prepare(reqs[:400])
    b.execute()
    base = b.results_flat(10)
    times = b.profile_execute()
    assert "plane_eval" in times and times["plane_eval"]["ms"] > 0 and "plane_seed" in times and "sparse_fill" in times
    again = b.results_flat(10)
    assert (again["ids"] == base["ids"]).all()
    w = b.work_stats()
    assert w["plane_items"] > 0 and w["tiles"] == (SHARDED["num_docs"] + 8191) // 8192 and w["sparse_entries"] > 0


@pytest.mark.skipif("__import__('veloci_b200').device_count() < 2")
def test_index_on_a_second_device(gpu, sharded_corpus):
    """Kernel attributes (opt-in shared memory) are per device: an index on device 1 must work in a process that already
    ran every kernel on device 0."""
    d, reqs = sharded_corpus
    first = gpu.Index(d, device=0).search_batch(reqs, k=10)
    second = gpu.Index(d, device=1).search_batch(reqs, k=10)
    assert (second["status"] == 0).all()
    for key in ("num_hits", "ids"):
        assert (first[key] == second[key]).all()


@pytest.mark.skipif("__import__('veloci_b200').device_count() < 2")
def test_in_library_exchange_over_nccl(gpu, sharded_corpus):
    """Two processes, one GPU each: vgpu_comm_init, the plan channel, then vgpu_batch_execute as a collective.  Every rank
    must end with the unsharded result (tools/sharded_worker.py checks it on both ranks)."""
    d, reqs = sharded_corpus
    req_path = tempfile.mktemp(suffix=".jsonl")
    open(req_path, "w").write("\n".join(reqs))
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tools", "sharded_worker.py"), d, req_path]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "sharded_worker ok" in r.stdout


def test_ids_set_ops_and_filter_steps(gpu, sharded_corpus):
    """union_hits_ids, intersect_hits_ids, IntersectScoresWithIds and the filtered ResolveTokenIdToAnchor as step entry points,
    against the oracle's restatement of set_op.rs (with its quirks: one input passes through unsorted, empty other inputs of
    an intersection are skipped, an empty id list keeps every scored hit)."""
    d, _ = sharded_corpus
    index, oracle = gpu.Index(d), helpers.Oracle(d)
    rng = np.random.default_rng(12)
    N = SHARDED["num_docs"]
    lists = [[int(x) for x in rng.choice(N, size=n, replace=False)] for n in (900, 300, 1500, 40)]
    lists[1] = lists[1] + lists[0][:200]  # overlaps
    lists[2] = lists[2] + lists[0][100:400] + lists[1][:120]
    lists[3] = lists[3] + lists[0][150:260]
    as_inputs = lambda ls: [{"hits_ids": l} for l in ls]
    for pick in ([0, 1], [0, 1, 2], [0, 1, 2, 3], [2], [1, 3]):
        ls = [lists[i] for i in pick]
        assert index.union_hits_ids(ls) == oracle.call("union_hits_ids", inputs=as_inputs(ls)), pick
        assert index.intersect_hits_ids(ls) == oracle.call("intersect_hits_ids", inputs=as_inputs(ls)), pick
    with_empty = [lists[0], [], lists[2]]
    assert index.intersect_hits_ids(with_empty) == oracle.call("intersect_hits_ids", inputs=as_inputs(with_empty))
    assert index.intersect_hits_ids([[], lists[2]]) == oracle.call("intersect_hits_ids", inputs=as_inputs([[], lists[2]])) == []
    assert index.union_hits_ids([]) == [] and index.intersect_hits_ids([]) == []
    hits = [(i, float(np.float32(rng.random() * 7))) for i in lists[0]]
    for ids in (lists[2], []):
        ref = oracle.call("intersect_score_hits_with_ids", inputs=[{"hits_scores": [[i, s] for i, s in hits]}, {"hits_ids": ids}])
        got = index.intersect_scores_with_ids(hits, ids)
        assert [(g[0], float(g[1])) for g in got] == [(r[0], float(np.float32(r[1]))) for r in ref]
    # the filtered resolve: the oracle's resolve, cut to the filter
    part = {"terms": ["abcd"], "path": "body", "levenshtein_distance": 2}
    term_hits, _ = index.field_search(part)
    full = oracle.call("resolve_token_to_anchor", part=part)["hits_scores"]
    assert len(full) > 20
    keep = sorted(r[0] for r in full[::3]) + [N - 1]
    got = index.resolve_to_anchor_filtered(part, term_hits, keep)
    want = [r for r in full if r[0] in set(keep)]
    assert [g[0] for g in got] == [r[0] for r in want]
    assert all(abs(float(g[1]) - r[1]) <= 1e-5 * abs(r[1]) for g, r in zip(got, want))
    assert index.resolve_to_anchor_filtered(part, term_hits, []) == []


def test_facet_step(gpu, sharded_corpus):
    """get_facet as a step entry point: groups of a facet field over an explicit id list (facet.rs:31-73)."""
    d, reqs = sharded_corpus
    index, oracle = gpu.Index(d), helpers.Oracle(d)
    field = json.loads(reqs[450])["facets"][0]["field"]
    rng = np.random.default_rng(4)
    ids = sorted(int(x) for x in rng.choice(SHARDED["num_docs"], size=5000, replace=False))
    for top in (3, 10, None):
        ref = oracle.call("get_facet", field=field, top=top, ids=ids)
        got = index.facet({"field": field, "top": top}, ids)
        assert [(t, c, i) for t, c, i in got] == [(r[0], r[1], r[2]) for r in ref], top
    assert index.facet({"field": field}, []) == []


def test_generated_requests_on_the_reference_corpus(gpu, native_libs):
    """Request generation through the C ABI (vgpu_search_query, vgpu_suggest_query) on the corpus of
    tests/all/test_query_generator.rs: the text the library generates equals the plain-Python oracle's request, the batch
    planner takes it as is, and the hits are the CPU oracle's (and the ones the reference's tests assert)."""
    import ref_fixtures as fx

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import query_generator as qg  # test infrastructure

    d = tempfile.mkdtemp(prefix="vb200_r2_qg_")
    helpers.create_index(d, fx.TEST_QG_DOCS, fx.TEST_QG_CONFIG)
    index, oracle = gpu.Index(d), helpers.Oracle(d)
    meta = json.load(open(os.path.join(d, "metaData.json")))
    cat = qg.Catalog(list(meta["columns"]), [f for f, c in meta["columns"].items() if any(i["path"] == f + ".textindex.to_anchor_id_score" for i in c["indices"])])
    cases = [
        ({"search_term": "urge"}, 1, "1587690"),                                            # test_query_generator.rs:170-179
        ({"search_term": "ent_seq:99999"}, 1, "99999"),                                     # :182-189
        ({"search_term": "ent_seq:99999", "parser_options": {"no_attributes": True}}, 1, "1337"),  # :192-204
        ({"search_term": "urge OR いよく"}, 3, "1587690"),                                   # :207-216
        ({"search_term": "urge AND いよく"}, 1, "1587690"),                                  # :219-228
        ({"search_term": "urge AND いよく AND awesome"}, 0, None),                            # :296-303
        ({"search_term": "awes*"}, 1, None), ({"search_term": "いよ*"}, 3, None), ({"search_term": "awesam*"}, 1, None),  # :306-326
        ({"search_term": "will", "top": 10, "facets": ["commonness", "kanji[].commonness"], "levenshtein": 0, "boost_fields": {"meanings.eng[]": 1.5}}, 2, None),  # :270-280
        ({"search_term": "will", "top": 10, "levenshtein": 0, "boost_fields": {"meanings.eng[]": 1.5}, "boost_terms": {"meanings.ger[]:majestätisches Aussehen (n)": 20.0}}, 2, "1337"),  # :282-294
        ({"search_term": "*wesom*", "fields": ["tags[]"]}, 1, "1587700"),                    # :328-336 (regex part)
        ({"search_term": "*we*some", "fields": ["tags[]"]}, 1, "1587700"),                   # :338-346
        ({"search_term": "tags[]:*wesam*"}, 0, None),                                        # :349-356
        ({"search_term": "will urge", "phrase_pairs": True, "fields": ["meanings.eng[]", "meanings.ger[]"], "text_locality": True,
          "boost_queries": [{"path": "commonness", "boost_fun": "Log10", "param": 1}], "filter": "commonness:20 OR ent_seq:(25 26)"}, None, None),
    ]
    texts = []
    for params, _, _ in cases:
        text = index.search_query(params)
        want = json.loads(json.dumps(qg.search_query(cat, params)))
        assert json.loads(text).keys() == want.keys() and json.loads(text)["search_req"] == want["search_req"], params
        texts.append(text)
    b = index.prepare(texts).execute()
    for q, (params, n, first) in enumerate(cases):
        assert b.status(q) == 0, (params, b.message(q))
        g, c = b.result(q), oracle.search(texts[q])
        assert g["num_hits"] == c["num_hits"], params
        ok, why = helpers.same_topk([(i, float(s)) for i, s in g["data"]], [(h[0], float(np.float32(h[1]))) for h in c["data"]])
        assert ok, (params, why)
        if n is not None:
            assert len(g["data"]) == n, params
        if first is not None:
            assert fx.TEST_QG_DOCS[g["data"][0][0]]["ent_seq"] == first, params
            assert index.get_doc(g["data"][0][0]) == fx.TEST_QG_DOCS[g["data"][0][0]]  # hits[0].doc of the reference's tests
    with pytest.raises(gpu.VelociGpuError) as e:
        index.search_query(search_term="notexistingfield:awes*")                              # :369-378
    assert e.value.status == 2 and "Field notexistingfield not found in" in str(e.value)
    with pytest.raises(gpu.VelociGpuError) as e:
        index.search_query(search_term="awes*", fields=["notexistingfield"])                  # :358-367
    assert e.value.status == 1 and "All fields filtered" in str(e.value)
    # suggest_query feeds vgpu_suggest
    text = index.suggest_query("Begeisteru", top=10, levenshtein=1, fields=["meanings.ger[]"])
    assert json.loads(text) == qg.suggest_query(cat, "Begeisteru", top=10, levenshtein=1, fields=["meanings.ger[]"])
    got = index.suggest_multi(json.loads(text))
    ref = oracle.call("suggest_multi", request=json.loads(text))
    assert [g[0] for g in got] == [r[0] for r in ref] and "begeisterung" in [g[0] for g in got]


def test_regex_parts(gpu, sharded_corpus):
    """`is_regex` parts (search_field.rs:72-83): regex_match_kernel runs the pattern's DFA over the dictionary; term hits
    (ids and scores) against the oracle's thread-list simulation, alone and mixed with fuzzy parts in one batch."""
    d, _ = sharded_corpus
    index, oracle = gpu.Index(d), helpers.Oracle(d)
    parts = [
        {"path": "body", "terms": [".*ab.*"], "is_regex": True},
        {"path": "body", "terms": ["a.*z"], "is_regex": True},
        {"path": "body", "terms": ["[a-c]{2}x?.*q"], "is_regex": True},
        {"path": "body", "terms": ["ab"], "is_regex": True, "starts_with": True},
        {"path": "body", "terms": ["AB.*"], "is_regex": True, "ignore_case": False},
        {"path": "body", "terms": ["AB.*"], "is_regex": True, "ignore_case": True, "boost": 2.5},
        {"path": "body", "terms": ["(ab|ba)+"], "is_regex": True, "levenshtein_distance": 1},
        {"path": "body", "terms": ["zzzzzzzzzzzz"], "is_regex": True},
    ]
    n_matched = 0
    for part in parts:
        got, _ = index.field_search(part)
        ref = oracle.call("field_search", part=part)["hits_scores"]
        assert sorted(g[0] for g in got) == sorted(r[0] for r in ref), part
        by_id = {r[0]: r[1] for r in ref}
        assert all(abs(float(s) - by_id[i]) <= 1e-5 * abs(by_id[i]) for i, s in got), part
        n_matched += len(got)
    assert n_matched > 100
    boost = [{"path": "commonness", "boost_fun": "Log10", "param": 1}]
    reqs = [
        {"search_req": {"search": parts[0]}, "boost": boost},
        {"search_req": {"or": {"queries": [{"search": parts[1]}, {"search": {"path": "body", "terms": ["abcd"], "levenshtein_distance": 1}}]}}, "boost": boost, "top": 20},
        {"search_req": {"and": {"queries": [{"search": parts[0]}, {"search": parts[3]}]}}},
        {"search_req": {"search": {**parts[0], "top": 3}}},
        {"search_req": {"search": {"path": "body", "terms": ["^ab"], "is_regex": True}}},
        {"search_req": {"search": {"path": "body", "terms": [r"\w+"], "is_regex": True}}},
    ]
    texts = [json.dumps(r) for r in reqs]
    b = index.prepare(texts).execute()
    for q in range(4):
        assert b.status(q) == 0, (reqs[q], b.message(q))
        g, c = b.result(q), oracle.search(texts[q])
        assert g["num_hits"] == c["num_hits"] and g["num_hits"] > 0, reqs[q]
        ok, why = helpers.same_topk([(i, float(s)) for i, s in g["data"]], [(h[0], float(np.float32(h[1]))) for h in c["data"]])
        assert ok, (reqs[q], why)
    assert b.status(4) == 1 and "anchors" in b.message(4)       # the reference panics on this build
    assert b.status(5) == 8                                      # valid for the reference, outside the implemented syntax
    # a plan with regex parts travels: export on one handle, import on another
    blob = index.prepare(texts[:3]).export_plan()
    other = gpu.Index(d)
    b2 = gpu.Batch(other, texts[:3], plan=blob).execute()
    for q in range(3):
        assert b2.result(q)["num_hits"] == b.result(q)["num_hits"]
        assert [i for i, _ in b2.result(q)["data"]] == [i for i, _ in b.result(q)["data"]]


def test_result_docs_and_why_found(gpu, native_libs):
    """search::to_search_result through the C ABI (vgpu_batch_result_docs): documents from the compressed store, why_found
    highlights from the terms the GPU matched; the reference's expected strings (tests/all/test_why_found.rs) and the
    plain-Python oracle (oracle/highlight.py) on the same matched terms."""
    import ref_fixtures as fx

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import highlight as ohl  # test infrastructure

    d = tempfile.mkdtemp(prefix="vb200_r2_wf_")
    helpers.create_index(d, fx.TEST_WHYFOUND_DOCS, fx.TEST_WHYFOUND_CONFIG)
    columns = json.load(open(os.path.join(d, "metaData.json")))["columns"]
    index, oracle = gpu.Index(d), helpers.Oracle(d)
    S = lambda term, path, **kw: {"search_req": {"search": {"terms": [term], "path": path, **kw}}, "why_found": True}
    cases = [
        (S("veloci", "url"), "url", [["https://github.com/PSeitz/<b>veloci</b>"]]),                                   # :73-92
        (S("test", "custom_tokenized"), "custom_tokenized", [["<b>test</b>§_ cool _"]]),                               # :94-105
        (S("_ cool _", "custom_tokenized"), "custom_tokenized", [["test§<b>_ cool _</b>"]]),                           # :118-129
        (S("<<cool>>", "custom_tokenized"), "custom_tokenized", [["<b><<cool>></b>"]]),                                # :149-161
        (S("ID1000", "not_tokenized"), "not_tokenized", [["<b>ID1000</b>"]]),                                          # :163-175
        (S("ID1000", "not_tokenized_1_n[]"), "not_tokenized_1_n[]", [["<b>ID1000</b>"]]),                              # :193-205
        (S("schön", "richtig", levenshtein_distance=1), "richtig", [["<b>schön</b> super"], ["<b>shön</b>"]]),         # :237-251
        (S("treffers", "viele[]", levenshtein_distance=1), "viele[]", [["<b>treffers</b>", "super <b>treffers</b>"]]), # :253-266 (first hit)
        (S("umsortiert", "viele[]", levenshtein_distance=0), "viele[]", [[" ... zu checken, dass da nicht <b>umsortiert</b> wird"]]),  # :285-299
        ({"search_req": {"or": {"queries": [{"search": {"terms": ["Taschenbuch"], "path": "buch", "levenshtein_distance": 1}},
                                            {"search": {"terms": ["kartoniert"], "path": "buch", "levenshtein_distance": 1}}]}}, "why_found": True},
         "buch", [["<b>Taschenbuch</b> (<b>kartoniert</b>)"]]),                                                        # :318-350
        ({"search_req": {"search": {"terms": ["schön"], "path": "richtig"}}}, "richtig", None),                        # why_found not asked for
    ]
    texts = [json.dumps(c[0], ensure_ascii=False) for c in cases]
    b = index.prepare(texts).execute()
    for q, (req, field, expected) in enumerate(cases):
        res = b.result_docs(q)
        ref = oracle.search(texts[q])
        assert res["num_hits"] == ref["num_hits"] and [h["hit"]["id"] for h in res["data"]] == [h[0] for h in ref["data"]], req
        for h in res["data"]:
            assert h["doc"] == fx.TEST_WHYFOUND_DOCS[h["hit"]["id"]]
        if expected is None:
            assert all(h["why_found"] == {} for h in res["data"])
            continue
        for h, want in zip(res["data"], expected):
            assert h["why_found"][field] == want, (req, h)
        # and the whole map equals the oracle's highlight over the oracle's matched terms
        parts = [req["search_req"]["search"]] if "search" in req["search_req"] else [p["search"] for p in req["search_req"]["or"]["queries"]]
        terms = {}
        for part in parts:
            terms.setdefault(part["path"] + ".textindex", set()).update(oracle.call("field_search", part=part)["terms"])
        for h in res["data"]:
            assert h["why_found"] == ohl.highlight_on_original_document(columns, h["doc"], terms), req
    with pytest.raises(gpu.VelociGpuError):
        index.get_doc(len(fx.TEST_WHYFOUND_DOCS))
    # requests with `select`: documents rebuilt from the indices, why_found by token ids (test_why_found.rs:177-221, :267-284)
    sel = [
        ({**S("ID1000", "not_tokenized"), "select": ["not_tokenized"]}, {"not_tokenized": "ID1000"}, {"not_tokenized": ["<b>ID1000</b>"]}),
        ({**S("ID1000", "not_tokenized_1_n[]"), "select": ["not_tokenized_1_n[]"]}, {"not_tokenized_1_n": ["ID1000"]}, {"not_tokenized_1_n[]": ["<b>ID1000</b>"]}),
        ({**S("umsortiert", "viele[]", levenshtein_distance=0), "select": ["richtig"]}, {"richtig": "shön"}, {"viele[]": [" ... zu checken, dass da nicht <b>umsortiert</b> wird"]}),
        ({"search_req": {"search": {"terms": ["schön"], "path": "richtig"}}, "select": ["richtig", "viele[]", "nofield"]}, {"richtig": "schön super", "viele": ["nette", "leute"]}, {}),
    ]
    b = index.prepare([json.dumps(r, ensure_ascii=False) for r, _, _ in sel]).execute()
    for q, (req, doc, why) in enumerate(sel):
        res = b.result_docs(q)
        assert res["data"][0]["doc"] == doc and res["data"][0]["why_found"] == why, (req, res)
    assert index.read_doc(2, ["custom_tokenized", "viele[]"]) == {"custom_tokenized": "<<cool>>", "viele": fx.TEST_WHYFOUND_DOCS[2]["viele"]}


def test_lowercase_with_expansion_and_final_sigma(gpu, native_libs):
    """Scoring runs on Rust-lowercased text (search_field.rs:284,312): U+0130 lower-cases to two scalars, a word-final capital
    sigma to U+03C2.  Terms and queries with them, fuzzy and exact, case-folding and not: hits and scores equal the oracle's
    (whose lower-casing is pinned against CPython's in tests/test_lowercase.py)."""
    docs = [{"title": t} for t in ("İstanbul", "istanbul", "ΟΔΟΣ ΑΘΗΝΑΣ", "οδος", "οδοσ", "Straße", "STRASSE", "ΣΟΦΙΑ", "ISTANBUL ΟΔΟΣ", "İİ", "ΑΣ")]
    d = tempfile.mkdtemp(prefix="vb200_r2_lower_")
    helpers.create_index(d, docs, {"title": {"fulltext": {"tokenize": True}}})
    index, oracle = gpu.Index(d), helpers.Oracle(d)
    reqs = []
    for term in ("istanbul", "İstanbul", "i̇stanbul", "ISTANBUL", "οδος", "οδοσ", "ΟΔΟΣ", "αθηνας", "strasse", "straße", "σοφια", "İİ", "ας", "ασ"):
        for lev in (0, 1, 2):
            for extra in ({}, {"ignore_case": False}, {"starts_with": True}):
                reqs.append({"search_req": {"search": {"terms": [term], "path": "title", "levenshtein_distance": lev, **extra}}, "top": 20})
    texts = [json.dumps(r, ensure_ascii=False) for r in reqs]
    b = index.prepare(texts).execute()
    n_hits = 0
    for q, r in enumerate(reqs):
        try:
            c = oracle.search(texts[q])
        except helpers.OracleError:
            assert b.status(q) != 0, r
            continue
        if b.status(q) == 8:  # case-sensitive search for a term whose lower-case form has another length: declared outside the path
            assert r["search_req"]["search"].get("ignore_case") is False and "İ" in r["search_req"]["search"]["terms"][0], (r, b.message(q))
            continue
        assert b.status(q) == 0, (r, b.message(q))
        g = b.result(q)
        assert g["num_hits"] == c["num_hits"], r
        ok, why = helpers.same_topk([(i, float(s)) for i, s in g["data"]], [(h[0], float(np.float32(h[1]))) for h in c["data"]])
        assert ok, (r, why)
        n_hits += g["num_hits"]
    assert n_hits > 100


def test_large_top_and_skip(gpu, sharded_corpus):
    """top + skip beyond 256 (the reference sorts any `top`, src/search.rs:210-218): up to 4096 keys per request the heap is
    merged in global memory on the tile path.  Ids, scores and counts against the oracle; mixed with small-k requests in one
    batch; refused above the limit and on sharded handles."""
    d, reqs = sharded_corpus
    index, oracle = gpu.Index(d), helpers.Oracle(d)
    picks = []
    for i, (top, skip) in enumerate([(1000, 0), (300, 700), (4096, 0), (10, 4000), (257, 0), (10, 0), (2000, 100), (64, 0)]):
        r = json.loads(reqs[i * 7 if i % 2 == 0 else 400 + i])  # or3 requests and `and` requests (with facets)
        r["top"], r["skip"] = top, skip
        picks.append(r)
    single = {"search_req": {"search": {"terms": ["ab"], "path": "body", "starts_with": True}}, "top": 3000}
    picks.append(single)
    # more skip_when_score values than the old limit of four (boost.rs:283-377 takes any number)
    picks.append({"search_req": {"search": {"terms": ["ab"], "path": "body", "starts_with": True}}, "top": 50,
                  "boost": [{"path": "commonness", "boost_fun": "Multiply", "skip_when_score": [1.5, 2.5, 3.0, 4.0, 5.0, 6.5, 7.0]}]})
    texts = [json.dumps(r) for r in picks]
    b = index.prepare(texts).execute()
    total = 0
    for q, r in enumerate(picks):
        assert b.status(q) == 0, (r["top"], b.message(q))
        g, c = b.result(q), oracle.search(texts[q])
        assert g["num_hits"] == c["num_hits"], (q, r["top"], r["skip"])
        assert len(g["data"]) == len(c["data"]), (q, r["top"], r["skip"], len(g["data"]), len(c["data"]))
        ok, why = helpers.same_topk([(i, float(s)) for i, s in g["data"]], [(h[0], float(np.float32(h[1]))) for h in c["data"]])
        assert ok, (q, r["top"], r["skip"], why)
        total += len(g["data"])
    assert total > 3000
    too_many = dict(picks[0], top=4097)
    b = index.prepare([json.dumps(too_many)]).execute()
    assert b.status(0) == 8 and "4096" in b.message(0)
    shard = gpu.Index(d, shard_rank=0, n_shards=2)
    b = shard.prepare([texts[0]])
    assert b.status(0) == 8 and "sharded" in b.message(0)
    # the top_n step takes the same sizes
    rng = np.random.default_rng(2)
    hits = [(int(i), float(np.float32(rng.random() * 9))) for i in rng.choice(SHARDED["num_docs"], size=6000, replace=False)]
    got = index.top_n(hits, 1500, 20)
    ref = sorted(hits, key=lambda h: (-h[1], -h[0]))[20:1520]
    assert [g[0] for g in got] == [r[0] for r in ref]


def test_facets_with_many_groups(gpu, native_libs):
    """A facet field with more groups than the old 1024 limit (src/facet.rs:31-73 takes any number): `top: null` returns them
    all, `top: 2000` the first 2000 in (count desc, value id asc) order."""
    rng = np.random.default_rng(6)
    docs = [{"body": "hit" if i % 3 else "miss", "tags": ["t%04d" % int(x) for x in rng.integers(0, 3000, size=3)]} for i in range(9000)]
    d = tempfile.mkdtemp(prefix="vb200_r2_facets_")
    helpers.create_index(d, docs, {"*GLOBAL*": {"features": ["All"]}, "tags[]": {"facet": True}})
    index, oracle = gpu.Index(d), helpers.Oracle(d)
    reqs = [{"search_req": {"search": {"terms": ["hit"], "path": "body"}}, "facets": [{"field": "tags[]", "top": top}]} for top in (None, 2000, 1025, 5)]
    texts = [json.dumps(r) for r in reqs]
    b = index.prepare(texts).execute()
    for q, r in enumerate(reqs):
        assert b.status(q) == 0, b.message(q)
        got = b.result(q)["facets"]["tags[]"]
        ref = oracle.search(texts[q])["facets"]["tags[]"]
        assert len(got) == len(ref) and len(got) == (r["facets"][0]["top"] or len(ref))
        assert [(g[0], g[1]) for g in got] == [(x[0], x[1]) for x in ref], q
    assert len(b.result(0)["facets"]["tags[]"]) > 2500


def test_boost_and_phrase_step_symbols(gpu, native_libs):
    """The remaining PlanStep kinds as step entry points (plan_steps.rs:174-217, 260-293): BoostToAnchor, ApplyAnchorBoost,
    PlanStepPhrasePairToAnchorId, BoostAnchorFromPhraseResults -- each against the oracle's restatement of the step, and chained
    (FieldSearch -> ResolveTokenIdToAnchor / BoostToAnchor -> ApplyAnchorBoost) against the whole request."""
    import ref_fixtures as fx

    rng = np.random.default_rng(11)
    syll = ["ka", "ki", "ku", "mi", "mo", "ra", "ri", "ru", "sa", "to"]
    words = ["".join(rng.choice(syll, size=int(rng.integers(2, 4)))) for _ in range(120)]
    docs = []
    for i in range(1500):
        n = int(rng.integers(0, 4))
        d = {"ent_seq": str(i), "title": " ".join(str(w) for w in rng.choice(words[:25], size=4))}
        if n:  # several boosted values per document: the position-dependent case of apply_boost_values_anchor
            d["kana"] = [{"text": str(rng.choice(words)), **({"commonness": int(rng.integers(1, 900))} if rng.random() < 0.75 else {})} for _ in range(n)]
        docs.append(d)
    d = tempfile.mkdtemp(prefix="vb200_r2_steps_")
    helpers.create_index(d, docs, {"kana[].text": {"fulltext": {"tokenize": True}}, "kana[].commonness": dict(fx.BOOST),
                                   "title": {"features": ["Search", "PhraseBoost", "BoostTextLocality"], "fulltext": {"tokenize": True}}})
    index, oracle = gpu.Index(d), helpers.Oracle(d)
    close = lambda a, b: abs(float(a) - float(b)) <= 1e-5 * max(abs(float(b)), 1e-30)
    n_boosted = n_multi = 0
    for term in [str(w) for w in rng.choice(words, size=10)] + ["mi", "ka", "ri"]:
        part = {"terms": [term], "path": "kana[].text", "levenshtein_distance": 1, "starts_with": True}
        term_hits, _ = index.field_search(part)
        hits, _ = index.resolve_to_anchor(part, term_hits)
        for fun in ("Log10", "Multiply", "Add"):
            boost = {"path": "kana[].commonness", "boost_fun": fun, "param": 1}
            # BoostToAnchor
            got_ids = index.boost_to_anchor(part, boost, term_hits=term_hits)
            ref_ids = oracle.call("boost_to_anchor", part=part, boost=boost, hits_scores=[[i, float(s)] for i, s in term_hits])
            assert [g[0] for g in got_ids] == [r[0] for r in ref_ids], (term, fun)
            assert all(float(g[1]) == float(np.float32(r[1])) for g, r in zip(got_ids, ref_ids))
            # ApplyAnchorBoost
            got = index.apply_anchor_boost(boost, hits, got_ids)
            ref = oracle.call("apply_boost_values_anchor", hits_scores=[[i, float(s)] for i, s in hits], boost=boost, boost_ids=[[i, float(v)] for i, v in got_ids])
            assert [g[0] for g in got] == [r[0] for r in ref], (term, fun)
            assert all(close(g[1], r[1]) for g, r in zip(got, ref)), (term, fun)
            n_boosted += len(got_ids)
            n_multi += len(got_ids) - len({g[0] for g in got_ids})
        # the chain equals the request with the boost on the part's own 1:n level
        req = {"search_req": {"search": part}, "boost": [{"path": "kana[].commonness", "boost_fun": "Log10", "param": 1}], "top": 4000}
        whole = oracle.search(json.dumps(req))
        chained = index.apply_anchor_boost({"path": "kana[].commonness", "boost_fun": "Log10", "param": 1}, hits, index.boost_to_anchor(part, req["boost"][0], term_hits=term_hits))
        by_id = {h[0]: h[1] for h in whole["data"]}
        assert len(chained) == whole["num_hits"] and all(close(s, by_id[i]) for i, s in chained), term
    assert n_boosted > 300 and n_multi > 20
    assert index.apply_anchor_boost({"path": "kana[].commonness", "boost_fun": "Add"}, [(3, 1.0), (9, 2.0)], []) == [(3, np.float32(1.0)), (9, np.float32(2.0))]
    with pytest.raises(gpu.VelociGpuError) as e:
        index.apply_anchor_boost({"path": "kana[].commonness", "boost_fun": "Add"}, [(3, 1.0)], [(4, 2.0)])
    assert e.value.status == 8

    # PlanStepPhrasePairToAnchorId, BoostAnchorFromPhraseResults
    n_pairs = 0
    phrase_results, phrase_names = [], []
    for w1, w2 in [(str(a), str(b)) for a, b in zip(rng.choice(words[:25], size=12), rng.choice(words[:25], size=12))]:
        p1 = {"terms": [w1], "path": "title", "levenshtein_distance": 1}
        p2 = {"terms": [w2], "path": "title", "levenshtein_distance": 1}
        ids1 = [i for i, _ in index.field_search(p1)[0]]
        ids2 = [i for i, _ in index.field_search(p2)[0]]
        got = index.phrase_pairs_to_anchor("title", ids1, ids2)
        assert got == oracle.call("get_anchor_for_phrases_in_field", path="title", ids1=ids1, ids2=ids2), (w1, w2)
        n_pairs += len(got)
        phrase_results.append(got), phrase_names.append([w1, w2])
    assert n_pairs > 50
    assert index.phrase_pairs_to_anchor("title.textindex", [], [1, 2]) == []
    with pytest.raises(gpu.VelociGpuError) as e:
        index.phrase_pairs_to_anchor("no_such_field", [1], [2])
    assert e.value.status == 3
    hits = [(int(i), float(np.float32(rng.random() * 9 + 0.5))) for i in sorted(rng.choice(1500, size=700, replace=False))]
    names = sorted({tuple(n) for n in phrase_names})
    groups = [names.index(tuple(n)) for n in phrase_names]
    got = index.boost_anchor_from_phrase_results(hits, phrase_results, groups)
    ref = oracle.call("boost_anchor_from_phrase_results", hits_scores=[[i, s] for i, s in hits], boosts=[{"hits_ids": r, "phrase": n} for r, n in zip(phrase_results, phrase_names)])
    assert [g[0] for g in got] == [r[0] for r in ref] and all(close(g[1], r[1]) for g, r in zip(got, ref))
    assert sum(1 for (i, s), g in zip(hits, got) if not close(g[1], s)) > 10  # some hits were boosted
    # the same phrase twice is one boost (merged and deduplicated), two phrases hitting one anchor are two
    twice = index.boost_anchor_from_phrase_results([(5, 2.0), (7, 1.0)], [[5], [5, 7], [5]], [0, 0, 1])
    assert twice == [(5, np.float32(50.0)), (7, np.float32(5.0))]

    # boost_text_locality as a step: the tokens several query terms matched in one text
    n_loc = 0
    for ws in ([words[0], words[1]], [words[2], words[3], words[4]], [words[5], words[5]], [words[6]], [words[7], words[8], words[9], words[10]]):
        token_ids = [[i for i, _ in index.field_search({"terms": [str(w)], "path": "title", "levenshtein_distance": 1})[0]] for w in ws]
        got = index.text_locality("title", token_ids)
        ref = oracle.call("boost_text_locality", path="title", terms={"t%d" % k: ids for k, ids in enumerate(token_ids)})
        assert [(g[0], float(g[1])) for g in got] == [(r[0], float(r[1])) for r in ref], ws
        n_loc += len(got)
    assert n_loc > 20


def test_explain(gpu, native_libs):
    """`explain` (result/explain.rs): vgpu_batch_explain and the "explain" of vgpu_batch_result_docs against the
    oracle's explain maps -- the reference's two explain tests (tests.rs:347-391), every shape of tests/test_explain.py on
    the reference corpus and on a 20k-term dictionary with token values.  The walk itself is also checked on the CPU
    (tests/test_explain.py); this test adds the device match of the bare parts and posting_lookup_kernel."""
    import ref_fixtures as fx
    from test_explain import UNSUPPORTED_REQUESTS, large_dictionary_requests, reference_corpus_requests
    from test_part_hits import make_valued_index

    d = tempfile.mkdtemp(prefix="vb200_gpu_explain_")
    helpers.create_index(d, fx.TEST_ALL_DOCS, fx.TEST_ALL_CONFIG)
    helpers.add_token_values(d, *fx.TEST_ALL_TOKEN_VALUES)
    index, oracle = gpu.Index(d), helpers.Oracle(d)

    def check(index, oracle, requests):
        texts = [json.dumps(r, ensure_ascii=False) for r in requests]
        b = index.prepare(texts).execute()
        n_items = 0
        for q, r in enumerate(requests):
            assert b.status(q) == 0, (texts[q], b.message(q))
            want_res = oracle.search(r)
            got_ids, want_ids = [h[0] for h in b.result(q)["data"]], [h[0] for h in want_res["data"]]
            common = set(got_ids) & set(want_ids)  # (hits that tie at the end of the top-k may differ: helpers.same_topk)
            assert len(got_ids) == len(want_ids) and len(common) + 2 >= len(want_ids), texts[q]
            got = b.explain(q)
            assert set(got) == set(got_ids), texts[q]
            for a in common:  # an anchor nothing was recorded for has no entry in the reference's map
                want = want_res.get("explain", {}).get(str(a), [])
                assert helpers.same_explain(got[a], want), (texts[q], a, got[a], want)
                n_items += len(want)
        return b, n_items

    requests = reference_corpus_requests()
    b, n_items = check(index, oracle, requests)
    assert n_items > 60
    assert len(b.explain(0)[b.result(0)["data"][0][0]]) == 2  # tests.rs:363
    assert len(b.explain(1)[b.result(1)["data"][0][0]]) == 5  # tests.rs:390
    # DocWithHit::explain (search.rs:86,96)
    docs = b.result_docs(1)["data"]
    assert [len(h["explain"]) for h in docs] == [5, 5] and docs[0]["doc"]["ent_seq"] == "1587690"
    assert helpers.same_explain({h["hit"]["id"]: h["explain"] for h in docs}, b.explain(1))
    plain = index.prepare([json.dumps({"search_req": {"search": {"terms": ["urge"], "path": "meanings.eng[]"}}})]).execute()
    assert "explain" not in plain.result_docs(0)["data"][0]
    with pytest.raises(gpu.VelociGpuError) as e:
        plain.explain(0)
    assert e.value.status == 8
    # outside the reconstruction: the search answers, the explanation is refused
    unsupported = UNSUPPORTED_REQUESTS[:1]  # the 1:n boost (the phrase boost case needs another corpus: CPU test only)
    bu = index.prepare([json.dumps(r, ensure_ascii=False) for r in unsupported]).execute()
    for q in range(len(unsupported)):
        assert bu.status(q) == 0, bu.message(q)
        with pytest.raises(gpu.VelociGpuError) as e:
            bu.explain(q)
        assert e.value.status == 8, (q, e.value)
        assert all(h["explain"] is None for h in bu.result_docs(q)["data"])
    # an imported plan does not carry the request; a shard does not hold every anchor's postings
    imported = index.prepare(len(requests), plan=b.export_plan()).execute()
    assert imported.result(1) == b.result(1)
    with pytest.raises(gpu.VelociGpuError) as e:
        imported.explain(1)
    assert e.value.status == 8

    d, oracle, words, _ = make_valued_index()
    index = gpu.Index(d)
    large = list(large_dictionary_requests(words))
    _, n_items = check(index, oracle, large)
    assert n_items > 500
    shard = gpu.Index(d, shard_rank=0, n_shards=2)
    bs = shard.prepare([json.dumps(large[0])]).execute()
    assert bs.status(0) == 0
    with pytest.raises(gpu.VelociGpuError) as e:
        bs.explain(0)
    assert e.value.status == 8


def test_device_resident_step_chain(gpu, sharded_corpus):
    """the step seam over hit lists that stay on the device (vgpu_dev_*).  FieldSearch's term hits ->
    ResolveTokenIdToAnchor per part -> Union / Intersect -> BoostPlanStepFromBoostRequest -> top_n, every intermediate
    list a handle: each step against the same step over host lists (ids and score bits identical), the chain's top-k
    against the whole request through vgpu_batch_execute and against the oracle."""
    from test_gpu_parity import assert_same_topk

    d, _ = sharded_corpus
    index, oracle = gpu.Index(d), helpers.Oracle(d)
    bits = lambda hits: [(int(a), int(np.float32(s).view(np.uint32))) for a, s in hits]
    boost = {"path": "commonness", "boost_fun": "Log10", "param": 1}
    reqs = helpers.synthetic_requests(num_queries=12, query_kind="or3", levenshtein=1, query_seed=61, **SHARDED)
    reqs += helpers.synthetic_requests(num_queries=12, query_kind="and", levenshtein=1, query_seed=62, **SHARDED)
    n_hits_seen = 0
    for text in reqs:
        r = json.loads(text)
        is_or = "or" in r["search_req"]
        parts = [q["search"] for q in r["search_req"]["or" if is_or else "and"]["queries"]]
        dev_lists, host_lists = [], []
        for part in parts:
            term_hits, _ = index.field_search(part)
            dl = index.dev_resolve_to_anchor(part, term_hits)
            hl, _ = index.resolve_to_anchor(part, term_hits)
            assert len(dl) == len(hl) and bits(dl.download()) == bits(hl), part
            dev_lists.append(dl), host_lists.append(hl)
        terms = [p["terms"][0] for p in parts]
        if is_or:
            merged = index.dev_union_hits_score(dev_lists, terms)
            assert bits(merged.download()) == bits(index.union_hits_score(host_lists, terms)), text
        else:
            merged = index.dev_intersect_hits_score(dev_lists)
            assert bits(merged.download()) == bits(index.intersect_hits_score(host_lists)), text
        n_hits_seen += len(merged)
        if "boost" in r:
            boosted = index.dev_add_boost(r["boost"][0], merged)
            assert bits(boosted.download()) == bits(index.add_boost(r["boost"][0], merged.download())), text
            merged = boosted
        top = index.dev_top_n(merged, 10)
        whole = index.search(json.dumps({k: v for k, v in r.items() if k != "facets"}))
        assert_same_topk(top, whole["data"], ctx=text)
        ref = oracle.search({k: v for k, v in r.items() if k != "facets"})
        assert len(merged) == ref["num_hits"] == whole["num_hits"], text
        assert_same_topk(top, [(h[0], np.float32(h[1])) for h in ref["data"]], ctx=text)
        assert bits(index.dev_top_n(merged, 4, 3)) == bits(index.top_n(merged.download(), 4, 3)), text
    assert n_hits_seen > 10000
    # host lists in and out: a repeated anchor keeps its largest score, order is by anchor id, negative scores survive
    ups = [(70, 2.5), (3, 1.0), (70, 4.0), (119999, 0.0), (5, -3.0)]
    up = index.dev_upload(ups)
    assert up.download() == [(3, 1.0), (5, -3.0), (70, 4.0), (119999, 0.0)] and len(up) == 4
    other = index.dev_upload([(5, 2.0), (9, 1.0)])
    assert bits(index.dev_union_hits_score([up, other], ["a", "b"]).download()) == bits(index.union_hits_score([up.download(), other.download()], ["a", "b"]))
    assert bits(index.dev_intersect_hits_score([up, other]).download()) == bits(index.intersect_hits_score([up.download(), other.download()]))
    assert index.dev_union_hits_score([], []).download() == [] and len(index.dev_upload([])) == 0
    assert bits(index.dev_union_hits_score([other], ["a"]).download()) == bits(other.download())  # one input passes through
    # a handle belongs to its index; a shard's lists hold the shard's anchors
    second = gpu.Index(d, shard_rank=1, n_shards=2)
    with pytest.raises(gpu.VelociGpuError) as e:
        second.dev_union_hits_score([up, other], ["a", "b"])
    assert e.value.status == 1
    part = json.loads(reqs[0])["search_req"]["or"]["queries"][0]["search"]
    term_hits, _ = index.field_search(part)
    first = gpu.Index(d, shard_rank=0, n_shards=2)
    halves = [s.dev_resolve_to_anchor(part, term_hits).download() for s in (first, second)]
    assert bits(halves[0] + halves[1]) == bits(index.resolve_to_anchor(part, term_hits)[0])
    assert all(a < SHARDED["num_docs"] // 2 for a, _ in halves[0]) and all(a >= SHARDED["num_docs"] // 2 for a, _ in halves[1])


def test_code_search_on_the_gpu(gpu, native_libs):
    """tests/all/test_code_search.rs on the device: regex parts given directly and `*` patterns through the query generator
    (four-field `or`s of regex parts), on the reference's one-line corpus; hits, scores and the document against the oracle
    (tests/test_code_search.py holds the same requests against the oracle alone)."""
    from test_code_search import API_CASES, GENERATOR_CASES, LINE, generated_request, make_code_index
    from test_gpu_parity import assert_same_topk

    d, cat = make_code_index()
    index, oracle = gpu.Index(d), helpers.Oracle(d)
    reqs = [r for r, _ in API_CASES] + [generated_request(d, cat, p) for p, _ in GENERATOR_CASES]
    want = [n for _, n in API_CASES] + [n for _, n in GENERATOR_CASES]
    b = index.prepare([json.dumps(r, ensure_ascii=False) for r in reqs]).execute()
    for q, (r, n) in enumerate(zip(reqs, want)):
        assert b.status(q) == 0, (r, b.message(q))
        got, ref = b.result(q), oracle.search(r)
        assert got["num_hits"] == ref["num_hits"] and len(got["data"]) == n, r
        assert_same_topk(got["data"], [(h[0], np.float32(h[1])) for h in ref["data"]], ctx=r)
        if n:
            assert b.result_docs(q)["data"][0]["doc"]["line"] == LINE


def test_field_highlight_on_the_gpu(gpu, native_libs):
    """search_field::highlight through the C ABI (vgpu_highlight): the part's terms matched and scored on the device, texts and
    snippets against the Python oracle (tests/test_field_highlight.py: its parts, the reference's tests.rs:1009-1045 first)."""
    from test_field_highlight import PARTS, STORY, make_index, oracle_highlight, same

    d = make_index()
    index, oracle = gpu.Index(d), helpers.Oracle(d)
    assert [t for t, _, _ in index.highlight(PARTS[0])] == [STORY] and [t for t, _, _ in index.highlight(PARTS[1])] == [STORY]
    for part in PARTS:
        got = [(t, float(s), i) for t, s, i in index.highlight(part)]
        assert same(got, oracle_highlight(d, oracle, part)), (part, got)
    for part in ({"terms": ["story"], "path": "mylongtext"}, {"terms": ["1587690"], "path": "nofulltext", "snippet": True}):
        with pytest.raises(gpu.VelociGpuError) as e:
            index.highlight(part)
        assert e.value.status == 1
    with pytest.raises(gpu.VelociGpuError) as e:
        index.highlight({"terms": ["story"], "path": "nope", "snippet": True})
    assert e.value.status == 2  # FstNotFound
