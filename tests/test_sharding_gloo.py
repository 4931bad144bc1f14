"""The N > 1 exchange protocol on CPU: two gloo ranks each hold the hits of their anchor range
(cut from the oracle's full result), build shard-local top-k rows, all-gather them and merge;
the result must equal the unsharded top-k and hit count.  (The device merge kernel itself is
covered by the GPU test test_sharded_index_equals_unsharded.)"""
import json
import os
import socket
import tempfile

import numpy as np
import pytest

import helpers
from veloci_b200 import sharding

WORLD = 2
K = 10
PARAMS = dict(num_docs=6000, vocab=800, seed=3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, port, index_dir, requests, out_dir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    oracle = helpers.Oracle(index_dir)
    lo, hi = sharding.shard_range(PARAMS["num_docs"], rank, WORLD)
    n = len(requests)
    rows = np.zeros((n, K), dtype=np.uint64)
    hits = np.zeros(n, dtype=np.int64)
    for q, r in enumerate(requests):
        req = json.loads(r)
        req["top"] = 1 << 20  # every hit, so that the shard's part can be cut out
        full = oracle.search(req)["data"]
        mine = [(h[0], np.uint32(h[2]).view(np.float32)) for h in full if lo <= h[0] < hi]
        hits[q] = len(mine)
        rows[q] = sharding.local_rows([m[0] for m in mine], [m[1] for m in mine], K, K)
    g_rows = [torch.empty(n * K, dtype=torch.int64) for _ in range(WORLD)]
    g_hits = [torch.empty(n, dtype=torch.int64) for _ in range(WORLD)]
    dist.all_gather(g_rows, torch.from_numpy(rows.view(np.int64).reshape(-1).copy()))
    dist.all_gather(g_hits, torch.from_numpy(hits))
    keys = np.stack([t.numpy().view(np.uint64).reshape(n, K) for t in g_rows])
    merged, total = sharding.merge_gathered_host(keys, np.stack([t.numpy() for t in g_hits]), K)
    np.save(os.path.join(out_dir, f"keys_{rank}.npy"), merged)
    np.save(os.path.join(out_dir, f"hits_{rank}.npy"), total)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition_the_anchors():
    for docs in (1, 7, 1000, 10_000_019):
        for world in (1, 2, 3, 8):
            edges = [sharding.shard_range(docs, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == docs
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))


def test_key_packing_orders_like_the_reference():
    rng = np.random.default_rng(1)
    scores = np.concatenate([rng.normal(size=200).astype(np.float32), np.float32([0.0, -0.0, 1.5, 1.5, 1.5])])
    ids = rng.permutation(len(scores)).astype(np.uint32)
    keys = sharding.pack_keys(ids, scores)
    order = np.argsort(keys)[::-1]
    ref = sorted(range(len(scores)), key=lambda i: (-float(scores[i]), -int(ids[i])))  # score desc, id desc (search.rs:123-130)
    same_score_as_ref = [float(scores[i]) for i in order] == [float(scores[i]) for i in ref]
    assert same_score_as_ref
    got_ids, got_scores = sharding.unpack_keys(keys)
    assert (got_ids == ids).all() and (got_scores.view(np.uint32) == scores.view(np.uint32))[scores != 0].all()


def test_two_rank_gloo_exchange_equals_unsharded(native_libs):
    import torch.multiprocessing as mp

    d = tempfile.mkdtemp(prefix="vb200_gloo_")
    helpers.create_synthetic_index(d, **PARAMS)
    reqs = helpers.synthetic_requests(num_queries=24, query_kind="or3", levenshtein=1, query_seed=5, **PARAMS)
    out_dir = tempfile.mkdtemp(prefix="vb200_gloo_out_")
    mp.spawn(_rank_main, args=(_free_port(), d, reqs, out_dir), nprocs=WORLD, join=True)
    oracle = helpers.Oracle(d)
    ref = oracle.search_batch(reqs, threads=2, k=K)
    for rank in range(WORLD):
        keys = np.load(os.path.join(out_dir, f"keys_{rank}.npy"))
        hits = np.load(os.path.join(out_dir, f"hits_{rank}.npy"))
        assert (hits == ref["num_hits"].astype(np.int64)).all()
        ids, scores = sharding.unpack_keys(keys)
        for q in range(len(reqs)):
            n = int(min(K, ref["num_hits"][q]))
            assert (scores[q, :n].view(np.uint32) == ref["scores"][q, :n].view(np.uint32)).all()
            assert (ids[q, :n] == ref["ids"][q, :n]).all()
