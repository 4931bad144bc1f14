#!/usr/bin/env python
"""One rank of a sharded search (launched by torchrun, one process per GPU): opens shard RANK of WORLD_SIZE of an index,
joins the library's communicator, receives the plans of local rank 0 through the shared-memory channel and executes the
batch as one collective call per step.  Every rank then holds the complete result; each compares it with the unsharded
index opened on its own GPU, and rank 0 also with the CPU oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/sharded_worker.py INDEX_DIR REQUESTS.jsonl
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np
    import torch
    import torch.distributed as dist

    import helpers
    import veloci_b200

    d, req_path = sys.argv[1], sys.argv[2]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    reqs = [l for l in open(req_path).read().split("\n") if l]

    def bcast(payload, n):
        t = torch.zeros(n, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
        dist.broadcast(t, src=0)
        return bytes(t.cpu().numpy().tobytes())

    shard = veloci_b200.Index(d, device=local, shard_rank=rank, n_shards=world)
    shard.comm_init(bcast(veloci_b200.comm_unique_id() if rank == 0 else None, 128))
    ch = veloci_b200.PlanChannel("/vb200_worker_" + bcast(os.urandom(8) if rank == 0 else None, 8).hex(), rank, world, capacity=64 << 20)
    # resident batch, executed three times (collective), then the stream path with three batches
    batch = shard.prepare(reqs if rank == 0 else [], channel=ch)
    for _ in range(3):
        batch.execute()
    got = batch.results_flat(10)
    streamed = list(shard.search_stream((reqs if rank == 0 else [] for _ in range(3)), k=10, channel=ch))
    whole_index = veloci_b200.Index(d, device=local)
    whole_batch = whole_index.prepare(reqs).execute()
    whole = whole_batch.results_flat(10)
    for name, r in [("resident", got)] + [(f"stream{i}", s) for i, s in enumerate(streamed)]:
        for key in ("status", "num_hits", "ids"):
            assert (r[key] == whole[key]).all(), (rank, name, key)
        assert (r["scores"].view(np.uint32) == whole["scores"].view(np.uint32)).all(), (rank, name)
    for q in range(0, len(reqs), 13):
        if whole["status"][q] == 0:
            assert batch.result(q).get("facets") == whole_batch.result(q).get("facets"), (rank, q)
    if rank == 0:
        rows = list(range(0, len(reqs), 7))
        ref = helpers.Oracle(d).search_batch([reqs[q] for q in rows], threads=4, k=10)
        par = helpers.batch_parity(got, ref, rows)
        assert par["equal"] == par["checked"], par
    dist.barrier()
    ch.close()
    shard.comm_destroy()
    if rank == 0:
        print(f"sharded_worker ok: {world} ranks, {len(reqs)} requests, phases {batch.phase_ms()}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
