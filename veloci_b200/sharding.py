"""Anchor-range sharding: the host-side arithmetic of the multi-GPU path.

Every rank owns the anchors [rank * num_docs / world, (rank + 1) * num_docs / world) of every
anchor-keyed structure (the same split vgpu_index_open applies), evaluates the whole batch on its
shard and produces, per request, `stride` 64-bit keys sorted descending plus its local num_hits.
One all-gather later every rank holds [world][n][stride] keys; because the order
(score desc, id desc) is total and the shards are disjoint, the global top-k is the top-k of the
union of the local top-k rows and num_hits is the sum.

The product path merges on the device (merge_heaps_kernel through vgpu_batch_merge_gathered);
`merge_gathered_host` is the numpy statement of the same merge, used by the CPU tests of the
exchange protocol.
"""
import numpy as np


def shard_range(num_docs, rank, world):
    """Anchor range of `rank`, identical to DeviceIndex::open (anchor_lo / anchor_hi)."""
    return num_docs * rank // world, num_docs * (rank + 1) // world


def pack_keys(ids, scores):
    """(orderable f32 score << 32) | anchor id: larger key = better hit under (score desc, id desc)."""
    bits = np.asarray(scores, dtype=np.float32).view(np.uint32).astype(np.uint64)
    neg = (bits & np.uint64(0x80000000)) != 0
    key = np.where(neg, (~bits) & np.uint64(0xFFFFFFFF), bits | np.uint64(0x80000000))
    key = np.where(key == 0, np.uint64(1), key)
    return (key << np.uint64(32)) | np.asarray(ids, dtype=np.uint64)


def unpack_keys(keys):
    keys = np.asarray(keys, dtype=np.uint64)
    k = (keys >> np.uint64(32)).astype(np.uint32)
    neg = (k & np.uint32(0x80000000)) == 0
    bits = np.where(neg, ~k, k & np.uint32(0x7FFFFFFF)).astype(np.uint32)
    return (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32), bits.view(np.float32)


def local_rows(ids, scores, k, stride):
    """One request's shard-local row: its best k hits as keys, descending, zero padded to `stride`."""
    keys = np.sort(pack_keys(ids, scores))[::-1][:k]
    row = np.zeros(stride, dtype=np.uint64)
    row[: len(keys)] = keys
    return row


def merge_gathered_host(gathered_keys, gathered_hits, k):
    """gathered_keys [world][n][stride], gathered_hits [world][n] -> (keys [n][stride], num_hits [n])."""
    world, n, stride = gathered_keys.shape
    flat = np.transpose(gathered_keys, (1, 0, 2)).reshape(n, world * stride)
    order = np.sort(flat, axis=1)[:, ::-1]
    out = np.zeros((n, stride), dtype=np.uint64)
    kk = min(k, stride)
    out[:, :kk] = order[:, :kk]
    return out, gathered_hits.sum(axis=0)
