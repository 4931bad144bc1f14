"""ctypes binding of include/veloci_b200.h (the calls a Rust `-sys` crate would make)."""
import ctypes
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

STATUS_NAMES = {
    0: "ok", 1: "InvalidRequest", 2: "FieldNotFound", 3: "PathNotFound", 4: "Io", 5: "Json",
    6: "Cuda", 7: "Nccl", 8: "Unsupported", 9: "Internal",
}


class VelociGpuError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"[{STATUS_NAMES.get(status, status)}] {message}")
        self.status = status
        self.message = message


class _Hit(ctypes.Structure):
    _fields_ = [("id", ctypes.c_uint32), ("score", ctypes.c_float)]


class _HitList(ctypes.Structure):
    _fields_ = [("hits", ctypes.POINTER(_Hit)), ("n_hits", ctypes.c_uint32), ("ids", ctypes.POINTER(ctypes.c_uint32)), ("n_ids", ctypes.c_uint32)]


class _Suggestion(ctypes.Structure):
    _fields_ = [("text", ctypes.c_char_p), ("score", ctypes.c_float), ("id", ctypes.c_uint32)]


class _Suggestions(ctypes.Structure):
    _fields_ = [("items", ctypes.POINTER(_Suggestion)), ("n", ctypes.c_uint32), ("text_block", ctypes.c_void_p)]


def lib_path():
    return os.path.join(_HERE, "lib", "libveloci_b200.so")


def load_library():
    """Loads the CUDA library; raises if it has not been built (there is no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise VelociGpuError(6, f"{path} is missing: build it with `python -m veloci_b200.build` (nvcc, sm_100a)")
    L = ctypes.CDLL(path)
    vp, u32, u64, i32, cp = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int32, ctypes.c_char_p
    P = ctypes.POINTER
    L.vgpu_last_error.restype = cp
    L.vgpu_device_count.restype = i32
    L.vgpu_index_open.argtypes = [cp, i32, u32, u32, P(vp)]
    L.vgpu_index_open_ex.argtypes = [cp, i32, u32, u32, u32, P(vp)]
    L.vgpu_comm_unique_id.argtypes = [vp]
    L.vgpu_comm_init.argtypes = [vp, vp]
    L.vgpu_comm_destroy.argtypes = [vp]
    L.vgpu_batch_export_plan.argtypes = [vp, P(vp), P(ctypes.c_size_t)]
    L.vgpu_batch_prepare_from_plan.argtypes = [vp, vp, ctypes.c_size_t, P(vp)]
    L.vgpu_plan_channel_open.argtypes = [cp, u32, u32, ctypes.c_size_t, P(vp)]
    L.vgpu_plan_channel_close.argtypes = [vp]
    L.vgpu_plan_channel_close.restype = None
    L.vgpu_batch_prepare_shared.argtypes = [vp, vp, u64, ctypes.c_char_p, ctypes.c_size_t, u32, P(vp)]
    L.vgpu_plan_channel_ticket.argtypes = [vp]
    L.vgpu_plan_channel_ticket.restype = u64
    L.vgpu_index_close.argtypes = [vp]
    L.vgpu_index_close.restype = None
    L.vgpu_index_info.argtypes = [vp, P(u64), P(u64), P(u64), P(u64)]
    L.vgpu_batch_prepare.argtypes = [vp, P(cp), u32, P(vp)]
    L.vgpu_batch_prepare_lines.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t, u32, P(vp)]
    L.vgpu_batch_prepare_jsonl.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t, P(u32), P(vp)]
    L.vgpu_batch_execute.argtypes = [vp]
    L.vgpu_batch_execute_begin.argtypes = [vp]
    L.vgpu_batch_execute_finish.argtypes = [vp]
    L.vgpu_batch_thresholds.argtypes = [vp, P(vp), P(u32)]
    L.vgpu_batch_facet_histograms.argtypes = [vp, P(vp), P(u64)]
    L.vgpu_batch_free.argtypes = [vp]
    L.vgpu_batch_free.restype = None
    L.vgpu_batch_size.argtypes = [vp, P(u32)]
    L.vgpu_batch_status.argtypes = [vp, u32]
    L.vgpu_batch_message.argtypes = [vp, u32]
    L.vgpu_batch_message.restype = cp
    L.vgpu_batch_result.argtypes = [vp, u32, P(u64), P(_Hit), u32, P(u32)]
    L.vgpu_batch_results_flat.argtypes = [vp, u32, vp, vp, vp, vp]
    L.vgpu_search_batch.argtypes = [vp, P(cp), u32, u32, vp, vp, vp, vp]
    L.vgpu_batch_facet_count.argtypes = [vp, u32, P(u32)]
    L.vgpu_batch_facet.argtypes = [vp, u32, u32, P(cp), P(u32)]
    L.vgpu_batch_facet_group.argtypes = [vp, u32, u32, u32, P(u32), P(u32), P(cp)]
    L.vgpu_batch_local_topk.argtypes = [vp, P(vp), P(vp), P(u32)]
    L.vgpu_batch_merge_gathered.argtypes = [vp, vp, vp, u32]
    L.vgpu_free.argtypes = [vp]
    L.vgpu_free.restype = None
    L.vgpu_hitlist_free.argtypes = [P(_HitList)]
    L.vgpu_hitlist_free.restype = None
    L.vgpu_field_search.argtypes = [vp, cp, i32, i32, P(_HitList)]
    L.vgpu_resolve_to_anchor.argtypes = [vp, cp, P(_HitList), P(_HitList)]
    L.vgpu_union_hits_score.argtypes = [vp, P(_HitList), P(cp), u32, P(_HitList)]
    L.vgpu_intersect_hits_score.argtypes = [vp, P(_HitList), u32, P(_HitList)]
    L.vgpu_resolve_to_anchor_filtered.argtypes = [vp, cp, P(_HitList), P(u32), u32, P(_HitList)]
    L.vgpu_union_hits_ids.argtypes = [vp, P(_HitList), u32, P(_HitList)]
    L.vgpu_intersect_hits_ids.argtypes = [vp, P(_HitList), u32, P(_HitList)]
    L.vgpu_intersect_scores_with_ids.argtypes = [vp, P(_HitList), P(_HitList), P(_HitList)]
    L.vgpu_phrase_pairs_to_anchor.argtypes = [vp, cp, P(ctypes.c_uint32), ctypes.c_uint32, P(ctypes.c_uint32), ctypes.c_uint32, P(_HitList)]
    L.vgpu_boost_anchor_from_phrase_results.argtypes = [vp, P(_HitList), P(_HitList), P(ctypes.c_uint32), ctypes.c_uint32, P(_HitList)]
    L.vgpu_boost_to_anchor.argtypes = [vp, cp, P(_HitList), cp, P(_HitList)]
    L.vgpu_apply_anchor_boost.argtypes = [vp, cp, P(_HitList), P(_HitList), P(_HitList)]
    L.vgpu_text_locality.argtypes = [vp, cp, P(_HitList), ctypes.c_uint32, P(_HitList)]
    L.vgpu_facet.argtypes = [vp, cp, P(u32), u32, P(_Suggestions)]
    L.vgpu_add_boost.argtypes = [vp, cp, P(_HitList)]
    L.vgpu_top_n.argtypes = [vp, P(_HitList), u32, u32, P(_HitList)]
    L.vgpu_dev_upload.argtypes = [vp, P(_HitList), P(vp)]
    L.vgpu_dev_download.argtypes = [vp, P(_HitList)]
    L.vgpu_dev_len.argtypes = [vp]
    L.vgpu_dev_len.restype = ctypes.c_uint32
    L.vgpu_dev_free.argtypes = [vp]
    L.vgpu_dev_free.restype = None
    L.vgpu_dev_resolve_to_anchor.argtypes = [vp, cp, P(_HitList), P(vp)]
    L.vgpu_dev_union_hits_score.argtypes = [vp, P(vp), P(cp), u32, P(vp)]
    L.vgpu_dev_intersect_hits_score.argtypes = [vp, P(vp), u32, P(vp)]
    L.vgpu_dev_add_boost.argtypes = [vp, cp, vp, P(vp)]
    L.vgpu_dev_top_n.argtypes = [vp, vp, u32, u32, P(_HitList)]
    L.vgpu_suggest.argtypes = [vp, cp, P(_Suggestions)]
    L.vgpu_suggest_part.argtypes = [vp, cp, P(_Suggestions)]
    L.vgpu_highlight.argtypes = [vp, cp, P(_Suggestions)]
    L.vgpu_get_doc.argtypes = [vp, ctypes.c_uint32, P(vp)]
    L.vgpu_batch_result_docs.argtypes = [vp, ctypes.c_uint32, P(vp)]
    L.vgpu_batch_explain.argtypes = [vp, ctypes.c_uint32, P(vp)]
    L.vgpu_explain_plan.argtypes = [cp, P(vp)]
    L.vgpu_read_doc.argtypes = [vp, ctypes.c_uint32, cp, P(vp)]
    L.vgpu_search_query.argtypes = [vp, cp, P(vp)]
    L.vgpu_suggest_query.argtypes = [vp, cp, P(vp)]
    L.vgpu_query_parse.argtypes = [cp, ctypes.c_uint32, P(vp)]
    L.vgpu_suggestions_free.argtypes = [P(_Suggestions)]
    L.vgpu_suggestions_free.restype = None
    L.vgpu_launch_count.restype = u64
    L.vgpu_batch_phase_ms.argtypes = [vp, P(ctypes.c_float), u32]
    L.vgpu_batch_traffic_model.argtypes = [vp, P(u64), P(u64), P(u64), P(u64)]
    L.vgpu_batch_path_stats.argtypes = [vp, P(u64), P(u64), P(u64)]
    L.vgpu_batch_io_bytes.argtypes = [vp, P(u64), P(u64)]
    L.vgpu_batch_set_profiling.argtypes = [vp, i32]
    L.vgpu_batch_kernel_times_json.argtypes = [vp]
    L.vgpu_batch_kernel_times_json.restype = cp
    L.vgpu_batch_work_stats.argtypes = [vp, P(u64), u32]
    _LIB = L
    return L


def _check(rc):
    if rc != 0:
        raise VelociGpuError(rc, load_library().vgpu_last_error().decode("utf-8", "replace"))


def _take_string(L, rc, out):
    _check(rc)
    try:
        return ctypes.string_at(out.value).decode("utf-8")
    finally:
        L.vgpu_free(out)


def query_parse(text, no_attributes=False, no_parentheses=False, no_levensthein=False):
    """query_parser::parse_with_opt: the Debug text of the parsed query tree."""
    L = load_library()
    out = ctypes.c_void_p()
    return _take_string(L, L.vgpu_query_parse(text.encode("utf-8"), int(no_attributes) | int(no_parentheses) << 1 | int(no_levensthein) << 2, ctypes.byref(out)), out)


def explain_plan(request):
    """search::explain_plan: the request's plan steps as a Graphviz dot graph."""
    L = load_library()
    out = ctypes.c_void_p()
    text = request if isinstance(request, (str, bytes)) else json.dumps(request, ensure_ascii=False)
    return _take_string(L, L.vgpu_explain_plan(text.encode("utf-8") if isinstance(text, str) else text, ctypes.byref(out)), out)


def device_count():
    return int(load_library().vgpu_device_count())


def launch_count():
    return int(load_library().vgpu_launch_count())


def _encode_requests(requests):
    enc = [r.encode("utf-8") if isinstance(r, str) else (r if isinstance(r, bytes) else json.dumps(r, ensure_ascii=False).encode("utf-8")) for r in requests]
    return (ctypes.c_char_p * len(enc))(*enc), enc


def _join_lines(requests):
    """The requests as one line-feed separated buffer.  Requests that are already bytes (what a server holds) are joined without
    another pass over their text; str requests are encoded."""
    if requests and isinstance(requests[0], bytes):
        return b"\n".join(requests)
    return "\n".join(requests).encode("utf-8")


def _to_hitlist(hits=None, ids=None):
    hits = list(hits or [])
    ids = list(ids or [])
    hl = _HitList()
    harr = (_Hit * max(1, len(hits)))()
    for i, (a, s) in enumerate(hits):
        harr[i].id, harr[i].score = int(a), float(s)
    iarr = (ctypes.c_uint32 * max(1, len(ids)))(*ids)
    hl.hits, hl.n_hits = ctypes.cast(harr, ctypes.POINTER(_Hit)), len(hits)
    hl.ids, hl.n_ids = ctypes.cast(iarr, ctypes.POINTER(ctypes.c_uint32)), len(ids)
    hl._keep = (harr, iarr)
    return hl


def _from_hitlist(L, hl):
    hits = [(hl.hits[i].id, np.float32(hl.hits[i].score)) for i in range(hl.n_hits)]
    ids = [hl.ids[i] for i in range(hl.n_ids)]
    L.vgpu_hitlist_free(ctypes.byref(hl))
    return hits, ids


class DeviceHitList:
    """hits_scores of a plan step kept on the index's device for the next step (vgpu_hitlist_dev)."""

    def __init__(self, index, handle):
        self.index, self.L, self.h = index, index.L, handle

    def __len__(self):
        return int(self.L.vgpu_dev_len(self.h))

    def download(self):
        """[(anchor id, score)] by ascending anchor id."""
        out = _HitList()
        _check(self.L.vgpu_dev_download(self.h, ctypes.byref(out)))
        return _from_hitlist(self.L, out)[0]

    def close(self):
        if self.h:
            self.L.vgpu_dev_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """A prepared batch of requests (vgpu_batch_prepare / execute / fetch)."""

    def __init__(self, index, requests, plan=None, channel=None, ticket=None):
        """`requests`: the batch.  `plan`: a blob from `export_plan()` of another handle of the same directory (the requests
        are then not parsed again; pass their count or the list).  `channel`: a PlanChannel -- local rank 0 plans and
        publishes, the other local ranks import (collective over the channel's ranks)."""
        self.L = load_library()
        self.index = index
        self.n = requests if isinstance(requests, int) else len(requests)
        self.h = ctypes.c_void_p()
        if plan is not None:
            _check(self.L.vgpu_batch_prepare_from_plan(index.h, plan, len(plan), ctypes.byref(self.h)))
            self.n = self._size()
            return
        if channel is not None:
            blob = b"" if channel.rank != 0 else _join_lines(requests)
            if ticket is None:
                ticket = channel.ticket()
            _check(self.L.vgpu_batch_prepare_shared(index.h, channel.h, ticket, blob if channel.rank == 0 else None, len(blob), self.n if channel.rank == 0 else 0, ctypes.byref(self.h)))
            self.n = self._size()
            return
        if self.n >= 64 and self._prepare_lines(index, requests):
            return
        arr, self._keep = _encode_requests(requests)
        _check(self.L.vgpu_batch_prepare(index.h, arr, self.n, ctypes.byref(self.h)))

    def _size(self):
        n = ctypes.c_uint32()
        _check(self.L.vgpu_batch_size(self.h, ctypes.byref(n)))
        return int(n.value)

    def result_docs(self, q):
        """The hits of request q as documents with why_found highlights (search::to_search_result)."""
        out = ctypes.c_void_p()
        return json.loads(_take_string(self.L, self.L.vgpu_batch_result_docs(self.h, q, ctypes.byref(out)), out))

    def explain(self, q):
        """{anchor id: [Explain, ...]} of the hits request q returns (a request with "explain": true)."""
        out = ctypes.c_void_p()
        text = _take_string(self.L, self.L.vgpu_batch_explain(self.h, q, ctypes.byref(out)), out)
        return {int(k): v for k, v in json.loads(text).items()}

    def export_plan(self):
        """The batch's plan as bytes without process-local addresses (vgpu_batch_export_plan)."""
        blob, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self.L.vgpu_batch_export_plan(self.h, ctypes.byref(blob), ctypes.byref(n)))
        try:
            return ctypes.string_at(blob.value, n.value)
        finally:
            self.L.vgpu_free(blob)

    def _prepare_lines(self, index, requests):
        """Many requests: one buffer of line-feed separated requests instead of one C string each.  False when some request
        is not a string or contains a line feed (the library checks the count of line feeds before doing anything)."""
        try:
            blob = _join_lines(requests)
        except TypeError:
            return False
        rc = self.L.vgpu_batch_prepare_lines(index.h, blob, len(blob), self.n, ctypes.byref(self.h))
        if rc == 1:  # VGPU_ERR_INVALID_REQUEST: not n - 1 line feeds
            return False
        _check(rc)
        return True

    def execute(self):
        _check(self.L.vgpu_batch_execute(self.h))
        return self

    def status(self, q):
        return int(self.L.vgpu_batch_status(self.h, q))

    def message(self, q):
        return self.L.vgpu_batch_message(self.h, q).decode("utf-8", "replace")

    def result(self, q, cap=4096):
        """-> {"num_hits", "data": [(id, score)], "facets": {...}} of request q (raises for a failed request)."""
        st = self.status(q)
        if st != 0:
            raise VelociGpuError(st, self.message(q))
        nh, n = ctypes.c_uint64(), ctypes.c_uint32()
        hits = (_Hit * cap)()
        _check(self.L.vgpu_batch_result(self.h, q, ctypes.byref(nh), hits, cap, ctypes.byref(n)))
        out = {"num_hits": int(nh.value), "data": [(hits[i].id, np.float32(hits[i].score)) for i in range(n.value)]}
        nf = ctypes.c_uint32()
        _check(self.L.vgpu_batch_facet_count(self.h, q, ctypes.byref(nf)))
        if nf.value:
            facets = {}
            for f in range(nf.value):
                name, ng = ctypes.c_char_p(), ctypes.c_uint32()
                _check(self.L.vgpu_batch_facet(self.h, q, f, ctypes.byref(name), ctypes.byref(ng)))
                groups = []
                for g in range(ng.value):
                    vid, cnt, text = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_char_p()
                    _check(self.L.vgpu_batch_facet_group(self.h, q, f, g, ctypes.byref(vid), ctypes.byref(cnt), ctypes.byref(text)))
                    groups.append((text.value.decode("utf-8", "replace"), int(cnt.value), int(vid.value)))
                facets[name.value.decode("utf-8")] = groups
            out["facets"] = facets
        return out

    def results_flat(self, k=10):
        ids = np.zeros((self.n, k), dtype=np.uint32)
        scores = np.zeros((self.n, k), dtype=np.float32)
        num_hits = np.zeros(self.n, dtype=np.uint64)
        status = np.zeros(self.n, dtype=np.int32)
        _check(self.L.vgpu_batch_results_flat(self.h, k, ids.ctypes.data, scores.ctypes.data, num_hits.ctypes.data, status.ctypes.data))
        return {"ids": ids, "scores": scores, "num_hits": num_hits, "status": status}

    def execute_begin(self):
        _check(self.L.vgpu_batch_execute_begin(self.h))
        return self

    def execute_finish(self):
        _check(self.L.vgpu_batch_execute_finish(self.h))
        return self

    def thresholds(self):
        """Device pointer and length of the per-request thresholds (64-bit order keys) after execute_begin."""
        ptr, n = ctypes.c_void_p(), ctypes.c_uint32()
        _check(self.L.vgpu_batch_thresholds(self.h, ctypes.byref(ptr), ctypes.byref(n)))
        return ptr.value, int(n.value)

    def facet_histograms(self):
        """Device pointer and length (u32 words) of the batch's facet histograms (None, 0 without facets)."""
        ptr, n = ctypes.c_void_p(), ctypes.c_uint64()
        _check(self.L.vgpu_batch_facet_histograms(self.h, ctypes.byref(ptr), ctypes.byref(n)))
        return ptr.value, int(n.value)

    def local_topk(self):
        """Device pointers of the shard-local result rows: (keys_ptr, num_hits_ptr, stride)."""
        keys, hits, stride = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_uint32()
        _check(self.L.vgpu_batch_local_topk(self.h, ctypes.byref(keys), ctypes.byref(hits), ctypes.byref(stride)))
        return keys.value, hits.value, int(stride.value)

    def merge_gathered(self, keys_ptr, hits_ptr, n_shards):
        _check(self.L.vgpu_batch_merge_gathered(self.h, ctypes.c_void_p(keys_ptr), ctypes.c_void_p(hits_ptr), n_shards))

    def phase_ms(self):
        ms = (ctypes.c_float * 6)()
        _check(self.L.vgpu_batch_phase_ms(self.h, ms, 6))
        return [float(x) for x in ms]

    def path_stats(self):
        a, b, c = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        _check(self.L.vgpu_batch_path_stats(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"plane_items": a.value, "general_items": b.value, "plane_evaluated": c.value}

    def traffic_model(self):
        a, b, c, d = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        _check(self.L.vgpu_batch_traffic_model(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d)))
        return {"posting_bytes": a.value, "boost_bytes": b.value, "postings": c.value, "union_hits": d.value}

    WORK_STATS = ["matched_terms", "postings", "union_hits", "sparse_entries", "plane_items", "general_items", "plane_evaluated", "plane_item_evals",
                  "plane_unconverged_sweeps", "plane_sweepless", "tiles", "parts", "planes", "plane_words"]

    def work_stats(self):
        v = (ctypes.c_uint64 * len(self.WORK_STATS))()
        _check(self.L.vgpu_batch_work_stats(self.h, v, len(self.WORK_STATS)))
        return dict(zip(self.WORK_STATS, [int(x) for x in v]))

    def profile_execute(self):
        """One execute with a CUDA event pair around every kernel launch -> {"kernel": {"launches", "ms"}}."""
        _check(self.L.vgpu_batch_set_profiling(self.h, 1))
        try:
            self.execute()
            return json.loads(self.L.vgpu_batch_kernel_times_json(self.h).decode())
        finally:
            _check(self.L.vgpu_batch_set_profiling(self.h, 0))

    def io_bytes(self):
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        _check(self.L.vgpu_batch_io_bytes(self.h, ctypes.byref(a), ctypes.byref(b)))
        return {"h2d": a.value, "d2h": b.value}

    def close(self):
        if self.h:
            self.L.vgpu_batch_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id():
    """128 bytes (an ncclUniqueId) made by one rank and handed to every rank's `Index.comm_init`."""
    buf = ctypes.create_string_buffer(128)
    _check(load_library().vgpu_comm_unique_id(ctypes.cast(buf, ctypes.c_void_p)))
    return buf.raw


class PlanChannel:
    """Shared-memory channel between the processes of one box: local rank 0 publishes plans, the others import them."""

    def __init__(self, name, local_rank, local_ranks, capacity=64 << 20):
        self.L = load_library()
        self.rank = local_rank
        self.h = ctypes.c_void_p()
        _check(self.L.vgpu_plan_channel_open(name.encode(), local_rank, local_ranks, capacity, ctypes.byref(self.h)))

    def ticket(self):
        """Number of the caller's next batch; taken once per batch in execution order (by every rank alike)."""
        return int(self.L.vgpu_plan_channel_ticket(self.h))

    def close(self):
        if self.h:
            self.L.vgpu_plan_channel_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Index:
    """An index directory loaded into the HBM of one GPU (one anchor-range shard of it)."""

    def __init__(self, directory, device=0, shard_rank=0, n_shards=1, planes=True, deletion_index=True):
        self.L = load_library()
        self.h = ctypes.c_void_p()
        flags = (0 if planes else 1) | (0 if deletion_index else 2)
        _check(self.L.vgpu_index_open_ex(os.fsencode(directory), device, shard_rank, n_shards, flags, ctypes.byref(self.h)))

    def comm_init(self, unique_id):
        """Joins the communicator of the shards (NCCL): afterwards `Batch.execute` is a collective that ends with the complete
        result on every rank."""
        _check(self.L.vgpu_comm_init(self.h, ctypes.cast(ctypes.create_string_buffer(unique_id, 128), ctypes.c_void_p)))

    def comm_destroy(self):
        _check(self.L.vgpu_comm_destroy(self.h))

    def info(self):
        a, b, c, d = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        _check(self.L.vgpu_index_info(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d)))
        return {"num_docs": a.value, "anchor_lo": b.value, "anchor_hi": c.value, "device_bytes": d.value}

    def prepare(self, requests, plan=None, channel=None, ticket=None):
        return Batch(self, requests, plan=plan, channel=channel, ticket=ticket)

    def search(self, request):
        b = Batch(self, [request]).execute()
        try:
            return b.result(0)
        finally:
            b.close()

    def search_batch(self, requests, k=10):
        """One vgpu_search_batch call: host JSON in, host arrays out."""
        n = len(requests)
        arr, keep = _encode_requests(requests)
        ids = np.zeros((n, k), dtype=np.uint32)
        scores = np.zeros((n, k), dtype=np.float32)
        num_hits = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        _check(self.L.vgpu_search_batch(self.h, arr, n, k, ids.ctypes.data, scores.ctypes.data, num_hits.ctypes.data, status.ctypes.data))
        return {"ids": ids, "scores": scores, "num_hits": num_hits, "status": status}

    def search_stream(self, batches, k=10, run=None, channel=None, depth=None):
        """Evaluates a sequence of request batches, yielding one `results_flat(k)` dict per batch, in order.

        Planner threads parse, plan and upload the next batches (`vgpu_batch_prepare`, host work + H2D on the batch's own
        stream) while the calling thread has batch i on the GPU and reads its rows back: the steady-state time per batch
        is max(prepare / depth, execute + fetch) instead of their sum.  `depth` batches are prepared ahead, each on its own
        thread (default: 1; 2 with a plan channel, where one process plans for every rank of the box and planning, not the
        GPUs, bounds the rate).  `run(batch)` replaces the plain `batch.execute()` when the caller has more to do per batch."""
        from collections import deque
        from concurrent.futures import ThreadPoolExecutor
        if depth is None:
            depth = 2 if channel is not None else 1
        it = iter(batches)
        with ThreadPoolExecutor(max_workers=depth, thread_name_prefix="veloci-plan") as planner:
            ahead = deque()

            def plan_next():
                try:
                    reqs = next(it)
                except StopIteration:
                    return False
                ticket = channel.ticket() if channel is not None else None  # in execution order, on this thread
                ahead.append(planner.submit(self.prepare, reqs, None, channel, ticket))
                return True

            while len(ahead) < depth and plan_next():
                pass
            while ahead:
                batch = ahead.popleft().result()
                plan_next()
                try:
                    if run is None:
                        batch.execute()
                    else:
                        run(batch)
                    yield batch.results_flat(k)
                finally:
                    batch.close()

    # ---- suggest (search_field.rs:178-228): [(text, score, term id)]
    def _suggest(self, fn, payload):
        out = _Suggestions()
        _check(fn(self.h, json.dumps(payload, ensure_ascii=False).encode("utf-8"), ctypes.byref(out)))
        try:
            return [(out.items[i].text.decode("utf-8"), np.float32(out.items[i].score), out.items[i].id) for i in range(out.n)]
        finally:
            self.L.vgpu_suggestions_free(ctypes.byref(out))

    def suggest_multi(self, request):
        """`request`: {"suggest": [part, ...], "top": .., "skip": ..}"""
        return self._suggest(self.L.vgpu_suggest, request)

    def suggest(self, part):
        """One RequestSearchPart; its top/skip bound the list."""
        return self._suggest(self.L.vgpu_suggest_part, part)

    def highlight(self, part):
        """search_field::highlight: [(highlighted text, score, text id)] of a part with "snippet": true."""
        return self._suggest(self.L.vgpu_highlight, part)

    # ---- documents
    def get_doc(self, doc_id):
        """The stored document of a hit (DocLoader::get_doc), parsed."""
        out = ctypes.c_void_p()
        return json.loads(_take_string(self.L, self.L.vgpu_get_doc(self.h, int(doc_id), ctypes.byref(out)), out))

    def read_doc(self, doc_id, fields):
        """The document rebuilt from the indices, `fields` only (read_data, the request's `select`)."""
        out = ctypes.c_void_p()
        return json.loads(_take_string(self.L, self.L.vgpu_read_doc(self.h, int(doc_id), json.dumps(list(fields)).encode("utf-8"), ctypes.byref(out)), out))

    # ---- request generation (query_generator::search_query / suggest_query)
    def search_query(self, params=None, **kw):
        """SearchQueryGeneratorParameters (dict and/or keywords, e.g. search_term="nice AND cool~1") -> the request's JSON text,
        ready for Batch / search_batch."""
        out = ctypes.c_void_p()
        text = json.dumps({**(params or {}), **kw}, ensure_ascii=False).encode("utf-8")
        return _take_string(self.L, self.L.vgpu_search_query(self.h, text, ctypes.byref(out)), out)

    def suggest_query(self, request, **kw):
        """suggest_query(request, top, skip, levenshtein, fields, levenshtein_auto_limit) -> JSON text for suggest_multi."""
        out = ctypes.c_void_p()
        text = json.dumps({"request": request, **kw}, ensure_ascii=False).encode("utf-8")
        return _take_string(self.L, self.L.vgpu_suggest_query(self.h, text, ctypes.byref(out)), out)

    # ---- step seam
    def field_search(self, part, get_scores=True, get_ids=False):
        out = _HitList()
        _check(self.L.vgpu_field_search(self.h, json.dumps(part, ensure_ascii=False).encode("utf-8"), int(get_scores), int(get_ids), ctypes.byref(out)))
        return _from_hitlist(self.L, out)

    def resolve_to_anchor(self, part, hits, ids=None):
        inp, out = _to_hitlist(hits, ids), _HitList()
        _check(self.L.vgpu_resolve_to_anchor(self.h, json.dumps(part, ensure_ascii=False).encode("utf-8"), ctypes.byref(inp), ctypes.byref(out)))
        return _from_hitlist(self.L, out)

    def union_hits_score(self, lists, terms):
        n = len(lists)
        keep = [_to_hitlist(l) for l in lists]
        arr = (_HitList * max(1, n))(*keep)
        tarr = (ctypes.c_char_p * max(1, n))(*[t.encode("utf-8") for t in terms])
        out = _HitList()
        _check(self.L.vgpu_union_hits_score(self.h, arr, tarr, n, ctypes.byref(out)))
        return _from_hitlist(self.L, out)[0]

    def intersect_hits_score(self, lists):
        n = len(lists)
        keep = [_to_hitlist(l) for l in lists]
        arr = (_HitList * max(1, n))(*keep)
        out = _HitList()
        _check(self.L.vgpu_intersect_hits_score(self.h, arr, n, ctypes.byref(out)))
        return _from_hitlist(self.L, out)[0]

    def resolve_to_anchor_filtered(self, part, hits, filter_ids):
        inp, out = _to_hitlist(hits), _HitList()
        arr = (ctypes.c_uint32 * max(1, len(filter_ids)))(*filter_ids)
        _check(self.L.vgpu_resolve_to_anchor_filtered(self.h, json.dumps(part, ensure_ascii=False).encode("utf-8"), ctypes.byref(inp), arr, len(filter_ids), ctypes.byref(out)))
        return _from_hitlist(self.L, out)[0]

    def _ids_op(self, fn, id_lists):
        n = len(id_lists)
        keep = [_to_hitlist(None, l) for l in id_lists]
        arr = (_HitList * max(1, n))(*keep)
        out = _HitList()
        _check(fn(self.h, arr, n, ctypes.byref(out)))
        return _from_hitlist(self.L, out)[1]

    def union_hits_ids(self, id_lists):
        return self._ids_op(self.L.vgpu_union_hits_ids, id_lists)

    def intersect_hits_ids(self, id_lists):
        return self._ids_op(self.L.vgpu_intersect_hits_ids, id_lists)

    def intersect_scores_with_ids(self, hits, ids):
        a, b, out = _to_hitlist(hits), _to_hitlist(None, ids), _HitList()
        _check(self.L.vgpu_intersect_scores_with_ids(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(out)))
        return _from_hitlist(self.L, out)[0]

    def facet(self, facet_request, ids):
        """get_facet over `ids` -> [(text, count, value id)]"""
        out = _Suggestions()
        arr = (ctypes.c_uint32 * max(1, len(ids)))(*ids)
        _check(self.L.vgpu_facet(self.h, json.dumps(facet_request).encode("utf-8"), arr, len(ids), ctypes.byref(out)))
        try:
            return [(out.items[i].text.decode("utf-8"), int(out.items[i].score), out.items[i].id) for i in range(out.n)]
        finally:
            self.L.vgpu_suggestions_free(ctypes.byref(out))

    def add_boost(self, boost, hits):
        inp = _to_hitlist(hits)
        _check(self.L.vgpu_add_boost(self.h, json.dumps(boost).encode("utf-8"), ctypes.byref(inp)))
        return [(inp.hits[i].id, np.float32(inp.hits[i].score)) for i in range(inp.n_hits)]

    def phrase_pairs_to_anchor(self, path, ids1, ids2):
        """PlanStepPhrasePairToAnchorId: sorted anchors of every (id1, id2) pair."""
        a1, a2 = (ctypes.c_uint32 * max(1, len(ids1)))(*ids1), (ctypes.c_uint32 * max(1, len(ids2)))(*ids2)
        out = _HitList()
        _check(self.L.vgpu_phrase_pairs_to_anchor(self.h, path.encode("utf-8"), a1, len(ids1), a2, len(ids2), ctypes.byref(out)))
        return _from_hitlist(self.L, out)[1]

    def boost_anchor_from_phrase_results(self, hits, phrase_results, groups):
        """BoostAnchorFromPhraseResults: `phrase_results` id lists, `groups` the phrase each belongs to."""
        inp, out = _to_hitlist(hits), _HitList()
        keep = [_to_hitlist([], ids) for ids in phrase_results]
        arr = (_HitList * max(1, len(keep)))(*keep)
        g = (ctypes.c_uint32 * max(1, len(groups)))(*groups)
        _check(self.L.vgpu_boost_anchor_from_phrase_results(self.h, ctypes.byref(inp), arr, g, len(keep), ctypes.byref(out)))
        return _from_hitlist(self.L, out)[0]

    def boost_to_anchor(self, part, boost, term_hits=None, text_ids=None):
        """BoostToAnchor: (anchor, boost value) pairs in value-id order."""
        inp, out = _to_hitlist(term_hits, text_ids), _HitList()
        _check(self.L.vgpu_boost_to_anchor(self.h, json.dumps(part, ensure_ascii=False).encode("utf-8"), ctypes.byref(inp), json.dumps(boost).encode("utf-8"), ctypes.byref(out)))
        return _from_hitlist(self.L, out)[0]

    def apply_anchor_boost(self, boost, hits, boost_ids):
        """ApplyAnchorBoost: `hits` with the boost values of `boost_ids` applied."""
        inp, vals, out = _to_hitlist(hits), _to_hitlist(boost_ids), _HitList()
        _check(self.L.vgpu_apply_anchor_boost(self.h, json.dumps(boost).encode("utf-8"), ctypes.byref(inp), ctypes.byref(vals), ctypes.byref(out)))
        return _from_hitlist(self.L, out)[0]

    def text_locality(self, path, term_token_ids):
        """boost_text_locality: `term_token_ids` = per query term the token ids it matched -> [(anchor, boost)]."""
        keep = [_to_hitlist([], ids) for ids in term_token_ids]
        arr = (_HitList * max(1, len(keep)))(*keep)
        out = _HitList()
        _check(self.L.vgpu_text_locality(self.h, path.encode("utf-8"), arr, len(keep), ctypes.byref(out)))
        return _from_hitlist(self.L, out)[0]

    # ---- step seam, device-resident: lists stay in HBM between the steps
    def _dev(self, call):
        out = ctypes.c_void_p()
        _check(call(ctypes.byref(out)))
        return DeviceHitList(self, out)

    def dev_upload(self, hits):
        inp = _to_hitlist(hits)
        return self._dev(lambda out: self.L.vgpu_dev_upload(self.h, ctypes.byref(inp), out))

    def dev_resolve_to_anchor(self, part, term_hits):
        inp = _to_hitlist(term_hits)
        text = json.dumps(part, ensure_ascii=False).encode("utf-8")
        return self._dev(lambda out: self.L.vgpu_dev_resolve_to_anchor(self.h, text, ctypes.byref(inp), out))

    def dev_union_hits_score(self, lists, terms):
        n = len(lists)
        arr = (ctypes.c_void_p * max(1, n))(*[l.h for l in lists])
        tarr = (ctypes.c_char_p * max(1, n))(*[t.encode("utf-8") for t in terms])
        return self._dev(lambda out: self.L.vgpu_dev_union_hits_score(self.h, arr, tarr, n, out))

    def dev_intersect_hits_score(self, lists):
        n = len(lists)
        arr = (ctypes.c_void_p * max(1, n))(*[l.h for l in lists])
        return self._dev(lambda out: self.L.vgpu_dev_intersect_hits_score(self.h, arr, n, out))

    def dev_add_boost(self, boost, hits):
        text = json.dumps(boost).encode("utf-8")
        return self._dev(lambda out: self.L.vgpu_dev_add_boost(self.h, text, hits.h, out))

    def dev_top_n(self, hits, top, skip=0):
        out = _HitList()
        _check(self.L.vgpu_dev_top_n(self.h, hits.h, top, skip, ctypes.byref(out)))
        return _from_hitlist(self.L, out)[0]

    def top_n(self, hits, top, skip=0):
        inp, out = _to_hitlist(hits), _HitList()
        _check(self.L.vgpu_top_n(self.h, ctypes.byref(inp), top, skip, ctypes.byref(out)))
        return _from_hitlist(self.L, out)[0]

    def close(self):
        if self.h:
            self.L.vgpu_index_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
