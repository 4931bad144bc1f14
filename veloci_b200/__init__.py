"""veloci_b200 -- B200-native query-time hit pipeline for Veloci indices.

The product is the C-ABI library `lib/libveloci_b200.so` (include/veloci_b200.h);
this package is the thin ctypes host binding used by the tests and bench.py.  It
mirrors the reference's two seams:

    Index(dir)                ~  Persistence::load(dir)           (src/persistence.rs:393)
    Index.search(request)     ~  search::search(request, &pers)   (src/search.rs:143)
    Index.search_batch([...]) ~  the same for a batch of requests
    Index.field_search / resolve_to_anchor / union_hits_score / intersect_hits_score /
    add_boost / top_n         ~  the plan steps (src/plan_creator/plan_steps.rs); Index.dev_* the same over
                                 hit lists that stay on the device (DeviceHitList)
    Index.suggest / suggest_multi / highlight  ~  search_field::suggest / suggest_multi / highlight
    Batch.result_docs / explain ~ search::to_search_result, SearchResult::explain;  explain_plan ~ search::explain_plan

There is no CPU fallback: without the CUDA library or a CUDA device every call raises.
"""
from .api import (  # noqa: F401
    Batch,
    DeviceHitList,
    Index,
    PlanChannel,
    VelociGpuError,
    comm_unique_id,
    device_count,
    explain_plan,
    launch_count,
    lib_path,
    load_library,
)
