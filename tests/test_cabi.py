"""The drop-in boundary on a machine without a GPU: the shared library loads, exports every symbol that
include/veloci_b200.h declares, answers the calls that need no device, and refuses -- loudly, with VGPU_ERR_CUDA --
every call that would compute.  No compute entry point is exercised here (that is tests/test_gpu_*.py, -m gpu)."""
import ctypes
import json
import os
import re
import tempfile

import pytest

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "veloci_b200.h")


@pytest.fixture(scope="module")
def lib():
    from veloci_b200 import build

    return ctypes.CDLL(build.build_main_lib())  # nvcc cross-compiles for sm_100a without a device


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_seams():
    names = declared_symbols()
    for must in ("vgpu_index_open", "vgpu_batch_prepare", "vgpu_batch_execute", "vgpu_batch_result", "vgpu_search_batch", "vgpu_field_search",
                 "vgpu_resolve_to_anchor", "vgpu_union_hits_score", "vgpu_intersect_hits_score", "vgpu_add_boost", "vgpu_top_n",
                 "vgpu_batch_local_topk", "vgpu_batch_merge_gathered", "vgpu_batch_facet_group"):
        assert must in names
    assert len(names) >= 30


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, f"declared in include/veloci_b200.h but not exported: {missing}"


def test_every_symbol_appears_in_the_integration_notes():
    """INTEGRATION.md shows the reference-side binding (the `-sys` crate's extern block and the Rust call sites): a symbol
    added to the header without its binding there would be a seam nobody can use."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in declared_symbols() if n not in doc]
    assert not missing, missing


def test_python_binding_binds_only_declared_symbols():
    src = open(os.path.join(ROOT, "veloci_b200", "api.py")).read()
    used = set(re.findall(r"\bL\.(vgpu_[a-z0-9_]+)", src))
    assert used and used <= set(declared_symbols()), used - set(declared_symbols())


def test_no_device_no_result(lib, native_libs):
    """Without a CUDA device the library reports zero devices and index_open fails with VGPU_ERR_CUDA: there is no CPU path."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("this check is for machines without a GPU")
    lib.vgpu_device_count.restype = ctypes.c_int32
    lib.vgpu_last_error.restype = ctypes.c_char_p
    assert lib.vgpu_device_count() == 0
    d = tempfile.mkdtemp(prefix="vb200_cabi_")
    helpers.create_synthetic_index(d, num_docs=200, vocab=50, seed=1)
    handle = ctypes.c_void_p()
    rc = lib.vgpu_index_open(d.encode(), 0, 0, 1, ctypes.byref(handle))
    assert rc == 6 and not handle.value, rc  # VGPU_ERR_CUDA
    assert b"CUDA" in lib.vgpu_last_error() or b"cuda" in lib.vgpu_last_error()
    # null handles are rejected, not dereferenced
    assert lib.vgpu_batch_execute(None) == 1
    n = ctypes.c_uint32()
    assert lib.vgpu_batch_prepare_jsonl(None, b"{}", 2, ctypes.byref(n), ctypes.byref(handle)) == 1


def test_python_loader_has_no_fallback(monkeypatch):
    import veloci_b200
    from veloci_b200 import api

    monkeypatch.setattr(api, "_LIB", None)
    monkeypatch.setattr(api, "lib_path", lambda: os.path.join(ROOT, "veloci_b200", "lib", "missing.so"))
    with pytest.raises(veloci_b200.api.VelociGpuError) as e:
        api.load_library()
    assert e.value.status == 6


def test_plan_channel_opens_between_ranks_without_a_device(lib):
    """The shared-memory plan channel is host-only: rank 0 creates it, another rank (here: a thread) attaches, tickets
    count batches from 1, a rank count that disagrees with rank 0's is refused."""
    import threading

    vp, u32, u64 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64
    lib.vgpu_plan_channel_open.argtypes = [ctypes.c_char_p, u32, u32, ctypes.c_size_t, ctypes.POINTER(vp)]
    lib.vgpu_plan_channel_close.argtypes = [vp]
    lib.vgpu_plan_channel_close.restype = None
    lib.vgpu_plan_channel_ticket.argtypes = [vp]
    lib.vgpu_plan_channel_ticket.restype = u64
    lib.vgpu_last_error.restype = ctypes.c_char_p
    name = f"/vb200_cabi_{os.getpid()}".encode()
    other, rc_other = vp(), []
    t = threading.Thread(target=lambda: rc_other.append(lib.vgpu_plan_channel_open(name, 1, 2, 1 << 20, ctypes.byref(other))))
    t.start()  # waits until rank 0 has created the segment
    mine = vp()
    assert lib.vgpu_plan_channel_open(name, 0, 2, 1 << 20, ctypes.byref(mine)) == 0, lib.vgpu_last_error()
    t.join(timeout=30)
    assert rc_other == [0] and other.value
    assert [lib.vgpu_plan_channel_ticket(mine) for _ in range(3)] == [1, 2, 3]
    assert lib.vgpu_plan_channel_ticket(other) == 1
    wrong = vp()
    assert lib.vgpu_plan_channel_open(name, 1, 3, 1 << 20, ctypes.byref(wrong)) != 0  # rank count differs from rank 0's
    assert lib.vgpu_plan_channel_open(name, 5, 2, 1 << 20, ctypes.byref(wrong)) != 0  # rank out of range
    lib.vgpu_plan_channel_close(other)
    lib.vgpu_plan_channel_close(mine)


def test_query_parse_needs_no_device(lib):
    """vgpu_query_parse is host code only (query_parser/src/parser.rs:26-28): it answers on a box without a GPU."""
    from veloci_b200 import api

    assert api.query_parse("a AND b:(c d~2)") == '("a" AND b:("c" OR "d"~2))'
    assert api.query_parse("(cool)", no_parentheses=True) == '"(cool)"'
    with pytest.raises(api.VelociGpuError) as e:
        api.query_parse("fancy~")
    assert e.value.status == 1 and "Expecting a levenshtein number" in str(e.value)
