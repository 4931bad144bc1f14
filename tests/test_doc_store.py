"""The document store (first half: fetching the documents of the top hits; doc_store/src/lib.rs).

The product's reader (csrc/host/doc_store.hpp, csrc/format/lz4_block.hpp; reached through the host-only helper library,
the C ABI entry is vgpu_get_doc) against an independent plain-Python reader (oracle/doc_store.py), on the reference's own
unit tests (round trips), on hand-built LZ4 blocks from the format description, and on index directories."""
import ctypes
import json
import os
import random
import sys
import tempfile

import pytest

import helpers
import ref_fixtures as fx

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import doc_store as ods  # noqa: E402  (test infrastructure)


@pytest.fixture(scope="module")
def lib(native_libs):
    L = helpers._index_lib()
    L.vidx_get_doc.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_size_t]
    L.vidx_write_doc_store.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    L.vidx_lz4_decompress.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    L.vidx_lz4_compress.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
    return L


def product_get_doc(lib, directory, doc_id):
    out = ctypes.create_string_buffer(1 << 20)
    rc = lib.vidx_get_doc(directory.encode(), doc_id, out, len(out))
    return rc, out.value.decode("utf-8")


def write_store(lib, docs):
    d = tempfile.mkdtemp(prefix="vb200_docs_")
    err = ctypes.create_string_buffer(1024)
    assert lib.vidx_write_doc_store(os.path.join(d, "data").encode(), json.dumps(docs).encode(), err, 1024) == 0, err.value
    return d


def product_lz4(lib, block, out_len):
    out = ctypes.create_string_buffer(max(1, out_len))
    err = ctypes.create_string_buffer(256)
    rc = lib.vidx_lz4_decompress(block, len(block), out, out_len, err, 256)
    return rc, out.raw[:out_len], err.value.decode()


# ---- LZ4 blocks built by hand from the block format
LZ4_KATS = [
    (bytes([0x50]) + b"hello", b"hello"),                                              # literals only
    (bytes([0x11]) + b"a" + bytes([1, 0]) + bytes([0x50]) + b"bcdef", b"a" * 6 + b"bcdef"),  # overlapping match (run), offset 1, length 5
    (bytes([0x4F]) + b"abcd" + bytes([4, 0, 3]) + bytes([0x10]) + b"!", b"abcd" + b"abcd" * 5 + b"ab" + b"!"),  # match length 15 + 3 + 4 = 22
    (bytes([0xF0, 0x05]) + b"x" * 20, b"x" * 20),                                      # literal length 15 + 5
    (bytes([0xF0, 0xFF, 0x00]) + b"y" * 270, b"y" * 270),                              # 15 + 255 + 0
    (b"", b""),
]


@pytest.mark.parametrize("block,plain", LZ4_KATS)
def test_lz4_block_vectors(lib, block, plain):
    assert ods.lz4_block_decompress(block, len(plain)) == plain
    rc, got, err = product_lz4(lib, block, len(plain))
    assert rc == 0 and got == plain, err


def test_lz4_damaged_blocks_are_refused(lib):
    for block, n in ((bytes([0x11]) + b"a" + bytes([2, 0]), 6), (bytes([0x50]) + b"hel", 5), (bytes([0x50]) + b"hello", 4), (bytes([0x11]) + b"a" + bytes([0, 0]) + bytes([0x00]), 6), (bytes([0x1F]) + b"a" + bytes([1, 0]), 30)):
        rc, _, err = product_lz4(lib, block, n)
        assert rc == 1 and "lz4" in err


def test_lz4_compressor_round_trips_through_both_decoders(lib):
    rng = random.Random(5)
    cases = [b"", b"a", b"abc" * 5, bytes(rng.randrange(256) for _ in range(5000)), b"x" * 70000, json.dumps(fx.TEST_ALL_DOCS, ensure_ascii=False).encode("utf-8") * 7,
             b"".join(rng.choice([b"tag", b"nice", b"cool", b"\xe9\xa3\x9f", b" "]) for _ in range(30000))]
    for data in cases:
        out = ctypes.create_string_buffer(len(data) + len(data) // 200 + 64)
        n = ctypes.c_size_t()
        assert lib.vidx_lz4_compress(data, len(data), out, len(out), ctypes.byref(n)) == 0
        packed = out.raw[:n.value]
        assert ods.decompress_size_prepended(packed) == data
        rc, got, err = product_lz4(lib, packed[4:], len(data))
        assert rc == 0 and got == data, err
        if len(data) > 20000 and len(set(data)) < 16:
            assert len(packed) < len(data) * 2 // 3  # it does compress


def test_doc_store_reference_unit_tests(lib):
    # doc_store/src/lib.rs:172-190
    docs = ['{"test":"ok"}', '{"test2":"ok"}', '{"test3":"ok"}']
    d = write_store(lib, docs)
    loader = ods.DocLoader(open(os.path.join(d, "data"), "rb").read())
    for i, doc in enumerate(docs):
        assert loader.get_doc(i) == doc and product_get_doc(lib, d, i) == (0, doc)
    # doc_store/src/lib.rs:64-81: 2640 copies of one document (several blocks)
    doc1 = '{"category": "superb", "tags": ["nice", "cool"] }'
    d = write_store(lib, [doc1] * 2640)
    data = open(os.path.join(d, "data"), "rb").read()
    loader = ods.DocLoader(data)
    assert len(loader.index) > 3 and len(data) < 2640 * len(doc1) // 10
    for i in range(2640):
        assert loader.get_doc(i) == doc1
    for i in (0, 1, 333, 334, 335, 1000, 2638, 2639):
        assert product_get_doc(lib, d, i) == (0, doc1)
    rc, msg = product_get_doc(lib, d, 2640)
    assert rc == 1 and "does not exist" in msg


def test_documents_of_varying_size_and_block_edges(lib):
    rng = random.Random(9)
    docs = []
    for i in range(400):
        n = rng.choice([1, 5, 50, 300, 3000, 17000 if i % 97 == 0 else 10])
        docs.append(json.dumps({"id": i, "text": "".join(rng.choice("abcdeé食 ") for _ in range(n))}, ensure_ascii=False))
    d = write_store(lib, docs)
    loader = ods.DocLoader(open(os.path.join(d, "data"), "rb").read())
    for i, doc in enumerate(docs):
        assert loader.get_doc(i) == doc
        assert product_get_doc(lib, d, i) == (0, doc), i


def test_index_directory_carries_its_documents(lib):
    d = tempfile.mkdtemp(prefix="vb200_docs_idx_")
    helpers.create_index(d, fx.TEST_QG_DOCS, fx.TEST_QG_CONFIG)
    loader = ods.DocLoader(open(os.path.join(d, "data"), "rb").read())
    for i, doc in enumerate(fx.TEST_QG_DOCS):
        assert json.loads(loader.get_doc(i)) == doc
        rc, text = product_get_doc(lib, d, i)
        assert rc == 0 and json.loads(text) == doc
    meta = json.load(open(os.path.join(d, "metaData.json")))
    assert meta["num_docs"] == len(fx.TEST_QG_DOCS) and meta["bytes_indexed"] > 0
    # the hits of a generated request, as documents (tests/all/test_query_generator.rs:170-179 asserts on hits[0].doc)
    rc, req, _ = helpers.generate_request(d, {"search_term": "urge"})
    hit = helpers.Oracle(d).search(req)["data"][0]
    doc = json.loads(product_get_doc(lib, d, hit[0])[1])
    assert doc["ent_seq"] == "1587690" and doc["commonness"] == 20 and doc["tags"] == ["nice"]


def test_damaged_store_is_refused(lib):
    d = write_store(lib, ['{"a":1}', '{"b":2}'])
    path = os.path.join(d, "data")
    data = bytearray(open(path, "rb").read())
    for cut in (1, 5, len(data) // 2):
        open(path, "wb").write(bytes(data[:-cut]))
        assert product_get_doc(lib, d, 0)[0] == 1
    data[6] ^= 0x40  # inside the first block
    open(path, "wb").write(bytes(data))
    rc, text = product_get_doc(lib, d, 0)
    assert rc == 1 or text != '{"a":1}'
