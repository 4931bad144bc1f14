"""A second, independent decoder of the index files (oracle/index_files.py, plain Python) against the product's reader
(csrc/format/{fst,vint,codecs}.hpp, which the C++ search oracle shares) -- VERDICT r1 item 4/9.  Every file of index
directories written by the C++ indexer is decoded by both; the C++ FST reader is fed dictionaries written by the Python
writer, which uses the one-byte COMMON_INPUTS encoding the C++ writer never emits."""
import ctypes
import os
import random
import sys
import tempfile

import pytest

import helpers
import ref_fixtures as fx

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import index_files as oif  # noqa: E402  (test infrastructure)


def cpp_fst_items(path):
    lib = helpers._index_lib()
    lib.vidx_fst_dump.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    out = ctypes.create_string_buffer(1 << 24)
    rc = lib.vidx_fst_dump(path.encode(), out, len(out))
    assert rc == 0, out.value
    lines = out.value.split(b"\n")
    assert lines[-1] == b"ok", lines[-1]
    items = []
    for line in lines[:-1]:
        key, value = line.rsplit(b"\t", 1)
        items.append((bytes.fromhex(key.decode()), int(value)))
    return items


def key_sets():
    rng = random.Random(17)
    urls = sorted({("http://%s.%s/%s?%s=%d" % (rng.choice(["www", "api", "test"]), rng.choice(["com", "net", "org"]), rng.choice(["a", "path/to", "index"]), rng.choice("pqx"), rng.randrange(50))).encode() for _ in range(400)})
    words = sorted({bytes(rng.choice(b"abcdefgh") for _ in range(rng.randint(1, 9))) for _ in range(3000)})
    wide = sorted({bytes([b]) + bytes([rng.randrange(256)]) for b in range(256) for _ in range(2)})
    utf8 = sorted({s.encode("utf-8") for s in ["食べる", "食べ物", "schön", "schon", "übung", "zebra", "a", "ab", "abc", "abcd", ""]})
    fan40 = sorted({bytes([0x30 + i]) + b"x" for i in range(40)} | {b"\x30"})
    return {"urls": urls, "words": words, "fanout 256": wide, "utf8 + empty key": utf8, "fanout 40 (transition index)": fan40}


@pytest.mark.parametrize("name", list(key_sets()))
@pytest.mark.parametrize("common_inputs", [True, False])
def test_python_written_fst_is_read_by_the_product_reader(native_libs, name, common_inputs):
    keys = key_sets()[name]
    rng = random.Random(len(keys))
    pairs, v = [], 0
    for k in keys:
        v += rng.choice([0, 1, 1, 2, 300, 70000]) if pairs else rng.choice([0, 5])
        pairs.append((k, v))
    data = oif.write_fst(pairs, common_inputs=common_inputs)
    if common_inputs and name == "urls":
        plain = oif.write_fst(pairs, common_inputs=False)
        assert len(data) < len(plain)  # the one-byte encoding is in use
    assert oif.Fst(data).items() == pairs
    path = os.path.join(tempfile.mkdtemp(prefix="vb200_fst_"), "t.fst")
    open(path, "wb").write(data)
    assert cpp_fst_items(path) == pairs


@pytest.fixture(scope="module")
def directories(native_libs):
    out = []
    for docs, config in ((fx.TEST_ALL_DOCS, fx.TEST_ALL_CONFIG), (fx.TEST_PHRASE_DOCS, fx.TEST_PHRASE_CONFIG), (fx.TEST_QG_DOCS, fx.TEST_QG_CONFIG)):
        d = tempfile.mkdtemp(prefix="vb200_dec_")
        helpers.create_index(d, docs, config)
        out.append(d)
    d = tempfile.mkdtemp(prefix="vb200_dec_syn_")
    helpers.create_synthetic_index(d, num_docs=3000, vocab=500, seed=5, tags=12)
    out.append(d)
    return out


def test_every_dictionary(directories):
    n = 0
    for d in directories:
        oracle, ix = helpers.Oracle(d), oif.IndexDirectory(d)
        for field in ix.meta["columns"]:
            if not os.path.exists(os.path.join(d, field + ".textindex.fst")):
                continue
            py = [(k.decode("utf-8"), v) for k, v in ix.dictionary(field)]
            cpp = [(t, i) for t, i in oracle.call("dict", path=field + ".textindex")]
            assert py == cpp, (d, field)
            assert cpp_fst_items(os.path.join(d, field + ".textindex.fst")) == [(k.encode("utf-8"), v) for k, v in py]
            n += len(py)
    assert n > 700


def test_every_store(directories):
    seen = {"KeyValue": 0, "Boost": 0, "AnchorScore": 0, "Phrase": 0}
    for d in directories:
        oracle, ix = helpers.Oracle(d), oif.IndexDirectory(d)
        for field, meta in ix.indices():
            path, cat = meta["path"], meta["index_category"]
            if cat == "AnchorScore":
                start, data = ix.read(path + ".indirect"), ix.read(path + ".data")
                wide = meta["data_type"] == "U64"
                n_ids = len(start) // (8 if wide else 4)
                for token in list(range(min(n_ids, 400))) + [n_ids, n_ids + 7]:
                    assert [list(p) for p in oif.anchor_scores(start, data, token, wide)] == oracle.call("postings", path=path[:-len(".to_anchor_id_score")], id=token), (path, token)
                seen[cat] += n_ids
            elif cat == "Phrase":
                if meta["is_empty"]:
                    continue
                table = oif.phrase_pair_records(ix.read(path + ".indirect"), ix.read(path + ".data"))
                for (t1, t2), anchors in list(table.items())[:500]:
                    assert oracle.call("phrase_pairs", path=path, t1=t1, t2=t2) == anchors, (path, t1, t2)
                assert oracle.call("phrase_pairs", path=path, t1=4000000, t2=1) is None
                seen[cat] += len(table)
            elif cat == "KeyValue":
                if meta["is_empty"]:
                    continue
                if meta["index_cardinality"] == "MultiValue":
                    start, data = ix.read(path + ".indirect"), ix.read(path + ".data")
                    n_ids = len(start) // 4
                    for vid in list(range(min(n_ids, 600))) + [n_ids, n_ids + 3]:
                        assert oif.indirect_get_values(start, data, vid) == oracle.call("get_values", path=path, id=vid), (path, vid)
                else:
                    raw = ix.read(path)
                    width = oif.packed_width(meta["metadata"]["max_value_id"])
                    n_ids = len(raw) // width
                    for vid in list(range(min(n_ids, 600))) + [n_ids + 1]:
                        v = oif.packed_get_value(raw, width, vid)
                        assert (None if v is None else [v]) == oracle.call("get_values", path=path, id=vid), (path, vid)
                seen[cat] += n_ids
            else:  # Boost columns are read through get_boost; the layout is a KeyValue store's
                seen[cat] += 1
    assert all(v > 0 for v in seen.values()), seen


def test_vint_vectors():
    # persistence_data_binary_search.rs:252-253: [5, 6] serialized at offset 1 ends at offset 4
    assert oif.vint_array(bytes([0, 2, 5, 6, 9]), 1) == [5, 6]
    assert oif.vint(bytes([0x80, 0x01]), 0) == (128, 2) and oif.vint(bytes([0xFF, 0xFF, 0xFF, 0xFF, 0x0F]), 0) == (0xFFFFFFFF, 5)
    # most-common array: common value 7, items [7, 3, 64, 7] -> 0x40, 0x03, (0x80 | 0, 0x01), 0x40
    assert oif.vint_common_array(bytes([7, 5, 0x40, 0x03, 0x80, 0x01, 0x40]), 0) == [7, 3, 64, 7]
