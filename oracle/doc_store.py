"""CPU oracle (TEST INFRASTRUCTURE, never imported by the product): an independent plain-Python reader of the compressed
document store `data` (doc_store/src/lib.rs) -- LZ4 block decoding from the published block format, the vint header,
the block index -- and of the store's writer (for the round trips the reference's own unit tests do,
doc_store/src/lib.rs:64-81, 172-190).  The product's reader is csrc/host/doc_store.hpp + csrc/format/lz4_block.hpp.
"""
import struct


def lz4_block_decompress(src, out_len):
    """LZ4 block format: token, literal length extension, literals, 2-byte offset, match length extension."""
    out = bytearray()
    i = 0
    n = len(src)
    while i < n:
        token = src[i]
        i += 1
        lit = token >> 4
        if lit == 15:
            while True:
                b = src[i]
                i += 1
                lit += b
                if b != 255:
                    break
        out += src[i:i + lit]
        i += lit
        if i >= n:
            break
        offset = src[i] | (src[i + 1] << 8)
        i += 2
        length = token & 15
        if length == 15:
            while True:
                b = src[i]
                i += 1
                length += b
                if b != 255:
                    break
        length += 4
        assert 0 < offset <= len(out), "match offset outside the output"
        for _ in range(length):  # may overlap
            out.append(out[-offset])
    assert len(out) == out_len, (len(out), out_len)
    return bytes(out)


def decompress_size_prepended(buf):  # lz4_flex::decompress_size_prepended
    (n,) = struct.unpack_from("<I", buf, 0)
    return lz4_block_decompress(buf[4:], n)


def read_vint(buf, pos):
    v = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7


class DocLoader:
    def __init__(self, data):  # lib.rs:15-23
        self.data = data
        (index_size,) = struct.unpack_from("<I", data, len(data) - 4)
        start = len(data) - 4 - index_size
        self.index = [struct.unpack_from("<II", data, start + 8 * k) for k in range(index_size // 8)]

    def get_doc(self, doc_id):  # lib.rs:26-62
        # binary_search_slice (lib.rs:210-240): the block with the largest first id <= doc_id
        k = max(j for j, (first, _) in enumerate(self.index[:-1]) if first <= doc_id)
        while k > 0 and self.index[k - 1][0] == self.index[k][0]:
            k -= 1
        start, end = self.index[k][1] - 1, self.index[k + 1][1] - 1
        block = decompress_size_prepended(self.data[start:end])
        arr_size, pos = read_vint(block, 0)
        arr_end = pos + arr_size
        first_id, pos = read_vint(block, pos)
        offsets = []
        while pos < arr_end:
            v, pos = read_vint(block, pos)
            offsets.append(v)
        at = doc_id - first_id
        return block[arr_end + offsets[at]:arr_end + offsets[at + 1]].decode("utf-8")
