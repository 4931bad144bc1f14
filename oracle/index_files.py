"""CPU oracle (TEST INFRASTRUCTURE, never imported by the product): a second, independent decoder of the files of an
index directory, in plain Python -- the `fst` 0.4 term dictionary (format version 3), vint arrays and "most common value"
vint arrays (`vint32` 0.3), Indirect 1:n stores, packed single arrays, token -> (anchor, score) postings and the
phrase-pair store.  Written from the same published formats as the product's reader (csrc/format/{fst,vint,codecs}.hpp)
but sharing no code with it: tests/test_index_decoder.py holds the two readers against each other on every file of index
directories the C++ writer produced, and feeds the C++ reader dictionaries this file's FST writer produced -- the writer
here uses the one-byte COMMON_INPUTS encoding of OneTrans / OneTransNext states, which the C++ writer never emits.

Neither reader has been run on a veloci-written index (no Rust toolchain here): SURVEY §8 f.1 stays open.
Reference call sites: src/persistence.rs:206-305 (which file is what), src/indices/indirect/indirect.rs:10-89,
src/indices/direct/single_array.rs:17-63, src/indices/persistence_score/token_to_anchor_score_vint.rs:128-204,
src/indices/persistence_data_binary_search.rs:126-203."""
import json
import os
import struct

# ------------------------------------------------------------------------------------------------ vint32
HIGH_BIT = 1 << 31


def vint(buf, pos):
    """little-endian base 128 -> (value, next position)"""
    value = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        value |= (b & 0x7F) << shift
        if b < 0x80:
            return value, pos
        shift += 7


def vint_array(buf, pos):
    """VIntArray::serialize: vint(byte length) || vints"""
    n, pos = vint(buf, pos)
    end = pos + n
    out = []
    while pos < end:
        v, pos = vint(buf, pos)
        out.append(v)
    return out


def vint_common_array(buf, pos):
    """VIntArrayEncodeMostCommon: vint(most common value) || vint(byte length) || items; the first byte of an item is
    [more][is most common][6 payload bits], continuation bytes are [more][7 payload bits]"""
    common, pos = vint(buf, pos)
    n, pos = vint(buf, pos)
    end = pos + n
    out = []
    while pos < end:
        b = buf[pos]
        pos += 1
        if b == 0x40:
            out.append(common)
            continue
        value, shift = b & 0x3F, 6
        while b & 0x80:
            b = buf[pos]
            pos += 1
            value |= (b & 0x7F) << shift
            shift += 7
        out.append(value)
    return out


# ------------------------------------------------------------------------------------------------ stores
def indirect_get_values(start_pos, data, value_id):
    """indirect.rs:56-89: None for an id out of range or an empty bucket"""
    if value_id * 4 + 4 > len(start_pos):
        return None
    (slot,) = struct.unpack_from("<I", start_pos, value_id * 4)
    if slot & HIGH_BIT:
        return [slot & ~HIGH_BIT]
    if slot == 0:
        return None
    if slot >= len(data):
        return []
    return vint_array(data, slot)


def packed_get_value(raw, width, value_id):
    """single_array.rs:93-147: `width` bytes per id, 0 = no value, else value + 1"""
    chunk = raw[value_id * width:(value_id + 1) * width]
    if not chunk:
        return None
    v = int.from_bytes(chunk, "little")
    return None if v == 0 else v - 1


def packed_width(max_value_id):
    """get_bytes_required (single_array.rs): the reference doubles the value"""
    val = (2 * max_value_id) & 0xFFFFFFFF
    return 1 if val < 1 << 8 else 2 if val < 1 << 16 else 3 if val < 1 << 24 else 4


def anchor_scores(start_pos, data, token_id, wide=False):
    """token_to_anchor_score_vint.rs:128-204: [(anchor, score)]; anchors delta coded"""
    w = 8 if wide else 4
    if token_id * w + w > len(start_pos):
        return []
    pos = int.from_bytes(start_pos[token_id * w:token_id * w + w], "little")
    if pos == 0 or pos >= len(data):
        return []
    vals = vint_common_array(data, pos)
    out, anchor = [], 0
    for i in range(0, len(vals) - 1, 2):
        anchor += vals[i]
        out.append((anchor, vals[i + 1]))
    return out


def phrase_pair_records(recs, data):
    """persistence_data_binary_search.rs:51-92: {(t1, t2): [anchors]} from 12-byte (t1, t2, offset) records"""
    out = {}
    for i in range(0, len(recs), 12):
        t1, t2, off = struct.unpack_from("<III", recs, i)
        out[(t1, t2)] = vint_array(data, off) if off < len(data) else []
    return out


# ------------------------------------------------------------------------------------------------ fst
COMMON_INPUTS_INV = b"te/oasripcnw.hlm-du012g=:bf3y5&_4v9678k%?xCDASFIBEjPTzRNM+LOqHGW"  # index (1-based) -> input byte
EMPTY_ADDRESS = 0


class Fst:
    """fst 0.4, format version 3: header u64 version, u64 type; nodes addressed by their LAST byte (the state byte),
    address 0 = the final node without transitions; footer u64 number of keys, u64 root address, u32 checksum."""

    def __init__(self, data):
        self.d = data
        version, _ty = struct.unpack_from("<QQ", data, 0)
        assert 1 <= version <= 3, version
        self.version = version
        end = len(data) - 4 if version >= 3 else len(data)
        self.n_keys, self.root = struct.unpack_from("<QQ", data, end - 16)

    def _uint(self, at, n):
        return int.from_bytes(self.d[at:at + n], "little")

    def node(self, addr):
        """-> (is_final, final_output, [(input byte, output, target address)])"""
        if addr == EMPTY_ADDRESS:
            return True, 0, []
        d = self.d
        state = d[addr]
        kind = state >> 6
        if kind in (2, 3):
            common = state & 0x3F
            at = addr  # walks down from the state byte
            if common:
                inp = COMMON_INPUTS_INV[common - 1]
            else:
                at -= 1
                inp = d[at]
            if kind == 3:  # OneTransNext: the target is the node right below
                return False, 0, [(inp, 0, at - 1)]
            at -= 1
            tsize, osize = d[at] >> 4, d[at] & 15
            at -= tsize
            delta = self._uint(at, tsize)
            at -= osize
            out = self._uint(at, osize) if osize else 0
            return False, 0, [(inp, out, EMPTY_ADDRESS if delta == 0 else at - delta)]  # `at` = the node's first byte
        is_final = bool(state & 0x40)
        ntrans = state & 0x3F
        at = addr
        if ntrans == 0:
            at -= 1
            ntrans = 256 if d[at] == 1 else d[at]
        at -= 1
        tsize, osize = d[at] >> 4, d[at] & 15
        if self.version >= 2 and ntrans > 32:
            at -= 256  # the transition index
        inputs_end = at
        at -= ntrans
        addrs_end = at
        at -= ntrans * tsize
        outs_end = at
        at -= ntrans * osize
        final_out = 0
        if is_final:
            at -= osize
            final_out = self._uint(at, osize) if osize else 0
        first = at
        trans = []
        for i in range(ntrans):  # transition i is stored i-th from the end of each block
            inp = d[inputs_end - 1 - i]
            delta = self._uint(addrs_end - (i + 1) * tsize, tsize)
            out = self._uint(outs_end - (i + 1) * osize, osize) if osize else 0
            trans.append((inp, out, EMPTY_ADDRESS if delta == 0 else first - delta))
        return is_final, final_out, trans

    def items(self):
        """every (key bytes, value) in key order"""
        out = []
        stack = [(self.root, b"", 0)]
        while stack:
            addr, key, acc = stack.pop()
            is_final, final_out, trans = self.node(addr)
            if is_final:
                out.append((key, acc + final_out))
            for inp, o, target in reversed(trans):
                stack.append((target, key + bytes([inp]), acc + o))
        return out

    def ord_to_term(self, ord_):
        """search_field.rs:36-51: the key whose value is `ord_` when the values are the keys' ranks.  -> (found, bytes); for a
        value no key carries (ids of texts too long to be stored, create_fulltext.rs:60-64) the walk ends somewhere and
        get_text_for_id (:520-526) takes the bytes gathered so far all the same."""
        out = bytearray()
        is_final, _, trans = self.node(self.root)
        while ord_ != 0 or not is_final:
            pick = None
            for t in trans:
                if t[1] <= ord_:
                    pick = t
                else:
                    break
            if pick is None:
                return False, bytes(out)
            ord_ -= pick[1]
            out.append(pick[0])
            is_final, _, trans = self.node(pick[2])
        return True, bytes(out)

    def get(self, key):
        addr, acc = self.root, 0
        for b in key:
            _, _, trans = self.node(addr)
            for inp, o, target in trans:
                if inp == b:
                    addr, acc = target, acc + o
                    break
            else:
                return None
        is_final, final_out, _ = self.node(addr)
        return acc + final_out if is_final else None


def _crc32c(data):
    table = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        table.append(c)
    c = 0xFFFFFFFF
    for b in data:
        c = table[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _pack_size(v):
    n = 1
    while n < 8 and v >> (8 * n):
        n += 1
    return n


def write_fst(pairs, common_inputs=True):
    """An FST builder for tests: keys (bytes) strictly ascending, values non-decreasing.  A trie without suffix sharing,
    outputs pushed onto the first transition where they are known; single-transition nodes use OneTransNext / OneTrans
    with the one-byte COMMON_INPUTS encoding when the input has one (`common_inputs`)."""
    common_index = {b: i + 1 for i, b in enumerate(COMMON_INPUTS_INV)}
    root = {"final": None, "kids": {}}
    for key, value in pairs:
        node = root
        for b in key:
            node = node["kids"].setdefault(b, {"final": None, "kids": {}})
        node["final"] = value
    buf = bytearray(struct.pack("<QQ", 3, 0))
    last_addr = [None]

    def smallest(node):  # the smallest value below a node: what an edge can carry as its output
        if node["final"] is not None:
            return node["final"]
        return smallest(node["kids"][min(node["kids"])])

    def compile_node(node, carried):
        """writes the node (children first) -> address; `carried` = the sum of outputs on the path to it"""
        kids = []
        for b in sorted(node["kids"]):
            child = node["kids"][b]
            out = smallest(child) - carried
            kids.append((b, out, compile_node(child, carried + out)))
        is_final = node["final"] is not None
        final_out = node["final"] - carried if is_final else 0
        if is_final and not kids and final_out == 0:
            return EMPTY_ADDRESS
        start = len(buf)
        if len(kids) == 1 and not is_final:
            b, out, target = kids[0]
            ci = common_index.get(b, 0) if common_inputs else 0
            if target == last_addr[0] and out == 0 and target + 1 == start:
                if not ci:
                    buf.append(b)
                buf.append(0xC0 | ci)
            else:
                osize = _pack_size(out) if out else 0
                if osize:
                    buf.extend(out.to_bytes(osize, "little"))
                delta = 0 if target == EMPTY_ADDRESS else start - target
                tsize = _pack_size(delta)
                buf.extend(delta.to_bytes(tsize, "little"))
                buf.append(tsize << 4 | osize)
                if not ci:
                    buf.append(b)
                buf.append(0x80 | ci)
        else:
            n = len(kids)
            deltas = [0 if t == EMPTY_ADDRESS else start - t for _, _, t in kids]
            tsize = max([_pack_size(x) for x in deltas], default=1)
            any_out = final_out != 0 or any(o for _, o, _ in kids)
            osize = max([_pack_size(o) for _, o, _ in kids] + ([_pack_size(final_out)] if is_final else []), default=0) if any_out else 0
            if osize:
                if is_final:
                    buf.extend(final_out.to_bytes(osize, "little"))
                for _, o, _ in reversed(kids):
                    buf.extend(o.to_bytes(osize, "little"))
            for x in reversed(deltas):
                buf.extend(x.to_bytes(tsize, "little"))
            for b, _, _ in reversed(kids):
                buf.append(b)
            if n > 32:
                index = bytearray([255] * 256)
                for i, (b, _, _) in enumerate(kids):
                    index[b] = i
                buf.extend(index)
            buf.append(tsize << 4 | osize)
            state = 0x40 if is_final else 0
            if 1 <= n <= 63:
                state |= n
            else:
                buf.append(1 if n == 256 else n)
            buf.append(state)
        last_addr[0] = len(buf) - 1
        return last_addr[0]

    root_addr = compile_node(root, 0)
    buf.extend(struct.pack("<QQ", len(pairs), root_addr))
    crc = _crc32c(bytes(buf))
    buf.extend(struct.pack("<I", (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF))
    return bytes(buf)


# ------------------------------------------------------------------------------------------------ a whole directory
class IndexDirectory:
    """metaData.json and the files it names (src/persistence.rs:206-305)"""

    def __init__(self, path):
        self.path = path
        self.meta = json.load(open(os.path.join(path, "metaData.json")))

    def read(self, name):
        with open(os.path.join(self.path, name), "rb") as f:
            return f.read()

    def indices(self):
        for field, col in self.meta["columns"].items():
            for ix in col["indices"]:
                yield field, ix

    def dictionary(self, field):
        return Fst(self.read(field + ".textindex.fst")).items()
