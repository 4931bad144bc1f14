// Term-dictionary file `<field>.textindex.fst`: a BurntSushi `fst` 0.4 (format
// version 3) Map from term bytes to term id.
//
// Reference call sites: writer src/create/create_fulltext.rs:53-69
// (fst::MapBuilder, keys inserted in byte order, value = term id), reader
// src/persistence.rs:293-305 (fst::Map::new), traversal
// src/search/search_field.rs:36-65 (ord_to_term / search().into_stream()).
//
// The `fst` crate source is NOT under /root/reference (Cargo.toml:29, no
// Cargo.lock).  What is restated here is its published on-disk format:
//   header  u64 version (3), u64 type (0)
//   nodes   compiled bottom-up; a node's address is the index of its LAST byte
//           (the state byte); address 0 = the empty final node (never written),
//           state byte top bits 11 = OneTransNext, 10 = OneTrans, 0x = AnyTrans
//           (bit 6 = final, low 6 bits = ntrans when 1..63).
//   footer  u64 number of keys, u64 root address, u32 masked CRC32C
// UNVERIFIED AGAINST UPSTREAM: no binary fixture in the reference pins it.  In
// particular the COMMON_INPUTS table (one-byte encoding of frequent inputs in
// OneTrans* states) is reproduced from memory for the *reader* only; the
// writer never emits a common-input index, which every conforming reader
// accepts.  The writer does full suffix sharing (upstream uses a bounded LRU
// registry), so files are valid but not byte-identical to upstream's.
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace vfmt {

static const uint64_t kFstVersion = 3;
static const size_t kFstEmptyAddr = 0;
static const size_t kFstNoneAddr = 1;
static const size_t kFstTransIndexThreshold = 32;

// index (1-based in the state byte) -> input byte
static const uint8_t kFstCommonInputsInv[64] = {
    't', 'e', '/', 'o', 'a', 's', 'r', 'i', 'p', 'c', 'n', 'w', '.', 'h', 'l', 'm',
    '-', 'd', 'u', '0', '1', '2', 'g', '=', ':', 'b', 'f', '3', 'y', '5', '&', '_',
    '4', 'v', '9', '6', '7', '8', 'k', '%', '?', 'x', 'C', 'D', 'A', 'S', 'F', 'I',
    'B', 'E', 'j', 'P', 'T', 'z', 'R', 'N', 'M', '+', 'L', 'O', 'q', 'H', 'G', 'W'};

inline uint32_t crc32c(const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);
            table[i] = c;
        }
        init = true;
    }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

inline int fst_pack_size(uint64_t v) {
    int n = 1;
    while (n < 8 && (v >> (8 * n)) != 0) ++n;
    return n;
}

// ------------------------------------------------------------------ writer --
class FstWriter {
  public:
    FstWriter() {
        put_u64(kFstVersion);
        put_u64(0);
        stack_.emplace_back();
    }
    // keys strictly ascending (byte order), values non-decreasing (term ids are)
    void insert(const uint8_t* key, size_t len, uint64_t val) {
        if (n_keys_ > 0 && std::string((const char*)key, len) <= last_key_)
            throw std::runtime_error("fst: keys must be inserted in strictly ascending order");
        size_t prefix = 0;
        uint64_t out = val;
        while (prefix < len && prefix + 1 < stack_.size() && stack_[prefix].has_last && stack_[prefix].last_inp == key[prefix]) {
            uint64_t o = stack_[prefix].last_out;
            if (o > out) throw std::runtime_error("fst: values must be non-decreasing in key order");
            out -= o;
            ++prefix;
        }
        freeze_below(prefix);
        // extend with the suffix
        for (size_t i = prefix; i < len; ++i) {
            stack_[i].has_last = true;
            stack_[i].last_inp = key[i];
            stack_[i].last_out = (i == prefix) ? out : 0;
            stack_.emplace_back();
        }
        if (len == prefix) {
            // key is a prefix of the previous key: impossible under strict ordering, except the empty first key
            stack_[len].is_final = true;
            stack_[len].final_out = out;
        } else {
            stack_.back().is_final = true;
        }
        last_key_.assign((const char*)key, len);
        ++n_keys_;
    }
    void insert(const std::string& k, uint64_t v) { insert((const uint8_t*)k.data(), k.size(), v); }

    std::vector<uint8_t> finish() {
        freeze_below(0);
        size_t root = compile(stack_[0]);
        put_u64(n_keys_);
        put_u64(root);
        uint32_t sum = crc32c(buf_.data(), buf_.size());
        uint32_t masked = ((sum >> 15) | (sum << 17)) + 0xA282EAD8u;
        uint8_t b[4];
        memcpy(b, &masked, 4);
        buf_.insert(buf_.end(), b, b + 4);
        return std::move(buf_);
    }

  private:
    struct Trans {
        uint8_t inp;
        uint64_t out;
        size_t addr;
    };
    struct Unfinished {
        bool is_final = false;
        uint64_t final_out = 0;
        std::vector<Trans> trans;
        bool has_last = false;
        uint8_t last_inp = 0;
        uint64_t last_out = 0;
    };
    std::vector<uint8_t> buf_;
    std::vector<Unfinished> stack_;
    std::unordered_map<std::string, size_t> registry_;
    std::string last_key_;
    uint64_t n_keys_ = 0;
    size_t last_addr_ = kFstNoneAddr;

    void put_u64(uint64_t v) {
        uint8_t b[8];
        memcpy(b, &v, 8);
        buf_.insert(buf_.end(), b, b + 8);
    }
    void put_packed(uint64_t v, int n) {
        for (int i = 0; i < n; ++i) buf_.push_back((uint8_t)(v >> (8 * i)));
    }
    // compile every unfinished node deeper than `depth` and hook it into its parent
    void freeze_below(size_t depth) {
        while (stack_.size() > depth + 1) {
            Unfinished node = std::move(stack_.back());
            stack_.pop_back();
            size_t addr = compile(node);
            Unfinished& parent = stack_.back();
            parent.trans.push_back(Trans{parent.last_inp, parent.last_out, addr});
            parent.has_last = false;
        }
    }
    size_t compile(const Unfinished& node) {
        if (node.is_final && node.trans.empty() && node.final_out == 0) return kFstEmptyAddr;
        std::string sig;
        sig.reserve(16 + node.trans.size() * 17);
        sig.push_back(node.is_final ? 1 : 0);
        sig.append((const char*)&node.final_out, 8);
        for (auto& t : node.trans) {
            sig.push_back((char)t.inp);
            sig.append((const char*)&t.out, 8);
            uint64_t a = t.addr;
            sig.append((const char*)&a, 8);
        }
        auto it = registry_.find(sig);
        if (it != registry_.end()) return it->second;
        size_t start = buf_.size();
        if (node.trans.size() == 1 && !node.is_final) {
            const Trans& t = node.trans[0];
            if (t.addr == last_addr_ && t.out == 0 && last_addr_ + 1 == start) {
                buf_.push_back(t.inp);
                buf_.push_back(0xC0);  // OneTransNext, no common-input index
            } else {
                int osize = t.out ? fst_pack_size(t.out) : 0;
                if (osize) put_packed(t.out, osize);
                uint64_t delta = t.addr == kFstEmptyAddr ? 0 : (uint64_t)(start - t.addr);
                int tsize = fst_pack_size(delta);
                put_packed(delta, tsize);
                buf_.push_back((uint8_t)((tsize << 4) | osize));
                buf_.push_back(t.inp);
                buf_.push_back(0x80);  // OneTrans
            }
        } else {
            size_t n = node.trans.size();
            int tsize = 0, osize = 0;
            bool any_out = node.final_out != 0;
            for (auto& t : node.trans) {
                uint64_t delta = t.addr == kFstEmptyAddr ? 0 : (uint64_t)(start - t.addr);
                tsize = std::max(tsize, fst_pack_size(delta));
                osize = std::max(osize, fst_pack_size(t.out));
                any_out = any_out || t.out != 0;
            }
            if (node.is_final) osize = std::max(osize, fst_pack_size(node.final_out));
            if (n == 0) tsize = 1;
            if (!any_out) osize = 0;
            if (osize) {
                if (node.is_final) put_packed(node.final_out, osize);
                for (size_t i = n; i-- > 0;) put_packed(node.trans[i].out, osize);
            }
            for (size_t i = n; i-- > 0;) {
                const Trans& t = node.trans[i];
                uint64_t delta = t.addr == kFstEmptyAddr ? 0 : (uint64_t)(start - t.addr);
                put_packed(delta, tsize);
            }
            for (size_t i = n; i-- > 0;) buf_.push_back(node.trans[i].inp);
            if (n > kFstTransIndexThreshold) {
                uint8_t index[256];
                memset(index, 255, sizeof index);
                for (size_t i = 0; i < n; ++i) index[node.trans[i].inp] = (uint8_t)i;
                buf_.insert(buf_.end(), index, index + 256);
            }
            buf_.push_back((uint8_t)((tsize << 4) | osize));
            uint8_t state = node.is_final ? 0x40 : 0x00;
            if (n >= 1 && n <= 63) {
                state |= (uint8_t)n;
            } else {
                buf_.push_back(n == 256 ? 1 : (uint8_t)n);
            }
            buf_.push_back(state);
        }
        size_t addr = buf_.size() - 1;
        last_addr_ = addr;
        registry_.emplace(std::move(sig), addr);
        return addr;
    }
};

// ------------------------------------------------------------------ reader --
class FstReader {
  public:
    struct Trans {
        uint8_t inp;
        uint64_t out;
        size_t addr;
    };
    struct Node {
        bool is_final = false;
        uint64_t final_out = 0;
        std::vector<Trans> trans;
    };

    FstReader() = default;
    FstReader(const uint8_t* data, size_t len) { open(data, len); }

    void open(const uint8_t* data, size_t len) {
        d_ = data;
        if (len < 32) throw std::runtime_error("fst: file too small");
        uint64_t version;
        memcpy(&version, data, 8);
        if (version < 1 || version > 3) throw std::runtime_error("fst: unsupported version " + std::to_string(version));
        version_ = version;
        size_t end = version >= 3 ? len - 4 : len;
        uint64_t nkeys, root;
        memcpy(&nkeys, data + end - 16, 8);
        memcpy(&root, data + end - 8, 8);
        n_keys_ = nkeys;
        root_ = (size_t)root;
        len_ = end - 16;
        if (root_ != kFstEmptyAddr && root_ >= len_) throw std::runtime_error("fst: root address out of range");
    }
    uint64_t len() const { return n_keys_; }
    size_t root() const { return root_; }

    void node(size_t addr, Node& n) const {
        n.trans.clear();
        n.is_final = false;
        n.final_out = 0;
        if (addr == kFstEmptyAddr) {
            n.is_final = true;
            return;
        }
        uint8_t state = d_[addr];
        uint8_t kind = state >> 6;
        if (kind == 3) {  // OneTransNext
            uint8_t ci = state & 0x3F;
            size_t input_len = ci ? 0 : 1;
            uint8_t inp = ci ? kFstCommonInputsInv[ci - 1] : d_[addr - 1];
            size_t first = addr - input_len;
            n.trans.push_back(Trans{inp, 0, first - 1});
        } else if (kind == 2) {  // OneTrans
            uint8_t ci = state & 0x3F;
            size_t input_len = ci ? 0 : 1;
            uint8_t inp = ci ? kFstCommonInputsInv[ci - 1] : d_[addr - 1];
            uint8_t sizes = d_[addr - input_len - 1];
            int tsize = sizes >> 4, osize = sizes & 0xF;
            size_t first = addr - input_len - 1 - (size_t)tsize - (size_t)osize;
            uint64_t delta = unpack(addr - input_len - 1 - (size_t)tsize, tsize);
            uint64_t out = osize ? unpack(first, osize) : 0;
            n.trans.push_back(Trans{inp, out, delta == 0 ? kFstEmptyAddr : first - (size_t)delta});
        } else {  // AnyTrans
            n.is_final = (state & 0x40) != 0;
            size_t ntrans = state & 0x3F;
            size_t ntrans_len = 0;
            if (ntrans == 0) {
                ntrans_len = 1;
                ntrans = d_[addr - 1];
                if (ntrans == 1) ntrans = 256;
            }
            uint8_t sizes = d_[addr - ntrans_len - 1];
            size_t tsize = sizes >> 4, osize = sizes & 0xF;
            size_t index_size = (version_ >= 2 && ntrans > kFstTransIndexThreshold) ? 256 : 0;
            size_t final_osize = n.is_final ? osize : 0;
            size_t first = addr - ntrans_len - 1 - index_size - ntrans - ntrans * tsize - ntrans * osize - final_osize;
            size_t inputs_end = addr - ntrans_len - 1 - index_size;  // one past the input of transition 0
            size_t trans_end = inputs_end - ntrans;
            size_t outs_end = trans_end - ntrans * tsize;
            n.trans.resize(ntrans);
            for (size_t i = 0; i < ntrans; ++i) {
                n.trans[i].inp = d_[inputs_end - i - 1];
                uint64_t delta = unpack(trans_end - i * tsize - tsize, (int)tsize);
                n.trans[i].addr = delta == 0 ? kFstEmptyAddr : first - (size_t)delta;
                n.trans[i].out = osize ? unpack(outs_end - i * osize - osize, (int)osize) : 0;
            }
            if (n.is_final && osize) n.final_out = unpack(outs_end - ntrans * osize - osize, (int)osize);
        }
    }

    // Visits every (key, value) in ascending key order.
    void for_each(const std::function<void(const std::string&, uint64_t)>& f) const {
        std::string key;
        walk(root_, 0, key, f);
    }

    // search_field.rs:36-51 ord_to_term (valid when values are the key ranks)
    bool ord_to_term(uint64_t ord, std::string& out) const {
        out.clear();
        Node n;
        node(root_, n);
        while (ord != 0 || !n.is_final) {
            const Trans* pick = nullptr;
            for (auto& t : n.trans) {
                if (t.out <= ord) pick = &t;
                else break;
            }
            if (!pick) return false;
            ord -= pick->out;
            out.push_back((char)pick->inp);
            size_t next = pick->addr;
            node(next, n);
        }
        return true;
    }

    bool get(const std::string& key, uint64_t& val) const {
        Node n;
        size_t addr = root_;
        uint64_t out = 0;
        for (unsigned char c : key) {
            node(addr, n);
            bool found = false;
            for (auto& t : n.trans)
                if (t.inp == c) {
                    out += t.out;
                    addr = t.addr;
                    found = true;
                    break;
                }
            if (!found) return false;
        }
        node(addr, n);
        if (!n.is_final) return false;
        val = out + n.final_out;
        return true;
    }

  private:
    const uint8_t* d_ = nullptr;
    size_t len_ = 0;
    size_t root_ = 0;
    uint64_t n_keys_ = 0;
    uint64_t version_ = 3;

    uint64_t unpack(size_t at, int n) const {
        uint64_t v = 0;
        for (int i = 0; i < n; ++i) v |= (uint64_t)d_[at + i] << (8 * i);
        return v;
    }
    void walk(size_t addr, uint64_t out, std::string& key, const std::function<void(const std::string&, uint64_t)>& f) const {
        Node n;
        node(addr, n);
        if (n.is_final) f(key, out + n.final_out);
        for (auto& t : n.trans) {
            key.push_back((char)t.inp);
            walk(t.addr, out + t.out, key, f);
            key.pop_back();
        }
    }
};

}  // namespace vfmt
