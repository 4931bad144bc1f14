// LZ4 block format (the payload of lz4_flex 0.11's `compress_prepend_size` / `decompress_size_prepended`, which the
// reference's doc store calls: doc_store/src/lib.rs:39,139; lz4_flex is a third-party crate absent from /root/reference).
// Restated from the published block format: a block is a run of sequences
//     token (1 byte: literal length in the high nibble, match length - 4 in the low nibble; 15 = "more length bytes
//     follow", each adding 0..255, the first byte below 255 ends the length)
//     literals, then a 2-byte little-endian match offset (1..65535 bytes back, may overlap the bytes being written)
// and the last sequence ends after its literals.  `size-prepended` = the uncompressed length as u32 LE in front.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace vfmt {

struct Lz4Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// Decodes one block of exactly `out_len` bytes.
inline void lz4_block_decompress(const uint8_t* src, size_t src_len, uint8_t* out, size_t out_len) {
    size_t ip = 0, op = 0;
    auto extended = [&](size_t len) {
        if (len != 15) return len;
        while (true) {
            if (ip >= src_len) throw Lz4Error("lz4 block: length runs past the input");
            const uint8_t b = src[ip++];
            len += b;
            if (b != 255) return len;
        }
    };
    while (ip < src_len) {
        const uint8_t token = src[ip++];
        const size_t lit = extended(token >> 4);
        if (lit > src_len - ip || lit > out_len - op) throw Lz4Error("lz4 block: literals run past the buffer");
        memcpy(out + op, src + ip, lit);
        ip += lit, op += lit;
        if (ip == src_len) break;  // the last sequence has no match
        if (src_len - ip < 2) throw Lz4Error("lz4 block: truncated match offset");
        const size_t offset = (size_t)src[ip] | ((size_t)src[ip + 1] << 8);
        ip += 2;
        const size_t len = extended(token & 15u) + 4;
        if (offset == 0 || offset > op) throw Lz4Error("lz4 block: match offset outside the output");
        if (len > out_len - op) throw Lz4Error("lz4 block: match runs past the output");
        for (size_t i = 0; i < len; ++i) out[op + i] = out[op + i - offset];  // byte-wise: the match may overlap its own output
        op += len;
    }
    if (op != out_len) throw Lz4Error("lz4 block: decoded length differs from the prepended size");
}

// decompress_size_prepended
inline std::vector<uint8_t> lz4_decompress_size_prepended(const uint8_t* src, size_t src_len) {
    if (src_len < 4) throw Lz4Error("lz4: missing size prefix");
    const uint32_t n = (uint32_t)src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24);
    if ((uint64_t)n > (uint64_t)src_len * 256 + 64) throw Lz4Error("lz4: implausible size prefix");  // a block expands at most 255x
    std::vector<uint8_t> out(n);
    lz4_block_decompress(src + 4, src_len - 4, out.data(), n);
    return out;
}

// A greedy single-probe compressor producing a valid block (any decoder reads it; the exact bytes need not equal
// lz4_flex's).  End-of-block rules of the format: the last 5 bytes are literals, no match starts in the last 12 bytes.
inline void lz4_compress_prepend_size(const uint8_t* src, size_t n, std::vector<uint8_t>& out) {
    for (int i = 0; i < 4; ++i) out.push_back((uint8_t)((uint32_t)n >> (8 * i)));
    auto put_len = [&](size_t len) {
        for (; len >= 255; len -= 255) out.push_back(255);
        out.push_back((uint8_t)len);
    };
    auto sequence = [&](size_t lit_begin, size_t lit_len, size_t offset, size_t match_len) {  // match_len 0: final literals
        const size_t ml = match_len ? match_len - 4 : 0;
        out.push_back((uint8_t)((lit_len >= 15 ? 15 : lit_len) << 4 | (ml >= 15 ? 15 : ml)));
        if (lit_len >= 15) put_len(lit_len - 15);
        out.insert(out.end(), src + lit_begin, src + lit_begin + lit_len);
        if (!match_len) return;
        out.push_back((uint8_t)(offset & 0xFF)), out.push_back((uint8_t)(offset >> 8));
        if (ml >= 15) put_len(ml - 15);
    };
    std::vector<uint32_t> table(1u << 14, 0xFFFFFFFFu);
    size_t anchor = 0, i = 0;
    while (n >= 13 && i + 12 <= n) {
        uint32_t word;
        memcpy(&word, src + i, 4);
        const uint32_t h = (word * 2654435761u) >> 18;
        const uint32_t cand = table[h];
        table[h] = (uint32_t)i;
        uint32_t cw = 0;
        if (cand != 0xFFFFFFFFu) memcpy(&cw, src + cand, 4);
        if (cand != 0xFFFFFFFFu && i - cand <= 65535 && cw == word) {
            size_t len = 4;
            while (i + len < n - 5 && src[cand + len] == src[i + len]) ++len;
            sequence(anchor, i - anchor, i - cand, len);
            i += len, anchor = i;
        } else {
            ++i;
        }
    }
    sequence(anchor, n - anchor, 0, 0);
}

}  // namespace vfmt
