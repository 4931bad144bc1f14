// CPU oracle (TEST INFRASTRUCTURE): Rust's `str::to_lowercase` (alloc::str, used at src/search/search_field.rs:284,312),
// restated over the oracle's own tables (lower_tables.hpp): char::to_lowercase for every scalar (U+0130 gives two), and
// `map_uppercase_sigma`: a capital sigma is written as the final form exactly when, skipping Case_Ignorable scalars, a Cased
// scalar comes before it and none comes after it.  Pinned against CPython's str.lower() (the same Unicode algorithm,
// implemented independently) in tests/test_lowercase.py.
#pragma once
#include <string>
#include <vector>

#include "lower_tables.hpp"

namespace olow {

inline bool in_table(const unsigned (*t)[2], int n, unsigned cp) {
    for (int lo = 0, hi = n - 1; lo <= hi;) {
        const int mid = (lo + hi) / 2;
        if (cp < t[mid][0]) hi = mid - 1;
        else if (cp > t[mid][1]) lo = mid + 1;
        else return true;
    }
    return false;
}
inline unsigned lower_one(unsigned cp) {  // scalars with a one-scalar lowercase
    for (int lo = 0, hi = lower_pairs_n - 1; lo <= hi;) {
        const int mid = (lo + hi) / 2;
        if (cp < lower_pairs[mid][0]) hi = mid - 1;
        else if (cp > lower_pairs[mid][0]) lo = mid + 1;
        else return lower_pairs[mid][1];
    }
    return cp;
}
inline std::vector<unsigned> decode(const std::string& s) {
    std::vector<unsigned> out;
    for (size_t i = 0; i < s.size();) {
        const unsigned char b = (unsigned char)s[i];
        const int n = b < 0x80 ? 1 : b >= 0xF0 ? 4 : b >= 0xE0 ? 3 : 2;
        unsigned cp = n == 1 ? b : b & (0xFFu >> (n + 1));
        for (int k = 1; k < n && i + k < s.size(); ++k) cp = (cp << 6) | ((unsigned char)s[i + k] & 0x3F);
        out.push_back(cp);
        i += (size_t)n;
    }
    return out;
}
inline void encode(std::string& out, unsigned cp) {
    if (cp < 0x80) out += (char)cp;
    else if (cp < 0x800) out += (char)(0xC0 | cp >> 6), out += (char)(0x80 | (cp & 0x3F));
    else if (cp < 0x10000) out += (char)(0xE0 | cp >> 12), out += (char)(0x80 | ((cp >> 6) & 0x3F)), out += (char)(0x80 | (cp & 0x3F));
    else out += (char)(0xF0 | cp >> 18), out += (char)(0x80 | ((cp >> 12) & 0x3F)), out += (char)(0x80 | ((cp >> 6) & 0x3F)), out += (char)(0x80 | (cp & 0x3F));
}

inline std::string to_lowercase(const std::string& s) {
    const std::vector<unsigned> in = decode(s);
    std::string out;
    for (size_t i = 0; i < in.size(); ++i) {
        if (in[i] == 0x3A3) {
            bool before = false, after = false;
            for (size_t j = i; j-- > 0;)
                if (!in_table(ignorable_ranges, ignorable_ranges_n, in[j])) {
                    before = in_table(cased_ranges, cased_ranges_n, in[j]);
                    break;
                }
            for (size_t j = i + 1; j < in.size(); ++j)
                if (!in_table(ignorable_ranges, ignorable_ranges_n, in[j])) {
                    after = in_table(cased_ranges, cased_ranges_n, in[j]);
                    break;
                }
            encode(out, before && !after ? 0x3C2 : 0x3C3);
        } else if (in[i] == 0x130) {
            encode(out, 0x69), encode(out, 0x307);
        } else {
            encode(out, lower_one(in[i]));
        }
    }
    return out;
}

}  // namespace olow
