// UTF-8 helpers and scalar-value lowercasing for the query-time path.
//
// The reference lower-cases with Rust's `str::to_lowercase()`
// (src/search/search_field.rs:284,312).  This table holds every scalar whose
// lowercase form is a single scalar (generated from the Unicode database by
// tools/gen_lower_table.py); `to_lowercase` adds the one scalar that expands
// (U+0130 -> "i" + U+0307) and the context rule for final sigma, with the Cased /
// Case_Ignorable tables of case_props.hpp.  `lower_scalar` alone stays context
// free.  ASCII is handled inline.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "case_props.hpp"

namespace vfmt {

struct LowerRun {
    uint32_t first, last;
    int step;
    int delta;
};
static const LowerRun kLowerRuns[] = {
    {0xC0, 0xD6, 1, 32},
    {0xD8, 0xDE, 1, 32},
    {0x100, 0x12E, 2, 1},
    {0x132, 0x136, 2, 1},
    {0x139, 0x147, 2, 1},
    {0x14A, 0x176, 2, 1},
    {0x178, 0x178, 1, -121},
    {0x179, 0x17D, 2, 1},
    {0x181, 0x181, 1, 210},
    {0x182, 0x184, 2, 1},
    {0x186, 0x186, 1, 206},
    {0x187, 0x187, 1, 1},
    {0x189, 0x18A, 1, 205},
    {0x18B, 0x18B, 1, 1},
    {0x18E, 0x18E, 1, 79},
    {0x18F, 0x18F, 1, 202},
    {0x190, 0x190, 1, 203},
    {0x191, 0x191, 1, 1},
    {0x193, 0x193, 1, 205},
    {0x194, 0x194, 1, 207},
    {0x196, 0x196, 1, 211},
    {0x197, 0x197, 1, 209},
    {0x198, 0x198, 1, 1},
    {0x19C, 0x19C, 1, 211},
    {0x19D, 0x19D, 1, 213},
    {0x19F, 0x19F, 1, 214},
    {0x1A0, 0x1A4, 2, 1},
    {0x1A6, 0x1A6, 1, 218},
    {0x1A7, 0x1A7, 1, 1},
    {0x1A9, 0x1A9, 1, 218},
    {0x1AC, 0x1AC, 1, 1},
    {0x1AE, 0x1AE, 1, 218},
    {0x1AF, 0x1AF, 1, 1},
    {0x1B1, 0x1B2, 1, 217},
    {0x1B3, 0x1B5, 2, 1},
    {0x1B7, 0x1B7, 1, 219},
    {0x1B8, 0x1B8, 1, 1},
    {0x1BC, 0x1BC, 1, 1},
    {0x1C4, 0x1C4, 1, 2},
    {0x1C5, 0x1C5, 1, 1},
    {0x1C7, 0x1C7, 1, 2},
    {0x1C8, 0x1C8, 1, 1},
    {0x1CA, 0x1CA, 1, 2},
    {0x1CB, 0x1DB, 2, 1},
    {0x1DE, 0x1EE, 2, 1},
    {0x1F1, 0x1F1, 1, 2},
    {0x1F2, 0x1F4, 2, 1},
    {0x1F6, 0x1F6, 1, -97},
    {0x1F7, 0x1F7, 1, -56},
    {0x1F8, 0x21E, 2, 1},
    {0x220, 0x220, 1, -130},
    {0x222, 0x232, 2, 1},
    {0x23A, 0x23A, 1, 10795},
    {0x23B, 0x23B, 1, 1},
    {0x23D, 0x23D, 1, -163},
    {0x23E, 0x23E, 1, 10792},
    {0x241, 0x241, 1, 1},
    {0x243, 0x243, 1, -195},
    {0x244, 0x244, 1, 69},
    {0x245, 0x245, 1, 71},
    {0x246, 0x24E, 2, 1},
    {0x370, 0x372, 2, 1},
    {0x376, 0x376, 1, 1},
    {0x37F, 0x37F, 1, 116},
    {0x386, 0x386, 1, 38},
    {0x388, 0x38A, 1, 37},
    {0x38C, 0x38C, 1, 64},
    {0x38E, 0x38F, 1, 63},
    {0x391, 0x3A1, 1, 32},
    {0x3A3, 0x3AB, 1, 32},
    {0x3CF, 0x3CF, 1, 8},
    {0x3D8, 0x3EE, 2, 1},
    {0x3F4, 0x3F4, 1, -60},
    {0x3F7, 0x3F7, 1, 1},
    {0x3F9, 0x3F9, 1, -7},
    {0x3FA, 0x3FA, 1, 1},
    {0x3FD, 0x3FF, 1, -130},
    {0x400, 0x40F, 1, 80},
    {0x410, 0x42F, 1, 32},
    {0x460, 0x480, 2, 1},
    {0x48A, 0x4BE, 2, 1},
    {0x4C0, 0x4C0, 1, 15},
    {0x4C1, 0x4CD, 2, 1},
    {0x4D0, 0x52E, 2, 1},
    {0x531, 0x556, 1, 48},
    {0x10A0, 0x10C5, 1, 7264},
    {0x10C7, 0x10C7, 1, 7264},
    {0x10CD, 0x10CD, 1, 7264},
    {0x13A0, 0x13EF, 1, 38864},
    {0x13F0, 0x13F5, 1, 8},
    {0x1C90, 0x1CBA, 1, -3008},
    {0x1CBD, 0x1CBF, 1, -3008},
    {0x1E00, 0x1E94, 2, 1},
    {0x1E9E, 0x1E9E, 1, -7615},
    {0x1EA0, 0x1EFE, 2, 1},
    {0x1F08, 0x1F0F, 1, -8},
    {0x1F18, 0x1F1D, 1, -8},
    {0x1F28, 0x1F2F, 1, -8},
    {0x1F38, 0x1F3F, 1, -8},
    {0x1F48, 0x1F4D, 1, -8},
    {0x1F59, 0x1F5F, 2, -8},
    {0x1F68, 0x1F6F, 1, -8},
    {0x1F88, 0x1F8F, 1, -8},
    {0x1F98, 0x1F9F, 1, -8},
    {0x1FA8, 0x1FAF, 1, -8},
    {0x1FB8, 0x1FB9, 1, -8},
    {0x1FBA, 0x1FBB, 1, -74},
    {0x1FBC, 0x1FBC, 1, -9},
    {0x1FC8, 0x1FCB, 1, -86},
    {0x1FCC, 0x1FCC, 1, -9},
    {0x1FD8, 0x1FD9, 1, -8},
    {0x1FDA, 0x1FDB, 1, -100},
    {0x1FE8, 0x1FE9, 1, -8},
    {0x1FEA, 0x1FEB, 1, -112},
    {0x1FEC, 0x1FEC, 1, -7},
    {0x1FF8, 0x1FF9, 1, -128},
    {0x1FFA, 0x1FFB, 1, -126},
    {0x1FFC, 0x1FFC, 1, -9},
    {0x2126, 0x2126, 1, -7517},
    {0x212A, 0x212A, 1, -8383},
    {0x212B, 0x212B, 1, -8262},
    {0x2132, 0x2132, 1, 28},
    {0x2160, 0x216F, 1, 16},
    {0x2183, 0x2183, 1, 1},
    {0x24B6, 0x24CF, 1, 26},
    {0x2C00, 0x2C2F, 1, 48},
    {0x2C60, 0x2C60, 1, 1},
    {0x2C62, 0x2C62, 1, -10743},
    {0x2C63, 0x2C63, 1, -3814},
    {0x2C64, 0x2C64, 1, -10727},
    {0x2C67, 0x2C6B, 2, 1},
    {0x2C6D, 0x2C6D, 1, -10780},
    {0x2C6E, 0x2C6E, 1, -10749},
    {0x2C6F, 0x2C6F, 1, -10783},
    {0x2C70, 0x2C70, 1, -10782},
    {0x2C72, 0x2C72, 1, 1},
    {0x2C75, 0x2C75, 1, 1},
    {0x2C7E, 0x2C7F, 1, -10815},
    {0x2C80, 0x2CE2, 2, 1},
    {0x2CEB, 0x2CED, 2, 1},
    {0x2CF2, 0x2CF2, 1, 1},
    {0xA640, 0xA66C, 2, 1},
    {0xA680, 0xA69A, 2, 1},
    {0xA722, 0xA72E, 2, 1},
    {0xA732, 0xA76E, 2, 1},
    {0xA779, 0xA77B, 2, 1},
    {0xA77D, 0xA77D, 1, -35332},
    {0xA77E, 0xA786, 2, 1},
    {0xA78B, 0xA78B, 1, 1},
    {0xA78D, 0xA78D, 1, -42280},
    {0xA790, 0xA792, 2, 1},
    {0xA796, 0xA7A8, 2, 1},
    {0xA7AA, 0xA7AA, 1, -42308},
    {0xA7AB, 0xA7AB, 1, -42319},
    {0xA7AC, 0xA7AC, 1, -42315},
    {0xA7AD, 0xA7AD, 1, -42305},
    {0xA7AE, 0xA7AE, 1, -42308},
    {0xA7B0, 0xA7B0, 1, -42258},
    {0xA7B1, 0xA7B1, 1, -42282},
    {0xA7B2, 0xA7B2, 1, -42261},
    {0xA7B3, 0xA7B3, 1, 928},
    {0xA7B4, 0xA7C2, 2, 1},
    {0xA7C4, 0xA7C4, 1, -48},
    {0xA7C5, 0xA7C5, 1, -42307},
    {0xA7C6, 0xA7C6, 1, -35384},
    {0xA7C7, 0xA7C9, 2, 1},
    {0xA7D0, 0xA7D0, 1, 1},
    {0xA7D6, 0xA7D8, 2, 1},
    {0xA7F5, 0xA7F5, 1, 1},
    {0xFF21, 0xFF3A, 1, 32},
    {0x10400, 0x10427, 1, 40},
    {0x104B0, 0x104D3, 1, 40},
    {0x10570, 0x1057A, 1, 39},
    {0x1057C, 0x1058A, 1, 39},
    {0x1058C, 0x10592, 1, 39},
    {0x10594, 0x10595, 1, 39},
    {0x10C80, 0x10CB2, 1, 64},
    {0x118A0, 0x118BF, 1, 32},
    {0x16E40, 0x16E5F, 1, 32},
    {0x1E900, 0x1E921, 1, 34}};
static const int kNumLowerRuns = (int)(sizeof(kLowerRuns) / sizeof(kLowerRuns[0]));

inline uint32_t lower_scalar(uint32_t cp) {
    if (cp < 0x80) return (cp >= 'A' && cp <= 'Z') ? cp + 32 : cp;
    int lo = 0, hi = kNumLowerRuns - 1;
    while (lo <= hi) {
        int mid = (lo + hi) / 2;
        const LowerRun& r = kLowerRuns[mid];
        if (cp < r.first) hi = mid - 1;
        else if (cp > r.last) lo = mid + 1;
        else {
            if ((cp - r.first) % (uint32_t)r.step == 0) return (uint32_t)((int64_t)cp + r.delta);
            return cp;
        }
    }
    return cp;
}

// Decodes one scalar starting at s[i]; advances i.  Malformed bytes decode as themselves.
inline uint32_t utf8_next(const uint8_t* s, size_t n, size_t& i) {
    uint8_t b = s[i++];
    if (b < 0x80) return b;
    int extra = (b >= 0xF0) ? 3 : (b >= 0xE0) ? 2 : (b >= 0xC0) ? 1 : 0;
    uint32_t cp = (extra == 3) ? (b & 0x07) : (extra == 2) ? (b & 0x0F) : (extra == 1) ? (b & 0x1F) : b;
    while (extra-- > 0 && i < n && (s[i] & 0xC0) == 0x80) cp = (cp << 6) | (s[i++] & 0x3F);
    return cp;
}

inline void utf8_decode(const std::string& s, std::vector<uint32_t>& out) {
    out.clear();
    size_t i = 0;
    while (i < s.size()) out.push_back(utf8_next((const uint8_t*)s.data(), s.size(), i));
}

inline void utf8_append(std::string& out, uint32_t cp) {
    if (cp < 0x80) {
        out.push_back((char)cp);
    } else if (cp < 0x800) {
        out.push_back((char)(0xC0 | (cp >> 6)));
        out.push_back((char)(0x80 | (cp & 0x3F)));
    } else if (cp < 0x10000) {
        out.push_back((char)(0xE0 | (cp >> 12)));
        out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
        out.push_back((char)(0x80 | (cp & 0x3F)));
    } else {
        out.push_back((char)(0xF0 | (cp >> 18)));
        out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
        out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
        out.push_back((char)(0x80 | (cp & 0x3F)));
    }
}

inline bool in_ranges(const uint32_t (*ranges)[2], int n, uint32_t cp) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) / 2;
        if (cp < ranges[mid][0]) hi = mid - 1;
        else if (cp > ranges[mid][1]) lo = mid + 1;
        else return true;
    }
    return false;
}

// Rust's `str::to_lowercase` on scalars: every scalar's lowercase mapping, U+0130 expanding to "i" + U+0307, and a capital
// sigma becoming the final form when a cased letter precedes it and none follows (case-ignorable scalars skipped on both
// sides: alloc::str::to_lowercase / map_uppercase_sigma).
inline void lowercase_scalars(const std::vector<uint32_t>& in, std::vector<uint32_t>& out) {
    out.clear();
    out.reserve(in.size());
    for (size_t i = 0; i < in.size(); ++i) {
        const uint32_t cp = in[i];
        if (cp < 0x80) {
            out.push_back((cp >= 'A' && cp <= 'Z') ? cp + 32 : cp);
        } else if (cp == 0x3A3) {
            auto cased_after_ignorables = [&](long from, long step) {
                for (long j = from; j >= 0 && j < (long)in.size(); j += step)
                    if (!in_ranges(kCaseIgnorableRanges, kNumCaseIgnorableRanges, in[(size_t)j])) return in_ranges(kCasedRanges, kNumCasedRanges, in[(size_t)j]);
                return false;
            };
            const bool word_final = cased_after_ignorables((long)i - 1, -1) && !cased_after_ignorables((long)i + 1, 1);
            out.push_back(word_final ? 0x3C2u : 0x3C3u);
        } else if (cp == 0x130) {
            out.push_back(0x69), out.push_back(0x307);
        } else {
            out.push_back(lower_scalar(cp));
        }
    }
}

inline std::string to_lowercase(const std::string& s) {
    bool ascii = true;
    for (unsigned char c : s) ascii = ascii && c < 0x80;
    std::string out;
    out.reserve(s.size());
    if (ascii) {
        for (char c : s) out.push_back((c >= 'A' && c <= 'Z') ? (char)(c + 32) : c);
        return out;
    }
    std::vector<uint32_t> in, low;
    utf8_decode(s, in);
    lowercase_scalars(in, low);
    for (uint32_t cp : low) utf8_append(out, cp);
    return out;
}

inline size_t utf8_count(const std::string& s) {
    size_t n = 0;
    for (unsigned char c : s) n += (c & 0xC0) != 0x80;
    return n;
}

}  // namespace vfmt
