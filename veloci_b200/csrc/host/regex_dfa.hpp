// Regex search parts (`is_regex`, src/search/search_field.rs:72-83): the reference compiles the pattern with
// regex-automata 0.1.9 (`dense::Builder::new().case_insensitive(..).build(pattern)`: unanchored, leftmost-first, Unicode,
// syntax of regex-syntax 0.6) and walks the term dictionary with the DFA as an fst::Automaton: a term matches when the
// DFA is in a match state after the term's last byte (`starts_with`: after any prefix).  regex-automata is a third-party
// crate absent from /root/reference; this file restates its published construction at the level of Unicode scalars:
//
//   pattern -> syntax tree -> Thompson NFA with ordered alternatives (an unanchored, lazy `(?s:.)*?` in front)
//           -> subset construction that keeps the NFA states of a DFA state in priority order and drops everything after
//              the first Match state (leftmost-first: `a|ab` never accepts "ab", the `a` branch has matched and cut `ab`)
//           -> scalar classes + a dense transition table, which regex_match_kernel (cuda/fuzzy.cu) runs over the
//              dictionary's symbol strings, one term per thread.
//
// The byte-level automaton of the crate and this scalar-level one accept the same valid-UTF-8 strings: a UTF-8 sequence
// is a deterministic expansion of its scalar, every thread leaves a character class through the same successor, and the
// order of threads is decided at scalar boundaries only.
//
// Syntax covered: literals, escapes of meta characters, \n \r \t \f \v \a \0, \xHH \x{H..} \u{H..} \uHHHH, `.`, classes
// `[a-z]` `[^...]` with escapes, \d \D \s \S (Unicode), `* + ? {m} {m,} {m,n}` and their lazy forms, `|`, groups `( )`
// `(?: )` `(?P<n> )`, flags i and s (`(?i)`, `(?s:..)`, `(?-i)`).  Anchors and word boundaries make the reference's
// `build(..).unwrap()` panic (the 0.1 DFA does not support them): RegexError.  Everything else (\w \p{..}, class set
// operations, POSIX classes, flags m x u U) is RegexUnsupported.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../format/case_fold.hpp"

namespace vregex {

struct RegexError : std::runtime_error {  // the reference cannot build this pattern either
    using std::runtime_error::runtime_error;
};
struct RegexUnsupported : std::runtime_error {  // valid for the reference, outside this implementation
    using std::runtime_error::runtime_error;
};

struct Range {
    uint32_t lo, hi;
};
using CharSet = std::vector<Range>;  // sorted, disjoint, not adjacent

static const uint32_t kMaxScalar = 0x10FFFF;

inline void normalize(CharSet& s) {
    std::sort(s.begin(), s.end(), [](const Range& a, const Range& b) { return a.lo < b.lo; });
    CharSet out;
    for (const Range& r : s) {
        if (!out.empty() && r.lo <= out.back().hi + 1) out.back().hi = std::max(out.back().hi, r.hi);
        else out.push_back(r);
    }
    s.swap(out);
}
inline void drop_surrogates(CharSet& s) {
    CharSet out;
    for (const Range& r : s) {
        if (r.hi < 0xD800 || r.lo > 0xDFFF) {
            out.push_back(r);
            continue;
        }
        if (r.lo < 0xD800) out.push_back(Range{r.lo, 0xD7FF});
        if (r.hi > 0xDFFF) out.push_back(Range{0xE000, r.hi});
    }
    s.swap(out);
}
inline CharSet negate(const CharSet& s) {  // over the scalar values
    CharSet out;
    uint32_t next = 0;
    for (const Range& r : s) {
        if (r.lo > next) out.push_back(Range{next, r.lo - 1});
        next = r.hi + 1;
    }
    if (next <= kMaxScalar) out.push_back(Range{next, kMaxScalar});
    drop_surrogates(out);
    return out;
}
inline uint32_t simple_fold(uint32_t c) {
    int lo = 0, hi = vfmt::kNumCaseFold - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) / 2;
        if (vfmt::kCaseFold[mid][0] == c) return vfmt::kCaseFold[mid][1];
        if (vfmt::kCaseFold[mid][0] < c) lo = mid + 1;
        else hi = mid - 1;
    }
    return c;
}
// regex-syntax's simple case folding of a class: every scalar that folds like a member joins it
inline void add_case_folds(CharSet& s) {
    normalize(s);
    auto has = [&](uint32_t c) {
        for (const Range& r : s)
            if (c >= r.lo && c <= r.hi) return true;
        return false;
    };
    CharSet extra;
    std::vector<uint32_t> targets;  // folds of the members
    for (int i = 0; i < vfmt::kNumCaseFold; ++i)
        if (has(vfmt::kCaseFold[i][0])) targets.push_back(vfmt::kCaseFold[i][1]);
    std::sort(targets.begin(), targets.end());
    for (uint32_t t : targets) extra.push_back(Range{t, t});
    for (int i = 0; i < vfmt::kNumCaseFold; ++i) {
        const uint32_t t = vfmt::kCaseFold[i][1];
        if (has(t) || std::binary_search(targets.begin(), targets.end(), t)) extra.push_back(Range{vfmt::kCaseFold[i][0], vfmt::kCaseFold[i][0]});
    }
    s.insert(s.end(), extra.begin(), extra.end());
    normalize(s);
}

// ------------------------------------------------------------------------------------------------ syntax tree
struct Ast {
    enum Kind { Empty, Set, Concat, Alternation, Repeat } kind = Empty;
    CharSet set;
    std::vector<std::unique_ptr<Ast>> kids;
    uint32_t min = 0, max = 0;  // Repeat; max == kNoMax: unbounded
    bool greedy = true;
    static const uint32_t kNoMax = 0xFFFFFFFFu;
};

class PatternParser {
   public:
    PatternParser(const std::string& pattern, bool case_insensitive) : s_(pattern) { flags_.i = case_insensitive; }

    std::unique_ptr<Ast> parse() {
        auto a = alternation(0);
        if (pos_ < s_.size()) throw RegexError("regex parse error: unopened group at byte " + std::to_string(pos_));
        return a;
    }

   private:
    struct Flags {
        bool i = false, s = false;
    };
    static const uint32_t kMaxRepeat = 1000, kMaxDepth = 200;

    bool eof() const { return pos_ >= s_.size(); }
    char cur() const { return s_[pos_]; }
    uint32_t next_scalar() {
        const uint8_t b = (uint8_t)s_[pos_];
        if (b < 0x80) return ++pos_, b;
        const uint32_t n = b >= 0xF0 ? 4 : b >= 0xE0 ? 3 : 2;
        uint32_t cp = b & (0xFF >> (n + 1));
        for (uint32_t k = 1; k < n && pos_ + k < s_.size(); ++k) cp = (cp << 6) | ((uint8_t)s_[pos_ + k] & 0x3F);
        pos_ = std::min(s_.size(), pos_ + n);
        return cp;
    }
    std::unique_ptr<Ast> set_node(CharSet set) const {
        if (flags_.i) add_case_folds(set);
        else normalize(set);
        auto a = std::make_unique<Ast>();
        a->kind = Ast::Set, a->set = std::move(set);
        return a;
    }
    static CharSet digits() {
        CharSet s;
        for (int i = 0; i < vfmt::kNumDecimalDigitRanges; ++i) s.push_back(Range{vfmt::kDecimalDigitRanges[i][0], vfmt::kDecimalDigitRanges[i][1]});
        return s;
    }
    static CharSet spaces() {  // White_Space
        return CharSet{{0x09, 0x0D}, {0x20, 0x20}, {0x85, 0x85}, {0xA0, 0xA0}, {0x1680, 0x1680}, {0x2000, 0x200A}, {0x2028, 0x2029}, {0x202F, 0x202F}, {0x205F, 0x205F}, {0x3000, 0x3000}};
    }

    std::unique_ptr<Ast> alternation(uint32_t depth) {
        if (depth > kMaxDepth) throw RegexError("regex parse error: nesting too deep");
        const Flags saved = flags_;  // flags set inside a group end with it
        std::vector<std::unique_ptr<Ast>> alts;
        alts.push_back(concat(depth));
        while (!eof() && cur() == '|') {
            ++pos_;
            alts.push_back(concat(depth));
        }
        flags_ = saved;
        if (alts.size() == 1) return std::move(alts[0]);
        auto a = std::make_unique<Ast>();
        a->kind = Ast::Alternation, a->kids = std::move(alts);
        return a;
    }

    std::unique_ptr<Ast> concat(uint32_t depth) {
        auto a = std::make_unique<Ast>();
        a->kind = Ast::Concat;
        while (!eof() && cur() != '|' && cur() != ')') {
            std::unique_ptr<Ast> item = atom(depth);
            if (!item) continue;  // a bare flag group "(?i)"
            while (!eof() && (cur() == '*' || cur() == '+' || cur() == '?' || cur() == '{')) {
                uint32_t lo = 0, hi = Ast::kNoMax;
                if (cur() == '{') {
                    if (!counted(lo, hi)) break;  // a '{' that does not open a counted repetition is a parse error in regex-syntax
                } else {
                    lo = cur() == '+' ? 1 : 0, hi = cur() == '?' ? 1 : Ast::kNoMax;
                    ++pos_;
                }
                auto rep = std::make_unique<Ast>();
                rep->kind = Ast::Repeat, rep->min = lo, rep->max = hi, rep->greedy = true;
                if (!eof() && cur() == '?') rep->greedy = false, ++pos_;
                rep->kids.push_back(std::move(item));
                item = std::move(rep);
            }
            a->kids.push_back(std::move(item));
        }
        if (a->kids.empty()) a->kind = Ast::Empty;
        return a;
    }

    bool counted(uint32_t& lo, uint32_t& hi) {  // at '{'
        size_t p = pos_ + 1;
        auto number = [&](uint32_t& v) {
            const size_t start = p;
            uint64_t acc = 0;
            while (p < s_.size() && s_[p] >= '0' && s_[p] <= '9') acc = std::min<uint64_t>(acc * 10 + (uint64_t)(s_[p] - '0'), 1u << 30), ++p;
            v = (uint32_t)acc;
            return p > start;
        };
        if (!number(lo)) throw RegexError("regex parse error: repetition quantifier expects a valid decimal");
        hi = lo;
        if (p < s_.size() && s_[p] == ',') {
            ++p;
            if (!number(hi)) hi = Ast::kNoMax;
        }
        if (p >= s_.size() || s_[p] != '}') throw RegexError("regex parse error: unclosed counted repetition");
        if (hi != Ast::kNoMax && hi < lo) throw RegexError("regex parse error: invalid repetition count range");
        if (lo > kMaxRepeat || (hi != Ast::kNoMax && hi > kMaxRepeat)) throw RegexUnsupported("counted repetitions above " + std::to_string(kMaxRepeat));
        pos_ = p + 1;
        return true;
    }

    std::unique_ptr<Ast> atom(uint32_t depth) {
        const char c = cur();
        if (c == '(') return group(depth);
        if (c == '[') return char_class();
        if (c == '.') {
            ++pos_;
            CharSet all = flags_.s ? CharSet{{0, kMaxScalar}} : CharSet{{0, 0x09}, {0x0B, kMaxScalar}};
            drop_surrogates(all);
            auto a = std::make_unique<Ast>();
            a->kind = Ast::Set, a->set = std::move(all);
            return a;
        }
        if (c == '^' || c == '$') throw RegexError("the reference's DFA does not support anchors (regex-automata 0.1: build() fails, the search panics)");
        if (c == '*' || c == '+' || c == '?') throw RegexError("regex parse error: repetition operator missing expression");
        if (c == '{') throw RegexError("regex parse error: repetition operator missing expression");
        if (c == '\\') {
            CharSet set;
            if (escape(set, false)) return set_node(std::move(set));
        }
        const uint32_t cp = next_scalar();
        return set_node(CharSet{{cp, cp}});
    }

    std::unique_ptr<Ast> group(uint32_t depth) {
        ++pos_;  // '('
        const Flags outer = flags_;
        if (!eof() && cur() == '?') {
            ++pos_;
            if (!eof() && cur() == 'P') {  // (?P<name>...)
                ++pos_;
                if (eof() || cur() != '<') throw RegexError("regex parse error: unrecognized flag");
                while (!eof() && cur() != '>') ++pos_;
                if (eof()) throw RegexError("regex parse error: unclosed group name");
                ++pos_;
            } else {
                bool on = true;
                Flags f = flags_;
                while (!eof() && cur() != ':' && cur() != ')') {
                    const char fc = cur();
                    if (fc == '-') on = false;
                    else if (fc == 'i') f.i = on;
                    else if (fc == 's') f.s = on;
                    else if (fc == 'm' || fc == 'x' || fc == 'u' || fc == 'U') throw RegexUnsupported(std::string("regex flag '") + fc + "'");
                    else throw RegexError("regex parse error: unrecognized flag");
                    ++pos_;
                }
                if (eof()) throw RegexError("regex parse error: unclosed group");
                if (cur() == ')') {  // "(?i)": the flags hold until the enclosing group ends
                    ++pos_;
                    flags_ = f;
                    return nullptr;
                }
                ++pos_;  // ':'
                flags_ = f;
            }
        }
        auto inner = alternation(depth + 1);
        if (eof() || cur() != ')') throw RegexError("regex parse error: unclosed group");
        ++pos_;
        flags_ = outer;
        return inner;
    }

    // after a backslash: a class (\d \s ...) or one scalar; false = not an escape this parser folds into `set`
    bool escape(CharSet& set, bool in_class) {
        ++pos_;  // '\\'
        if (eof()) throw RegexError("regex parse error: incomplete escape sequence");
        const char c = cur();
        auto hex = [&](size_t digits_n, bool braces) {
            uint32_t v = 0;
            size_t n = 0;
            if (braces) {
                ++pos_;  // '{'
                while (!eof() && cur() != '}') {
                    const int d = hex_digit(cur());
                    if (d < 0 || ++n > 8) throw RegexError("regex parse error: invalid hexadecimal digit");
                    v = v * 16 + (uint32_t)d, ++pos_;
                }
                if (eof() || n == 0) throw RegexError("regex parse error: incomplete hexadecimal escape");
                ++pos_;
            } else {
                for (; n < digits_n; ++n) {
                    if (eof() || hex_digit(cur()) < 0) throw RegexError("regex parse error: invalid hexadecimal digit");
                    v = v * 16 + (uint32_t)hex_digit(cur()), ++pos_;
                }
            }
            if (v > kMaxScalar || (v >= 0xD800 && v <= 0xDFFF)) throw RegexError("regex parse error: invalid Unicode scalar value");
            return v;
        };
        switch (c) {
            case 'd': ++pos_, set = digits(); return true;
            case 'D': ++pos_, set = digits(), normalize(set), set = negate(set); return true;
            case 's': ++pos_, set = spaces(); return true;
            case 'S': ++pos_, set = negate(spaces()); return true;
            case 'w': case 'W': case 'p': case 'P': throw RegexUnsupported(std::string("regex class \\") + c);
            case 'b': case 'B': case 'A': case 'z':
                if (in_class) throw RegexError("regex parse error: unrecognized escape sequence");
                throw RegexError("the reference's DFA does not support anchors and word boundaries (regex-automata 0.1: build() fails, the search panics)");
            case 'n': ++pos_, set = {{'\n', '\n'}}; return true;
            case 'r': ++pos_, set = {{'\r', '\r'}}; return true;
            case 't': ++pos_, set = {{'\t', '\t'}}; return true;
            case 'f': ++pos_, set = {{0x0C, 0x0C}}; return true;
            case 'v': ++pos_, set = {{0x0B, 0x0B}}; return true;
            case 'a': ++pos_, set = {{0x07, 0x07}}; return true;
            case '0': ++pos_, set = {{0, 0}}; return true;
            case 'x': case 'u': case 'U': {
                ++pos_;
                const uint32_t v = (!eof() && cur() == '{') ? hex(0, true) : hex(c == 'x' ? 2 : c == 'u' ? 4 : 8, false);
                set = {{v, v}};
                return true;
            }
            default: break;
        }
        if (strchr("\\.+*?()|[]{}^$#&-~", c) != nullptr || (in_class && c == ':')) {
            ++pos_;
            set = {{(uint32_t)(uint8_t)c, (uint32_t)(uint8_t)c}};
            return true;
        }
        throw RegexError("regex parse error: unrecognized escape sequence");
    }
    static int hex_digit(char c) { return c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1; }

    std::unique_ptr<Ast> char_class() {
        ++pos_;  // '['
        bool negated = false;
        if (!eof() && cur() == '^') negated = true, ++pos_;
        CharSet set;
        bool first = true;
        while (true) {
            if (eof()) throw RegexError("regex parse error: unclosed character class");
            if (cur() == ']' && !first) break;
            first = false;
            if (cur() == '[') throw RegexUnsupported("nested / POSIX character classes");
            if ((cur() == '&' && s_.compare(pos_, 2, "&&") == 0) || (cur() == '~' && s_.compare(pos_, 2, "~~") == 0) || (cur() == '-' && s_.compare(pos_, 2, "--") == 0))
                throw RegexUnsupported("character class set operations");
            CharSet item;
            bool single = true;  // one scalar: may start a range
            if (cur() == '\\') {
                const char e = pos_ + 1 < s_.size() ? s_[pos_ + 1] : 0;
                single = !(e == 'd' || e == 'D' || e == 's' || e == 'S');
                escape(item, true);
            } else {
                const uint32_t cp = next_scalar();
                item = {{cp, cp}};
            }
            if (single && pos_ + 1 < s_.size() && cur() == '-' && s_[pos_ + 1] != ']') {
                ++pos_;  // '-'
                CharSet hi_item;
                if (cur() == '\\') {
                    const char e = pos_ + 1 < s_.size() ? s_[pos_ + 1] : 0;
                    if (e == 'd' || e == 'D' || e == 's' || e == 'S') throw RegexError("regex parse error: invalid character class range");
                    escape(hi_item, true);
                } else {
                    if (cur() == '[') throw RegexUnsupported("nested / POSIX character classes");
                    const uint32_t cp = next_scalar();
                    hi_item = {{cp, cp}};
                }
                if (hi_item[0].lo < item[0].lo) throw RegexError("regex parse error: invalid character class range");
                item = {{item[0].lo, hi_item[0].lo}};
            }
            set.insert(set.end(), item.begin(), item.end());
        }
        ++pos_;  // ']'
        if (flags_.i) add_case_folds(set);
        else normalize(set);
        drop_surrogates(set);
        if (negated) set = negate(set);
        auto a = std::make_unique<Ast>();
        a->kind = Ast::Set, a->set = std::move(set);
        return a;
    }

    const std::string& s_;
    size_t pos_ = 0;
    Flags flags_;
};

// ------------------------------------------------------------------------------------------------ NFA
struct NfaState {
    enum Kind : uint8_t { Set, Union, Match } kind = Match;
    uint32_t set = 0;           // Set: index into Nfa::sets
    uint32_t next = 0;          // Set
    std::vector<uint32_t> alt;  // Union: alternatives, best first
};
struct Nfa {
    std::vector<NfaState> states;
    std::vector<CharSet> sets;
    uint32_t start = 0;
    static const size_t kMaxStates = 200000;

    uint32_t add(NfaState s) {
        if (states.size() >= kMaxStates) throw RegexUnsupported("regex too large (more than 200000 NFA states)");
        states.push_back(std::move(s));
        return (uint32_t)states.size() - 1;
    }
    uint32_t add_set(const CharSet& set, uint32_t next) {
        sets.push_back(set);
        NfaState s;
        s.kind = NfaState::Set, s.set = (uint32_t)sets.size() - 1, s.next = next;
        return add(std::move(s));
    }
    uint32_t add_union() {
        NfaState s;
        s.kind = NfaState::Union;
        return add(std::move(s));
    }

    // the fragment for `a` that continues at `next` (the shapes of regex-automata 0.1's nfa/compiler.rs: c_at_least,
    // c_bounded, c_zero_or_one, alternation and concatenation in order)
    uint32_t compile(const Ast& a, uint32_t next) {
        switch (a.kind) {
            case Ast::Empty: return next;
            case Ast::Set:
                if (a.set.empty()) {  // a class nothing is in: no way through
                    return add_union();
                }
                return add_set(a.set, next);
            case Ast::Concat: {
                uint32_t at = next;
                for (size_t i = a.kids.size(); i-- > 0;) at = compile(*a.kids[i], at);
                return at;
            }
            case Ast::Alternation: {
                const uint32_t u = add_union();
                std::vector<uint32_t> alt;
                for (auto& k : a.kids) alt.push_back(compile(*k, next));
                states[u].alt = std::move(alt);
                return u;
            }
            case Ast::Repeat: {
                const Ast& body = *a.kids[0];
                uint32_t after = next;
                if (a.max == Ast::kNoMax) {
                    // a* : U -> [a -> U, next];  a+ : a -> U -> [a, next]  (lazy: the alternatives swapped)
                    const uint32_t u = add_union();
                    const uint32_t loop = compile(body, u);
                    states[u].alt = a.greedy ? std::vector<uint32_t>{loop, next} : std::vector<uint32_t>{next, loop};
                    if (a.min == 0) return u;
                    after = loop;
                    for (uint32_t i = 1; i < a.min; ++i) after = compile(body, after);
                    return after;
                }
                // a{m,n}: m copies, then (a(a(a)?)?)? with every exit going straight to `next`
                for (uint32_t i = a.min; i < a.max; ++i) {
                    const uint32_t copy = compile(body, after);
                    const uint32_t u = add_union();
                    states[u].alt = a.greedy ? std::vector<uint32_t>{copy, next} : std::vector<uint32_t>{next, copy};
                    after = u;
                }
                for (uint32_t i = 0; i < a.min; ++i) after = compile(body, after);
                return after;
            }
        }
        return next;
    }

    static Nfa of(const Ast& ast) {
        Nfa n;
        NfaState m;
        m.kind = NfaState::Match;
        const uint32_t match = n.add(std::move(m));
        const uint32_t body = n.compile(ast, match);
        // unanchored: `(?s:.)*?` in front, lazy, so a restart is the thread of lowest priority
        const uint32_t u = n.add_union();
        CharSet all{{0, kMaxScalar}};
        drop_surrogates(all);
        const uint32_t any = n.add_set(all, u);
        n.states[u].alt = {body, any};
        n.start = u;
        return n;
    }
};

// ------------------------------------------------------------------------------------------------ DFA
// States: 0 is the dead state.  `trans[state * n_classes + class]` = next state | kMatchBit when that state is a match
// state; `class_ranges` maps scalars to classes (sorted, covering every scalar).
struct Dfa {
    static const uint16_t kMatchBit = 0x8000;
    static const size_t kMaxStates = 0x7FFF;
    struct ClassRange {
        uint32_t lo, hi;
        uint16_t cls;
    };
    std::vector<ClassRange> class_ranges;
    uint32_t n_classes = 0, n_states = 0;
    uint16_t start = 0;  // with kMatchBit when the start state matches (the empty string)
    std::vector<uint16_t> trans;

    uint16_t class_of(uint32_t scalar) const {
        size_t lo = 0, hi = class_ranges.size();
        while (lo < hi) {
            const size_t mid = (lo + hi) / 2;
            if (class_ranges[mid].hi < scalar) lo = mid + 1;
            else hi = mid;
        }
        return class_ranges[lo].cls;
    }
    // the reference's test on a dictionary term (host-side twin of regex_match_kernel, for tests and tiny dictionaries)
    bool matches(const std::vector<uint32_t>& scalars, bool starts_with) const {
        uint16_t st = start;
        if (starts_with && (st & kMatchBit)) return true;
        for (uint32_t c : scalars) {
            st = trans[(size_t)(st & ~kMatchBit) * n_classes + class_of(c)];
            if ((st & ~kMatchBit) == 0) return false;
            if (starts_with && (st & kMatchBit)) return true;
        }
        return !starts_with && (st & kMatchBit);
    }

    static Dfa of(const Nfa& nfa) {
        Dfa d;
        // scalar classes: scalars no character set of the pattern tells apart
        std::vector<uint32_t> cuts{0, kMaxScalar + 1};
        for (auto& set : nfa.sets)
            for (const Range& r : set) cuts.push_back(r.lo), cuts.push_back(r.hi + 1);
        std::sort(cuts.begin(), cuts.end());
        cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
        std::map<std::vector<bool>, uint16_t> class_of_signature;
        std::vector<uint32_t> representative;
        for (size_t i = 0; i + 1 < cuts.size(); ++i) {
            const uint32_t lo = cuts[i], hi = cuts[i + 1] - 1;
            std::vector<bool> sig(nfa.sets.size());
            for (size_t s = 0; s < nfa.sets.size(); ++s) {
                const CharSet& set = nfa.sets[s];
                auto it = std::upper_bound(set.begin(), set.end(), lo, [](uint32_t v, const Range& r) { return v < r.lo; });
                sig[s] = it != set.begin() && (it - 1)->hi >= lo;
            }
            auto ins = class_of_signature.emplace(std::move(sig), (uint16_t)class_of_signature.size());
            if (ins.second) representative.push_back(lo);
            if (!d.class_ranges.empty() && d.class_ranges.back().cls == ins.first->second) d.class_ranges.back().hi = hi;
            else d.class_ranges.push_back(ClassRange{lo, hi, ins.first->second});
        }
        d.n_classes = (uint32_t)class_of_signature.size();
        // which sets hold each class's representative
        std::vector<std::vector<bool>> in_set(d.n_classes, std::vector<bool>(nfa.sets.size()));
        for (auto& kv : class_of_signature) in_set[kv.second] = kv.first;

        struct State {
            bool is_match = false;
            std::vector<uint32_t> nfa_states;  // Set states in priority order, nothing after the first Match
            bool operator<(const State& o) const { return is_match != o.is_match ? is_match < o.is_match : nfa_states < o.nfa_states; }
        };
        std::vector<uint32_t> order;  // insertion-ordered set of NFA states
        std::vector<uint8_t> seen(nfa.states.size(), 0);
        std::vector<uint32_t> stack;
        auto closure = [&](uint32_t from) {  // determinize.rs epsilon_closure: depth first, best alternative first
            stack.push_back(from);
            while (!stack.empty()) {
                uint32_t id = stack.back();
                stack.pop_back();
                while (true) {
                    if (seen[id]) break;
                    seen[id] = 1, order.push_back(id);
                    const NfaState& s = nfa.states[id];
                    if (s.kind != NfaState::Union || s.alt.empty()) break;
                    for (size_t k = s.alt.size(); k-- > 1;) stack.push_back(s.alt[k]);
                    id = s.alt[0];
                }
            }
        };
        auto freeze = [&]() {  // determinize.rs new_state
            State st;
            for (uint32_t id : order) {
                const NfaState& s = nfa.states[id];
                if (s.kind == NfaState::Set) st.nfa_states.push_back(id);
                else if (s.kind == NfaState::Match) {
                    st.is_match = true;
                    break;  // leftmost-first: threads of lower priority are dropped
                }
            }
            for (uint32_t id : order) seen[id] = 0;
            order.clear();
            return st;
        };
        std::map<State, uint16_t> ids;
        std::vector<State> states;
        auto intern = [&](State st) {
            auto it = ids.find(st);
            if (it != ids.end()) return it->second;
            if (states.size() >= kMaxStates) throw RegexUnsupported("regex too large (more than 32767 DFA states)");
            const uint16_t id = (uint16_t)states.size();
            ids.emplace(st, id);
            states.push_back(std::move(st));
            return id;
        };
        intern(State());  // dead
        closure(nfa.start);
        const uint16_t start = intern(freeze());
        for (size_t at = 0; at < states.size(); ++at) {
            d.trans.resize((at + 1) * d.n_classes);
            for (uint32_t c = 0; c < d.n_classes; ++c) {
                const std::vector<uint32_t> from = states[at].nfa_states;  // (copy: `states` grows below)
                for (uint32_t id : from) {
                    const NfaState& s = nfa.states[id];
                    if (in_set[c][s.set]) closure(s.next);
                }
                const uint16_t to = intern(freeze());
                d.trans[at * d.n_classes + c] = (uint16_t)(to | (states[to].is_match ? kMatchBit : 0));
            }
        }
        d.n_states = (uint32_t)states.size();
        d.start = (uint16_t)(start | (states[start].is_match ? kMatchBit : 0));
        return d;
    }
};

// dense::Builder::new().case_insensitive(ci).build(pattern)
inline Dfa compile(const std::string& pattern, bool case_insensitive) {
    const std::unique_ptr<Ast> ast = PatternParser(pattern, case_insensitive).parse();
    return Dfa::of(Nfa::of(*ast));
}

}  // namespace vregex
