// CPU oracle (TEST INFRASTRUCTURE) for `is_regex` search parts (src/search/search_field.rs:72-83): does the reference's
// DFA accept this dictionary term?  The reference builds the DFA with regex-automata 0.1.9
// (`dense::Builder::new().case_insensitive(ci).build(pattern)`; third-party, absent from /root/reference) and runs it as
// an fst::Automaton over the term dictionary.  Restated here without ever building a DFA: the pattern is compiled to a
// prioritised thread program and a term is simulated against it, keeping the thread list of regex-automata's subset
// construction (nfa states in priority order, cut after the first match: determinize.rs `new_state`, `next`,
// `epsilon_closure`).  The list after the term's last scalar is exactly the DFA state the reference would be in.
//
// PARITY: pinned by the reference's two unit tests (search_field.rs:101-141) and the wildcard cases of
// tests/all/test_query_generator.rs:328-356; the leftmost-first cut itself has no vector in the reference: unpinned.
//
// The product's implementation (csrc/host/regex_dfa.hpp) is a different program: syntax tree -> NFA by continuation
// passing -> dense DFA over scalar classes, run on the GPU.  Shared with it: only the generated Unicode tables
// (csrc/format/case_fold.hpp).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../veloci_b200/csrc/format/case_fold.hpp"

namespace oracle_regex {

struct BadPattern : std::runtime_error {  // the reference's `.unwrap()` on the build would panic
    using std::runtime_error::runtime_error;
};
struct OutsideSubset : std::runtime_error {  // valid regex-syntax this restatement does not cover
    using std::runtime_error::runtime_error;
};

// One instruction of the thread program.
struct Inst {
    enum Op { Chars, Split, Jump, Match } op;
    std::vector<std::pair<uint32_t, uint32_t>> ranges;  // Chars: inclusive scalar ranges (unsorted is fine)
    bool negated = false;                               // Chars: the complement of `ranges`
    int x = 0, y = 0;                                   // Chars/Jump: next = x; Split: try x first, then y
};

class Program {
   public:
    Program(const std::string& pattern, bool case_insensitive) : src_(pattern), ci_(case_insensitive) {
        // unanchored: lazily skip any scalars first (lowest priority: the restart thread)
        const int split = emit(Inst{Inst::Split, {}, false, 0, 0});
        const int any = emit(Inst{Inst::Chars, {{0, 0x10FFFF}}, false, split, 0});
        code_[split].x = (int)code_.size(), code_[split].y = any;
        alternation();
        if (at_ < src_.size()) throw BadPattern("unopened group");
        emit(Inst{Inst::Match, {}, false, 0, 0});
    }

    // is the reference's DFA in a match state after the whole term (or, `starts_with`, after some prefix of it)?
    bool accepts(const std::vector<uint32_t>& term, bool starts_with) const {
        std::vector<int> cur, nxt;
        std::vector<char> on(code_.size(), 0);
        bool matched = closure(0, cur, on);
        cut(cur, matched);
        if (starts_with && matched) return true;
        for (uint32_t c : term) {
            std::fill(on.begin(), on.end(), 0);
            nxt.clear();
            bool m = false;
            for (int pc : cur) {
                const Inst& in = code_[pc];
                if (in.op == Inst::Chars && holds(in, c)) m = closure(in.x, nxt, on) || m;
            }
            cut(nxt, m);
            matched = m;
            cur.swap(nxt);
            if (starts_with && matched) return true;
            if (cur.empty() && !matched) return false;
        }
        return !starts_with && matched;
    }

   private:
    // ---- simulation
    static bool holds(const Inst& in, uint32_t c) {
        bool inside = false;
        for (auto& r : in.ranges) inside = inside || (c >= r.first && c <= r.second);
        return inside != in.negated;
    }
    // threads reachable from pc without consuming, best first, appended to `list` (Match included, as pc of the Match)
    bool closure(int pc, std::vector<int>& list, std::vector<char>& on) const {
        if (on[pc]) return false;
        on[pc] = 1;
        const Inst& in = code_[pc];
        if (in.op == Inst::Jump) return closure(in.x, list, on);
        if (in.op == Inst::Split) {
            const bool a = closure(in.x, list, on);
            const bool b = closure(in.y, list, on);
            return a || b;
        }
        list.push_back(pc);
        return in.op == Inst::Match;
    }
    // leftmost-first: nothing after the first Match survives; the Match itself is not a thread
    void cut(std::vector<int>& list, bool& matched) const {
        matched = false;
        for (size_t i = 0; i < list.size(); ++i)
            if (code_[list[i]].op == Inst::Match) {
                matched = true;
                list.resize(i);
                break;
            }
    }

    // ---- compilation: one pass over the pattern text, a repeated atom is compiled again from its source span
    int emit(Inst in) {
        if (code_.size() > 400000) throw OutsideSubset("pattern too large");
        code_.push_back(std::move(in));
        return (int)code_.size() - 1;
    }
    bool end() const { return at_ >= src_.size(); }
    uint32_t scalar() {
        const unsigned char b = (unsigned char)src_[at_];
        int n = b < 0x80 ? 1 : b >= 0xF0 ? 4 : b >= 0xE0 ? 3 : 2;
        uint32_t cp = n == 1 ? b : (b & (0xFF >> (n + 1)));
        for (int k = 1; k < n && at_ + k < src_.size(); ++k) cp = (cp << 6) | ((unsigned char)src_[at_ + k] & 0x3F);
        at_ += (size_t)n;
        return cp;
    }

    void alternation() {
        const bool ci_saved = ci_, dot_saved = dotall_;
        // a | b | c  ->  split(a, split(b, c)), every branch jumping to the common end
        std::vector<int> jumps;
        while (true) {
            const int split = emit(Inst{Inst::Split, {}, false, 0, 0});
            code_[split].x = (int)code_.size();
            concatenation();
            if (!end() && src_[at_] == '|') {
                ++at_;
                jumps.push_back(emit(Inst{Inst::Jump, {}, false, 0, 0}));
                code_[split].y = (int)code_.size();
                continue;
            }
            code_[split].op = Inst::Jump;  // the last branch has no alternative
            break;
        }
        for (int j : jumps) code_[j].x = (int)code_.size();
        ci_ = ci_saved, dotall_ = dot_saved;
    }

    void concatenation() {
        while (!end() && src_[at_] != '|' && src_[at_] != ')') {
            const size_t span_begin = at_;
            const bool ci_before = ci_, dot_before = dotall_;
            const int begin = (int)code_.size();
            if (!atom()) continue;  // "(?i)"
            size_t span_end = at_;
            int first = begin;  // the quantified expression is code_[first..]
            bool ci_item = ci_before, dot_item = dot_before;
            while (!end() && (src_[at_] == '*' || src_[at_] == '+' || src_[at_] == '?' || src_[at_] == '{')) {
                uint32_t lo, hi;
                const size_t q_begin = at_;
                quantifier(lo, hi);
                bool greedy = true;
                if (!end() && src_[at_] == '?') greedy = false, ++at_;
                // compile `lo` mandatory copies, then the optional tail, from the source text of the item
                code_.resize((size_t)first);
                const size_t resume = at_;
                const bool ci_now = ci_, dot_now = dotall_;
                auto copy = [&]() {
                    at_ = span_begin, ci_ = ci_item, dotall_ = dot_item;
                    quantified_item(span_end);
                };
                if (hi == kInf) {
                    if (lo == 0) {  // L: split(body -> L, out)
                        const int l = emit(Inst{Inst::Split, {}, false, 0, 0});
                        const int body = (int)code_.size();
                        copy();
                        emit(Inst{Inst::Jump, {}, false, l, 0});
                        set_split(l, body, (int)code_.size(), greedy);
                    } else {
                        for (uint32_t i = 1; i < lo; ++i) copy();
                        const int body = (int)code_.size();  // body; split(body, out)
                        copy();
                        const int l = emit(Inst{Inst::Split, {}, false, 0, 0});
                        set_split(l, body, (int)code_.size(), greedy);
                    }
                } else {
                    for (uint32_t i = 0; i < lo; ++i) copy();
                    std::vector<int> exits;  // (a(a(a)?)?)?: every optional copy may leave to the common end
                    for (uint32_t i = lo; i < hi; ++i) {
                        const int l = emit(Inst{Inst::Split, {}, false, 0, 0});
                        exits.push_back(l);
                        copy();
                    }
                    for (int l : exits) set_split(l, l + 1, (int)code_.size(), greedy);
                }
                at_ = resume, ci_ = ci_now, dotall_ = dot_now;
                // a further quantifier applies to everything just emitted: its "source" is the item with this quantifier
                span_end = at_;
                (void)q_begin;
            }
        }
    }
    static const uint32_t kInf = 0xFFFFFFFFu;
    void set_split(int l, int body, int out, bool greedy) {
        code_[l].op = Inst::Split;
        code_[l].x = greedy ? body : out, code_[l].y = greedy ? out : body;
    }
    // re-compiles the source span [at_, span_end): one atom followed by the quantifiers already applied to it
    void quantified_item(size_t span_end) {
        Program sub(*this, span_end);
        (void)sub;
    }
    // private "sub-compiler" constructor: compiles src_[at_ .. span_end) as a concatenation into the parent's code
    Program(Program& parent, size_t span_end) : src_(parent.src_), ci_(parent.ci_), dotall_(parent.dotall_), at_(parent.at_) {
        code_.swap(parent.code_);
        const std::string saved = src_;
        src_.resize(span_end);
        try {
            concatenation();
        } catch (...) {
            src_ = saved;
            code_.swap(parent.code_);
            throw;
        }
        src_ = saved;
        code_.swap(parent.code_);
        parent.at_ = span_end;
    }

    void quantifier(uint32_t& lo, uint32_t& hi) {
        const char c = src_[at_];
        if (c != '{') {
            ++at_;
            lo = c == '+' ? 1 : 0, hi = c == '?' ? 1 : kInf;
            return;
        }
        ++at_;
        auto number = [&](uint32_t& v) {
            size_t b = at_;
            unsigned long long acc = 0;
            while (!end() && src_[at_] >= '0' && src_[at_] <= '9') acc = acc * 10 + (unsigned)(src_[at_] - '0'), ++at_, acc = acc > 100000 ? 100000 : acc;
            v = (uint32_t)acc;
            return at_ > b;
        };
        if (!number(lo)) throw BadPattern("counted repetition without a number");
        hi = lo;
        if (!end() && src_[at_] == ',') {
            ++at_;
            if (!number(hi)) hi = kInf;
        }
        if (end() || src_[at_] != '}') throw BadPattern("unclosed counted repetition");
        ++at_;
        if (hi != kInf && hi < lo) throw BadPattern("invalid repetition range");
        if (lo > 1000 || (hi != kInf && hi > 1000)) throw OutsideSubset("counted repetition above 1000");
    }

    void fold(std::vector<std::pair<uint32_t, uint32_t>>& ranges) const {  // simple case folding closure of a set
        auto inside = [&](uint32_t c) {
            for (auto& r : ranges)
                if (c >= r.first && c <= r.second) return true;
            return false;
        };
        std::vector<uint32_t> folds;
        for (int i = 0; i < vfmt::kNumCaseFold; ++i)
            if (inside(vfmt::kCaseFold[i][0])) folds.push_back(vfmt::kCaseFold[i][1]);
        std::vector<std::pair<uint32_t, uint32_t>> more;
        for (uint32_t f : folds) more.push_back({f, f});
        for (int i = 0; i < vfmt::kNumCaseFold; ++i) {
            const uint32_t f = vfmt::kCaseFold[i][1];
            bool wanted = inside(f);
            for (uint32_t g : folds) wanted = wanted || g == f;
            if (wanted) more.push_back({vfmt::kCaseFold[i][0], vfmt::kCaseFold[i][0]});
        }
        ranges.insert(ranges.end(), more.begin(), more.end());
    }
    void emit_set(std::vector<std::pair<uint32_t, uint32_t>> ranges, bool negated) {
        if (ci_) fold(ranges);
        Inst in{Inst::Chars, std::move(ranges), negated, 0, 0};
        const int pc = emit(std::move(in));
        code_[pc].x = pc + 1;
    }

    // \d \s and single-scalar escapes; returns false and leaves `one` set for a single scalar
    bool escape_class(std::vector<std::pair<uint32_t, uint32_t>>& ranges, bool& negated, uint32_t& one, bool in_class) {
        ++at_;
        if (end()) throw BadPattern("incomplete escape");
        const char c = src_[at_];
        auto hexval = [&](int fixed) {
            uint32_t v = 0;
            int n = 0;
            if (!end() && src_[at_] == '{') {
                ++at_;
                while (!end() && src_[at_] != '}') {
                    const char h = src_[at_++];
                    const int d = h >= '0' && h <= '9' ? h - '0' : h >= 'a' && h <= 'f' ? h - 'a' + 10 : h >= 'A' && h <= 'F' ? h - 'A' + 10 : -1;
                    if (d < 0 || ++n > 8) throw BadPattern("bad hex escape");
                    v = v * 16 + (uint32_t)d;
                }
                if (end() || n == 0) throw BadPattern("bad hex escape");
                ++at_;
            } else {
                for (; n < fixed; ++n) {
                    if (end()) throw BadPattern("bad hex escape");
                    const char h = src_[at_++];
                    const int d = h >= '0' && h <= '9' ? h - '0' : h >= 'a' && h <= 'f' ? h - 'a' + 10 : h >= 'A' && h <= 'F' ? h - 'A' + 10 : -1;
                    if (d < 0) throw BadPattern("bad hex escape");
                    v = v * 16 + (uint32_t)d;
                }
            }
            if (v > 0x10FFFF || (v >= 0xD800 && v <= 0xDFFF)) throw BadPattern("not a scalar value");
            return v;
        };
        negated = false;
        switch (c) {
            case 'd': case 'D':
                ++at_;
                for (int i = 0; i < vfmt::kNumDecimalDigitRanges; ++i) ranges.push_back({vfmt::kDecimalDigitRanges[i][0], vfmt::kDecimalDigitRanges[i][1]});
                negated = c == 'D';
                return true;
            case 's': case 'S':
                ++at_;
                ranges = {{0x09, 0x0D}, {0x20, 0x20}, {0x85, 0x85}, {0xA0, 0xA0}, {0x1680, 0x1680}, {0x2000, 0x200A}, {0x2028, 0x2029}, {0x202F, 0x202F}, {0x205F, 0x205F}, {0x3000, 0x3000}};
                negated = c == 'S';
                return true;
            case 'w': case 'W': case 'p': case 'P': throw OutsideSubset("Unicode class escape");
            case 'b': case 'B': case 'A': case 'z':
                if (in_class) throw BadPattern("unrecognized escape");
                throw BadPattern("anchors / word boundaries: regex-automata 0.1 cannot build this DFA");
            case 'n': ++at_, one = '\n'; return false;
            case 'r': ++at_, one = '\r'; return false;
            case 't': ++at_, one = '\t'; return false;
            case 'f': ++at_, one = 0x0C; return false;
            case 'v': ++at_, one = 0x0B; return false;
            case 'a': ++at_, one = 0x07; return false;
            case '0': ++at_, one = 0; return false;
            case 'x': ++at_, one = hexval(2); return false;
            case 'u': ++at_, one = hexval(4); return false;
            case 'U': ++at_, one = hexval(8); return false;
            default: break;
        }
        if (strchr("\\.+*?()|[]{}^$#&-~", c) || (in_class && c == ':')) {
            ++at_, one = (unsigned char)c;
            return false;
        }
        throw BadPattern("unrecognized escape");
    }

    bool atom() {  // false: a flag group that emitted nothing
        const char c = src_[at_];
        if (c == '(') {
            ++at_;
            const bool ci_outer = ci_, dot_outer = dotall_;
            if (!end() && src_[at_] == '?') {
                ++at_;
                if (!end() && src_[at_] == 'P') {
                    ++at_;
                    if (end() || src_[at_] != '<') throw BadPattern("unrecognized flag");
                    while (!end() && src_[at_] != '>') ++at_;
                    if (end()) throw BadPattern("unclosed group name");
                    ++at_;
                } else {
                    bool on = true, ci = ci_, dot = dotall_;
                    while (!end() && src_[at_] != ':' && src_[at_] != ')') {
                        const char f = src_[at_++];
                        if (f == '-') on = false;
                        else if (f == 'i') ci = on;
                        else if (f == 's') dot = on;
                        else if (f == 'm' || f == 'x' || f == 'u' || f == 'U') throw OutsideSubset("regex flag");
                        else throw BadPattern("unrecognized flag");
                    }
                    if (end()) throw BadPattern("unclosed group");
                    ci_ = ci, dotall_ = dot;
                    if (src_[at_++] == ')') return false;
                }
            }
            alternation();  // (restores the flags it started with)
            if (end() || src_[at_] != ')') throw BadPattern("unclosed group");
            ++at_;
            ci_ = ci_outer, dotall_ = dot_outer;
            return true;
        }
        if (c == '.') {
            ++at_;
            Inst in{Inst::Chars, {}, true, 0, 0};  // everything but '\n' (or everything, with the s flag)
            if (!dotall_) in.ranges = {{'\n', '\n'}};
            const int pc = emit(std::move(in));
            code_[pc].x = pc + 1;
            return true;
        }
        if (c == '^' || c == '$') throw BadPattern("anchors: regex-automata 0.1 cannot build this DFA");
        if (c == '*' || c == '+' || c == '?' || c == '{') throw BadPattern("repetition operator missing expression");
        if (c == '[') return char_class(), true;
        if (c == '\\') {
            std::vector<std::pair<uint32_t, uint32_t>> ranges;
            bool negated;
            uint32_t one = 0;
            if (escape_class(ranges, negated, one, false)) {
                if (negated) {  // fold first, then complement
                    if (ci_) fold(ranges);
                    Inst in{Inst::Chars, std::move(ranges), true, 0, 0};
                    const int pc = emit(std::move(in));
                    code_[pc].x = pc + 1;
                } else {
                    emit_set(std::move(ranges), false);
                }
            } else {
                emit_set({{one, one}}, false);
            }
            return true;
        }
        const uint32_t cp = scalar();
        emit_set({{cp, cp}}, false);
        return true;
    }

    void char_class() {
        ++at_;
        bool negated = false;
        if (!end() && src_[at_] == '^') negated = true, ++at_;
        std::vector<std::pair<uint32_t, uint32_t>> ranges;
        std::vector<std::pair<uint32_t, uint32_t>> excluded_sets;  // \D \S inside a class: complements, expanded below
        bool first = true;
        auto one_scalar = [&](uint32_t& v) {  // a class member that is a single scalar; false for \d \s ...
            if (src_[at_] == '\\') {
                std::vector<std::pair<uint32_t, uint32_t>> set;
                bool neg;
                if (escape_class(set, neg, v, true)) {
                    if (neg) {  // complement of `set`
                        std::sort(set.begin(), set.end());
                        uint32_t next = 0;
                        for (auto& r : set) {
                            if (r.first > next) ranges.push_back({next, r.first - 1});
                            next = r.second + 1;
                        }
                        if (next <= 0x10FFFF) ranges.push_back({next, 0x10FFFF});
                    } else {
                        ranges.insert(ranges.end(), set.begin(), set.end());
                    }
                    return false;
                }
                return true;
            }
            if (src_[at_] == '[') throw OutsideSubset("nested / POSIX class");
            v = scalar();
            return true;
        };
        while (true) {
            if (end()) throw BadPattern("unclosed class");
            if (src_[at_] == ']' && !first) break;
            first = false;
            if (src_.compare(at_, 2, "&&") == 0 || src_.compare(at_, 2, "~~") == 0 || src_.compare(at_, 2, "--") == 0) throw OutsideSubset("class set operation");
            uint32_t lo = 0;
            if (!one_scalar(lo)) continue;
            uint32_t hi = lo;
            if (at_ + 1 < src_.size() && src_[at_] == '-' && src_[at_ + 1] != ']') {
                ++at_;
                if (!one_scalar(hi)) throw BadPattern("invalid class range");
                if (hi < lo) throw BadPattern("invalid class range");
            }
            ranges.push_back({lo, hi});
        }
        ++at_;
        if (ci_) fold(ranges);
        Inst in{Inst::Chars, std::move(ranges), negated, 0, 0};
        const int pc = emit(std::move(in));
        code_[pc].x = pc + 1;
        (void)excluded_sets;
    }

    std::string src_;
    bool ci_ = false, dotall_ = false;
    size_t at_ = 0;
    std::vector<Inst> code_;
};

}  // namespace oracle_regex
