// Minimal JSON value / parser / writer used by the host side of veloci-b200.
//
// Plays the role serde_json plays in the reference (request/mod.rs derives
// Deserialize; metadata.rs:18-26 reads metaData.json).  Object keys keep their
// insertion order and can also be walked in sorted order (serde_json's default
// Map is a BTreeMap, which is what json_converter iterates).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace vjson {

struct ParseError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    bool num_is_u64 = false;   // serde_json Number::is_u64
    bool num_is_i64 = false;   // negative integer
    uint64_t u64 = 0;
    int64_t i64 = 0;
    std::string str;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;

    bool is_null() const { return kind == Null; }
    bool is_bool() const { return kind == Bool; }
    bool is_number() const { return kind == Number; }
    bool is_string() const { return kind == String; }
    bool is_array() const { return kind == Array; }
    bool is_object() const { return kind == Object; }

    const Value* get(const char* key, size_t len) const {
        if (kind != Object) return nullptr;
        for (auto& kv : obj)
            if (kv.first.size() == len && memcmp(kv.first.data(), key, len) == 0) return &kv.second;
        return nullptr;
    }
    const Value* get(const char* key) const { return get(key, strlen(key)); }
    const Value* get(const std::string& key) const { return get(key.data(), key.size()); }
    bool has(const std::string& key) const { return get(key) != nullptr; }

    static Value make_string(std::string s) {
        Value v;
        v.kind = String;
        v.str = std::move(s);
        return v;
    }
    static Value make_number(double d) {
        Value v;
        v.kind = Number;
        v.num = d;
        if (d >= 0 && d == std::floor(d) && d < 1.8e19) {
            v.num_is_u64 = true;
            v.u64 = (uint64_t)d;
        }
        return v;
    }
    static Value make_u64(uint64_t u) {
        Value v;
        v.kind = Number;
        v.num = (double)u;
        v.num_is_u64 = true;
        v.u64 = u;
        return v;
    }
    static Value make_bool(bool b) {
        Value v;
        v.kind = Bool;
        v.b = b;
        return v;
    }
    static Value make_array() {
        Value v;
        v.kind = Array;
        return v;
    }
    static Value make_object() {
        Value v;
        v.kind = Object;
        return v;
    }
    Value& set(const std::string& key, Value v) {
        kind = Object;
        for (auto& kv : obj)
            if (kv.first == key) {
                kv.second = std::move(v);
                return kv.second;
            }
        obj.emplace_back(key, std::move(v));
        return obj.back().second;
    }
};

// Shortest decimal text that round-trips the double, without exponent for the
// magnitudes that occur in documents (mirrors Rust's `f64::to_string`).
inline std::string f64_to_string(double d) {
    if (std::isnan(d)) return "NaN";
    if (std::isinf(d)) return d > 0 ? "inf" : "-inf";
    char buf[64];
    for (int prec = 1; prec <= 17; ++prec) {
        snprintf(buf, sizeof buf, "%.*g", prec, d);
        if (strtod(buf, nullptr) == d) break;
    }
    std::string s(buf);
    if (s.find('e') != std::string::npos || s.find('E') != std::string::npos) {
        // expand exponent form
        snprintf(buf, sizeof buf, "%.*f", 17, d);
        std::string best;
        for (int prec = 0; prec <= 340; ++prec) {
            char big[512];
            snprintf(big, sizeof big, "%.*f", prec, d);
            if (strtod(big, nullptr) == d) {
                best = big;
                break;
            }
        }
        if (!best.empty()) s = best;
    }
    return s;
}

class Parser {
  public:
    Parser(const char* p, size_t n) : p_(p), end_(p + n) {}

    Value parse_document() {
        skip_ws();
        Value v = parse_value();
        skip_ws();
        if (p_ != end_) fail("trailing characters");
        return v;
    }
    // Parses one value and leaves the cursor after it (for streams of values).
    bool parse_next(Value& out) {
        skip_ws();
        if (p_ == end_) return false;
        out = parse_value();
        return true;
    }

  protected:  // (the tokenizer is shared with readers that fill their own structs without a DOM)
    const char* p_;
    const char* end_;

    [[noreturn]] void fail(const char* msg) { throw ParseError(std::string("json: ") + msg); }
    void skip_ws() {
        while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) ++p_;
    }
    Value parse_value() {
        Value v;
        parse_value_into(v);
        return v;
    }
    // Values are built where they will live: no moves of half-built subtrees.
    void parse_value_into(Value& v) {
        if (p_ == end_) fail("unexpected end");
        switch (*p_) {
            case '{': parse_object(v); return;
            case '[': parse_array(v); return;
            case '"':
                v.kind = Value::String;
                parse_string(v.str);
                return;
            case 't':
                expect("true");
                v.kind = Value::Bool, v.b = true;
                return;
            case 'f':
                expect("false");
                v.kind = Value::Bool, v.b = false;
                return;
            case 'n':
                expect("null");
                return;
            default: parse_number(v);
        }
    }
    void expect(const char* lit) {
        size_t n = strlen(lit);
        if ((size_t)(end_ - p_) < n || memcmp(p_, lit, n) != 0) fail("bad literal");
        p_ += n;
    }
    void parse_number(Value& v) {
        // RFC 8259 grammar, as serde_json reads it: -? (0 | [1-9][0-9]*) (. [0-9]+)? ([eE] [+-]? [0-9]+)?
        const char* s = p_;
        auto digits = [&]() {
            const char* a = p_;
            while (p_ < end_ && *p_ >= '0' && *p_ <= '9') ++p_;
            if (p_ == a) fail("bad number");
        };
        if (p_ < end_ && *p_ == '-') ++p_;
        if (p_ < end_ && *p_ == '0') ++p_;
        else digits();
        bool is_float = false;
        if (p_ < end_ && *p_ == '.') {
            ++p_, is_float = true;
            digits();
        }
        if (p_ < end_ && (*p_ == 'e' || *p_ == 'E')) {
            ++p_, is_float = true;
            if (p_ < end_ && (*p_ == '+' || *p_ == '-')) ++p_;
            digits();
        }
        v.kind = Value::Number;
        const size_t len = (size_t)(p_ - s);
        if (!is_float && *s != '-' && len <= 15) {  // short unsigned integers: exact in a double, no strtod
            uint64_t u = 0;
            for (const char* c = s; c < p_; ++c) u = u * 10 + (uint64_t)(*c - '0');
            v.num = (double)u, v.num_is_u64 = true, v.u64 = u;
            return;
        }
        char small[48];
        std::string big;
        const char* txt = small;
        if (len < sizeof small) {
            memcpy(small, s, len);
            small[len] = 0;
        } else {
            big.assign(s, len);
            txt = big.c_str();
        }
        v.num = strtod(txt, nullptr);
        if (!is_float) {
            if (txt[0] == '-') {
                v.num_is_i64 = true;
                v.i64 = strtoll(txt, nullptr, 10);
            } else {
                v.num_is_u64 = true;
                v.u64 = strtoull(txt, nullptr, 10);
            }
        }
    }
    static void append_utf8(std::string& out, uint32_t cp) {
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
    uint32_t parse_hex4() {
        if (end_ - p_ < 4) fail("bad \\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; ++i) {
            char c = *p_++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= c - '0';
            else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
            else fail("bad hex digit");
        }
        return v;
    }
    std::string parse_string() {
        std::string out;
        parse_string(out);
        return out;
    }
    void parse_string(std::string& out) {
        ++p_;  // opening quote
        {   // the common case: no escapes before the closing quote
            const char* q = p_;
            while (q < end_ && *q != '"' && *q != '\\') ++q;
            if (q < end_ && *q == '"') {
                out.assign(p_, (size_t)(q - p_));
                p_ = q + 1;
                return;
            }
        }
        while (true) {
            if (p_ == end_) fail("unterminated string");
            char c = *p_++;
            if (c == '"') break;
            if (c == '\\') {
                if (p_ == end_) fail("bad escape");
                char e = *p_++;
                switch (e) {
                    case '"': out.push_back('"'); break;
                    case '\\': out.push_back('\\'); break;
                    case '/': out.push_back('/'); break;
                    case 'b': out.push_back('\b'); break;
                    case 'f': out.push_back('\f'); break;
                    case 'n': out.push_back('\n'); break;
                    case 'r': out.push_back('\r'); break;
                    case 't': out.push_back('\t'); break;
                    case 'u': {
                        uint32_t cp = parse_hex4();
                        if (cp >= 0xD800 && cp <= 0xDBFF && end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
                            p_ += 2;
                            uint32_t lo = parse_hex4();
                            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        }
                        append_utf8(out, cp);
                        break;
                    }
                    default: fail("bad escape char");
                }
            } else {
                out.push_back(c);
            }
        }
    }
    void parse_array(Value& v) {
        ++p_;
        v.kind = Value::Array;
        skip_ws();
        if (p_ < end_ && *p_ == ']') {
            ++p_;
            return;
        }
        v.arr.reserve(4);
        while (true) {
            skip_ws();
            v.arr.emplace_back();
            parse_value_into(v.arr.back());
            skip_ws();
            if (p_ == end_) fail("unterminated array");
            if (*p_ == ',') {
                ++p_;
                continue;
            }
            if (*p_ == ']') {
                ++p_;
                break;
            }
            fail("expected , or ]");
        }
    }
    void parse_object(Value& v) {
        ++p_;
        v.kind = Value::Object;
        skip_ws();
        if (p_ < end_ && *p_ == '}') {
            ++p_;
            return;
        }
        v.obj.reserve(4);
        while (true) {
            skip_ws();
            if (p_ == end_ || *p_ != '"') fail("expected object key");
            v.obj.emplace_back();
            parse_string(v.obj.back().first);
            skip_ws();
            if (p_ == end_ || *p_ != ':') fail("expected :");
            ++p_;
            skip_ws();
            parse_value_into(v.obj.back().second);
            // a repeated key keeps its first position and takes the last value
            const std::string& key = v.obj.back().first;
            for (size_t i = 0; i + 1 < v.obj.size(); ++i)
                if (v.obj[i].first == key) {
                    v.obj[i].second = std::move(v.obj.back().second);
                    v.obj.pop_back();
                    break;
                }
            skip_ws();
            if (p_ == end_) fail("unterminated object");
            if (*p_ == ',') {
                ++p_;
                continue;
            }
            if (*p_ == '}') {
                ++p_;
                break;
            }
            fail("expected , or }");
        }
    }
};

inline Value parse(const std::string& s) { return Parser(s.data(), s.size()).parse_document(); }
inline Value parse(const char* s, size_t n) { return Parser(s, n).parse_document(); }

inline void write_string(std::string& out, const std::string& s) {
    out.push_back('"');
    for (unsigned char c : s) {
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            default:
                if (c < 0x20) {
                    char buf[8];
                    snprintf(buf, sizeof buf, "\\u%04x", c);
                    out += buf;
                } else {
                    out.push_back((char)c);
                }
        }
    }
    out.push_back('"');
}

inline void write(std::string& out, const Value& v, int indent = -1, int depth = 0) {
    auto nl = [&](int d) {
        if (indent >= 0) {
            out.push_back('\n');
            out.append((size_t)(indent * d), ' ');
        }
    };
    switch (v.kind) {
        case Value::Null: out += "null"; break;
        case Value::Bool: out += v.b ? "true" : "false"; break;
        case Value::Number:
            if (v.num_is_u64) out += std::to_string(v.u64);
            else if (v.num_is_i64) out += std::to_string(v.i64);
            else {
                std::string s = f64_to_string(v.num);
                if (s.find('.') == std::string::npos && s.find('n') == std::string::npos && s.find('N') == std::string::npos) s += ".0";
                out += s;
            }
            break;
        case Value::String: write_string(out, v.str); break;
        case Value::Array:
            out.push_back('[');
            for (size_t i = 0; i < v.arr.size(); ++i) {
                if (i) out.push_back(',');
                nl(depth + 1);
                write(out, v.arr[i], indent, depth + 1);
            }
            if (!v.arr.empty()) nl(depth);
            out.push_back(']');
            break;
        case Value::Object:
            out.push_back('{');
            for (size_t i = 0; i < v.obj.size(); ++i) {
                if (i) out.push_back(',');
                nl(depth + 1);
                write_string(out, v.obj[i].first);
                out.push_back(':');
                if (indent >= 0) out.push_back(' ');
                write(out, v.obj[i].second, indent, depth + 1);
            }
            if (!v.obj.empty()) nl(depth);
            out.push_back('}');
            break;
    }
}

inline std::string to_string(const Value& v, int indent = -1) {
    std::string s;
    write(s, v, indent);
    return s;
}

}  // namespace vjson
