// The query language of the reference's `query_parser` crate (query_parser/src/lib.rs:1-33): whitespace-separated
// terms are OR-connected, `AND` / `OR` between spaces are operators, "quoted phrases" are one term, `attr:term` and
// `attr:(...)` restrict to a field, `( )` group, `term~2` sets the edit distance.  No precedence: every operator is
// right-associative (query_parser/src/parser.rs:16-20,103-139).
//
// Layout here: one pass over the UTF-8 bytes produces tokens (byte spans into the query text), a recursive descent over
// the tokens fills an arena of nodes (indices instead of boxes); the tree is what query_generator.hpp lowers to a
// `search::Request`.  Every quirk of the reference that changes the tree is kept and marked "as the reference".
#pragma once
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace vquery {

struct ParseError : std::runtime_error {  // query_parser/src/error.rs:1-6; `what()` is the Debug text of the variant
    using std::runtime_error::runtime_error;
};

struct ParserOptions {  // query_parser/src/lib.rs:43-54
    bool no_attributes = false;
    bool no_parentheses = false;
    bool no_levensthein = false;
};

enum class TokenType : uint8_t { AttributeLiteral, Literal, ParenthesesOpen, ParenthesesClose, Tilde, Or, And };

inline const char* token_type_name(TokenType t) {
    static const char* names[] = {"AttributeLiteral", "Literal", "ParenthesesOpen", "ParenthesesClose", "Tilde", "Or", "And"};
    return names[(int)t];
}

struct Token {
    uint32_t begin, end;  // bytes of the query text
    TokenType type;
};

// char::is_whitespace: the Unicode White_Space property
inline bool is_whitespace(uint32_t c) {
    return (c >= 0x09 && c <= 0x0D) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) || c == 0x2028 || c == 0x2029 ||
           c == 0x202F || c == 0x205F || c == 0x3000;
}

// Tokens of `text` (query_parser/src/lexer.rs:107-195).
class Lexer {
   public:
    Lexer(const std::string& text, ParserOptions opt) : s_(text), opt_(opt) {}

    std::vector<Token> tokens() {
        std::vector<Token> out;
        Token t;
        while (next(t)) out.push_back(t);
        return out;
    }

   private:
    // scalar at byte `i` and its length; 0 length at the end of the text (the text is valid UTF-8: it came through the JSON reader)
    uint32_t peek(size_t i, uint32_t& len) const {
        if (i >= s_.size()) return len = 0, 0;
        const uint8_t b = (uint8_t)s_[i];
        if (b < 0x80) return len = 1, b;
        const uint32_t n = b >= 0xF0 ? 4 : b >= 0xE0 ? 3 : 2;
        uint32_t cp = b & (0xFF >> (n + 1));
        for (uint32_t k = 1; k < n && i + k < s_.size(); ++k) cp = (cp << 6) | ((uint8_t)s_[i + k] & 0x3F);
        return len = n, cp;
    }
    bool at(size_t i, char c) const { return i < s_.size() && s_[i] == c; }
    void skip_scalar() {
        uint32_t len;
        peek(pos_, len);
        prev_ws_ = false;
        pos_ += len;
    }
    bool separator(uint32_t c) const {
        return ((c == '(' || c == ')') && !opt_.no_parentheses) || (c == '~' && !opt_.no_levensthein) || (c == ':' && !opt_.no_attributes);
    }
    // `AND ` / `OR ` count as operators only after whitespace, and the byte after them must be a plain space (lexer.rs:114-124)
    bool keyword(const char* kw, size_t n) const { return prev_ws_ && s_.compare(pos_, n, kw) == 0 && at(pos_ + n, ' '); }
    TokenType literal_kind() {  // a literal directly followed by ':' names an attribute; the colon is consumed
        if (!opt_.no_attributes && at(pos_, ':')) {
            ++pos_, prev_ws_ = false;
            return TokenType::AttributeLiteral;
        }
        return TokenType::Literal;
    }

    bool next(Token& t) {
        uint32_t len;
        for (uint32_t c = peek(pos_, len); len && is_whitespace(c); c = peek(pos_, len)) pos_ += len, prev_ws_ = true;
        if (pos_ >= s_.size()) return false;
        const uint32_t start = (uint32_t)pos_;
        if (keyword("AND", 3)) return pos_ += 3, prev_ws_ = false, t = {start, (uint32_t)pos_, TokenType::And}, true;
        if (keyword("OR", 2)) return pos_ += 2, prev_ws_ = false, t = {start, (uint32_t)pos_, TokenType::Or}, true;
        if (s_[pos_] == '"') {  // a phrase: up to the next quote (or the end of the text); quotes cannot be escaped
            const size_t close = s_.find('"', pos_ + 1);
            const uint32_t stop = (uint32_t)(close == std::string::npos ? s_.size() : close);
            pos_ = close == std::string::npos ? s_.size() : close + 1;
            prev_ws_ = false;
            const TokenType kind = literal_kind();
            return t = {start + 1, stop, kind}, true;
        }
        const char c = s_[pos_];
        if ((c == '(' || c == ')') && !opt_.no_parentheses) return ++pos_, prev_ws_ = false, t = {start, start + 1, c == '(' ? TokenType::ParenthesesOpen : TokenType::ParenthesesClose}, true;
        if (c == '~' && !opt_.no_levensthein) return ++pos_, prev_ws_ = false, t = {start, start + 1, TokenType::Tilde}, true;
        skip_scalar();  // the first scalar of a literal is taken whatever it is (a lone ':' starts a literal)
        for (uint32_t d = peek(pos_, len); len && !is_whitespace(d) && !separator(d); d = peek(pos_, len)) pos_ += len;
        const uint32_t stop = (uint32_t)pos_;
        const TokenType kind = literal_kind();
        return t = {start, stop, kind}, true;
    }

    const std::string& s_;
    ParserOptions opt_;
    size_t pos_ = 0;
    bool prev_ws_ = false;  // the scalar before pos_ exists and is whitespace
};

enum class Operator : uint8_t { Or, And };

// One node of the tree (query_parser/src/ast.rs:9-14, ast/leaf.rs:1-7): children are indices into Ast::nodes.
struct Node {
    enum Kind : uint8_t { Leaf, Attributed, Binary } kind = Leaf;
    Operator op = Operator::Or;
    int32_t levenshtein = -1;  // Leaf: -1 = not given (u8 otherwise)
    int32_t left = -1;         // Attributed: the subtree; Binary: left operand
    int32_t right = -1;        // Binary: right operand
    std::string text;          // Leaf: phrase; Attributed: attribute name
};

struct Ast {
    std::vector<Node> nodes;
    int32_t root = -1;

    int32_t leaf(std::string phrase, int32_t lev = -1) {
        Node n;
        n.kind = Node::Leaf, n.text = std::move(phrase), n.levenshtein = lev;
        nodes.push_back(std::move(n));
        return (int32_t)nodes.size() - 1;
    }
    int32_t attributed(std::string attr, int32_t sub) {
        Node n;
        n.kind = Node::Attributed, n.text = std::move(attr), n.left = sub;
        nodes.push_back(std::move(n));
        return (int32_t)nodes.size() - 1;
    }
    int32_t binary(int32_t a, Operator op, int32_t b) {
        Node n;
        n.kind = Node::Binary, n.op = op, n.left = a, n.right = b;
        nodes.push_back(std::move(n));
        return (int32_t)nodes.size() - 1;
    }

    // the Debug text of the reference's UserAST (ast.rs:51-59, ast/leaf.rs:9-17): what its parser tests compare
    void debug(int32_t i, std::string& out) const {
        const Node& n = nodes[i];
        if (n.kind == Node::Leaf) {
            out += '"', out += n.text, out += '"';
            if (n.levenshtein >= 0) out += '~', out += std::to_string(n.levenshtein);
        } else if (n.kind == Node::Attributed) {
            out += n.text, out += ':';
            debug(n.left, out);
        } else {
            out += '(';
            debug(n.left, out);
            out += n.op == Operator::Or ? " OR " : " AND ";
            debug(n.right, out);
            out += ')';
        }
    }
    std::string debug() const {
        std::string out;
        if (root >= 0) debug(root, out);
        return out;
    }

    // UserAST::walk_terms (ast.rs:147-162): the phrases in text order
    void walk_terms(int32_t i, const std::function<void(const std::string&)>& cb) const {
        const Node& n = nodes[i];
        if (n.kind == Node::Leaf) return cb(n.text);
        walk_terms(n.left, cb);
        if (n.kind == Node::Binary) walk_terms(n.right, cb);
    }

    // UserAST::get_phrase_pairs (ast.rs:118-145): adjacent terms, for phrase boosts.  A set in the reference; here the
    // distinct pairs in order of first appearance.  Leaving an attribute for a different one starts a new run; the run
    // is NOT cut when an attributed subtree ends (as the reference: "a:(x y) z" pairs y with z).
    std::vector<std::pair<std::string, std::string>> phrase_pairs() const {
        std::vector<std::pair<std::string, std::string>> out;
        int32_t last = -1;
        if (root >= 0) pairs(root, -1, last, out);
        return out;
    }

    // UserAST::filter_ast (ast.rs:68-95): the tree without the subtrees `drop(node, attribute or nullptr)` names;
    // -1 when nothing is left.  Nodes are appended to this arena.
    int32_t filter(int32_t i, const std::function<bool(const Ast&, int32_t, const std::string*)>& drop, const std::string* attr = nullptr) {
        if (drop(*this, i, attr)) return -1;
        const Node n = nodes[i];
        if (n.kind == Node::Attributed) {
            const int32_t sub = filter(n.left, drop, &n.text);
            return sub < 0 ? -1 : attributed(n.text, sub);
        }
        if (n.kind == Node::Binary) {
            const int32_t a = filter(n.left, drop, attr), b = filter(n.right, drop, attr);
            if (a >= 0 && b >= 0) return binary(a, n.op, b);
            return a >= 0 ? a : b;
        }
        return i;
    }

   private:
    void pairs(int32_t i, int32_t cur_attr, int32_t& last, std::vector<std::pair<std::string, std::string>>& out) const {
        const Node& n = nodes[i];
        if (n.kind == Node::Attributed) {
            if (cur_attr < 0 || nodes[cur_attr].text == n.text) {
                pairs(n.left, i, last, out);
            } else {
                int32_t fresh = -1;
                pairs(n.left, i, fresh, out);
            }
        } else if (n.kind == Node::Binary) {
            pairs(n.left, cur_attr, last, out);
            pairs(n.right, cur_attr, last, out);
        } else {
            if (last >= 0) {
                std::pair<std::string, std::string> p(nodes[last].text, n.text);
                bool seen = false;
                for (auto& q : out) seen = seen || q == p;
                if (!seen) out.push_back(std::move(p));
            }
            last = i;
        }
    }
};

class Parser {
   public:
    Parser(const std::string& text, ParserOptions opt) : s_(text), toks_(Lexer(text, opt).tokens()) {}

    Ast parse() {
        ast_.root = expression(0);
        return std::move(ast_);  // tokens after an unmatched ')' are dropped, as the reference (parser.rs:130,234: "\"cool\")" parses)
    }

   private:
    static const int kMaxDepth = 2000;  // the reference recurses without a bound; a bound here instead of a stack overflow

    bool is(TokenType t) const { return pos_ < toks_.size() && toks_[pos_].type == t; }
    bool at_end() const { return pos_ >= toks_.size(); }
    std::string text_of(const Token& t) const { return s_.substr(t.begin, t.end - t.begin); }

    // ParseError::UnexpectedTokenType(marked text, message) (parser.rs:44-66, error.rs:8-10)
    [[noreturn]] void unexpected(const std::string& message, const char* allowed) const {
        const size_t b = at_end() ? s_.size() : toks_[pos_].begin, e = at_end() ? s_.size() : toks_[pos_].end;
        const std::string marked = s_.substr(0, b) + "\xEF\xB9\x8F" + s_.substr(b, e - b) + "\xEF\xB9\x8F" + s_.substr(e);
        std::string msg = message;
        if (msg.empty()) {
            msg = std::string(" Unexpected token_type, got ") + (at_end() ? "EOF" : token_type_name(toks_[pos_].type));
            msg += allowed ? std::string("\" allowed_types: ") + allowed + "\"" : std::string("\"\"");  // an Option<String> printed with {:?} in the reference
        }
        throw ParseError("UnexpectedTokenType(" + quoted(marked) + ", " + quoted(msg) + ")");
    }
    static std::string quoted(const std::string& s) {  // Rust's {:?} of a str, for the characters that occur here
        std::string out = "\"";
        for (char c : s) {
            if (c == '"' || c == '\\') out += '\\';
            if (c == '\n') {
                out += "\\n";
                continue;
            }
            out += c;
        }
        return out + "\"";
    }

    // a term with its optional "~n" (parser.rs:80-101)
    int32_t user_filter(const Token& tok) {
        int32_t lev = -1;
        if (is(TokenType::Tilde)) {
            ++pos_;
            if (!is(TokenType::Literal)) unexpected("Expecting a levenshtein number after a '~' ", nullptr);
            const Token num = toks_[pos_++];
            const std::string digits = text_of(num);
            size_t i = !digits.empty() && digits[0] == '+' ? 1 : 0;  // u8::from_str: an optional '+', then digits, at most 255
            int32_t v = 0;
            bool ok = i < digits.size();
            for (; ok && i < digits.size(); ++i) {
                ok = digits[i] >= '0' && digits[i] <= '9' && (v = v * 10 + (digits[i] - '0')) <= 255;
            }
            if (!ok)
                throw ParseError("ExpectedNumber(\"Expected number after tilde to define levenshtein distance but got Token { byte_start_pos: " + std::to_string(num.begin) +
                                 ", byte_stop_pos: " + std::to_string(num.end) + ", token_type: Literal }\")");
            lev = v;
        }
        return ast_.leaf(text_of(tok), lev);
    }

    // what may follow a complete operand (parser.rs:103-139)
    int32_t continuation(int32_t left, int depth) {
        if (at_end()) return left;
        switch (toks_[pos_].type) {
            case TokenType::AttributeLiteral:
            case TokenType::Literal:
                return ast_.binary(left, Operator::Or, expression(depth + 1));
            case TokenType::Or:
                ++pos_;
                return ast_.binary(left, Operator::Or, expression(depth + 1));
            case TokenType::And:
                ++pos_;
                return ast_.binary(left, Operator::And, expression(depth + 1));
            case TokenType::ParenthesesClose:
                return left;  // left for the caller that opened it
            case TokenType::ParenthesesOpen:  // "a (b)": unimplemented!() in the reference
                throw ParseError("Unimplemented(\"an opening parenthesis directly after a term\")");
            case TokenType::Tilde:
                unexpected("", "[Some(AttributeLiteral), Some(Literal), Some(ParenthesesOpen), Some(ParenthesesClose), Some(And), Some(Or), None]");
        }
        return left;
    }

    int32_t expression(int depth) {  // parser.rs:141-190
        if (depth > kMaxDepth) throw ParseError("TooDeep(\"query nests deeper than " + std::to_string(kMaxDepth) + "\")");
        if (at_end()) throw ParseError("UnexpectedEnd(\"the query ends where a term is expected\")");  // a panic in the reference (parser.rs:75)
        const Token tok = toks_[pos_++];
        switch (tok.type) {
            case TokenType::AttributeLiteral: {
                if (is(TokenType::ParenthesesOpen)) return ast_.attributed(text_of(tok), expression(depth + 1));  // the attribute covers all that follows, as the reference
                if (is(TokenType::Literal)) {
                    const Token term = toks_[pos_++];
                    const int32_t leaf = user_filter(term);
                    return continuation(ast_.attributed(text_of(tok), leaf), depth);
                }
                unexpected("only token or ( allowed after attribute ('attr:') ", nullptr);
            }
            case TokenType::Literal:
                return continuation(user_filter(tok), depth);
            case TokenType::ParenthesesOpen: {
                const int32_t inner = expression(depth + 1);
                if (!is(TokenType::ParenthesesClose)) unexpected("", "[Some(ParenthesesClose)]");
                ++pos_;
                return continuation(inner, depth);
            }
            case TokenType::Tilde:
                unexpected("", nullptr);  // (marks the token after the tilde, as the reference)
            case TokenType::ParenthesesClose:
            case TokenType::Or:
            case TokenType::And:  // unimplemented!() in the reference
                throw ParseError(std::string("Unimplemented(\"") + token_type_name(tok.type) + " where a term is expected\")");
        }
        throw ParseError("Unreachable");
    }

    const std::string& s_;
    std::vector<Token> toks_;
    size_t pos_ = 0;
    Ast ast_;
};

// query_parser::parse_with_opt (parser.rs:26-28)
inline Ast parse(const std::string& text, ParserOptions opt = ParserOptions()) { return Parser(text, opt).parse(); }

}  // namespace vquery
