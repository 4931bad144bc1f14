"""CPU oracle (TEST INFRASTRUCTURE, never imported by the product) for the request-generation row of the scope table:
a plain-Python restatement of the reference's `query_parser` crate and of `query_generator::search_query` /
`suggest_query`.  Only tests import this file; the product's own implementation is csrc/host/query_parser.hpp and
csrc/host/query_generator.hpp.

Pinned against the reference's own unit vectors (query_parser/src/lexer.rs:248-325, parser.rs:203-480,
ast.rs:166-312, src/query_generator/query_parser_to_veloci_request.rs:181-223) in tests/test_query_generator.py.

Every function cites the reference lines it follows.  Where the reference walks a hash map (columns, boost_terms, the
phrase-pair set) the order is unspecified there; this file uses key order for the maps and first appearance for the
pairs, and the tests compare those as sets.
"""
import json
import struct

# ------------------------------------------------------------------------------------------------ lexer (lexer.rs)

ATTRIBUTE_LITERAL, LITERAL, PAREN_OPEN, PAREN_CLOSE, TILDE, OR, AND = "AttributeLiteral", "Literal", "ParenthesesOpen", "ParenthesesClose", "Tilde", "Or", "And"


class ParseError(Exception):
    """query_parser/src/error.rs:1-6; str() is the Debug text of the variant"""


class Options:  # query_parser/src/lib.rs:43-54
    def __init__(self, no_attributes=False, no_parentheses=False, no_levensthein=False):
        self.no_attributes, self.no_parentheses, self.no_levensthein = no_attributes, no_parentheses, no_levensthein


def _is_whitespace(c):  # char::is_whitespace = White_Space; str.isspace() also takes U+001C..U+001F, which Rust does not
    return c.isspace() and not ("\x1c" <= c <= "\x1f")


def _single_char_type(c, o):  # lexer.rs:24-35
    if c == "(" and not o.no_parentheses:
        return PAREN_OPEN
    if c == ")" and not o.no_parentheses:
        return PAREN_CLOSE
    if c == "~" and not o.no_levensthein:
        return TILDE
    return None


def _is_separator(c, o):  # lexer.rs:37-44
    return (c in "()" and not o.no_parentheses) or (c == "~" and not o.no_levensthein) or (c == ":" and not o.no_attributes)


def lex(text, o=None):
    """lexer.rs:107-195: [(byte_start, byte_stop, type)]"""
    o = o or Options()
    chars = list(text)
    pos = 0
    byte_pos = 0
    tokens = []

    def eat():
        nonlocal pos, byte_pos
        if pos < len(chars):
            byte_pos += len(chars[pos].encode("utf-8"))
            pos += 1

    def attr_colon():
        return not o.no_attributes and pos < len(chars) and chars[pos] == ":"

    while True:
        while pos < len(chars) and _is_whitespace(chars[pos]):
            eat()
        if pos >= len(chars):
            return tokens
        c = chars[pos]
        start = byte_pos
        prev_ws = pos != 0 and _is_whitespace(chars[pos - 1])
        ttype = None
        if chars[pos:pos + 4] == ["A", "N", "D", " "] and prev_ws:
            eat(), eat(), eat()
            ttype = AND
        elif chars[pos:pos + 3] == ["O", "R", " "] and prev_ws:
            eat(), eat()
            ttype = OR
        if pos < len(chars) and chars[pos] == '"':
            eat()
            start += 1
            while pos < len(chars) and chars[pos] != '"':
                eat()
            stop = byte_pos
            eat()
            if attr_colon():
                eat()
                tokens.append((start, stop, ATTRIBUTE_LITERAL))
            else:
                tokens.append((start, stop, LITERAL))
            continue
        single = _single_char_type(c, o)
        if single:
            ttype = single
            eat()
        if ttype:
            tokens.append((start, byte_pos, ttype))
            continue
        eat()
        while pos < len(chars) and not _is_whitespace(chars[pos]) and not _is_separator(chars[pos], o):
            eat()
        stop = byte_pos
        if attr_colon():
            eat()
            tokens.append((start, stop, ATTRIBUTE_LITERAL))
        else:
            tokens.append((start, stop, LITERAL))


# ------------------------------------------------------------------------------------------------ tree (ast.rs)
# ("leaf", phrase, levenshtein or None) | ("attr", name, tree) | ("bin", left, "OR" | "AND", right)


def debug(ast):  # ast.rs:51-59, ast/leaf.rs:9-17, ast/operator.rs:6-13
    if ast[0] == "leaf":
        return '"%s"' % ast[1] + ("" if ast[2] is None else "~%d" % ast[2])
    if ast[0] == "attr":
        return "%s:%s" % (ast[1], debug(ast[2]))
    return "(%s %s %s)" % (debug(ast[1]), ast[2], debug(ast[3]))


def filter_ast(ast, should_filter, attr=None):  # ast.rs:68-95
    if should_filter(ast, attr):
        return None
    if ast[0] == "attr":
        sub = filter_ast(ast[2], should_filter, ast[1])
        return None if sub is None else ("attr", ast[1], sub)
    if ast[0] == "bin":
        a, b = filter_ast(ast[1], should_filter, attr), filter_ast(ast[3], should_filter, attr)
        if a is not None and b is not None:
            return ("bin", a, ast[2], b)
        return a if a is not None else b
    return ast


def phrase_pairs(ast):  # ast.rs:118-145 (a HashSet there; first-appearance order here)
    out = []

    def walk(node, last, cur_attr):  # `last` is a one-element list: the shared &mut Option<&str>
        if node[0] == "attr":
            if cur_attr is None or cur_attr == node[1]:
                walk(node[2], last, node[1])
            else:
                walk(node[2], [None], node[1])
        elif node[0] == "bin":
            walk(node[1], last, cur_attr)
            walk(node[3], last, cur_attr)
        else:
            if last[0] is not None and (last[0], node[1]) not in out:
                out.append((last[0], node[1]))
            last[0] = node[1]

    walk(ast, [None], None)
    return out


def walk_terms(ast):  # ast.rs:147-162
    if ast[0] == "leaf":
        return [ast[1]]
    if ast[0] == "attr":
        return walk_terms(ast[2])
    return walk_terms(ast[1]) + walk_terms(ast[3])


# ------------------------------------------------------------------------------------------------ parser (parser.rs)


def _rust_str_debug(s):
    return '"' + s.replace("\\", "\\\\").replace('"', '\\"').replace("\n", "\\n") + '"'


class _Parser:
    def __init__(self, text, o):
        self.text = text
        self.raw = text.encode("utf-8")
        self.tokens = lex(text, o)
        self.pos = 0

    def tok_text(self, t):
        return self.raw[t[0]:t[1]].decode("utf-8")

    def get_type(self):
        return self.tokens[self.pos][2] if self.pos < len(self.tokens) else None

    def unexpected(self, message, allowed):  # parser.rs:44-66
        if self.pos < len(self.tokens):
            start, stop = self.tokens[self.pos][0], self.tokens[self.pos][1]
        else:
            start = stop = len(self.raw)
        marked = (self.raw[:start] + "﹏".encode() + self.raw[start:stop] + "﹏".encode() + self.raw[stop:]).decode("utf-8")
        if message == "":
            got = self.get_type() or "EOF"
            allowed_text = "" if allowed is None else " allowed_types: [%s]" % ", ".join("None" if a is None else "Some(%s)" % a for a in allowed)
            message = " Unexpected token_type, got %s%s" % (got, _rust_str_debug(allowed_text))
        raise ParseError("UnexpectedTokenType(%s, %s)" % (_rust_str_debug(marked), _rust_str_debug(message)))

    def next_token(self):  # parser.rs:74-78 (`.unwrap()` on a missing token: a panic there, an error here)
        if self.pos >= len(self.tokens):
            raise ParseError("panic: next_token at the end of the query")
        t = self.tokens[self.pos]
        self.pos += 1
        return t

    def user_filter(self, tok):  # parser.rs:80-101
        lev = None
        if self.get_type() == TILDE:
            self.next_token()
            if self.get_type() != LITERAL:
                self.unexpected("Expecting a levenshtein number after a '~' ", [LITERAL])
            lt = self.next_token()
            digits = self.tok_text(lt)
            body = digits[1:] if digits.startswith("+") else digits
            if not body or any(ch not in "0123456789" for ch in body) or int(body) > 255:  # u8::from_str
                raise ParseError('ExpectedNumber("Expected number after tilde to define levenshtein distance but got Token { byte_start_pos: %d, byte_stop_pos: %d, token_type: Literal }")' % (lt[0], lt[1]))
            lev = int(body)
        return ("leaf", self.tok_text(tok), lev)

    def sub_expression(self, cur):  # parser.rs:103-139
        allowed = [ATTRIBUTE_LITERAL, LITERAL, PAREN_OPEN, PAREN_CLOSE, AND, OR, None]
        t = self.get_type()
        if t not in allowed:
            self.unexpected("", allowed)
        if t in (ATTRIBUTE_LITERAL, LITERAL):
            return ("bin", cur, "OR", self.parse())
        if t == OR:
            self.next_token()
            return ("bin", cur, "OR", self.parse())
        if t == AND:
            self.next_token()
            return ("bin", cur, "AND", self.parse())
        if t == PAREN_OPEN:
            raise ParseError("panic: unimplemented (ParenthesesOpen after an operand)")
        return cur  # ParenthesesClose or the end

    def parse(self):  # parser.rs:141-190
        tok = self.next_token()
        if tok[2] == ATTRIBUTE_LITERAL:
            nt = self.get_type()
            if nt == PAREN_OPEN:
                return ("attr", self.tok_text(tok), self.parse())
            if nt == LITERAL:
                leaf = self.user_filter(self.next_token())
                return self.sub_expression(("attr", self.tok_text(tok), leaf))
            self.unexpected("only token or ( allowed after attribute ('attr:') ", [LITERAL, PAREN_OPEN])
        if tok[2] == LITERAL:
            return self.sub_expression(self.user_filter(tok))
        if tok[2] == PAREN_OPEN:
            inner = self.parse()
            if self.get_type() != PAREN_CLOSE:
                self.unexpected("", [PAREN_CLOSE])
            self.next_token()
            return self.sub_expression(inner)
        if tok[2] == TILDE:
            self.unexpected("", None)
        raise ParseError("panic: unimplemented (%s where an operand is expected)" % tok[2])


def parse(text, o=None):  # parser.rs:23-28
    return _Parser(text, o or Options()).parse()


# ------------------------------------------------------------------------------------------------ generator


class GeneratorError(Exception):
    """VelociError::FieldNotFound / AllFieldsFiltered (src/error.rs:13-17)"""


def f32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


def default_levenshtein(term, auto_limit, wildcard):  # src/query_generator.rs:85-99
    n = len(term)
    if wildcard:
        return 0 if n <= 3 else min(1, auto_limit) if n <= 5 else min(2, auto_limit)
    return 0 if n <= 2 else min(1, auto_limit) if n <= 5 else min(2, auto_limit)


def get_levenshtein(term, levenshtein, auto_limit, wildcard):  # src/query_generator.rs:129-132
    lev = levenshtein if levenshtein is not None else default_levenshtein(term, 1 if auto_limit is None else auto_limit, wildcard)
    cap = (len(term) - 1) % (1 << 64)  # usize arithmetic of a release build
    return min(lev, cap)


REGEX_META = set("\\.+*?()|[]{}^$#&-~")  # regex_syntax::is_meta_character


def regex_escape(s):
    return "".join("\\" + c if c in REGEX_META else c for c in s)


class Catalog:
    """`all_fields` = metadata.get_all_fields(); `search_fields` = those with a token_to_anchor index
    (src/metadata.rs:28-30, src/persistence.rs:329-332)"""

    def __init__(self, all_fields, search_fields):
        self.all_fields, self.search_fields = sorted(all_fields), sorted(search_fields)


def search_field_names(cat, whitelist):  # src/query_generator.rs:101-127
    res = [f for f in cat.all_fields if f in whitelist] if whitelist is not None else list(cat.search_fields)
    if not res:
        raise GeneratorError("All fields filtered all_fields: %s filter: %s" % (json.dumps(cat.all_fields, ensure_ascii=False), "None" if whitelist is None else "Some(%s)" % json.dumps(whitelist, ensure_ascii=False)))
    return res


def check_field(field, all_fields):  # src/query_generator.rs:134-144
    if field not in all_fields:
        raise GeneratorError("Field %s not found in %s" % (field, json.dumps(all_fields, ensure_ascii=False)))


def expand_fields(ast, fields):  # query_parser_to_veloci_request.rs:85-114
    if ast[0] == "bin":
        return ("bin", expand_fields(ast[1], fields), ast[2], expand_fields(ast[3], fields))
    if ast[0] == "leaf":
        cur = ("attr", fields[0], ast)
        for name in fields[1:]:
            cur = ("bin", ("attr", name, ast), "OR", cur)
        return cur
    check_field(ast[1], fields)
    return ast


def ast_to_request(ast, opt, field_name=None):  # query_parser_to_veloci_request.rs:23-83
    if ast[0] == "bin":
        return {"and" if ast[2] == "AND" else "or": {"queries": [ast_to_request(ast[1], opt, field_name), ast_to_request(ast[3], opt, field_name)]}}
    if ast[0] == "attr":
        return ast_to_request(ast[2], opt, ast[1])
    term = ast[1]
    part = {"path": field_name}
    starts_with = term.endswith("*") and term.count("*") == 1
    if starts_with:
        term = term[:-1]
    is_regex = "*" in term
    lev = None
    if is_regex:
        term = ".*".join(regex_escape(p) for p in term.split("*"))
    else:
        lev = ast[2] if ast[2] is not None else get_levenshtein(term, opt.get("levenshtein"), opt.get("levenshtein_auto_limit"), starts_with)
    part["terms"] = [term]
    if lev is not None:
        part["levenshtein_distance"] = lev
    if starts_with:
        part["starts_with"] = True
    if is_regex:
        part["is_regex"] = True
    boost = (opt.get("boost_fields") or {}).get(field_name)
    if boost is not None:
        part["boost"] = f32(boost)
    if opt.get("ignore_case") is not None:
        part["ignore_case"] = opt["ignore_case"]
    return {"search": part}


def simplify(req):  # src/search/request/search_request.rs:27-76
    for kind in ("or", "and"):
        if kind in req:
            tree = req[kind]
            for q in tree["queries"]:
                simplify(q)
            lifted = []
            for i in reversed(range(len(tree["queries"]))):
                q = tree["queries"][i]
                if kind in q and q[kind].get("options") is None:
                    lifted.extend(tree["queries"].pop(i)[kind]["queries"])
            tree["queries"].extend(lifted)


def ast_to_search_request(ast, fields, opt):  # query_parser_to_veloci_request.rs:11-15 (the stop-word filter's result is dropped there)
    req = ast_to_request(expand_fields(ast, fields), opt)
    simplify(req)
    return req


def boost_term_parts(cat, spec, value):  # src/query_generator.rs:146-168
    term, only = spec, None
    if ":" in spec:
        pieces = spec.split(":")
        term = pieces.pop(1)
        only = pieces
    return [{"path": f, "terms": [term], "boost": f32(value)} for f in search_field_names(cat, only)]


def phrase_boosts(cat, fields, pairs, levenshtein, auto_limit, boost_fields):  # src/query_generator.rs:268-295
    out = []
    for a, b in pairs:
        for f in search_field_names(cat, fields):
            def part(t):
                p = {"path": f, "terms": [t], "levenshtein_distance": get_levenshtein(t, levenshtein, auto_limit, False)}
                if boost_fields and f in boost_fields:
                    p["boost"] = f32(boost_fields[f])
                return p
            out.append({"search1": part(a), "search2": part(b)})
    return out


def _options(d):
    d = d or {}
    return Options(d.get("no_attributes", False), d.get("no_parentheses", False), d.get("no_levensthein", False))


def search_query(cat, opt):
    """src/query_generator.rs:175-257; `opt` = SearchQueryGeneratorParameters as a dict.  Returns the Request as the dict
    `serde_json::to_value(&request)` would give (keys the serde derive skips are absent, `select` is always there)."""
    facetlimit = opt.get("facetlimit") if opt.get("facetlimit") is not None else 5
    search_fields = search_field_names(cat, opt.get("fields"))
    ast = parse(opt.get("search_term", ""), _options(opt.get("parser_options")))
    req = {"search_req": ast_to_search_request(ast, search_fields, opt)}
    if opt.get("boost_queries") is not None:
        req["boost"] = [{"path": b["path"], "boost_fun": b.get("boost_fun"), "param": None if b.get("param") is None else f32(b["param"]),
                         "skip_when_score": None if b.get("skip_when_score") is None else [f32(x) for x in b["skip_when_score"]], "expression": b.get("expression")} for b in opt["boost_queries"]]
    if opt.get("boost_terms") is not None:
        req["boost_term"] = [p for spec in sorted(opt["boost_terms"]) for p in boost_term_parts(cat, spec, opt["boost_terms"][spec])]
    if opt.get("facets") is not None:
        for f in opt["facets"]:
            check_field(f, cat.all_fields)
        req["facets"] = [{"field": f, "top": facetlimit} for f in opt["facets"]]
    pairs = phrase_pairs(ast)
    if opt.get("phrase_pairs") and pairs:
        req["phrase_boosts"] = phrase_boosts(cat, opt.get("fields"), pairs, opt.get("levenshtein"), opt.get("levenshtein_auto_limit"), opt.get("boost_fields"))
    req["select"] = None
    if opt.get("filter") is not None:
        fast = parse(opt["filter"], _options(opt.get("filter_parser_options")))
        req["filter"] = ast_to_search_request(fast, cat.all_fields, {"levenshtein": 0})
    if opt.get("top") is not None:
        req["top"] = opt["top"]
    if opt.get("skip") is not None:
        req["skip"] = opt["skip"]
    for flag in ("why_found", "text_locality", "explain"):
        if opt.get(flag):
            req[flag] = True
    return req


def suggest_query(cat, request, top=None, skip=None, levenshtein=None, fields=None, levenshtein_auto_limit=None):  # src/query_generator.rs:297-322
    if top is None:
        top = 10
    parts = []
    for f in search_field_names(cat, fields):
        lev = levenshtein if levenshtein is not None else default_levenshtein(request, 1 if levenshtein_auto_limit is None else levenshtein_auto_limit, True)
        p = {"path": f, "terms": [request], "levenshtein_distance": lev, "starts_with": True, "top": top}
        if skip is not None:
            p["skip"] = skip
        parts.append(p)
    req = {"suggest": parts, "select": None, "top": top}
    if skip is not None:
        req["skip"] = skip
    return req
