"""CPU oracle (TEST INFRASTRUCTURE, never imported by the product): plain-Python restatement of the reference's
why_found highlighting on the stored document (src/highlight_field.rs) with its tokenizer
(src/tokenizer/simple_tokenizer_group.rs).  Pinned against tests/all/test_why_found.rs in tests/test_highlight.py.
The product's implementation is csrc/host/highlight.hpp."""

DEFAULT_SEPERATORS = [" ", "\t", "\n", "\r", ":", "(", ")", ",", ".", "…", ";", "・", "’", "—", "-", "\\", "[", "]", "{", "}", "<", ">", "'", '"', "“", "™"]  # tokenizer/mod.rs:17-19


def tokenize(text, seperators):
    """SimpleTokenizerGroupTokenIter (simple_tokenizer_group.rs:48-82): [(piece, is_seperator)]"""
    out = []
    last_returned = 0
    last_was_token = False
    for pos, ch in enumerate(text):
        if ch in seperators:
            if pos == 0:
                last_was_token = True
            elif not last_was_token:
                out.append((text[last_returned:pos], False))
                last_was_token = True
                last_returned = pos
        elif last_was_token:
            out.append((text[last_returned:pos], True))
            last_was_token = False
            last_returned = pos
    if last_returned != len(text):
        out.append((text[last_returned:], last_was_token))
    return out


def group_hit_positions_for_snippet(hit_pos, num_words_around):  # highlight_field.rs:19-38
    around = num_words_around * 2
    grouped = []
    previous = -around
    for pos in hit_pos:
        if pos - previous >= around:
            grouped.append([])
        previous = pos
        grouped[-1].append(pos)
    return grouped


def highlight_text(text, terms, tokenizer_seperators, num_words_around=5, start="<b>", end="</b>", connector=" ... "):
    """highlight_field.rs:98-141; `tokenizer_seperators` None = the field has no tokenizer"""
    if len(terms) == 1 and text in terms:
        return start + text + end
    if tokenizer_seperators is None:
        return None
    tokens = [t for t, _ in tokenize(text, tokenizer_seperators)]
    hit_pos = [i for i, t in enumerate(tokens) if t in terms]
    around = num_words_around * 2
    parts = []
    for group in group_hit_positions_for_snippet(hit_pos, num_words_around):
        lo = max(group[0] - around, 0)                      # grouped_to_positions_for_snippet :40-44
        hi = min(group[-1] + around + 1, len(tokens))
        parts.append("".join(start + tokens[i] + end if tokens[i] in terms else tokens[i] for i in range(lo, hi)))  # build_snippet :46-77
    snippet = connector.join(parts)
    if hit_pos:                                              # ellipsis_snippet :80-96
        if hit_pos[0] > around:
            snippet = connector + snippet
        if hit_pos[-1] < len(tokens) - around:
            snippet = snippet + connector
    return snippet if hit_pos else None


def _texts(value, path, name, out):
    """json_converter::for_each_element's text callback: (text, field path) in document order"""
    if isinstance(value, list):
        for el in value:
            _texts(el, path + name + "[]", "", out)
    elif isinstance(value, dict):
        prefix = path + name
        if prefix:
            prefix += "."
        for k, v in value.items():
            _texts(v, prefix, k, out)
    elif value is not None:
        if isinstance(value, bool):
            text = "true" if value else "false"
        else:
            text = value if isinstance(value, str) else repr(value) if isinstance(value, float) else str(value)
        out.append((text, path + name))


def highlight_on_original_document(columns, doc, why_found_terms):
    """highlight_field.rs:148-186.  `columns`: metaData.json's columns; `why_found_terms`: path.textindex -> set of terms"""
    out = {}
    texts = []
    _texts(doc, "", "", texts)
    for text, field in texts:
        terms = why_found_terms.get(field + ".textindex")
        if terms is None:
            continue
        options = columns[field]["textindex_metadata"]["options"]
        seps = None
        if options["tokenize"]:
            seps = options["tokenize_on_chars"] if options.get("tokenize_on_chars") is not None else DEFAULT_SEPERATORS
        h = highlight_text(text, set(terms), seps)
        if h is not None:
            out.setdefault(field, []).append(h)
    return out
