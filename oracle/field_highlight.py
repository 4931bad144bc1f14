"""CPU oracle (TEST INFRASTRUCTURE, never imported by the product): search_field::highlight
(src/search/search_field.rs:232-245) = get_term_ids_in_field + resolve_token_hits_to_text_id with snippets (:549-636) +
get_text_score_id_from_result(false, ..) (:160-190), in plain Python.  The part's term hits come from the C++ oracle's
get_term_ids_in_field; the joins and the snippets run over the oracle's own decoder of the index files
(oracle/read_document.py).  The product's implementation is csrc/host/field_highlight.hpp."""
import re

_REPLACEMENTS = [  # util::normalize_text, src/util.rs:11-30
    (re.compile(r"\([fmn\d]\)"), " "),
    (re.compile(r"[\(\)]"), " "),
    (re.compile("[{}'\"“]"), ""),
    (re.compile(r"\s\s+"), " "),
    (re.compile("[,.…;・’-]"), ""),
]


def normalize_text(text):
    for rx, to in _REPLACEMENTS:
        text = rx.sub(to, text)
    return text.lower().strip()


def highlight(reader, get_term_ids_in_field, part):
    """`get_term_ids_in_field(part)` -> [(term id, score)] in hits_scores order; `reader`: oracle/read_document.Reader."""
    part = dict(part)
    part["terms"] = [normalize_text(t) for t in part["terms"]]  # :234
    snippet = part.pop("snippet", None) or False
    info = part.pop("snippet_info", None) or {}
    hits = get_term_ids_in_field(part)
    path = part["path"] if part["path"].endswith(".textindex") else part["path"] + ".textindex"
    token_hits = []
    for term_id, score in hits:  # :571-584
        for parent in reader.get_values(path + ".tokens_to_text_id", term_id) or []:
            token_hits.append((parent, score, term_id))
    token_hits.sort(key=lambda h: h[0])
    assert snippet, "without snippets the reference indexes an empty highlight map"
    out = []
    i = 0
    while i < len(token_hits):  # :605-632, hits_scores cleared first (:599-601)
        j = i
        while j < len(token_hits) and token_hits[j][0] == token_hits[i][0]:
            j += 1
        group = token_hits[i:j]
        best = group[0][1]
        for _, score, _ in group:  # max_by_key(|score|): the last of equal maxima
            if abs(score) >= abs(best):
                best = score
        text = reader.highlight_document(path, group[0][0], {t for _, _, t in group}, num_words_around=info.get("num_words_around_snippet", 5),
                                         start=info.get("snippet_start_tag", "<b>"), end=info.get("snippet_end_tag", "</b>"), connector=info.get("snippet_connector", " ... "))
        if text is not None:
            out.append((text, best, group[0][0]))
        i = j
    out.sort(key=lambda h: -h[1])  # :186 (ties: Python's sort is stable, the reference's is not)
    skip, top = part.get("skip"), part.get("top")
    if skip:
        out = out[skip:]
    if top is not None:
        out = out[:top]
    return out
