"""CPU oracle (TEST INFRASTRUCTURE, never imported by the product): `select` -- documents rebuilt from the indices
(src/search/read_document.rs:8-59, src/search.rs:242-279) -- and the token-id based why_found that goes with it
(src/search/why_found.rs:11-49, highlight_document src/highlight_field.rs:187-271), in plain Python over the oracle's own
decoder of the index files (oracle/index_files.py).  The product's implementation is csrc/host/read_document.hpp."""
import index_files as oif
from highlight import group_hit_positions_for_snippet

PARENT_TO_VALUE_ID = ".parent_to_value_id"


class Reader:
    def __init__(self, directory):
        self.ix = oif.IndexDirectory(directory)
        self.stores = {}
        for _field, meta in self.ix.indices():
            if meta["index_category"] == "KeyValue":
                self.stores[meta["path"]] = meta
        self._terms = {}
        self._files = {}

    def _file(self, name):
        if name not in self._files:
            self._files[name] = self.ix.read(name)
        return self._files[name]

    def has_index(self, path):
        return path in self.stores

    def get_values(self, path, value_id):  # IndexIdToParent::get_values
        meta = self.stores[path]
        if meta["is_empty"]:
            return None
        if meta["index_cardinality"] == "MultiValue":
            return oif.indirect_get_values(self._file(path + ".indirect"), self._file(path + ".data"), value_id)
        v = oif.packed_get_value(self._file(path), oif.packed_width(meta["metadata"]["max_value_id"]), value_id)
        return None if v is None else [v]

    def get_value(self, path, value_id):
        vals = self.get_values(path, value_id)
        return vals[0] if vals else None

    def text_for_id(self, field, term_id):  # get_text_for_id (search_field.rs:520-526): ord_to_term on the field's FST, its result ignored
        if field not in self._terms:
            self._terms[field] = oif.Fst(self._file(field + ".fst"))
        return self._terms[field].ord_to_term(term_id)[1].decode("utf-8", "replace")

    def join_and_get_text_for_ids(self, value_id, prop):  # src/search.rs:242-269
        field = prop + ".textindex"
        text_id = self.get_value(field + PARENT_TO_VALUE_ID, value_id)
        if text_id is None:
            return None
        if text_id >= self.ix.meta["columns"][prop]["textindex_metadata"]["num_text_ids"]:
            tokens = self.get_values(field + ".text_id_to_token_ids", text_id)
            assert tokens is not None, "MissingTextId"
            return "".join(self.text_for_id(field, t) for t in tokens)
        return self.text_for_id(field, text_id)

    def read_tree_from_fields(self, fields):  # src/search.rs:272-279, util.rs:175-229
        paths = []
        for f in fields:
            if self.has_index(f + ".textindex" + PARENT_TO_VALUE_ID):
                parts = f.split(".")
                paths.append([".".join(parts[:i + 1]) for i in range(len(parts))])
        return to_node_tree(paths)

    def read_tree(self, value_id, tree):  # src/search/read_document.rs:13-59
        out = {}
        for prop, sub in tree.items():
            current = prop + PARENT_TO_VALUE_ID
            is_array = prop.endswith("[]")
            name = prop.split(".")[-1]
            name = name[:-2] if name.endswith("[]") else name
            if sub is None:
                if is_array:
                    sub_ids = self.get_values(current, value_id)
                    if sub_ids is not None:
                        out[name] = [t for t in (self.join_and_get_text_for_ids(s, prop) for s in sub_ids) if t is not None]
                else:
                    text = self.join_and_get_text_for_ids(value_id, prop)
                    if text is not None:
                        out[name] = text
            elif not self.has_index(current):
                out[name] = self.read_tree(value_id, sub)
            else:
                sub_ids = self.get_values(current, value_id)
                if sub_ids is not None:
                    if is_array:
                        out[name] = [self.read_tree(s, sub) for s in sub_ids]
                    elif sub_ids:
                        out[name] = self.read_tree(sub_ids[0], sub)
        return out

    def read_data(self, doc_id, fields):  # read_document.rs:8-11
        return self.read_tree(doc_id, self.read_tree_from_fields(fields))

    # ---- why_found with select
    def highlight_document(self, path, text_id, token_ids, num_words_around=5, start="<b>", end="</b>", connector=" ... "):  # highlight_field.rs:187-271
        doc = self.get_values(path + ".text_id_to_token_ids", text_id)
        if doc is None:
            if text_id in token_ids:
                return start + self.text_for_id(path, text_id) + end
            return None
        hits = sorted(pos for pos, t in enumerate(doc) if t in token_ids)
        if not hits:
            return None
        around = num_words_around * 2
        parts = []
        for group in group_hit_positions_for_snippet(hits, num_words_around):
            lo, hi = max(group[0] - around, 0), min(group[-1] + around + 1, len(doc))
            parts.append("".join(start + self.text_for_id(path, doc[i]) + end if doc[i] in token_ids else self.text_for_id(path, doc[i]) for i in range(lo, hi)))
        snippet = connector.join(parts)
        if hits[0] > around:
            snippet = connector + snippet
        if hits[-1] < len(doc) - around:
            snippet = snippet + connector
        return snippet

    def get_why_found(self, anchor, term_ids_in_field):  # why_found.rs:11-49, one anchor; {"<field>.textindex": ids}
        out = {}
        for path, ids in term_ids_in_field.items():
            if not ids:
                continue
            field = path[:-len(".textindex")]
            steps, cur = [], []
            for part in field.split("."):  # util.rs:147-162 get_steps_to_anchor
                cur.append(part)
                if part.endswith("[]"):
                    steps.append(".".join(cur))
            steps.append(field + ".textindex")
            value_ids = [anchor]
            for step in steps:  # facet.rs:75-93 join_anchor_to_leaf
                nxt = []
                for v in value_ids:
                    nxt.extend(self.get_values(step + PARENT_TO_VALUE_ID, v) or [])
                value_ids = nxt
            for text_id in value_ids:
                h = self.highlight_document(steps[-1], text_id, set(ids))
                if h is not None:
                    out.setdefault(field, []).append(h)
        return out


def to_node_tree(paths):  # util.rs:201-229; a leaf is None
    tree = {}
    keys = sorted({p[0] for p in paths})
    for key in keys:
        rest = [p[1:] for p in paths if p[0] == key]
        is_leaf = any(len(r) == 0 for r in rest)
        rest = [r for r in rest if r]
        tree[key] = None if (not rest or is_leaf) else to_node_tree(rest)
    return tree
